"""TEST INFRASTRUCTURE ONLY (tests/host_shim/asan.sh): walks the inference entry points of the C ABI on the host engine (fp32 plan, inert
stand-ins for the eval-mode AttnBN mixture and the decode kernel) so that their HOST logic -- staging buffers, the two-slot host
pipeline, stage tables, debug read-back -- runs under AddressSanitizer / UBSan.  Values are not checked here (the GPU parity tests do).
    python run_infer_entry_points.py <libmonocon_host_engine.so>"""
import sys, ctypes as C, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from monocon_pytorch_b200 import engine as E
from oracle import fixtures as FX
L=C.CDLL(sys.argv[1]); E.declare_signatures(L)
vp=C.c_void_p
sd=FX.make_state_dict(0)
B,H,W=2,64,128
def ok(rc,h=None):
    assert rc==0, L.mc_last_error(h).decode()
h=vp(); ok(L.mc_create(C.byref(h),0,B,H,W,E.MC_PREC_FP32))
keep=[]
for k,v in sd.items():
    if not torch.is_floating_point(v): continue
    a=np.ascontiguousarray(v.numpy().astype(np.float32)); keep.append(a)
    ok(L.mc_set_param(h,k.encode(),a.ctypes.data,(C.c_int64*max(1,a.ndim))(*a.shape),a.ndim),h)
ok(L.mc_finalize_params(h,0),h)
ok(L.mc_set_option(h,b'use_graph',0),h)
img=np.ascontiguousarray(FX.make_images(B,H,W,seed=1).numpy().astype(np.float32))
pred=[np.zeros((B,c,H//4,W//4),np.float32) for c in (3,9,2,2,2,18,3,2,12,12)]
parr=(vp*10)(*[p.ctypes.data for p in pred])
ok(L.mc_forward(h,img.ctypes.data,B,parr,None),h)
ok(L.mc_forward(h,img.ctypes.data,1,parr,None),h)
P2=np.tile(np.array([[700,0,600,40],[0,700,180,2],[0,0,1,0.003]],np.float32),(B,1,1)); invP=np.tile(np.eye(4,dtype=np.float32),(B,1,1))
K=30
b2=np.zeros((B,K,5),np.float32); b3=np.zeros((B,K,7),np.float32); lb=np.zeros((B,K),np.int64); ix=np.zeros((B,K),np.int64); vl=np.zeros((B,K),np.uint8)
ok(L.mc_decode(h,parr,B,P2.ctypes.data,invP.ctypes.data,H,W,K,0.4,b2.ctypes.data,b3.ctypes.data,lb.ctypes.data,ix.ctypes.data,vl.ctypes.data,None),h)
ok(L.mc_infer_device(h,img.ctypes.data,B,P2.ctypes.data,invP.ctypes.data,K,0.4,b2.ctypes.data,b3.ctypes.data,lb.ctypes.data,ix.ctypes.data,vl.ctypes.data,None),h)
ok(L.mc_infer_host(h,img.ctypes.data,B,P2.ctypes.data,invP.ctypes.data,K,0.4,b2.ctypes.data,b3.ctypes.data,lb.ctypes.data,ix.ctypes.data,vl.ctypes.data,None),h)
for rep in range(3):
    for slot in (0,1):
        ok(L.mc_infer_host_submit(h,slot,img.ctypes.data,B,P2.ctypes.data,invP.ctypes.data,K,0.4,b2.ctypes.data,b3.ctypes.data,lb.ctypes.data,ix.ctypes.data,vl.ctypes.data),h)
    for slot in (0,1): ok(L.mc_infer_host_wait(h,slot),h)
ptrs=(vp*10)(); ok(L.mc_get_pred_ptrs(h,ptrs),h); ok(L.mc_copy_pred(h,B,parr,None),h)
n=L.mc_num_stages(h); ms=(C.c_float*n)()
for s in range(n):
    name=C.create_string_buffer(96); f=C.c_double(); by=C.c_double(); impl=C.c_int()
    ok(L.mc_stage_info(h,s,name,96,C.byref(f),C.byref(by),C.byref(impl)),h)
ok(L.mc_profile_stages(h,img.ctypes.data,B,P2.ctypes.data,invP.ctypes.data,1,ms,None),h)
c_,h_,w_=C.c_int(),C.c_int(),C.c_int()
ok(L.mc_debug_tensor_shape(h,b'neck.feat',C.byref(c_),C.byref(h_),C.byref(w_)),h)
out=np.zeros((B,c_.value,h_.value,w_.value),np.float32); ok(L.mc_debug_tensor(h,b'neck.feat',B,out.ctypes.data,None),h)
# uint8 input path
u8=np.random.RandomState(0).randint(0,255,(B,H-3,W-5,3)).astype(np.uint8); hw=np.array([[H-3,W-5],[H-10,W-20]],np.int32)
print('workspace',L.mc_workspace_bytes(h), 'stages',n, 'launches',L.mc_num_kernel_launches(h))
L.mc_destroy(h)
print('inference host logic under asan/ubsan: clean')
