#!/bin/sh
# TEST INFRASTRUCTURE ONLY: the host-shim build of csrc/train_backward.cu under AddressSanitizer + UBSan, then every kernel case of
# tests/backward_cases.py and the full-graph backward on numpy buffers (malloc'ed, so ASan's red zones surround them).  An
# out-of-bounds index that is harmless on the host would be an illegal address on the device; this is the closest check this
# GPU-less container offers.  Last run: clean (round 1).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="${TMPDIR:-/tmp}/mc_asan"
mkdir -p "$out"
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -DMC_HOST_SHIM -I"$here" \
    -I"$root/monocon_pytorch_b200/csrc" -x c++ "$root/monocon_pytorch_b200/csrc/train_backward.cu" -o "$out/libtrain_backward_host.so"
cat > "$out/run.py" <<PY
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, "$root"); sys.path.insert(0, "$root/tests")
import backward_cases as BC
import test_backward_graph_host as T
from oracle import fixtures as FX, train_fixtures as TF
L = C.CDLL("$out/libtrain_backward_host.so"); L.mc_bw_last_error.restype = C.c_char_p; L.mc_bw_heads_scratch_bytes.restype = C.c_longlong
bk = BC.HostBackend(L)
for c in BC.CONV_CASES: BC.conv_case(bk, *c)
for c in BC.BN_CASES: BC.bn_case(bk, *c)
BC.colsum_case(bk); BC.maxpool_case(bk); BC.upsample_case(bk)
for c in BC.HEAD_CASES: BC.heads_case(bk, *c)
sd = FX.make_state_dict(0)
B, hw = 2, (64, 128)
img, label = FX.make_images(B, *hw, seed=41), TF.make_labels(B, hw, seed=42)
with torch.no_grad():
    G = T.Graph({k: v.clone() for k, v in sd.items()}, B); ts = G.build(img.float())
args, hb, keep, losses = G.heads(ts, label, hw)
n = G.tensors[ts].H * G.tensors[ts].W
scratch = np.zeros(int(L.mc_bw_heads_scratch_bytes(B, n)) // 8 + 64, np.float64); args.scratch = (scratch.ctypes.data + 255) // 256 * 256
op = T.Op(); op.type, op.nsrc, op.heads = T.HEADS, 1, C.pointer(args); op.src[0] = ts; G.ops.append(op)
tensors, ops = (T.Tensor * len(G.tensors))(*G.tensors), (T.Op * len(G.ops))(*G.ops)
assert L.mc_bw_run_graph(tensors, len(G.tensors), ops, len(G.ops), B, None) == 0
print("asan/ubsan: clean")
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python "$out/run.py"

# ---- the engine's host logic (api.cu + engine.cu against the stand-in runtime) under the same sanitizers --------------------------
# (round 1: found a use-after-free in Net::add_conv -- a TensorInfo reference held across add_tensor's push_back -- fixed)
src="$root/monocon_pytorch_b200/csrc"
F="-O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -I${CUDA_HOME:-/usr/local/cuda}/include -I$src"
g++ $F -x c++ -c "$src/api.cu" -o "$out/api.o"
g++ $F -x c++ -c "$src/engine.cu" -o "$out/engine.o"
g++ $F -c "$here/host_engine_stubs.cpp" -o "$out/stubs.o"
g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -DMC_HOST_SHIM -I"$here" -I"$src" -x c++ -c "$src/train_backward.cu" -o "$out/tb.o"
g++ -shared -Wl,-Bsymbolic -fsanitize=address,undefined -o "$out/libmonocon_host_engine.so" "$out/api.o" "$out/engine.o" "$out/stubs.o" "$out/tb.o"
cat > "$out/run_engine.py" <<PY
import sys, ctypes as C
sys.path.insert(0, "$root"); sys.path.insert(0, "$root/tests")
import backward_cases as BC
import test_host_engine as T
from monocon_pytorch_b200 import engine as E
from oracle import fixtures as FX
L = C.CDLL("$out/libmonocon_host_engine.so")
L.mc_bw_last_error.restype = C.c_char_p; L.mc_bw_heads_scratch_bytes.restype = C.c_longlong
E.declare_signatures(L)
L.mc_debug_bw_graph.argtypes = [C.c_void_p, C.POINTER(C.POINTER(BC.Tensor)), C.POINTER(C.c_int), C.POINTER(C.POINTER(BC.Op)), C.POINTER(C.c_int)]
T.test_engine_driven_training_step_on_the_host(L, FX.make_state_dict(0))
print("host engine under asan/ubsan: clean")
PY
MC_ASAN=1 LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python "$out/run_engine.py"
# ---- and the inference entry points' host logic (staging, host pipeline slots, stage tables) -------------------------------------------
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python "$here/run_infer_entry_points.py" "$out/libmonocon_host_engine.so"
