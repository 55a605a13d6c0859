"""One profiled device-resident training iteration (warm-up outside the profiled range): ncu --profile-from-start off target.
PROF_BATCH (default 8), PROF_PREC (bf16 = the tensor-core step, fp32_simt = the FFMA twin)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                               # noqa: E402
import monocon_pytorch_b200 as M                           # noqa: E402
from monocon_pytorch_b200 import engine as E               # noqa: E402
from monocon_pytorch_b200 import train_ops as T            # noqa: E402
from oracle import train_fixtures as TF                    # noqa: E402  (synthetic labels only)

B, H, W = int(os.environ.get('PROF_BATCH', '8')), 384, 1280
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
eng = E.Engine(dev, B, H, W, os.environ.get('PROF_PREC', 'bf16'))
eng.load_state_dict(model.state_dict(), training=2)
opt = T.ResidentClipAdamW(eng)
label = TF.make_labels(B, (H, W), seed=21, max_objs_per_image=8)
img = (torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(1)) * 0.5).to(dev)
data = {'img': img, 'img_metas': {'pad_shape': [(H, W)] * B}, 'label': {k: torch.from_numpy(v).to(dev) for k, v in label.items()}}
gen = T.TargetGenerator()
pred = eng.alloc_pred(B)


def iteration():
    eng.forward_train(img, out=pred)
    tgt = gen(data, (B, 64, H // 4, W // 4))
    loss, grad = T.get_losses(dict(zip(E.PRED_NAMES, pred)), tgt, with_grad=True, check_empty=False)
    eng.backward_train(pred, [grad[k] for k in E.PRED_NAMES])
    opt.step()


iteration()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
iteration()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
