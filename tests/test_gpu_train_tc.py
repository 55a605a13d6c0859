"""GPU tests of the bf16 tensor-core training kernels (run with ``-m gpu`` on a B200), through the C ABI.

Kernel level: the weight gradient (csrc/wgrad_tc.cu) against float64 autograd of torch.nn.functional.conv2d on identical
bf16-rounded operands (the kernel multiplies bf16 x bf16 exactly and accumulates in fp32, so the only differences are the
accumulation order and the tensor core's accumulate rounding).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from monocon_pytorch_b200 import engine as E          # noqa: E402

DEV = torch.device('cuda', 0)

WGRAD_CASES = [
    # B, Cin, H, W, Cout, k, split
    (2, 64, 16, 16, 64, 3, 1),          # one 64-channel chunk, one Cout tile with 64 real rows (SWIZZLE_128B both operands)
    (2, 64, 24, 40, 128, 3, 1),         # two dy boxes per step (M = 128), H = 24 -> 12-row tiles
    (3, 128, 12, 40, 128, 3, 2),        # two sources (IDAUp node, dla_neck.py:104), odd batch
    (2, 256, 12, 24, 512, 3, 1),        # level5 conv1 shape class: 4 Cout tiles x 4 chunks x 2 tap groups
    (1, 64, 24, 80, 576, 3, 1),         # nine head stems as one convolution: last Cout tile has 64 rows
    (2, 32, 32, 64, 64, 3, 1),          # 32 input channels: SWIZZLE_64B halo tile, all nine taps in one group
    (2, 16, 32, 64, 16, 3, 1),          # level0: 16 -> 16, SWIZZLE_32B on both sides
    (2, 16, 32, 64, 32, 3, 1),          # level1 (as the stride-1 problem on the zero-inserted gradient)
    (2, 32, 18, 40, 64, 3, 1),          # H = 18: 16-row tiles with a zero-filled remainder
    (2, 512, 12, 24, 128, 1, 4),        # Root 1x1 over four children (dla.py:126)
    (2, 128, 24, 40, 64, 1, 2),         # level2 Root
    (2, 32, 24, 40, 64, 1, 1),          # project 1x1 (dla.py:181-185)
    (5, 128, 48, 160, 128, 3, 1),       # more pixel tiles than SMs x stages: the ring wraps, split-K over many CTAs
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_wgrad_tc_kernel_parity(case):
    B, Cin, H, W, Cout, k, split = case
    g = torch.Generator().manual_seed(hash(case) & 0xffff)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    dy = torch.randn(B, Cout, H, W, generator=g).bfloat16().float()
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, stride=1, padding=(k - 1) // 2).backward(dy.double())
    ref = w.grad
    got = E.conv2d_wgrad_tc(x.to(DEV), dy.to(DEV), k, split=split).cpu().double()
    err = float((got - ref).abs().max() / ref.abs().max())
    # per-tap diagnostics make a wrong descriptor interpretation readable in the log
    if err >= 1e-4:
        per_tap = [(float((got[:, :, i // k, i % k] - ref[:, :, i // k, i % k]).abs().max() / ref.abs().max())) for i in range(k * k)]
        per_co = (got - ref).abs().amax(dim=(1, 2, 3)) / ref.abs().max()
        per_ci = (got - ref).abs().amax(dim=(0, 2, 3)) / ref.abs().max()
        print(f'{case}: per-tap {per_tap}\n bad co {torch.nonzero(per_co > 1e-4).flatten().tolist()[:40]}\n bad ci {torch.nonzero(per_ci > 1e-4).flatten().tolist()[:40]}')
    assert err < 1e-4, f'{case}: rel-to-max error {err:.3e}'
