"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the modulated deformable convolution (DCNv2) that BASELINE.json's north_star
names for the IDAUp nodes.  The reference repository itself contains no deformable convolution, so
the algorithm restated here is the published one of its usual dependency: torchvision.ops.deform_conv2d, pinned version
torchvision 0.26.0 (torchvision/ops/deform_conv.py:14-96; CPU kernel torchvision/csrc/ops/cpu/deform_conv2d_kernel.cpp:
bilinear_interpolate and deformable_im2col).  Pinned: tests/golden/gen_dcn_golden.py runs torchvision's own operator in the build
container on seeded inputs and stores its outputs in tests/golden/dcn.npz; tests/test_dcn_oracle.py holds this file to them.

Only tests/ may import this module; the product path (monocon_pytorch_b200/csrc/dcn.cu) never does.

Semantics (3x3, stride 1, padding 1, dilation 1, one offset group, as DLA's DCN nodes use it):
    y[n, co, h, w] = bias[co] + sum_{k = (i, j)} sum_ci w[co, ci, i, j] * mask[n, k, h, w] * x~[n, ci, h - 1 + i + dy, w - 1 + j + dx]
    dy = offset[n, 2k, h, w], dx = offset[n, 2k + 1, h, w]
    x~ = bilinear interpolation with zeros outside the image; a sample at py <= -1, py >= H, px <= -1 or px >= W is 0.
"""
import numpy as np
import torch


def bilinear_sample(x: torch.Tensor, py: torch.Tensor, px: torch.Tensor) -> torch.Tensor:
    """x (B, C, H, W); py, px (B, H, W) sample positions -> (B, C, H, W) (deform_conv2d_kernel.cpp: bilinear_interpolate)."""
    B, C, H, W = x.shape
    inside = (py > -1) & (py < H) & (px > -1) & (px < W)
    h_low, w_low = torch.floor(py), torch.floor(px)
    lh, lw = py - h_low, px - w_low
    hh, hw = 1 - lh, 1 - lw
    h_low, w_low = h_low.long(), w_low.long()
    h_high, w_high = h_low + 1, w_low + 1
    flat = x.reshape(B, C, H * W)

    def corner(hi, wi, ok):
        ok = ok & inside
        idx = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).reshape(B, 1, -1).expand(B, C, -1)
        v = torch.gather(flat, 2, idx).reshape(B, C, *py.shape[1:])
        return v * ok.unsqueeze(1).to(x.dtype)
    v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))
    v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
    v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
    v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
    return (hh * hw).unsqueeze(1) * v1 + (hh * lw).unsqueeze(1) * v2 + (lh * hw).unsqueeze(1) * v3 + (lh * lw).unsqueeze(1) * v4


def deform_conv2d(x: torch.Tensor, offset: torch.Tensor, mask: torch.Tensor, weight: torch.Tensor, bias=None, col_round=None) -> torch.Tensor:
    """3x3 / stride 1 / padding 1 / dilation 1 modulated deformable convolution; tensors as torchvision.ops.deform_conv2d takes them.
    ``col_round``: optional rounding of the sampled columns (the bf16-emulating checker: the engine's throughput mode stores them as bf16)."""
    B, C, H, W = x.shape
    Cout = weight.shape[0]
    assert weight.shape[1:] == (C, 3, 3) and offset.shape == (B, 18, H, W) and mask.shape == (B, 9, H, W)
    ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    out = torch.zeros(B, Cout, H, W, dtype=x.dtype)
    for k in range(9):
        i, j = k // 3, k % 3
        py = ys - 1 + i + offset[:, 2 * k]
        px = xs - 1 + j + offset[:, 2 * k + 1]
        col = bilinear_sample(x, py, px) * mask[:, k].unsqueeze(1)            # deformable_im2col: one column block per tap
        if col_round is not None:
            col = col_round(col)
        out += torch.einsum('oc,bchw->bohw', weight[:, :, i, j], col)
    if bias is not None:
        out += bias.view(1, -1, 1, 1)
    return out


def deform_columns(x: torch.Tensor, offset: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """The column tensor deform_conv2d contracts with the weights, (B, 9 C, H, W) with channel index tap * C + c (deformable_im2col);
    what the engine's dcn_columns_kernel materialises (stage-wise tests)."""
    B, C, H, W = x.shape
    ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
    xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
    cols = []
    for k in range(9):
        i, j = k // 3, k % 3
        cols.append(bilinear_sample(x, ys - 1 + i + offset[:, 2 * k], xs - 1 + j + offset[:, 2 * k + 1]) * mask[:, k].unsqueeze(1))
    return torch.cat(cols, 1)


def dcn_pack(x: torch.Tensor, weight: torch.Tensor, offset_weight: torch.Tensor, offset_bias: torch.Tensor, round_fn=None) -> torch.Tensor:
    """The DCNv2 "pack" block as DLA's deformable necks use it (mmcv ModulatedDeformConv2dPack.forward / CenterNet's DCN.forward):
    a plain 3x3 convolution with bias yields 27 channels from the block's own input; channels 0..17 are the offsets ((dy, dx) per
    tap), channels 18..26 the mask logits (sigmoid); no bias on the deformable convolution itself.
    ``round_fn``: storage rounding of the offset field and the columns (bf16-emulating checker only)."""
    q = round_fn if round_fn is not None else (lambda t: t)
    om = q(torch.nn.functional.conv2d(x, q(offset_weight), offset_bias, stride=1, padding=1))
    return deform_conv2d(x, om[:, :18], torch.sigmoid(om[:, 18:27]), q(weight), None, col_round=round_fn)


def make_case(B, C, H, W, Cout, seed):
    """Seeded inputs: offsets of a few pixels (so that samples leave the image at the borders), masks in (0, 1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    offset = torch.randn(B, 18, H, W, generator=g) * 2.0
    mask = torch.sigmoid(torch.randn(B, 9, H, W, generator=g))
    w = torch.randn(Cout, C, 3, 3, generator=g) * (2.0 / (C * 9)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    return x, offset, mask, w, b


GOLDEN_CASES = [(2, 16, 12, 20, 32, 1), (1, 64, 9, 13, 64, 2), (2, 8, 6, 8, 16, 3)]
# channel counts of the DCN neck (9 C a multiple of 64: the tensor-core 1x1 layer over the columns), checked against this oracle on the GPU
GPU_CASES = [(1, 64, 9, 13, 64, 2), (2, 64, 24, 40, 64, 5), (2, 128, 16, 24, 64, 6), (1, 256, 12, 20, 128, 7), (2, 512, 8, 12, 256, 8)]
