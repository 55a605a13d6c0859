"""GPU tests of the entry points added in round 2, all through the C ABI:
* mc_infer_host_u8_submit / _wait   -- the pipelined host API fed with uint8 frames (bench.py's end-to-end path)
* mc_refresh_params                  -- new weights into the same device buffers (what the module does after an optimiser step)
* mc_calibrate_scales / mc_scale_status -- the range management of the fp32-accurate tensor-core mode
"""
import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import monocon_oracle as O

pytestmark = pytest.mark.gpu

from monocon_pytorch_b200 import engine as E          # noqa: E402

DEV = torch.device('cuda', 0)


def _rel_to_max(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_pipelined_uint8_host_api_matches_device_path(fixture_sd, precision):
    """Two slots, three batches of pinned uint8 HWC frames: every batch's host results equal what mc_infer_device_u8 returns
    for the same frames (bit for bit: same kernels, same inputs)."""
    H, W, B = 128, 256, 2
    eng = E.Engine(DEV, B, H, W, precision)
    eng.load_state_dict(fixture_sd)
    rng = np.random.RandomState(4)
    frames = [torch.from_numpy(rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)).pin_memory() for _ in range(3)]
    hw_h = torch.tensor([[H, W], [H - 7, W - 12]], dtype=torch.int32).pin_memory()      # the second frame is smaller than the canvas
    P2_np = FX.kitti_p2(B, 5)
    P2_h, invP_h = torch.from_numpy(P2_np), E.inverse_viewpad(P2_np)
    if eng.tensor_core_fp32:
        eng.calibrate_scales(torch.from_numpy(O.preprocess_u8([f.numpy() for f in frames[0]])).to(DEV).contiguous()[:, :, :H, :W].contiguous())
    want = []
    for f in frames:
        out = eng.infer_device_u8(f.to(DEV), hw_h.to(DEV), P2_h.to(DEV), invP_h.to(DEV), topk=30, thres=0.0)
        want.append({k: v.cpu() for k, v in out.items()})
    outs = [E.Engine.alloc_host_out(B, 30), E.Engine.alloc_host_out(B, 30)]
    got = []
    eng.infer_host_u8_submit(0, frames[0], hw_h, P2_h, invP_h, outs[0], topk=30, thres=0.0)
    for i in range(3):
        if i + 1 < 3:
            eng.infer_host_u8_submit((i + 1) & 1, frames[i + 1], hw_h, P2_h, invP_h, outs[(i + 1) & 1], topk=30, thres=0.0)
        eng.infer_host_wait(i & 1)
        got.append({k: v.clone() for k, v in outs[i & 1].items()})
    for w, g in zip(want, got):
        for k in w:
            assert torch.equal(w[k], g[k]), k
    with pytest.raises(E.EngineError):
        eng.infer_host_u8_submit(0, frames[0].to(DEV), hw_h, P2_h, invP_h, outs[0])      # device tensor: this is the HOST api
    eng.close()


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'fp32_simt'])
def test_refresh_state_dict_equals_a_fresh_engine(fixture_sd, precision):
    """mc_refresh_params: the same handle, the same buffers (and the same captured CUDA graph), new weights -- bit-identical to an
    engine created from those weights."""
    H, W, B = 128, 256, 2
    img = FX.make_images(B, H, W, seed=7).to(DEV)
    P2_np = FX.kitti_p2(B, 5)
    P2, invP = torch.from_numpy(P2_np).to(DEV), E.inverse_viewpad(P2_np).to(DEV)
    sd_b = {k: (v * 1.25 if (torch.is_floating_point(v) and k.endswith('conv1.weight')) else v.clone()) for k, v in fixture_sd.items()}
    eng = E.Engine(DEV, B, H, W, precision)
    eng.load_state_dict(fixture_sd)
    eng.set_option('use_graph', 1)
    first = {k: v.clone() for k, v in eng.infer_device(img, P2, invP, topk=30, thres=0.0).items()}
    ws = eng.workspace_bytes
    eng.refresh_state_dict(sd_b)
    assert eng.workspace_bytes == ws                                     # nothing was allocated
    again = {k: v.clone() for k, v in eng.infer_device(img, P2, invP, topk=30, thres=0.0).items()}   # replays the captured graph
    maps = [t.clone() for t in eng.pred_views(B)]
    fresh = E.Engine(DEV, B, H, W, precision)
    fresh.load_state_dict(sd_b)
    ref = fresh.infer_device(img, P2, invP, topk=30, thres=0.0)
    ref_maps = fresh.pred_views(B)
    assert not torch.equal(first['box3d'], again['box3d'])               # the new weights took effect
    for k in ref:
        assert torch.equal(ref[k], again[k]), k
    for a, b in zip(maps, ref_maps):
        assert torch.equal(a, b)
    eng.refresh_state_dict(fixture_sd)                                   # and back
    back = eng.infer_device(img, P2, invP, topk=30, thres=0.0)
    for k in first:
        assert torch.equal(first[k], back[k]), k
    eng.close(); fresh.close()


def test_scale_calibration_and_saturation_report(fixture_sd):
    """fp32-accurate tensor-core mode: frames 30000x larger than anything the uncalibrated scales (2^0) can hold saturate the fp16
    planes -- reported by mc_scale_status, never inf / nan; after mc_calibrate_scales on that batch nothing saturates and the
    maps are back inside the 1e-3 gate against the oracle on the same frames."""
    H, W, B = 128, 256, 2
    img = FX.make_images(B, H, W, seed=19) * 3.0e4
    eng = E.Engine(DEV, B, H, W, 'fp32')
    eng.load_state_dict(fixture_sd)
    out = eng.forward(img.to(DEV))
    frac, nsat = eng.scale_status()
    assert nsat > 0 and frac >= 0.99
    assert all(bool(torch.isfinite(t).all()) for t in out)
    eng.calibrate_scales(img.to(DEV))
    out = eng.forward(img.to(DEV))
    frac, nsat = eng.scale_status()
    assert nsat == 0 and 0.01 < frac < 0.2, (frac, nsat)                # maxima sit 2^-4 ... 2^-5 below the fp16 limit
    ref, inter = O.forward(fixture_sd, img, return_intermediates=True)
    assert _rel_to_max(eng.debug_tensor('neck.feat', B).cpu().numpy(), inter['feat'].numpy()) < 1e-3
    for k, t in zip(E.PRED_NAMES, out):
        got, want = t.cpu().numpy(), ref[k].numpy()
        if k in ('center_heatmap_pred', 'kpt_heatmap_pred', 'depth_pred'):
            continue        # logits 3e4 times larger than usual turn the sigmoid into a step function: not a numerics check
        assert _rel_to_max(got, want) < 1e-3, k
    # tiny frames: uncalibrated they fall into the fp16 subnormals (lo pieces vanish), calibrated they are exact again
    tiny = FX.make_images(B, H, W, seed=19) * 1.0e-3
    eng.calibrate_scales(tiny.to(DEV))
    out = eng.forward(tiny.to(DEV))
    ref = O.forward(fixture_sd, tiny)
    assert eng.scale_status()[1] == 0
    for k, t in zip(E.PRED_NAMES, out):
        assert _rel_to_max(t.cpu().numpy(), ref[k].numpy()) < 1e-3, k
    eng.close()
