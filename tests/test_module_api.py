"""CPU: the drop-in nn.Module surface (state_dict layout, constructor, error behaviour) and the C ABI."""
import ctypes
import os
import re

import pytest
import torch

import monocon_pytorch_b200 as M
from monocon_pytorch_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_matches_reference():
    m = M.MonoConDetector(num_dla_layers=34, pretrained_backbone=False)
    ref = [l.strip().split(' ', 1) for l in open(os.path.join(ROOT, 'tests', 'golden', 'keys.txt'))]
    mine = [(k, f'{tuple(v.shape)} {v.dtype}') for k, v in m.state_dict().items()]
    assert [k for k, _ in mine] == [k for k, _ in ref]
    assert [s for _, s in mine] == [s for _, s in ref]
    assert len(list(m.parameters())) == 242
    assert sum(p.numel() for p in m.parameters()) == 19620261


def test_load_state_dict_roundtrip(fixture_sd):
    m = M.MonoConDetector(pretrained_backbone=False)
    res = m.load_state_dict(fixture_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, fixture_sd[k]), k


def test_reference_init_statistics():
    torch.manual_seed(0)
    m = M.MonoConDetector(pretrained_backbone=False)
    sd = m.state_dict()
    assert abs(float(sd['head.heatmap_head.3.bias'][0]) + 2.1972246) < 1e-5            # monocon_heads.py:134-137
    assert float(sd['head.wh_head.0.weight'].std()) < 2e-3                              # N(0, 0.001), :139-146
    w = sd['backbone.level2.tree1.conv2.weight']                                        # N(0, sqrt(2/(9*64))), dla.py:264-271
    assert abs(float(w.std()) - (2.0 / (9 * 64)) ** 0.5) < 5e-3
    up = sd['neck.ida_0.up_1.weight'][0, 0]                                             # bilinear, dla_neck.py:83-92
    assert torch.allclose(up[0], torch.tensor([0.0625, 0.1875, 0.1875, 0.0625]))


def test_error_behaviour_without_gpu_or_in_train_mode():
    m = M.MonoConDetector(pretrained_backbone=False)
    with pytest.raises(E.EngineError):
        m.train()({'img': torch.zeros(2, 3, 64, 64)})                                   # train mode: CUDA only, too
    with pytest.raises(Exception, match='training mode'):                               # monocon_detector.py:72-73
        m.train().batch_eval({'img': torch.zeros(1, 3, 64, 64)})
    with pytest.raises(E.EngineError):
        m.eval()({'img': torch.zeros(1, 3, 64, 64)})                                    # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        M.MonoConDetector(num_dla_layers=60, pretrained_backbone=False)


def test_cabi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/monocon_b200.h declares."""
    lib = E.load_library()
    hdr = open(os.path.join(ROOT, 'include', 'monocon_b200.h')).read()
    declared = sorted(set(re.findall(r'MC_API\s+[\w\s\*]+?\b(mc_\w+)\s*\(', hdr)))
    assert declared == sorted(E.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_cabi_fails_loudly_without_gpu():
    lib = E.load_library()
    h = ctypes.c_void_p()
    assert lib.mc_create(ctypes.byref(h), 0, 1, 64, 64, 0) != 0
    assert len(lib.mc_last_error(None)) > 0
