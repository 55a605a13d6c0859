"""Host-side mirror of the reference's training-side interfaces over the C ABI (include/monocon_b200.h, csrc/train_ops.cu).

* ``TargetGenerator``   -- same constructor and ``__call__(input_dict, feat_shape)`` as utils/target_generator.py:17-138
* ``get_losses``        -- ``MonoConDenseHeads._get_losses(pred_dict, target_dict)`` (monocon_heads.py:203-310); optionally
                           also d(sum of the losses)/d(pred) for the ten maps
* ``ClipAdamW``         -- ``clip_grad_norm_(params, 35, 2)`` + ``torch.optim.AdamW.step`` of engine/monocon_engine.py:39-53,94-100
                           as one fused step over all parameter tensors (``param_groups`` / ``state_dict`` kept so that
                           solver/cyclic_scheduler.py can rewrite ``lr`` / ``betas`` as it does for torch's optimiser)

No CPU fallback: CUDA tensors only, the shared library must be present.
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, Iterable, List, Optional, Tuple

import torch

from .engine import EngineError, PRED_CHANNELS, PRED_NAMES, load_library

LOSS_NAMES = ('loss_center_heatmap', 'loss_wh', 'loss_offset', 'loss_dim', 'loss_center2kpt_offset', 'loss_kpt_heatmap',
              'loss_kpt_heatmap_offset', 'loss_alpha_cls', 'loss_alpha_reg', 'loss_depth')      # monocon_heads.py:299-309

_vp = ctypes.c_void_p


class _Labels(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ('gt_bboxes', 'gt_labels', 'gt_bboxes_3d', 'depths', 'gt_kpts_2d', 'gt_kpts_valid_mask', 'mask')]


_TARGET_FIELDS = (('center_heatmap', 'center_heatmap_target'), ('kpt_heatmap', 'kpt_heatmap_target'), ('wh', 'wh_target'),
                  ('offset', 'offset_target'), ('dim', 'dim_target'), ('alpha_cls', 'alpha_cls_target'),
                  ('alpha_offset', 'alpha_offset_target'), ('depth', 'depth_target'),
                  ('center2kpt_offset', 'center2kpt_offset_target'), ('kpt_heatmap_offset', 'kpt_heatmap_offset_target'),
                  ('indices', 'indices'), ('indices_kpt', 'indices_kpt'), ('mask_target', 'mask_target'),
                  ('mask_center2kpt_offset', 'mask_center2kpt_offset'), ('mask_kpt_heatmap_offset', 'mask_kpt_heatmap_offset'))


class _Targets(ctypes.Structure):
    _fields_ = [(n, _vp) for n, _ in _TARGET_FIELDS]


_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        ci, cd = ctypes.c_int, ctypes.c_double
        lib.mc_generate_targets.argtypes = [ci, ci, ci, ci, ci, ci, ci, ctypes.POINTER(_Labels), ctypes.POINTER(_Targets), _vp]
        lib.mc_losses_workspace_bytes.restype = ctypes.c_size_t
        lib.mc_losses.argtypes = [ci, ci, ci, ci, ci, ctypes.POINTER(_vp), ctypes.POINTER(_Targets), _vp, ctypes.POINTER(_vp), _vp, _vp]
        lib.mc_optimizer_create.argtypes = [ctypes.POINTER(_vp), ci, ci, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                            ctypes.POINTER(ctypes.c_int64)]
        lib.mc_optimizer_step.argtypes = [_vp, ctypes.POINTER(_vp), ci, cd, cd, cd, cd, cd, cd, _vp, _vp]
        lib.mc_optimizer_destroy.argtypes = [_vp]
        lib.mc_optimizer_destroy.restype = None
        lib.mc_train_last_error.restype = ctypes.c_char_p
        _bound = True
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        raise EngineError(f'{what}: ' + _lib().mc_train_last_error().decode())


def _dev(t: torch.Tensor) -> torch.device:
    if not t.is_cuda:
        raise EngineError('training-side kernels run on CUDA tensors only (no CPU fallback)')
    return t.device


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _targets_struct(t: Dict[str, torch.Tensor]) -> _Targets:
    s = _Targets()
    for field, key in _TARGET_FIELDS:
        setattr(s, field, t[key].data_ptr())
    return s


class TargetGenerator:
    """Drop-in for utils/target_generator.py:TargetGenerator (same arguments, same 15-tensor dict; ``mask_target`` is a
    bool tensor on the label's device instead of the CPU, which is where monocon_heads.py:214 moves it anyway)."""

    def __init__(self, num_classes: int = 3, max_objs: int = 30, num_kpt: int = 9, num_alpha_bins: int = 12):
        if (num_classes, num_kpt, num_alpha_bins) != (3, 9, 12):
            raise EngineError('the kernels are built for num_classes 3, num_kpt 9, num_alpha_bins 12 (monocon_detector.py:12-17)')
        self.num_classes, self.max_objs, self.num_kpt, self.num_alpha_bins = num_classes, max_objs, num_kpt, num_alpha_bins

    def __call__(self, input_dict: Dict[str, Any], feat_shape: Tuple[int, ...]) -> Dict[str, torch.Tensor]:
        label = input_dict['label']
        dev = _dev(label['gt_bboxes'])
        pad_h, pad_w = input_dict['img_metas']['pad_shape'][0]
        B, _, fh, fw = feat_shape
        M, K = self.max_objs, self.num_kpt
        f32 = dict(dtype=torch.float32, device=dev)
        lab = {'gt_bboxes': label['gt_bboxes'].to(torch.float32), 'gt_labels': label['gt_labels'].to(torch.uint8),
               'gt_bboxes_3d': label['gt_bboxes_3d'].to(torch.float32), 'depths': label['depths'].to(torch.float32),
               'gt_kpts_2d': label['gt_kpts_2d'].to(torch.float32), 'gt_kpts_valid_mask': label['gt_kpts_valid_mask'].to(torch.uint8),
               'mask': label['mask'].to(torch.uint8)}
        lab = {k: v.contiguous() for k, v in lab.items()}
        assert lab['mask'].shape == (B, M), (tuple(lab['mask'].shape), B, M)
        t = {'center_heatmap_target': torch.empty(B, 3, fh, fw, **f32), 'wh_target': torch.empty(B, M, 2, **f32),
             'offset_target': torch.empty(B, M, 2, **f32), 'dim_target': torch.empty(B, M, 3, **f32),
             'alpha_cls_target': torch.empty(B, M, 1, **f32), 'alpha_offset_target': torch.empty(B, M, 1, **f32),
             'depth_target': torch.empty(B, M, 1, **f32), 'center2kpt_offset_target': torch.empty(B, M, 2 * K, **f32),
             'kpt_heatmap_target': torch.empty(B, K, fh, fw, **f32), 'kpt_heatmap_offset_target': torch.empty(B, M, 2 * K, **f32),
             'indices': torch.empty(B, M, dtype=torch.int64, device=dev), 'indices_kpt': torch.empty(B, M * K, dtype=torch.int64, device=dev),
             'mask_target': torch.empty(B, M, dtype=torch.bool, device=dev),
             'mask_center2kpt_offset': torch.empty(B, M, 2 * K, **f32), 'mask_kpt_heatmap_offset': torch.empty(B, M, 2 * K, **f32)}
        ls = _Labels(**{k: v.data_ptr() for k, v in lab.items()})
        ts = _targets_struct(t)
        _check(_lib().mc_generate_targets(dev.index, B, M, fh, fw, int(pad_h), int(pad_w), ctypes.byref(ls), ctypes.byref(ts), _stream(dev)),
               'mc_generate_targets')
        return t


def get_losses(pred_dict: Dict[str, torch.Tensor], target_dict: Dict[str, torch.Tensor], max_objs: int = 30, with_grad: bool = False,
               check_empty: bool = True):
    """The reference's ten losses (dict of 0-dim tensors in its order).  ``with_grad=True`` additionally returns
    {pred key: d(sum of the ten losses)/d(pred)}.  ``check_empty`` reads one flag back (a host sync) and raises
    AssertionError for a batch without objects, like losses/l1_loss.py:15 does."""
    dev = _dev(pred_dict[PRED_NAMES[0]])
    B, _, fh, fw = pred_dict[PRED_NAMES[0]].shape
    preds = [pred_dict[k].to(torch.float32).contiguous() for k in PRED_NAMES]
    for p, c in zip(preds, PRED_CHANNELS):
        assert p.shape == (B, c, fh, fw), (tuple(p.shape), c)
    tgt = dict(target_dict)
    tgt['mask_target'] = tgt['mask_target'].to(dev).to(torch.bool).contiguous()
    grads = [torch.empty_like(p) for p in preds] if with_grad else None
    losses = torch.empty(len(LOSS_NAMES), dtype=torch.float32, device=dev)
    lib = _lib()
    ws = torch.empty(lib.mc_losses_workspace_bytes() // 8, dtype=torch.float64, device=dev)
    pp = (_vp * len(preds))(*[p.data_ptr() for p in preds])
    gp = (_vp * len(preds))(*[g.data_ptr() for g in grads]) if with_grad else None
    ts = _targets_struct(tgt)
    _check(lib.mc_losses(dev.index, B, max_objs, fh, fw, pp, ctypes.byref(ts), losses.data_ptr(), gp, ws.data_ptr(), _stream(dev)), 'mc_losses')
    if check_empty:
        assert float(ws[6].item()) == 0.0, 'no valid object in the batch (the reference asserts here too)'
    out = {k: losses[i] for i, k in enumerate(LOSS_NAMES)}
    if with_grad:
        return out, {k: g for k, g in zip(PRED_NAMES, grads)}
    return out


class AdamW(torch.optim.Optimizer):
    """``torch.nn.utils.clip_grad_norm_(params, max_norm, 2)`` followed by ``torch.optim.AdamW.step`` as one fused update
    (csrc/train_ops.cu: grad_sumsq_kernel + adamw_kernel over all tensors in two launches).

    A real ``torch.optim.Optimizer`` whose class is called ``AdamW``, because that is what the reference checks: its
    ``CyclicScheduler`` asserts ``optimizer.__class__.__name__ == 'AdamW'`` (solver/cyclic_scheduler.py:16) and ``_LRScheduler``
    insists on an ``Optimizer`` instance; both read and write ``param_groups[i]['lr' / 'betas' / 'initial_lr']``.  The state is
    kept in torch's own layout (``state[p] = {'step', 'exp_avg', 'exp_avg_sq'}``), so ``state_dict()`` / ``load_state_dict()``
    interchange with ``torch.optim.AdamW`` checkpoints (engine/base_engine.py:171-187).  One parameter group (the reference's
    solver has one, monocon_engine.py:39-44).  ``max_norm=None`` disables the clip.  Also exported as ``ClipAdamW``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2.25e-4, betas=(0.95, 0.99), eps: float = 1e-8,
                 weight_decay: float = 1e-5, max_norm: Optional[float] = 35.0):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        assert len(self.param_groups) == 1, 'one parameter group (as the reference solver builds it)'
        self.params: List[torch.Tensor] = list(self.param_groups[0]['params'])
        assert self.params, 'no parameters'
        dev = self.params[0].device          # construction is device-agnostic (schedulers, checkpoints); step() needs CUDA
        for p in self.params:
            assert p.device == dev and p.dtype == torch.float32 and p.is_contiguous()
        self.device = dev
        self.max_norm = max_norm
        self.total_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self._h = _vp()
        self._bound = None           # data pointers the C handle was created for

    # ---- state in torch's layout -----------------------------------------------------------------
    def _init_state(self) -> None:
        for p in self.params:
            st = self.state[p]
            if 'exp_avg' not in st:
                st['step'] = torch.zeros((), dtype=torch.float32)
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _bind(self) -> None:
        """(Re)create the C handle when the tensors behind the state changed (first step, load_state_dict)."""
        _dev(self.params[0])                 # raises on CPU tensors: the fused update has no CPU fallback
        self._init_state()
        ptrs = tuple(t.data_ptr() for p in self.params for t in (p, self.state[p]['exp_avg'], self.state[p]['exp_avg_sq']))
        if self._bound == ptrs and self._h:
            return
        self.close()
        n = len(self.params)
        arr = lambda ts: (_vp * n)(*[t.data_ptr() for t in ts])
        numel = (ctypes.c_int64 * n)(*[p.numel() for p in self.params])
        self._h = _vp()
        _check(_lib().mc_optimizer_create(ctypes.byref(self._h), self.device.index, n, arr(self.params),
                                          arr([self.state[p]['exp_avg'] for p in self.params]),
                                          arr([self.state[p]['exp_avg_sq'] for p in self.params]), numel), 'mc_optimizer_create')
        self._bound = ptrs

    @property
    def step_count(self) -> int:
        self._init_state()
        return int(self.state[self.params[0]]['step'].item())

    @property
    def exp_avg(self) -> List[torch.Tensor]:
        self._init_state()
        return [self.state[p]['exp_avg'] for p in self.params]

    @property
    def exp_avg_sq(self) -> List[torch.Tensor]:
        self._init_state()
        return [self.state[p]['exp_avg_sq'] for p in self.params]

    @torch.no_grad()
    def step(self, closure=None) -> torch.Tensor:
        """Returns the pre-clip total gradient norm (0-dim device tensor), as clip_grad_norm_ does."""
        if closure is not None:
            with torch.enable_grad():
                closure()
        self._bind()
        g = self.param_groups[0]
        step = self.step_count + 1
        grads = []
        for p in self.params:
            if p.grad is None:
                grads.append(None)
            else:
                assert p.grad.is_contiguous() and p.grad.dtype == torch.float32
                grads.append(p.grad.data_ptr())
        gp = (_vp * len(grads))(*grads)
        max_norm = float(self.max_norm) if self.max_norm is not None else 3.0e38
        _check(_lib().mc_optimizer_step(self._h, gp, step, float(g['lr']), float(g['betas'][0]), float(g['betas'][1]),
                                        float(g['eps']), float(g['weight_decay']), max_norm, self.total_norm.data_ptr(),
                                        _stream(self.device)), 'mc_optimizer_step')
        for p in self.params:
            self.state[p]['step'] += 1
        # the kernel wrote through raw pointers: tell torch the parameters changed (autograd's version counters are what
        # MonoConDetector watches to refresh its engine, and what torch's own saved-tensor checks rely on)
        touched = [p for p in self.params if p.grad is not None]
        bump = getattr(torch._C, '_increment_version', None)
        if bump is not None:
            for p in touched:
                bump(p)
        else:                                                    # pragma: no cover - older torch
            for p in touched:
                p.add_(0)
        return self.total_norm

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        super().load_state_dict(state_dict)
        self.params = list(self.param_groups[0]['params'])
        for p in self.params:                                    # the fused kernel needs dense fp32 moments on the parameter's device
            st = self.state.get(p)
            if st and 'exp_avg' in st:
                st['exp_avg'] = st['exp_avg'].to(p.device, torch.float32).contiguous()
                st['exp_avg_sq'] = st['exp_avg_sq'].to(p.device, torch.float32).contiguous()
                st['step'] = torch.as_tensor(float(st['step']), dtype=torch.float32)
        self._bound = None

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            _lib().mc_optimizer_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


ClipAdamW = AdamW


class ResidentClipAdamW:
    """EXPERIMENTAL.  The same fused clip + AdamW step (csrc/train_ops.cu) over the ENGINE's own trainable buffers and gradient
    buffers (``Engine.train_tensors()`` of an engine loaded with ``training=2``): both are element-wise, so the packed layouts
    need no unpacking and the whole iteration -- ``forward_train`` -> targets -> ``get_losses(with_grad=True)`` ->
    ``backward_train`` -> ``step`` -- stays on the device.  ``param_groups[0]`` carries ``lr`` / ``betas`` like torch's optimiser
    (solver/cyclic_scheduler.py:36-71 drives it unchanged); ``Engine.get_param`` / ``get_buffer`` give the module its
    state_dict back.  Reference: engine/monocon_engine.py:39-53 (solver), :94-100 (clip + step)."""

    def __init__(self, engine, lr: float = 2.25e-4, betas=(0.95, 0.99), eps: float = 1e-8, weight_decay: float = 1e-5, max_norm: float = 35.0):
        self.engine = engine
        self.device = engine.device
        self.tensors = engine.train_tensors()
        self.param_groups = [{'lr': lr, 'betas': tuple(betas), 'eps': eps, 'weight_decay': weight_decay}]
        self.max_norm = max_norm
        self.exp_avg = [torch.zeros(m, dtype=torch.float32, device=self.device) for _, _, _, m in self.tensors]
        self.exp_avg_sq = [torch.zeros(m, dtype=torch.float32, device=self.device) for _, _, _, m in self.tensors]
        self.step_count = 0
        self.total_norm = torch.zeros((), dtype=torch.float32, device=self.device)
        n = len(self.tensors)
        self._grads = (_vp * n)(*[g for _, _, g, _ in self.tensors])
        params = (_vp * n)(*[p for _, p, _, _ in self.tensors])
        numel = (ctypes.c_int64 * n)(*[m for _, _, _, m in self.tensors])
        arr = lambda ts: (_vp * n)(*[t.data_ptr() for t in ts])
        self._h = _vp()
        _check(_lib().mc_optimizer_create(ctypes.byref(self._h), self.device.index, n, params, arr(self.exp_avg), arr(self.exp_avg_sq), numel),
               'mc_optimizer_create')

    def step(self) -> torch.Tensor:
        g = self.param_groups[0]
        self.step_count += 1
        _check(_lib().mc_optimizer_step(self._h, self._grads, self.step_count, float(g['lr']), float(g['betas'][0]), float(g['betas'][1]),
                                        float(g['eps']), float(g['weight_decay']), float(self.max_norm), self.total_norm.data_ptr(),
                                        _stream(self.device)), 'mc_optimizer_step')
        return self.total_norm

    def zero_grad(self, set_to_none: bool = True):
        pass                                   # the backward pass zeroes what it accumulates into

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            _lib().mc_optimizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
