"""ctypes binding of ``libmonocon_b200.so`` (C ABI in ``include/monocon_b200.h``).

There is deliberately no fallback: if the shared library is missing or no B200 is visible the
calls raise -- the product path never silently runs on the CPU or through PyTorch eager ops.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmonocon_b200.so')

MC_PREC_BF16 = 0
MC_PREC_FP32 = 1          # fp32 storage, FFMA convolutions (training; the slow twin of the mode below)
MC_PREC_FP32_TC = 2       # fp32-accurate results on the tensor cores: fp16 hi + lo planes, three tcgen05 MMAs per K-block
# precision names of Engine / conv2d: 'fp32' is the reference's arithmetic (TF32 off, test.py:30-33) to ~1e-6 per layer and runs
# on the tensor cores; 'fp32_simt' keeps the FFMA kernels (what train-mode engines use); 'bf16' is the throughput mode
PRECISIONS = {'bf16': MC_PREC_BF16, 'fp32': MC_PREC_FP32_TC, 'fp32_simt': MC_PREC_FP32}
MC_CONV_AUTO = 0
MC_CONV_SIMT = 1
# neck_variant of mc_create_ex: the reference's plain-convolution IDAUp blocks, or their DCNv2 variant (north_star)
MC_NECK_CONV = 0
MC_NECK_DCN = 1

PRED_NAMES = ('center_heatmap_pred', 'kpt_heatmap_pred', 'wh_pred', 'offset_pred', 'kpt_heatmap_offset_pred',
              'center2kpt_offset_pred', 'dim_pred', 'depth_pred', 'alpha_cls_pred', 'alpha_offset_pred')
PRED_CHANNELS = (3, 9, 2, 2, 2, 18, 3, 2, 12, 12)

EXPORTS = ('mc_create', 'mc_create_ex', 'mc_deform_conv2d', 'mc_set_param', 'mc_finalize_params', 'mc_refresh_params', 'mc_forward', 'mc_decode', 'mc_infer_host',
           'mc_infer_device', 'mc_infer_host_submit', 'mc_infer_host_wait', 'mc_infer_host_u8_submit', 'mc_get_pred_ptrs', 'mc_copy_pred', 'mc_set_option', 'mc_workspace_bytes', 'mc_num_kernel_launches',
           'mc_flops_per_image', 'mc_bytes_per_image', 'mc_last_error', 'mc_destroy', 'mc_debug_tensor_shape',
           'mc_debug_tensor', 'mc_conv2d', 'mc_num_stages', 'mc_stage_info', 'mc_profile_stages',
           'mc_kitti_boxes', 'mc_calibrate_scales', 'mc_scale_status',
           # uint8 input pipeline (Normalize + Pad + ToTensor fused into the input packing)
           'mc_set_normalization', 'mc_forward_u8', 'mc_infer_device_u8',
           # peer-memory all-gather of the decode outputs (dist.PeerGather)
           'mc_gather_create', 'mc_gather_connect', 'mc_gather_slot_bytes', 'mc_gather_buffer', 'mc_infer_device_gather',
           'mc_gather_wait',
           # train-mode forward (first half of the training step)
           'mc_forward_train', 'mc_get_buffer', 'mc_train_generation',
           # KITTI evaluation overlaps (eval_ops.py)
           'mc_rotate_iou', 'mc_box3d_overlap', 'mc_eval_last_error',
           # training-side rows (train_ops.py)
           'mc_generate_targets', 'mc_losses', 'mc_losses_workspace_bytes', 'mc_optimizer_create', 'mc_optimizer_step',
           'mc_optimizer_destroy', 'mc_train_last_error',
           # backward kernels of the training step (experimental; csrc/train_backward.cu)
           'mc_bw_conv', 'mc_bw_batchnorm', 'mc_bw_colsum', 'mc_bw_maxpool2', 'mc_bw_upsample2', 'mc_bw_heads_scratch_bytes',
           'mc_bw_heads', 'mc_bw_last_error', 'mc_bw_run_graph', 'mc_backward_train', 'mc_get_grad', 'mc_get_param',
           'mc_num_train_tensors', 'mc_train_tensor', 'mc_debug_bw_graph', 'mc_bw_run_graph_range',
           'mc_num_backward_stages', 'mc_backward_train_segment',
           # bf16 tensor-core training kernels (csrc/wgrad_tc.cu, csrc/train_tc.cu)
           'mc_conv2d_wgrad_tc', 'mc_debug_train_dump')

_lib = None


class EngineError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise EngineError(f'{LIB_PATH} is missing; run `python -m monocon_pytorch_b200.build`')
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    declare_signatures(lib)
    _lib = lib
    return lib


def declare_signatures(lib: ctypes.CDLL) -> None:
    """ctypes prototypes of the C ABI (include/monocon_b200.h).  Separate from load_library so that the CPU test of the
    engine's host logic (tests/test_host_engine.py) can apply the same prototypes to its stand-in build."""
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.mc_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci]
    if hasattr(lib, 'mc_create_ex'):                 # absent from the host stand-in of the tests
        lib.mc_create_ex.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, ci]
        lib.mc_deform_conv2d.argtypes = [ci, ci, vp, ci, ci, ci, ci, vp, vp, vp, vp, ci, ci, vp, vp, ctypes.c_char_p, ci]
    lib.mc_set_param.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(ctypes.c_int64), ci]
    lib.mc_finalize_params.argtypes = [vp, ci]
    lib.mc_refresh_params.argtypes = [vp]
    lib.mc_forward.argtypes = [vp, vp, ci, ctypes.POINTER(vp), vp]
    lib.mc_decode.argtypes = [vp, ctypes.POINTER(vp), ci, vp, vp, ci, ci, ci, cf, vp, vp, vp, vp, vp, vp]
    lib.mc_infer_host.argtypes = [vp, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp]
    lib.mc_infer_device.argtypes = [vp, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp]
    lib.mc_forward_train.argtypes = [vp, vp, ci, ctypes.POINTER(vp), vp]
    lib.mc_get_buffer.argtypes = [vp, ctypes.c_char_p, vp, ci]
    lib.mc_train_generation.argtypes = [vp]
    lib.mc_train_generation.restype = ctypes.c_longlong
    lib.mc_backward_train.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, vp]
    lib.mc_get_grad.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int64]
    lib.mc_get_param.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int64]
    lib.mc_num_train_tensors.argtypes = [vp]
    lib.mc_train_tensor.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ci),
                                    ctypes.c_char_p, ci]
    lib.mc_num_backward_stages.argtypes = [vp]
    lib.mc_backward_train_segment.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, ci, vp]
    lib.mc_kitti_boxes.argtypes = [ci, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp]
    lib.mc_set_normalization.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.mc_forward_u8.argtypes = [vp, vp, vp, ci, ci, ci, ctypes.POINTER(vp), vp]
    lib.mc_infer_device_u8.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp]
    lib.mc_gather_create.argtypes = [vp, ci, ci, ci, vp]
    lib.mc_gather_connect.argtypes = [vp, vp]
    lib.mc_gather_slot_bytes.argtypes = [vp]
    lib.mc_gather_slot_bytes.restype = ctypes.c_size_t
    lib.mc_gather_buffer.argtypes = [vp, ci, ctypes.POINTER(vp)]
    lib.mc_infer_device_gather.argtypes = [vp, vp, ci, vp, vp, cf, ci, vp]
    lib.mc_gather_wait.argtypes = [vp, ci, vp]
    lib.mc_infer_host_submit.argtypes = [vp, ci, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp]
    lib.mc_infer_host_wait.argtypes = [vp, ci]
    lib.mc_infer_host_u8_submit.argtypes = [vp, ci, vp, vp, ci, ci, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp]
    lib.mc_get_pred_ptrs.argtypes = [vp, ctypes.POINTER(vp)]
    lib.mc_copy_pred.argtypes = [vp, ci, ctypes.POINTER(vp), vp]
    lib.mc_set_option.argtypes = [vp, ctypes.c_char_p, ci]
    lib.mc_calibrate_scales.argtypes = [vp, vp, ci, vp]
    lib.mc_scale_status.argtypes = [vp, ctypes.POINTER(cf), ctypes.POINTER(ci)]
    lib.mc_workspace_bytes.argtypes = [vp]
    lib.mc_workspace_bytes.restype = ctypes.c_size_t
    lib.mc_num_kernel_launches.argtypes = [vp]
    lib.mc_flops_per_image.argtypes = [vp]
    lib.mc_flops_per_image.restype = ctypes.c_double
    lib.mc_bytes_per_image.argtypes = [vp]
    lib.mc_bytes_per_image.restype = ctypes.c_double
    lib.mc_last_error.argtypes = [vp]
    lib.mc_last_error.restype = ctypes.c_char_p
    lib.mc_destroy.argtypes = [vp]
    lib.mc_destroy.restype = None
    lib.mc_debug_tensor_shape.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.mc_debug_tensor.argtypes = [vp, ctypes.c_char_p, ci, vp, vp]
    lib.mc_conv2d.argtypes = [ci, ci, ci, vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, vp, vp, vp, ci, ci, vp, vp,
                              ctypes.c_char_p, ci]
    if hasattr(lib, 'mc_conv2d_wgrad_tc'):           # stand-alone tensor-core operator entries: absent from the host stand-in of the tests
        lib.mc_conv2d_wgrad_tc.argtypes = [ci, vp, ci, ci, ci, ci, vp, ci, ci, ci, vp, vp, ctypes.c_char_p, ci]
    lib.mc_num_stages.argtypes = [vp]
    lib.mc_stage_info.argtypes = [vp, ci, ctypes.c_char_p, ci, ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ci)]
    lib.mc_profile_stages.argtypes = [vp, vp, ci, vp, vp, ci, ctypes.POINTER(ctypes.c_float), vp]


def kitti_boxes(box3d: torch.Tensor, valid: torch.Tensor, P2: torch.Tensor, img_hw: torch.Tensor):
    """Device-side get_valid_bboxes_3d / convert_to_kitti_3d arithmetic (mc_kitti_boxes): (B,K,7) boxes, (B,K) valid, (B,3,4)
    P2, (B,2) int32 original image sizes -> bbox (B,K,4) float64, alpha (B,K) float32, keep (B,K) uint8."""
    lib = load_library()
    if not box3d.is_cuda:
        raise EngineError('kitti_boxes runs on CUDA tensors only (no CPU fallback)')
    dev = box3d.device
    B, K = box3d.shape[:2]
    box3d = box3d.to(torch.float32).contiguous()
    valid = valid.to(torch.uint8).contiguous()
    P2 = P2.to(dev, torch.float32).contiguous()
    img_hw = img_hw.to(dev, torch.int32).contiguous()
    assert tuple(P2.shape) == (B, 3, 4) and tuple(img_hw.shape) == (B, 2)
    bbox = torch.empty(B, K, 4, dtype=torch.float64, device=dev)
    alpha = torch.empty(B, K, dtype=torch.float32, device=dev)
    keep = torch.empty(B, K, dtype=torch.uint8, device=dev)
    rc = lib.mc_kitti_boxes(dev.index, box3d.data_ptr(), valid.data_ptr(), P2.data_ptr(), img_hw.data_ptr(), B, K, bbox.data_ptr(),
                            alpha.data_ptr(), keep.data_ptr(), _stream_ptr(dev))
    if rc != 0:
        raise EngineError('mc_kitti_boxes: ' + lib.mc_last_error(None).decode())
    return bbox, alpha, keep


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """One engine = one (device, max_batch, H, W, precision) plan with packed weights."""

    def __init__(self, device: torch.device, max_batch: int, H: int, W: int, precision: str = 'bf16',
                 conv_impl: int = MC_CONV_AUTO, use_dcn: bool = False):
        self.lib = load_library()
        self.use_dcn = bool(use_dcn)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise EngineError('the MonoCon B200 engine runs on CUDA devices only (no CPU fallback)')
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', self.index)
        self.max_batch, self.H, self.W = int(max_batch), int(H), int(W)
        self.precision = precision
        self.conv_impl = conv_impl
        prec = PRECISIONS[precision]
        if prec == MC_PREC_FP32_TC and conv_impl == MC_CONV_SIMT:
            prec = MC_PREC_FP32                 # the FFMA kernels work on fp32 storage
        self._h = ctypes.c_void_p()
        self._create(prec)
        self.fh, self.fw = self.H // 4, self.W // 4
        self.finalized = False

    def _create(self, prec: int) -> None:
        self.close()
        self.prec = prec
        if self.use_dcn:
            rc = self.lib.mc_create_ex(ctypes.byref(self._h), self.index, self.max_batch, self.H, self.W, prec, MC_NECK_DCN)
        else:
            rc = self.lib.mc_create(ctypes.byref(self._h), self.index, self.max_batch, self.H, self.W, prec)
        if rc != 0:
            raise EngineError('mc_create: ' + self.lib.mc_last_error(None).decode())
        if self.conv_impl != MC_CONV_AUTO:
            self._check(self.lib.mc_set_option(self._h, b'conv_impl', self.conv_impl), 'mc_set_option')

    @property
    def tensor_core_fp32(self) -> bool:
        return self.prec == MC_PREC_FP32_TC

    def calibrate_scales(self, img: torch.Tensor) -> None:
        """fp32-accurate tensor-core engines: fit the per-tensor scales of the fp16 planes to this sample batch (synchronises)."""
        self._check_img(img)
        self._check(self.lib.mc_calibrate_scales(self._h, img.data_ptr(), img.shape[0], _stream_ptr(self.device)), 'mc_calibrate_scales')

    def scale_status(self):
        """(largest stored value / fp16 limit, number of saturated tensors) since the previous call (synchronises)."""
        mf, ns = ctypes.c_float(), ctypes.c_int()
        self._check(self.lib.mc_scale_status(self._h, ctypes.byref(mf), ctypes.byref(ns)), 'mc_scale_status')
        return float(mf.value), int(ns.value)

    # ------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise EngineError(f'{what}: ' + self.lib.mc_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.mc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], training=False) -> None:
        """Hand every floating-point entry of a reference-layout state_dict to the engine and fold.  ``training=True``
        (fp32 engines only) keeps the BatchNorm parameters separate for ``forward_train``; ``training=2`` (experimental)
        also keeps what ``backward_train`` needs (raw convolution outputs, batch statistics, gradient buffers)."""
        if training and getattr(self, 'prec', None) == MC_PREC_FP32_TC:
            self._create(MC_PREC_FP32)          # train-mode engines run the fp32 FFMA kernels
        self._stage(sd)
        self._check(self.lib.mc_finalize_params(self._h, int(training)), 'mc_finalize_params')
        self.finalized = True
        self.training = bool(training)

    def _stage(self, sd: Dict[str, torch.Tensor]) -> None:
        for key, val in sd.items():
            if not torch.is_floating_point(val):
                continue
            t = val.detach().to(dtype=torch.float32).contiguous()
            shape = (ctypes.c_int64 * max(1, t.dim()))(*t.shape)
            self._check(self.lib.mc_set_param(self._h, key.encode(), t.data_ptr(), shape, t.dim()), f'mc_set_param({key})')

    def refresh_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """New values for a loaded engine (same plan, same mode): repacked into the SAME device buffers (mc_refresh_params), so
        captured CUDA graphs, ``train_tensors()`` pointers and optimiser handles stay valid.  What a module-level training loop
        calls after ``optimizer.step()`` instead of rebuilding the engine."""
        if not self.finalized:
            raise EngineError('refresh_state_dict follows load_state_dict')
        self._stage(sd)
        self._check(self.lib.mc_refresh_params(self._h), 'mc_refresh_params')

    def backward_train(self, pred: List[torch.Tensor], dpred: List[torch.Tensor], segments=None, on_segment=None) -> None:
        """EXPERIMENTAL.  The backward pass of the batch ``forward_train`` just ran, on an engine loaded with ``training=2``:
        ``pred`` are the maps it returned, ``dpred`` dL/dpred in the same order (``train_ops.get_losses(..., with_grad=True)``).
        Parameter gradients stay in the engine; read them with ``get_grad``.  ``segments`` = [(first, last), ...] walks the stage
        list in several calls (from its end down to 0) and calls ``on_segment(i)`` after each -- where data-parallel training
        launches the all-reduce of the gradients that are already final (``dist.OverlappedGradientAverager``)."""
        for t in list(pred) + list(dpred):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise EngineError('backward_train: contiguous float32 maps on the engine device')
        pa = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in pred])
        da = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in dpred])
        if segments is None:
            self._check(self.lib.mc_backward_train(self._h, pa, da, pred[0].shape[0], _stream_ptr(self.device)), 'mc_backward_train')
            return
        n = self.num_backward_stages
        assert segments and segments[0][1] == n and segments[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(segments, segments[1:])), \
            'segments walk the stage list from its end to 0 without gaps: [(k1, n), (k2, k1), ..., (0, kj)]'
        for i, (first, last) in enumerate(segments):
            self._check(self.lib.mc_backward_train_segment(self._h, pa, da, pred[0].shape[0], first, last, _stream_ptr(self.device)),
                        'mc_backward_train_segment')
            if on_segment is not None:
                on_segment(i)                  # the gradients of stages [first, n) are final (stream-ordered): launch their exchange here

    @property
    def num_backward_stages(self) -> int:
        n = self.lib.mc_num_backward_stages(self._h)
        if n < 0:
            raise EngineError('num_backward_stages: load the state_dict with training=2')
        return n

    def get_grad(self, key: str, shape) -> torch.Tensor:
        out = torch.empty(tuple(shape), dtype=torch.float32)
        self._check(self.lib.mc_get_grad(self._h, key.encode(), out.data_ptr(), out.numel()), f'mc_get_grad({key})')
        return out

    def get_param(self, key: str, shape) -> torch.Tensor:
        """Current value of one parameter in state_dict layout (the resident optimiser updates the engine's packed copy)."""
        out = torch.empty(tuple(shape), dtype=torch.float32)
        self._check(self.lib.mc_get_param(self._h, key.encode(), out.data_ptr(), out.numel()), f'mc_get_param({key})')
        return out

    def train_tensors(self):
        """[(key, param_ptr, grad_ptr, numel)]: the trainable buffers in the engine's own layout (``training=2`` engines);
        ``self.train_tensor_stages`` holds, in the same order, the stage whose backward finishes each gradient."""
        n = self.lib.mc_num_train_tensors(self._h)
        if n < 0:
            raise EngineError('train_tensors: load the state_dict with training=2')
        out, self.train_tensor_stages = [], []
        for i in range(n):
            p, g, m, st = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int()
            key = ctypes.create_string_buffer(160)
            self._check(self.lib.mc_train_tensor(self._h, i, ctypes.byref(p), ctypes.byref(g), ctypes.byref(m), ctypes.byref(st), key, 160),
                        'mc_train_tensor')
            out.append((key.value.decode(), p.value, g.value, m.value))
            self.train_tensor_stages.append(st.value)
        return out

    def forward_train(self, img: torch.Tensor, out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """MonoConDetector.forward in train() mode up to the prediction maps: batch-statistic BatchNorm, running statistics
        updated inside the engine (read them back with ``get_buffer``).  2 <= B."""
        self._check_img(img)
        B = img.shape[0]
        out = out if out is not None else self.alloc_pred(B)
        arr = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in out])
        self._check(self.lib.mc_forward_train(self._h, img.data_ptr(), B, arr, _stream_ptr(self.device)), 'mc_forward_train')
        return out

    @property
    def train_generation(self) -> int:
        """Counts forward_train calls; backward_train differentiates the LAST one."""
        return int(self.lib.mc_train_generation(self._h))

    def get_buffer(self, key: str, n: int) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.float32)
        self._check(self.lib.mc_get_buffer(self._h, key.encode(), out.data_ptr(), n), f'mc_get_buffer({key})')
        return out

    def set_option(self, name: str, value: int) -> None:
        self._check(self.lib.mc_set_option(self._h, name.encode(), int(value)), 'mc_set_option')

    # ------------------------------------------------------------------------------------------
    def alloc_pred(self, B: int) -> List[torch.Tensor]:
        return [torch.empty((B, c, self.fh, self.fw), dtype=torch.float32, device=self.device) for c in PRED_CHANNELS]

    def forward(self, img: torch.Tensor, out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        self._check_img(img)
        B = img.shape[0]
        out = out if out is not None else self.alloc_pred(B)
        arr = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in out])
        self._check(self.lib.mc_forward(self._h, img.data_ptr(), B, arr, _stream_ptr(self.device)), 'mc_forward')
        return out

    def alloc_decode(self, B: int, topk: int):
        dev = self.device
        return {'box2d': torch.empty((B, topk, 5), dtype=torch.float32, device=dev),
                'box3d': torch.empty((B, topk, 7), dtype=torch.float32, device=dev),
                'labels': torch.empty((B, topk), dtype=torch.int64, device=dev),
                'inds': torch.empty((B, topk), dtype=torch.int64, device=dev),
                'valid': torch.empty((B, topk), dtype=torch.uint8, device=dev)}

    def decode(self, pred: Sequence[torch.Tensor], P2: torch.Tensor, invP: torch.Tensor, img_hw, topk: int = 30,
               thres: float = 0.4, out=None):
        B = pred[0].shape[0]
        for t, c in zip(pred, PRED_CHANNELS):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (B, c, self.fh, self.fw)):
                raise EngineError('decode: prediction maps must be contiguous fp32 CUDA tensors of the engine geometry')
        self._check_calib(P2, invP, B)
        out = out if out is not None else self.alloc_decode(B, topk)
        arr = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in pred])
        self._check(self.lib.mc_decode(self._h, arr, B, P2.data_ptr(), invP.data_ptr(), int(img_hw[0]), int(img_hw[1]),
                                       topk, float(thres), out['box2d'].data_ptr(), out['box3d'].data_ptr(),
                                       out['labels'].data_ptr(), out['inds'].data_ptr(), out['valid'].data_ptr(),
                                       _stream_ptr(self.device)), 'mc_decode')
        return out

    def infer_device(self, img: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, topk: int = 30, thres: float = 0.4,
                     out=None):
        self._check_img(img)
        B = img.shape[0]
        self._check_calib(P2, invP, B)
        out = out if out is not None else self.alloc_decode(B, topk)
        self._check(self.lib.mc_infer_device(self._h, img.data_ptr(), B, P2.data_ptr(), invP.data_ptr(), topk, float(thres),
                                             out['box2d'].data_ptr(), out['box3d'].data_ptr(), out['labels'].data_ptr(),
                                             out['inds'].data_ptr(), out['valid'].data_ptr(), _stream_ptr(self.device)),
                    'mc_infer_device')
        return out

    # ---- uint8 frames in: Normalize + Pad + ToTensor of the reference's test pipeline inside the input-packing kernel ----
    def set_normalization(self, mean: Sequence[float], std: Sequence[float]) -> None:
        m = (ctypes.c_double * 3)(*[float(v) for v in mean])
        s = (ctypes.c_double * 3)(*[float(v) for v in std])
        self._check(self.lib.mc_set_normalization(self._h, m, s), 'mc_set_normalization')

    def _check_u8(self, img_u8: torch.Tensor, hw: torch.Tensor):
        if not (img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] == 3 and img_u8.is_contiguous()):
            raise EngineError('img_u8 must be a contiguous CUDA uint8 tensor of shape (B, H0, W0, 3)')
        B, H0, W0, _ = img_u8.shape
        if not (hw.is_cuda and hw.dtype == torch.int32 and tuple(hw.shape) == (B, 2) and hw.is_contiguous()):
            raise EngineError('hw must be a contiguous CUDA int32 tensor of shape (B, 2): valid (height, width) per frame')
        if B > self.max_batch or H0 > self.H or W0 > self.W:
            raise EngineError(f'frames ({B}, {H0}, {W0}) exceed the engine geometry ({self.max_batch}, {self.H}, {self.W})')
        return B, H0, W0

    def forward_u8(self, img_u8: torch.Tensor, hw: torch.Tensor) -> List[torch.Tensor]:
        """(B, H0, W0, 3) uint8 frames + per-frame valid sizes -> the ten prediction maps (as `forward`)."""
        B, H0, W0 = self._check_u8(img_u8, hw)
        outs = [torch.empty(B, c, self.H // 4, self.W // 4, dtype=torch.float32, device=self.device) for c in PRED_CHANNELS]
        arr = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        self._check(self.lib.mc_forward_u8(self._h, img_u8.data_ptr(), hw.data_ptr(), B, H0, W0, arr, _stream_ptr(self.device)), 'mc_forward_u8')
        return outs

    def infer_device_u8(self, img_u8: torch.Tensor, hw: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, topk: int = 30,
                        thres: float = 0.4, out=None):
        B, H0, W0 = self._check_u8(img_u8, hw)
        self._check_calib(P2, invP, B)
        out = out if out is not None else self.alloc_decode(B, topk)
        self._check(self.lib.mc_infer_device_u8(self._h, img_u8.data_ptr(), hw.data_ptr(), B, H0, W0, P2.data_ptr(), invP.data_ptr(), topk,
                                                float(thres), out['box2d'].data_ptr(), out['box3d'].data_ptr(), out['labels'].data_ptr(),
                                                out['inds'].data_ptr(), out['valid'].data_ptr(), _stream_ptr(self.device)), 'mc_infer_device_u8')
        return out

    # ---- peer-memory all-gather of the decode outputs (multi-GPU inference; see dist.PeerGather) -------------------
    def gather_create(self, world: int, rank: int, topk: int) -> bytes:
        """Allocates this rank's gather block; returns its 64-byte CUDA IPC handle."""
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.mc_gather_create(self._h, world, rank, topk, ctypes.cast(buf, ctypes.c_void_p)), 'mc_gather_create')
        return buf.raw

    def gather_connect(self, handles: bytes) -> None:
        buf = ctypes.create_string_buffer(handles, len(handles))
        self._check(self.lib.mc_gather_connect(self._h, ctypes.cast(buf, ctypes.c_void_p)), 'mc_gather_connect')

    def gather_slot_bytes(self) -> int:
        return int(self.lib.mc_gather_slot_bytes(self._h))

    def gather_buffer_ptr(self, buf: int) -> int:
        p = ctypes.c_void_p()
        self._check(self.lib.mc_gather_buffer(self._h, buf, ctypes.byref(p)), 'mc_gather_buffer')
        return int(p.value)

    def infer_device_gather(self, img: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, buf: int, thres: float = 0.4) -> None:
        self._check_img(img)
        B = img.shape[0]
        self._check_calib(P2, invP, B)
        self._check(self.lib.mc_infer_device_gather(self._h, img.data_ptr(), B, P2.data_ptr(), invP.data_ptr(), float(thres), buf,
                                                    _stream_ptr(self.device)), 'mc_infer_device_gather')

    def gather_wait(self, buf: int) -> None:
        self._check(self.lib.mc_gather_wait(self._h, buf, _stream_ptr(self.device)), 'mc_gather_wait')

    def infer_host(self, img: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, topk: int = 30, thres: float = 0.4,
                   out=None):
        """Host tensors in (ideally pinned), host tensors out; synchronises the current stream."""
        B = img.shape[0]
        if img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous() or tuple(img.shape[1:]) != (3, self.H, self.W):
            raise EngineError('infer_host: img must be a contiguous fp32 host tensor (B,3,H,W) of the engine geometry')
        if out is None:
            out = {'box2d': torch.empty((B, topk, 5), dtype=torch.float32).pin_memory(),
                   'box3d': torch.empty((B, topk, 7), dtype=torch.float32).pin_memory(),
                   'labels': torch.empty((B, topk), dtype=torch.int64).pin_memory(),
                   'inds': torch.empty((B, topk), dtype=torch.int64).pin_memory(),
                   'valid': torch.empty((B, topk), dtype=torch.uint8).pin_memory()}
        P2 = P2.to(torch.float32).contiguous()
        invP = invP.to(torch.float32).contiguous()
        self._check(self.lib.mc_infer_host(self._h, img.data_ptr(), B, P2.data_ptr(), invP.data_ptr(), topk, float(thres),
                                           out['box2d'].data_ptr(), out['box3d'].data_ptr(), out['labels'].data_ptr(),
                                           out['inds'].data_ptr(), out['valid'].data_ptr(), _stream_ptr(self.device)),
                    'mc_infer_host')
        return out

    @staticmethod
    def alloc_host_out(B: int, topk: int):
        return {'box2d': torch.empty((B, topk, 5), dtype=torch.float32).pin_memory(),
                'box3d': torch.empty((B, topk, 7), dtype=torch.float32).pin_memory(),
                'labels': torch.empty((B, topk), dtype=torch.int64).pin_memory(),
                'inds': torch.empty((B, topk), dtype=torch.int64).pin_memory(),
                'valid': torch.empty((B, topk), dtype=torch.uint8).pin_memory()}

    def infer_host_submit(self, slot: int, img: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, out, topk: int = 30,
                          thres: float = 0.4) -> None:
        """Asynchronous, double-buffered host path: enqueue H2D -> forward + decode -> D2H for `slot` (0 or 1) and
        return; `infer_host_wait(slot)` blocks until `out` (pinned host tensors from alloc_host_out) is filled."""
        B = img.shape[0]
        if img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous() or tuple(img.shape[1:]) != (3, self.H, self.W):
            raise EngineError('infer_host_submit: img must be a contiguous fp32 host tensor (B,3,H,W) of the engine geometry')
        for t in (P2, invP):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise EngineError('infer_host_submit: P2 / invP must be contiguous fp32 host tensors')
        self._check(self.lib.mc_infer_host_submit(self._h, slot, img.data_ptr(), B, P2.data_ptr(), invP.data_ptr(), topk,
                                                  float(thres), out['box2d'].data_ptr(), out['box3d'].data_ptr(),
                                                  out['labels'].data_ptr(), out['inds'].data_ptr(), out['valid'].data_ptr()),
                    'mc_infer_host_submit')

    def infer_host_u8_submit(self, slot: int, img_u8: torch.Tensor, hw: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, out,
                             topk: int = 30, thres: float = 0.4) -> None:
        """As infer_host_submit, fed with host uint8 HWC frames (B, H0, W0, 3) + int32 valid sizes (B, 2): the reference's
        Normalize + Pad + ToTensor run on the device inside the input packing; a quarter of the H2D bytes."""
        if img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[3] != 3 or not img_u8.is_contiguous():
            raise EngineError('infer_host_u8_submit: img_u8 must be a contiguous uint8 host tensor (B, H0, W0, 3)')
        B, H0, W0, _ = img_u8.shape
        if hw.is_cuda or hw.dtype != torch.int32 or tuple(hw.shape) != (B, 2) or not hw.is_contiguous():
            raise EngineError('infer_host_u8_submit: hw must be a contiguous int32 host tensor (B, 2)')
        if B > self.max_batch or H0 > self.H or W0 > self.W:
            raise EngineError(f'frames ({B}, {H0}, {W0}) exceed the engine geometry ({self.max_batch}, {self.H}, {self.W})')
        for t in (P2, invP):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise EngineError('infer_host_u8_submit: P2 / invP must be contiguous fp32 host tensors')
        self._check(self.lib.mc_infer_host_u8_submit(self._h, slot, img_u8.data_ptr(), hw.data_ptr(), B, H0, W0, P2.data_ptr(),
                                                     invP.data_ptr(), topk, float(thres), out['box2d'].data_ptr(), out['box3d'].data_ptr(),
                                                     out['labels'].data_ptr(), out['inds'].data_ptr(), out['valid'].data_ptr()),
                    'mc_infer_host_u8_submit')

    def infer_host_wait(self, slot: int) -> None:
        self._check(self.lib.mc_infer_host_wait(self._h, slot), 'mc_infer_host_wait')

    def pred_views(self, B: int) -> List[torch.Tensor]:
        """Copies of the engine-owned maps of the last infer_* call."""
        outs = self.alloc_pred(B)
        arr = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in outs])
        self._check(self.lib.mc_copy_pred(self._h, B, arr, _stream_ptr(self.device)), 'mc_copy_pred')
        return outs

    def profile_stages(self, img: torch.Tensor, P2: torch.Tensor, invP: torch.Tensor, iters: int = 5):
        """Per-stage device time (CUDA events, eager launches): list of dicts name/ms/flops/bytes/tensor_core."""
        self._check_img(img)
        B = img.shape[0]
        self._check_calib(P2, invP, B)
        n = int(self.lib.mc_num_stages(self._h))
        ms = (ctypes.c_float * n)()
        self._check(self.lib.mc_profile_stages(self._h, img.data_ptr(), B, P2.data_ptr(), invP.data_ptr(), iters, ms,
                                               _stream_ptr(self.device)), 'mc_profile_stages')
        out = []
        for i in range(n):
            name = ctypes.create_string_buffer(128)
            fl, by, tc = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
            self._check(self.lib.mc_stage_info(self._h, i, name, 128, ctypes.byref(fl), ctypes.byref(by), ctypes.byref(tc)),
                        'mc_stage_info')
            out.append({'name': name.value.decode(), 'ms': float(ms[i]), 'flops': fl.value * B, 'bytes': by.value * B,
                        'tensor_core': bool(tc.value), 'impl': int(tc.value)})
        return out

    def debug_tensor(self, name: str, B: int) -> torch.Tensor:
        c, hh, ww = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.mc_debug_tensor_shape(self._h, name.encode(), ctypes.byref(c), ctypes.byref(hh), ctypes.byref(ww)),
                    'mc_debug_tensor_shape')
        out = torch.empty((B, c.value, hh.value, ww.value), dtype=torch.float32, device=self.device)
        self._check(self.lib.mc_debug_tensor(self._h, name.encode(), B, out.data_ptr(), _stream_ptr(self.device)), 'mc_debug_tensor')
        return out

    # ------------------------------------------------------------------------------------------
    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.mc_workspace_bytes(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.mc_num_kernel_launches(self._h))

    @property
    def flops_per_image(self) -> float:
        return float(self.lib.mc_flops_per_image(self._h))

    @property
    def bytes_per_image(self) -> float:
        return float(self.lib.mc_bytes_per_image(self._h))

    def _check_img(self, img: torch.Tensor) -> None:
        if not (img.is_cuda and img.dtype == torch.float32 and img.is_contiguous() and img.dim() == 4
                and tuple(img.shape[1:]) == (3, self.H, self.W) and img.shape[0] <= self.max_batch
                and img.device.index == self.index):
            raise EngineError(f'img must be a contiguous fp32 CUDA tensor (B<={self.max_batch},3,{self.H},{self.W}) '
                              f'on cuda:{self.index}; got {tuple(img.shape)} {img.dtype} {img.device}')

    def _check_calib(self, P2: torch.Tensor, invP: torch.Tensor, B: int) -> None:
        ok = (P2.is_cuda and invP.is_cuda and P2.dtype == torch.float32 and invP.dtype == torch.float32
              and P2.is_contiguous() and invP.is_contiguous() and tuple(P2.shape) == (B, 3, 4) and tuple(invP.shape) == (B, 4, 4))
        if not ok:
            raise EngineError('P2 must be (B,3,4) and invP (B,4,4), contiguous fp32 CUDA tensors')


def inverse_viewpad(P2: np.ndarray) -> torch.Tensor:
    """inv(4x4-padded P2) per image, computed on the CPU in fp32 exactly as the reference does
    (model/dense_heads/monocon_heads.py:543-546)."""
    out = []
    for p in np.asarray(P2, dtype=np.float32).reshape(-1, 3, 4):
        viewpad = torch.eye(4)
        viewpad[:3, :4] = torch.from_numpy(p.copy())
        out.append(torch.inverse(viewpad))
    return torch.stack(out, 0).contiguous()


def conv2d(x: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, stride: int = 1, pad: int = 0,
           residual: Optional[torch.Tensor] = None, relu: bool = False, split: int = 1, precision: str = 'bf16',
           conv_impl: int = MC_CONV_AUTO) -> torch.Tensor:
    """Stand-alone operator entry (kernel-level parity tests): conv + folded BN + residual + ReLU."""
    lib = load_library()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    B, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    err = ctypes.create_string_buffer(1024)
    prec = PRECISIONS[precision]
    if prec == MC_PREC_FP32_TC and conv_impl == MC_CONV_SIMT:
        prec = MC_PREC_FP32
    w = w.contiguous().float(); scale = scale.contiguous().float(); shift = shift.contiguous().float()
    if residual is not None:
        residual = residual.contiguous().float()
    rc = lib.mc_conv2d(x.device.index or 0, prec, conv_impl, x.data_ptr(), B, Cin, H, W, w.data_ptr(), Cout, k, stride, pad,
                       scale.data_ptr(), shift.data_ptr(), _ptr(residual), int(relu), split, y.data_ptr(),
                       _stream_ptr(x.device), err, 1024)
    if rc != 0:
        raise EngineError('mc_conv2d: ' + err.value.decode())
    return y


def deform_conv2d(x: torch.Tensor, offset: torch.Tensor, mask: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                  split: int = 1, precision: str = 'fp32') -> torch.Tensor:
    """Stand-alone operator entry (kernel-level parity tests): the modulated deformable 3x3 convolution of the DCN neck variant,
    arguments as torchvision.ops.deform_conv2d(x, offset, w, bias, padding=1, mask=mask) takes them (fp32 CUDA tensors)."""
    lib = load_library()
    for t in (x, offset, mask, w):
        assert t.is_cuda and t.dtype == torch.float32
    x, offset, mask, w = x.contiguous(), offset.contiguous(), mask.contiguous(), w.contiguous()
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    assert tuple(w.shape) == (Cout, Cin, 3, 3) and tuple(offset.shape) == (B, 18, H, W) and tuple(mask.shape) == (B, 9, H, W)
    if bias is not None:
        bias = bias.contiguous().float()
    y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    err = ctypes.create_string_buffer(1024)
    rc = lib.mc_deform_conv2d(x.device.index or 0, PRECISIONS[precision], x.data_ptr(), B, Cin, H, W, offset.data_ptr(), mask.data_ptr(),
                              w.data_ptr(), _ptr(bias), Cout, split, y.data_ptr(), _stream_ptr(x.device), err, 1024)
    if rc != 0:
        raise EngineError('mc_deform_conv2d: ' + err.value.decode())
    return y


def conv2d_wgrad_tc(x: torch.Tensor, dy: torch.Tensor, k: int, split: int = 1) -> torch.Tensor:
    """Stand-alone operator entry (kernel-level parity tests): tensor-core weight gradient of a k x k / stride 1 convolution in the
    bf16 training arithmetic.  x (B,Cin,H,W), dy (B,Cout,H,W) fp32 CUDA -> dw (Cout,Cin,k,k) fp32."""
    lib = load_library()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and dy.is_cuda and dy.dtype == torch.float32 and dy.is_contiguous()
    B, Cin, H, W = x.shape
    Cout = dy.shape[1]
    assert tuple(dy.shape) == (B, Cout, H, W)
    Cst = 8 if (k == 7 and Cin == 3) else Cin       # the stem's image is stored with 8 channels (3 + 5 zeros)
    dw = torch.empty((k * k, Cst, Cout), dtype=torch.float32, device=x.device)
    err = ctypes.create_string_buffer(1024)
    rc = lib.mc_conv2d_wgrad_tc(x.device.index or 0, x.data_ptr(), B, Cin, H, W, dy.data_ptr(), Cout, k, split, dw.data_ptr(),
                                _stream_ptr(x.device), err, 1024)
    if rc != 0:
        raise EngineError('mc_conv2d_wgrad_tc: ' + err.value.decode())
    return dw[:, :Cin].permute(2, 1, 0).reshape(Cout, Cin, k, k).contiguous()
