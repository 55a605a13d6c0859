"""Drop-ins for the reference's KITTI evaluation overlap functions (engine/kitti_eval/rotate_iou.py:337-379,
engine/kitti_eval/eval.py:121-164) over the C ABI (csrc/kernels_eval.cu): same names, numpy in / numpy out, same argument
meaning.  They need a CUDA device like the reference's numba.cuda kernel does; there is no CPU fallback."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .engine import EngineError, load_library

_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.mc_rotate_iou.argtypes = [ci, vp, vp, ci, ci, ci, vp, vp]
        lib.mc_box3d_overlap.argtypes = [ci, vp, vp, ci, ci, ci, vp, vp]
        lib.mc_eval_last_error.restype = ctypes.c_char_p
        _bound = True
    return lib


def _run(fn_name, boxes, query_boxes, np_dtype, width, criterion, device_id):
    if not torch.cuda.is_available():
        raise EngineError('the evaluation overlap kernels need a CUDA device (no CPU fallback)')
    lib = _lib()
    dev = torch.device('cuda', device_id)
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np_dtype).reshape(-1, width))
    q = np.ascontiguousarray(np.asarray(query_boxes, dtype=np_dtype).reshape(-1, width))
    N, K = b.shape[0], q.shape[0]
    out = torch.zeros((N, K), dtype=torch.float32, device=dev)
    if N and K:
        tb, tq = torch.from_numpy(b).to(dev), torch.from_numpy(q).to(dev)
        rc = getattr(lib, fn_name)(device_id, tb.data_ptr(), tq.data_ptr(), N, K, int(criterion), out.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise EngineError(f'{fn_name}: ' + lib.mc_eval_last_error().decode())
    return out.cpu().numpy()


def rotate_iou_gpu_eval(boxes, query_boxes, criterion: int = -1, device_id: int = 0) -> np.ndarray:
    """engine/kitti_eval/rotate_iou.py:337: (N,5), (K,5) BEV boxes [cx, cy, dx, dy, angle] -> (N,K) float32 (the reference
    casts its inputs to float32 first, so its `.astype(boxes.dtype)` is float32 too)."""
    return _run('mc_rotate_iou', boxes, query_boxes, np.float32, 5, criterion, device_id)


def bev_box_overlap(boxes, qboxes, criterion: int = -1) -> np.ndarray:
    """engine/kitti_eval/eval.py:121-125."""
    return rotate_iou_gpu_eval(boxes, qboxes, criterion)


def d3_box_overlap(boxes, qboxes, criterion: int = -1, device_id: int = 0) -> np.ndarray:
    """engine/kitti_eval/eval.py:159-164: camera boxes (N,7), (K,7) [x, y, z, l, h, w, ry] -> (N,K) float32."""
    return _run('mc_box3d_overlap', boxes, qboxes, np.float64, 7, criterion, device_id)
