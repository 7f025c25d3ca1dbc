"""Drop-in for the three hot symbols of the reference's pybind module `maskrcnn_benchmark._C`
(maskrcnn_benchmark/csrc/vision.cpp:7-15) with identical signatures and tensor layouts, backed by the
C ABI.  A reference checkout uses it with `from dadetect_b200 import _C` in place of
`from maskrcnn_benchmark import _C` (layers/roi_align.py:8, layers/nms.py:3) — see INTEGRATION.md.

  roi_align_forward(input[N,C,H,W], rois[K,5], spatial_scale, pooled_h, pooled_w, sampling_ratio) -> [K,C,ph,pw]
  roi_align_backward(grad[K,C,ph,pw], rois, spatial_scale, pooled_h, pooled_w, N, C, H, W, sampling_ratio) -> [N,C,H,W]
  nms(dets[N,4], scores[N], threshold) -> int64 kept indices, ascending (GPU semantics: IoU > threshold)

Errors surface as RuntimeError like the reference's AT_ASSERTM / THCudaCheck.  CUDA tensors only: the
ops that are unused by the DA path (roi_pool_*, sigmoid_focalloss_*) raise.
"""
import torch

from . import _lib
from . import ops as _ops
from .ops import _chk, _ptr, _stream


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    x = _chk(input, name="input")
    r = _chk(rois, name="rois")
    n, c, h, w = x.shape
    k = r.shape[0]
    out = torch.empty((k, c, pooled_height, pooled_width), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    _lib.call("dd_roi_align_forward_nchw", _ptr(x), _ptr(r), _ptr(out), n, c, h, w, k, float(spatial_scale),
              int(pooled_height), int(pooled_width), int(sampling_ratio), _stream())
    return out


def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
                       sampling_ratio):
    g = _chk(grad, name="grad")
    r = _chk(rois, name="rois")
    out = torch.zeros((batch_size, channels, height, width), dtype=torch.float32, device=g.device)
    if g.numel() == 0:
        return out
    _lib.call("dd_roi_align_backward_nchw", _ptr(g), _ptr(r), _ptr(out), int(batch_size), int(channels), int(height),
              int(width), r.shape[0], float(spatial_scale), int(pooled_height), int(pooled_width), int(sampling_ratio),
              _stream())
    return out


def nms(dets, scores, threshold):
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device="cpu")   # the reference returns a CPU empty tensor (nms.h:17-18)
    return _ops.nms(dets, scores, threshold)


def _unused(*args, **kwargs):
    raise RuntimeError("this op is not on the DA Faster R-CNN path and is not provided by dadetect_b200")


roi_pool_forward = roi_pool_backward = sigmoid_focalloss_forward = sigmoid_focalloss_backward = _unused
