"""TEST INFRASTRUCTURE ONLY (oracle) — CPU restatement of the reference's per-sample input pipeline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path
(da-detect_b200/) never does.

What it restates (reference file:line):
  * maskrcnn_benchmark/data/transforms/build.py:5-32      — Compose([Resize, RandomHorizontalFlip, ToTensor, Normalize])
  * maskrcnn_benchmark/data/transforms/transforms.py:35-69 — Resize.get_size (min/max size rule) + F.resize
  * transforms.py:72-81                                   — RandomHorizontalFlip (random.random() < prob)
  * transforms.py:84-98                                   — ToTensor, Normalize(to_bgr255): image[[2,1,0]] * 255, (x - mean) / std
  * maskrcnn_benchmark/structures/image_list.py:49-91      — to_image_list: zero padding up to SIZE_DIVISIBILITY
  * maskrcnn_benchmark/data/collate_batch.py:39-56         — BatchCollator

Third-party arithmetic on this path that is NOT under /root/reference: `F.resize` on a PIL image is
`PIL.Image.resize(size, BILINEAR)` (torchvision.transforms.functional.resize -> _functional_pil.resize); the
reference pins no version (requirements.txt), this container has Pillow 12.2.0 / torchvision 0.26.0.  Pillow's
published algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc, two passes — horizontal first — with a uint8
intermediate) is restated below in numpy integer arithmetic.  Pinned by tests/test_preprocess_cpu.py against
Pillow itself and against the real reference `build_transforms` pipeline (tests/golden/preprocess_*.pt, made by
oracle/make_golden.py preprocess).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are fixed point with 22 fractional bits


def bilinear_filter(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc.
    Returns (bounds int32 [out,2] = (xmin, count), kk int32 [out, ksize], ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale                      # bilinear support = 1
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = bilinear_filter((np.arange(xmax) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                   # the C loop accumulates left to right in double
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        q = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS)).astype(np.int64)
        kk[xx, :xmax] = q.astype(np.int32)            # C cast: truncation toward zero (astype does the same)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img, out_size, axis):
    """One pass of ImagingResample{Horizontal,Vertical}_8bpc on a uint8 [H,W,C] array."""
    in_size = img.shape[axis]
    bounds, kk, ksize = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(img, out_h, out_w):
    """PIL.Image.resize((out_w, out_h), BILINEAR) on a uint8 [H,W,3] array: horizontal pass, then vertical pass;
    a pass whose size does not change is skipped (ImagingResample: need_horizontal / need_vertical)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    if img.shape[1] != out_w:
        img = resample_axis(img, out_w, 1)
    if img.shape[0] != out_h:
        img = resample_axis(img, out_h, 0)
    return img


def get_size(image_size, min_size, max_size):
    """Resize.get_size (transforms.py:43-63) for one already-chosen min_size; image_size = (w, h); returns (oh, ow)."""
    w, h = image_size
    size = min_size
    if max_size is not None:
        min_original_size = float(min((w, h)))
        max_original_size = float(max((w, h)))
        if max_original_size / min_original_size * size > max_size:
            size = int(round(max_size * min_original_size / max_original_size))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        ow = size
        oh = int(size * h / w)
    else:
        oh = size
        ow = int(size * w / h)
    return (oh, ow)


def to_tensor_normalize(img_u8, mean, std, to_bgr255=True):
    """ToTensor + Normalize (transforms.py:84-98) with torch's fp32 operation order: u8 -> f32, / 255,
    channel swap, * 255, - mean, / std.  Returns float32 [3,H,W]."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    if to_bgr255:
        x = x[[2, 1, 0]] * np.float32(255)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return ((x - m) / s).astype(np.float32)


def transform_image(img_u8, out_hw, flip, mean, std, to_bgr255=True):
    """Resize -> (flip) -> ToTensor -> Normalize for one RGB uint8 [H,W,3] image."""
    r = pil_resize_bilinear(img_u8, out_hw[0], out_hw[1])
    if flip:
        r = r[:, ::-1]
    return to_tensor_normalize(np.ascontiguousarray(r), mean, std, to_bgr255)


def collate(tensors, size_divisible):
    """to_image_list on a list of [3,h,w] arrays (image_list.py:66-88): zero padding to the per-axis maximum
    rounded up to size_divisible.  Returns (batch [N,3,Hp,Wp], [(h,w)...])."""
    h = max(t.shape[1] for t in tensors)
    w = max(t.shape[2] for t in tensors)
    if size_divisible > 0:
        h = int(math.ceil(h / size_divisible) * size_divisible)
        w = int(math.ceil(w / size_divisible) * size_divisible)
    out = np.zeros((len(tensors), 3, h, w), dtype=np.float32)
    for o, t in zip(out, tensors):
        o[:, :t.shape[1], :t.shape[2]] = t
    return out, [tuple(t.shape[1:]) for t in tensors]


def resize_boxes(boxes, src_wh, dst_wh):
    """BoxList.resize (structures/bounding_box.py:91-129): per-axis ratios; equal ratios take the scalar path
    `box * ratio` (same arithmetic)."""
    rw, rh = float(dst_wh[0]) / float(src_wh[0]), float(dst_wh[1]) / float(src_wh[1])
    b = np.asarray(boxes, dtype=np.float32).copy()
    if rw == rh:
        return b * np.float32(rw)
    b[:, 0::2] *= np.float32(rw)
    b[:, 1::2] *= np.float32(rh)
    return b


def hflip_boxes(boxes, width):
    """BoxList.transpose(FLIP_LEFT_RIGHT) (bounding_box.py:131-164): TO_REMOVE = 1."""
    b = np.asarray(boxes, dtype=np.float32).copy()
    x1 = np.float32(width) - b[:, 2] - np.float32(1)
    x2 = np.float32(width) - b[:, 0] - np.float32(1)
    b[:, 0], b[:, 2] = x1, x2
    return b
