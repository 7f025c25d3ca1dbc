"""GPU input pipeline: the reference's per-sample CPU transform chain and the batch padding as one
kernel launch per image.

Reference                                                        here
  data/transforms/build.py:5-32      build_transforms(cfg, is_train)  -> build_transforms(cfg, is_train, device)
  transforms.py:35-69                Resize (PIL bilinear)            \
  transforms.py:72-81                RandomHorizontalFlip              |  dd_preprocess_image: uint8 HWC in,
  transforms.py:84-86                ToTensor                          |  padded float [3,Hp,Wp] batch slot out
  transforms.py:89-98                Normalize(to_bgr255)              |
  structures/image_list.py:66-88     to_image_list zero padding       /
  data/collate_batch.py:39-56        BatchCollator                    -> DeviceBatchCollator

The host keeps only the decisions: the output size rule (`Resize.get_size`), the flip draw, the box transforms on the
targets.  Both random draws use Python's `random` in the reference's order (per sample: `random.choice(min_size)`,
then `random.random()`), so a seeded run makes the same decisions as the reference pipeline.  The decoded image
travels to the GPU as uint8 (4x fewer H2D bytes than the reference's fp32 tensor; 1/7 at 2048x1024 -> 1200x600).
"""
import ctypes
import math
import random

import numpy as np
import torch

from .. import _lib
from ..structures import ImageList
from ..structures.bounding_box import FLIP_LEFT_RIGHT


def get_size(image_size, size, max_size):
    """Resize.get_size (transforms.py:43-63) for an already chosen `size`; image_size = (w, h) -> (oh, ow)."""
    w, h = image_size
    if max_size is not None:
        min_original_size = float(min((w, h)))
        max_original_size = float(max((w, h)))
        if max_original_size / min_original_size * size > max_size:
            size = int(round(max_size * min_original_size / max_original_size))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


def resample_coeffs(in_size, out_size):
    """Pillow's bilinear coefficient table of one axis as host int32 tensors (bounds [out,2], kk [out,ksize])."""
    lib = _lib.load()
    ks = lib.dd_resample_ksize(int(in_size), int(out_size))
    if ks < 0:
        raise ValueError("resample_coeffs: sizes must be positive, got {} -> {}".format(in_size, out_size))
    bounds = torch.empty((out_size, 2), dtype=torch.int32)
    kk = torch.empty((out_size, ks), dtype=torch.int32)
    _lib.call("dd_resample_coeffs", int(in_size), int(out_size), ctypes.c_void_p(bounds.data_ptr()),
              ctypes.c_void_p(kk.data_ptr()))
    return bounds, kk


class DeviceTransform(object):
    """Compose([Resize(min_size, max_size), RandomHorizontalFlip(flip_prob), ToTensor(), Normalize(mean, std,
    to_bgr255)]) with the image arithmetic on the GPU.  `plan` makes the host decisions for one sample, `run` enqueues
    the kernel that writes one image into its slot of a batch tensor."""

    def __init__(self, min_size, max_size, flip_prob, mean, std, to_bgr255=True, device="cuda"):
        if not isinstance(min_size, (list, tuple)):
            min_size = (min_size,)
        self.min_size = tuple(min_size)
        self.max_size = max_size
        self.flip_prob = flip_prob
        self.mean = (ctypes.c_float * 3)(*[float(m) for m in mean])
        self.std = (ctypes.c_float * 3)(*[float(s) for s in std])
        self.to_bgr255 = bool(to_bgr255)
        self.device = torch.device(device)
        self._coeffs = {}                      # (in, out) -> device (bounds, kk); a handful of sizes per dataset

    def plan(self, image_size):
        """image_size = (w, h) of the decoded image -> ((oh, ow), flip) with the reference's draw order."""
        size = random.choice(self.min_size)
        out_hw = get_size(image_size, size, self.max_size)
        flip = random.random() < self.flip_prob
        return out_hw, flip

    def transform_target(self, target, out_hw, flip):
        if target is None:
            return None
        target = target.resize((out_hw[1], out_hw[0]))
        if flip:
            target = target.transpose(FLIP_LEFT_RIGHT)
        return target

    def _device_coeffs(self, in_size, out_size):
        key = (int(in_size), int(out_size))
        ent = self._coeffs.get(key)
        if ent is None:
            b, k = resample_coeffs(*key)
            ent = (b.to(self.device), k.to(self.device))
            self._coeffs[key] = ent
        return ent

    def run(self, image_u8, out_hw, flip, dst):
        """image_u8: uint8 [h, w, 3|4] on self.device (row-contiguous pixels); dst: float32 [3, Hp, Wp] slot."""
        if image_u8.dtype != torch.uint8 or image_u8.dim() != 3 or image_u8.shape[2] not in (3, 4):
            raise RuntimeError("dadetect_b200: images must be uint8 [H,W,3|4], got {} {}".format(
                image_u8.dtype, tuple(image_u8.shape)))
        if not image_u8.is_cuda or not dst.is_cuda:
            raise RuntimeError("dadetect_b200: the input pipeline runs on the GPU — there is no CPU path")
        if image_u8.stride(2) != 1 or image_u8.stride(1) != image_u8.shape[2]:
            image_u8 = image_u8.contiguous()
        h, w, ps = image_u8.shape
        oh, ow = out_hw
        xb, xk = self._device_coeffs(w, ow)
        yb, yk = self._device_coeffs(h, oh)
        assert dst.dtype == torch.float32 and dst.is_contiguous() and dst.shape[0] == 3
        _lib.call("dd_preprocess_image", ctypes.c_void_p(image_u8.data_ptr()), h, w, ps, int(image_u8.stride(0)),
                  ctypes.c_void_p(xb.data_ptr()), ctypes.c_void_p(xk.data_ptr()), xk.shape[1],
                  ctypes.c_void_p(yb.data_ptr()), ctypes.c_void_p(yk.data_ptr()), yk.shape[1], oh, ow,
                  1 if flip else 0, 1 if self.to_bgr255 else 0, self.mean, self.std,
                  ctypes.c_void_p(dst.data_ptr()), dst.shape[1], dst.shape[2],
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


def build_transforms(cfg, is_train=True, device="cuda"):
    """data/transforms/build.py:5-32 (flip_prob is hard-coded to 0.5 for training there, :9)."""
    if is_train:
        min_size, max_size, flip_prob = cfg.INPUT.MIN_SIZE_TRAIN, cfg.INPUT.MAX_SIZE_TRAIN, 0.5
    else:
        min_size, max_size, flip_prob = cfg.INPUT.MIN_SIZE_TEST, cfg.INPUT.MAX_SIZE_TEST, 0
    return DeviceTransform(min_size, max_size, flip_prob, cfg.INPUT.PIXEL_MEAN, cfg.INPUT.PIXEL_STD,
                           cfg.INPUT.TO_BGR255, device)


class DeviceBatchCollator(object):
    """BatchCollator (data/collate_batch.py:39-56) over RAW samples: a batch is a list of
    (uint8 HWC image — numpy array or torch tensor, host or device —, BoxList | None, id); returns
    (ImageList on the device, tuple of transformed targets, tuple of ids).  Host images are staged through pinned
    memory and copied asynchronously on the current stream; one dd_preprocess_image launch per image writes
    straight into the zero-padded batch tensor."""

    def __init__(self, transform, size_divisible=0):
        self.transform = transform
        self.size_divisible = int(size_divisible)
        self._pinned = {}                       # slot -> pinned staging buffer (grow-only)
        self._copied = {}                       # slot -> event recorded after the H2D copy out of that buffer
        self.h2d_bytes = 0                      # bytes copied host -> device by the last call

    def _to_device(self, slot, img):
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img))
        if img.is_cuda:
            return img
        img = img.contiguous()
        busy = self._copied.get(slot)
        if busy is not None:
            busy.synchronize()                   # the previous batch's DMA out of this staging buffer has finished
        pin = self._pinned.get(slot)
        if pin is None or pin.numel() < img.numel():
            pin = torch.empty(img.numel(), dtype=torch.uint8).pin_memory()
            self._pinned[slot] = pin
        stage = pin[: img.numel()].view(img.shape)
        stage.copy_(img)
        self.h2d_bytes += img.numel()
        dev = stage.to(self.transform.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._copied[slot] = ev
        return dev

    def __call__(self, batch):
        images, targets, ids = [], [], []
        for sample in batch:
            images.append(sample[0])
            targets.append(sample[1] if len(sample) > 1 else None)
            ids.append(sample[2] if len(sample) > 2 else None)
        plans = [self.transform.plan((im.shape[1], im.shape[0])) for im in images]
        images, new_targets = self.collate_planned(images, targets, plans)
        return images, new_targets, tuple(ids)

    def collate_planned(self, images, targets, plans, slot_base=0):
        """One padded device batch from raw images whose (output size, flip) decisions are already made."""
        hp = max(p[0][0] for p in plans)
        wp = max(p[0][1] for p in plans)
        if self.size_divisible > 0:
            hp = int(math.ceil(hp / self.size_divisible) * self.size_divisible)
            wp = int(math.ceil(wp / self.size_divisible) * self.size_divisible)
        out = torch.empty((len(images), 3, hp, wp), dtype=torch.float32, device=self.transform.device)
        if slot_base == 0:
            self.h2d_bytes = 0
        new_targets = []
        for i, (im, tg, (out_hw, flip)) in enumerate(zip(images, targets, plans)):
            self.transform.run(self._to_device(slot_base + i, im), out_hw, flip, out[i])
            new_targets.append(self.transform.transform_target(tg, out_hw, flip))
        return ImageList(out, [torch.Size(p[0]) for p in plans]), tuple(new_targets)


class DeviceBatchCollatorTriplet(DeviceBatchCollator):
    """BatchCollator_triplet (data/collate_batch.py:14-36): samples are 9-tuples (image, target, image_p, target_p,
    image_n, target_n, idx1, idx2, idx3) — source, target-domain and auxiliary-domain views; three padded device
    batches come back in the reference's tuple order.  The reference transforms the three images of a sample one after
    the other inside the dataset, so the random draws are made sample by sample in (image, image_p, image_n) order."""

    def plan_batch(self, batch):
        return [[self.transform.plan((s[j].shape[1], s[j].shape[0])) for j in (0, 2, 4)] for s in batch]

    def __call__(self, batch):
        plans = self.plan_batch(batch)
        n = len(batch)
        out = []
        for v, j in enumerate((0, 2, 4)):
            imgs, tgs = self.collate_planned([s[j] for s in batch], [s[j + 1] for s in batch],
                                             [plans[i][v] for i in range(n)], slot_base=v * n)
            out.extend([imgs, tgs])
        return tuple(out) + tuple(tuple(s[k] for s in batch) for k in (6, 7, 8))
