"""ImageList: a zero-padded image batch tensor plus the un-padded (h, w) of each
image (maskrcnn_benchmark/structures/image_list.py:9-91).  ``a + b`` concatenates
two batches along N after zero-padding to the common max size (:36-47) — this is
how the trainer builds the [source..., target...(, aux...)] batch
(maskrcnn_benchmark/engine/trainer.py:215,223).
"""
import math

import torch


class ImageList(object):
    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, *args, **kwargs):
        return ImageList(self.tensors.to(*args, **kwargs), self.image_sizes)

    def __add__(self, other):
        a, b = self.tensors, other.tensors
        h, w = max(a.shape[2], b.shape[2]), max(a.shape[3], b.shape[3])
        out = a.new_zeros((a.shape[0] + b.shape[0], a.shape[1], h, w))
        out[: a.shape[0], :, : a.shape[2], : a.shape[3]].copy_(a)
        out[a.shape[0]:, :, : b.shape[2], : b.shape[3]].copy_(b)
        return ImageList(out, list(self.image_sizes) + list(other.image_sizes))


def to_image_list(tensors, size_divisible=0):
    """Tensor [N,C,H,W] / [C,H,W] / list of [C,H,W] / ImageList -> ImageList
    (image_list.py:49-91); lists are padded to the per-dim max, rounded up to
    ``size_divisible``."""
    if isinstance(tensors, torch.Tensor) and size_divisible > 0:
        tensors = [tensors] if tensors.dim() == 3 else list(tensors)
    if isinstance(tensors, ImageList):
        return tensors
    if isinstance(tensors, torch.Tensor):
        if tensors.dim() == 3:
            tensors = tensors[None]
        assert tensors.dim() == 4
        return ImageList(tensors, [t.shape[-2:] for t in tensors])
    if isinstance(tensors, (tuple, list)):
        c, h, w = (max(s) for s in zip(*[img.shape for img in tensors]))
        if size_divisible > 0:
            h = int(math.ceil(h / size_divisible) * size_divisible)
            w = int(math.ceil(w / size_divisible) * size_divisible)
        batched = tensors[0].new_zeros((len(tensors), c, h, w))
        for img, pad in zip(tensors, batched):
            pad[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
        return ImageList(batched, [im.shape[-2:] for im in tensors])
    raise TypeError("Unsupported type for to_image_list: {}".format(type(tensors)))
