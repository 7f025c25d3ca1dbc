"""BoxList: a box tensor + per-box fields + the image size — the boundary type the
reference passes between data loader, model and losses
(maskrcnn_benchmark/structures/bounding_box.py:9-266).  API-compatible subset:
bbox/size/mode/extra_fields, add_field/get_field/has_field/fields, convert,
clip_to_image, area, indexing, to(), copy_with_fields, len().  Pixel boxes use the
legacy inclusive "+1" width convention (bounding_box.py:67-86,215-230).
"""
import torch

FLIP_LEFT_RIGHT = 0
FLIP_TOP_BOTTOM = 1


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2:
            raise ValueError("bbox should have 2 dimensions, got {}".format(bbox.ndimension()))
        if bbox.size(-1) != 4:
            raise ValueError("last dimension of bbox should have a size of 4, got {}".format(bbox.size(-1)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (image_width, image_height)
        self.mode = mode
        self.extra_fields = {}

    # fields -----------------------------------------------------------------
    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, other):
        for k, v in other.extra_fields.items():
            self.extra_fields[k] = v

    # geometry ---------------------------------------------------------------
    def _split_into_xyxy(self):
        if self.mode == "xyxy":
            return self.bbox.split(1, dim=-1)
        x, y, w, h = self.bbox.split(1, dim=-1)
        return x, y, x + (w - 1).clamp(min=0), y + (h - 1).clamp(min=0)

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        x1, y1, x2, y2 = self._split_into_xyxy()
        if mode == "xyxy":
            box = torch.cat((x1, y1, x2, y2), dim=-1)
        else:
            box = torch.cat((x1, y1, x2 - x1 + 1, y2 - y1 + 1), dim=-1)
        out = BoxList(box, self.size, mode=mode)
        out._copy_extra_fields(self)
        return out

    def clip_to_image(self, remove_empty=True):
        w, h = self.size
        self.bbox[:, 0].clamp_(min=0, max=w - 1)
        self.bbox[:, 1].clamp_(min=0, max=h - 1)
        self.bbox[:, 2].clamp_(min=0, max=w - 1)
        self.bbox[:, 3].clamp_(min=0, max=h - 1)
        if remove_empty:
            b = self.bbox
            return self[(b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])]
        return self

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    def transpose(self, method):
        if method not in (FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM):
            raise NotImplementedError("Only FLIP_LEFT_RIGHT and FLIP_TOP_BOTTOM implemented")
        w, h = self.size
        x1, y1, x2, y2 = self._split_into_xyxy()
        if method == FLIP_LEFT_RIGHT:
            x1, x2 = w - x2 - 1, w - x1 - 1
        else:
            y1, y2 = h - y2, h - y1
        out = BoxList(torch.cat((x1, y1, x2, y2), dim=-1), self.size, mode="xyxy")
        for k, v in self.extra_fields.items():
            out.add_field(k, v if isinstance(v, torch.Tensor) else v.transpose(method))
        return out.convert(self.mode)

    def resize(self, size, *args, **kwargs):
        rw, rh = (float(s) / float(o) for s, o in zip(size, self.size))
        if rw == rh:                     # bounding_box.py:99-108: equal ratios scale the box in its own mode
            out = BoxList(self.bbox * rw, size, mode=self.mode)
            for k, v in self.extra_fields.items():
                out.add_field(k, v if isinstance(v, torch.Tensor) else v.resize(size, *args, **kwargs))
            return out
        x1, y1, x2, y2 = self._split_into_xyxy()
        out = BoxList(torch.cat((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=-1), size, mode="xyxy")
        for k, v in self.extra_fields.items():
            out.add_field(k, v if isinstance(v, torch.Tensor) else v.resize(size, *args, **kwargs))
        return out.convert(self.mode)

    # container behaviour ----------------------------------------------------
    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def copy_with_fields(self, fields, skip_missing=False):
        out = BoxList(self.bbox, self.size, self.mode)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for field in fields:
            if self.has_field(field):
                out.add_field(field, self.get_field(field))
            elif not skip_missing:
                raise KeyError("Field '{}' not found in {}".format(field, self))
        return out

    def __repr__(self):
        return "BoxList(num_boxes={}, image_width={}, image_height={}, mode={})".format(
            len(self), self.size[0], self.size[1], self.mode)


def is_source_image(target):
    """True when the image's boxes carry is_source=True (reference: `is_source.any()`, e.g. rpn/loss.py:63-67).
    The answer is cached on the BoxList so the device->host read happens once per image, not once per use."""
    flag = getattr(target, "_is_source_image", None)
    if flag is None:
        flag = bool(target.get_field("is_source").any())
        target._is_source_image = flag
    return flag


def cache_source_flags(targets):
    """One device->host transfer for the whole batch instead of one sync per image per consumer."""
    import torch
    todo = [t for t in targets if getattr(t, "_is_source_image", None) is None]
    if not todo:
        return
    flags = torch.stack([t.get_field("is_source").any() for t in todo]).tolist()
    for t, f in zip(todo, flags):
        t._is_source_image = bool(f)
