"""Seeded synthetic inputs and weights for parity tests and bench.py.

Images: U[0,255) minus the BGR pixel mean (reference defaults.py:52-56,
data/transforms/transforms.py:88-98).  Targets: M boxes per image with x1~U[0,0.93W),
y1~U[0,0.88H), w~U[16,0.2W), h~U[16,0.3H) clipped to the image, labels~U{1..C-1};
image 0 is the source domain, the others target/aux ([source, target(, aux)] order,
reference engine/trainer.py:215,223).  Weights: per-tensor generators keyed by the
state-dict NAME, so the values do not depend on module construction order and the
same dict can be loaded into the reference model, the oracle and this package.
"""
import zlib

import torch

PIXEL_MEAN = (102.9801, 115.9465, 122.7717)


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) & 0x7FFFFFFF)
    return g


def make_batch(n_images, height, width, num_classes=9, boxes_per_image=20, seed=1029):
    """Returns (images f32 [N,3,H,W], targets list of dict(boxes, labels, is_source))."""
    g = _gen(seed)
    images = torch.rand((n_images, 3, height, width), generator=g) * 255.0
    images -= torch.tensor(PIXEL_MEAN).view(1, 3, 1, 1)
    targets = []
    for i in range(n_images):
        m = boxes_per_image
        x1 = torch.rand(m, generator=g) * (0.93 * width)
        y1 = torch.rand(m, generator=g) * (0.88 * height)
        w = 16 + torch.rand(m, generator=g) * max(0.2 * width - 16, 1.0)
        h = 16 + torch.rand(m, generator=g) * max(0.3 * height - 16, 1.0)
        x2 = torch.minimum(x1 + w, torch.tensor(width - 1.0))
        y2 = torch.minimum(y1 + h, torch.tensor(height - 1.0))
        boxes = torch.stack([x1, y1, x2, y2], dim=1).floor()
        labels = torch.randint(1, num_classes, (m,), generator=g, dtype=torch.int64)
        targets.append(dict(boxes=boxes, labels=labels, is_source=(i == 0)))
    return images, targets


def make_state_dict(shapes, seed=100, random_bn=True):
    """shapes: {name: shape}.  Conv/FC weights follow the reference initialisers'
    distributions (kaiming_uniform(a=1) for backbone/res5 convs, resnet.py:254,291,328;
    normal(std) heads, rpn.py:35-37, roi_box_predictors.py:22-26, da_heads.py:28-30,54-58);
    biases get a small normal instead of 0 so that bias paths are exercised.  FrozenBN
    buffers are identity when random_bn=False (batch_norm.py:14-17); otherwise random
    affine values chosen so activations stay O(1..10) through the residual trunk."""
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        g = _gen(seed * 1000003 + zlib.crc32(name.encode()))
        if ".cell_anchors." in name:
            continue
        leaf = name.rsplit(".", 1)[-1]
        is_bn = ".bn" in name or ".downsample.1." in name
        if is_bn:
            c = shape[0]
            if not random_bn:
                v = {"weight": torch.ones(c), "bias": torch.zeros(c),
                     "running_mean": torch.zeros(c), "running_var": torch.ones(c)}[leaf]
            elif leaf == "weight":
                base = 0.02 if "stem" in name else (0.35 if (".bn3" in name or ".downsample.1." in name) else 1.0)
                v = base * (0.75 + 0.5 * torch.rand(c, generator=g))
            elif leaf == "bias":
                v = 0.1 * torch.randn(c, generator=g)
            elif leaf == "running_mean":
                v = 0.1 * torch.randn(c, generator=g)
            else:
                v = 0.5 + torch.rand(c, generator=g)
            sd[name] = v
            continue
        if leaf == "bias":
            sd[name] = 0.01 * torch.randn(shape, generator=g)
            continue
        if name.startswith("rpn.head"):
            std = 0.01
        elif "cls_score" in name:
            std = 0.01
        elif "bbox_pred" in name:
            std = 0.001
        elif "imghead" in name:
            std = 0.001
        elif "fc3_da" in name:
            std = 0.05
        elif "inshead" in name:
            std = 0.01
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            bound = (3.0 / fan_in) ** 0.5
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            continue
        sd[name] = std * torch.randn(shape, generator=g)
    return sd
