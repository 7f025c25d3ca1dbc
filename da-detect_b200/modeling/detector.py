"""GeneralizedRCNN — the drop-in model boundary.

Same constructor contract, forward signature, loss-dict keys and state-dict names as
maskrcnn_benchmark/modeling/detector/{detectors.py:8-10, generalized_rcnn.py:37-156}, so the
reference's tools/train_net_triplet.py / engine/trainer.py drive it unchanged:

    model = build_detection_model(cfg); model.to("cuda")
    loss_dict = model(images, targets); sum(loss_dict.values()).backward()

`images`: ImageList | list[Tensor[3,H,W]] | Tensor[N,3,H,W] (NCHW, mean-subtracted BGR — the reference's
input convention); `targets`: list[BoxList] with fields `labels` and `is_source`, ordered
[source..., target...(, aux...)].  Internally everything runs NHWC through libdadetect_b200.so.
"""
import torch
from torch import nn

from .. import ops
from ..structures import to_image_list
from ..utils.random_source import RandomSource
from .backbone import build_backbone
from .da_heads import build_da_heads, build_da_heads_triplet
from .roi_heads import build_roi_heads
from .rpn import build_rpn


class GeneralizedRCNN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.rng = RandomSource()
        self.backbone = build_backbone(cfg)
        self.rpn = build_rpn(cfg, self.rng)
        self.roi_heads = build_roi_heads(cfg, self.rng)
        self.da_heads = build_da_heads(cfg, self.rng)
        self.triplet_use = cfg.MODEL.DA_HEADS.TRIPLET_USE
        self.da_heads_triplet = build_da_heads_triplet(cfg, self.rng) if self.triplet_use else False
        self.Aligned = cfg.MODEL.DA_HEADS.ALIGNMENT
        self.size_divisible = cfg.DATALOADER.SIZE_DIVISIBILITY

    def set_random_source(self, rng):
        """Swap the source of randperm/dropout draws (parity tests replay oracle draws)."""
        self.rng = rng
        for m in self.modules():
            if hasattr(m, "rng"):
                m.rng = rng
            if hasattr(m, "loss_evaluator") and hasattr(m.loss_evaluator, "rng"):
                m.loss_evaluator.rng = rng

    def forward(self, images, targets=None):
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        images = to_image_list(images)
        x = ops.nchw_to_nhwc(images.tensors)
        features = self.backbone(x)
        proposals, proposal_losses = self.rpn(images, features, targets)
        da_losses = {}
        if self.roi_heads:
            if self.training:
                feat = features[0]
                if self.da_heads_triplet:
                    # generalized_rcnn.py:88-122 — exactly one image per domain per rank
                    ori_features = [feat[0:2]]
                    ori_targets = targets[0:2]
                    box = self.roi_heads.box
                    _, _, detector_losses, _, dom = self.roi_heads(ori_features, proposals[0:2], ori_targets)
                    pooled = box.last_pooled
                    n_src = sum(len(p) for p, t in zip(box.loss_evaluator._proposals, ori_targets)
                                if bool(t.get_field("is_source").any()))
                    pooled_set = [0, 0, 0]
                    if self.Aligned:                         # :109-114 — all three with the TARGET image's proposals
                        pooled_set = []
                        for i in range(3):
                            self.roi_heads([feat[i:i + 1]], [proposals[1]], [targets[i]])
                            pooled_set.append(box.last_pooled)
                    img_set = [feat[0:1], feat[1:2], feat[2:3]]
                    da_losses = self.da_heads_triplet(ori_features, pooled, dom, n_src, pooled_set, img_set,
                                                      ori_targets)
                elif self.da_heads:
                    box = self.roi_heads.box
                    _, _, detector_losses, _, dom = self.roi_heads(features, proposals, targets)
                    n_src = sum(len(p) for p, t in zip(box.loss_evaluator._proposals, targets)
                                if bool(t.get_field("is_source").any()))
                    da_losses = self.da_heads(features, box.last_pooled, dom, n_src, targets)
                else:
                    # The reference leaves `detector_losses` unbound here; plain Faster R-CNN
                    # training is the obvious intent.
                    _, _, detector_losses, _, _ = self.roi_heads(features, proposals, targets)
            else:
                _, result, detector_losses, _, _ = self.roi_heads(features, proposals, targets)
        else:
            result = proposals
            detector_losses = {}
        if self.training:
            losses = {}
            losses.update(detector_losses)
            losses.update(proposal_losses)
            losses.update(da_losses)
            return losses
        return result


_META_ARCHITECTURES = {"GeneralizedRCNN": GeneralizedRCNN}


def build_detection_model(cfg):
    return _META_ARCHITECTURES[cfg.MODEL.META_ARCHITECTURE](cfg)
