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
from ..structures import cache_source_flags, is_source_image, to_image_list
from ..utils.random_source import RandomSource
from ..utils.sections import section
from .backbone import build_backbone
from .da_heads import build_da_heads, build_da_heads_triplet
from .roi_heads import build_roi_heads
from .rpn import build_rpn


class _Trunk(nn.Module):
    """Static-shape segment 1: backbone + RPN head.  Not registered in the model tree (shares its modules)."""

    def __init__(self, backbone, rpn_head):
        super().__init__()
        self.backbone, self.rpn_head = backbone, rpn_head

    def forward(self, x):
        feat = self.backbone(x)[0]
        logits, deltas = self.rpn_head(feat)
        return feat, logits, deltas


class _BoxBranch(nn.Module):
    """Static-shape segment 2 (for a fixed number of ROIs): ROIAlign -> res5 -> 7x7 average -> predictor."""

    def __init__(self, feature_extractor, predictor):
        super().__init__()
        self.fe, self.predictor = feature_extractor, predictor

    def forward(self, feat, rois):
        x = ops.roi_align(feat, rois, self.fe.pooler.scale, self.fe.pooler.output_size, self.fe.pooler.sampling_ratio,
                          2 if self.fe.even_bins else 1)
        pooled = self.fe.head(x, self.fe.even_bins, pooled=True)
        cls, box = self.predictor(pooled)
        return pooled, cls, box


class SegmentRunner(object):
    """Runs static-shape segments either eagerly or as CUDA graphs (torch.cuda.make_graphed_callables captures
    the forward AND backward kernel sequences of a segment once per input-shape signature; replay removes the
    per-launch Python/ctypes/autograd overhead, which is ~3x the GPU time of the step otherwise)."""

    def __init__(self):
        self.enabled = False
        self.cache = {}

    def run(self, name, factory, args):
        if not self.enabled or not torch.is_grad_enabled():
            return factory()(*args)
        key = (name,) + tuple((tuple(a.shape), a.requires_grad) for a in args)
        fn = self.cache.get(key)
        if fn is None:
            sample = tuple(a.detach().clone().requires_grad_(a.requires_grad) for a in args)
            torch.cuda.synchronize()
            fn = torch.cuda.make_graphed_callables(factory(), sample)
            self.cache[key] = fn
        return fn(*args)


class GeneralizedRCNN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.__dict__["segments"] = SegmentRunner()
        self.rng = RandomSource()
        self.backbone = build_backbone(cfg)
        self.rpn = build_rpn(cfg, self.rng)
        self.roi_heads = build_roi_heads(cfg, self.rng)
        self.da_heads = build_da_heads(cfg, self.rng)
        self.triplet_use = cfg.MODEL.DA_HEADS.TRIPLET_USE
        self.da_heads_triplet = build_da_heads_triplet(cfg, self.rng) if self.triplet_use else False
        self.Aligned = cfg.MODEL.DA_HEADS.ALIGNMENT
        self.size_divisible = cfg.DATALOADER.SIZE_DIVISIBILITY
        # FPN: multi-level features, host-driven control flow.  FPN + DA follows the intent of
        # da_heads_fpn.py (DomainAdaptationModuleFPN, parity unpinned: the reference's own combination cannot run,
        # SURVEY §9.9); the triplet module has no FPN counterpart in the reference at all.
        self.fpn = cfg.MODEL.BACKBONE.CONV_BODY.endswith("-FPN")
        if self.fpn and self.da_heads_triplet:
            raise NotImplementedError("the auxiliary-domain triplet module on an FPN backbone: the reference has no "
                                      "FPN variant of DomainAdaptationModule_triplet; set MODEL.DA_HEADS.TRIPLET_USE False")
        self.static_shapes = not self.fpn
        # run the backward pass of the RPN losses during the forward pass, beside the proposal chain (rpn.py); the
        # caller must zero the gradients BEFORE the forward pass and sum the losses with unit weights: set by
        # FlatSGDTrainer, off under any other driver
        self.early_backward = False
        self.__dict__["_meta_cache"] = {}

    def enable_static_shapes(self, flag=True):
        """Training without host reads (plain DA / no-DA modes): proposals and sampled ROIs live in fixed-capacity
        device buffers with device-side counts, so a whole step can be captured into one CUDA graph."""
        self.static_shapes = bool(flag)

    def _batch_meta(self, targets):
        """Per-batch-signature device constants (GT offsets, which images take GT proposals, per-image domain
        labels); built once per signature, outside any graph capture."""
        key = tuple((len(t), bool(is_source_image(t))) for t in targets)
        meta = self._meta_cache.get(key)
        if meta is None:
            dev = targets[0].bbox.device
            offs = [0]
            for n, _ in key:
                offs.append(offs[-1] + n)
            meta = dict(gt_offsets=torch.tensor(offs, dtype=torch.int32, device=dev),
                        append_gt=torch.tensor([1 if s else 0 for _, s in key], dtype=torch.uint8, device=dev),
                        seg=torch.tensor([1 if s else 0 for _, s in key], dtype=torch.uint8, device=dev),
                        src_index=torch.tensor([i for i, (_, s) in enumerate(key) if s], dtype=torch.int32, device=dev),
                        max_gt=max([n for n, s in key if s] + [0]))
            self._meta_cache[key] = meta
        meta = dict(meta)
        meta["gt_cat"] = torch.cat([t.convert("xyxy").bbox.to(torch.float32) for t in targets], dim=0)
        counts = [getattr(t, "_gt_count_dev", None) for t in targets]
        if all(c is not None for c in counts):       # GT padded to a capacity: the live row counts are on the device
            meta["gt_counts"] = torch.cat(counts)
        return meta

    def _forward_static(self, images, targets, feat, head_out):
        """head_out: the RPN head's (logits, deltas), or None in early-backward mode (rpn.py::_forward_static_early)."""
        meta = self._batch_meta(targets)
        features = [feat]
        early = head_out is None
        early_img, after_head = [None], None
        if early and self.da_heads and not self.da_heads_triplet:
            def after_head(ready):
                early_img[0] = self.da_heads.early_image_loss(feat, targets, meta["seg"], after=ready)
                return early_img[0][0] if early_img[0] is not None else None
        props, proposal_losses, pending = self.rpn.forward_static(images, features, targets, head_out, meta,
                                                                  early_backward=early, after_head=after_head)
        losses = self._forward_static_heads(features, targets, props, proposal_losses, meta, early_img[0])
        if pending is not None:
            self.rpn.finish_early_backward(losses, pending)
        return losses

    def _forward_static_fpn(self, images, targets, features):
        """FPN training without host reads (enable_static_shapes(True); what FlatSGDTrainer's step graph captures):
        fixed-capacity proposals over the levels (rpn.py::proposals_fpn_static), fixed ROI slots through the multi-level
        pooler and the MLP head, per-level DA heads on all slots with level masks."""
        meta = self._batch_meta(targets)
        props, proposal_losses = self.rpn.forward_fpn_static(images, features, targets, meta)
        box = self.roi_heads.box
        detector_losses, pooled, dom, row_valid = box.forward_static(features, props, targets)
        losses = {}
        losses.update(detector_losses)
        losses.update(proposal_losses)
        if self.da_heads:
            losses.update(self.da_heads(features, pooled, dom, box.loss_evaluator.batch, targets,
                                        box.feature_extractor.pooler.last_levels, row_valid=row_valid, seg=meta["seg"]))
        return losses

    def _forward_static_heads(self, features, targets, props, proposal_losses, meta, early_img=None):
        feat = features[0]
        losses = {}
        box = self.roi_heads.box
        B = box.loss_evaluator.batch
        if self.da_heads_triplet:
            # generalized_rcnn.py:88-122 — exactly one image per domain per rank: [source, target, auxiliary]
            ori_features, ori_targets = [feat[0:2]], targets[0:2]
            detector_losses, pooled, dom, row_valid = box.forward_static(ori_features, props.slice(0, 2), ori_targets)
            pooled_set, set_valid = [0, 0, 0], None
            if self.Aligned:                             # :109-114 — all three with the TARGET image's proposals
                pooled_set, set_valid = [], []
                p1 = props.slice(1, 2)
                for i in range(3):
                    _, pooled_i, _, valid_i = box.forward_static([feat[i:i + 1]], p1, [targets[i]])
                    pooled_set.append(pooled_i)
                    set_valid.append(valid_i)
            img_set = [feat[0:1], feat[1:2], feat[2:3]]
            da_losses = self.da_heads_triplet(ori_features, pooled, dom, B, pooled_set, img_set, ori_targets,
                                              row_valid=row_valid, seg=meta["seg"][0:2], set_valid=set_valid)
            losses.update(detector_losses)
            losses.update(proposal_losses)
            losses.update(da_losses)
            return losses
        detector_losses, pooled, dom, row_valid = box.forward_static(features, props, targets)
        losses.update(detector_losses)
        losses.update(proposal_losses)
        if self.da_heads:
            losses.update(self.da_heads(features, pooled, dom, B, targets, row_valid=row_valid, seg=meta["seg"],
                                        early_img=early_img))
        return losses

    def enable_cuda_graphs(self, flag=True):
        """Replay the static-shape segments (backbone + RPN head; box branch per ROI count) as CUDA graphs."""
        self.segments.enabled = bool(flag)
        if self.roi_heads:
            self.roi_heads.box.__dict__["segments"] = self.segments

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
        if targets is not None:
            cache_source_flags(targets)
        if self.fpn:
            x = ops._chk(images.tensors, name="images")
            features = self.backbone(x)
            if self.training and self.static_shapes and self.roi_heads:
                return self._forward_static_fpn(images, targets, features)
            proposals, proposal_losses = self.rpn(images, features, targets)
        else:
            static = self.training and self.static_shapes and bool(self.roi_heads)
            with section("trunk_fwd"):
                x = ops._chk(images.tensors, name="images")        # NCHW; the stem consumes it directly
                if static and self.early_backward and torch.is_grad_enabled() and not self.segments.enabled:
                    return self._forward_static(images, targets, self.backbone(x)[0], None)
                feat, logits, deltas = self.segments.run("trunk", lambda: _Trunk(self.backbone, self.rpn.head), (x,))
            if static:
                return self._forward_static(images, targets, feat, (logits, deltas))
            features = [feat]
            with section("rpn_proposals_and_loss"):
                proposals, proposal_losses = self.rpn(images, features, targets, head_out=(logits, deltas))
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
                                if is_source_image(t))
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
                    with section("box_head_fwd"):
                        _, _, detector_losses, _, dom = self.roi_heads(features, proposals, targets)
                    n_src = sum(len(p) for p, t in zip(box.loss_evaluator._proposals, targets)
                                if is_source_image(t))
                    with section("da_heads_fwd"):
                        if self.fpn:      # per-level heads; the ROIs' pyramid levels come from the multi-level pooler
                            da_losses = self.da_heads(features, box.last_pooled, dom, n_src, targets,
                                                      box.feature_extractor.pooler.last_levels)
                        else:
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
