"""Box head: proposal sampling, ROIAlign -> res5 -> predictor, Fast R-CNN losses masked to source ROIs,
and the instance features / domain labels handed to the DA heads.

Mirrors maskrcnn_benchmark/modeling/roi_heads/{roi_heads.py, box_head/{box_head,loss,
roi_box_feature_extractors,roi_box_predictors,inference}.py}; state-dict names
``roi_heads.box.feature_extractor.head.layer4.*`` and ``roi_heads.box.predictor.{cls_score,bbox_pred}.*``.

De-duplication: the reference runs ROIAlign + res5 + predictor a second time on
``subsample_for_da`` proposals that are provably the same boxes in the same order
(box_head/box_head.py:104-110).  Here the features are computed once and used by both consumers;
autograd adds the two gradient contributions exactly as the reference's two passes would.  The two
extra ``randperm`` draws per image of that second pass are still made, to keep the RNG stream aligned.
"""
import torch
from torch import nn

from .. import ops
from ..structures import BoxList, is_source_image
from .backbone import ResNetHead
from ..utils.sections import section
from .sampling import BELOW_LOW_THRESHOLD, BETWEEN_THRESHOLDS, balanced_sample


class Pooler(nn.Module):
    """Single-level Pooler (poolers.py:45-121).  `even_bins` selects the stride-2-aware ROIAlign."""

    def __init__(self, output_size, scales, sampling_ratio):
        super().__init__()
        if len(scales) != 1:
            raise NotImplementedError("multi-level pooling (FPN) is a 'next' row (SURVEY §8f)")
        self.output_size = int(output_size)
        self.scale = float(scales[0])
        self.sampling_ratio = int(sampling_ratio)

    @staticmethod
    def convert_to_roi_format(boxes):
        rows = []
        for i, b in enumerate(boxes):
            rows.append(torch.cat([torch.full((len(b), 1), float(i), dtype=b.bbox.dtype, device=b.bbox.device),
                                   b.bbox], dim=1))
        return torch.cat(rows, dim=0)

    def forward(self, feats, boxes, even_bins):
        rois = self.convert_to_roi_format(boxes)
        return ops.roi_align(feats[0], rois, self.scale, self.output_size, self.sampling_ratio, 2 if even_bins else 1)


class ResNet50Conv5ROIFeatureExtractor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        B = cfg.MODEL.ROI_BOX_HEAD
        self.pooler = Pooler(B.POOLER_RESOLUTION, B.POOLER_SCALES, B.POOLER_SAMPLING_RATIO)
        self.head = ResNetHead(cfg)
        # res5's first block reads only even bins when its 1x1 convs carry the stride
        self.even_bins = bool(cfg.MODEL.RESNETS.STRIDE_IN_1X1 and cfg.MODEL.RESNETS.RES5_DILATION == 1
                              and B.POOLER_RESOLUTION % 2 == 0)

    def forward(self, feats, proposals):
        x = self.pooler(feats, proposals, self.even_bins)
        return self.head(x, self.even_bins)


class FastRCNNPredictor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        num_inputs = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS * 8
        nc = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES
        self.cls_score = nn.Linear(num_inputs, nc)
        self.bbox_pred = nn.Linear(num_inputs, (2 if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG else nc) * 4)
        nn.init.normal_(self.cls_score.weight, mean=0, std=0.01)
        nn.init.constant_(self.cls_score.bias, 0)
        nn.init.normal_(self.bbox_pred.weight, mean=0, std=0.001)
        nn.init.constant_(self.bbox_pred.bias, 0)

    def forward(self, pooled):
        """pooled: [K, 2048] (the 7x7 average is shared with the DA instance head)."""
        return (ops.linear(pooled, self.cls_score.weight, self.cls_score.bias),
                ops.linear(pooled, self.bbox_pred.weight, self.bbox_pred.bias))


class FastRCNNLossComputation(object):
    """box_head/loss.py:16-221."""

    def __init__(self, cfg, rng):
        H = cfg.MODEL.ROI_HEADS
        self.high, self.low = H.FG_IOU_THRESHOLD, H.BG_IOU_THRESHOLD
        self.weights = H.BBOX_REG_WEIGHTS
        self.batch, self.pos_fraction = H.BATCH_SIZE_PER_IMAGE, H.POSITIVE_FRACTION
        if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG:
            raise NotImplementedError("CLS_AGNOSTIC_BBOX_REG is not used by any DA config")
        self.rng = rng

    def prepare_targets(self, proposals, targets, sample_for_da=False):
        labels, regs, domains = [], [], []
        for p, t in zip(proposals, targets):
            src = is_source_image(t)
            gt = t.convert("xyxy").bbox
            m, _ = ops.match(gt, p.bbox, self.high, self.low, False)
            gl = t.get_field("labels")
            sel = m.clamp(min=0) if src else m               # loss.py:47-51 (negative wrap for target images)
            lab = gl[sel].to(torch.int64).clone()
            lab[m == BELOW_LOW_THRESHOLD] = 0
            lab[m == BETWEEN_THRESHOLDS] = -1
            regs.append(ops.box_encode(gt, p.bbox, m, self.weights, wrap_negative=not src))
            domains.append(torch.full_like(lab, src, dtype=torch.bool))
            if not src or sample_for_da:
                lab[:] = 0
            labels.append(lab)
        return labels, regs, domains

    @torch.no_grad()
    def subsample(self, proposals, targets):
        labels, regs, domains = self.prepare_targets(proposals, targets)
        pos_m, neg_m = balanced_sample(labels, self.batch, self.pos_fraction, self.rng)
        out = []
        for p, lab, rg, dm, pm, nm in zip(proposals, labels, regs, domains, pos_m, neg_m):
            idx = torch.nonzero(pm | nm).squeeze(1)
            q = BoxList(p.bbox[idx], p.size, p.mode)
            if p.has_field("objectness"):
                q.add_field("objectness", p.get_field("objectness")[idx])
            q.add_field("labels", lab[idx])
            q.add_field("regression_targets", rg[idx])
            q.add_field("domain_labels", dm[idx])
            out.append(q)
        self._proposals = out
        return out

    @torch.no_grad()
    def subsample_for_da(self, proposals, targets):
        """subsample_for_da (loss.py:132-163) selects every already-sampled proposal in order; only its
        RNG side effect (2 randperm per image) is reproduced."""
        for p in proposals:
            n = len(p)
            self.rng.randperm(0, p.bbox.device)
            self.rng.randperm(n, p.bbox.device)
            if n > self.batch:
                raise AssertionError("subsample_for_da would drop proposals; not reachable from subsample()")
        return proposals

    def __call__(self, class_logits, box_regression):
        props = self._proposals
        labels = torch.cat([p.get_field("labels") for p in props], dim=0)
        regt = torch.cat([p.get_field("regression_targets") for p in props], dim=0)
        dom = torch.cat([p.get_field("domain_labels") for p in props], dim=0)
        mask = dom.to(torch.uint8)
        cls_loss = ops.softmax_ce_mean(class_logits, labels, mask)
        box_loss = ops.box_reg_loss(box_regression, regt, labels, mask)
        return cls_loss, box_loss, dom


class ROIBoxHead(nn.Module):
    def __init__(self, cfg, rng):
        super().__init__()
        if cfg.MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR != "ResNet50Conv5ROIFeatureExtractor":
            raise NotImplementedError("only the C4 res5 box head is on the accelerated path")
        self.feature_extractor = ResNet50Conv5ROIFeatureExtractor(cfg)
        self.predictor = FastRCNNPredictor(cfg)
        self.loss_evaluator = FastRCNNLossComputation(cfg, rng)
        self.cfg = cfg.clone()

    def forward(self, features, proposals, targets=None):
        """Training: returns (x, proposals, losses, da_ins_feas, da_ins_labels) like
        box_head/box_head.py:36-117, plus the shared pooled vector as attribute `last_pooled`."""
        if self.training:
            with section("  box_subsample"):
                proposals = self.loss_evaluator.subsample(proposals, targets)
        segments = self.__dict__.get("segments")
        if segments is not None and self.training:
            from .detector import _BoxBranch
            rois = Pooler.convert_to_roi_format(proposals)
            pooled, class_logits, box_regression = segments.run(
                "box", lambda: _BoxBranch(self.feature_extractor, self.predictor), (features[0], rois))
            x = None          # the [K,7,7,2048] map stays inside the segment; consumers use `pooled`
        else:
            x = self.feature_extractor(features, proposals)
            pooled = ops.avgpool_hw(x)
            class_logits, box_regression = self.predictor(pooled)
        self.last_pooled = pooled
        if not self.training:
            from .inference import box_post_process
            return x, box_post_process(self.cfg, class_logits, box_regression, proposals), {}, x, None
        loss_classifier, loss_box_reg, dom = self.loss_evaluator(class_logits, box_regression)
        self.loss_evaluator.subsample_for_da(proposals, targets)
        self.last = dict(class_logits=class_logits, box_regression=box_regression)
        return x, proposals, dict(loss_classifier=loss_classifier, loss_box_reg=loss_box_reg), x, dom


class CombinedROIHeads(nn.ModuleDict):
    def __init__(self, cfg, heads):
        super().__init__(heads)
        self.cfg = cfg.clone()

    def forward(self, features, proposals, targets=None):
        return self.box(features, proposals, targets)


def build_roi_heads(cfg, rng):
    if cfg.MODEL.RPN_ONLY:
        return []
    if cfg.MODEL.MASK_ON or cfg.MODEL.KEYPOINT_ON:
        raise NotImplementedError("mask / keypoint heads are outside the DA path")
    return CombinedROIHeads(cfg, [("box", ROIBoxHead(cfg, rng))])
