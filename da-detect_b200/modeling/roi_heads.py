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


class MultiLevelPooler(nn.Module):
    """Pooler over FPN levels (poolers.py:45-121): LevelMapper + one ROIAlign per level, written straight into the
    rows of one output tensor by ops.roi_align_levels (no nonzero / index_put round trips)."""

    def __init__(self, output_size, scales, sampling_ratio):
        super().__init__()
        self.output_size = int(output_size)
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = int(sampling_ratio)

    def forward(self, feats, boxes):
        rois = Pooler.convert_to_roi_format(boxes)
        out, self.last_levels = ops.roi_align_levels(list(feats[:len(self.scales)]), rois, self.scales,
                                                     self.output_size, self.sampling_ratio)
        return out


class FPN2MLPFeatureExtractor(nn.Module):
    """roi_box_feature_extractors.py:48-79: multi-level 7x7 ROIAlign -> fc6 -> ReLU -> fc7 -> ReLU.  The reference
    flattens its NCHW ROI map as (C, h, w); ours is NHWC, so fc6's columns are re-ordered to (h, w, C) on the fly —
    the parameter keeps the reference's [1024, C*49] shape and column order (checkpoint compatible)."""

    def __init__(self, cfg):
        super().__init__()
        B = cfg.MODEL.ROI_BOX_HEAD
        if B.USE_GN:
            raise NotImplementedError("ROI_BOX_HEAD.USE_GN is not used by any DA config")
        self.pooler = MultiLevelPooler(B.POOLER_RESOLUTION, B.POOLER_SCALES, B.POOLER_SAMPLING_RATIO)
        self.channels, self.resolution = cfg.MODEL.BACKBONE.OUT_CHANNELS, B.POOLER_RESOLUTION
        self.fc6 = nn.Linear(self.channels * self.resolution ** 2, B.MLP_HEAD_DIM)
        self.fc7 = nn.Linear(B.MLP_HEAD_DIM, B.MLP_HEAD_DIM)
        for fc in (self.fc6, self.fc7):                      # make_fc (make_layers.py:83-94)
            nn.init.kaiming_uniform_(fc.weight, a=1)
            nn.init.constant_(fc.bias, 0)
        self.even_bins = False

    def forward(self, feats, proposals):
        return self._head(self.pooler(feats, proposals))     # [K, r, r, C] -> [K, MLP_HEAD_DIM]

    def forward_rois(self, feats, rois):
        """The same on a fixed-capacity ROI tensor [K, 5] (sync-free training path)."""
        p = self.pooler
        x, p.last_levels = ops.roi_align_levels(list(feats[:len(p.scales)]), rois, p.scales, p.output_size,
                                                p.sampling_ratio)
        return self._head(x)

    def _head(self, x):
        k, r, c = x.shape[0], self.resolution, self.channels
        w6 = self.fc6.weight.view(-1, c, r * r).permute(0, 2, 1).reshape(-1, r * r * c)
        x = ops.linear(x.reshape(k, r * r * c), w6, self.fc6.bias, relu=True)
        return ops.linear(x, self.fc7.weight, self.fc7.bias, relu=True)


class FPNPredictor(nn.Module):
    """roi_box_predictors.py:37-57."""

    def __init__(self, cfg):
        super().__init__()
        rep, nc = cfg.MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM, cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES
        self.cls_score = nn.Linear(rep, nc)
        self.bbox_pred = nn.Linear(rep, (2 if cfg.MODEL.CLS_AGNOSTIC_BBOX_REG else nc) * 4)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in (self.cls_score, self.bbox_pred):
            nn.init.constant_(l.bias, 0)

    def forward(self, x):
        cls, box = ops.fused_heads(x, [self.cls_score.weight, self.bbox_pred.weight],
                                   [self.cls_score.bias, self.bbox_pred.bias])
        return cls, box


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
        cls, box = ops.fused_heads(pooled, [self.cls_score.weight, self.bbox_pred.weight],
                                   [self.cls_score.bias, self.bbox_pred.bias])       # 9 + 36 -> one 64-column GEMM
        return cls, box


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
        self._static_const = {}

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

    @torch.no_grad()
    def subsample_static(self, props, targets):
        """subsample (loss.py:100-130) on a ProposalBatch with the sampler on the device: every image contributes
        exactly `batch` ROI slots; `valid` marks the slots that hold a sampled proposal (all of them unless an
        image has fewer than `batch` labelled candidates).  Returns the flat per-ROI tensors of the batch."""
        boxes, nprop = props.boxes, props.count
        n_img, cap = boxes.shape[0], boxes.shape[1]
        B = self.batch
        dev = boxes.device
        # per-signature device constants (GT offsets, domain flags): built once, outside any graph capture
        key = tuple((len(t), bool(is_source_image(t))) for t in targets)
        const = self._static_const.get(key)
        if const is None:
            offs = [0]
            for n, _ in key:
                offs.append(offs[-1] + n)
            const = (torch.tensor(offs, dtype=torch.int32, device=dev),
                     torch.tensor([1 if s else 0 for _, s in key], dtype=torch.uint8, device=dev))
            self._static_const[key] = const
        gt_offsets, src_flags = const
        gts = [t.convert("xyxy").bbox for t in targets]
        counts_dev = [getattr(t, "_gt_count_dev", None) for t in targets]
        lab_all = torch.empty((n_img, cap), dtype=torch.int32, device=dev)
        ms, keys = [], []
        for i, t in enumerate(targets):                      # three launches per image: Matcher, labels, random keys
            m, _ = ops.match(gts[i], boxes[i], self.high, self.low, False, m_dev=counts_dev[i])
            ops.roi_labels(m, t.get_field("labels"), is_source_image(t), nprop[i:i + 1], out=lab_all[i])
            ms.append(m)
            keys.append(self.rng.sample_keys(lab_all[i]))
        one = n_img == 1
        m_all = ms[0].view(1, -1) if one else torch.stack(ms)
        key_all = keys[0].view(1, -1) if one else torch.stack(keys)
        sel, cnt = ops.balanced_sample(lab_all, nprop, key_all, B, int(B * self.pos_fraction))
        gt_cat = gts[0] if one else torch.cat(gts, dim=0)
        gt_counts = None
        if all(c is not None for c in counts_dev):           # GT padded to a capacity: live row counts on the device
            gt_counts = counts_dev[0] if one else torch.cat(counts_dev)
        st = ops.roi_gather_sampled(boxes, props.objectness, sel, cnt, lab_all, m_all, gt_cat, gt_offsets, gt_counts,
                                    src_flags, self.weights)
        st.update(counts=cnt[:, 1], sizes=props.sizes)
        self._static = st
        self.rng.consume_da_draws(st["counts"])
        return st

    def static_proposals(self):
        """The sampled proposals of the last static step as list[BoxList] (host reads; for tests / inspection)."""
        st = self._static
        B = self.batch
        out = []
        for i, c in enumerate(st["counts"].tolist()):
            sl = slice(i * B, i * B + c)
            q = BoxList(st["rois"][sl, 1:], st["sizes"][i], mode="xyxy")
            for f in ("objectness", "labels", "regression_targets", "domain_labels"):
                q.add_field(f, st[f][sl])
            out.append(q)
        return out

    def loss_static(self, class_logits, box_regression):
        st = self._static
        mask = (st["valid"] & st["domain_labels"]).to(torch.uint8)
        cls_loss = ops.softmax_ce_mean(class_logits, st["labels"], mask)
        box_loss = ops.box_reg_loss(box_regression, st["regression_targets"], st["labels"], mask)
        return cls_loss, box_loss

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
        fe, pred = cfg.MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR, cfg.MODEL.ROI_BOX_HEAD.PREDICTOR
        self.mlp_head = fe == "FPN2MLPFeatureExtractor"
        if self.mlp_head and pred == "FPNPredictor":
            self.feature_extractor = FPN2MLPFeatureExtractor(cfg)
            self.predictor = FPNPredictor(cfg)
        elif fe == "ResNet50Conv5ROIFeatureExtractor" and pred == "FastRCNNPredictor":
            self.feature_extractor = ResNet50Conv5ROIFeatureExtractor(cfg)
            self.predictor = FastRCNNPredictor(cfg)
        else:
            raise NotImplementedError("box head {} + {} is outside the accelerated path".format(fe, pred))
        self.loss_evaluator = FastRCNNLossComputation(cfg, rng)
        self.cfg = cfg.clone()

    def forward_static(self, features, props, targets):
        """Training on a ProposalBatch without host reads: returns (losses, pooled [K,2048], domain labels [K],
        row_valid uint8 [K]) with K = images x BATCH_SIZE_PER_IMAGE."""
        st = self.loss_evaluator.subsample_static(props, targets)
        segments = self.__dict__.get("segments")
        if self.mlp_head:                                     # FPN: multi-level pooling + MLP head on the ROI slots
            pooled = self.feature_extractor.forward_rois(features, st["rois"])
            class_logits, box_regression = self.predictor(pooled)
        elif segments is not None:
            from .detector import _BoxBranch
            pooled, class_logits, box_regression = segments.run(
                "box", lambda: _BoxBranch(self.feature_extractor, self.predictor), (features[0], st["rois"]))
        else:
            fe = self.feature_extractor
            x = ops.roi_align(features[0], st["rois"], fe.pooler.scale, fe.pooler.output_size,
                              fe.pooler.sampling_ratio, 2 if fe.even_bins else 1)
            pooled = fe.head(x, fe.even_bins, pooled=True)
            class_logits, box_regression = self.predictor(pooled)
        loss_classifier, loss_box_reg = self.loss_evaluator.loss_static(class_logits, box_regression)
        # detached: a kept reference into the autograd graph would pin its AccumulateGrad nodes across steps
        self.last = dict(class_logits=class_logits.detach(), box_regression=box_regression.detach())
        return (dict(loss_classifier=loss_classifier, loss_box_reg=loss_box_reg), pooled, st["domain_labels"],
                st["valid"].to(torch.uint8))

    def forward(self, features, proposals, targets=None):
        """Training: returns (x, proposals, losses, da_ins_feas, da_ins_labels) like
        box_head/box_head.py:36-117, plus the shared pooled vector as attribute `last_pooled`."""
        if self.training:
            with section("  box_subsample"):
                proposals = self.loss_evaluator.subsample(proposals, targets)
        segments = self.__dict__.get("segments")
        if self.mlp_head:                                     # FPN: the extractor output IS the [K,1024] vector
            x = pooled = self.feature_extractor(features, proposals)
            class_logits, box_regression = self.predictor(pooled)
        elif segments is not None and self.training:
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
