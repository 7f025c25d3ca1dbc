"""Domain-adaptation heads and losses.

Mirrors maskrcnn_benchmark/modeling/da_heads/{da_heads.py,loss.py}, layers/gradient_scalar_layer.py and
layers/consistency_loss.py.  State-dict names: ``da_heads.{imghead.conv1_da,imghead.conv2_da,
inshead.fc1_da,inshead.fc2_da,inshead.fc3_da}.{weight,bias}`` and the same under ``da_heads_triplet``.

Fusions relative to the reference: BCE/consistency/triplet are single fused forward+gradient kernels;
the sigmoid of the consistency branch is folded into the loss kernel; the AdvGRL weight is computed and
consumed on the device (no `.numpy()` / `.cpu()` sync); the AdvGRL "probe" pass of the image
head is not recomputed because it has the same forward values as the GRL pass.
"""
import contextlib

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..structures import is_source_image
from .backbone import Conv2dParams


class DAImgHead(nn.Module):
    """Two 1x1 convs on the C4 map (da_heads.py:12-37); output NHWC [N,h,w,1]."""

    def __init__(self, in_channels):
        super().__init__()
        self.conv1_da = Conv2dParams(in_channels, 512, 1, bias=True)
        self.conv2_da = Conv2dParams(512, 1, 1, bias=True)
        for l in (self.conv1_da, self.conv2_da):
            nn.init.normal_(l.weight, std=0.001)
            nn.init.constant_(l.bias, 0)

    def forward(self, feat):
        t = ops.conv_bn_act(feat, self.conv1_da.weight, None, self.conv1_da.bias, relu=True)
        return ops.fused_heads(t, [self.conv2_da.weight], [self.conv2_da.bias])[0]     # 1 channel, padded to 32


class DAInsHead(nn.Module):
    """FC 2048-1024-1024-1 with ReLU + dropout(0.5) (da_heads.py:40-68)."""

    def __init__(self, in_channels, rng):
        super().__init__()
        self.fc1_da = nn.Linear(in_channels, 1024)
        self.fc2_da = nn.Linear(1024, 1024)
        self.fc3_da = nn.Linear(1024, 1)
        for l in (self.fc1_da, self.fc2_da):
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0)
        nn.init.normal_(self.fc3_da.weight, std=0.05)
        nn.init.constant_(self.fc3_da.bias, 0)
        self.rng = rng

    def skip_draws(self, x, row_valid=None):
        """Consume the two dropout draws of a pass that is not evaluated — only a replaying random source cares
        (the oracle recorded them); the default source draws nothing."""
        if self.training and self.rng.replay:
            shape = (x.shape[0], self.fc1_da.weight.shape[0])
            self.rng.dropout_keep(shape, x.device, row_valid)
            self.rng.dropout_keep((x.shape[0], self.fc2_da.weight.shape[0]), x.device, row_valid)

    def forward(self, x, row_valid=None):
        x = ops.linear(x, self.fc1_da.weight, self.fc1_da.bias, relu=True)
        if self.training:
            x = ops.dropout_with_mask(x, self.rng.dropout_keep(tuple(x.shape), x.device, row_valid))
        x = ops.linear(x, self.fc2_da.weight, self.fc2_da.bias, relu=True)
        if self.training:
            x = ops.dropout_with_mask(x, self.rng.dropout_keep(tuple(x.shape), x.device, row_valid))
        return ops.fused_heads(x, [self.fc3_da.weight], [self.fc3_da.bias])[0]


def _image_domain_labels(targets, device):
    """prepare_masks (da_heads/loss.py:45-53): one bit per image, is_source.any()."""
    return torch.tensor([1 if is_source_image(t) else 0 for t in targets],
                        dtype=torch.uint8, device=device)


def da_img_loss(img_logits, targets, seg=None):
    """Per-pixel BCE against the image's domain label, mean over N*h*w (loss.py:140-168)."""
    n = img_logits.shape[0]
    if seg is None:
        seg = _image_domain_labels(targets, img_logits.device)
    return ops.bce_with_logits_mean(img_logits, None, seg, img_logits.numel() // n)


def da_ins_loss(ins_logits, dom, row_valid=None):
    """Per-ROI BCE against the ROI's domain (loss.py:170-174).  row_valid: slots of the fixed-capacity ROI
    layout that hold no proposal are neutralised (logit +100 vs target 1: exactly zero loss and gradient) and
    the mean is rescaled to the existing ROIs."""
    x = ins_logits.reshape(-1)
    t = dom.to(torch.float32)
    if row_valid is None:
        return ops.bce_with_logits_mean(x, t)
    v = row_valid.bool()
    x = torch.where(v, x, torch.full_like(x[:1], 100.0))
    t = torch.where(v, t, torch.ones_like(t))
    return ops.bce_with_logits_mean(x, t) * (float(x.numel()) / v.sum().to(torch.float32))


def _num_source(dom_sizes):
    return int(dom_sizes)


class _Base(nn.Module):
    def __init__(self, cfg, rng):
        super().__init__()
        self.cfg = cfg.clone()
        D = cfg.MODEL.DA_HEADS
        if not cfg.MODEL.BACKBONE.CONV_BODY.startswith("R"):
            raise NotImplementedError("only ResNet bodies are on the DA path")
        self.img_weight, self.ins_weight, self.cst_weight = D.DA_IMG_LOSS_WEIGHT, D.DA_INS_LOSS_WEIGHT, D.DA_CST_LOSS_WEIGHT
        self.imghead = DAImgHead(cfg.MODEL.BACKBONE.OUT_CHANNELS)
        self.inshead = DAInsHead(cfg.MODEL.RESNETS.RES2_OUT_CHANNELS * 8, rng)


class DomainAdaptationModule(_Base):
    """The original DA module (da_heads.py:354-440 + DALossComputation loss.py:28-104)."""

    def early_image_loss(self, feat, targets, seg=None, after=None):
        """The image-level domain loss with its backward pass run EARLY (see rpn.py::_forward_static_early): it depends
        on the trunk features only, so its forward AND backward run on their own stream right after the trunk, beside
        the latency-bound RPN loss / proposal chains.  Returns None when the term is off, else the pending
        (stream, feat, cut, weighted loss) for forward(..., early_img=...).  after: a CUDA event the pass waits for (the
        end of the RPN head, so that its dense kernels fill the time of the RPN loss chain instead of delaying the head).
        Same preconditions as the RPN's early pass: gradients zeroed before the forward pass, unit loss weights."""
        if not (self.training and self.img_weight > 0):
            return None
        D = self.cfg.MODEL.DA_HEADS
        side = None
        ctx = contextlib.nullcontext()
        if feat.is_cuda:
            if self.__dict__.get("_early_stream") is None:
                self.__dict__["_early_stream"] = torch.cuda.Stream(device=feat.device)
            side = self.__dict__["_early_stream"]
            side.wait_stream(torch.cuda.current_stream())
            if after is not None:
                side.wait_event(after)
            ctx = torch.cuda.stream(side)
        # its dense kernels leave room for what runs beside them: the proposal chain's one-CTA-per-image kernels and
        # the 8-CTA cluster of the RPN sampler
        budget = ops.sm_budget(ops.NUM_SMS - 2 * feat.shape[0] - 8) if feat.is_cuda else contextlib.nullcontext()
        with ctx, budget:
            cut = feat.detach().requires_grad_(True)
            da_img = self.imghead(ops.gradient_scalar(cut, -1.0 * D.DA_IMG_GRL_WEIGHT))
            loss = self.img_weight * da_img_loss(da_img, targets, seg)
            torch.autograd.backward([loss])
            loss = loss.detach()
        return side, feat, cut, loss

    def forward(self, img_features, pooled_ins, dom, n_src, targets, row_valid=None, seg=None, early_img=None):
        """row_valid / seg: fixed-capacity ROI slots that exist, and the cached per-image domain labels (both
        optional; supplied by the sync-free training path).  early_img: the pending result of early_image_loss()."""
        if not self.training:
            return {}
        D = self.cfg.MODEL.DA_HEADS
        feat = img_features[0]
        # The reference evaluates all four head passes and all three losses and then drops the ones whose weight
        # is 0 (da_heads.py:417-436); a dropped term has no effect on the loss dict or on any gradient, so its
        # passes are skipped here.  The instance head's dropout draws are still consumed (RNG-stream parity).
        need_img, need_ins, need_cst = self.img_weight > 0, self.ins_weight > 0, self.cst_weight > 0
        l_img = l_ins = l_cst = None
        if need_img and early_img is None:
            da_img = self.imghead(ops.gradient_scalar(feat, -1.0 * D.DA_IMG_GRL_WEIGHT))
            l_img = da_img_loss(da_img, targets, seg)
        if need_ins:
            da_ins = self.inshead(ops.gradient_scalar(pooled_ins, -1.0 * D.DA_INS_GRL_WEIGHT), row_valid)
            l_ins = da_ins_loss(da_ins, dom, row_valid)
        else:
            self.inshead.skip_draws(pooled_ins, row_valid)
        if need_cst:
            da_img_c = self.imghead(ops.gradient_scalar(feat, 1.0 * D.DA_IMG_GRL_WEIGHT))
            da_ins_c = self.inshead(ops.gradient_scalar(pooled_ins, 1.0 * D.DA_INS_GRL_WEIGHT), row_valid)
            l_cst = ops.consistency_loss(da_img_c.reshape(da_img_c.shape[0], -1), da_ins_c.reshape(-1), n_src, row_valid)
        else:
            self.inshead.skip_draws(pooled_ins, row_valid)
        losses = {}
        if self.img_weight > 0:
            if early_img is not None:            # join the early pass: stream, then the autograd graph of `feat`
                side, feat0, cut, loss = early_img
                if side is not None:
                    torch.cuda.current_stream().wait_stream(side)
                losses["loss_da_image"] = ops.inject_grad(loss, feat0, cut)
            else:
                losses["loss_da_image"] = self.img_weight * l_img
        if self.ins_weight > 0:
            losses["loss_da_instance"] = self.ins_weight * l_ins
        if self.cst_weight > 0:
            losses["loss_da_consistency"] = self.cst_weight * l_cst
        return losses


class DAImgHeadFPN(nn.Module):
    """One (1x1 conv -> ReLU -> 1x1 conv) image-level domain classifier per pyramid level with the parameter names of
    da_heads_fpn.py:37-66 (``da_img_conv{1,2}_level{i}``)."""

    def __init__(self, in_channels, levels=5):
        super().__init__()
        self.levels = levels
        for i in range(levels):
            c1, c2 = Conv2dParams(in_channels, 512, 1, bias=True), Conv2dParams(512, 1, 1, bias=True)
            for l in (c1, c2):
                nn.init.normal_(l.weight, std=0.001)
                nn.init.constant_(l.bias, 0)
            self.add_module("da_img_conv1_level{}".format(i), c1)
            self.add_module("da_img_conv2_level{}".format(i), c2)

    def forward(self, feats):
        outs = []
        for i, f in enumerate(feats):
            c1, c2 = getattr(self, "da_img_conv1_level{}".format(i)), getattr(self, "da_img_conv2_level{}".format(i))
            t = ops.conv_bn_act(f, c1.weight, None, c1.bias, relu=True)
            outs.append(ops.fused_heads(t, [c2.weight], [c2.bias])[0])
        return outs


class DAInsHeadFPN(nn.Module):
    """One FC triple (1024-1024-1024-1, ReLU + dropout 0.5) per pooler level (da_heads_fpn.py:146-207,
    ``da_ins_fc{1,2,3}_level{i}``); every ROI goes through the head of its LevelMapper level."""

    def __init__(self, in_channels, rng, levels=4):
        super().__init__()
        self.levels, self.rng = levels, rng
        for i in range(levels):
            for j, (cin, cout) in enumerate(((in_channels, 1024), (1024, 1024), (1024, 1)), 1):
                fc = nn.Linear(cin, cout)
                nn.init.normal_(fc.weight, std=0.01)
                nn.init.constant_(fc.bias, 0)
                self.add_module("da_ins_fc{}_level{}".format(j, i), fc)

    def forward(self, x, roi_levels):
        """x [K, C]; roi_levels int [K].  Host-driven like the rest of the FPN path (one size read per level)."""
        out = torch.zeros((x.shape[0], 1), dtype=x.dtype, device=x.device)
        counts = torch.bincount(roi_levels.to(torch.int64), minlength=self.levels).tolist()
        for lvl in range(self.levels):
            if counts[lvl] == 0:                                  # da_heads_fpn.py:194
                continue
            idx = torch.nonzero(roi_levels == lvl).squeeze(1)
            xs = x[idx]
            for j in (1, 2):
                fc = getattr(self, "da_ins_fc{}_level{}".format(j, lvl))
                xs = ops.linear(xs, fc.weight, fc.bias, relu=True)
                if self.training:
                    xs = ops.dropout_with_mask(xs, self.rng.dropout_keep(tuple(xs.shape), xs.device))
            fc3 = getattr(self, "da_ins_fc3_level{}".format(lvl))
            out = out.index_put((idx,), ops.fused_heads(xs, [fc3.weight], [fc3.bias])[0])
        return out


def _ins_head_fpn_static(head, x, roi_levels, row_valid):
    """DAInsHeadFPN on fixed-capacity ROI slots, without host reads: every level head runs on ALL rows and a row keeps
    the output of the head of its own level (the other products are multiplied by zero, forward and backward).  A
    replaying random source skips the levels without ROIs like the reference does (da_heads_fpn.py:194) — a host read,
    test-only — and its recorded masks are scattered into the rows of the level."""
    out = torch.zeros((x.shape[0], 1), dtype=x.dtype, device=x.device)
    valid = None if row_valid is None else row_valid.bool()
    for lvl in range(head.levels):
        m = roi_levels == lvl
        if valid is not None:
            m = m & valid
        if head.rng.replay and int(m.sum()) == 0:
            continue
        xs = x
        for j in (1, 2):
            fc = getattr(head, "da_ins_fc{}_level{}".format(j, lvl))
            xs = ops.linear(xs, fc.weight, fc.bias, relu=True)
            if head.training:
                xs = ops.dropout_with_mask(xs, head.rng.dropout_keep(tuple(xs.shape), xs.device, m))
        fc3 = getattr(head, "da_ins_fc3_level{}".format(lvl))
        out = out + ops.fused_heads(xs, [fc3.weight], [fc3.bias])[0] * m.to(x.dtype).unsqueeze(1)
    return out


class DomainAdaptationModuleFPN(nn.Module):
    """DA heads on an FPN backbone (BASELINE configs[4]).  PARITY UNPINNED: the reference has no runnable FPN + DA
    combination; this is the intent of da_heads_fpn.py:209-295 with its defects resolved as listed in
    oracle/fpn_ref.py (per-level image heads, per-level instance heads routed by LevelMapper without the stray
    `return`, COS_WEIGHT := 1.0, image BCE over the pixels of all levels, layers/consistency_loss.py over the list of
    levels, DA_*_LOSS_WEIGHT applied as the C4 module does).  Checked against that oracle restatement."""

    def __init__(self, cfg, rng):
        super().__init__()
        self.cfg = cfg.clone()
        D = cfg.MODEL.DA_HEADS
        self.img_weight, self.ins_weight, self.cst_weight = D.DA_IMG_LOSS_WEIGHT, D.DA_INS_LOSS_WEIGHT, D.DA_CST_LOSS_WEIGHT
        self.imghead = DAImgHeadFPN(cfg.MODEL.BACKBONE.OUT_CHANNELS, levels=5)
        self.inshead = DAInsHeadFPN(cfg.MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM, rng, levels=len(cfg.MODEL.ROI_BOX_HEAD.POOLER_SCALES))

    def forward(self, img_features, ins_feas, dom, n_src, targets, roi_levels, row_valid=None, seg=None):
        """row_valid / seg: the ROI slots that exist and the cached per-image domain labels — given by the sync-free
        training path (fixed-capacity ROI slots, n_src = slots per image); then no step reads a size on the host."""
        if not self.training:
            return {}
        D = self.cfg.MODEL.DA_HEADS
        need_img, need_ins, need_cst = self.img_weight > 0, self.ins_weight > 0, self.cst_weight > 0
        losses = {}
        static = row_valid is not None
        ins = (lambda z: _ins_head_fpn_static(self.inshead, z, roi_levels, row_valid)) if static else \
              (lambda z: self.inshead(z, roi_levels))
        # (same evaluation order as the oracle: instance GRL pass, instance consistency pass, image passes —
        # the dropout draws of the instance passes are consumed in that order)
        da_ins = ins(ops.gradient_scalar(ins_feas, -1.0 * D.DA_INS_GRL_WEIGHT))
        da_ins_c = ins(ops.gradient_scalar(ins_feas, 1.0 * D.DA_INS_GRL_WEIGHT))
        if need_img:
            da_img = self.imghead([ops.gradient_scalar(f, -1.0 * D.DA_IMG_GRL_WEIGHT) for f in img_features])
            n = da_img[0].shape[0]
            flat = torch.cat([t.reshape(n, -1) for t in da_img], dim=1)
            losses["loss_da_image"] = self.img_weight * da_img_loss(flat, targets, seg)
        if need_ins:
            losses["loss_da_instance"] = self.ins_weight * da_ins_loss(da_ins, dom, row_valid)
        if need_cst:
            da_img_c = self.imghead([ops.gradient_scalar(f, 1.0 * D.DA_IMG_GRL_WEIGHT) for f in img_features])
            # mean over ROIs x levels of |image-level mean probability - ROI probability| = the mean of the
            # per-level consistency losses (every level has the same K rows)
            terms = [ops.consistency_loss(t.reshape(t.shape[0], -1), da_ins_c.reshape(-1), n_src, row_valid)
                     for t in da_img_c]
            losses["loss_da_consistency"] = self.cst_weight * (sum(terms) / float(len(terms)))
        return losses


ADV_BCE = float(F.binary_cross_entropy_with_logits(torch.tensor([[0.7, 0.3]]), torch.tensor([[1.0, 0.0]])))


class DomainAdaptationModule_triplet(_Base):
    """DA module with AdvGRL and the auxiliary-domain triplet losses (da_heads.py:72-344 +
    DALossComputation_Component loss.py:108-222)."""

    def __init__(self, cfg, rng):
        super().__init__(cfg, rng)
        D = cfg.MODEL.DA_HEADS
        self.triplet_img_weight, self.triplet_ins_weight = D.DA_TRIPLET_IMG_WEIGHT, D.DA_TRIPLET_INS_WEIGHT
        # State of the reference (da_heads.py:107-113, loss.py:128-130).  The instance margin is not adaptive (a
        # constant).  The adaptive IMAGE margin and the previous image-triplet loss it consults live on the device
        # (non-persistent buffers: they are not part of the reference's state dict), so that no configuration reads
        # a loss on the host and every one of them can be captured into the whole-step graph.
        self.margin_ins = 0.0
        self.register_buffer("margin_img_state", torch.zeros(1, dtype=torch.float64), persistent=False)
        self.register_buffer("margin_img_dev", torch.zeros(1, dtype=torch.float32), persistent=False)
        self.register_buffer("prev_img", torch.ones(1, dtype=torch.float32), persistent=False)   # triplet_img = [1]

    @property
    def margin_img(self):
        """The current image-level margin (host read; for logging and tests)."""
        return float(self.margin_img_state)

    def _grl_weight(self, loss, lam, lam_adv):
        D = self.cfg.MODEL.DA_HEADS
        if D.DA_ADV_GRL:
            return ops.adv_grl_weight(loss, ADV_BCE, lam, lam_adv, D.DA_ADV_GRL_THRESHOLD)
        return torch.full((1,), -1.0 * lam, dtype=torch.float32, device=loss.device)

    def forward(self, img_features, pooled_ins, dom, n_src, pooled_set, img_fea_set, targets, row_valid=None, seg=None,
                set_valid=None):
        """row_valid / set_valid: validity of the fixed-capacity ROI slots of `pooled_ins` / of each member of
        `pooled_set`; seg: cached per-image domain labels (all optional, supplied by the sync-free path)."""
        if not self.training:
            return {}
        D = self.cfg.MODEL.DA_HEADS
        feat = img_features[0]
        losses = {}
        if self.triplet_ins_weight > 0:                       # Domainlevel_Ins_component
            s, p, n = pooled_set
            self.margin_ins = D.TRIPLET_MARGIN_INS           # adaptive=False (da_heads.py:266, loss.py:214-216)
            if set_valid is not None:
                # slots that hold no ROI: anchor = positive and a far-away negative put the hinge at exactly zero
                # (no loss, no gradient); the mean is rescaled to the rows that exist
                v = (set_valid[0] & set_valid[1] & set_valid[2]).bool().unsqueeze(1)
                zero = torch.zeros_like(s[:1])
                s, p, n = torch.where(v, s, zero), torch.where(v, p, zero), torch.where(v, n, zero + 1.0e3)
                l = ops.triplet_margin_loss(s, p, n, self.margin_ins, s.shape[0], s.shape[1], 1)
                l = l * (float(s.shape[0]) / v.sum().to(torch.float32))
            else:
                l = ops.triplet_margin_loss(s, p, n, self.margin_ins, s.shape[0], s.shape[1], 1)
            losses["triplet_loss_instance"] = self.triplet_ins_weight * l
        if self.triplet_img_weight > 0:                       # Domainlevel_Img_component
            s, p, n = img_fea_set
            # adaptive margin (loss.py:182-200): += 0.001 whenever the previous step's loss was exactly 0 and
            # int(margin) != int(max margin) — evaluated on the device from the kept previous loss
            margin = ops.adaptive_margin_update(self.margin_img_state, self.prev_img, D.TRIPLET_MARGIN_IMG, 0.001,
                                                D.TRIPLET_MAX_MARGIN, out=self.margin_img_dev)
            _, h, w, c = s.shape                              # NHWC [1,h,w,C]: distance over w per (c,h)
            l = ops.triplet_margin_loss(s, p, n, margin, h * c, w, c)
            losses["triplet_loss_image"] = self.triplet_img_weight * l
            with torch.no_grad():
                self.prev_img.copy_(l.detach().reshape(1))
        if self.img_weight > 0:                               # DA_Img_component
            wdev = torch.empty(1, dtype=torch.float32, device=feat.device)
            da_img = self.imghead(ops.gradient_scalar_dev(feat, wdev))
            l_img = da_img_loss(da_img, targets, seg)
            wdev.copy_(self._grl_weight(l_img, D.DA_IMG_GRL_WEIGHT, D.DA_IMG_advGRL_WEIGHT))
            losses["loss_da_image"] = self.img_weight * l_img
        if self.ins_weight > 0:                               # DA_Ins_component
            with torch.no_grad():
                cur = da_ins_loss(self.inshead(pooled_ins.detach(), row_valid), dom, row_valid)
            w_ins = self._grl_weight(cur, D.DA_INS_GRL_WEIGHT, D.DA_INS_advGRL_WEIGHT)
            da_ins = self.inshead(ops.gradient_scalar_dev(pooled_ins, w_ins), row_valid)
            losses["loss_da_instance"] = self.ins_weight * da_ins_loss(da_ins, dom, row_valid)
        if self.cst_weight > 0:                               # Consistency_component
            img_c = self.imghead(ops.gradient_scalar(feat, 1.0 * D.DA_IMG_GRL_WEIGHT))
            ins_c = self.inshead(ops.gradient_scalar(pooled_ins, 1.0 * D.DA_INS_GRL_WEIGHT), row_valid)
            l_cst = ops.consistency_loss(img_c.reshape(img_c.shape[0], -1), ins_c.reshape(-1), n_src, row_valid)
            losses["loss_da_consistency"] = self.cst_weight * l_cst
        return losses


def build_da_heads(cfg, rng):
    if cfg.MODEL.DOMAIN_ADAPTATION_ON:
        if cfg.MODEL.BACKBONE.CONV_BODY.endswith("-FPN"):
            return DomainAdaptationModuleFPN(cfg, rng)
        return DomainAdaptationModule(cfg, rng)
    return []


def build_da_heads_triplet(cfg, rng):
    if cfg.MODEL.DOMAIN_ADAPTATION_ON:
        return DomainAdaptationModule_triplet(cfg, rng)
    return []
