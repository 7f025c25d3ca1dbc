"""TEST INFRASTRUCTURE ONLY — CPU oracle for the DA Faster R-CNN training hot path.

A functional, plain-torch (fp32, NCHW, CPU) restatement of what one call of the
reference's ``GeneralizedRCNN.forward(images, targets)`` computes in training mode
(maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:61-156) for the R-50-C4
domain-adaptive configurations, plus the eval-mode proposal/box post-processing.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
import this module; the product path (da-detect_b200/) never does.

Pinning: tests/test_oracle_pins.py checks this file against
  * the reference's golden vectors (tests/test_nms.py, tests/test_box_coder.py,
    the anchor table in modeling/rpn/anchor_generator.py:201-219), and
  * tests/golden/*.pt — outputs of the REAL reference modules imported from
    /root/reference in the build container by oracle/make_golden.py (committed).

Every function cites the reference lines it follows.  Parameters are a flat dict
with the reference's state-dict names.  Randomness is injected
through ``Hooks`` so that a CUDA run and this CPU run can consume identical draws.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    """ctypes handle on oracle/liboracle_ref.so (built by oracle/Makefile)."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_ref.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = ctypes.CDLL(so)
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        ci, cf = ctypes.c_int, ctypes.c_float
        L.ref_roi_align_forward.argtypes = [f32p, f32p, ci, ci, ci, ci, cf, ci, ci, ci, f32p]
        L.ref_roi_align_backward.argtypes = [f32p, f32p, ci, ci, ci, ci, cf, ci, ci, ci, f32p]
        L.ref_nms.argtypes = [f32p, i64p, ci, cf, ci, i64p]
        L.ref_nms.restype = ci
        _LIB = L
    return _LIB


def _fp(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ip(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int64))


# --------------------------------------------------------------------------- native ops
class _RoiAlignRef(torch.autograd.Function):
    """layers/roi_align.py:11-44 over csrc ROIAlign (see roialign_nms_ref.c)."""

    @staticmethod
    def forward(ctx, x, rois, scale, ph, pw, sampling_ratio):
        x = x.contiguous().float()
        rois = rois.contiguous().float()
        n, c, h, w = x.shape
        out = torch.empty(rois.shape[0], c, ph, pw)
        lib().ref_roi_align_forward(_fp(x), _fp(rois), rois.shape[0], c, h, w, scale, ph, pw,
                                    sampling_ratio, _fp(out))
        ctx.save_for_backward(rois)
        ctx.meta = (x.shape, scale, ph, pw, sampling_ratio)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        shape, scale, ph, pw, sr = ctx.meta
        g = g.contiguous().float()
        gin = torch.zeros(shape)
        lib().ref_roi_align_backward(_fp(g), _fp(rois), rois.shape[0], shape[1], shape[2], shape[3],
                                     scale, ph, pw, sr, _fp(gin))
        return gin, None, None, None, None, None


def roi_align(x, rois, scale, ph, pw, sampling_ratio):
    if x.is_cuda:
        # "reference graph on the same GPU" timing leg only (bench.py torch_cudnn_baseline): the reference's csrc
        # CUDA kernels cannot be built (THC removed from torch); torchvision's roi_align(aligned=False)
        # implements the same ROIAlign arithmetic.  Never used as a parity checker.
        import torchvision
        return torchvision.ops.roi_align(x, rois, (int(ph), int(pw)), float(scale), int(sampling_ratio), aligned=False)
    return _RoiAlignRef.apply(x, rois, float(scale), int(ph), int(pw), int(sampling_ratio))


def nms(boxes, scores, thresh, strict=True):
    """csrc/nms.h:10-28.  strict=True is the CUDA reference (IoU > thresh, nms.cu:60),
    strict=False the CPU one (>=, nms_cpu.cpp:60).  Returns ascending kept indices."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64)
    if boxes.is_cuda:                 # same-GPU timing leg only (see roi_align): stock torchvision NMS
        import torchvision
        return torchvision.ops.nms(boxes.float(), scores.float(), float(thresh)).sort()[0]
    boxes = boxes.contiguous().float()
    order = torch.sort(scores.float(), descending=True, stable=True)[1].contiguous()
    keep = torch.empty(n, dtype=torch.int64)
    m = lib().ref_nms(_fp(boxes), _ip(order), n, float(thresh), 1 if strict else 0, _ip(keep))
    return keep[:m].clone()


# --------------------------------------------------------------------------- hooks
class Hooks(object):
    """Source of the random draws on the path, in reference call order:
    ``randperm(n)`` (balanced_positive_negative_sampler.py:57-58) and
    ``dropout_keep(shape)`` -> float {0,1} keep mask with p_keep = 0.5
    (da_heads.py:63,65 through F.dropout).  Defaults consume torch's global CPU
    generator exactly like the reference's own calls do."""

    def randperm(self, n):
        return torch.randperm(n)

    def dropout_keep(self, shape):
        return torch.empty(shape).bernoulli_(0.5)


class RecordingHooks(Hooks):
    """Records every draw so that another implementation can replay them."""

    def __init__(self, base=None):
        self.base = base or Hooks()
        self.perms, self.masks = [], []

    def randperm(self, n):
        p = self.base.randperm(n)
        self.perms.append(p.clone())
        return p

    def dropout_keep(self, shape):
        m = self.base.dropout_keep(shape)
        self.masks.append(m.clone())
        return m


class ReplayHooks(Hooks):
    def __init__(self, perms, masks):
        self.perms, self.masks = list(perms), list(masks)

    def randperm(self, n):
        p = self.perms.pop(0)
        assert p.numel() == n, "randperm replay out of step: want {} have {}".format(n, p.numel())
        return p.clone()

    def dropout_keep(self, shape):
        m = self.masks.pop(0)
        assert tuple(m.shape) == tuple(shape)
        return m.clone()


# --------------------------------------------------------------------------- anchors
def cell_anchors(stride, sizes, ratios):
    """generate_anchors/_generate_anchors/_ratio_enum/_scale_enum
    (rpn/anchor_generator.py:222-291): float64 numpy, np.round, then .float()."""
    scales = np.array(sizes, dtype=np.float64) / stride
    ratios = np.array(ratios, dtype=np.float64)
    base = np.array([1, 1, stride, stride], dtype=np.float64) - 1

    def whc(a):
        w = a[2] - a[0] + 1
        h = a[3] - a[1] + 1
        return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)

    def mk(ws, hs, xc, yc):
        ws, hs = ws[:, None], hs[:, None]
        return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))

    w, h, xc, yc = whc(base)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    by_ratio = mk(ws, hs, xc, yc)
    rows = []
    for a in by_ratio:
        w, h, xc, yc = whc(a)
        rows.append(mk(w * scales, h * scales, xc, yc))
    return torch.from_numpy(np.vstack(rows)).float()


def grid_anchors(fh, fw, stride, cell):
    """AnchorGenerator.grid_anchors (anchor_generator.py:73-95): order (y, x, anchor)."""
    sx = torch.arange(0, fw * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(0, fh * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    xx, yy = xx.reshape(-1), yy.reshape(-1)
    shifts = torch.stack((xx, yy, xx, yy), dim=1)
    return (shifts.view(-1, 1, 4) + cell.to(shifts.device).view(1, -1, 4)).reshape(-1, 4)


def anchor_visibility(anchors, img_w, img_h, straddle):
    """add_visibility_to (anchor_generator.py:97-111)."""
    if straddle < 0:
        return torch.ones(anchors.shape[0], dtype=torch.bool)
    return ((anchors[:, 0] >= -straddle) & (anchors[:, 1] >= -straddle)
            & (anchors[:, 2] < img_w + straddle) & (anchors[:, 3] < img_h + straddle))


# --------------------------------------------------------------------------- box math
BBOX_XFORM_CLIP = math.log(1000.0 / 16)


def box_encode(ref, prop, weights):
    """BoxCoder.encode (box_coder.py:22-50)."""
    ew = prop[:, 2] - prop[:, 0] + 1
    eh = prop[:, 3] - prop[:, 1] + 1
    ex = prop[:, 0] + 0.5 * ew
    ey = prop[:, 1] + 0.5 * eh
    gw = ref[:, 2] - ref[:, 0] + 1
    gh = ref[:, 3] - ref[:, 1] + 1
    gx = ref[:, 0] + 0.5 * gw
    gy = ref[:, 1] + 0.5 * gh
    wx, wy, ww, wh = weights
    return torch.stack((wx * (gx - ex) / ew, wy * (gy - ey) / eh,
                        ww * torch.log(gw / ew), wh * torch.log(gh / eh)), dim=1)


def box_decode(codes, boxes, weights):
    """BoxCoder.decode (box_coder.py:52-95); codes [R, 4*k]."""
    boxes = boxes.to(codes.dtype)
    w = boxes[:, 2] - boxes[:, 0] + 1
    h = boxes[:, 3] - boxes[:, 1] + 1
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx = codes[:, 0::4] / wx
    dy = codes[:, 1::4] / wy
    dw = torch.clamp(codes[:, 2::4] / ww, max=BBOX_XFORM_CLIP)
    dh = torch.clamp(codes[:, 3::4] / wh, max=BBOX_XFORM_CLIP)
    pcx = dx * w[:, None] + cx[:, None]
    pcy = dy * h[:, None] + cy[:, None]
    pw = torch.exp(dw) * w[:, None]
    ph = torch.exp(dh) * h[:, None]
    out = torch.zeros_like(codes)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw - 1
    out[:, 3::4] = pcy + 0.5 * ph - 1
    return out


def box_iou(a, b):
    """boxlist_iou (structures/boxlist_ops.py:56-91), '+1' areas."""
    area_a = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    area_b = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    lt = torch.max(a[:, None, :2], b[:, :2])
    rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (area_a[:, None] + area_b - inter)


def clip_boxes(b, img_w, img_h):
    """BoxList.clip_to_image(remove_empty=False) (bounding_box.py:214-225)."""
    b = b.clone()
    b[:, 0].clamp_(min=0, max=img_w - 1)
    b[:, 1].clamp_(min=0, max=img_h - 1)
    b[:, 2].clamp_(min=0, max=img_w - 1)
    b[:, 3].clamp_(min=0, max=img_h - 1)
    return b


BELOW_LOW, BETWEEN = -1, -2


def matcher(q, high, low, allow_low_quality):
    """Matcher.__call__ + set_low_quality_matches_ (matcher.py:42-112); q is [M gt, N pred]."""
    if q.numel() == 0:
        raise ValueError("No ground-truth or proposal boxes available for one of the images during training")
    vals, matches = q.max(dim=0)
    all_matches = matches.clone()
    below = vals < low
    between = (vals >= low) & (vals < high)
    matches[below] = BELOW_LOW
    matches[between] = BETWEEN
    if allow_low_quality:
        best_per_gt, _ = q.max(dim=1)
        pairs = torch.nonzero(q == best_per_gt[:, None])
        upd = pairs[:, 1]
        matches[upd] = all_matches[upd]
    return matches


def balanced_sampler(labels_per_image, batch_size, pos_fraction, hooks):
    """BalancedPositiveNegativeSampler.__call__ (balanced_positive_negative_sampler.py:27-76)."""
    pos_masks, neg_masks = [], []
    for lab in labels_per_image:
        positive = torch.nonzero(lab >= 1).squeeze(1)
        negative = torch.nonzero(lab == 0).squeeze(1)
        num_pos = min(positive.numel(), int(batch_size * pos_fraction))
        num_neg = min(negative.numel(), batch_size - num_pos)
        perm1 = hooks.randperm(positive.numel())[:num_pos]
        perm2 = hooks.randperm(negative.numel())[:num_neg]
        pm = torch.zeros_like(lab, dtype=torch.bool)
        nm = torch.zeros_like(lab, dtype=torch.bool)
        pm[positive[perm1]] = True
        nm[negative[perm2]] = True
        pos_masks.append(pm)
        neg_masks.append(nm)
    return pos_masks, neg_masks


def smooth_l1(x, t, beta, size_average):
    """layers/smooth_l1_loss.py:6-16."""
    n = torch.abs(x - t)
    loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    return loss.mean() if size_average else loss.sum()


# --------------------------------------------------------------------------- backbone
def frozen_bn(x, P, name):
    """FrozenBatchNorm2d.forward (layers/batch_norm.py:19-24)."""
    scale = P[name + ".weight"] * P[name + ".running_var"].rsqrt()
    bias = P[name + ".bias"] - P[name + ".running_mean"] * scale
    return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


def bottleneck(x, P, name, stride):
    """Bottleneck.forward (backbone/resnet.py:294-314), STRIDE_IN_1X1=True (:262-270)."""
    identity = x
    out = F.relu(frozen_bn(F.conv2d(x, P[name + ".conv1.weight"], stride=stride), P, name + ".bn1"))
    out = F.relu(frozen_bn(F.conv2d(out, P[name + ".conv2.weight"], padding=1), P, name + ".bn2"))
    out = frozen_bn(F.conv2d(out, P[name + ".conv3.weight"]), P, name + ".bn3")
    if (name + ".downsample.0.weight") in P:
        identity = frozen_bn(F.conv2d(x, P[name + ".downsample.0.weight"], stride=stride), P,
                             name + ".downsample.1")
    return F.relu(out + identity)


def stage(x, P, name, blocks, first_stride):
    for i in range(blocks):
        x = bottleneck(x, P, "{}.{}".format(name, i), first_stride if i == 0 else 1)
    return x


STAGE_BLOCKS = {"R-50-C4": (3, 4, 6), "R-101-C4": (3, 4, 23)}


def backbone_c4(images, P, conv_body="R-50-C4"):
    """BaseStem (resnet.py:331-336) + ResNet.forward (:138-145) for *-C4: one level out."""
    x = F.conv2d(images, P["backbone.body.stem.conv1.weight"], stride=2, padding=3)
    x = F.relu(frozen_bn(x, P, "backbone.body.stem.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    b = STAGE_BLOCKS[conv_body]
    x = stage(x, P, "backbone.body.layer1", b[0], 1)
    x = stage(x, P, "backbone.body.layer2", b[1], 2)
    x = stage(x, P, "backbone.body.layer3", b[2], 2)
    return x


def res5_head(x, P):
    """ResNetHead with stage index 4, 3 blocks, first stride 2
    (resnet.py:148-194 as built at roi_box_feature_extractors.py:27-37)."""
    return stage(x, P, "roi_heads.box.feature_extractor.head.layer4", 3, 2)


# --------------------------------------------------------------------------- RPN
def rpn_head(feat, P):
    """RPNHead.forward (rpn/rpn.py:39-46)."""
    t = F.relu(F.conv2d(feat, P["rpn.head.conv.weight"], P["rpn.head.conv.bias"], padding=1))
    logits = F.conv2d(t, P["rpn.head.cls_logits.weight"], P["rpn.head.cls_logits.bias"])
    deltas = F.conv2d(t, P["rpn.head.bbox_pred.weight"], P["rpn.head.bbox_pred.bias"])
    return logits, deltas


def permute_and_flatten(layer, n, c, h, w):
    """rpn/utils.py:10-14: [N, A*C, H, W] -> [N, H*W*A, C]."""
    return layer.view(n, -1, c, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, c)


def rpn_proposals(anchors, logits, deltas, image_sizes, pre_nms, post_nms, nms_thresh, min_size,
                  nms_strict=True):
    """RPNPostProcessor.forward_for_single_feature_map (rpn/inference.py:76-123).
    Returns per image (boxes [P,4], objectness [P])."""
    n, a, h, w = logits.shape
    obj = permute_and_flatten(logits, n, 1, h, w).view(n, -1).sigmoid()
    reg = permute_and_flatten(deltas, n, 4, h, w)
    k = min(pre_nms, a * h * w)
    obj, idx = obj.topk(k, dim=1, sorted=True)
    out = []
    for i in range(n):
        props = box_decode(reg[i][idx[i]], anchors[idx[i]], (1.0, 1.0, 1.0, 1.0))
        ih, iw = image_sizes[i]
        props = clip_boxes(props, iw, ih)
        ws = props[:, 2] - props[:, 0] + 1
        hs = props[:, 3] - props[:, 1] + 1
        keep = torch.nonzero((ws >= min_size) & (hs >= min_size)).squeeze(1)  # boxlist_ops.py:37-51
        props, sc = props[keep], obj[i][keep]
        keep = nms(props, sc, nms_thresh, strict=nms_strict)                   # boxlist_ops.py:11-34
        if post_nms > 0:
            keep = keep[:post_nms]
        out.append((props[keep], sc[keep]))
    return out


def rpn_loss(anchors, visibility, logits, deltas, gt_boxes, is_source_img, cfg, hooks, aux=None):
    """RPNLossComputation.prepare_targets + __call__ (rpn/loss.py:57-143): labels only
    for source images; predictions flattened over ALL images (index alignment relies on
    source-first order)."""
    R = cfg.MODEL.RPN
    labels, reg_targets = [], []
    for i, (gt, src) in enumerate(zip(gt_boxes, is_source_img)):
        if not src:
            continue
        vis_i = visibility[i] if isinstance(visibility, (list, tuple)) else visibility
        q = box_iou(gt, anchors)
        m = matcher(q, R.FG_IOU_THRESHOLD, R.BG_IOU_THRESHOLD, True)
        matched = gt[m.clamp(min=0)]
        lab = (m >= 0).to(torch.float32)
        lab[m == BELOW_LOW] = 0
        lab[~vis_i] = -1
        lab[m == BETWEEN] = -1
        labels.append(lab)
        reg_targets.append(box_encode(matched, anchors, (1.0, 1.0, 1.0, 1.0)))
    pos_m, neg_m = balanced_sampler(labels, R.BATCH_SIZE_PER_IMAGE, R.POSITIVE_FRACTION, hooks)
    pos = torch.nonzero(torch.cat(pos_m, dim=0)).squeeze(1)
    neg = torch.nonzero(torch.cat(neg_m, dim=0)).squeeze(1)
    sampled = torch.cat([pos, neg], dim=0)
    n, a, h, w = logits.shape
    obj = permute_and_flatten(logits, n, 1, h, w).reshape(-1)
    reg = permute_and_flatten(deltas, n, 4, h, w).reshape(-1, 4)
    labels = torch.cat(labels, dim=0)
    reg_targets = torch.cat(reg_targets, dim=0)
    box_loss = smooth_l1(reg[pos], reg_targets[pos], 1.0 / 9, False) / sampled.numel()
    obj_loss = F.binary_cross_entropy_with_logits(obj[sampled], labels[sampled])
    if aux is not None:
        aux.update(rpn_labels=labels, rpn_pos=pos, rpn_neg=neg, rpn_reg_targets=reg_targets)
    return obj_loss, box_loss


# --------------------------------------------------------------------------- box head
def box_head_prepare(props, gt_boxes, gt_labels, is_source_img, cfg, sample_for_da=False):
    """FastRCNNLossComputation.prepare_targets (box_head/loss.py:55-93)."""
    H = cfg.MODEL.ROI_HEADS
    labels, regs, domains = [], [], []
    for p, gt, gl, src in zip(props, gt_boxes, gt_labels, is_source_img):
        m = matcher(box_iou(gt, p), H.FG_IOU_THRESHOLD, H.BG_IOU_THRESHOLD, False)
        sel = m.clamp(min=0) if src else m          # loss.py:47-51: raw -1/-2 wrap for target images
        lab = gl[sel].to(torch.int64).clone()
        lab[m == BELOW_LOW] = 0
        lab[m == BETWEEN] = -1
        regs.append(box_encode(gt[sel], p, H.BBOX_REG_WEIGHTS))
        domains.append(torch.full_like(lab, bool(src), dtype=torch.bool))
        if not src or sample_for_da:
            lab[:] = 0
        labels.append(lab)
    return labels, regs, domains


def box_head_subsample(props, gt_boxes, gt_labels, is_source_img, cfg, hooks):
    """subsample (box_head/loss.py:95-130) followed by subsample_for_da (:132-163).  The
    second call re-selects every proposal in order; it is executed here
    only for its two randperm draws per image."""
    H = cfg.MODEL.ROI_HEADS
    labels, regs, domains = box_head_prepare([b for b, _ in props], gt_boxes, gt_labels, is_source_img, cfg)
    pos_m, neg_m = balanced_sampler(labels, H.BATCH_SIZE_PER_IMAGE, H.POSITIVE_FRACTION, hooks)
    out = []
    for (b, s), lab, rg, dm, pm, nm in zip(props, labels, regs, domains, pos_m, neg_m):
        idx = torch.nonzero(pm | nm).squeeze(1)
        out.append(dict(boxes=b[idx], objectness=s[idx], labels=lab[idx], regression_targets=rg[idx],
                        domain_labels=dm[idx], sampled_idx=idx))
    lab2, _, _ = box_head_prepare([(o["boxes"]) for o in out], gt_boxes, gt_labels, is_source_img, cfg, True)
    pm2, nm2 = balanced_sampler(lab2, H.BATCH_SIZE_PER_IMAGE, H.POSITIVE_FRACTION, hooks)
    for o, a, b in zip(out, pm2, nm2):
        assert bool((a | b).all()), "subsample_for_da must re-select every proposal"
    return out


def rois_from(samples):
    """Pooler.convert_to_roi_format (poolers.py:78-89)."""
    return torch.cat([torch.cat([torch.full((len(s["boxes"]), 1), float(i)), s["boxes"]], dim=1)
                      for i, s in enumerate(samples)], dim=0)


def box_feature_extractor(feat, rois, P, cfg):
    """ResNet50Conv5ROIFeatureExtractor.forward (roi_box_feature_extractors.py:42-45)."""
    B = cfg.MODEL.ROI_BOX_HEAD
    x = roi_align(feat, rois, B.POOLER_SCALES[0], B.POOLER_RESOLUTION, B.POOLER_RESOLUTION,
                  B.POOLER_SAMPLING_RATIO)
    return res5_head(x, P)


def box_predictor(x, P):
    """FastRCNNPredictor.forward (roi_box_predictors.py:28-33)."""
    v = F.avg_pool2d(x, 7, 7).view(x.size(0), -1)
    return (F.linear(v, P["roi_heads.box.predictor.cls_score.weight"], P["roi_heads.box.predictor.cls_score.bias"]),
            F.linear(v, P["roi_heads.box.predictor.bbox_pred.weight"], P["roi_heads.box.predictor.bbox_pred.bias"]))


def fastrcnn_loss(cls_logits, box_reg, samples):
    """FastRCNNLossComputation.__call__ (box_head/loss.py:165-221)."""
    labels = torch.cat([s["labels"] for s in samples])
    regt = torch.cat([s["regression_targets"] for s in samples])
    dom = torch.cat([s["domain_labels"] for s in samples])
    cls_logits, box_reg, labels, regt = cls_logits[dom], box_reg[dom], labels[dom], regt[dom]
    cls_loss = F.cross_entropy(cls_logits, labels)
    pos = torch.nonzero(labels > 0).squeeze(1)
    cols = 4 * labels[pos][:, None] + torch.tensor([0, 1, 2, 3])
    box_loss = smooth_l1(box_reg[pos[:, None], cols], regt[pos], 1.0, False) / labels.numel()
    return cls_loss, box_loss, dom


def roi_box_head_train(feat, props, gt_boxes, gt_labels, is_source_img, P, cfg, hooks, aux=None, tag=""):
    """ROIBoxHead.forward, training branch (box_head/box_head.py:36-117).  The second,
    numerically identical feature-extractor pass (:104-110) is computed once here and
    used twice, so autograd sums both gradient contributions exactly like the reference."""
    samples = box_head_subsample(props, gt_boxes, gt_labels, is_source_img, cfg, hooks)
    rois = rois_from(samples)
    x = box_feature_extractor(feat, rois, P, cfg)
    cls_logits, box_reg = box_predictor(x, P)
    cls_loss, box_loss, dom = fastrcnn_loss(cls_logits, box_reg, samples)
    if aux is not None:
        aux[tag + "samples"] = samples
        aux[tag + "rois"] = rois
        aux[tag + "class_logits"] = cls_logits
        aux[tag + "box_regression"] = box_reg
    return dict(loss_classifier=cls_loss, loss_box_reg=box_loss), x, dom


# --------------------------------------------------------------------------- DA heads
class _GRL(torch.autograd.Function):
    """layers/gradient_scalar_layer.py:4-13."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.w = w
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.w * g, None


def grl(x, w):
    return _GRL.apply(x, float(w))


def da_img_head(feat, P, pre):
    """DAImgHead.forward (da_heads/da_heads.py:32-37)."""
    t = F.relu(F.conv2d(feat, P[pre + ".imghead.conv1_da.weight"], P[pre + ".imghead.conv1_da.bias"]))
    return F.conv2d(t, P[pre + ".imghead.conv2_da.weight"], P[pre + ".imghead.conv2_da.bias"])


def da_ins_head(x, P, pre, hooks):
    """DAInsHead.forward (da_heads.py:61-68): dropout p=0.5 in training, scale 1/(1-p)."""
    x = F.relu(F.linear(x, P[pre + ".inshead.fc1_da.weight"], P[pre + ".inshead.fc1_da.bias"]))
    x = x * hooks.dropout_keep(x.shape) * 2.0
    x = F.relu(F.linear(x, P[pre + ".inshead.fc2_da.weight"], P[pre + ".inshead.fc2_da.bias"]))
    x = x * hooks.dropout_keep(x.shape) * 2.0
    return F.linear(x, P[pre + ".inshead.fc3_da.weight"], P[pre + ".inshead.fc3_da.bias"])


def da_img_loss(da_img, is_source_img):
    """da_heads/loss.py:140-168 (== :66-97): per-pixel BCE-with-logits, mean over N*H*W."""
    n, a, h, w = da_img.shape
    x = da_img.permute(0, 2, 3, 1)
    lab = torch.zeros_like(x)
    lab[torch.tensor(is_source_img, dtype=torch.bool), :] = 1
    return F.binary_cross_entropy_with_logits(x.reshape(n, -1), lab.reshape(n, -1))


def da_ins_loss(da_ins, dom):
    """da_heads/loss.py:170-174 (== :98-100) with the CPU patch .float()."""
    return F.binary_cross_entropy_with_logits(torch.squeeze(da_ins), dom.float())


def consistency_loss(img_sig, ins_sig, dom):
    """layers/consistency_loss.py:3-27, one level, N must be 2; ROI->image assignment by
    the COUNT of source ROIs first."""
    n = img_sig.shape[0]
    assert n == 2, "only batch size=2 is supported for consistency loss now, received batch size: {}".format(n)
    k = ins_sig.size(0)
    n_src = int(torch.nonzero(dom).size(0))
    means = img_sig.reshape(n, -1).mean(1)
    per_roi = torch.cat([means[0].view(1, 1).repeat(n_src, 1), means[1].view(1, 1).repeat(k - n_src, 1)], dim=0)
    return torch.abs(per_roi - ins_sig).mean()


def da_heads_original(feat, da_ins_feas, dom, is_source_img, P, cfg, hooks, pre="da_heads"):
    """DomainAdaptationModule.forward (da_heads.py:388-440) + DALossComputation (loss.py:55-104)."""
    D = cfg.MODEL.DA_HEADS
    v = F.avg_pool2d(da_ins_feas, 7, 7).view(da_ins_feas.size(0), -1)
    img_g = grl(feat, -1.0 * D.DA_IMG_GRL_WEIGHT)
    ins_g = grl(v, -1.0 * D.DA_INS_GRL_WEIGHT)
    img_c = grl(feat, 1.0 * D.DA_IMG_GRL_WEIGHT)
    ins_c = grl(v, 1.0 * D.DA_INS_GRL_WEIGHT)
    da_img = da_img_head(img_g, P, pre)
    da_ins = da_ins_head(ins_g, P, pre, hooks)
    da_img_c = da_img_head(img_c, P, pre).sigmoid()
    da_ins_c = da_ins_head(ins_c, P, pre, hooks).sigmoid()
    l_img = da_img_loss(da_img, is_source_img)
    l_ins = da_ins_loss(da_ins, dom)
    l_cst = consistency_loss(da_img_c, da_ins_c, dom)
    out = {}
    if D.DA_IMG_LOSS_WEIGHT > 0:
        out["loss_da_image"] = D.DA_IMG_LOSS_WEIGHT * l_img
    if D.DA_INS_LOSS_WEIGHT > 0:
        out["loss_da_instance"] = D.DA_INS_LOSS_WEIGHT * l_ins
    if D.DA_CST_LOSS_WEIGHT > 0:
        out["loss_da_consistency"] = D.DA_CST_LOSS_WEIGHT * l_cst
    return out


ADV_BCE = float(F.binary_cross_entropy_with_logits(torch.tensor([[0.7, 0.3]]), torch.tensor([[1.0, 0.0]])))


def adv_grl_weight(current_loss, lam, lam_adv, threshold):
    """Adv_GRL (da_heads.py:173-195), evident intent: if L <= BCE([.7,.3],[1,0])
    the GRL weight is -lam_adv * min(T, 1/L), else the plain -lam."""
    L = float(current_loss)
    if L <= ADV_BCE:
        return -1.0 * lam_adv * float(min(torch.tensor(float(threshold)), 1.0 / torch.tensor(L)))
    return -1.0 * lam


class TripletState(object):
    """Host-side state of DomainAdaptationModule_triplet / DALossComputation_Component:
    previous triplet losses (da_heads.py:107-110,320,325) and adaptive margins (loss.py:128-130)."""

    def __init__(self):
        self.prev_img, self.prev_ins = 1, 1
        self.margin_img, self.margin_ins = 0.0, 0.0


def triplet_margin_loss(a, p, n, margin):
    """nn.TripletMarginLoss(margin, p=2) (loss.py:198-200): pairwise_distance over the LAST
    dim with eps=1e-6 added to the difference, mean over the remaining elements."""
    dp = torch.sqrt(((a - p + 1e-6) ** 2).sum(-1))
    dn = torch.sqrt(((a - n + 1e-6) ** 2).sum(-1))
    return torch.clamp(margin + dp - dn, min=0).mean()


def _adaptive_margin(cur, prev_loss, adaptive, lr, max_margin, margin):
    """triplet_img_loss / triplet_ins_loss margin bookkeeping (loss.py:180-196, 202-218)."""
    if cur == 0.0:
        cur = margin
    if adaptive:
        if float(prev_loss) == 0.0 and int(cur) != int(max_margin):
            cur = cur + lr
    else:
        cur = margin
    return cur


def da_heads_triplet(feat2, da_ins_feas, dom, ins_set, img_set, is_source_img2, P, cfg, hooks, state,
                     pre="da_heads_triplet"):
    """DomainAdaptationModule_triplet.forward (da_heads.py:293-344), terms in reference order."""
    D = cfg.MODEL.DA_HEADS
    out = {}
    if D.DA_TRIPLET_INS_WEIGHT > 0:                                   # Domainlevel_Ins_component :251-274
        s, p, n = [F.avg_pool2d(t, 7, 7).view(t.size(0), -1) for t in ins_set]
        state.margin_ins = _adaptive_margin(state.margin_ins, state.prev_ins, False, 0.001,
                                            D.TRIPLET_MAX_MARGIN, D.TRIPLET_MARGIN_INS)
        l = triplet_margin_loss(s, p, n, state.margin_ins)
        out["triplet_loss_instance"] = D.DA_TRIPLET_INS_WEIGHT * l
        state.prev_ins = float(l.detach())
    if D.DA_TRIPLET_IMG_WEIGHT > 0:                                   # Domainlevel_Img_component :236-249
        state.margin_img = _adaptive_margin(state.margin_img, state.prev_img, True, 0.001,
                                            D.TRIPLET_MAX_MARGIN, D.TRIPLET_MARGIN_IMG)
        l = triplet_margin_loss(img_set[0], img_set[1], img_set[2], state.margin_img)
        out["triplet_loss_image"] = D.DA_TRIPLET_IMG_WEIGHT * l
        state.prev_img = float(l.detach())
    if D.DA_IMG_LOSS_WEIGHT > 0:                                      # DA_Img_component :125-143
        cur = da_img_loss(da_img_head(feat2, P, pre).detach(), is_source_img2)
        w = adv_grl_weight(cur, D.DA_IMG_GRL_WEIGHT, D.DA_IMG_advGRL_WEIGHT, D.DA_ADV_GRL_THRESHOLD) \
            if D.DA_ADV_GRL else -1.0 * D.DA_IMG_GRL_WEIGHT
        out["loss_da_image"] = D.DA_IMG_LOSS_WEIGHT * da_img_loss(da_img_head(grl(feat2, w), P, pre), is_source_img2)
    v = None
    if D.DA_INS_LOSS_WEIGHT > 0:                                      # DA_Ins_component :147-169
        v = F.avg_pool2d(da_ins_feas, 7, 7).view(da_ins_feas.size(0), -1)
        cur = da_ins_loss(da_ins_head(v.detach(), P, pre, hooks), dom)
        w = adv_grl_weight(cur, D.DA_INS_GRL_WEIGHT, D.DA_INS_advGRL_WEIGHT, D.DA_ADV_GRL_THRESHOLD) \
            if D.DA_ADV_GRL else -1.0 * D.DA_INS_GRL_WEIGHT
        out["loss_da_instance"] = D.DA_INS_LOSS_WEIGHT * da_ins_loss(da_ins_head(grl(v, w), P, pre, hooks), dom)
    if D.DA_CST_LOSS_WEIGHT > 0:                                      # Consistency_component :276-291
        v = F.avg_pool2d(da_ins_feas, 7, 7).view(da_ins_feas.size(0), -1)
        img_c = da_img_head(grl(feat2, 1.0 * D.DA_IMG_GRL_WEIGHT), P, pre).sigmoid()
        ins_c = da_ins_head(grl(v, 1.0 * D.DA_INS_GRL_WEIGHT), P, pre, hooks).sigmoid()
        out["loss_da_consistency"] = D.DA_CST_LOSS_WEIGHT * consistency_loss(img_c, ins_c, dom)
    return out


# --------------------------------------------------------------------------- detector
def forward_train(P, cfg, images, targets, hooks=None, triplet_state=None, nms_strict=True, aux=None,
                  image_sizes=None):
    """GeneralizedRCNN.forward, training (generalized_rcnn.py:61-156).

    images  : float32 [N,3,H,W] already mean-subtracted/padded (ImageList.tensors); without `image_sizes` all
              images are taken to be H x W (the synthetic batches of SURVEY §8d are unpadded)
    image_sizes : optional list of un-padded (h, w) per image (ImageList.image_sizes): proposals are clipped to
              and anchor visibility is evaluated against each image's OWN size (anchor_generator.py:113-125
              builds one anchor BoxList per image with that image's size; rpn/inference.py:101-103)
    targets : list of dict(boxes f32 [M,4] xyxy, labels int64 [M], is_source bool) in
              [source..., target...(, aux...)] order
    Returns the reference's loss dict (same keys) of 0-d tensors attached to P's autograd graph.
    """
    hooks = hooks or Hooks()
    aux = aux if aux is not None else {}
    n, _, ih, iw = images.shape
    R = cfg.MODEL.RPN
    gt_boxes = [t["boxes"] for t in targets]
    gt_labels = [t["labels"] for t in targets]
    is_src = [bool(t["is_source"]) for t in targets]
    image_sizes = [(ih, iw)] * n if image_sizes is None else [(int(h), int(w)) for h, w in image_sizes]

    feat = backbone_c4(images, P, cfg.MODEL.BACKBONE.CONV_BODY)
    logits, deltas = rpn_head(feat, P)
    fh, fw = feat.shape[-2:]
    cell = cell_anchors(R.ANCHOR_STRIDE[0], R.ANCHOR_SIZES, R.ASPECT_RATIOS)
    anchors = grid_anchors(fh, fw, R.ANCHOR_STRIDE[0], cell)
    vis = [anchor_visibility(anchors, w_i, h_i, R.STRADDLE_THRESH) for h_i, w_i in image_sizes]
    if len(set(image_sizes)) == 1:
        vis = vis[0]
    with torch.no_grad():
        props = rpn_proposals(anchors, logits, deltas, image_sizes, R.PRE_NMS_TOP_N_TRAIN,
                              R.POST_NMS_TOP_N_TRAIN, R.NMS_THRESH, R.MIN_SIZE, nms_strict)
        # add_gt_proposals: source images only (rpn/inference.py:51-74)
        props = [(torch.cat([b, g]), torch.cat([s, torch.ones(len(g))])) if src else (b, s)
                 for (b, s), g, src in zip(props, gt_boxes, is_src)]
    loss_obj, loss_rpn_box = rpn_loss(anchors, vis, logits, deltas, gt_boxes, is_src, cfg, hooks, aux)
    aux.update(features=feat, anchors=anchors, visibility=vis, objectness=logits, rpn_box_regression=deltas,
               proposals=props)

    D = cfg.MODEL.DA_HEADS
    losses = {}
    if D.TRIPLET_USE:                                                  # generalized_rcnn.py:88-122
        state = triplet_state or TripletState()
        f2 = feat[0:2]
        det, da_feas, dom = roi_box_head_train(f2, props[0:2], gt_boxes[0:2], gt_labels[0:2], is_src[0:2],
                                               P, cfg, hooks, aux)
        img_set = [feat[0:1], feat[1:2], feat[2:3]]
        ins_set = [0, 0, 0]
        if D.ALIGNMENT:                                                # :109-114: all with proposals[1]
            ins_set = []
            for i in range(3):
                _, fe, _ = roi_box_head_train(feat[i:i + 1], [props[1]], [gt_boxes[i]], [gt_labels[i]], [is_src[i]],
                                              P, cfg, hooks, aux, tag="aligned{}_".format(i))
                ins_set.append(fe)
        da = da_heads_triplet(f2, da_feas, dom, ins_set, img_set, is_src[0:2], P, cfg, hooks, state)
    else:                                                              # :124-128
        det, da_feas, dom = roi_box_head_train(feat, props, gt_boxes, gt_labels, is_src, P, cfg, hooks, aux)
        da = da_heads_original(feat, da_feas, dom, is_src, P, cfg, hooks)
    aux.update(da_ins_labels=dom)
    losses.update(det)
    losses.update(loss_objectness=loss_obj, loss_rpn_box_reg=loss_rpn_box)
    losses.update(da)
    return losses


# --------------------------------------------------------------------------- parameters
def param_shapes(cfg):
    """Reference state-dict names/shapes for R-50/101-C4 with DA heads."""
    shapes = {}
    nc = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES

    def bn(name, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            shapes[name + "." + s] = (c,)

    def block(name, cin, mid, cout, has_down):
        shapes[name + ".conv1.weight"] = (mid, cin, 1, 1)
        bn(name + ".bn1", mid)
        shapes[name + ".conv2.weight"] = (mid, mid, 3, 3)
        bn(name + ".bn2", mid)
        shapes[name + ".conv3.weight"] = (cout, mid, 1, 1)
        bn(name + ".bn3", cout)
        if has_down:
            shapes[name + ".downsample.0.weight"] = (cout, cin, 1, 1)
            bn(name + ".downsample.1", cout)

    shapes["backbone.body.stem.conv1.weight"] = (64, 3, 7, 7)
    bn("backbone.body.stem.bn1", 64)
    cin = 64
    for li, nb in enumerate(STAGE_BLOCKS[cfg.MODEL.BACKBONE.CONV_BODY]):
        mid, cout = 64 * 2 ** li, 256 * 2 ** li
        for b in range(nb):
            block("backbone.body.layer{}.{}".format(li + 1, b), cin, mid, cout, b == 0)
            cin = cout
    for b in range(3):
        block("roi_heads.box.feature_extractor.head.layer4.{}".format(b), cin, 512, 2048, b == 0)
        cin = 2048
    na = len(cfg.MODEL.RPN.ANCHOR_SIZES) * len(cfg.MODEL.RPN.ASPECT_RATIOS)
    shapes["rpn.anchor_generator.cell_anchors.0"] = (na, 4)
    shapes["rpn.head.conv.weight"] = (1024, 1024, 3, 3)
    shapes["rpn.head.conv.bias"] = (1024,)
    shapes["rpn.head.cls_logits.weight"] = (na, 1024, 1, 1)
    shapes["rpn.head.cls_logits.bias"] = (na,)
    shapes["rpn.head.bbox_pred.weight"] = (4 * na, 1024, 1, 1)
    shapes["rpn.head.bbox_pred.bias"] = (4 * na,)
    shapes["roi_heads.box.predictor.cls_score.weight"] = (nc, 2048)
    shapes["roi_heads.box.predictor.cls_score.bias"] = (nc,)
    shapes["roi_heads.box.predictor.bbox_pred.weight"] = (4 * nc, 2048)
    shapes["roi_heads.box.predictor.bbox_pred.bias"] = (4 * nc,)
    pres = ["da_heads"] + (["da_heads_triplet"] if cfg.MODEL.DA_HEADS.TRIPLET_USE else [])
    if cfg.MODEL.DOMAIN_ADAPTATION_ON:
        for pre in pres:
            shapes[pre + ".imghead.conv1_da.weight"] = (512, 1024, 1, 1)
            shapes[pre + ".imghead.conv1_da.bias"] = (512,)
            shapes[pre + ".imghead.conv2_da.weight"] = (1, 512, 1, 1)
            shapes[pre + ".imghead.conv2_da.bias"] = (1,)
            shapes[pre + ".inshead.fc1_da.weight"] = (1024, 2048)
            shapes[pre + ".inshead.fc1_da.bias"] = (1024,)
            shapes[pre + ".inshead.fc2_da.weight"] = (1024, 1024)
            shapes[pre + ".inshead.fc2_da.bias"] = (1024,)
            shapes[pre + ".inshead.fc3_da.weight"] = (1, 1024)
            shapes[pre + ".inshead.fc3_da.bias"] = (1,)
    return shapes


def is_trainable(name):
    """Frozen: stem + res2 (FREEZE_CONV_BODY_AT=2, resnet.py:127-136) and every
    FrozenBatchNorm2d buffer; cell_anchors is a buffer."""
    if ".bn" in name or ".downsample.1." in name or "cell_anchors" in name:
        return False
    if name.startswith("backbone.body.stem") or name.startswith("backbone.body.layer1."):
        return False
    return True
