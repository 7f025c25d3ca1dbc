"""TEST INFRASTRUCTURE ONLY (oracle) — CPU restatement of the reference's FPN detector path: eval mode (pinned to the
real reference model) and plain Faster R-CNN training (restated from the reference's loss classes; the reference cannot
run it, see forward_train_fpn) (BASELINE configs[4], R-101-FPN).  Functional, fp32, NCHW, on top of the helpers of
oracle/da_frcnn_ref.py.  Only tests/ may import this file; the product path never does.

Follows (reference file:line):
  * modeling/backbone/resnet.py:41-77,138-145      ResNet body with return_features on every stage (C2..C5)
  * modeling/backbone/fpn.py:43-85                 FPN.forward: laterals, nearest 2x top-down path, LastLevelMaxPool
  * modeling/backbone/backbone.py:21-43            build_resnet_fpn_backbone (channel lists)
  * modeling/rpn/anchor_generator.py:47-66,73-125  one anchor size per level, strides (4, 8, 16, 32, 64)
  * modeling/rpn/rpn.py:39-46                      the SAME RPNHead applied to every level
  * modeling/rpn/inference.py:76-181               per-level post-processing, cat over levels, select_over_all_levels
  * modeling/poolers.py:11-42,104-121              LevelMapper, multi-level Pooler.forward
  * roi_heads/box_head/roi_box_feature_extractors.py:48-79   FPN2MLPFeatureExtractor
  * roi_heads/box_head/roi_box_predictors.py:37-57           FPNPredictor
  * roi_heads/box_head/inference.py:43-150                   PostProcessor (softmax, decode, clip, per-class NMS, top-k)

Pinned by tests/test_fpn_cpu.py against outputs of the REAL reference R-101-FPN model run on CPU
(tests/golden/eval_faster_rcnn_r101_fpn.pt, made by `python oracle/make_golden.py fpn`): pyramid probes, RPN
proposals, per-ROI pyramid levels and final detections.
"""
import torch
import torch.nn.functional as F

import da_frcnn_ref as orc

FPN_STAGE_BLOCKS = {"R-50-FPN": (3, 4, 6, 3), "R-101-FPN": (3, 4, 23, 3), "R-152-FPN": (3, 8, 36, 3)}


def resnet_body_all_stages(images, P, conv_body):
    """BaseStem + the four stages; every stage output is returned (resnet.py:138-145)."""
    x = F.conv2d(images, P["backbone.body.stem.conv1.weight"], stride=2, padding=3)
    x = F.relu(orc.frozen_bn(x, P, "backbone.body.stem.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nb in enumerate(FPN_STAGE_BLOCKS[conv_body]):
        x = orc.stage(x, P, "backbone.body.layer{}".format(li + 1), nb, 1 if li == 0 else 2)
        outs.append(x)
    return outs


def fpn_forward(feats, P):
    """FPN.forward (fpn.py:43-74) with LastLevelMaxPool (:80-82): returns [P2, P3, P4, P5, P6]."""
    def conv(name, x, pad):
        return F.conv2d(x, P["backbone.fpn.%s.weight" % name], P["backbone.fpn.%s.bias" % name], padding=pad)

    n = len(feats)
    last_inner = conv("fpn_inner%d" % n, feats[-1], 0)
    results = [conv("fpn_layer%d" % n, last_inner, 1)]
    for i in range(n - 2, -1, -1):
        top_down = F.interpolate(last_inner, scale_factor=2, mode="nearest")
        last_inner = conv("fpn_inner%d" % (i + 1), feats[i], 0) + top_down
        results.insert(0, conv("fpn_layer%d" % (i + 1), last_inner, 1))
    results.append(F.max_pool2d(results[-1], 1, 2, 0))
    return results


def select_over_all_levels(per_image, fpn_post_nms_top_n, training):
    """RPNPostProcessor.select_over_all_levels (rpn/inference.py:154-181).  per_image: list of (boxes, objectness)
    already concatenated over the levels.  Training: ONE top-k over the whole batch, original order kept; test: a
    sorted top-k per image."""
    if training:
        allsc = torch.cat([s for _, s in per_image])
        k = min(fpn_post_nms_top_n, allsc.numel())
        _, top = torch.topk(allsc, k, dim=0, sorted=True)
        mask = torch.zeros_like(allsc, dtype=torch.bool)
        mask[top] = True
        out = []
        for (b, s), m in zip(per_image, mask.split([len(s) for _, s in per_image])):
            out.append((b[m], s[m]))
        return out
    out = []
    for b, s in per_image:
        _, top = torch.topk(s, min(fpn_post_nms_top_n, s.numel()), dim=0, sorted=True)
        out.append((b[top], s[top]))
    return out


def rpn_fpn_proposals(pyramid, P, cfg, image_sizes, training=False, nms_strict=True):
    """RPNModule.forward in test mode over five levels (rpn.py:88-140, inference.py:126-152)."""
    R = cfg.MODEL.RPN
    pre = R.PRE_NMS_TOP_N_TRAIN if training else R.PRE_NMS_TOP_N_TEST
    post = R.POST_NMS_TOP_N_TRAIN if training else R.POST_NMS_TOP_N_TEST
    fpn_post = R.FPN_POST_NMS_TOP_N_TRAIN if training else R.FPN_POST_NMS_TOP_N_TEST
    n = pyramid[0].shape[0]
    per_image = [[] for _ in range(n)]
    for feat, stride, size in zip(pyramid, R.ANCHOR_STRIDE, R.ANCHOR_SIZES):
        logits, deltas = orc.rpn_head(feat, P)
        cell = orc.cell_anchors(stride, (size,), R.ASPECT_RATIOS)
        anchors = orc.grid_anchors(feat.shape[2], feat.shape[3], stride, cell)
        lvl = orc.rpn_proposals(anchors, logits, deltas, image_sizes, pre, post, R.NMS_THRESH, R.MIN_SIZE,
                                nms_strict=nms_strict)
        for i in range(n):
            per_image[i].append(lvl[i])
    cat = [(torch.cat([b for b, _ in lv]), torch.cat([s for _, s in lv])) for lv in per_image]
    return select_over_all_levels(cat, fpn_post, training)


def level_mapper(boxes, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
    """LevelMapper.__call__ (poolers.py:34-42); areas with the +1 convention (bounding_box.py:227-230)."""
    area = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
    s = torch.sqrt(area)
    lvls = torch.floor(canonical_level + torch.log2(s / canonical_scale + eps))
    return torch.clamp(lvls, min=k_min, max=k_max).to(torch.int64) - int(k_min)


def multilevel_pool(pyramid, rois, scales, resolution, sampling_ratio):
    """Pooler.forward (poolers.py:104-121): rois [K,5]; returns ([K,C,r,r], levels)."""
    k_min = -torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item()
    k_max = -torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item()
    levels = level_mapper(rois[:, 1:], k_min, k_max)
    out = torch.zeros((rois.shape[0], pyramid[0].shape[1], resolution, resolution), dtype=pyramid[0].dtype)
    for lvl, (feat, scale) in enumerate(zip(pyramid, scales)):
        idx = torch.nonzero(levels == lvl).squeeze(1)
        if idx.numel():
            out[idx] = orc.roi_align(feat, rois[idx], scale, resolution, resolution, sampling_ratio)
    return out, levels


def fpn2mlp_head(x, P):
    """FPN2MLPFeatureExtractor.forward after pooling (roi_box_feature_extractors.py:71-79) + FPNPredictor (:50-57)."""
    pre = "roi_heads.box.feature_extractor."
    x = x.reshape(x.shape[0], -1)
    x = F.relu(F.linear(x, P[pre + "fc6.weight"], P[pre + "fc6.bias"]))
    x = F.relu(F.linear(x, P[pre + "fc7.weight"], P[pre + "fc7.bias"]))
    pp = "roi_heads.box.predictor."
    return (F.linear(x, P[pp + "cls_score.weight"], P[pp + "cls_score.bias"]),
            F.linear(x, P[pp + "bbox_pred.weight"], P[pp + "bbox_pred.bias"]))


def box_postprocess(cfg, class_logits, box_regression, proposals, image_sizes, nms_strict=True):
    """PostProcessor.forward + filter_results (box_head/inference.py:43-150).  proposals: list of boxes [Pi,4]."""
    H = cfg.MODEL.ROI_HEADS
    prob = F.softmax(class_logits, -1)
    sizes = [len(p) for p in proposals]
    decoded = orc.box_decode(box_regression, torch.cat(proposals), H.BBOX_REG_WEIGHTS)
    nc = prob.shape[1]
    results = []
    for pr, bx, (ih, iw) in zip(prob.split(sizes), decoded.split(sizes), image_sizes):
        bx = orc.clip_boxes(bx.reshape(-1, 4), iw, ih).reshape(-1, nc * 4)
        boxes, scores, labels = [], [], []
        for j in range(1, nc):
            inds = torch.nonzero(pr[:, j] > H.SCORE_THRESH).squeeze(1)
            sj, bj = pr[inds, j], bx[inds, j * 4:(j + 1) * 4]
            keep = orc.nms(bj, sj, H.NMS, strict=nms_strict) if len(inds) else inds
            boxes.append(bj[keep])
            scores.append(sj[keep])
            labels.append(torch.full((len(keep),), j, dtype=torch.int64))
        b, s, l = torch.cat(boxes), torch.cat(scores), torch.cat(labels)
        if len(s) > H.DETECTIONS_PER_IMG > 0:
            thr, _ = torch.kthvalue(s, len(s) - H.DETECTIONS_PER_IMG + 1)
            keep = torch.nonzero(s >= thr.item()).squeeze(1)
            b, s, l = b[keep], s[keep], l[keep]
        results.append(dict(boxes=b, scores=s, labels=l))
    return results


def forward_eval_fpn(P, cfg, images, nms_strict=True):
    """GeneralizedRCNN.forward in eval mode for an R-*-FPN model (generalized_rcnn.py:61-156 with
    targets=None).  Returns dict(pyramid, proposals, levels, detections)."""
    n, _, ih, iw = images.shape
    image_sizes = [(ih, iw)] * n
    pyramid = fpn_forward(resnet_body_all_stages(images, P, cfg.MODEL.BACKBONE.CONV_BODY), P)
    props = rpn_fpn_proposals(pyramid, P, cfg, image_sizes, training=False, nms_strict=nms_strict)
    rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), b], dim=1) for i, (b, _) in enumerate(props)])
    B = cfg.MODEL.ROI_BOX_HEAD
    pooled, levels = multilevel_pool(pyramid[:len(B.POOLER_SCALES)], rois, B.POOLER_SCALES, B.POOLER_RESOLUTION,
                                     B.POOLER_SAMPLING_RATIO)
    logits, reg = fpn2mlp_head(pooled, P)
    dets = box_postprocess(cfg, logits, reg, [b for b, _ in props], image_sizes, nms_strict=nms_strict)
    return dict(pyramid=pyramid, proposals=props, levels=levels, detections=dets)


# --------------------------------------------------------------------------------------------------- training (no DA)
def rpn_loss_fpn(pyramid, heads, cfg, gt_boxes, is_source_img, image_size, hooks):
    """RPNLossComputation over five levels (rpn/loss.py:57-143): per image the anchors of all levels are
    concatenated (cat_boxlist, :119), the predictions are flattened level by level per image
    (concat_box_prediction_layers, rpn/utils.py:17-45)."""
    R = cfg.MODEL.RPN
    ih, iw = image_size
    anchors, vis = [], []
    for feat, stride, size in zip(pyramid, R.ANCHOR_STRIDE, R.ANCHOR_SIZES):
        a = orc.grid_anchors(feat.shape[2], feat.shape[3], stride, orc.cell_anchors(stride, (size,), R.ASPECT_RATIOS))
        anchors.append(a)
        vis.append(orc.anchor_visibility(a, iw, ih, R.STRADDLE_THRESH))
    anchors, vis = torch.cat(anchors), torch.cat(vis)
    labels, reg_targets = [], []
    for gt, src in zip(gt_boxes, is_source_img):
        if not src:
            continue
        m = orc.matcher(orc.box_iou(gt, anchors), R.FG_IOU_THRESHOLD, R.BG_IOU_THRESHOLD, True)
        lab = (m >= 0).to(torch.float32)
        lab[m == orc.BELOW_LOW] = 0
        lab[~vis] = -1
        lab[m == orc.BETWEEN] = -1
        labels.append(lab)
        reg_targets.append(orc.box_encode(gt[m.clamp(min=0)], anchors, (1.0, 1.0, 1.0, 1.0)))
    pos_m, neg_m = orc.balanced_sampler(labels, R.BATCH_SIZE_PER_IMAGE, R.POSITIVE_FRACTION, hooks)
    pos = torch.nonzero(torch.cat(pos_m)).squeeze(1)
    neg = torch.nonzero(torch.cat(neg_m)).squeeze(1)
    sampled = torch.cat([pos, neg])
    n = pyramid[0].shape[0]
    obj = torch.cat([orc.permute_and_flatten(lg, n, 1, lg.shape[2], lg.shape[3]) for lg, _ in heads], dim=1).reshape(-1)
    reg = torch.cat([orc.permute_and_flatten(dl, n, 4, dl.shape[2], dl.shape[3]) for _, dl in heads], dim=1).reshape(-1, 4)
    labels, reg_targets = torch.cat(labels), torch.cat(reg_targets)
    box_loss = orc.smooth_l1(reg[pos], reg_targets[pos], 1.0 / 9, False) / sampled.numel()
    obj_loss = F.binary_cross_entropy_with_logits(obj[sampled], labels[sampled])
    return obj_loss, box_loss


def forward_train_fpn(P, cfg, images, targets, hooks):
    """GeneralizedRCNN.forward in training mode for an R-*-FPN model WITHOUT DA heads.  The reference leaves
    `detector_losses` unbound in this configuration; the evident intent — the four Faster R-CNN losses
    — is restated from the same loss classes the DA path uses (rpn/loss.py, box_head/loss.py), with the FPN-specific
    proposal selection of rpn/inference.py:154-181 (training branch).  Returns the loss dict."""
    n, _, ih, iw = images.shape
    gt_boxes = [t["boxes"] for t in targets]
    gt_labels = [t["labels"] for t in targets]
    is_src = [bool(t["is_source"]) for t in targets]
    pyramid = fpn_forward(resnet_body_all_stages(images, P, cfg.MODEL.BACKBONE.CONV_BODY), P)
    heads = [orc.rpn_head(f, P) for f in pyramid]
    with torch.no_grad():
        props = rpn_fpn_proposals([f.detach() for f in pyramid], {k: v.detach() for k, v in P.items()}, cfg,
                                  [(ih, iw)] * n, training=True, nms_strict=True)
        props = [(torch.cat([b, g]), torch.cat([s, torch.ones(len(g))])) if src else (b, s)      # add_gt_proposals
                 for (b, s), g, src in zip(props, gt_boxes, is_src)]
    obj_loss, rpn_box_loss = rpn_loss_fpn(pyramid, heads, cfg, gt_boxes, is_src, (ih, iw), hooks)
    samples = orc.box_head_subsample(props, gt_boxes, gt_labels, is_src, cfg, hooks)
    B = cfg.MODEL.ROI_BOX_HEAD
    pooled, _ = multilevel_pool(pyramid[:len(B.POOLER_SCALES)], orc.rois_from(samples), B.POOLER_SCALES,
                                B.POOLER_RESOLUTION, B.POOLER_SAMPLING_RATIO)
    cls_logits, box_reg = fpn2mlp_head(pooled, P)
    cls_loss, box_loss, _ = orc.fastrcnn_loss(cls_logits, box_reg, samples)
    return dict(loss_classifier=cls_loss, loss_box_reg=box_loss, loss_objectness=obj_loss,
                loss_rpn_box_reg=rpn_box_loss)


# ----------------------------------------------------------------------------------- FPN + DA (BASELINE configs[4])
# PARITY UNPINNED.  The reference ships no runnable FPN + DA combination: GeneralizedRCNN imports da_heads.py, whose
# heads are sized for C4 (da_heads.py:368-370), and the FPN variant da_heads_fpn.py cannot even be
# imported (it imports names loss.py does not define, :12; reads MODEL.DA_HEADS.COS_WEIGHT, :214, which
# config/defaults.py does not have; its forward signature, :236, is not the one generalized_rcnn.py:126 calls; its
# DAInsHead.forward returns from inside the level loop, :205; and DALossComputation concatenates the per-level image
# logits along dim 0, loss.py:82, which fails for maps of different sizes).  What follows restates the INTENT of
# da_heads_fpn.py with each of those defects resolved in the most conservative way, and says so line by line:
#   * image head     one (conv 1x1 256->512, ReLU, conv 1x1 512->1) pair PER pyramid level (:37-66), on the GRL'd map
#   * instance head  one (fc 1024->1024, ReLU, dropout, fc, ReLU, dropout, fc ->1) triple per POOLER level (:146-207),
#                    every ROI through the head of ITS level (LevelMapper over the sampled proposals, :222-225,254) —
#                    all four levels, i.e. without the stray `return` of :205
#   * GRL weights    -DA_IMG_GRL_WEIGHT / -DA_INS_GRL_WEIGHT, consistency branches +1.0 x the same (COS_WEIGHT := 1.0,
#                    the value da_heads.py:381-382 hard-codes for C4)
#   * image loss     BCE-with-logits against the image's domain label, mean over ALL pixels of ALL levels (the per-level
#                    [N, H_l*W_l] logits concatenated along dim 1 instead of dim 0: identical for one level)
#   * instance loss  BCE-with-logits against the ROI's domain (loss.py:98-100)
#   * consistency    layers/consistency_loss.py:3-27 as written — it already takes a list of levels: |mean prob of the
#                    ROI's image at level l - ROI prob|, mean over ROIs x levels
#   * loss weights   DA_IMG/INS/CST_LOSS_WEIGHT applied and zero-weight terms dropped as da_heads.py:417-436 does.
def da_img_heads_fpn(pyramid, P, pre="da_heads"):
    outs = []
    for i, f in enumerate(pyramid):
        t = F.relu(F.conv2d(f, P["%s.imghead.da_img_conv1_level%d.weight" % (pre, i)],
                            P["%s.imghead.da_img_conv1_level%d.bias" % (pre, i)]))
        outs.append(F.conv2d(t, P["%s.imghead.da_img_conv2_level%d.weight" % (pre, i)],
                             P["%s.imghead.da_img_conv2_level%d.bias" % (pre, i)]))
    return outs


def da_ins_heads_fpn(x, levels, P, hooks, n_levels, pre="da_heads"):
    """x [K, 1024]; ROI k goes through the FC triple of levels[k] (dropout draws in level order, only for levels
    that hold ROIs, like the reference's `if len(idx_in_level) > 0`, :194)."""
    result = torch.zeros((x.shape[0], 1), dtype=x.dtype)
    for lvl in range(n_levels):
        idx = torch.nonzero(levels == lvl).squeeze(1)
        if idx.numel() == 0:
            continue
        xs = x[idx]
        for j in (1, 2):
            w, b = P["%s.inshead.da_ins_fc%d_level%d.weight" % (pre, j, lvl)], P["%s.inshead.da_ins_fc%d_level%d.bias" % (pre, j, lvl)]
            xs = F.relu(F.linear(xs, w, b))
            xs = xs * hooks.dropout_keep(tuple(xs.shape)) * 2.0                  # F.dropout(p=0.5, training=True)
        out = F.linear(xs, P["%s.inshead.da_ins_fc3_level%d.weight" % (pre, lvl)], P["%s.inshead.da_ins_fc3_level%d.bias" % (pre, lvl)])
        result = result.index_put((idx,), out)
    return result


def da_heads_fpn(pyramid, ins_feas, dom, levels, is_source_img, P, cfg, hooks, pre="da_heads"):
    D = cfg.MODEL.DA_HEADS
    n_ins_levels = len(cfg.MODEL.ROI_BOX_HEAD.POOLER_SCALES)
    img_g = [orc.grl(f, -1.0 * D.DA_IMG_GRL_WEIGHT) for f in pyramid]
    ins_g = orc.grl(ins_feas, -1.0 * D.DA_INS_GRL_WEIGHT)
    img_c = [orc.grl(f, 1.0 * D.DA_IMG_GRL_WEIGHT) for f in pyramid]
    ins_c = orc.grl(ins_feas, 1.0 * D.DA_INS_GRL_WEIGHT)
    da_ins = da_ins_heads_fpn(ins_g, levels, P, hooks, n_ins_levels, pre)
    da_ins_c = da_ins_heads_fpn(ins_c, levels, P, hooks, n_ins_levels, pre).sigmoid()
    da_img = da_img_heads_fpn(img_g, P, pre)
    da_img_c = [t.sigmoid() for t in da_img_heads_fpn(img_c, P, pre)]
    n = pyramid[0].shape[0]
    flat = torch.cat([t.permute(0, 2, 3, 1).reshape(n, -1) for t in da_img], dim=1)
    lab = torch.zeros_like(flat)
    lab[torch.tensor(is_source_img, dtype=torch.bool), :] = 1
    l_img = F.binary_cross_entropy_with_logits(flat, lab)
    l_ins = orc.da_ins_loss(da_ins, dom)
    # consistency_loss.py:3-27 over the list of levels: [K, L] absolute differences, mean
    k = da_ins_c.size(0)
    n_src = int(torch.nonzero(dom).size(0))
    assert n == 2, "only batch size=2 is supported for consistency loss now, received batch size: {}".format(n)
    cols = []
    for t in da_img_c:
        means = t.reshape(n, -1).mean(1)
        per_roi = torch.cat([means[0].view(1, 1).repeat(n_src, 1), means[1].view(1, 1).repeat(k - n_src, 1)], dim=0)
        cols.append(torch.abs(per_roi - da_ins_c))
    l_cst = torch.cat(cols, dim=1).mean()
    out = {}
    if D.DA_IMG_LOSS_WEIGHT > 0:
        out["loss_da_image"] = D.DA_IMG_LOSS_WEIGHT * l_img
    if D.DA_INS_LOSS_WEIGHT > 0:
        out["loss_da_instance"] = D.DA_INS_LOSS_WEIGHT * l_ins
    if D.DA_CST_LOSS_WEIGHT > 0:
        out["loss_da_consistency"] = D.DA_CST_LOSS_WEIGHT * l_cst
    return out


def forward_train_fpn_da(P, cfg, images, targets, hooks):
    """GeneralizedRCNN.forward, training, R-*-FPN WITH the DA heads (the configuration BASELINE configs[4] names):
    forward_train_fpn's detector losses — RPN labels for source images only, detection losses masked to source ROIs
    (rpn/loss.py:66-67, box_head/loss.py:84-85,200-219, as on the C4 path) — plus da_heads_fpn on the five pyramid
    maps and the [K, 1024] MLP features of the sampled ROIs (generalized_rcnn.py:124-128).  Parity unpinned (above)."""
    n, _, ih, iw = images.shape
    gt_boxes = [t["boxes"] for t in targets]
    gt_labels = [t["labels"] for t in targets]
    is_src = [bool(t["is_source"]) for t in targets]
    pyramid = fpn_forward(resnet_body_all_stages(images, P, cfg.MODEL.BACKBONE.CONV_BODY), P)
    heads = [orc.rpn_head(f, P) for f in pyramid]
    with torch.no_grad():
        props = rpn_fpn_proposals([f.detach() for f in pyramid], {k: v.detach() for k, v in P.items()}, cfg,
                                  [(ih, iw)] * n, training=True, nms_strict=True)
        props = [(torch.cat([b, g]), torch.cat([s, torch.ones(len(g))])) if src else (b, s)
                 for (b, s), g, src in zip(props, gt_boxes, is_src)]
    obj_loss, rpn_box_loss = rpn_loss_fpn(pyramid, heads, cfg, gt_boxes, is_src, (ih, iw), hooks)
    samples = orc.box_head_subsample(props, gt_boxes, gt_labels, is_src, cfg, hooks)
    B = cfg.MODEL.ROI_BOX_HEAD
    rois = orc.rois_from(samples)
    pooled, levels = multilevel_pool(pyramid[:len(B.POOLER_SCALES)], rois, B.POOLER_SCALES, B.POOLER_RESOLUTION,
                                     B.POOLER_SAMPLING_RATIO)
    pre = "roi_heads.box.feature_extractor."
    x = pooled.reshape(pooled.shape[0], -1)
    x = F.relu(F.linear(x, P[pre + "fc6.weight"], P[pre + "fc6.bias"]))
    x = F.relu(F.linear(x, P[pre + "fc7.weight"], P[pre + "fc7.bias"]))
    pp = "roi_heads.box.predictor."
    cls_logits = F.linear(x, P[pp + "cls_score.weight"], P[pp + "cls_score.bias"])
    box_reg = F.linear(x, P[pp + "bbox_pred.weight"], P[pp + "bbox_pred.bias"])
    cls_loss, box_loss, dom = orc.fastrcnn_loss(cls_logits, box_reg, samples)
    losses = dict(loss_classifier=cls_loss, loss_box_reg=box_loss, loss_objectness=obj_loss,
                  loss_rpn_box_reg=rpn_box_loss)
    losses.update(da_heads_fpn(pyramid, x, dom, levels, is_src, P, cfg, hooks))
    return losses
