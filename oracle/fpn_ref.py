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
