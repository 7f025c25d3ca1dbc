"""TEST INFRASTRUCTURE ONLY — torch-CPU stand-ins for the kernel-calling functions of dadetect_b200.ops, installed
by the `cpu_ops` fixture (monkeypatch, undone after the test) so that the PYTHON control flow of the model code
(FPN pyramid wiring, multi-level RPN selection, level mapping, head plumbing, post-processing) can be exercised in
the `-m "not gpu"` suite.  Nothing here ships: the product has no CPU path (ops raise on CPU tensors), and the
arithmetic of every stand-in comes from torch / the oracle, never from the product."""
import pytest
import torch
import torch.nn.functional as F

import da_frcnn_ref as orc


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def conv_bn_act(x, weight, scale=None, bias=None, residual=None, stride=1, pad=0, relu=False):
    w = weight if weight.dim() == 4 else weight[:, :, None, None]
    y = F.conv2d(nchw(x), w.contiguous(), stride=stride, padding=pad)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    y = nhwc(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def linear(x, weight, bias=None, relu=False):
    y = F.linear(x, weight, bias)
    return F.relu(y) if relu else y


def fused_heads(x, weights, biases):
    return [F.linear(x, w.reshape(w.shape[0], -1), b) for w, b in zip(weights, biases)]


def bottleneck_stage(x, blocks, strides, input_is_relu=False, grad_premasked=False, pool_output=False):
    for b, s in zip(blocks, strides):
        y = conv_bn_act(x, b["w1"], b["s1"], b["b1"], stride=s, relu=True)
        y = conv_bn_act(y, b["w2"], b["s2"], b["b2"], pad=1, relu=True)
        idt = conv_bn_act(x, b["wd"], b["sd"], b["bd"], stride=s) if "wd" in b else x
        x = conv_bn_act(y, b["w3"], b["s3"], b["b3"], residual=idt, relu=True)
    return x.mean(dim=(1, 2)) if pool_output else x


def anchor_grid(cell, fh, fw, stride, img_w, img_h, straddle):
    a = orc.grid_anchors(fh, fw, stride, cell)
    return a, orc.anchor_visibility(a, img_w, img_h, straddle).to(torch.uint8)


def rpn_topk_decode(logits, deltas, anchors, k, img_w, img_h, min_size):
    n = logits.shape[0]
    obj = logits.reshape(n, -1).sigmoid()
    reg = deltas.reshape(n, -1, 4)
    sc, idx = obj.topk(k, dim=1, sorted=True)
    boxes, scores, valid = torch.zeros(n, k, 4), torch.zeros(n, k), torch.zeros(n, dtype=torch.int32)
    for i in range(n):
        bx = orc.clip_boxes(orc.box_decode(reg[i][idx[i]], anchors[idx[i]], (1.0, 1.0, 1.0, 1.0)), img_w, img_h)
        keep = ((bx[:, 2] - bx[:, 0] + 1) >= min_size) & ((bx[:, 3] - bx[:, 1] + 1) >= min_size)
        m = int(keep.sum())
        boxes[i, :m], scores[i, :m], valid[i] = bx[keep], sc[i][keep], m
    return boxes, scores, idx.to(torch.int32), valid


def nms_sorted_batched(boxes, valid, thresh, max_keep):
    n = boxes.shape[0]
    keep, cnt = torch.zeros(n, max_keep, dtype=torch.int64), torch.zeros(n, dtype=torch.int32)
    for i in range(n):
        v = int(valid[i])
        k = orc.nms(boxes[i, :v], torch.arange(v, 0, -1).float(), thresh, strict=True)[:max_keep]
        keep[i, :len(k)], cnt[i] = k, len(k)
    return keep, cnt


def roi_align_levels(feats, rois, scales, pooled, sampling_ratio):
    import fpn_ref
    out, lv = fpn_ref.multilevel_pool([nchw(f) for f in feats], rois, scales, pooled, sampling_ratio)
    return nhwc(out), lv.to(torch.int32)


def roi_align(feat, rois, scale, pooled, sampling_ratio, bin_step=1):
    return nhwc(orc.roi_align(nchw(feat), rois, scale, pooled, pooled, sampling_ratio))[:, ::bin_step, ::bin_step]


# ---- training path (host-driven control flow)
class _Grl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.w = w
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        w = ctx.w
        return g * (float(w) if not torch.is_tensor(w) else w.reshape(()).to(g.dtype)), None   # a device weight is read NOW


def match(gt, pred, high, low, allow_low_quality, m_dev=None):
    if m_dev is not None:                      # GT padded to a capacity: only the first *m_dev rows exist
        gt = gt[: int(m_dev)]
    if gt.shape[0] == 0 or pred.shape[0] == 0:
        raise ValueError("No ground-truth or proposal boxes available for one of the images during training")
    q = orc.box_iou(gt, pred)
    return orc.matcher(q, high, low, allow_low_quality), q.max(dim=0)[0]


def box_encode(gt, pred, matches, weights, wrap_negative=False, m_dev=None):
    if m_dev is not None:
        gt = gt[: int(m_dev)]
    idx = torch.where(matches < 0, matches + gt.shape[0] if wrap_negative else torch.zeros_like(matches), matches)
    return orc.box_encode(gt[idx.clamp(min=0)], pred, weights)


def bce_with_logits_mean(x, targets=None, seg_labels=None, seg_len=0):
    if targets is None:
        targets = seg_labels.to(torch.float32).repeat_interleave(int(seg_len)).view_as(x)
    return F.binary_cross_entropy_with_logits(x, targets.view_as(x))


def smooth_l1_sum(x, t, beta, divisor):
    n = (x - t).abs()
    return torch.where(n < beta, 0.5 * n * n / beta, n - 0.5 * beta).sum() / divisor


def softmax_ce_mean(logits, labels, row_mask):
    m = row_mask.bool()
    return F.cross_entropy(logits[m], labels[m])


def box_reg_loss(box_reg, reg_targets, labels, row_mask):
    m = row_mask.bool()                                            # box_head/loss.py:205-219
    pos = torch.nonzero(m & (labels > 0)).squeeze(1)
    cols = 4 * labels[pos][:, None] + torch.arange(4)
    return smooth_l1_sum(box_reg[pos[:, None], cols], reg_targets[pos], 1.0, float(m.sum()))


def consistency_loss(img_logits, ins_logits, n_src, row_valid=None):
    assert row_valid is None
    k = ins_logits.numel()
    dom = torch.arange(k) < int(n_src)
    return orc.consistency_loss(img_logits.sigmoid(), ins_logits.sigmoid().reshape(-1, 1), dom)


def triplet_margin_loss(a, p, n, margin, rows, d, inner):
    def rows_of(t):
        return t.reshape(-1, d, inner).permute(0, 2, 1).reshape(rows, d)
    return F.triplet_margin_loss(rows_of(a), rows_of(p), rows_of(n), margin=float(margin), p=2)


def adaptive_margin_update(state, prev_loss, margin_cfg, lr, max_margin, out=None):
    state.fill_(orc._adaptive_margin(float(state), 1.0 if prev_loss is None else float(prev_loss), True, lr, max_margin,
                                     margin_cfg))
    m = state.to(torch.float32)
    if out is not None:
        out.copy_(m)
        return out
    return m


def adv_grl_weight(loss, bce, lam, lam_adv, threshold, out=None):
    w = torch.tensor([orc.adv_grl_weight(loss.detach(), lam, lam_adv, threshold)], dtype=torch.float32)
    if out is not None:
        out.copy_(w)
        return out
    return w


# ---- sync-free (fixed-capacity) path
def proposals_gather(boxes, scores, keep, keep_count, gt_cat, gt_offsets, append_gt, cap, gt_counts=None):
    n, post = boxes.shape[0], keep.shape[1]
    out_b, out_s = torch.zeros(n, cap, 4), torch.zeros(n, cap)
    out_c = torch.zeros(n, dtype=torch.int32)
    for g in range(n):
        c = min(int(keep_count[g]), post)
        sel = keep[g, :c]
        out_b[g, :c], out_s[g, :c] = boxes[g][sel], scores[g][sel]
        if int(append_gt[g]):
            a, b = int(gt_offsets[g]), int(gt_offsets[g + 1])
            if gt_counts is not None:
                b = a + min(int(gt_counts[g]), b - a)
            out_b[g, c:c + b - a], out_s[g, c:c + b - a] = gt_cat[a:b], 1.0
            c += b - a
        out_c[g] = c
    return out_b, out_s, out_c


def balanced_sample(labels, n_dev, keys, batch, max_pos):
    images, n_cap = labels.shape
    sel = torch.zeros(images, batch, dtype=torch.int64)
    cnt = torch.zeros(images, 2, dtype=torch.int32)
    for g in range(images):
        n = n_cap if n_dev is None else int(n_dev[g])
        lab, k = labels[g, :n], keys[g, :n]
        pos, neg = torch.nonzero(lab >= 1).squeeze(1), torch.nonzero(lab == 0).squeeze(1)
        npos = min(pos.numel(), max_pos)
        nneg = min(neg.numel(), batch - npos)
        p = pos[torch.argsort(k[pos], stable=True)[:npos]]           # smallest keys; ties -> lower index
        q = neg[torch.argsort(k[neg], stable=True)[:nneg]]
        chosen = torch.sort(torch.cat([p, q]))[0]
        sel[g, :chosen.numel()] = chosen
        cnt[g, 0], cnt[g, 1] = npos, npos + nneg
    return sel, cnt


def consistency_loss_masked(img_logits, ins_logits, n_src, row_valid=None):
    if row_valid is None:
        return consistency_loss(img_logits, ins_logits, n_src)
    means = img_logits.sigmoid().reshape(2, -1).mean(1)
    k = ins_logits.numel()
    per_roi = torch.where(torch.arange(k) < int(n_src), means[0], means[1])
    v = row_valid.bool()
    return (torch.abs(per_roi - ins_logits.sigmoid().reshape(-1)) * v).sum() / v.sum()


def rpn_anchor_labels(matches, visibility):
    lab = (matches >= 0).to(torch.int32)
    lab[(matches == -2) | ~visibility.bool()] = -1
    return lab


def rpn_sampled_losses(logits, deltas, anchors, sel, counts, labels, matches, gt_cat, gt_offsets, src_img, beta):
    """torch restatement of rpn/loss.py:118-141 on the sampled anchors (autograd provides the gradients)."""
    S, B = sel.shape
    A = anchors.shape[0]
    obj, reg = logits.reshape(-1), deltas.reshape(-1, 4)
    total = int(counts[:, 1].sum())
    bce = l1 = 0.0
    for s in range(S):
        sl = sel[s, : int(counts[s, 1])]
        lab = labels[s][sl]
        rows = sl + s * A
        bce = bce + F.binary_cross_entropy_with_logits(obj[rows], (lab == 1).float(), reduction="sum")
        p = sl[lab == 1]
        g0 = int(gt_offsets[int(src_img[s])])
        tg = orc.box_encode(gt_cat[g0 + matches[s][p]], anchors[p], (1.0, 1.0, 1.0, 1.0))
        l1 = l1 + orc.smooth_l1(reg[p + s * A], tg, beta, size_average=False)
    return bce / total, l1 / total


def roi_labels(matches, gt_labels, is_source, n_prop, out=None):
    m = matches
    if is_source:
        lab = gt_labels.to(torch.int64)[m.clamp(min=0)]
        lab = torch.where(m == -1, torch.zeros_like(lab), lab)
        lab = torch.where(m == -2, torch.full_like(lab, -1), lab)
    else:
        lab = torch.zeros_like(m)
    lab = torch.where(torch.arange(m.numel()) < int(n_prop), lab, torch.full_like(lab, -1)).to(torch.int32)
    if out is not None:
        out.copy_(lab)
        return out
    return lab


def roi_gather_sampled(boxes, objectness, sel, counts, labels, matches, gt_cat, gt_offsets, gt_counts, is_source,
                       weights):
    """torch restatement of the per-image gathers of box_head/loss.py:100-130 (the former Python body of
    subsample_static)."""
    n_img, cap = objectness.shape
    B = sel.shape[1]
    valid = torch.arange(B).unsqueeze(0) < counts[:, 1:2]
    rois, labs, regs, doms, objs = [], [], [], [], []
    for i in range(n_img):
        src = bool(is_source[i])
        bx = boxes[i][sel[i]]
        rois.append(torch.cat([torch.full((B, 1), float(i)), bx], dim=1))
        labs.append(torch.where(valid[i], labels[i][sel[i]].to(torch.int64), torch.zeros_like(sel[i])))
        g0, g1 = int(gt_offsets[i]), int(gt_offsets[i + 1])
        live = g1 - g0 if gt_counts is None else min(int(gt_counts[i]), g1 - g0)
        m = matches[i][sel[i]]
        m = torch.where(m < 0, (m + max(live, 1)) if not src else torch.zeros_like(m), m).clamp(min=0)
        regs.append(orc.box_encode(gt_cat[g0 + m], bx, weights))
        doms.append(torch.full((B,), src, dtype=torch.bool))
        objs.append(objectness[i][sel[i]])
    return dict(rois=torch.cat(rois), labels=torch.cat(labs), regression_targets=torch.cat(regs),
                domain_labels=torch.cat(doms), valid=valid.reshape(-1), objectness=torch.cat(objs))


TRAINING_STAND_INS = dict(
    roi_labels=roi_labels, roi_gather_sampled=roi_gather_sampled,
    rpn_anchor_labels=rpn_anchor_labels, rpn_sampled_losses=rpn_sampled_losses,
    proposals_gather=proposals_gather, balanced_sample=balanced_sample,
    gradient_scalar=lambda x, w: _Grl.apply(x, w), gradient_scalar_dev=lambda x, wdev: _Grl.apply(x, wdev),
    dropout_with_mask=lambda x, keep: x * keep * 2.0, match=match, box_encode=box_encode,
    bce_with_logits_mean=bce_with_logits_mean, smooth_l1_sum=smooth_l1_sum, softmax_ce_mean=softmax_ce_mean,
    box_reg_loss=box_reg_loss, consistency_loss=consistency_loss_masked, triplet_margin_loss=triplet_margin_loss,
    adv_grl_weight=adv_grl_weight, adaptive_margin_update=adaptive_margin_update,
)

STAND_INS = dict(
    _chk=lambda t, dtype=torch.float32, name="tensor": t,
    conv_bn_act=conv_bn_act, linear=linear, bottleneck_stage=bottleneck_stage, fused_heads=fused_heads,
    stem_tc_supported=lambda x, w: False, nchw_to_nhwc=nhwc,
    maxpool3x3s2=lambda x: nhwc(F.max_pool2d(nchw(x), 3, 2, 1)),
    upsample2x=lambda x: nhwc(F.interpolate(nchw(x), scale_factor=2, mode="nearest")),
    subsample2=lambda x: nhwc(F.max_pool2d(nchw(x), 1, 2, 0)),
    avgpool_hw=lambda x: x.mean(dim=(1, 2)),
    anchor_grid=anchor_grid, rpn_topk_decode=rpn_topk_decode, nms_sorted_batched=nms_sorted_batched,
    nms=lambda boxes, scores, thresh: orc.nms(boxes, scores, thresh, strict=True),
    box_decode=lambda codes, boxes, weights: orc.box_decode(codes, boxes, weights),
    roi_align_levels=roi_align_levels, roi_align=roi_align,
)


@pytest.fixture
def cpu_ops(monkeypatch):
    from dadetect_b200 import ops
    for name, fn in list(STAND_INS.items()) + list(TRAINING_STAND_INS.items()):
        assert hasattr(ops, name), name                 # a renamed op must not silently lose its stand-in
        monkeypatch.setattr(ops, name, fn)
    return ops
