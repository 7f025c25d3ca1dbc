"""Region proposal network: head convs, on-device proposal generation, RPN losses.

Mirrors maskrcnn_benchmark/modeling/rpn/{rpn,anchor_generator,inference,loss}.py: the single-level (C4) case on
the sync-free path, the multi-level (FPN) case with the reference's host-driven control flow; state-dict names
``rpn.head.{conv,cls_logits,bbox_pred}.{weight,bias}`` and the buffers ``rpn.anchor_generator.cell_anchors.{i}``
are preserved.
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from ..structures import BoxList, is_source_image
from .backbone import Conv2dParams
from ..utils.sections import section
from .sampling import BELOW_LOW_THRESHOLD, BETWEEN_THRESHOLDS, balanced_sample


def generate_cell_anchors(stride, sizes, aspect_ratios):
    """The classic ratio-then-scale enumeration around the (0,0,stride-1,stride-1) window, float64 with
    np.round, cast to float32 (rpn/anchor_generator.py:222-291)."""
    scales = np.asarray(sizes, dtype=np.float64) / stride
    ratios = np.asarray(aspect_ratios, dtype=np.float64)

    def centre(box):
        w, h = box[2] - box[0] + 1, box[3] - box[1] + 1
        return w, h, box[0] + 0.5 * (w - 1), box[1] + 0.5 * (h - 1)

    def windows(ws, hs, cx, cy):
        ws, hs = ws.reshape(-1, 1), hs.reshape(-1, 1)
        return np.hstack([cx - 0.5 * (ws - 1), cy - 0.5 * (hs - 1), cx + 0.5 * (ws - 1), cy + 0.5 * (hs - 1)])

    w, h, cx, cy = centre(np.array([0, 0, stride - 1, stride - 1], dtype=np.float64))
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    out = []
    for box in windows(ws, hs, cx, cy):
        w, h, cx, cy = centre(box)
        out.append(windows(w * scales, h * scales, cx, cy))
    return torch.from_numpy(np.vstack(out)).float()


class BufferList(nn.Module):
    def __init__(self, buffers):
        super().__init__()
        for i, b in enumerate(buffers):
            self.register_buffer(str(i), b)

    def __iter__(self):
        return iter(self._buffers.values())

    def __len__(self):
        return len(self._buffers)


class AnchorGenerator(nn.Module):
    """AnchorGenerator (anchor_generator.py:34-125), single- or multi-level.  The grid and its visibility mask are
    produced by one kernel and cached per (feature size, image size) instead of being rebuilt from
    arange/meshgrid/stack every iteration."""

    def __init__(self, sizes, aspect_ratios, stride, straddle_thresh):
        """stride: one int (single level: all `sizes` at every location) or a tuple of strides, one per entry of
        `sizes` (FPN: one anchor size per level, anchor_generator.py:47-66)."""
        super().__init__()
        self.straddle_thresh = int(straddle_thresh)
        if isinstance(stride, (tuple, list)) and len(stride) > 1:
            if len(stride) != len(sizes):
                raise RuntimeError("FPN should have #anchor_strides == #sizes")
            self.strides = tuple(int(s) for s in stride)
            cells = [generate_cell_anchors(st, sz if isinstance(sz, (tuple, list)) else (sz,), aspect_ratios)
                     for st, sz in zip(self.strides, sizes)]
        else:
            self.strides = (int(stride[0] if isinstance(stride, (tuple, list)) else stride),)
            cells = [generate_cell_anchors(self.strides[0], sizes, aspect_ratios)]
        self.stride = self.strides[0]
        self.cell_anchors = BufferList(cells)
        self._cache = {}

    def num_anchors_per_location(self):
        return [len(c) for c in self.cell_anchors]

    def grid(self, fh, fw, img_w, img_h, level=0):
        cell = list(self.cell_anchors)[level]
        key = (level, fh, fw, img_w, img_h, cell.device, cell._version)
        if key not in self._cache:       # entries are kept: a captured step graph holds the address of its grid
            self._cache[key] = ops.anchor_grid(cell, fh, fw, self.strides[level], img_w, img_h, self.straddle_thresh)
        return self._cache[key]


class RPNHead(nn.Module):
    """3x3 conv + ReLU, then the two 1x1 predictors (rpn/rpn.py:13-46); outputs stay NHWC, which is
    already the (h, w, anchor) order permute_and_flatten (rpn/utils.py:10-14) produces."""

    def __init__(self, in_channels, num_anchors):
        super().__init__()
        self.conv = Conv2dParams(in_channels, in_channels, 3, 1, 1, bias=True)
        self.cls_logits = Conv2dParams(in_channels, num_anchors, 1, bias=True)
        self.bbox_pred = Conv2dParams(in_channels, num_anchors * 4, 1, bias=True)
        for l in (self.conv, self.cls_logits, self.bbox_pred):
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0)

    def forward(self, feat):
        t = ops.conv_bn_act(feat, self.conv.weight, None, self.conv.bias, pad=1, relu=True)
        # both 1x1 predictors as ONE tensor-core GEMM (15 + 60 output channels padded to 96): ops.fused_heads
        logits, deltas = ops.fused_heads(t, [self.cls_logits.weight, self.bbox_pred.weight],
                                         [self.cls_logits.bias, self.bbox_pred.bias])
        return logits, deltas


class ProposalBatch(object):
    """Fixed-capacity proposals of a batch: boxes [N,cap,4], objectness [N,cap], count int32 [N] (device);
    sizes: the un-padded (w, h) of every image (host)."""

    def __init__(self, boxes, objectness, count, sizes):
        self.boxes, self.objectness, self.count, self.sizes = boxes, objectness, count, list(sizes)

    def slice(self, a, b):
        """Images [a, b) of the batch (views)."""
        return ProposalBatch(self.boxes[a:b], self.objectness[a:b], self.count[a:b], self.sizes[a:b])

    def to_boxlists(self):
        """list[BoxList] like RPNPostProcessor returns (reads the counts on the host)."""
        out = []
        for i, c in enumerate(self.count.tolist()):
            bl = BoxList(self.boxes[i, :c], self.sizes[i], mode="xyxy")
            bl.add_field("objectness", self.objectness[i, :c])
            out.append(bl)
        return out


class RPNModule(nn.Module):
    def __init__(self, cfg, rng):
        super().__init__()
        self.cfg = cfg.clone()
        R = cfg.MODEL.RPN
        self.fpn = bool(R.USE_FPN)
        if self.fpn != (len(R.ANCHOR_STRIDE) > 1):           # make_anchor_generator's asserts (anchor_generator.py:133-138)
            raise AssertionError("USE_FPN needs one ANCHOR_STRIDE per ANCHOR_SIZE; non-FPN a single ANCHOR_STRIDE")
        if R.RPN_HEAD != "SingleConvRPNHead":
            raise NotImplementedError("RPN head {} is outside the accelerated path".format(R.RPN_HEAD))
        self.anchor_generator = AnchorGenerator(R.ANCHOR_SIZES, R.ASPECT_RATIOS,
                                                tuple(R.ANCHOR_STRIDE) if self.fpn else R.ANCHOR_STRIDE[0],
                                                R.STRADDLE_THRESH)
        self.head = RPNHead(cfg.MODEL.BACKBONE.OUT_CHANNELS, self.anchor_generator.num_anchors_per_location()[0])
        self.rng = rng
        self.proposal_hook = None
        self.keep_debug = False       # tests: keep label / sample index tensors of the last step (host reads)
        self.overlap_loss = True      # sync-free path: RPN loss chain on a side stream beside the proposal chain
        self.__dict__["_side"] = None

    def set_proposal_hook(self, fn):
        """fn(list[BoxList]) -> list[BoxList], called on the proposals of every forward (None = off).
        Like set_random_source this exists for parity tests: the TF32 arm is compared with the fp32 arm on
        IDENTICAL hard decisions (top-k order, NMS survivors), so that the comparison measures arithmetic
        error instead of a flipped near-tie."""
        self.proposal_hook = fn

    @staticmethod
    def _topk_decode(logits, deltas, anchors, k, image_sizes, min_size):
        """ops.rpn_topk_decode with every image clipped to ITS un-padded size (rpn/inference.py:101-103 builds one
        BoxList per image with that image's size; a batch may mix datasets of different sizes, e.g.
        cityscapes -> kitti).  Equal sizes: one launch for the batch; otherwise one launch per image."""
        sizes = [(int(h), int(w)) for h, w in image_sizes]
        if len(set(sizes)) == 1:
            ih, iw = sizes[0]
            return ops.rpn_topk_decode(logits, deltas, anchors, k, iw, ih, min_size)
        parts = [ops.rpn_topk_decode(logits[i:i + 1], deltas[i:i + 1], anchors, k, iw, ih, min_size)
                 for i, (ih, iw) in enumerate(sizes)]
        return tuple(torch.cat(t, dim=0) for t in zip(*parts))

    def _anchors_and_visibility(self, fh, fw, image_sizes, level=0):
        """(anchors [A,4], visibility): the anchor coordinates depend on the feature map only; the visibility mask
        (anchor_generator.py:98-111) on each image's un-padded size — one mask when all images share a size, else a
        list with one mask per image."""
        sizes = [(int(h), int(w)) for h, w in image_sizes]
        grids = {sz: self.anchor_generator.grid(fh, fw, sz[1], sz[0], level=level) for sz in set(sizes)}
        anchors = grids[sizes[0]][0]
        if len(grids) == 1:
            return anchors, grids[sizes[0]][1]
        return anchors, [grids[sz][1] for sz in sizes]

    # ---- proposals (RPNPostProcessor, rpn/inference.py:76-152) -------------------------------------
    @torch.no_grad()
    def proposals(self, anchors, logits, deltas, image_sizes, targets):
        R = self.cfg.MODEL.RPN
        train = self.training
        pre = R.PRE_NMS_TOP_N_TRAIN if train else R.PRE_NMS_TOP_N_TEST
        post = R.POST_NMS_TOP_N_TRAIN if train else R.POST_NMS_TOP_N_TEST
        n, fh, fw, a = logits.shape
        k = min(pre, fh * fw * a)
        wh = [(int(w), int(h)) for h, w in image_sizes]
        boxes, scores, _, valid = self._topk_decode(logits, deltas, anchors, k, image_sizes, R.MIN_SIZE)
        out = []
        if R.NMS_THRESH > 0:
            keep, cnt = ops.nms_sorted_batched(boxes, valid, R.NMS_THRESH, post)     # whole batch, 2 launches
            counts = cnt.tolist()                                                    # the one host read
            for i in range(n):
                sel = keep[i, : counts[i]]
                bl = BoxList(boxes[i][sel], wh[i], mode="xyxy")
                bl.add_field("objectness", scores[i][sel])
                out.append(bl)
        else:
            valid_h = valid.tolist()
            for i in range(n):
                bl = BoxList(boxes[i, : valid_h[i]], wh[i], mode="xyxy")
                bl.add_field("objectness", scores[i, : valid_h[i]])
                out.append(bl)
        if train and targets is not None:       # add_gt_proposals: source images only (:51-74)
            for i, t in enumerate(targets):
                if is_source_image(t):
                    gt = t.bbox.to(out[i].bbox.dtype)
                    bl = BoxList(torch.cat([out[i].bbox, gt]), out[i].size, mode="xyxy")
                    bl.add_field("objectness", torch.cat([out[i].get_field("objectness"),
                                                          torch.ones(len(gt), device=gt.device)]))
                    out[i] = bl
        return out

    # ---- losses (RPNLossComputation, rpn/loss.py:57-143) -------------------------------------------
    def losses(self, anchors, visibility, logits, deltas, targets):
        """visibility: one uint8 [A] mask, or a list with one mask per image (images of different sizes)."""
        R = self.cfg.MODEL.RPN
        labels, reg_targets = [], []
        for i, t in enumerate(targets):                     # labels exist for source images only (:66-67)
            if not is_source_image(t):
                continue
            vis = (visibility[i] if isinstance(visibility, (list, tuple)) else visibility).bool()
            gt = t.convert("xyxy").bbox
            m, _ = ops.match(gt, anchors, R.FG_IOU_THRESHOLD, R.BG_IOU_THRESHOLD, True)
            lab = (m >= 0).to(torch.float32)
            lab[m == BELOW_LOW_THRESHOLD] = 0
            lab[~vis] = -1
            lab[m == BETWEEN_THRESHOLDS] = -1
            labels.append(lab)
            reg_targets.append(ops.box_encode(gt, anchors, m, (1.0, 1.0, 1.0, 1.0)))
        pos_m, neg_m = balanced_sample(labels, R.BATCH_SIZE_PER_IMAGE, R.POSITIVE_FRACTION, self.rng)
        pos = torch.nonzero(torch.cat(pos_m, dim=0)).squeeze(1)
        neg = torch.nonzero(torch.cat(neg_m, dim=0)).squeeze(1)
        sampled = torch.cat([pos, neg], dim=0)
        obj = logits.reshape(-1)                             # (n, h, w, a) order == permute_and_flatten
        reg = deltas.reshape(-1, 4)
        labels = torch.cat(labels, dim=0)
        reg_targets = torch.cat(reg_targets, dim=0)
        box_loss = ops.smooth_l1_sum(reg[pos], reg_targets[pos], 1.0 / 9, float(sampled.numel()))
        obj_loss = ops.bce_with_logits_mean(obj[sampled], labels[sampled])
        self.last = dict(labels=labels, pos=pos, neg=neg, reg_targets=reg_targets)
        return obj_loss, box_loss

    # ---- fixed-capacity, sync-free variants (training) --------------------------------------------------
    @torch.no_grad()
    def proposals_static(self, anchors, logits, deltas, image_sizes, meta):
        """RPNPostProcessor without host reads: (proposals [N,cap,4], objectness [N,cap], count int32 [N]) with
        cap = POST_NMS_TOP_N_TRAIN + max GT boxes per image; 4 launches for the whole batch."""
        R = self.cfg.MODEL.RPN
        n, fh, fw, a = logits.shape
        k = min(R.PRE_NMS_TOP_N_TRAIN, fh * fw * a)
        post = min(R.POST_NMS_TOP_N_TRAIN, k)
        boxes, scores, _, valid = self._topk_decode(logits, deltas, anchors, k, image_sizes, R.MIN_SIZE)
        keep, cnt = ops.nms_sorted_batched(boxes, valid, R.NMS_THRESH, post)
        return ProposalBatch(*ops.proposals_gather(boxes, scores, keep, cnt, meta["gt_cat"], meta["gt_offsets"],
                                                   meta["append_gt"], post + meta["max_gt"],
                                                   gt_counts=meta.get("gt_counts")),
                             sizes=[(int(w), int(h)) for h, w in image_sizes])

    def losses_static(self, anchors, visibility, logits, deltas, targets, meta):
        """RPNLossComputation (rpn/loss.py:57-143) without host reads, as five launches per source image plus two for
        the batch: Matcher (2), labels, random keys, then ONE sampler launch for all source images (per image the 256
        sampled anchors sit in a fixed index vector with a device-side count) and ONE fused loss kernel (BCE +
        smooth-L1 + both gradients)."""
        R = self.cfg.MODEL.RPN
        B = R.BATCH_SIZE_PER_IMAGE
        max_pos = int(B * R.POSITIVE_FRACTION)
        A = anchors.shape[0]
        src = [i for i, t in enumerate(targets) if is_source_image(t)]      # labels exist for source images only (:66-67)
        if src != list(range(len(src))):
            raise AssertionError("source images must come first in a batch (rpn/loss.py indexes the objectness of the "
                                 "first len(labels) images)")
        labs, ms, keys = [], [], []
        for i in src:
            t = targets[i]
            vis = visibility[i] if isinstance(visibility, (list, tuple)) else visibility
            gt = t.convert("xyxy").bbox
            m_dev = getattr(t, "_gt_count_dev", None)        # GT padded to a capacity (signature-free step graph)
            m, _ = ops.match(gt, anchors, R.FG_IOU_THRESHOLD, R.BG_IOU_THRESHOLD, True, m_dev=m_dev)
            lab = ops.rpn_anchor_labels(m, vis)
            labs.append(lab)
            ms.append(m)
            keys.append(self.rng.sample_keys(lab))
        one = len(src) == 1
        lab = labs[0].view(1, -1) if one else torch.stack(labs)
        m = ms[0].view(1, -1) if one else torch.stack(ms)
        key = keys[0].view(1, -1) if one else torch.stack(keys)
        sel, cnt = ops.balanced_sample(lab, None, key, B, max_pos)
        src_img = meta.get("src_index")
        if src_img is None:
            src_img = torch.tensor(src, dtype=torch.int32, device=anchors.device)
        obj_loss, box_loss = ops.rpn_sampled_losses(logits, deltas, anchors, sel, cnt, lab, m, meta["gt_cat"],
                                                    meta["gt_offsets"], src_img, 1.0 / 9)
        if self.keep_debug:                                  # tests: host reads
            pos, neg = [], []
            for s in range(len(src)):
                sl = sel[s, : int(cnt[s, 1])]
                ls = lab[s][sl]
                pos.append(sl[ls == 1] + s * A)
                neg.append(sl[ls == 0] + s * A)
            self.last = dict(labels=lab.reshape(-1).to(torch.float32), pos=torch.cat(pos), neg=torch.cat(neg))
        return obj_loss, box_loss

    def forward_static(self, images, features, targets, head_out, meta, early_backward=False, after_head=None):
        """head_out: (logits, deltas) of self.head(features[0]), or None when early_backward is set (the head then runs
        here, on a detached copy of the features).  Returns (proposals, losses, pending): `pending` is None or the
        (stream, features, cut) triple finish_early_backward() needs."""
        feat = features[0]
        n, fh, fw, _ = feat.shape
        anchors, vis = self._anchors_and_visibility(fh, fw, images.image_sizes)
        if early_backward:
            return self._forward_static_early(images, feat, targets, meta, anchors, vis, after_head)
        logits, deltas = head_out
        if self.overlap_loss:
            # The proposal chain (top-k, NMS, gather: one or two CTAs per image, ~1.2 ms) and the RPN loss chain
            # (match, sampler, encode, losses: ~0.45 ms, also a few CTAs) are independent: the loss chain runs on a
            # side stream beside the proposals and joins before the box head.  Inside a step graph the fork / join
            # become graph edges.  The random draws keep their order (RPN sampler before box-head sampler).
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=feat.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                obj_loss, box_loss = self.losses_static(anchors, vis, logits, deltas, targets, meta)
            props = self.proposals_static(anchors, logits.detach(), deltas.detach(), images.image_sizes, meta)
            main.wait_stream(self._side)
        else:
            props = self.proposals_static(anchors, logits.detach(), deltas.detach(), images.image_sizes, meta)
            obj_loss, box_loss = self.losses_static(anchors, vis, logits, deltas, targets, meta)
        if self.proposal_hook is not None:
            props = self.proposal_hook(props)
        return props, {"loss_objectness": obj_loss, "loss_rpn_box_reg": box_loss}, None

    def _forward_static_early(self, images, feat, targets, meta, anchors, vis, after_head=None):
        """The RPN branch with its backward pass run EARLY.  The proposal chain and the box-head sampler that follows it
        are latency-bound (one CTA per image: ~1.4 ms during which the GPU is otherwise idle), and nothing dense in the
        forward pass can fill that time — everything after depends on the proposals.  The RPN losses, however, depend
        on the trunk features only, so their whole backward (the 3x3 conv's data and weight gradients: the largest
        GEMMs of the step) can run right there: the head is evaluated on `cut = feat.detach()`, the losses are
        back-propagated on the side stream as soon as they exist (weight gradients accumulate into the zeroed flat
        buffer, d loss / d feat lands in cut.grad), and finish_early_backward() hands cut.grad to the graph of `feat`.
        The dense kernels of that pass leave two SMs per image to the one-CTA-per-image kernels of the main stream
        (ops.sm_budget).  Requires gradients zeroed BEFORE the forward pass and unit loss weights (FlatSGDTrainer).
        after_head(event): called once the head has been launched, with the event that marks its end; may return the
        stream its own dense work runs on."""
        cuda = feat.is_cuda
        if cuda:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=feat.device)
            side = self._side
            side.wait_stream(main)
            on_side = lambda: torch.cuda.stream(side)
        else:
            import contextlib
            main = side = None
            on_side = contextlib.nullcontext
        with on_side():
            cut = feat.detach().requires_grad_(True)
            logits, deltas = self.head(cut)
            ready = None
            if cuda:
                ready = torch.cuda.Event()
                ready.record(side)
        other = None
        if after_head is not None:          # other early passes (DA image head) start when the RPN head has finished
            other = after_head(ready)
        with on_side():
            obj_loss, box_loss = self.losses_static(anchors, vis, logits, deltas, targets, meta)
            if cuda and other is not None:
                # two streams of persistent dense kernels beside each other only take turns on the SMs (the second
                # kernel's CTAs wait for the first's to retire and then run as a late wave): this backward pass
                # follows the other early pass, which has been running during the loss chain above
                side.wait_stream(other)
            if cuda:
                with ops.sm_budget(ops.NUM_SMS - 2 * feat.shape[0]):
                    torch.autograd.backward([obj_loss + box_loss])
            else:
                torch.autograd.backward([obj_loss + box_loss])
            obj_loss, box_loss = obj_loss.detach(), box_loss.detach()
        if cuda:
            main.wait_event(ready)
            logits.record_stream(main)
            deltas.record_stream(main)
        props = self.proposals_static(anchors, logits.detach(), deltas.detach(), images.image_sizes, meta)
        if self.proposal_hook is not None:
            props = self.proposal_hook(props)
        return props, {"loss_objectness": obj_loss, "loss_rpn_box_reg": box_loss}, (side, feat, cut)

    @staticmethod
    def finish_early_backward(losses, pending):
        """Join the early backward pass (stream and autograd graph): after this, back-propagating through
        losses["loss_objectness"] with a gradient of 1 adds d(RPN losses)/d feat to the gradient of the features."""
        side, feat, cut = pending
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        losses["loss_objectness"] = ops.inject_grad(losses["loss_objectness"], feat, cut)
        return losses

    # ---- FPN: per-level post-processing + select_over_all_levels (rpn/inference.py:126-181) ----------
    @torch.no_grad()
    def proposals_fpn(self, grids, outs, image_sizes, targets):
        R = self.cfg.MODEL.RPN
        train = self.training
        pre = R.PRE_NMS_TOP_N_TRAIN if train else R.PRE_NMS_TOP_N_TEST
        post = R.POST_NMS_TOP_N_TRAIN if train else R.POST_NMS_TOP_N_TEST
        fpn_post = R.FPN_POST_NMS_TOP_N_TRAIN if train else R.FPN_POST_NMS_TOP_N_TEST
        wh = [(int(w), int(h)) for h, w in image_sizes]
        n = outs[0][0].shape[0]
        boxes_i, scores_i = [[] for _ in range(n)], [[] for _ in range(n)]
        for (anchors, _), (logits, deltas) in zip(grids, outs):
            _, fh, fw, a = logits.shape
            k = min(pre, fh * fw * a)
            boxes, scores, _, valid = self._topk_decode(logits.detach(), deltas.detach(), anchors, k, image_sizes,
                                                        R.MIN_SIZE)
            keep, cnt = ops.nms_sorted_batched(boxes, valid, R.NMS_THRESH, min(post, k))
            for i, c in enumerate(cnt.tolist()):             # one host read per level
                sel = keep[i, :c]
                boxes_i[i].append(boxes[i][sel])
                scores_i[i].append(scores[i][sel])
        boxes_i = [torch.cat(b) for b in boxes_i]
        scores_i = [torch.cat(s) for s in scores_i]
        if train:            # :160-171 — ONE top-k over the proposals of the whole batch (the reference's known quirk)
            allsc = torch.cat(scores_i)
            _, top = torch.topk(allsc, min(fpn_post, allsc.numel()), dim=0, sorted=True)
            mask = torch.zeros_like(allsc, dtype=torch.bool)
            mask[top] = True
            picks = [torch.nonzero(m).squeeze(1) for m in mask.split([len(s) for s in scores_i])]
        else:                # :172-180 — per image, returned in descending objectness order
            picks = [torch.topk(s, min(fpn_post, s.numel()), dim=0, sorted=True)[1] for s in scores_i]
        out = []
        for i, pk in enumerate(picks):
            bl = BoxList(boxes_i[i][pk], wh[i], mode="xyxy")
            bl.add_field("objectness", scores_i[i][pk])
            out.append(bl)
        if train and targets is not None:
            for i, t in enumerate(targets):
                if is_source_image(t):
                    gt = t.bbox.to(out[i].bbox.dtype)
                    bl = BoxList(torch.cat([out[i].bbox, gt]), out[i].size, mode="xyxy")
                    bl.add_field("objectness", torch.cat([out[i].get_field("objectness"),
                                                          torch.ones(len(gt), device=gt.device)]))
                    out[i] = bl
        return out

    # ---- FPN, fixed-capacity and sync-free (training): per level top-k / NMS into padded buffers, the batch-wide
    # select_over_all_levels cut (rpn/inference.py:160-171) as a masked top-k + stable partition on the device
    @torch.no_grad()
    def proposals_fpn_static(self, grids, outs, image_sizes, meta):
        R = self.cfg.MODEL.RPN
        pre, post, fpn_post = R.PRE_NMS_TOP_N_TRAIN, R.POST_NMS_TOP_N_TRAIN, R.FPN_POST_NMS_TOP_N_TRAIN
        n = outs[0][0].shape[0]
        dev = outs[0][0].device
        no_gt = torch.zeros(n, dtype=torch.uint8, device=dev)
        bl, sl, vl = [], [], []
        for (anchors, _), (logits, deltas) in zip(grids, outs):
            _, fh, fw, a = logits.shape
            k = min(pre, fh * fw * a)
            p_l = min(post, k)
            boxes, scores, _, valid = self._topk_decode(logits, deltas, anchors, k, image_sizes, R.MIN_SIZE)
            keep, cnt = ops.nms_sorted_batched(boxes, valid, R.NMS_THRESH, p_l)
            b, s_, c = ops.proposals_gather(boxes, scores, keep, cnt, meta["gt_cat"], meta["gt_offsets"], no_gt, p_l,
                                            gt_counts=meta.get("gt_counts"))
            bl.append(b)
            sl.append(s_)
            vl.append(torch.arange(p_l, device=dev).unsqueeze(0) < c.unsqueeze(1))
        allb, alls, valid = torch.cat(bl, dim=1), torch.cat(sl, dim=1), torch.cat(vl, dim=1)     # [N, P, ...]
        P = alls.shape[1]
        # ONE top-k over the proposals of the whole batch (the reference's known quirk), then every image keeps its
        # selected proposals in their concatenated (level-major) order: mask + stable partition, no nonzero
        masked = torch.where(valid, alls, torch.full_like(alls, float("-inf"))).reshape(-1)
        top_v, top_i = torch.topk(masked, min(fpn_post, n * P), dim=0, sorted=True)
        sel = torch.zeros(n * P, dtype=torch.bool, device=dev)
        sel.scatter_(0, top_i, top_v > float("-inf"))
        sel = sel.view(n, P)
        order = torch.sort((~sel).to(torch.uint8), dim=1, stable=True)[1]
        cap_keep = min(P, fpn_post)
        boxes, obj, count = ops.proposals_gather(allb, alls, order[:, :cap_keep].contiguous(),
                                                 sel.sum(dim=1).to(torch.int32), meta["gt_cat"], meta["gt_offsets"],
                                                 meta["append_gt"], cap_keep + meta["max_gt"],
                                                 gt_counts=meta.get("gt_counts"))
        return ProposalBatch(boxes, obj, count, sizes=[(int(w), int(h)) for h, w in image_sizes])

    def forward_fpn_static(self, images, features, targets, meta):
        outs = [self.head(f) for f in features]              # the same head on every level (rpn.py:39-46)
        grids = [self._anchors_and_visibility(f.shape[1], f.shape[2], images.image_sizes, level=l)
                 for l, f in enumerate(features)]
        props = self.proposals_fpn_static(grids, [(lg.detach(), dl.detach()) for lg, dl in outs], images.image_sizes,
                                          meta)
        if self.proposal_hook is not None:
            props = self.proposal_hook(props)
        n = outs[0][0].shape[0]
        obj = torch.cat([lg.reshape(n, -1) for lg, _ in outs], dim=1)
        reg = torch.cat([dl.reshape(n, -1, 4) for _, dl in outs], dim=1)
        anchors = torch.cat([g[0] for g in grids], dim=0)
        if isinstance(grids[0][1], list):                    # images of different sizes: one mask per image
            vis = [torch.cat([g[1][i] for g in grids], dim=0) for i in range(n)]
        else:
            vis = torch.cat([g[1] for g in grids], dim=0)
        obj_loss, box_loss = self.losses_static(anchors, vis, obj, reg, targets, meta)
        return props, {"loss_objectness": obj_loss, "loss_rpn_box_reg": box_loss}

    def forward_fpn(self, images, features, targets=None):
        outs = [self.head(f) for f in features]              # the same head on every level (rpn.py:39-46)
        grids = [self._anchors_and_visibility(f.shape[1], f.shape[2], images.image_sizes, level=l)
                 for l, f in enumerate(features)]
        boxes = self.proposals_fpn(grids, outs, images.image_sizes, targets)
        if self.proposal_hook is not None:
            boxes = self.proposal_hook(boxes)
        if not self.training:
            return boxes, {}
        # concat_box_prediction_layers (rpn/utils.py:17-45): per image, the levels one after the other
        n = outs[0][0].shape[0]
        obj = torch.cat([lg.reshape(n, -1) for lg, _ in outs], dim=1)
        reg = torch.cat([dl.reshape(n, -1, 4) for _, dl in outs], dim=1)
        anchors = torch.cat([g[0] for g in grids], dim=0)
        if isinstance(grids[0][1], list):                    # images of different sizes: one mask per image
            vis = [torch.cat([g[1][i] for g in grids], dim=0) for i in range(n)]
        else:
            vis = torch.cat([g[1] for g in grids], dim=0)
        obj_loss, box_loss = self.losses(anchors, vis, obj, reg, targets)
        return boxes, {"loss_objectness": obj_loss, "loss_rpn_box_reg": box_loss}

    def forward(self, images, features, targets=None, head_out=None):
        if self.fpn:
            return self.forward_fpn(images, features, targets)
        feat = features[0]
        logits, deltas = head_out if head_out is not None else self.head(feat)
        n, fh, fw, _ = feat.shape
        anchors, vis = self._anchors_and_visibility(fh, fw, images.image_sizes)
        with section("  rpn_proposals"):
            boxes = self.proposals(anchors, logits.detach(), deltas.detach(), images.image_sizes, targets)
            if self.proposal_hook is not None:
                boxes = self.proposal_hook(boxes)
        if not self.training:
            return boxes, {}
        with section("  rpn_loss"):
            obj_loss, box_loss = self.losses(anchors, vis, logits, deltas, targets)
        return boxes, {"loss_objectness": obj_loss, "loss_rpn_box_reg": box_loss}


def build_rpn(cfg, rng):
    if cfg.MODEL.RETINANET_ON:
        raise NotImplementedError("RetinaNet is outside the accelerated path")
    return RPNModule(cfg, rng)
