"""Eval-mode box post-processing (box_head/inference.py:12-150): softmax, per-class decode with weights (10,10,5,5),
clip, score threshold, per-class NMS, top-k detections.

The reference loops over the classes on the host (one `nonzero`, one `_C.nms` and one `torch.full` per class,
:117-131) and reads the detection threshold back through `.cpu()` / `.item()` (:139-143).  Here the classes of an image
are ONE batched NMS problem: every class is a "row" of dd_nms_sorted_batched (its candidates sorted by score, the
count on the device), the survivors are scattered into a [classes, proposals] mask whose `nonzero` is already the
reference's output order (class-major, ascending proposal index within a class, :117-133), and the top-k threshold
stays on the device.  One host read per image remains: the size of the variable-length result."""
import torch

from .. import ops
from ..structures import BoxList


@torch.no_grad()
def box_post_process(cfg, class_logits, box_regression, proposals):
    H = cfg.MODEL.ROI_HEADS
    probs = torch.softmax(class_logits, -1)
    sizes = [len(p) for p in proposals]
    boxes = torch.cat([p.bbox for p in proposals], dim=0)
    decoded = ops.box_decode(box_regression.contiguous(), boxes, H.BBOX_REG_WEIGHTS)
    nc = probs.shape[1]
    results = []
    for p, pr, bx in zip(proposals, probs.split(sizes, 0), decoded.split(sizes, 0)):
        w, h = p.size
        r = pr.shape[0]
        bx = bx.reshape(-1, 4).clone()                       # clip_to_image(remove_empty=False), :84
        bx[:, 0].clamp_(min=0, max=w - 1)
        bx[:, 1].clamp_(min=0, max=h - 1)
        bx[:, 2].clamp_(min=0, max=w - 1)
        bx[:, 3].clamp_(min=0, max=h - 1)
        bx = bx.reshape(r, nc, 4)
        if r == 0 or nc <= 1:
            results.append(_empty(p, bx.device))
            continue
        # classes 1 .. nc-1 as rows: candidates (score > thresh, :116) first, by descending score, ties by lower index
        sc = pr[:, 1:].t().contiguous()                                     # [nc-1, r]
        cand = sc > H.SCORE_THRESH
        key = torch.where(cand, sc, torch.full_like(sc, -1.0))
        order = torch.sort(key, dim=1, descending=True, stable=True)[1]     # [nc-1, r] proposal index per rank
        cls_boxes = bx[:, 1:, :].permute(1, 0, 2)                           # [nc-1, r, 4]
        sorted_boxes = torch.gather(cls_boxes, 1, order.unsqueeze(-1).expand(-1, -1, 4)).contiguous()
        counts = cand.sum(dim=1).to(torch.int32)
        keep, kcnt = ops.nms_sorted_batched(sorted_boxes, counts, H.NMS, r)  # positions in rank order, per class
        live = torch.arange(r, device=keep.device).unsqueeze(0) < kcnt.unsqueeze(1)
        # survivors by proposal index (scatter_add of 0/1: rows beyond the count carry arbitrary positions)
        hits = torch.zeros((nc - 1, r), dtype=torch.int32, device=keep.device)
        hits.scatter_add_(1, torch.gather(order, 1, keep.clamp(min=0, max=r - 1)), live.to(torch.int32))
        kept = hits > 0
        if H.DETECTIONS_PER_IMG > 0:
            # :135-146 — keep the detections whose score reaches the (n - D + 1)-th smallest, when there are more than D
            s_all = torch.where(kept, sc, torch.full_like(sc, -1.0)).reshape(-1)
            n_det = kept.sum()
            d = int(H.DETECTIONS_PER_IMG)
            if s_all.numel() > d:
                thr = torch.topk(s_all, d, sorted=True)[0][d - 1]           # D-th largest == (n - D + 1)-th smallest
                thr = torch.where(n_det > d, thr, torch.full_like(thr, -1.0))
                kept &= sc >= thr
        cj = torch.nonzero(kept)                                            # class-major, ascending proposal index
        cls_idx, prop_idx = cj[:, 0], cj[:, 1]
        out = BoxList(cls_boxes[cls_idx, prop_idx], p.size, mode="xyxy")
        out.add_field("scores", sc[cls_idx, prop_idx])
        out.add_field("labels", cls_idx + 1)
        results.append(out)
    return results


def _empty(p, device):
    out = BoxList(torch.zeros((0, 4), dtype=torch.float32, device=device), p.size, mode="xyxy")
    out.add_field("scores", torch.zeros((0,), dtype=torch.float32, device=device))
    out.add_field("labels", torch.zeros((0,), dtype=torch.int64, device=device))
    return out
