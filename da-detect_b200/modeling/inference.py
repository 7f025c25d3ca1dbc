"""Eval-mode box post-processing (box_head/inference.py:12-150): softmax, per-class decode with
weights (10,10,5,5), clip, score threshold, per-class NMS, top-k detections.  A 'next' row:
decode and NMS run on our kernels; softmax/top-k glue is still torch."""
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
        bx = bx.reshape(-1, 4).clone()
        bx[:, 0].clamp_(min=0, max=w - 1)
        bx[:, 1].clamp_(min=0, max=h - 1)
        bx[:, 2].clamp_(min=0, max=w - 1)
        bx[:, 3].clamp_(min=0, max=h - 1)
        bx = bx.reshape(-1, nc * 4)
        out_b, out_s, out_l = [], [], []
        for j in range(1, nc):
            inds = torch.nonzero(pr[:, j] > H.SCORE_THRESH).squeeze(1)
            sj, bj = pr[inds, j], bx[inds, j * 4:(j + 1) * 4].contiguous()
            keep = ops.nms(bj, sj.contiguous(), H.NMS)
            out_b.append(bj[keep])
            out_s.append(sj[keep])
            out_l.append(torch.full((len(keep),), j, dtype=torch.int64, device=bj.device))
        b, s, l = torch.cat(out_b), torch.cat(out_s), torch.cat(out_l)
        if len(s) > H.DETECTIONS_PER_IMG > 0:
            thr, _ = torch.kthvalue(s.cpu(), len(s) - H.DETECTIONS_PER_IMG + 1)
            k = torch.nonzero(s >= thr.item()).squeeze(1)
            b, s, l = b[k], s[k], l[k]
        r = BoxList(b, p.size, mode="xyxy")
        r.add_field("scores", s)
        r.add_field("labels", l)
        results.append(r)
    return results
