"""Eval-side callers of the hot path: the inference loop, the COCO-json detection writer and a
box-AP evaluator, so that the headline's second half (Foggy-val mAP@0.5 at equal iterations) can be computed without
leaving this package.  Host code, no kernels.

Reference call sites:
  * engine/inference.py:18-52        compute_on_dataset (eval loop, outputs moved to the CPU, keyed by image id)
  * data/datasets/evaluation/coco/coco_eval.py:81-112   prepare_for_coco_detection (resize to the original image
    size, xyxy -> xywh, contiguous label -> json category id)
  * coco_eval.py:324-337             COCOeval(coco_gt, coco_dt, "bbox").evaluate() / accumulate() / summarize()

Third-party arithmetic: `pycocotools.cocoeval.COCOeval` is NOT under /root/reference, is un-pinned by the reference
(requirements.txt) and is absent from this container.  `evaluate_bbox_ap` restates its published algorithm for
iouType="bbox", area range "all": per image and category, detections in descending score order (stable), at most
`max_dets` per image, greedy matching to the not-yet-matched ground truth of highest IoU >= threshold (crowd regions
may absorb any number of detections, which are then ignored; IoU against a crowd = intersection / detection area),
precision made monotone from the right and sampled at 101 recall thresholds, averaged over the categories that have
ground truth.  PARITY UNPINNED: there is no pycocotools here to check against — the tests hold hand-computed cases.
"""
import numpy as np
import torch


@torch.no_grad()
def compute_on_dataset(model, data_loader, device):
    """engine/inference.py:18-52: {image_id: BoxList on the CPU} for every batch (images, targets, image_ids)."""
    model.eval()
    results = {}
    cpu = torch.device("cpu")
    for batch in data_loader:
        images, image_ids = batch[0], batch[2]
        output = model(images.to(device))
        results.update({img_id: o.to(cpu) for img_id, o in zip(image_ids, output)})
    return results


def prepare_for_coco_detection(predictions, dataset):
    """coco_eval.py:81-112.  predictions: list of BoxList (fields scores, labels) indexed like the dataset;
    dataset: id_to_img_map, get_img_info(i) -> {width, height}, contiguous_category_id_to_json_id."""
    coco_results = []
    for image_id, prediction in enumerate(predictions):
        original_id = dataset.id_to_img_map[image_id]
        if len(prediction) == 0:
            continue
        info = dataset.get_img_info(image_id)
        prediction = prediction.resize((info["width"], info["height"])).convert("xywh")
        boxes = prediction.bbox.tolist()
        scores = prediction.get_field("scores").tolist()
        labels = prediction.get_field("labels").tolist()
        mapped = [dataset.contiguous_category_id_to_json_id[i] for i in labels]
        coco_results.extend({"image_id": original_id, "category_id": mapped[k], "bbox": box, "score": scores[k]}
                            for k, box in enumerate(boxes))
    return coco_results


def _iou_xywh(dts, gts, crowd):
    """maskUtils.iou for boxes: dts [D,4], gts [G,4] xywh (float areas, no +1); crowd [G] bool."""
    d, g = np.asarray(dts, dtype=np.float64).reshape(-1, 4), np.asarray(gts, dtype=np.float64).reshape(-1, 4)
    iw = np.minimum(d[:, None, 0] + d[:, None, 2], g[None, :, 0] + g[None, :, 2]) - np.maximum(d[:, None, 0], g[None, :, 0])
    ih = np.minimum(d[:, None, 1] + d[:, None, 3], g[None, :, 1] + g[None, :, 3]) - np.maximum(d[:, None, 1], g[None, :, 1])
    inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
    da, ga = (d[:, 2] * d[:, 3])[:, None], (g[:, 2] * g[:, 3])[None, :]
    union = np.where(np.asarray(crowd, dtype=bool)[None, :], da, da + ga - inter)
    return np.where(union > 0, inter / np.maximum(union, 1e-300), 0.0)


def _match_image(dt_boxes, gt_boxes, gt_crowd, thr):
    """COCOeval.evaluateImg for one (image, category) at one IoU threshold.  dt in descending score order; gt sorted
    with the crowd (ignored) ones last.  Returns (matched [D] bool, ignored [D] bool)."""
    D, G = len(dt_boxes), len(gt_boxes)
    matched, ignored = np.zeros(D, dtype=bool), np.zeros(D, dtype=bool)
    if D == 0 or G == 0:
        return matched, ignored
    ious = _iou_xywh(dt_boxes, gt_boxes, gt_crowd)
    taken = np.zeros(G, dtype=bool)
    for di in range(D):
        best, m = min(thr, 1 - 1e-10), -1
        for gi in range(G):
            if taken[gi] and not gt_crowd[gi]:
                continue
            if m > -1 and not gt_crowd[m] and gt_crowd[gi]:
                break                                   # a regular match is kept rather than traded for a crowd
            if ious[di, gi] < best:
                continue
            best, m = ious[di, gi], gi
        if m == -1:
            continue
        matched[di], ignored[di], taken[m] = True, bool(gt_crowd[m]), True
    return matched, ignored


def evaluate_bbox_ap(coco_results, gt_annotations, iou_thrs=(0.5,), max_dets=100):
    """Box AP from COCO-json detections (`prepare_for_coco_detection`) and COCO annotation dicts
    ({image_id, category_id, bbox xywh, iscrowd}).  Returns {"AP": mean over thresholds and categories,
    "AP@<thr>": ..., "per_category": {cat: {thr: ap}}}.  iou_thrs=(0.5,) is the Foggy-Cityscapes mAP@0.5 of
    BASELINE.json; np.arange(0.5, 1.0, 0.05) gives COCO's primary metric."""
    rec_thrs = np.linspace(0.0, 1.0, 101)
    cats = sorted({g["category_id"] for g in gt_annotations})
    gts, dts = {}, {}
    for g in gt_annotations:
        gts.setdefault((g["image_id"], g["category_id"]), []).append(g)
    for d in coco_results:
        dts.setdefault((d["image_id"], d["category_id"]), []).append(d)
    images = sorted({k[0] for k in gts} | {k[0] for k in dts})
    per_cat = {}
    for cat in cats:
        per_img = []
        for img in images:
            g = sorted(gts.get((img, cat), []), key=lambda a: int(bool(a.get("iscrowd", 0))))     # stable: crowd last
            d = dts.get((img, cat), [])
            if not g and not d:
                continue
            order = np.argsort([-x["score"] for x in d], kind="mergesort")[:max_dets]
            d = [d[i] for i in order]
            per_img.append((np.array([x["score"] for x in d], dtype=np.float64), [x["bbox"] for x in d],
                            [x["bbox"] for x in g], np.array([bool(x.get("iscrowd", 0)) for x in g], dtype=bool)))
        npig = sum(int((~crowd).sum()) for _, _, _, crowd in per_img)
        if npig == 0:
            continue                                   # no ground truth of this category: left out of the mean
        scores = np.concatenate([s for s, _, _, _ in per_img]) if per_img else np.zeros(0)
        order = np.argsort(-scores, kind="mergesort")
        per_cat[cat] = {}
        for thr in iou_thrs:
            res = [_match_image(db, gb, crowd, thr) for _, db, gb, crowd in per_img]
            matched = np.concatenate([m for m, _ in res])[order] if res else np.zeros(0, dtype=bool)
            ignored = np.concatenate([i for _, i in res])[order] if res else np.zeros(0, dtype=bool)
            tp = np.cumsum(matched & ~ignored).astype(np.float64)
            fp = np.cumsum(~matched & ~ignored).astype(np.float64)
            rc = tp / npig
            pr = tp / (fp + tp + np.spacing(1))
            for i in range(len(pr) - 1, 0, -1):
                if pr[i] > pr[i - 1]:
                    pr[i - 1] = pr[i]
            q = np.zeros(len(rec_thrs))
            idx = np.searchsorted(rc, rec_thrs, side="left")
            ok = idx < len(pr)
            q[ok] = pr[idx[ok]]
            per_cat[cat][float(thr)] = float(q.mean())
    out = {"per_category": per_cat}
    for thr in iou_thrs:
        vals = [v[float(thr)] for v in per_cat.values()]
        out["AP@%.2f" % thr] = float(np.mean(vals)) if vals else -1.0
    allv = [x for v in per_cat.values() for x in v.values()]
    out["AP"] = float(np.mean(allv)) if allv else -1.0
    return out
