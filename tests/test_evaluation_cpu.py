"""CPU: the eval-side callers (da-detect_b200/engine/evaluation.py) — COCO-json writer against the reference's
prepare_for_coco_detection semantics, and the box-AP evaluator on hand-computed cases of COCOeval's algorithm
(pycocotools is absent here: parity with it is unpinned, see the module docstring)."""
import numpy as np
import pytest
import torch


def det(img, cat, box, score):
    return {"image_id": img, "category_id": cat, "bbox": list(map(float, box)), "score": float(score)}


def gt(img, cat, box, crowd=0):
    return {"image_id": img, "category_id": cat, "bbox": list(map(float, box)), "iscrowd": crowd}


def test_perfect_detections_give_ap_one():
    from dadetect_b200.engine.evaluation import evaluate_bbox_ap
    gts = [gt(1, 1, (10, 10, 50, 40)), gt(1, 2, (100, 20, 30, 30)), gt(2, 1, (5, 5, 20, 20))]
    dts = [det(g["image_id"], g["category_id"], g["bbox"], 0.9 - 0.1 * i) for i, g in enumerate(gts)]
    r = evaluate_bbox_ap(dts, gts, iou_thrs=(0.5, 0.75))
    assert r["AP"] == pytest.approx(1.0) and r["AP@0.50"] == pytest.approx(1.0) and r["AP@0.75"] == pytest.approx(1.0)


def test_hand_computed_precision_recall_curve():
    """2 GT, detections TP(.9) FP(.8) TP(.7): recall [.5,.5,1], precision [1,.5,.667] -> envelope [1,.667,.667];
    the 51 recall thresholds <= 0.5 read 1.0, the other 50 read 2/3: AP = (51 + 50 * 2/3) / 101."""
    from dadetect_b200.engine.evaluation import evaluate_bbox_ap
    gts = [gt(1, 1, (0, 0, 10, 10)), gt(1, 1, (100, 100, 10, 10))]
    dts = [det(1, 1, (0, 0, 10, 10), .9), det(1, 1, (50, 50, 10, 10), .8), det(1, 1, (100, 100, 10, 10), .7)]
    r = evaluate_bbox_ap(dts, gts)
    assert r["AP@0.50"] == pytest.approx((51 + 50 * 2.0 / 3.0) / 101, abs=1e-12)


def test_duplicates_crowds_thresholds_and_max_dets():
    from dadetect_b200.engine.evaluation import evaluate_bbox_ap
    g = [gt(1, 1, (0, 0, 10, 10))]
    # a second detection of the same object is a false positive: precision envelope [1, .5], recall reached at det 1
    r = evaluate_bbox_ap([det(1, 1, (0, 0, 10, 10), .9), det(1, 1, (0, 0, 10, 9), .8)], g)
    assert r["AP@0.50"] == pytest.approx(1.0)
    r = evaluate_bbox_ap([det(1, 1, (0, 0, 10, 9), .9), det(1, 1, (0, 0, 10, 10), .8)], g)     # order does not matter
    assert r["AP@0.50"] == pytest.approx(1.0)
    # IoU 0.6 box: a hit at 0.5, a miss at 0.75
    r = evaluate_bbox_ap([det(1, 1, (0, 0, 10, 6), .9)], g, iou_thrs=(0.5, 0.75))
    assert r["AP@0.50"] == pytest.approx(1.0) and r["AP@0.75"] == pytest.approx(0.0)
    # detections inside a crowd region are ignored (neither TP nor FP); IoU against a crowd = inter / det area
    gc = [gt(1, 1, (0, 0, 10, 10)), gt(1, 1, (100, 100, 200, 200), crowd=1)]
    dc = [det(1, 1, (120, 120, 10, 10), .95), det(1, 1, (130, 130, 10, 10), .94), det(1, 1, (0, 0, 10, 10), .5)]
    assert evaluate_bbox_ap(dc, gc)["AP@0.50"] == pytest.approx(1.0)
    # without the crowd annotation the same two detections are false positives ranked above the hit
    assert evaluate_bbox_ap(dc, gc[:1])["AP@0.50"] == pytest.approx(1.0 / 3.0, abs=1e-9)
    # max_dets: only the top-scoring detections of an image are considered
    many = [det(1, 1, (300 + 20 * i, 0, 10, 10), .9 - .001 * i) for i in range(5)] + [det(1, 1, (0, 0, 10, 10), .1)]
    assert evaluate_bbox_ap(many, g, max_dets=5)["AP@0.50"] == pytest.approx(0.0)
    assert evaluate_bbox_ap(many, g, max_dets=6)["AP@0.50"] == pytest.approx(1.0 / 6.0, abs=1e-9)


def test_categories_without_gt_are_left_out_and_missed_categories_count_zero():
    from dadetect_b200.engine.evaluation import evaluate_bbox_ap
    gts = [gt(1, 1, (0, 0, 10, 10)), gt(1, 2, (50, 50, 10, 10))]
    dts = [det(1, 1, (0, 0, 10, 10), .9), det(1, 3, (0, 0, 10, 10), .9)]          # category 3 has no ground truth
    r = evaluate_bbox_ap(dts, gts)
    assert set(r["per_category"]) == {1, 2}
    assert r["per_category"][1][0.5] == pytest.approx(1.0) and r["per_category"][2][0.5] == 0.0
    assert r["AP@0.50"] == pytest.approx(0.5)
    assert evaluate_bbox_ap([], [])["AP"] == -1.0


def test_prepare_for_coco_detection_matches_reference_semantics():
    """coco_eval.py:81-112: boxes are resized to the ORIGINAL image size, converted to xywh (+1 widths), labels are
    mapped to json category ids, empty predictions are skipped."""
    from dadetect_b200.engine.evaluation import prepare_for_coco_detection
    from dadetect_b200.structures import BoxList

    class DS(object):
        id_to_img_map = {0: 11, 1: 22, 2: 33}
        contiguous_category_id_to_json_id = {1: 24, 2: 25}

        def get_img_info(self, i):
            return {"width": 2048, "height": 1024}

    p0 = BoxList(torch.tensor([[100.0, 50.0, 299.0, 149.0]]), (1200, 600), mode="xyxy")
    p0.add_field("scores", torch.tensor([0.75]))
    p0.add_field("labels", torch.tensor([2]))
    p1 = BoxList(torch.zeros((0, 4)), (1200, 600), mode="xyxy")
    p1.add_field("scores", torch.zeros(0))
    p1.add_field("labels", torch.zeros(0, dtype=torch.int64))
    out = prepare_for_coco_detection([p0, p1], DS())
    assert len(out) == 1 and out[0]["image_id"] == 11 and out[0]["category_id"] == 25 and out[0]["score"] == 0.75
    s = 2048 / 1200
    x1, y1, x2, y2 = 100 * s, 50 * s, 299 * s, 149 * s
    np.testing.assert_allclose(out[0]["bbox"], [x1, y1, x2 - x1 + 1, y2 - y1 + 1], rtol=1e-6)
