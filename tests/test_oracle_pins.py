"""Pins oracle/ (the CPU restatement) to the reference: its own known-answer vectors and
fixtures produced by the REAL reference modules (oracle/make_golden.py).  CPU only."""
import os

import pytest
import torch

import da_frcnn_ref as orc
from make_golden import SCENARIOS, unpack_masks
from dadetect_b200.config import get_cfg_defaults
from dadetect_b200.utils.synthetic import make_batch, make_state_dict

CONFIGS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


# ------------------------------------------------------------------ reference KATs
def test_nms_kat_5_boxes(golden_dir):
    k = load(golden_dir, "ref_kats.pt")
    for thr, want in zip(k["nms5_thresh"], k["nms5_keep"]):
        for strict in (False, True):
            got = orc.nms(k["nms5_boxes"], k["nms5_scores"], thr, strict=strict)
            assert got.tolist() == sorted(want), (thr, strict)


def test_nms_kat_53_boxes(golden_dir):
    k = load(golden_dir, "ref_kats.pt")
    for strict in (False, True):
        got = orc.nms(k["nms53_boxes"], k["nms53_scores"], 0.5, strict=strict)
        assert got.tolist() == k["nms53_keep"].tolist()


def test_box_decode_kat(golden_dir):
    k = load(golden_dir, "ref_kats.pt")
    got = orc.box_decode(k["coder_deltas"], k["coder_boxes"], (1.0, 1.0, 1.0, 1.0))
    torch.testing.assert_close(got, k["coder_decoded"], atol=1e-4, rtol=0)


def test_anchor_table_kat(golden_dir):
    k = load(golden_dir, "ref_kats.pt")
    got = orc.cell_anchors(16, (128, 256, 512), (0.5, 1.0, 2.0))
    # the table in the reference comment (anchor_generator.py:201-219) is the MATLAB (1-based)
    # anchors.mat dump; the 0-based code output is that table minus one pixel.
    assert torch.equal(got + 1, k["anchors_stride16_128_256_512"])


# ------------------------------------------------------------------ op-level goldens from the real reference
def test_ops_against_reference(golden_dir):
    o = load(golden_dir, "ref_ops.pt")
    assert torch.equal(orc.cell_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0)), o["cell_anchors_da"])
    gt, pr = o["iou_gt"], o["iou_pr"]
    iou = orc.box_iou(gt, pr)
    assert torch.equal(iou, o["iou"])
    assert torch.equal(orc.matcher(iou.clone(), 0.7, 0.3, True), o["match_rpn"])
    assert torch.equal(orc.matcher(iou.clone(), 0.5, 0.5, False), o["match_box"])
    for wts, tag in (((1.0, 1.0, 1.0, 1.0), "rpn"), ((10.0, 10.0, 5.0, 5.0), "box")):
        enc = orc.box_encode(gt[torch.arange(300) % 7], pr, wts)
        assert torch.equal(enc, o["encode_" + tag])
        assert torch.equal(orc.box_decode(enc * 0.7 + 0.05, pr, wts), o["decode_" + tag])
    assert torch.equal(orc.box_decode(o["decode_clamp_in"], pr, (1.0, 1.0, 1.0, 1.0)), o["decode_clamp"])
    torch.testing.assert_close(orc.roi_align(o["ra_feat"], o["ra_rois"], 1 / 16, 14, 14, 0), o["ra_out_s0"],
                               atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(orc.roi_align(o["ra_feat"], o["ra_rois"], 1 / 16, 7, 7, 2), o["ra_out_s2"],
                               atol=1e-6, rtol=1e-6)
    for thr in (0.3, 0.5, 0.7):
        assert torch.equal(orc.nms(pr, o["nms_scores"], thr, strict=False), o["nms_ge_%.1f" % thr])
    torch.testing.assert_close(orc.smooth_l1(o["sl1_x"], o["sl1_t"], 1 / 9, False), o["sl1_b9_sum"])
    torch.testing.assert_close(orc.smooth_l1(o["sl1_x"], o["sl1_t"], 1.0, False), o["sl1_b1_sum"])
    torch.testing.assert_close(orc.consistency_loss(o["cst_img"], o["cst_ins"], o["cst_dom"]), o["cst"])
    for key in ("trip4", "trip2"):
        a, p, n, want = o[key]
        margin = 1.0 if key == "trip4" else 0.7
        torch.testing.assert_close(orc.triplet_margin_loss(a, p, n, margin), want, atol=1e-6, rtol=1e-6)
    assert abs(orc.ADV_BCE - o["adv_bce"]) < 1e-7
    for L, w in o["adv_grl"]:
        assert abs(orc.adv_grl_weight(L, 0.1, 0.1, 30) - w) < 1e-6 * max(1.0, abs(w)), (L, w)


def test_roi_align_backward_matches_torchvision():
    """The reference has no CPU ROIAlign backward (csrc/ROIAlign.h:44) and its CUDA file cannot be
    built here; torchvision's roi_align(aligned=False) implements the same math."""
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 5, 12, 20, generator=g, requires_grad=True)
    rois = torch.tensor([[0, 10.0, 12.0, 150.0, 100.0], [1, -20.0, 5.0, 90.0, 250.0], [1, 200.0, 40.0, 200.3, 41.0]])
    go = torch.randn(3, 5, 14, 14, generator=g)
    for sr in (0, 2):
        out = orc.roi_align(feat, rois, 1 / 16, 14, 14, sr)
        (gin,) = torch.autograd.grad(out, feat, go)
        ref = tv.ops.roi_align(feat, rois, (14, 14), 1 / 16, sr, aligned=False)
        (gref,) = torch.autograd.grad(ref, feat, go)
        torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(gin, gref, atol=1e-5, rtol=1e-5)


def test_strict_nms_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(5)
    xy = torch.rand(400, 2, generator=g) * 200
    wh = 5 + torch.rand(400, 2, generator=g) * 80
    boxes = torch.cat([xy, xy + wh], 1)
    scores = torch.rand(400, generator=g)
    # torchvision uses areas without the +1; shift x2,y2 by +1 to express the reference's convention
    b1 = boxes.clone()
    b1[:, 2:] += 1
    want = tv.ops.nms(b1, scores, 0.7).sort()[0]
    assert torch.equal(orc.nms(boxes, scores, 0.7, strict=True), want)


# ------------------------------------------------------------------ full-path scenarios from the real reference
def scenario_cfg(fx):
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(CONFIGS, fx["yaml"]))
    cfg.merge_from_list(fx["opts"])
    return cfg


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario_losses_and_grads(golden_dir, name):
    fx = load(golden_dir, "scenario_{}.pt".format(name))
    cfg = scenario_cfg(fx)
    sd = make_state_dict(orc.param_shapes(cfg))
    P = {k: v.clone().requires_grad_(orc.is_trainable(k)) for k, v in sd.items()}
    images, targets = make_batch(fx["n_images"], fx["height"], fx["width"],
                                 num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, boxes_per_image=fx["boxes_per_image"])
    hooks = orc.ReplayHooks(fx["perms"], unpack_masks(fx["masks"]))
    losses = orc.forward_train(P, cfg, images, targets, hooks=hooks, nms_strict=False)
    assert list(losses.keys()) == fx["loss_order"]
    for k, want in fx["losses"].items():
        assert abs(float(losses[k]) - want) <= 1e-6 * max(1.0, abs(want)), (k, float(losses[k]), want)
    assert not hooks.perms and not hooks.masks, "every recorded reference draw must be consumed"
    sum(losses.values()).backward()
    for k, probe in fx["grads"].items():
        g = P[k].grad.double().reshape(-1)
        assert abs(float(g.norm()) - probe["norm"]) <= 2e-4 * probe["norm"] + 1e-12, k
        torch.testing.assert_close(g[:16].float(), probe["head"], rtol=2e-3, atol=1e-7 + 1e-4 * probe["norm"] / g.numel() ** 0.5)
    without = sorted(k for k, p in P.items() if p.requires_grad and p.grad is None)
    assert without == [k for k in fx["params_without_grad"] if k in P]
