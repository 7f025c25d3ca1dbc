"""CPU: host-side checks of the FPN slice against facts recorded from the REAL reference model
(tests/golden/eval_faster_rcnn_r101_fpn.pt, made by `python oracle/make_golden.py fpn`): the state dict of our
R-101-FPN model has exactly the reference's names and shapes (checkpoint compatibility), the per-level cell
anchors equal the reference's, and the DA + FPN combination the reference cannot run is refused loudly."""
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fpn_cfg(opts=()):
    from dadetect_b200.config import get_cfg_defaults
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "e2e_faster_rcnn_R_101_FPN_1x.yaml"))
    cfg.merge_from_list(list(opts))
    return cfg


@pytest.fixture(scope="module")
def fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "eval_faster_rcnn_r101_fpn.pt"), weights_only=False)


def test_state_dict_names_and_shapes_equal_reference(fx):
    from dadetect_b200.modeling import build_detection_model
    model = build_detection_model(fpn_cfg(fx["opts"]))
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ours == fx["shapes"], (sorted(set(ours) ^ set(fx["shapes"]))[:10],
                                  [k for k in ours if k in fx["shapes"] and ours[k] != fx["shapes"][k]][:10])
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert trainable > 55e6                                   # R-101-FPN: ~60 M parameters, stem + res2 frozen


def test_cell_anchors_per_level_match_generate_anchors():
    """anchor_generator.py:47-66: one size per level, strides (4, 8, 16, 32, 64); values from the reference's
    generate_anchors (ratio enumeration with np.round) — level 2 is the classic 128-px row of the table at
    anchor_generator.py:201-219 scaled to stride 16."""
    from dadetect_b200.modeling.rpn import AnchorGenerator
    ag = AnchorGenerator((32, 64, 128, 256, 512), (0.5, 1.0, 2.0), (4, 8, 16, 32, 64), 0)
    cells = list(ag.cell_anchors)
    assert [tuple(c.shape) for c in cells] == [(3, 4)] * 5
    assert cells[2].tolist() == [[-84., -40., 99., 55.], [-56., -56., 71., 71.], [-36., -80., 51., 95.]]
    assert cells[4].tolist() == [[-332., -152., 395., 215.], [-224., -224., 287., 287.], [-148., -328., 211., 391.]]
    with pytest.raises(RuntimeError):
        AnchorGenerator((32, 64), (1.0,), (4, 8, 16), 0)


def test_fpn_with_da_heads_is_refused():
    from dadetect_b200.modeling import build_detection_model
    with pytest.raises(NotImplementedError):
        build_detection_model(fpn_cfg(["MODEL.DOMAIN_ADAPTATION_ON", True]))


def test_golden_exercises_every_pyramid_level(fx):
    assert torch.bincount(fx["levels"], minlength=4).min() > 0
    assert [p["shape"][2:] for p in fx["pyramid"]] == [(120, 160), (60, 80), (30, 40), (15, 20), (8, 10)]
    assert all(len(p["objectness"]) == 1200 for p in fx["proposals"])      # the select_over_all_levels cut is active


@pytest.mark.timeout(600)
def test_fpn_oracle_matches_real_reference_golden(fx):
    """oracle/fpn_ref.py (the CPU restatement of the FPN eval path) against the recorded outputs of the REAL
    reference R-101-FPN model: pyramid probes bit-level close, identical per-ROI levels, the same proposals and the
    same detections.  nms_strict=False: the golden was made with the reference's CPU NMS (>=)."""
    import fpn_ref
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    cfg = fpn_cfg(fx["opts"])
    P = make_state_dict(fx["shapes"])
    for k, f in fx["scale"].items():
        P[k] = P[k] * f
    images, _ = make_batch(2, fx["height"], fx["width"], num_classes=81, boxes_per_image=1, seed=fx["seed"])
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = fpn_ref.forward_eval_fpn(P, cfg, images, nms_strict=False)
    for t, want in zip(out["pyramid"], fx["pyramid"]):
        n, c, h, w = want["shape"]
        assert tuple(t.shape) == (n, c, h, w)
        for name, (y, x) in (("corner", (0, 0)), ("centre", (h // 2, w // 2)), ("last", (h - 1, w - 1))):
            torch.testing.assert_close(t[:, :8, y, x], want[name], atol=1e-5, rtol=1e-5)
    for (gb, gs), want in zip(out["proposals"], fx["proposals"]):
        assert len(gs) == len(want["objectness"])
        torch.testing.assert_close(gs, want["objectness"], atol=1e-6, rtol=0)
        torch.testing.assert_close(gb, want["boxes"], atol=1e-3, rtol=0)
    assert torch.equal(out["levels"], fx["levels"])
    for got, want in zip(out["detections"], fx["detections"]):
        assert len(got["scores"]) == len(want["scores"])
        assert torch.equal(got["labels"], want["labels"])
        torch.testing.assert_close(got["scores"], want["scores"], atol=1e-6, rtol=0)
        torch.testing.assert_close(got["boxes"], want["boxes"], atol=1e-3, rtol=0)
