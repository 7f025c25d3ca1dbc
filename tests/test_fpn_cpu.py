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


def proposal_batch(boxlists):
    """list[BoxList] (with `objectness`) -> the fixed-capacity ProposalBatch of the sync-free path."""
    from dadetect_b200.modeling.rpn import ProposalBatch
    cap = max(len(b) for b in boxlists)
    n = len(boxlists)
    boxes, obj = torch.zeros(n, cap, 4), torch.zeros(n, cap)
    for i, b in enumerate(boxlists):
        boxes[i, : len(b)], obj[i, : len(b)] = b.bbox, b.get_field("objectness")
    return ProposalBatch(boxes, obj, torch.tensor([len(b) for b in boxlists], dtype=torch.int32),
                         [b.size for b in boxlists])


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


def test_fpn_triplet_module_is_refused_and_fpn_da_heads_have_the_reference_names():
    """The reference has no FPN variant of the triplet module; its FPN DA module (da_heads_fpn.py, not importable) names
    its parameters da_img_conv{1,2}_level{0..4} (:44-57) and da_ins_fc{1,2,3}_level{0..3} (:164-180)."""
    from dadetect_b200.modeling import build_detection_model
    with pytest.raises(NotImplementedError):
        build_detection_model(fpn_cfg(["MODEL.DOMAIN_ADAPTATION_ON", True, "MODEL.DA_HEADS.TRIPLET_USE", True]))
    model = build_detection_model(fpn_cfg(["MODEL.BACKBONE.CONV_BODY", "R-50-FPN", "MODEL.DOMAIN_ADAPTATION_ON", True,
                                           "MODEL.DA_HEADS.TRIPLET_USE", False]))
    names = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("da_heads.")}
    want = {}
    for i in range(5):
        want["da_heads.imghead.da_img_conv1_level%d.weight" % i] = (512, 256, 1, 1)
        want["da_heads.imghead.da_img_conv1_level%d.bias" % i] = (512,)
        want["da_heads.imghead.da_img_conv2_level%d.weight" % i] = (1, 512, 1, 1)
        want["da_heads.imghead.da_img_conv2_level%d.bias" % i] = (1,)
    for i in range(4):
        for j, (cin, cout) in enumerate(((1024, 1024), (1024, 1024), (1024, 1)), 1):
            want["da_heads.inshead.da_ins_fc%d_level%d.weight" % (j, i)] = (cout, cin)
            want["da_heads.inshead.da_ins_fc%d_level%d.bias" % (j, i)] = (cout,)
    assert names == want


@pytest.mark.timeout(900)
@pytest.mark.parametrize("static", [False, True])
def test_product_fpn_da_training_orchestration_matches_oracle(static, cpu_ops):
    """FPN + DA (BASELINE configs[4]; parity unpinned, see oracle/fpn_ref.py): the product's python path — per-level
    image heads on the GRL'd pyramid, per-level instance heads routed by the pooler's LevelMapper levels, image BCE over
    all levels, consistency over the list of levels, source-only detection losses — with kernel stand-ins against
    oracle/fpn_ref.py::forward_train_fpn_da on replayed draws: losses and the gradient of every trainable tensor."""
    import da_frcnn_ref as orc
    import fpn_ref
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    from dadetect_b200.utils.random_source import ReplaySource
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    torch.set_num_threads(os.cpu_count())
    cfg = fpn_cfg(["MODEL.BACKBONE.CONV_BODY", "R-50-FPN", "MODEL.ROI_BOX_HEAD.NUM_CLASSES", 9,
                   "MODEL.RPN.FPN_POST_NMS_TOP_N_TRAIN", 4000, "MODEL.DOMAIN_ADAPTATION_ON", True,
                   "MODEL.DA_HEADS.TRIPLET_USE", False])
    model = build_detection_model(cfg)
    sd = make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    sd["rpn.head.cls_logits.weight"] = sd["rpn.head.cls_logits.weight"] * 20.0
    for k in sd:                                            # sensitive DA losses (std 0.001 heads sit at log 2 otherwise)
        if ".da_img_conv2_" in k:
            sd[k] = sd[k] * 100.0
        if ".da_ins_fc3_" in k:
            sd[k] = sd[k] * 10.0
    model.load_state_dict(sd, strict=False)
    model.train()
    H, W = 288, 416
    images, targets = make_batch(2, H, W, num_classes=9, boxes_per_image=4, seed=33)      # [source, target]
    torch.manual_seed(5)
    rec = orc.RecordingHooks()
    frozen = ("backbone.body.stem.", "backbone.body.layer1.")
    P = {k: v.clone().requires_grad_(v.is_floating_point() and ".bn" not in k and ".downsample.1." not in k
                                     and not k.startswith(frozen)) for k, v in sd.items()}
    want = fpn_ref.forward_train_fpn_da(P, cfg, images, targets, rec)
    assert set(want) == {"loss_classifier", "loss_box_reg", "loss_objectness", "loss_rpn_box_reg", "loss_da_image",
                         "loss_da_instance", "loss_da_consistency"}
    sum(want.values()).backward()
    # The proposals are the oracle's (equal-objectness proposals may come out of the top-k in either order, and the
    # dropout masks of the instance heads are positional): handed in through the proposal hook.
    with torch.no_grad():
        pyramid = fpn_ref.fpn_forward(fpn_ref.resnet_body_all_stages(images, sd, "R-50-FPN"), sd)
        props = fpn_ref.rpn_fpn_proposals(pyramid, sd, cfg, [(H, W)] * 2, training=True, nms_strict=True)
    forced = []
    for (b, s_), t in zip(props, targets):
        if t["is_source"]:
            b, s_ = torch.cat([b, t["boxes"]]), torch.cat([s_, torch.ones(len(t["boxes"]))])
        bl = BoxList(b, (W, H), mode="xyxy")
        bl.add_field("objectness", s_)
        forced.append(bl)
    # static: the sync-free path (fixed-capacity proposals / ROI slots, per-level DA heads on all slots with level
    # masks) that FlatSGDTrainer's step graph captures
    model.enable_static_shapes(static)
    model.rpn.set_proposal_hook((lambda props: proposal_batch(forced)) if static else (lambda boxes: forced))
    replay = ReplaySource(rec.perms, rec.masks)
    model.set_random_source(replay)
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].clone(), (W, H), mode="xyxy")
        b.add_field("labels", t["labels"].clone())
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool))
        tg.append(b)
    got = model(images, tg)
    assert set(got) == set(want)
    assert not replay.perms and not replay.masks
    for k in want:
        g, w = float(got[k].detach()), float(want[k].detach())
        assert abs(g - w) <= 2e-5 * max(1.0, abs(w)), (k, g, w)
    sum(got.values()).backward()
    checked = 0
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert P[k].requires_grad, k
        if P[k].grad is None:                               # an instance head of a level without ROIs
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        a, b = p.grad.double(), P[k].grad.double()
        assert float((a - b).norm()) <= 1e-3 * float(b.norm()) + 1e-9, (k, float((a - b).norm()), float(b.norm()))
        checked += 1
    assert checked > 80
    levels = model.roi_heads.box.feature_extractor.pooler.last_levels
    assert int((torch.bincount(levels.to(torch.int64).clamp(min=0), minlength=4) > 0).sum()) >= 2   # several instance heads in use


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


# ---------------------------------------------------------------------------------------------------------------------
# The PYTHON control flow of the product's FPN path (modeling/backbone.py::FPN, rpn.py::proposals_fpn / forward_fpn,
# roi_heads.py::MultiLevelPooler / FPN2MLPFeatureExtractor / FPNPredictor, inference.py) with the kernels replaced by
# torch stand-ins (tests/cpu_ops_emulation.py): wiring, ordering and selection logic are checked on CPU every round;
# the kernels themselves are checked on the GPU (tests/test_gpu_zfpn.py).
from cpu_ops_emulation import cpu_ops  # noqa: E402,F401  (pytest fixture)


@pytest.mark.timeout(900)
def test_product_fpn_eval_control_flow_matches_real_reference_golden(fx, cpu_ops):
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    cfg = fpn_cfg(fx["opts"])
    sd = make_state_dict(fx["shapes"])
    for k, f in fx["scale"].items():
        sd[k] = sd[k] * f
    images, _ = make_batch(2, fx["height"], fx["width"], num_classes=81, boxes_per_image=1, seed=fx["seed"])
    model = build_detection_model(cfg)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("cell_anchors" in k for k in missing.missing_keys)
    model.eval()
    seen = {}
    model.rpn.set_proposal_hook(lambda boxes: seen.setdefault("p", boxes) or boxes)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = model(images)
    for got, want in zip(seen["p"], fx["proposals"]):
        assert len(got) == len(want["objectness"])
        torch.testing.assert_close(got.get_field("objectness"), want["objectness"], atol=1e-6, rtol=0)
        d = torch.cdist(want["boxes"].double(), got.bbox.double(), p=float("inf"))      # equal scores may swap places
        assert int((d.min(dim=1)[0] < 1e-2).sum()) >= len(want["boxes"]) - 2
    lv = model.roi_heads.box.feature_extractor.pooler.last_levels.to(torch.int64)
    assert int((torch.bincount(lv, minlength=4) - torch.bincount(fx["levels"], minlength=4)).abs().sum()) <= 2
    for got, want in zip(out, fx["detections"]):
        assert abs(len(got) - len(want["scores"])) <= 1          # product NMS is strict >, the golden's is >=
        gb, gs, gl = got.bbox, got.get_field("scores"), got.get_field("labels")
        hits = 0
        for i in range(len(want["scores"])):
            hits += int(((gl == want["labels"][i]) & ((gb - want["boxes"][i]).abs().max(dim=1)[0] < 1e-2)
                         & ((gs - want["scores"][i]).abs() < 1e-5)).any())
        assert hits >= len(want["scores"]) - 1


def test_ops_refuse_cpu_tensors_without_the_stand_ins():
    """The stand-ins above are a test fixture; the product itself has no CPU path."""
    from dadetect_b200 import ops
    with pytest.raises(RuntimeError):
        ops.upsample2x(torch.zeros(1, 2, 2, 4))
    with pytest.raises(RuntimeError):
        ops.roi_align_levels([torch.zeros(1, 4, 4, 4)], torch.zeros(1, 5), (0.25,), 7, 2)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("static", [False, True])
def test_product_fpn_training_orchestration_matches_oracle(static, cpu_ops):
    """FPN training (plain Faster R-CNN losses): the product's python path (multi-level RPN loss over concatenated
    levels, batch-wide proposal cut, GT append, box-head sampling, multi-level pooling, MLP head, losses) with kernel
    stand-ins against oracle/fpn_ref.py::forward_train_fpn on replayed random draws: losses and gradients."""
    import da_frcnn_ref as orc
    import fpn_ref
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    from dadetect_b200.utils.random_source import ReplaySource
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    torch.set_num_threads(os.cpu_count())
    cfg = fpn_cfg(["MODEL.BACKBONE.CONV_BODY", "R-50-FPN", "MODEL.ROI_BOX_HEAD.NUM_CLASSES", 9,
                   "MODEL.RPN.FPN_POST_NMS_TOP_N_TRAIN", 600])
    model = build_detection_model(cfg)
    sd = make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    sd["rpn.head.cls_logits.weight"] = sd["rpn.head.cls_logits.weight"] * 20.0
    model.load_state_dict(sd, strict=False)
    model.train()
    H, W = 128, 192
    images, targets = make_batch(2, H, W, num_classes=9, boxes_per_image=4, seed=33)
    for t in targets:
        t["is_source"] = True                               # without DA heads every image is a labelled one

    torch.manual_seed(5)
    rec = orc.RecordingHooks()
    frozen = ("backbone.body.stem.", "backbone.body.layer1.")
    P = {k: v.clone().requires_grad_(v.is_floating_point() and ".bn" not in k and ".downsample.1." not in k
                                     and not k.startswith(frozen)) for k, v in sd.items()}
    want = fpn_ref.forward_train_fpn(P, cfg, images, targets, rec)
    sum(want.values()).backward()

    model.enable_static_shapes(static)       # True: the sync-free path (device-side select_over_all_levels, ROI slots)
    replay = ReplaySource(rec.perms, rec.masks)
    model.set_random_source(replay)
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].clone(), (W, H), mode="xyxy")
        b.add_field("labels", t["labels"].clone())
        b.add_field("is_source", torch.ones(len(t["labels"]), dtype=torch.bool))
        tg.append(b)
    got = model(images, tg)
    assert set(got) == set(want)
    assert not replay.perms and not replay.masks
    for k in want:
        g, w = float(got[k].detach()), float(want[k].detach())
        assert abs(g - w) <= 2e-5 * max(1.0, abs(w)), (k, g, w)
    sum(got.values()).backward()
    params = dict(model.named_parameters())
    checked = 0
    for k, p in params.items():
        if not p.requires_grad:
            continue
        assert P[k].requires_grad, k                        # the same parameters are trainable (FREEZE_CONV_BODY_AT 2)
        assert p.grad is not None and P[k].grad is not None, k
        a, b = p.grad.double(), P[k].grad.double()
        assert float((a - b).norm()) <= 1e-3 * float(b.norm()) + 1e-9, (k, float((a - b).norm()), float(b.norm()))
        checked += 1
    assert checked > 60
