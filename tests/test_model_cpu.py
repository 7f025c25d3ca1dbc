"""CPU: the PYTHON orchestration of the training path (modeling/detector.py, rpn.py, roi_heads.py, da_heads.py on
the host-driven control flow) against the CPU oracle, with the kernels replaced by torch stand-ins
(tests/cpu_ops_emulation.py — a test fixture; the product has no CPU path).  What this pins every round without a
GPU: call order, loss-dict keys and weights, GRL signs, which tensors feed which head, sampler / RNG draw order
(the oracle's recorded draws are replayed and must be consumed exactly), de-duplicated passes adding up to the
reference's gradients.  The kernels themselves are pinned on the GPU (tests/test_gpu_*.py)."""
import os

import pytest
import torch

import da_frcnn_ref as orc
from cpu_ops_emulation import cpu_ops  # noqa: F401  (pytest fixture)
from make_golden import SCENARIOS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scenario(name):
    from dadetect_b200.config import get_cfg_defaults
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    yaml_name, opts, n, H, W, m = SCENARIOS[name]
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(ROOT, "configs", yaml_name))
    cfg.merge_from_list(list(opts))
    sd = make_state_dict(orc.param_shapes(cfg))
    images, targets = make_batch(n, H, W, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, boxes_per_image=m)
    return cfg, sd, images, targets, (H, W)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name,static,early", [("da_img_ins_cst", False, False), ("da_img_ins_cst", True, False),
                                               ("da_img_ins_cst", True, True),
                                               ("triplet_aligned_advgrl", True, True)])
def test_training_orchestration_matches_oracle(name, static, early, cpu_ops):
    """static=False: the reference-like host-driven control flow; static=True: the production sync-free path
    (fixed-capacity proposal / ROI buffers with validity masks, device sampler on replayed keys) — its second
    stream is switched off here, streams being a CUDA notion.  early=True: the RPN losses are back-propagated during
    the forward pass (what FlatSGDTrainer's step graph runs): same loss dict, same total gradients."""
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    from dadetect_b200.utils.random_source import ReplaySource
    torch.set_num_threads(os.cpu_count())
    cfg, sd, images, targets, (H, W) = scenario(name)
    torch.manual_seed(77)
    rec = orc.RecordingHooks()
    P = {k: v.clone().requires_grad_(orc.is_trainable(k)) for k, v in sd.items()}
    want = orc.forward_train(P, cfg, images, targets, hooks=rec, nms_strict=True)
    sum(want.values()).backward()

    model = build_detection_model(cfg)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("cell_anchors" in k for k in missing.missing_keys)
    model.train()
    model.enable_static_shapes(static)
    model.early_backward = early
    model.rpn.overlap_loss = False
    replay = ReplaySource(rec.perms, rec.masks)
    model.set_random_source(replay)
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].clone(), (W, H), mode="xyxy")
        b.add_field("labels", t["labels"].clone())
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool))
        tg.append(b)
    got = model(images, tg)
    assert list(got.keys()) == list(want.keys())
    assert not replay.perms and not replay.masks            # every recorded draw consumed, in order
    for k in want:
        g, w = float(got[k].detach()), float(want[k].detach())
        assert abs(g - w) <= 2e-5 * max(1.0, abs(w)), (k, g, w)
    sum(got.values()).backward()
    params = dict(model.named_parameters())
    checked = 0
    for k, p in P.items():
        if not p.requires_grad:
            continue
        if p.grad is None:
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
            continue
        assert params[k].grad is not None, k
        a, b = params[k].grad.double(), p.grad.double()
        assert float((a - b).norm()) <= 1e-3 * float(b.norm()) + 1e-9, (k, float((a - b).norm()), float(b.norm()))
        checked += 1
    assert checked > 50


@pytest.mark.timeout(900)
@pytest.mark.parametrize("static", [False, True])
def test_images_of_different_unpadded_sizes_in_one_batch(static, cpu_ops):
    """A batch that mixes image sizes (cityscapes -> kitti, multi-scale MIN_SIZE_TRAIN): to_image_list pads to the
    common size (SIZE_DIVISIBILITY 32), proposals are clipped to and anchor visibility is evaluated against each
    image's OWN un-padded size (anchor_generator.py:113-125, rpn/inference.py:101-103)."""
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList, to_image_list
    from dadetect_b200.utils.random_source import ReplaySource
    from dadetect_b200.utils.synthetic import make_batch
    torch.set_num_threads(os.cpu_count())
    cfg, sd, _, _, _ = scenario("da_img_ins_cst")
    sizes = [(150, 250), (136, 200)]                         # (h, w): padded to 160 x 256
    imgs, tg_raw = [], []
    for i, (h, w) in enumerate(sizes):
        im, t = make_batch(2, h, w, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, boxes_per_image=4, seed=40 + i)
        imgs.append(im[i])
        tg_raw.append(t[i])
    il = to_image_list(imgs, 32)
    assert tuple(il.tensors.shape[-2:]) == (160, 256)
    torch.manual_seed(3)
    rec = orc.RecordingHooks()
    P = {k: v.clone().requires_grad_(orc.is_trainable(k)) for k, v in sd.items()}
    with torch.no_grad():
        want = orc.forward_train(P, cfg, il.tensors, tg_raw, hooks=rec, nms_strict=True, image_sizes=il.image_sizes)
    model = build_detection_model(cfg)
    model.load_state_dict(sd, strict=False)
    model.train()
    model.enable_static_shapes(static)
    model.rpn.overlap_loss = False
    replay = ReplaySource(rec.perms, rec.masks)
    model.set_random_source(replay)
    tg = []
    for t, (h, w) in zip(tg_raw, sizes):
        b = BoxList(t["boxes"].clone(), (w, h), mode="xyxy")
        b.add_field("labels", t["labels"].clone())
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool))
        tg.append(b)
    got = model(il, tg)
    assert not replay.perms and not replay.masks
    for k in want:
        g, w_ = float(got[k].detach()), float(want[k].detach())
        assert abs(g - w_) <= 2e-5 * max(1.0, abs(w_)), (k, g, w_)
    # and the answer does depend on the sizes: with both images taken as 160 x 256 other anchors are visible
    a0, a1 = {}, {}
    with torch.no_grad():
        plain = {k: v.clone() for k, v in sd.items()}
        orc.forward_train(plain, cfg, il.tensors, tg_raw, nms_strict=True, aux=a0, image_sizes=il.image_sizes)
        orc.forward_train(plain, cfg, il.tensors, tg_raw, nms_strict=True, aux=a1)
    assert not torch.equal(a0["rpn_labels"], a1["rpn_labels"])


@pytest.mark.timeout(900)
def test_eval_orchestration_matches_real_reference_golden_c4(cpu_ops):
    """BASELINE configs[0] (plain R-50-C4 Faster R-CNN, eval mode, 2 x 800x800): the product's eval control flow
    (test-mode RPN post-processing, box head, PostProcessor) with kernel stand-ins against the detections of the REAL
    reference model (tests/golden/eval_faster_rcnn_c4.pt)."""
    from dadetect_b200.config import get_cfg_defaults
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    torch.set_num_threads(os.cpu_count())
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "eval_faster_rcnn_c4.pt"), weights_only=False)
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(ROOT, "configs", fx["yaml"]))
    sd = make_state_dict(orc.param_shapes(cfg))
    for k, f in fx["scale"].items():
        sd[k] = sd[k] * f
    images, _ = make_batch(2, fx["height"], fx["width"], num_classes=81, boxes_per_image=1, seed=fx["seed"])
    model = build_detection_model(cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    with torch.no_grad():
        out = model(images)
    assert len(out) == len(fx["detections"])
    for got, want in zip(out, fx["detections"]):
        gb, gs, gl = got.bbox, got.get_field("scores"), got.get_field("labels")
        assert abs(len(gs) - len(want["scores"])) <= 2
        hits = 0
        for i in range(len(want["scores"])):
            hits += int(((gl == want["labels"][i]) & ((gb - want["boxes"][i]).abs().max(dim=1)[0] < 0.05)
                         & ((gs - want["scores"][i]).abs() < 2e-4)).any())
        assert hits >= len(want["scores"]) - 3, (hits, len(want["scores"]))
