"""TEST INFRASTRUCTURE — parity of ONE full-size training step (BASELINE.json configs[1..3]: 2 or 3 synthetic
1024 x 2048 images, 20 GT boxes each, 12 000 pre-NMS / 2 000 post-NMS proposals, 256 ROIs per image) of the product on
the GPU against the CPU oracle (oracle/da_frcnn_ref.py, pinned to the real reference) with the oracle's random draws
replayed.  These are the shapes bench.py times: the 64 x 128 feature map with its 1 x 8 x 16 tile mapping, K = 9216
in the RPN conv, 512-ROI res5.  Used by tests/test_gpu_fullsize.py and by `bench.py --check` (outside any timed
region).  The reference loop being matched: engine/trainer.py:228-239 over generalized_rcnn.py:79-153.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W = 1024, 2048
_DA = "da_faster_rcnn/e2e_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml"
_TRI = "da_faster_rcnn/e2e_triplet_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml"
# BASELINE.json configs index -> (yaml, overrides, images per GPU)
CONFIGS = {
    1: (_DA, ["MODEL.DA_HEADS.DA_INS_LOSS_WEIGHT", 0.0, "MODEL.DA_HEADS.DA_CST_LOSS_WEIGHT", 0.0], 2),
    2: (_DA, [], 2),
    3: (_TRI, ["MODEL.DA_HEADS.ALIGNMENT", True, "MODEL.DA_HEADS.DA_TRIPLET_INS_WEIGHT", 1.0], 3),
}
GRAD_PROBES = ["backbone.body.layer2.0.conv1.weight", "backbone.body.layer3.5.conv3.weight", "rpn.head.conv.weight",
               "rpn.head.cls_logits.weight", "roi_heads.box.feature_extractor.head.layer4.0.conv1.weight",
               "roi_heads.box.feature_extractor.head.layer4.2.conv2.weight", "roi_heads.box.predictor.cls_score.weight",
               "roi_heads.box.predictor.bbox_pred.weight"]


def load_cfg(index):
    from dadetect_b200.config import get_cfg_defaults
    yaml_name, opts, n = CONFIGS[index]
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(ROOT, "configs", yaml_name))
    cfg.merge_from_list(list(opts))
    return cfg, n


def run(index, dense="mixed", with_grads=False, height=H, width=W, seed=1029, device="cuda"):
    """One step of configs[index] on the GPU vs the oracle.  Returns a report dict:
      losses      {key: (got, want, rel)}           rel = |got - want| / max(|want|, 0.05)
      rpn_labels_equal, rpn_pos_equal, rpn_neg_equal   (bit-exact index tier)
      roi_labels_equal, roi_domain_equal, roi_boxes_moved (boxes off by > 2e-3 px that are not equal-score ties)
      grads       {name: rel-L2}  (with_grads only), grad_global
      oracle_s, gpu_s
    """
    import da_frcnn_ref as orc
    from dadetect_b200 import ops
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    from dadetect_b200.utils.random_source import ReplaySource
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    impl = {"simt": ops.IMPL_SIMT, "tcgen05": ops.IMPL_TCGEN05, "tcgen05x3": ops.IMPL_TCGEN05_X3,
            "mixed": ops.IMPL_TCGEN05_MIXED}[dense]
    cfg, n = load_cfg(index)
    sd = make_state_dict(orc.param_shapes(cfg))
    images, targets = make_batch(n, height, width, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, seed=seed)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(77)
    rec = orc.RecordingHooks()
    P = {k: v.clone().requires_grad_(with_grads and orc.is_trainable(k)) for k, v in sd.items()}
    aux = {}
    t0 = time.perf_counter()
    if with_grads:
        want = orc.forward_train(P, cfg, images, targets, hooks=rec, nms_strict=True, aux=aux)
        sum(want.values()).backward()
    else:
        with torch.no_grad():
            want = orc.forward_train(P, cfg, images, targets, hooks=rec, nms_strict=True, aux=aux)
    oracle_s = time.perf_counter() - t0
    want = {k: float(v) for k, v in want.items()}
    ref_samples = aux.get("samples")
    ref = dict(rpn_labels=aux["rpn_labels"], rpn_pos=aux["rpn_pos"], rpn_neg=aux["rpn_neg"])
    ref_grads = {k: P[k].grad.clone() for k in P if with_grads and P[k].grad is not None}
    del aux, P

    dev = torch.device(device)
    prev = ops.get_default_impl()
    ops.set_default_impl(impl)
    try:
        model = build_detection_model(cfg).to(dev)
        model.load_state_dict(sd, strict=False)
        model.train()
        replay = ReplaySource(rec.perms, rec.masks)
        model.set_random_source(replay)
        model.rpn.keep_debug = True
        tg = []
        for t in targets:
            b = BoxList(t["boxes"].to(dev), (width, height), mode="xyxy")
            b.add_field("labels", t["labels"].to(dev))
            b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
            tg.append(b)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = model(images.to(dev), tg)
        if with_grads:
            sum(got.values()).backward()
        torch.cuda.synchronize()
        gpu_s = time.perf_counter() - t0
        rep = dict(config=index, dense=dense, shape=[n, height, width], oracle_s=oracle_s, gpu_s=gpu_s,
                   keys_equal=list(got.keys()) == list(want.keys()), draws_consumed=not replay.perms and not replay.masks)
        rep["losses"] = {k: (float(got[k]), want[k], abs(float(got[k]) - want[k]) / max(abs(want[k]), 0.05))
                         for k in want if k in got}
        last = model.rpn.last
        rep["rpn_labels_equal"] = bool(torch.equal(last["labels"].cpu(), ref["rpn_labels"]))
        rep["rpn_pos_equal"] = bool(torch.equal(last["pos"].cpu(), ref["rpn_pos"]))
        rep["rpn_neg_equal"] = bool(torch.equal(last["neg"].cpu(), ref["rpn_neg"]))
        if ref_samples is not None and not cfg.MODEL.DA_HEADS.ALIGNMENT:
            box = model.roi_heads.box
            sampled = box.loss_evaluator.static_proposals() if model.static_shapes else box.loss_evaluator._proposals
            lab_ok = dom_ok = len(sampled) == len(ref_samples)
            moved = 0
            for p_, s_ in zip(sampled, ref_samples):
                if len(p_) != len(s_["labels"]):
                    lab_ok = dom_ok = False
                    moved += abs(len(p_) - len(s_["labels"]))
                    continue
                lab_ok &= bool(torch.equal(p_.get_field("labels").cpu(), s_["labels"]))
                dom_ok &= bool(torch.equal(p_.get_field("domain_labels").cpu(), s_["domain_labels"]))
                diff = (p_.bbox.cpu() - s_["boxes"]).abs().max(dim=1)[0] > 2e-3
                tie = (p_.get_field("objectness").cpu() - s_["objectness"]).abs() <= 1e-6
                moved += int((diff & ~tie).sum())
            rep.update(roi_labels_equal=bool(lab_ok), roi_domain_equal=bool(dom_ok), roi_boxes_moved=moved)
        if with_grads:
            named = dict(model.named_parameters())
            num = den = 0.0
            rep["grads"] = {}
            for k, gref in ref_grads.items():
                g = named[k].grad
                a, b = g.detach().cpu().double().reshape(-1), gref.double().reshape(-1)
                num += float((a - b).pow(2).sum())
                den += float(b.pow(2).sum())
                if k in GRAD_PROBES or ".imghead." in k or ".inshead." in k:
                    rep["grads"][k] = float((a - b).norm() / (b.norm() + 1e-30))
            rep["grad_global"] = (num / max(den, 1e-30)) ** 0.5
        return rep
    finally:
        ops.set_default_impl(prev)


def verdict(rep, loss_tol=1e-4):
    """List of human-readable failures (empty = pass)."""
    bad = []
    if not rep["keys_equal"]:
        bad.append("loss-dict keys differ")
    if not rep["draws_consumed"]:
        bad.append("recorded random draws were not consumed exactly")
    for k, (g, w, rel) in rep["losses"].items():
        if rel > loss_tol:
            bad.append("{}: got {:.7f} want {:.7f} rel {:.2e}".format(k, g, w, rel))
    for k in ("rpn_labels_equal", "rpn_pos_equal", "rpn_neg_equal", "roi_labels_equal", "roi_domain_equal"):
        if k in rep and not rep[k]:
            bad.append(k + " is False")
    if rep.get("roi_boxes_moved", 0) > 0:
        bad.append("{} sampled ROI boxes differ beyond an equal-score tie".format(rep["roi_boxes_moved"]))
    return bad
