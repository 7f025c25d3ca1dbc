"""TEST INFRASTRUCTURE — parity of ONE full-size training step (BASELINE.json configs[1..3]: 2 or 3 synthetic
1024 x 2048 images, 20 GT boxes each, 12 000 pre-NMS / 2 000 post-NMS proposals, 256 ROIs per image) of the product on
the GPU against the CPU oracle (oracle/da_frcnn_ref.py, pinned to the real reference) with the oracle's random draws
replayed.  These are the shapes bench.py times: the 64 x 128 feature map with its 1 x 8 x 16 tile mapping, K = 9216
in the RPN conv, 512-ROI res5.  Used by tests/test_gpu_fullsize.py and by `bench.py --check` (outside any timed
region).  The reference loop being matched: engine/trainer.py:228-239 over generalized_rcnn.py:79-153.

Method.  Two fp32 implementations of a 12 000-way score sort never agree on the order of scores that differ by a few
ulps (neither do the reference's own CPU and GPU paths), and one flipped pair can change an NMS survivor.  The step
is therefore pinned in three parts, each of them exact or tight:
  1. arithmetic   — the RPN head outputs (objectness logits, box deltas: the end of the dense trunk) against the
                    oracle's, fp32-grade tolerance;
  2. decisions    — the oracle's proposal procedure (top-k, decode, clip, small-box filter, NMS, top-n, GT append)
                    applied ON THE CPU TO THE PRODUCT'S OWN LOGITS must reproduce the product's proposals position
                    by position: no arithmetic noise is involved, so this is exact (boxes to 2e-3 px: expf);
                    candidates are ranked by the logit, see proposals_by_logit_order;
  3. downstream   — with the oracle's proposals handed to the product, everything after them (box-head sampling with
                    replayed draws, ROIAlign, res5, predictor, DA heads, every loss) against the oracle: losses within
                    1e-4, sampled ROIs and labels identical.
The end-to-end run on the product's own decisions is reported beside it (`losses_with_own_decisions`).
The seeded synthetic weights are conditioned (HEAD_SCALE) so that scores and losses are as spread and as sensitive
as a trained model's: at plain random initialisation every objectness score sits within 0.04 of 0.5 and the
classification loss is log(9) whatever ROI is sampled.
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


NEAR_TIE = 2.0e-7     # objectness gap (about 3 fp32 ulps of a 0.5 .. 1 score) below which two proposals count as tied
HEAD_SCALE = {"rpn.head.cls_logits.": 40.0, "rpn.head.bbox_pred.": 4.0, "roi_heads.box.predictor.cls_score.": 30.0,
              "roi_heads.box.predictor.bbox_pred.": 30.0, ".imghead.conv2_da.": 300.0, ".inshead.fc3_da.": 5.0}


def conditioned_state_dict(shapes):
    from dadetect_b200.utils.synthetic import make_state_dict
    sd = make_state_dict(shapes)
    for k in sd:
        for frag, f in HEAD_SCALE.items():
            if frag in k:
                sd[k] = sd[k] * f
    return sd


def oracle_proposal_batch(oracle_props, like):
    """The oracle's per-image proposals [(boxes [P,4], objectness [P])] as a fixed-capacity ProposalBatch shaped like
    `like` (the product's own batch): rows beyond the count are zero."""
    from dadetect_b200.modeling.rpn import ProposalBatch
    boxes, obj = torch.zeros_like(like.boxes), torch.zeros_like(like.objectness)
    cnt = torch.zeros_like(like.count)
    for i, (b, s_) in enumerate(oracle_props):
        n = min(len(b), boxes.shape[1])
        boxes[i, :n] = b[:n].to(boxes.device)
        obj[i, :n] = s_[:n].to(obj.device)
        cnt[i] = n
    return ProposalBatch(boxes, obj, cnt, like.sizes)


def compare_proposals(got, oracle_props):
    """Position-by-position comparison of the product's proposals with the oracle's.  Returns (mismatched positions,
    the largest objectness gap at a mismatched position, counts equal, boxes present in only one of the two lists).  The reference's top-k order of (nearly)
    equal scores is unspecified (SURVEY §10.3 "Ties"): a mismatch whose scores differ by less than NEAR_TIE is a
    swap inside a tie group, not an arithmetic disagreement."""
    mism, gap, counts_ok, set_diff = 0, 0.0, True, 0
    cnt = got.count.tolist()
    for i, (b, s_) in enumerate(oracle_props):
        if cnt[i] != len(b):
            counts_ok = False
        n = min(cnt[i], len(b))
        gb, gs = got.boxes[i, :n].cpu(), got.objectness[i, :n].cpu()
        bad = (gb - b[:n]).abs().max(dim=1)[0] > 2e-3
        mism += int(bad.sum()) + abs(cnt[i] - len(b))
        if bool(bad.any()):
            gap = max(gap, float((gs - s_[:n]).abs()[bad].max()))
            # membership: boxes (rounded to 0.01 px) present in one list but not in the other
            a = set(map(tuple, (got.boxes[i, :cnt[i]].cpu() * 100).round().long().tolist()))
            r = set(map(tuple, (b * 100).round().long().tolist()))
            set_diff += len(a ^ r)
    return mism, gap, counts_ok, set_diff


def proposals_by_logit_order(orc, anchors, logits, deltas, image_sizes, pre_nms, post_nms, nms_thresh, min_size):
    """oracle.rpn_proposals (RPNPostProcessor.forward_for_single_feature_map, rpn/inference.py:76-123) with ONE
    difference: candidates are ranked by the objectness LOGIT (stable, lower index first among equal logits) instead
    of by its fp32 sigmoid.  The two rankings agree except inside groups of distinct logits whose fp32 sigmoids
    coincide, where the reference's own order is whatever its top-k happens to return; ranking by the logit is what
    dd_rpn_topk_decode does, and it takes the device's sigmoid implementation (1 ulp off the host's) out of the
    comparison.  Everything else — decode, clip, small-box filter, NMS, top-n — is the oracle's code."""
    n, a, h, w = logits.shape
    obj = orc.permute_and_flatten(logits, n, 1, h, w).view(n, -1)
    reg = orc.permute_and_flatten(deltas, n, 4, h, w)
    k = min(pre_nms, a * h * w)
    out = []
    for i in range(n):
        order = torch.sort(obj[i], descending=True, stable=True)[1][:k]
        sc = obj[i][order].sigmoid()
        props = orc.box_decode(reg[i][order], anchors[order], (1.0, 1.0, 1.0, 1.0))
        ih, iw = image_sizes[i]
        props = orc.clip_boxes(props, iw, ih)
        ws = props[:, 2] - props[:, 0] + 1
        hs = props[:, 3] - props[:, 1] + 1
        keep = torch.nonzero((ws >= min_size) & (hs >= min_size)).squeeze(1)
        props, sc = props[keep], sc[keep]
        keep = orc.nms(props, sc, nms_thresh, strict=True)
        if post_nms > 0:
            keep = keep[:post_nms]
        out.append((props[keep], sc[keep]))
    return out


def run(index, dense="mixed", with_grads=False, height=H, width=W, seed=1029, device="cuda"):
    """One step of configs[index] on the GPU vs the oracle.  Returns a report dict:
      losses      {key: (got, want, rel)}           rel = |got - want| / max(|want|, 0.05)
      rpn_labels_equal, rpn_pos_equal, rpn_neg_equal   (bit-exact index tier)
      roi_labels_equal, roi_domain_equal, roi_boxes_moved (sampled ROI boxes off by > 2e-3 px)
      arithmetic, decisions, vs_oracle_proposals, losses_with_own_decisions   (see the module docstring)
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
    sd = conditioned_state_dict(orc.param_shapes(cfg))
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
    ref_props = [(b.detach().clone(), s_.detach().clone()) for b, s_ in aux["proposals"]]
    ref = dict(rpn_labels=aux["rpn_labels"], rpn_pos=aux["rpn_pos"], rpn_neg=aux["rpn_neg"],
               logits=aux["objectness"].detach().clone(), deltas=aux["rpn_box_regression"].detach().clone(),
               anchors=aux["anchors"].clone())
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
        seen, heads = [], {}
        model.rpn.set_proposal_hook(lambda p: seen.append(p) or p)
        hook = model.rpn.head.register_forward_hook(lambda m_, i_, o_: heads.update(logits=o_[0].detach(), deltas=o_[1].detach()))
        x_dev = images.to(dev)
        # ---- the arm's own decisions, end to end (informational: see the module docstring)
        try:
            got = model(x_dev, tg)
            own = {k: float(v) for k, v in got.items()}
            del got
        except AssertionError as e:          # the replayed draws went out of step: a candidate set differs in size
            own = {"error": str(e)[:200]}
        hook.remove()
        R = cfg.MODEL.RPN
        # ---- part 1: arithmetic of the dense trunk + RPN head
        lg = heads["logits"].permute(0, 3, 1, 2).contiguous().cpu()          # [N, A, FH, FW] like the reference
        dl = heads["deltas"].permute(0, 3, 1, 2).contiguous().cpu()          # [N, 4A, FH, FW]

        def rel(a, b):
            rms = float(b.double().pow(2).mean().sqrt())
            return float((a.double() - b.double()).pow(2).mean().sqrt()) / rms, float((a - b).abs().max()) / rms
        arith = dict(zip(("logits_rms_rel", "logits_max_rel"), rel(lg, ref["logits"])))
        arith.update(zip(("deltas_rms_rel", "deltas_max_rel"), rel(dl, ref["deltas"])))
        # ---- part 2: the oracle's decision procedure on the product's own logits == the product's proposals
        with torch.no_grad():
            props2 = proposals_by_logit_order(orc, ref["anchors"], lg, dl, [(height, width)] * n, R.PRE_NMS_TOP_N_TRAIN,
                                              R.POST_NMS_TOP_N_TRAIN, R.NMS_THRESH, R.MIN_SIZE)
            props2 = [(torch.cat([b, t["boxes"]]), torch.cat([s_, torch.ones(len(t["boxes"]))])) if t["is_source"]
                      else (b, s_) for (b, s_), t in zip(props2, targets)]
        d_mism, d_gap, d_counts, d_set = compare_proposals(seen[0], props2)
        # (and, informational, against the oracle's own proposals: differences here are arithmetic noise in the scores)
        mism, gap, counts_ok, set_diff = compare_proposals(seen[0], ref_props)
        # ---- part 3: everything downstream of the proposals, on the oracle's proposals
        replay = ReplaySource(rec.perms, rec.masks)
        model.set_random_source(replay)
        model.rpn.set_proposal_hook(lambda p: oracle_proposal_batch(ref_props, p))
        got = model(x_dev, tg)
        if with_grads:
            sum(got.values()).backward()
        torch.cuda.synchronize()
        gpu_s = time.perf_counter() - t0
        rep = dict(config=index, dense=dense, shape=[n, height, width], oracle_s=oracle_s, gpu_s=gpu_s,
                   keys_equal=list(got.keys()) == list(want.keys()), draws_consumed=not replay.perms and not replay.masks,
                   arithmetic=arith,
                   decisions=dict(positions_differing=d_mism, max_score_gap_at_difference=d_gap, counts_equal=d_counts,
                                  membership_differences=d_set),
                   vs_oracle_proposals=dict(positions_differing=mism, max_score_gap_at_difference=gap,
                                            counts_equal=counts_ok, membership_differences=set_diff),
                   losses_with_own_decisions=own)
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
                moved += int(((p_.bbox.cpu() - s_["boxes"]).abs().max(dim=1)[0] > 2e-3).sum())
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
    a = rep["arithmetic"]
    if a["logits_rms_rel"] > 2e-5 or a["deltas_rms_rel"] > 2e-5 or a["logits_max_rel"] > 5e-4 or a["deltas_max_rel"] > 5e-4:
        bad.append("RPN head outputs differ from the oracle beyond fp32 grade: {}".format(a))
    d = rep["decisions"]
    if not d["counts_equal"] or d["membership_differences"] > 0 or d["positions_differing"] > 0:
        bad.append("proposal decisions on identical logits differ from the oracle procedure: {}".format(d))
    if rep.get("roi_boxes_moved", 0) > 0:
        bad.append("{} sampled ROI boxes differ from the oracle's".format(rep["roi_boxes_moved"]))
    return bad
