"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt in the BUILD container by running
the REAL reference (imported from /root/reference through oracle/ref_harness.py).
The GPU box has no /root/reference, so the fixtures are committed together with this script.

  python oracle/make_golden.py            # (re)writes tests/golden/

Fixtures
  ref_kats.pt            the reference's own known-answer vectors, lifted by parsing its test
                         sources: tests/test_nms.py:11-58 and :60-217, tests/test_box_coder.py:11-105,
                         and the anchor table comment in modeling/rpn/anchor_generator.py:209-219.
  ref_ops.pt             outputs of reference sub-modules on seeded inputs: generate_anchors for the
                         DA anchor set, BoxCoder.encode/decode, boxlist_iou, Matcher (both modes),
                         compiled csrc/cpu ROIAlign forward and nms (>=), smooth_l1, consistency_loss,
                         TripletMarginLoss (4-D and 2-D), Adv_GRL weights.
  scenario_<name>.pt     one full GeneralizedRCNN.forward (training) + backward of the reference
                         model on a small synthetic batch: loss dict, recorded random draws
                         (randperm / dropout masks, bit-packed), proposal counts, sampled indices,
                         and gradient probes (norm + leading values) of selected parameters.
"""
import ast
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_harness as rh          # noqa: E402
import da_frcnn_ref as orc        # noqa: E402
from dadetect_b200.utils.synthetic import make_batch, make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SCENARIOS = {
    # name: (yaml, opts, n_images, H, W, boxes_per_image)
    "da_img_ins_cst": ("da_faster_rcnn/e2e_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml",
                       [], 2, 192, 320, 6),
    "da_img_only": ("da_faster_rcnn/e2e_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml",
                    ["MODEL.DA_HEADS.DA_INS_LOSS_WEIGHT", 0.0, "MODEL.DA_HEADS.DA_CST_LOSS_WEIGHT", 0.0],
                    2, 160, 256, 5),
    "triplet_aligned_advgrl": ("da_faster_rcnn/e2e_triplet_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml",
                               ["MODEL.DA_HEADS.ALIGNMENT", True, "MODEL.DA_HEADS.DA_TRIPLET_INS_WEIGHT", 1.0,
                                "MODEL.DA_HEADS.DA_CST_LOSS_WEIGHT", 1.0],
                               3, 160, 256, 5),
    "triplet_yaml_default": ("da_faster_rcnn/e2e_triplet_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml",
                             [], 3, 160, 256, 5),
}

GRAD_PROBES = [
    "backbone.body.layer2.0.conv1.weight", "backbone.body.layer2.3.conv2.weight",
    "backbone.body.layer3.0.downsample.0.weight", "backbone.body.layer3.5.conv3.weight",
    "rpn.head.conv.weight", "rpn.head.conv.bias", "rpn.head.cls_logits.weight", "rpn.head.bbox_pred.bias",
    "roi_heads.box.feature_extractor.head.layer4.0.conv1.weight",
    "roi_heads.box.feature_extractor.head.layer4.0.downsample.0.weight",
    "roi_heads.box.feature_extractor.head.layer4.2.conv2.weight",
    "roi_heads.box.predictor.cls_score.weight", "roi_heads.box.predictor.bbox_pred.weight",
    "{da}.imghead.conv1_da.weight", "{da}.imghead.conv2_da.bias",
    "{da}.inshead.fc1_da.weight", "{da}.inshead.fc3_da.weight",
]


# ----------------------------------------------------------------------------- KATs
def _literal_arrays(path, func_name):
    """Evaluate the `name = <np.array/list literal>` assignments inside one test function."""
    tree = ast.parse(open(path).read())
    env = {"np": np, "torch": torch}
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func_name:
            for st in node.body:
                if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                    name = st.targets[0].id
                    src = ast.get_source_segment(open(path).read(), st.value)
                    if "box_nms" in src or "box_coder" in src or "BoxCoder" in src or "np.sort" in src:
                        continue
                    try:
                        val = eval(compile(ast.Expression(st.value), path, "eval"), dict(env, **out))
                    except Exception:
                        continue
                    out[name] = val
    return out


def make_kats():
    t = os.path.join(rh.REF, "tests")
    a = _literal_arrays(os.path.join(t, "test_nms.py"), "test_nms_cpu")
    b = _literal_arrays(os.path.join(t, "test_nms.py"), "test_nms1_cpu")
    c = _literal_arrays(os.path.join(t, "test_box_coder.py"), "test_box_decoder")
    src = open(os.path.join(rh.REF, "maskrcnn_benchmark/modeling/rpn/anchor_generator.py")).read()
    m = re.search(r"# array\(\[\[(.*?)\]\]\)", src, re.S)
    nums = [float(x) for x in re.findall(r"-?\d+\.", m.group(1))]
    kats = dict(
        nms5_boxes=torch.as_tensor(a["inputs"][:, :4]), nms5_scores=torch.as_tensor(a["inputs"][:, 4]),
        nms5_thresh=list(a["test_thresh"]), nms5_keep=[list(x) for x in a["gt_indices"]],
        nms53_boxes=torch.as_tensor(b["boxes"]), nms53_scores=torch.as_tensor(b["scores"]),
        nms53_keep=torch.as_tensor(b["gt_indices"]),
        coder_boxes=torch.as_tensor(c["bbox"]), coder_deltas=torch.as_tensor(c["deltas"]),
        coder_decoded=torch.as_tensor(c["gt_bbox"]),
        anchors_stride16_128_256_512=torch.tensor(nums).view(9, 4),
    )
    assert kats["nms53_boxes"].shape == (53, 4) and kats["nms53_keep"].numel() == 26
    torch.save(kats, os.path.join(OUT, "ref_kats.pt"))
    return kats


# ----------------------------------------------------------------------------- op-level goldens
def make_ops():
    rh.install()
    from maskrcnn_benchmark import _C
    from maskrcnn_benchmark.layers import consistency_loss, smooth_l1_loss
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.matcher import Matcher
    from maskrcnn_benchmark.modeling.rpn.anchor_generator import generate_anchors
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.boxlist_ops import boxlist_iou
    g = torch.Generator().manual_seed(4242)
    o = {}
    o["cell_anchors_da"] = generate_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0)).float()

    def rand_boxes(n, w, h):
        x1 = torch.rand(n, generator=g) * w * 0.8
        y1 = torch.rand(n, generator=g) * h * 0.8
        bw = 4 + torch.rand(n, generator=g) * w * 0.4
        bh = 4 + torch.rand(n, generator=g) * h * 0.4
        return torch.stack([x1, y1, (x1 + bw).clamp(max=w - 1), (y1 + bh).clamp(max=h - 1)], 1)

    W, H = 320, 192
    gt, pr = rand_boxes(7, W, H), rand_boxes(300, W, H)
    pr[:7] = gt + 0.25                                    # a few high-IoU pairs
    o["iou_gt"], o["iou_pr"] = gt, pr
    iou = boxlist_iou(BoxList(gt, (W, H)), BoxList(pr, (W, H)))
    o["iou"] = iou
    o["match_rpn"] = Matcher(0.7, 0.3, allow_low_quality_matches=True)(iou.clone())
    o["match_box"] = Matcher(0.5, 0.5, allow_low_quality_matches=False)(iou.clone())
    for wts, tag in (((1.0, 1.0, 1.0, 1.0), "rpn"), ((10.0, 10.0, 5.0, 5.0), "box")):
        coder = BoxCoder(weights=wts)
        enc = coder.encode(gt[torch.arange(300) % 7], pr)
        o["encode_" + tag] = enc
        o["decode_" + tag] = coder.decode(enc * 0.7 + 0.05, pr)
    deltas_big = torch.randn(300, 4, generator=g) * 3.0    # exercises the log(1000/16) clamp
    o["decode_clamp_in"] = deltas_big
    o["decode_clamp"] = BoxCoder(weights=(1.0, 1.0, 1.0, 1.0)).decode(deltas_big, pr)

    feat = torch.randn(2, 8, 12, 20, generator=g)
    rois = torch.cat([torch.randint(0, 2, (40, 1), generator=g).float(), rand_boxes(40, W, H)], 1)
    rois[0, 1:] = torch.tensor([-40.0, -30.0, 500.0, 400.0])   # spills outside the map
    rois[1, 1:] = torch.tensor([50.0, 60.0, 50.2, 60.1])       # degenerate -> min size 1
    o["ra_feat"], o["ra_rois"] = feat, rois
    o["ra_out_s0"] = _C.roi_align_forward(feat, rois, 1.0 / 16, 14, 14, 0)
    o["ra_out_s2"] = _C.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 2)
    sc = torch.rand(300, generator=g)
    o["nms_scores"] = sc
    for thr in (0.3, 0.5, 0.7):
        o["nms_ge_%.1f" % thr] = _C.nms(pr, sc, thr)

    x, t = torch.randn(50, 4, generator=g), torch.randn(50, 4, generator=g) * 0.2
    o["sl1_x"], o["sl1_t"] = x, t
    o["sl1_b9_sum"] = smooth_l1_loss(x, t, beta=1.0 / 9, size_average=False)
    o["sl1_b1_sum"] = smooth_l1_loss(x, t, beta=1.0, size_average=False)

    img_sig = torch.rand(2, 1, 12, 20, generator=g)
    ins_sig = torch.rand(37, 1, generator=g)
    dom = torch.cat([torch.ones(21, dtype=torch.bool), torch.zeros(16, dtype=torch.bool)])
    o["cst_img"], o["cst_ins"], o["cst_dom"] = img_sig, ins_sig, dom
    o["cst"] = consistency_loss([img_sig], ins_sig, dom, size_average=True)

    a4, p4, n4 = (torch.randn(1, 6, 5, 8, generator=g) for _ in range(3))
    a2, p2, n2 = (torch.randn(9, 16, generator=g) for _ in range(3))
    o["trip4"] = (a4, p4, n4, torch.nn.TripletMarginLoss(margin=1.0, p=2)(a4, p4, n4))
    o["trip2"] = (a2, p2, n2, torch.nn.TripletMarginLoss(margin=0.7, p=2)(a2, p2, n2))

    # Adv_GRL weights straight from the reference method (da_heads.py:173-195)
    from maskrcnn_benchmark.modeling.da_heads.da_heads import DomainAdaptationModule_triplet
    cfg = rh.reference_cfg(SCENARIOS["triplet_yaml_default"][0])
    mod = DomainAdaptationModule_triplet(cfg)
    adv = []
    for L in (0.70, 0.62877, 0.5, 0.2, 0.05):
        mod.Adv_GRL(torch.tensor(L), [torch.zeros(1, 1, 1, 1)], list_option=True)
        w = mod.advGRL_optimized.weight if L <= float(mod.bce) else mod.grl_img.weight
        adv.append((L, float(w)))
    o["adv_grl"] = adv
    o["adv_bce"] = float(mod.bce)
    torch.save(o, os.path.join(OUT, "ref_ops.pt"))
    return o


# ----------------------------------------------------------------------------- scenarios
class _Recorder(object):
    """Captures the reference's own torch.randperm / F.dropout draws while it runs."""

    def __init__(self):
        self.perms, self.masks = [], []
        self._randperm, self._dropout = torch.randperm, torch.nn.functional.dropout

    def __enter__(self):
        rec = self

        def randperm(n, *a, **k):
            p = rec._randperm(n, *a, **k)
            rec.perms.append(p.clone())
            return p

        def dropout(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            keep = torch.empty_like(x).bernoulli_(1 - p)      # same generator consumption as F.dropout
            rec.masks.append(keep.clone())
            return x * keep / (1 - p)

        torch.randperm = randperm
        torch.nn.functional.dropout = dropout
        return self

    def __exit__(self, *exc):
        torch.randperm, torch.nn.functional.dropout = self._randperm, self._dropout


def pack_masks(masks):
    return [(tuple(m.shape), torch.from_numpy(np.packbits(m.numpy().astype(np.uint8).reshape(-1)))) for m in masks]


def unpack_masks(packed):
    out = []
    for shape, bits in packed:
        n = int(np.prod(shape))
        out.append(torch.from_numpy(np.unpackbits(bits.numpy())[:n].astype(np.float32)).view(shape))
    return out


def grad_probe(g):
    g = g.detach().double().reshape(-1)
    return dict(norm=float(g.norm()), head=g[:16].float().clone(), sum=float(g.sum()))


def make_scenario(name):
    yaml_name, opts, n, H, W, m = SCENARIOS[name]
    cfg = rh.reference_cfg(yaml_name, opts)
    sd = make_state_dict(orc.param_shapes(cfg))
    model = rh.build_reference_model(cfg, sd)
    model.train()
    images, targets = make_batch(n, H, W, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, boxes_per_image=m)
    torch.manual_seed(20231017)
    with _Recorder() as rec:
        losses = model(images, rh.to_reference_targets(targets, (H, W)))
        total = sum(losses.values())
        total.backward()
    named = dict(model.named_parameters())
    da = "da_heads_triplet" if cfg.MODEL.DA_HEADS.TRIPLET_USE else "da_heads"
    grads = {}
    for k in GRAD_PROBES:
        k = k.format(da=da)
        if named[k].grad is not None:
            grads[k] = grad_probe(named[k].grad)
    none_grad = sorted(k for k, p in named.items() if p.requires_grad and p.grad is None)
    fx = dict(
        name=name, yaml=yaml_name, opts=list(opts), n_images=n, height=H, width=W, boxes_per_image=m,
        losses={k: float(v) for k, v in losses.items()}, loss_order=list(losses.keys()),
        perms=rec.perms, masks=pack_masks(rec.masks), grads=grads, params_without_grad=none_grad,
        nms="cpu_ge",
    )
    torch.save(fx, os.path.join(OUT, "scenario_{}.pt".format(name)))
    return fx


def main():
    assert rh.available(), "reference not mounted"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    make_kats()
    make_ops()
    for name in SCENARIOS:
        fx = make_scenario(name)
        print(name, {k: round(v, 6) for k, v in fx["losses"].items()}, "perms", len(fx["perms"]),
              "masks", len(fx["masks"]), "no-grad params", len(fx["params_without_grad"]))
    print("fixture bytes:", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)))




# ----------------------------------------------------------------------------- eval-mode golden (BASELINE configs[0])
EVAL_YAML = "e2e_faster_rcnn_R_50_C4_1x.yaml"
EVAL_HW = (800, 800)
EVAL_SCALE = {"roi_heads.box.predictor.cls_score.weight": 40.0, "roi_heads.box.predictor.bbox_pred.weight": 40.0}


def eval_state_dict(shapes):
    """make_state_dict with the predictor weights amplified: with the stock std 0.01 / 0.001 initialisers every
    class score sits at 1/81 and no detection would pass SCORE_THRESH — nothing would be tested."""
    sd = make_state_dict(shapes)
    for k, f in EVAL_SCALE.items():
        sd[k] = sd[k] * f
    return sd


def make_eval():
    """The reference's OWN eval-mode forward (RPNPostProcessor test mode + box_head PostProcessor) on CPU for the
    plain R-50-C4 Faster R-CNN YAML (81 classes, DA off) on 2 synthetic 800x800 images — the one configuration the
    unpatched reference model can run end to end on CPU (BASELINE.json configs[0])."""
    cfg = rh.reference_cfg(EVAL_YAML, [])
    sd = eval_state_dict(orc.param_shapes(cfg))
    model = rh.build_reference_model(cfg, sd)
    model.eval()
    images, _ = make_batch(2, EVAL_HW[0], EVAL_HW[1], num_classes=81, boxes_per_image=1, seed=4242)
    with torch.no_grad():
        out = model(images)
    fx = dict(yaml=EVAL_YAML, height=EVAL_HW[0], width=EVAL_HW[1], seed=4242, scale=EVAL_SCALE, nms="cpu_ge",
              detections=[dict(boxes=o.bbox.clone(), scores=o.get_field("scores").clone(),
                               labels=o.get_field("labels").clone()) for o in out])
    torch.save(fx, os.path.join(OUT, "eval_faster_rcnn_c4.pt"))
    return fx


# ----------------------------------------------------------------------------- FPN eval-mode golden (BASELINE configs[4])
FPN_YAML = "e2e_faster_rcnn_R_101_FPN_1x.yaml"
FPN_HW = (480, 640)
# fewer pre-NMS candidates per level and a cut at 1500, so that proposals of every pyramid level survive the
# select_over_all_levels top-k (with the YAML's 1000/1000 the random-weight model keeps P2/P3 boxes only and the
# LevelMapper would see one level)
FPN_OPTS = ["MODEL.RPN.PRE_NMS_TOP_N_TEST", 500, "MODEL.RPN.FPN_POST_NMS_TOP_N_TEST", 1200]
FPN_SCALE = {"roi_heads.box.predictor.cls_score.weight": 12.0, "roi_heads.box.predictor.bbox_pred.weight": 40.0}


def tensor_probe(t):
    """Small fingerprint of a big activation: moments + the first channels at three pixels."""
    t = t.detach().float()
    n, c, h, w = t.shape
    return dict(shape=(n, c, h, w), mean=float(t.double().mean()), absmean=float(t.double().abs().mean()),
                corner=t[:, :8, 0, 0].clone(), centre=t[:, :8, h // 2, w // 2].clone(), last=t[:, :8, h - 1, w - 1].clone())


def make_fpn_eval():
    """The reference's OWN eval-mode forward of the R-101-FPN Faster R-CNN YAML on CPU (2 synthetic 256x320 images,
    81 classes): FPN top-down path, 5-level RPN + select_over_all_levels (test branch), LevelMapper pooling,
    FPN2MLP head, PostProcessor.  Records the detections plus probes of the pyramid and the RPN proposals."""
    import contextlib
    import io
    cfg = rh.reference_cfg(FPN_YAML, FPN_OPTS)
    rh.install()
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    with contextlib.redirect_stdout(io.StringIO()):
        model = build_detection_model(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = make_state_dict(shapes)
    for k, f in FPN_SCALE.items():
        sd[k] = sd[k] * f
    missing = model.load_state_dict(sd, strict=False)
    assert all("cell_anchors" in k for k in missing.missing_keys) and not missing.unexpected_keys
    model.eval()
    images, _ = make_batch(2, FPN_HW[0], FPN_HW[1], num_classes=81, boxes_per_image=1, seed=777)
    rec = {}
    h1 = model.backbone.register_forward_hook(lambda m, i, o: rec.__setitem__("pyramid", [tensor_probe(t) for t in o]))
    h2 = model.rpn.register_forward_hook(lambda m, i, o: rec.__setitem__(
        "proposals", [dict(boxes=b.bbox.clone(), objectness=b.get_field("objectness").clone()) for b in o[0]]))
    h3 = model.roi_heads.box.feature_extractor.pooler.register_forward_hook(
        lambda m, i, o: rec.__setitem__("levels", m.map_levels(i[1]).to(torch.int64)))
    with torch.no_grad():
        out = model(images)
    for h in (h1, h2, h3):
        h.remove()
    fx = dict(yaml=FPN_YAML, opts=FPN_OPTS, height=FPN_HW[0], width=FPN_HW[1], seed=777, scale=FPN_SCALE, nms="cpu_ge",
              shapes=shapes, pyramid=rec["pyramid"], proposals=rec["proposals"], levels=rec["levels"],
              detections=[dict(boxes=o.bbox.clone(), scores=o.get_field("scores").clone(),
                               labels=o.get_field("labels").clone()) for o in out])
    torch.save(fx, os.path.join(OUT, "eval_faster_rcnn_r101_fpn.pt"))
    return fx


# ----------------------------------------------------------------------------- config surface + structures (boundary B3 / B1)
def _flatten(node, prefix=""):
    out = {}
    for k, v in node.items():
        if hasattr(v, "items"):
            out.update(_flatten(v, prefix + k + "."))
        else:
            out[prefix + k] = list(v) if isinstance(v, tuple) else v
    return out


def make_boundary():
    """(1) The reference's config/defaults.py tree merged with each of its DA YAMLs (+ the two plain detectors this
    repo pins): YAML text and the flattened effective configuration.  (2) BoxList / ImageList behaviour of the real
    reference classes on seeded inputs (resize, transpose, convert, clip_to_image, area, to_image_list, __add__)."""
    rh.install()
    import glob
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.image_list import to_image_list
    cfgs = {}
    names = sorted(glob.glob(os.path.join(rh.REF, "configs", "da_faster_rcnn", "*.y*ml")))
    names += [os.path.join(rh.REF, "configs", n) for n in ("e2e_faster_rcnn_R_50_C4_1x.yaml",
                                                           "e2e_faster_rcnn_R_101_FPN_1x.yaml")]
    for path in names:
        rel = os.path.relpath(path, os.path.join(rh.REF, "configs"))
        c = rh.reference_cfg(rel, [])
        flat = _flatten(c)
        flat["MODEL.DEVICE"] = "cuda"                       # reference_cfg forces cpu for the CPU runs
        cfgs[rel] = dict(text=open(path).read(), effective=flat)
    g = torch.Generator().manual_seed(17)
    x1 = torch.rand(12, generator=g) * 300 - 20
    y1 = torch.rand(12, generator=g) * 200 - 20
    boxes = torch.stack([x1, y1, x1 + torch.rand(12, generator=g) * 150, y1 + torch.rand(12, generator=g) * 120], 1)
    b = BoxList(boxes.clone(), (320, 200), mode="xyxy")
    b.add_field("labels", torch.arange(12))
    st = dict(boxes=boxes, size=(320, 200))
    st["xywh"] = b.convert("xywh").bbox.clone()
    st["xywh_back"] = b.convert("xywh").convert("xyxy").bbox.clone()
    st["area"] = b.area().clone()
    st["resize_same_ratio"] = b.resize((640, 400)).bbox.clone()
    st["resize_two_ratios"] = b.resize((500, 333)).bbox.clone()
    st["resize_same_ratio_xywh"] = b.convert("xywh").resize((640, 400)).bbox.clone()    # scaled in its own mode (:99-108)
    st["resize_two_ratios_xywh"] = b.convert("xywh").resize((500, 333)).bbox.clone()
    st["flip_lr"] = b.transpose(0).bbox.clone()
    st["flip_tb"] = b.transpose(1).bbox.clone()
    cl = BoxList(boxes.clone(), (320, 200), mode="xyxy")
    cl.add_field("labels", torch.arange(12))
    cl = cl.clip_to_image(remove_empty=True)
    st["clip_boxes"], st["clip_labels"] = cl.bbox.clone(), cl.get_field("labels").clone()
    imgs = [torch.rand(3, 37, 53, generator=g), torch.rand(3, 41, 50, generator=g)]
    il = to_image_list(imgs, 32)
    il2 = to_image_list([torch.rand(3, 70, 20, generator=g)], 0)
    both = il + il2
    st["images"] = imgs + [il2.tensors[0].clone()]
    st["padded"], st["padded_sizes"] = il.tensors.clone(), [tuple(s) for s in il.image_sizes]
    st["added"], st["added_sizes"] = both.tensors.clone(), [tuple(s) for s in both.image_sizes]
    torch.save(dict(configs=cfgs, structures=st), os.path.join(OUT, "boundary_ref.pt"))
    return cfgs, st


# ----------------------------------------------------------------------------- input pipeline golden
PREPROCESS_CASES = [
    # (name, cfg opts, [(h, w) of the decoded images], random seed)
    ("down_2to1", ["INPUT.MIN_SIZE_TRAIN", (40,), "INPUT.MAX_SIZE_TRAIN", 80, "DATALOADER.SIZE_DIVISIBILITY", 32],
     [(70, 140), (70, 140)], 3),
    ("multi_scale_ragged", ["INPUT.MIN_SIZE_TRAIN", (30, 44, 52), "INPUT.MAX_SIZE_TRAIN", 90,
                            "DATALOADER.SIZE_DIVISIBILITY", 16], [(61, 97), (97, 61), (50, 50)], 11),
    ("upsample_no_divisor", ["INPUT.MIN_SIZE_TRAIN", (64,), "INPUT.MAX_SIZE_TRAIN", 200,
                             "DATALOADER.SIZE_DIVISIBILITY", 0], [(24, 37), (31, 29)], 5),
    ("identity_size", ["INPUT.MIN_SIZE_TRAIN", (48,), "INPUT.MAX_SIZE_TRAIN", 96, "DATALOADER.SIZE_DIVISIBILITY", 32],
     [(48, 96), (48, 80)], 8),
    ("max_size_rule_rgb", ["INPUT.MIN_SIZE_TRAIN", (60,), "INPUT.MAX_SIZE_TRAIN", 100, "INPUT.TO_BGR255", False,
                           "INPUT.PIXEL_MEAN", [0.485, 0.456, 0.406], "INPUT.PIXEL_STD", [0.229, 0.224, 0.225],
                           "DATALOADER.SIZE_DIVISIBILITY", 32], [(45, 130), (130, 45)], 21),
]


def make_preprocess():
    """The reference's OWN build_transforms(cfg, is_train=True) chain (PIL resize, flip, ToTensor, Normalize) and
    BatchCollator on synthetic decoded images + BoxLists, Python `random` seeded per case."""
    import random
    from PIL import Image
    rh.install()
    # `maskrcnn_benchmark/data/__init__.py` pulls in the dataset catalogue (pycocotools, torch._six: both absent
    # here); register the package by path so that only the two modules under test are executed.
    import types
    pkg = types.ModuleType("maskrcnn_benchmark.data")
    pkg.__path__ = [os.path.join(rh.REF, "maskrcnn_benchmark", "data")]
    sys.modules.setdefault("maskrcnn_benchmark.data", pkg)
    from maskrcnn_benchmark.data.transforms import build_transforms
    from maskrcnn_benchmark.data.collate_batch import BatchCollator
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    cases = {}
    for name, opts, sizes, seed in PREPROCESS_CASES:
        cfg = rh.reference_cfg("da_faster_rcnn/e2e_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml", opts)
        g = torch.Generator().manual_seed(seed)
        raws, boxes = [], []
        for (h, w) in sizes:
            raws.append(torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8))
            x1 = torch.rand(5, generator=g) * (w - 8)
            y1 = torch.rand(5, generator=g) * (h - 8)
            boxes.append(torch.stack([x1, y1, x1 + 2 + torch.rand(5, generator=g) * 5,
                                      y1 + 2 + torch.rand(5, generator=g) * 5], 1))
        tf = build_transforms(cfg, is_train=True)
        random.seed(seed)
        samples = []
        for i, (raw, bx) in enumerate(zip(raws, boxes)):
            t = BoxList(bx.clone(), (raw.shape[1], raw.shape[0]), mode="xyxy")
            t.add_field("labels", torch.arange(1, 6))
            img, t = tf(Image.fromarray(raw.numpy(), mode="RGB"), t)
            samples.append((img, t, i))
        images, targets, ids = BatchCollator(cfg.DATALOADER.SIZE_DIVISIBILITY)(samples)
        cases[name] = dict(opts=opts, seed=seed, raw=raws, boxes=boxes, batch=images.tensors.clone(),
                           image_sizes=[tuple(int(v) for v in s) for s in images.image_sizes],
                           target_boxes=[t.bbox.clone() for t in targets], target_sizes=[t.size for t in targets])
    torch.save(cases, os.path.join(OUT, "preprocess_ref.pt"))
    return cases


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "boundary":
        assert rh.available(), "reference not mounted"
        cfgs, st = make_boundary()
        print(len(cfgs), "configs;", sorted(cfgs)[:3], "...;", len(next(iter(cfgs.values()))["effective"]), "keys each")
    elif len(sys.argv) > 1 and sys.argv[1] == "fpn":
        assert rh.available(), "reference not mounted"
        torch.set_num_threads(os.cpu_count())
        fx = make_fpn_eval()
        print("pyramid", [(p["shape"], round(p["absmean"], 4)) for p in fx["pyramid"]])
        print("proposals", [len(p["objectness"]) for p in fx["proposals"]], "levels", torch.bincount(fx["levels"]).tolist())
        for d in fx["detections"]:
            print(len(d["scores"]), "detections; labels", sorted(set(d["labels"].tolist()))[:12], "scores",
                  [round(float(v), 4) for v in d["scores"][:6]])
    elif len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        assert rh.available(), "reference not mounted"
        for k, c in make_preprocess().items():
            print(k, tuple(c["batch"].shape), c["image_sizes"])
    elif len(sys.argv) > 1 and sys.argv[1] == "eval":       # only the eval-mode fixture
        assert rh.available(), "reference not mounted"
        torch.set_num_threads(os.cpu_count())
        fx = make_eval()
        for d in fx["detections"]:
            print(len(d["scores"]), "detections; labels", sorted(set(d["labels"].tolist()))[:12], "scores",
                  [round(float(v), 4) for v in d["scores"][:6]])
    else:
        main()
        make_eval()
        make_fpn_eval()
        make_preprocess()
        make_boundary()
