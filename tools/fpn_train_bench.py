"""Developer tool (GPU): training throughput of the R-101-FPN detector with the per-level DA heads (BASELINE
configs[4] family) at 800x1344 (the padded size of a 1333x800 image), 1 source + 1 target image per step, on
FlatSGDTrainer's whole-step CUDA graph (sync-free FPN path) and, for comparison, on the host-driven control flow.

  python tools/fpn_train_bench.py [--dense mixed] [--no-da]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dadetect_b200 import ops
from dadetect_b200.config import get_cfg_defaults
from dadetect_b200.engine import FlatSGDTrainer
from dadetect_b200.modeling import build_detection_model
from dadetect_b200.structures import BoxList
from dadetect_b200.utils.synthetic import make_batch, make_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--dense", default="mixed")
ap.add_argument("--no-da", action="store_true")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
impl = {"simt": ops.IMPL_SIMT, "tcgen05": ops.IMPL_TCGEN05, "tcgen05x3": ops.IMPL_TCGEN05_X3,
        "mixed": ops.IMPL_TCGEN05_MIXED}[args.dense]
ops.set_default_impl(impl)
cfg = get_cfg_defaults()
cfg.merge_from_file(os.path.join(ROOT, "configs", "e2e_faster_rcnn_R_101_FPN_1x.yaml"))
opts = ["MODEL.ROI_BOX_HEAD.NUM_CLASSES", 9]
if not args.no_da:
    opts += ["MODEL.DOMAIN_ADAPTATION_ON", True, "MODEL.DA_HEADS.TRIPLET_USE", False]
cfg.merge_from_list(opts)
dev = torch.device("cuda")
H, W = 800, 1344
images, targets = make_batch(2, H, W, num_classes=9, boxes_per_image=20, seed=3)
if args.no_da:
    for t in targets:
        t["is_source"] = True
x = images.to(dev)


def boxlists():
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].to(dev), (W, H), mode="xyxy")
        b.add_field("labels", t["labels"].to(dev))
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
        b._is_source_image = bool(t["is_source"])
        tg.append(b)
    return tg


sd = None
for mode in ("host-driven, eager", "sync-free, whole-step CUDA graph"):
    model = build_detection_model(cfg).to(dev)
    if sd is None:
        sd = make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd, strict=False)
    model.train()
    tr = FlatSGDTrainer(model, cfg, world_size=1)
    if mode.startswith("sync-free"):
        tr.enable_step_graph(True)
    tg = boxlists()
    for _ in range(4):
        ld = tr.step(x, tg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ld = tr.step(x, tg)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print("R-101-FPN %s training, 2 x %dx%d, %s arm, %s: %.2f ms/step, %.1f images/s; losses %s" % (
        "DA" if not args.no_da else "plain", H, W, args.dense, mode, ms, 2000.0 / ms,
        {k: round(float(v), 4) for k, v in ld.items()}))
    del tr, model
    torch.cuda.empty_cache()
