"""Developer tool (GPU): training throughput of the OTHER BASELINE.json configurations at full size (1024x2048),
batch resident in HBM, TF32 arm — the driver-facing bench.py measures configs[1].

  python tools/config_bench.py 3     # image + instance + consistency DA (whole-step graph)
  python tools/config_bench.py 4     # aligned triplet with AdvGRL, 3 images (host-driven path, eager)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dadetect_b200 import ops
from dadetect_b200.config import get_cfg_defaults
from dadetect_b200.engine import FlatSGDTrainer
from dadetect_b200.modeling import build_detection_model
from dadetect_b200.structures import BoxList
from dadetect_b200.utils.synthetic import make_batch, make_state_dict

which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ROOT = bench.ROOT
cfg = get_cfg_defaults()
if which == 3:
    cfg.merge_from_file(bench.YAML)
    n_img, label = 2, "configs[2]: image + instance + consistency DA, 1 source + 1 target"
else:
    cfg.merge_from_file(os.path.join(ROOT, "configs", "da_faster_rcnn",
                                     "e2e_triplet_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml"))
    cfg.merge_from_list(["MODEL.DA_HEADS.ALIGNMENT", True, "MODEL.DA_HEADS.DA_TRIPLET_INS_WEIGHT", 1.0])
    n_img, label = 3, "configs[3]: aligned triplet (source / target / auxiliary), AdvGRL"
dev = torch.device("cuda")
ops.set_default_impl(ops.IMPL_TCGEN05)
model = build_detection_model(cfg).to(dev)
model.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}), strict=False)
model.train()
trainer = FlatSGDTrainer(model, cfg, world_size=1)
trainer.enable_step_graph(True)        # falls back to eager launches in the triplet modes
batches = []
for s in range(4):
    images, targets = make_batch(n_img, bench.H, bench.W, num_classes=9, seed=1029 + s)
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].to(dev), (bench.W, bench.H), mode="xyxy")
        b.add_field("labels", t["labels"].to(dev))
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
        b._is_source_image = bool(t["is_source"])
        tg.append(b)
    batches.append((images.to(dev), tg))
for s in range(8):
    ld = trainer.step(*batches[s % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 12
e0.record()
for s in range(steps):
    ld = trainer.step(*batches[s % 4])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("%s: %.2f ms/step, %.1f images/s (TF32 arm, 1x B200, graph=%s) losses=%s" % (
    label, ms, n_img * 1000.0 / ms, bool(trainer.step_graphs), {k: round(float(v), 4) for k, v in ld.items()}))
