"""Developer tool (GPU): eval-mode throughput of the R-101-FPN detector (BASELINE configs[4]) at 800x1344
(the padded size of a 1333x800 COCO-style image), batch of 2, TF32 arm.

  python tools/fpn_bench.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dadetect_b200 import _lib, ops
from dadetect_b200.config import get_cfg_defaults
from dadetect_b200.modeling import build_detection_model
from dadetect_b200.utils.synthetic import make_batch, make_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = get_cfg_defaults()
cfg.merge_from_file(os.path.join(ROOT, "configs", "e2e_faster_rcnn_R_101_FPN_1x.yaml"))
dev = torch.device("cuda")
ops.set_default_impl(ops.IMPL_TCGEN05)
model = build_detection_model(cfg).to(dev)
sd = make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
sd["roi_heads.box.predictor.cls_score.weight"] *= 12.0
model.load_state_dict(sd, strict=False)
model.eval()
H, W = 800, 1344
images, _ = make_batch(2, H, W, num_classes=81, boxes_per_image=1, seed=3)
x = images.to(dev)
with torch.no_grad():
    for _ in range(3):
        out = model(x)
    torch.cuda.synchronize()
    before = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        out = model(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("R-101-FPN eval, 2 x %dx%d, TF32 arm: %.2f ms/batch, %.1f images/s, %d of our kernels per batch, detections %s" % (
    H, W, ms, 2000.0 / ms, (_lib.launch_count() - before) // steps, [len(o) for o in out]))
