"""Developer tool (GPU): run warm-up steps, then exactly ONE eager training step (config 2, 2 x 1024 x 2048)
between cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists:

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/one_step.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dadetect_b200 import ops
from dadetect_b200.engine import FlatSGDTrainer
from dadetect_b200.modeling import build_detection_model
from dadetect_b200.structures import BoxList
from dadetect_b200.utils.synthetic import make_batch, make_state_dict

dev = torch.device("cuda")
ops.set_default_impl(ops.IMPL_TCGEN05)
cfg = bench.load_cfg()
model = build_detection_model(cfg).to(dev)
model.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}), strict=False)
model.train()
trainer = FlatSGDTrainer(model, cfg, world_size=1)
images, targets = make_batch(2, bench.H, bench.W, num_classes=9, seed=1029)
tg = []
for t in targets:
    b = BoxList(t["boxes"].to(dev), (bench.W, bench.H), mode="xyxy")
    b.add_field("labels", t["labels"].to(dev))
    b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
    tg.append(b)
img = images.to(dev)
for _ in range(3):
    trainer.step(img, tg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
trainer.step(img, tg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
