"""Developer tool (GPU): where does one training step spend its wall time?  Synchronising section timers
plus an un-instrumented step time and the CPU-only launch time (no sync until the end)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dadetect_b200 import ops
from dadetect_b200.engine import FlatSGDTrainer
from dadetect_b200.modeling import build_detection_model
from dadetect_b200.structures import BoxList
from dadetect_b200.utils import sections
from dadetect_b200.utils.synthetic import make_batch, make_state_dict

graphs = "--no-graphs" not in sys.argv
dev = torch.device("cuda")
ops.set_default_impl(ops.IMPL_TCGEN05)
cfg = bench.load_cfg()
model = build_detection_model(cfg).to(dev)
model.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}), strict=False)
model.train()
model.enable_cuda_graphs(graphs)
trainer = FlatSGDTrainer(model, cfg, world_size=1)
batches = []
for s in range(4):
    images, targets = make_batch(2, bench.H, bench.W, num_classes=9, seed=1029 + s)
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].to(dev), (bench.W, bench.H), mode="xyxy")
        b.add_field("labels", t["labels"].to(dev))
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
        tg.append(b)
    batches.append((images.to(dev), tg))
for s in range(8):
    trainer.step(*batches[s % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(8):
    trainer.step(*batches[s % 4])
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("graphs=%s  step %.2f ms   (host-side issue time %.2f ms)" % (graphs, t_all / 8 * 1e3, t_cpu / 8 * 1e3))
sections.enable(True)
for s in range(8):
    trainer.step(*batches[s % 4])
for k, v in sections.totals.items():
    print("  %-28s %8.2f ms" % (k, v / 8 * 1e3))
