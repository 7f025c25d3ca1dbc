"""Developer tool (GPU): device timeline of ONE replay of the whole-step CUDA graph — every kernel with its stream,
start and duration (CUPTI activity records through torch.profiler; no ncu serialisation, so overlap between the
streams of the step is visible).  Prints the timeline, the per-kernel totals and the time during which fewer than
`--busy` kernels' worth of the GPU was occupied.

  python tools/step_timeline.py [--config 1] [--dense mixed] [--out gpurun_out/timeline.txt]
"""
import argparse
import collections
import os
import re
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


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("(anonymous namespace)::", "").replace("at::native::", "at::")
    name = re.sub(r"\(.*", "", name)
    return name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--dense", default="mixed")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda")
    impl = {"simt": ops.IMPL_SIMT, "tcgen05": ops.IMPL_TCGEN05, "tcgen05x3": ops.IMPL_TCGEN05_X3,
            "mixed": ops.IMPL_TCGEN05_MIXED}[args.dense]
    ops.set_default_impl(impl)
    yaml_path, opts, n_img, _, _ = bench.BENCH_CONFIGS[args.config]
    cfg = get_cfg_defaults()
    cfg.merge_from_file(yaml_path)
    cfg.merge_from_list(list(opts))
    model = build_detection_model(cfg).to(dev)
    model.load_state_dict(make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}), strict=False)
    model.train()
    trainer = FlatSGDTrainer(model, cfg, world_size=1)
    trainer.enable_step_graph(True)
    images, targets = make_batch(n_img, bench.H, bench.W, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES,
                                 boxes_per_image=20, seed=1029)
    tg = []
    for t in targets:
        b = BoxList(t["boxes"].to(dev), (bench.W, bench.H), mode="xyxy")
        b.add_field("labels", t["labels"].to(dev))
        b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
        b._is_source_image = bool(t["is_source"])
        tg.append(b)
    img = images.to(dev)
    for _ in range(5):
        trainer.step(img, tg)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        trainer.step(img, tg)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    rows = []
    for e in evs:
        rng = e.time_range
        rows.append((rng.start, rng.end, getattr(e, "device_resource_id", -1), short(e.name)))
    rows.sort()
    if not rows:
        raise SystemExit("no device activity records (CUPTI unavailable?)")
    t0 = rows[0][0]
    streams = {s: i for i, s in enumerate(sorted({r[2] for r in rows}))}
    lines = ["# one replay of the whole-step graph: config %d, dense arm %s; %d device activities, span %.1f us"
             % (args.config, args.dense, len(rows), rows[-1][1] - t0),
             "# start_us   dur_us  stream  kernel"]
    for a, b, s, n in rows:
        lines.append("%9.1f %8.1f  s%-2d  %s" % (a - t0, b - a, streams[s], n))
    # occupancy profile: time during which k activities overlap
    pts = []
    for a, b, _, _ in rows:
        pts.append((a, 1))
        pts.append((b, -1))
    pts.sort()
    depth, last, hist = 0, pts[0][0], collections.defaultdict(float)
    for t, d in pts:
        hist[depth] += t - last
        last = t
        depth += d
    lines.append("# time with k concurrent activities: " + ", ".join("k=%d: %.1f us" % (k, v) for k, v in sorted(hist.items())))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for a, b, _, n in rows:
        agg[n][0] += 1
        agg[n][1] += b - a
    lines.append("# per kernel (sum of durations; overlapping kernels both count)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append("#   %-70s n=%4d %9.1f us" % (n, c, t))
    text = "\n".join(lines)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
        print("\n".join(lines[-50:]))
    else:
        print(text)


if __name__ == "__main__":
    main()
