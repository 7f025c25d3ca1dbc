#!/usr/bin/env python
"""bench.py — training throughput of the DA Faster R-CNN hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dense mixed|tcgen05x3|tcgen05|simt]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one training iteration of the reference's do_da_train loop (engine/trainer.py:196-242):
forward + backward of GeneralizedRCNN with the DA heads on one (source, target) pair of synthetic
1024x2048 images per GPU, gradient all-reduce (N > 1), SGD-momentum update.  Workload = BASELINE.json
configs[1] (R-50-C4, image-level DA); `--config 2|3` selects configs[2] / configs[3].  One JSON line (rank 0).

The benchmarked dense arm is `mixed`: forward products 3xTF32 on the tensor cores (fp32-grade: losses within 1e-4
of the CPU oracle WITH ITS OWN hard decisions — top-k, NMS, samplers — tests/test_gpu_model.py, and at the
benchmarked shape tests/test_gpu_fullsize.py), data- and weight-gradient products plain TF32 (gradient tolerance
1e-2 of the global norm, checked over three SGD iterations).  The all-TF32 arm and the all-3xTF32 arm are timed
beside it and reported as side keys.

  value     images/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e       images/s through the public API (FlatSGDTrainer.step) with the batch in pinned HOST memory:
            H2D copy of images+targets and a D2H read of the loss vector inside the timed region
  sustained the same resident-batch loop run for >= 5 s (clocks sampled during it)
  roofline  the dominant kernel of the step (RPN 3x3 1024->1024 conv forward, 3xTF32) timed alone with CUDA
            events; algorithmic FLOPs / time vs the measured dense bf16 peak
  parity_check  one full-size step of this workload against the CPU oracle (outside every timed region)
  torch_cudnn_baseline  the reference's graph through stock torch/cuDNN ops on this GPU, TF32 on / off
  cpu_baseline  the CPU oracle (port of the reference path, oracle/) timed on the host cores: one full step of the
            same 2 x 1024 x 2048 workload (rank 0, N = 1 only)
`--impl reference` times that CPU port alone (the reference itself cannot be installed: SURVEY §8c).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1024, 2048
YAML = os.path.join(ROOT, "configs", "da_faster_rcnn", "e2e_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml")
OPTS = ["MODEL.DA_HEADS.DA_INS_LOSS_WEIGHT", 0.0, "MODEL.DA_HEADS.DA_CST_LOSS_WEIGHT", 0.0]   # configs[1]: image-level DA only
WORKLOAD = "da_faster_rcnn R-50-C4 image-level DA (BASELINE configs[1]); 1 source + 1 target 1024x2048 per GPU, 20 GT boxes/img"
TFLOP_PER_IMAGE = 2.27


def load_cfg():
    from dadetect_b200.config import get_cfg_defaults
    cfg = get_cfg_defaults()
    cfg.merge_from_file(YAML)
    cfg.merge_from_list(OPTS)
    return cfg


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (the B200_PROFILING.md clocks line), read
    in-process through NVML: forking `nvidia-smi` five times a second from a process that holds a CUDA context
    stalls the launching thread and roughly doubles the measured step time."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def run(self):
        n = self.nvml
        while not self.stop_flag and n is not None:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, mx, reasons))
            except Exception:
                pass
            time.sleep(0.25)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = [k for k, b in bits.items() if any(s[2] & b for s in self.samples)]
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.samples[0][1],
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU port arm
def cpu_port_step_seconds(cfg, h, w, steps=1, warmup=0):
    """Wall-clock seconds per training step (fwd+bwd+SGD) of the oracle port on the host cores."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import da_frcnn_ref as orc
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    torch.set_num_threads(os.cpu_count())
    sd = make_state_dict(orc.param_shapes(cfg))
    P = {k: v.clone().requires_grad_(orc.is_trainable(k)) for k, v in sd.items()}
    train = [p for p in P.values() if p.requires_grad]
    opt = torch.optim.SGD(train, lr=cfg.SOLVER.BASE_LR, momentum=cfg.SOLVER.MOMENTUM, weight_decay=cfg.SOLVER.WEIGHT_DECAY)
    times = []
    for s in range(warmup + steps):
        images, targets = make_batch(2, h, w, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, seed=1029 + s)
        t0 = time.perf_counter()
        losses = orc.forward_train(P, cfg, images, targets, nms_strict=True)
        opt.zero_grad()
        sum(losses.values()).backward()
        opt.step()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def cpu_baseline(cfg, budget_s=30.0):
    """One full training step (fwd + bwd + SGD) of the same 2 x 1024 x 2048 workload on all host threads — measured,
    not extrapolated (about 25 s on the 16-core box)."""
    t = cpu_port_step_seconds(cfg, H, W)
    return dict(value=2.0 / t, unit="images/s", cores=os.cpu_count(), kind="port", seconds_per_step=t,
                sample="1 full training step (fwd+bwd+SGD) of the same 2x1024x2048 workload, torch CPU fp32, all host threads")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg()
    base = cpu_baseline(cfg, budget_s=40.0)
    k = max(1, args.steps)
    line = {
        "impl": "reference", "metric": "DA-FRCNN R-50-C4 train images/sec", "value": base["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": 2000.0 / base["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "baseline_config_index": 1, "tflop_per_image": TFLOP_PER_IMAGE,
                   "note": "CPU oracle port of the reference path (the reference cannot be installed, SURVEY §8c); one "
                           "full-size step (fwd+bwd+SGD, all host threads) stands for every requested step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
DENSE_DTYPE = {"simt": "f32", "tcgen05": "tf32", "tcgen05x3": "tf32x3",
               "mixed": "tf32x3-forward/tf32-backward (fp32 storage, fp32 accumulate)"}
# BASELINE.json configs index -> (yaml, overrides, images per GPU, TFLOP per image, label)
TRIPLET_YAML = os.path.join(ROOT, "configs", "da_faster_rcnn",
                            "e2e_triplet_da_faster_rcnn_R_50_C4_cityscapes_to_foggy_cityscapes.yaml")
BENCH_CONFIGS = {
    1: (YAML, OPTS, 2, 2.27, WORKLOAD),
    2: (YAML, [], 2, 2.275, "da_faster_rcnn R-50-C4 image + instance + consistency DA (BASELINE configs[2]); 1 source + "
                            "1 target 1024x2048 per GPU (IMS_PER_BATCH = 2 x GPUs), 20 GT boxes/img"),
    3: (TRIPLET_YAML, ["MODEL.DA_HEADS.ALIGNMENT", True, "MODEL.DA_HEADS.DA_TRIPLET_INS_WEIGHT", 1.0], 3, 3.0,
        "e2e_triplet_da_faster_rcnn R-50-C4 aligned triplet + AdvGRL (BASELINE configs[3]); source + target + auxiliary "
        "1024x2048 per GPU, 20 GT boxes/img"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dense", default=os.environ.get("DADETECT_DENSE", "mixed"),
                    choices=["mixed", "simt", "tcgen05", "tcgen05x3"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3], help="BASELINE.json configs index")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-arms", action="store_true", help="skip the all-TF32 and all-3xTF32 side measurements")
    ap.add_argument("--no-check", action="store_true", help="skip the full-size parity check against the CPU oracle")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the stock torch/cuDNN same-GPU baseline")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 5 s sustained leg")
    ap.add_argument("--vary-gt", action="store_true", help="a different number of GT boxes per image every step")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly (no whole-step CUDA graph)")
    args = ap.parse_args()
    if os.environ.get("DD_BENCH_WATCHDOG"):          # developer aid: dump all stacks and exit if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DD_BENCH_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from dadetect_b200 import _lib, ops
    from dadetect_b200.config import get_cfg_defaults
    from dadetect_b200.engine import DevicePrefetcher, FlatSGDTrainer
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if not ops.tcgen05_available() and args.dense != "simt":
        raise SystemExit("bench.py: the tcgen05 arm is not built into libdadetect_b200.so")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:                                   # run (and pin host memory) on the CPUs next to this GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        if os.environ.get("DD_OVERLAP_EXCHANGE", "0") == "1" and int(os.environ.get("DD_EXCHANGE_CTAS", "0")) > 0:
            # the overlapped exchange leaves DD_EXCHANGE_CTAS SMs to NCCL: keep its kernels within them
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ["DD_EXCHANGE_CTAS"])
        dist.init_process_group("nccl", init_method="env://")
    assert world == args.gpus, "launch with torchrun --nproc-per-node {} for --gpus {}".format(args.gpus, args.gpus)
    warmup = max(3, args.warmup)
    impl_of = {"simt": ops.IMPL_SIMT, "tcgen05": ops.IMPL_TCGEN05, "tcgen05x3": ops.IMPL_TCGEN05_X3,
               "mixed": ops.IMPL_TCGEN05_MIXED}
    dense = args.dense

    yaml_path, opts, n_img, tflop_per_image, workload = BENCH_CONFIGS[args.config]
    cfg = get_cfg_defaults()
    cfg.merge_from_file(yaml_path)
    cfg.merge_from_list(list(opts))
    shapes_holder = {}

    def make_trainer(arm):
        """A fresh model + trainer on dense arm `arm` with the seeded synthetic weights."""
        ops.set_default_impl(impl_of[arm])
        model = build_detection_model(cfg).to(dev)
        if not shapes_holder:
            shapes_holder.update({k: tuple(v.shape) for k, v in model.state_dict().items()})
        model.load_state_dict(make_state_dict(shapes_holder), strict=False)
        model.train()
        tr = FlatSGDTrainer(model, cfg, world_size=world)
        if not args.no_graphs:
            tr.enable_step_graph(True)       # zero_grad + fwd + bwd + all-reduce + SGD as ONE CUDA graph
        return tr

    def host_batch(step):
        m = 20 if not args.vary_gt else 4 + (7 * step) % 29       # 4 .. 32 GT boxes per image, different every step
        images, targets = make_batch(n_img, H, W, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, boxes_per_image=m,
                                     seed=1029 + 1000 * rank + step)
        return images.pin_memory(), [dict(boxes=t["boxes"].pin_memory(), labels=t["labels"].pin_memory(),
                                          is_source=t["is_source"]) for t in targets]

    def to_device(images, targets):
        img = images.to(dev, non_blocking=True)
        tg = []
        for t in targets:
            b = BoxList(t["boxes"].to(dev, non_blocking=True), (W, H), mode="xyxy")
            b.add_field("labels", t["labels"].to(dev, non_blocking=True))
            b.add_field("is_source", torch.full((len(t["labels"]),), bool(t["is_source"]), dtype=torch.bool, device=dev))
            b._is_source_image = bool(t["is_source"])     # known on the host: no device read to find the domain
            tg.append(b)
        return img, tg

    n_host = 4
    host = [host_batch(s) for s in range(n_host)]
    resident = [to_device(*host[s]) for s in range(n_host)]
    h2d_bytes = host[0][0].numel() * 4 + sum(t["boxes"].numel() * 4 + t["labels"].numel() * 8 for t in host[0][1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_step, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for s in range(steps):
            run_step(s)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    state = {"trainer": make_trainer(dense), "n": 0}

    def step_resident(s):
        img, tg = resident[s % n_host]
        state["trainer"].step(img, tg)

    loss_host = torch.empty(16, dtype=torch.float32).pin_memory()
    prefetch = DevicePrefetcher(dev, (W, H))
    e2e_tag = [0]

    def step_e2e(_s):
        # public API end to end: every step moves its own batch from pinned host memory (on the copy stream, one
        # batch ahead of the compute stream), runs trainer.step and reads the loss vector back to the host
        s = e2e_tag[0]
        e2e_tag[0] += 1
        if s == 0:
            prefetch.put(0, *host[0])
        img, tg = prefetch.get(s)
        prefetch.put(s + 1, *host[(s + 1) % n_host])
        ld = state["trainer"].step(img, tg)
        prefetch.release(s)
        vec = torch.stack([v.detach() for v in ld.values()])
        loss_host[: vec.numel()].copy_(vec, non_blocking=False)       # D2H read of the step's result
        state["n"] = vec.numel()

    # ---- dominant kernel alone, timed FIRST (a cool GPU at its burst clocks: the peak it is held against is the
    # burst figure of MEASURED_PEAKS.json): RPN 3x3 conv 1024->1024 forward on [2,64,128,1024], the largest GEMM of the
    # step, on the arm the step runs it on (3xTF32 under `mixed`), and the TF32 kernel the backward uses beside it
    model = state["trainer"].model
    feat = torch.randn(n_img, H // 16, W // 16, 1024, device=dev)
    wt = ops.weight_ohwi(model.rpn.head.conv.weight.detach())
    bias = model.rpn.head.conv.bias.detach()
    k_flops = 2.0 * n_img * (H // 16) * (W // 16) * 9 * 1024 * 1024

    def time_kernel(impl, reps=10):
        for _ in range(3):
            ops.conv2d_forward_raw(feat, wt, None, bias, None, 3, 3, 1, 1, True, impl=impl)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            ops.conv2d_forward_raw(feat, wt, None, bias, None, 3, 3, 1, 1, True, impl=impl)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps

    fwd_arm = ops.fwd_impl(impl_of[dense])
    k_ms = time_kernel(fwd_arm)
    k2_ms = time_kernel(ops.IMPL_TCGEN05) if fwd_arm == ops.IMPL_TCGEN05_X3 else None
    del feat

    # untimed warm-up: at least two passes over the distinct batches so that the caching allocator has seen
    # every tensor size before the timed region
    for s in range(max(warmup, 2 * n_host)):
        step_resident(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + state["trainer"].graph_launches
    ms_step = timed(step_resident, args.steps)
    launches = _lib.launch_count() + state["trainer"].graph_launches - launches0
    n_graphs = len(state["trainer"].step_graphs) if state["trainer"].step_graphs is not None else 0
    for s in range(max(warmup, 3)):      # untimed: the prefetcher's device slots and copy stream come into being here
        step_e2e(s)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    images_per_step = n_img * world
    pk = peaks()

    line = {
        "metric": "DA-FRCNN R-50-C4 train images/sec", "value": images_per_step * 1000.0 / ms_step, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DENSE_DTYPE[dense], "data": "synthetic",
        "config": {"workload": workload, "baseline_config_index": args.config, "dense_impl": dense,
                   "whole_step_cuda_graph": not args.no_graphs, "step_graphs_captured": n_graphs,
                   "gt_boxes_per_image": "4..32, different every step" if args.vary_gt else 20,
                   "parallelism": "dp{}".format(world),
                   "l2": "per-step working set (>4 GB of activations) far exceeds the 126 MB L2; no flush needed",
                   "tflop_per_image": tflop_per_image,
                   "early_backward": bool(getattr(state["trainer"], "early_backward", False)),
                   "launches_per_step": launches / max(args.steps, 1),
                   "parity": "losses within 1e-4 of the CPU oracle with the arm's own hard decisions (tests/test_gpu_model.py "
                             "[mixed], tests/test_gpu_fullsize.py at 1024x2048); gradients TF32-grade (1e-2 of the global norm)"},
        "clocks": sampler.summary(),
        "e2e": {"value": images_per_step * 1000.0 / ms_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4 * state["n"], "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "step_frac_of_flop_roofline": (n_img * 1000.0 / ms_step) * tflop_per_image / pk["bf16_sustained"],
    }

    # ---- >= 5 s sustained leg (clocks sampled over it): the burst figure above divided by a sustained peak would
    # flatter the step; this is the number to hold against bf16_tflops_sustained
    if not args.no_sustained:
        sus = ClockSampler(local)
        if rank == 0:
            sus.start()
        n_sus = max(args.steps, int(5200.0 / ms_step) + 1)
        ms_sus = timed(step_resident, n_sus)
        sus.stop_flag = True
        line["sustained"] = {"steps": n_sus, "seconds": n_sus * ms_sus / 1000.0, "ms_per_step": ms_sus,
                             "value": images_per_step * 1000.0 / ms_sus, "unit": "images/s", "clocks": sus.summary(),
                             "step_frac_of_flop_roofline": (n_img * 1000.0 / ms_sus) * tflop_per_image / pk["bf16_sustained"]}

    # ---- roofline of the dominant kernel (timed alone at the start of the run, see above)
    def traffic_of(tag):
        tpath = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")     # dram bytes/launch from ncu --set full
        if os.path.exists(tpath):
            return json.load(open(tpath)).get(tag)
        return None

    x3 = fwd_arm == ops.IMPL_TCGEN05_X3
    achieved_tf = k_flops / (k_ms * 1e-3) / 1e12
    line["roofline"] = {
        "bound": "tensor", "achieved": achieved_tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
        "frac": achieved_tf / pk["bf16_burst"], "traffic": traffic_of("rpn3x3_fwd_x3" if x3 else "rpn3x3_fwd_tf32"),
        "traffic_source": "profiles/r02_roofline_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
        "kernel": ("conv_tc_kernel<X3> (RPN 3x3 1024->1024 fwd, M={} N=1024 K=9216; ALGORITHMIC flops: each product costs "
                   "three TF32 MMAs, so the ceiling of this kernel is 1/6 of the bf16 peak)" if x3 else
                   "conv_tc_kernel<256,0> (RPN 3x3 1024->1024 fwd, M={} N=1024 K=9216, TF32: ceiling = 0.5 of the bf16 "
                   "peak)").format(n_img * (H // 16) * (W // 16)),
        "ms": k_ms, "ceiling_frac_of_peak": (1.0 / 6.0) if x3 else 0.5,
        "peak_source": pk["source"] + ", dense bf16 burst"}
    if k2_ms is not None:
        a2 = k_flops / (k2_ms * 1e-3) / 1e12
        line["roofline_tf32_kernel"] = {"bound": "tensor", "achieved": a2, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                                        "frac": a2 / pk["bf16_burst"], "traffic": traffic_of("rpn3x3_fwd_tf32"), "ms": k2_ms,
                                        "kernel": "conv_tc_kernel<256,0> the same GEMM in plain TF32 (the arm the backward "
                                                  "products run on); ceiling = 0.5 of the bf16 peak"}

    # ---- side arms on the same workload (not the headline): all-TF32 and all-3xTF32
    import gc

    def drop_trainer():
        tr = state.pop("trainer", None)
        if tr is not None:
            tr.step_graphs = None
        del tr
        gc.collect()
        torch.cuda.empty_cache()

    if not args.no_side_arms and world == 1:
        for arm, key, note in (("tcgen05", "tf32_arm", "all products plain TF32 (what stock PyTorch/cuDNN computes by default); "
                                                      "losses within 2e-3 only, hard decisions not pinned: NOT parity-grade"),
                               ("tcgen05x3", "fp32_grade_arm", "all products 3xTF32: losses AND gradients at the fp32 tolerances")):
            if arm == dense:
                continue
            drop_trainer()
            state["trainer"] = make_trainer(arm)
            for s_ in range(2 * n_host):
                step_resident(s_)
            ms_a = timed(step_resident, max(4, args.steps // 2))
            line[key] = {"dense_impl": arm, "dtype": DENSE_DTYPE[arm], "value": images_per_step * 1000.0 / ms_a,
                         "unit": "images/s", "ms_per_step": ms_a, "note": note}
    drop_trainer()
    prefetch = None
    ops.set_default_impl(impl_of[dense])

    if rank == 0 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        if not args.no_check:
            # one full-size step of THIS workload against the CPU oracle, outside every timed region
            import fullsize_parity
            rep = fullsize_parity.run(args.config, dense=dense, with_grads=False)
            bad = fullsize_parity.verdict(rep)
            line["parity_check"] = {"passed": not bad, "failures": bad, "shape": rep["shape"], "dense": dense,
                                    "method": "tests/fullsize_parity.py: (1) RPN-head arithmetic vs oracle, (2) oracle decision "
                                              "procedure on the product's own logits == product proposals, (3) downstream of the "
                                              "proposals vs oracle with replayed draws",
                                    "max_loss_rel": max(v[2] for v in rep["losses"].values()),
                                    "losses": {k: [round(v[0], 7), round(v[1], 7)] for k, v in rep["losses"].items()},
                                    "arithmetic": rep["arithmetic"], "decisions": rep["decisions"],
                                    "vs_oracle_proposals": rep["vs_oracle_proposals"],
                                    "losses_with_own_decisions": rep["losses_with_own_decisions"],
                                    "index_tier": {k: rep[k] for k in rep if k.endswith("_equal") or k == "roi_boxes_moved"},
                                    "oracle_seconds": rep["oracle_s"]}
        if not args.no_torch_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import torch_baseline
            try:
                line["torch_cudnn_baseline"] = torch_baseline.baseline(args.config, steps=3)
            except Exception as e:       # a baseline, never a reason to lose the bench line
                line["torch_cudnn_baseline"] = {"unavailable": repr(e)[:300]}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(load_cfg())
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in dependency order: captured graphs hold NCCL kernels, so they go before the communicator;
        # the destroy itself runs under a deadline (a wedged communicator must not hang the launcher).
        dist.barrier()
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=20.0)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
