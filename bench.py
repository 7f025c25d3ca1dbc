#!/usr/bin/env python
"""bench.py — training throughput of the DA Faster R-CNN hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dense simt|tcgen05]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one training iteration of the reference's do_da_train loop (engine/trainer.py:196-242):
forward + backward of GeneralizedRCNN with the DA heads on one (source, target) pair of synthetic
1024x2048 images per GPU, gradient all-reduce (N > 1), SGD-momentum update.  Workload = BASELINE.json
configs[1] (R-50-C4, image-level DA).  One JSON line on stdout (rank 0).

  value     images/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e       images/s through the public API (FlatSGDTrainer.step) with the batch in pinned HOST memory:
            H2D copy of images+targets and a D2H read of the loss vector inside the timed region
  roofline  the dominant kernel (the RPN 3x3 1024->1024 conv forward, the largest single GEMM of the step)
            timed alone with CUDA events; algorithmic FLOPs / time vs the measured dense bf16 peak
  cpu_baseline  the CPU oracle (port of the reference path, oracle/) timed on the host cores on a bounded
            sample (rank 0, N = 1 only)
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
    """Bounded sample: the largest of 1024x2048 / 512x1024 / 256x512 whose projected step time fits the budget
    (projection from a 256x512 probe, cost ~ pixels)."""
    t_probe = cpu_port_step_seconds(cfg, 256, 512)
    for (h, w) in ((1024, 2048), (512, 1024)):
        scale = (h * w) / (256.0 * 512.0)
        if t_probe * scale <= budget_s:
            t = cpu_port_step_seconds(cfg, h, w)
            if (h, w) == (H, W):
                return dict(value=2.0 / t, unit="images/s", cores=os.cpu_count(), kind="port",
                            sample="1 full training step (fwd+bwd+SGD) of the same 2x1024x2048 workload, torch CPU fp32, all host threads")
            return dict(value=2.0 / (t * (H * W) / (h * w)), unit="images/s", cores=os.cpu_count(), kind="port",
                        sample="1 training step at {}x{} (1/{} of the pixels), scaled linearly in pixels to 1024x2048".format(
                            h, w, (H * W) // (h * w)))
    return dict(value=2.0 / (t_probe * 16.0), unit="images/s", cores=os.cpu_count(), kind="port",
                sample="1 training step at 256x512 (1/16 of the pixels), scaled linearly in pixels to 1024x2048")


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
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference path (the reference cannot be installed, "
                   "SURVEY §8c); one bounded sample stands for every requested step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dense", default=os.environ.get("DADETECT_DENSE", "auto"), choices=["auto", "simt", "tcgen05", "tcgen05x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-x3", action="store_true", help="skip the extra measurement of the fp32-grade (3xTF32) arm")
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
    from dadetect_b200.engine import DevicePrefetcher, FlatSGDTrainer
    from dadetect_b200.modeling import build_detection_model
    from dadetect_b200.structures import BoxList
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:                                   # run (and pin host memory) on the CPUs next to this GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", init_method="env://")
    assert world == args.gpus, "launch with torchrun --nproc-per-node {} for --gpus {}".format(args.gpus, args.gpus)
    warmup = max(3, args.warmup)

    dense = args.dense
    if dense == "auto":
        dense = "tcgen05" if ops.tcgen05_available() else "simt"
    ops.set_default_impl({"tcgen05": ops.IMPL_TCGEN05, "tcgen05x3": ops.IMPL_TCGEN05_X3}.get(dense, ops.IMPL_SIMT))

    cfg = load_cfg()
    # param_shapes mirrors the reference state dict; kept inside the package-independent oracle file only for
    # tests, so here the model's own state dict provides the shapes.
    model = build_detection_model(cfg).to(dev)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(make_state_dict(shapes), strict=False)
    model.train()
    trainer = FlatSGDTrainer(model, cfg, world_size=world)
    if not args.no_graphs:
        trainer.enable_step_graph(True)       # zero_grad + fwd + bwd + all-reduce + SGD as ONE CUDA graph

    def host_batch(step):
        images, targets = make_batch(2, H, W, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES,
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

    last_losses = {}

    def step_resident(s):
        img, tg = resident[s % n_host]
        last_losses["d"] = trainer.step(img, tg)          # `trainer` is re-bound for the fp32-grade arm below

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
        ld = trainer.step(img, tg)
        prefetch.release(s)
        vec = torch.stack([v.detach() for v in ld.values()])
        loss_host[: vec.numel()].copy_(vec, non_blocking=False)       # D2H read of the step's result
        last_losses["n"] = vec.numel()

    # untimed warm-up: at least two passes over the distinct batches so that the caching allocator has seen
    # every tensor size (proposal counts vary per batch) before the timed region
    for s in range(max(warmup, 2 * n_host)):
        step_resident(s)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + trainer.graph_launches
    ms_step = timed(step_resident, args.steps)
    launches = _lib.launch_count() + trainer.graph_launches - launches0
    for s in range(max(warmup, 3)):      # untimed: the prefetcher's device slots and copy stream come into being here
        step_e2e(s)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True

    # ---- dominant kernel alone: RPN 3x3 conv 1024->1024 forward on [2,64,128,1024] (largest GEMM of the step)
    feat = torch.randn(2, H // 16, W // 16, 1024, device=dev)
    wt = ops.weight_ohwi(model.rpn.head.conv.weight.detach())
    bias = model.rpn.head.conv.bias.detach()
    if dense == "tcgen05x3":               # the roofline kernel is the TF32 kernel of the throughput arm
        ops.set_default_impl(ops.IMPL_TCGEN05)
    for _ in range(3):
        ops.conv2d_forward_raw(feat, wt, None, bias, None, 3, 3, 1, 1, True)
    torch.cuda.synchronize()
    reps = 10
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        ops.conv2d_forward_raw(feat, wt, None, bias, None, 3, 3, 1, 1, True)
    ev1.record()
    torch.cuda.synchronize()
    k_ms = ev0.elapsed_time(ev1) / reps
    k_flops = 2.0 * 2 * (H // 16) * (W // 16) * 9 * 1024 * 1024
    pk = peaks()
    achieved_tf = k_flops / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01b_roofline_traffic.json")      # dram bytes/launch from the ncu capture
    if os.path.exists(tpath) and dense == "tcgen05":
        traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                "frac": achieved_tf / pk["bf16_burst"], "traffic": traffic,
                "kernel": "conv_tc_kernel<256,0> (RPN 3x3 1024->1024 fwd, M=16384 N=1024 K=9216, TF32: ceiling = 0.5 of the bf16 peak) via " + dense,
                "peak_source": pk["source"] + ", dense bf16 burst",
                "step_frac_of_flop_roofline": (2.0 * 1000.0 / ms_step) * TFLOP_PER_IMAGE / pk["bf16_sustained"]}

    images_per_step = 2 * world
    line = {
        "metric": "DA-FRCNN R-50-C4 train images/sec", "value": images_per_step * 1000.0 / ms_step, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"simt": "f32", "tcgen05": "tf32", "tcgen05x3": "tf32x3"}[dense], "data": "synthetic",
        "config": {"workload": WORKLOAD, "dense_impl": dense, "whole_step_cuda_graph": not args.no_graphs, "parallelism": "dp{}".format(world),
                   "l2": "per-step working set (>4 GB of activations) far exceeds the 126 MB L2; no flush needed",
                   "tflop_per_image": TFLOP_PER_IMAGE},
        "clocks": sampler.summary(),
        "e2e": {"value": images_per_step * 1000.0 / ms_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4 * last_losses.get("n", 0), "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    # ---- the fp32-grade tensor-core arm (3xTF32: losses AND gradients within the fp32 tolerances of the oracle,
    # tests/test_gpu_model.py) on the same workload, reported beside the default TF32 arm
    if dense == "tcgen05" and not args.no_x3:
        import gc
        trainer.step_graphs = None
        trainer = model = None
        prefetch = None
        gc.collect()
        torch.cuda.empty_cache()
        ops.set_default_impl(ops.IMPL_TCGEN05_X3)
        model3 = build_detection_model(cfg).to(dev)
        model3.load_state_dict(make_state_dict(shapes), strict=False)
        model3.train()
        trainer = FlatSGDTrainer(model3, cfg, world_size=world)
        if not args.no_graphs:
            trainer.enable_step_graph(True)
        for s_ in range(4):
            step_resident(s_)
        ms3 = timed(step_resident, max(4, args.steps // 2))
        line["fp32_grade_arm"] = {"dense_impl": "tcgen05x3", "dtype": "tf32x3 (operands split hi/lo, fp32-grade)",
                                  "value": images_per_step * 1000.0 / ms3, "unit": "images/s", "ms_per_step": ms3,
                                  "parity": "losses and gradients within the fp32 tolerances (1e-4 on losses) of the CPU oracle"}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in dependency order: captured graphs hold NCCL kernels, so they go before the communicator;
        # the destroy itself runs under a deadline (a wedged communicator must not hang the launcher).
        import gc
        dist.barrier()
        trainer.step_graphs = None
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=20.0)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
