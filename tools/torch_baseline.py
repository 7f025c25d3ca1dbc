"""The "reference on the same B200" baseline of SURVEY §8(d) / BASELINE.md §4: the reference's graph for one
training iteration (oracle/da_frcnn_ref.py, the line-by-line restatement of GeneralizedRCNN.forward, pinned to the real
reference) executed on the GPU through STOCK torch ops — cuDNN convolutions, cuBLAS linears, ATen elementwise /
top-k / sort / losses, torchvision's roi_align and nms standing in for the reference's csrc CUDA kernels (unbuildable:
THC is gone) — plus torch.optim.SGD with the reference's parameter groups (solver/build.py:7-20).
It is what a user gets by running the reference itself on this GPU (same op sequence, same host syncs), with TF32
allowed (torch's default for convolutions) or forbidden (the reference's fp32 arithmetic).  A stated baseline: none of
it is on the product path.

  python tools/torch_baseline.py [config-index 1|2|3] [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def step_ms(index=1, tf32=True, steps=5, warmup=2, height=1024, width=2048, device="cuda"):
    """Device-timed ms per training step (forward + backward + SGD) of the reference graph on `device`."""
    import torch
    import da_frcnn_ref as orc
    from fullsize_parity import load_cfg
    from dadetect_b200.utils.synthetic import make_batch, make_state_dict
    cfg, n = load_cfg(index)
    S = cfg.SOLVER
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.benchmark = True
    dev = torch.device(device)
    try:
        sd = make_state_dict(orc.param_shapes(cfg))   # (CPU generators: outside the device context)
        host_batches = [make_batch(n, height, width, num_classes=cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, seed=1029 + s)
                        for s in range(2)]
        with torch.device(dev):                       # the oracle's factory calls (arange, zeros, ...) follow
            P = {k: v.to(dev).requires_grad_(orc.is_trainable(k)) for k, v in sd.items()}
            groups = []
            for k, p in P.items():
                if p.requires_grad:
                    bias = "bias" in k
                    groups.append({"params": [p], "lr": S.BASE_LR * (S.BIAS_LR_FACTOR if bias else 1.0),
                                   "weight_decay": S.WEIGHT_DECAY_BIAS if bias else S.WEIGHT_DECAY})
            opt = torch.optim.SGD(groups, S.BASE_LR, momentum=S.MOMENTUM)
            batches = []
            for images, targets in host_batches:
                batches.append((images.to(dev), [dict(boxes=t["boxes"].to(dev), labels=t["labels"].to(dev),
                                                      is_source=t["is_source"]) for t in targets]))
            state = orc.TripletState()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for s in range(warmup + steps):
                if s == warmup:
                    torch.cuda.synchronize()
                    ev0.record()
                images, targets = batches[s % 2]
                losses = orc.forward_train(P, cfg, images, targets, triplet_state=state, nms_strict=True)
                opt.zero_grad()
                sum(losses.values()).backward()
                opt.step()
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) / steps, n
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


def baseline(index=1, steps=5):
    out = {"what": "reference graph (oracle restatement) through stock torch/cuDNN/torchvision ops + torch.optim.SGD on "
                   "this GPU, same 1024x2048 batch; not on the product path", "unit": "images/s"}
    for name, tf32 in (("tf32", True), ("fp32", False)):
        t0 = time.perf_counter()
        ms, n = step_ms(index, tf32, steps=steps)
        out[name] = {"ms_per_step": ms, "value": n * 1000.0 / ms, "wall_s": time.perf_counter() - t0}
    return out


if __name__ == "__main__":
    import json
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    print(json.dumps(baseline(idx, int(sys.argv[2]) if len(sys.argv) > 2 else 5)))
