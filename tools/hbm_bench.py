"""Developer probe (GPU): the HBM- and latency-bound kernels of one training step at the benchmarked shapes
(2 x 1024 x 2048, 512 ROIs), each timed alone with CUDA events: algorithmic bytes / time against the measured HBM
copy peak.  `--once` launches every kernel exactly once (for `ncu --set full -k regex:...` captures).

  python tools/hbm_bench.py [--once]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dadetect_b200 import ops as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def main():
    once = "--once" in sys.argv
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    peak = 6544.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    feat = torch.randn(2, 64, 128, 1024, device=dev, generator=g)
    K = 512
    x1 = torch.rand(K, device=dev, generator=g) * 1700
    y1 = torch.rand(K, device=dev, generator=g) * 800
    w = 32 + torch.rand(K, device=dev, generator=g) * 300
    h = 32 + torch.rand(K, device=dev, generator=g) * 200
    rois = torch.stack([(torch.arange(K, device=dev) % 2).float(), x1, y1, (x1 + w).clamp(max=2047), (y1 + h).clamp(max=1023)], 1)
    roi_out = o.roi_align(feat, rois, 1.0 / 16, 14, 0, 2)
    g_roi = torch.randn_like(roi_out)
    res5 = torch.randn(K, 7, 7, 2048, device=dev, generator=g)
    pooled = torch.empty(K, 2048, device=dev)
    g_pool = torch.randn(K, 2048, device=dev, generator=g)
    g_res5 = torch.empty_like(res5)
    stem_out = torch.randn(2, 512, 1024, 64, device=dev, generator=g)
    img = torch.randn(2, 3, 1024, 2048, device=dev, generator=g)
    w7 = torch.randn(64, 3, 7, 7, device=dev, generator=g) * 0.05
    sc, bi = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
    logits = torch.randn(2, 64, 128, 15, device=dev, generator=g) * 2
    deltas = torch.randn(2, 64, 128, 60, device=dev, generator=g) * 0.2
    import numpy as np
    from dadetect_b200.modeling.rpn import generate_cell_anchors
    cell = generate_cell_anchors(16, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0)).to(dev)
    anchors, _ = o.anchor_grid(cell, 64, 128, 16, 2048, 1024, 0)
    boxes, scores, _, valid = o.rpn_topk_decode(logits, deltas, anchors, 12000, 2048, 1024, 0.0)
    labels = (torch.rand(1, 122880, device=dev, generator=g) < 0.01).to(torch.int32) - (torch.rand(1, 122880, device=dev, generator=g) < 0.3).to(torch.int32)
    keys = torch.rand(1, 122880, device=dev, generator=g)
    gfeat = torch.randn_like(feat)
    gout = torch.empty_like(feat)
    wdev = torch.full((1,), -0.1, device=dev)
    from dadetect_b200 import _lib
    P = lambda t: None if t is None else __import__("ctypes").c_void_p(t.data_ptr())
    S = lambda: __import__("ctypes").c_void_p(torch.cuda.current_stream().cuda_stream)
    MB = 1e6
    cases = [
        ("roi_align_fwd_nhwc", lambda: o.roi_align(feat, rois, 1.0 / 16, 14, 0, 2),
         (roi_out.numel() * 4 + feat.numel() * 4) / MB, "512 ROIs x 49 even bins x 1024 ch written + the [2,64,128,1024] map read once"),
        ("roi_align_bwd_nhwc", lambda: _lib.call("dd_roi_align_backward", P(g_roi), P(rois), P(gout), 2, 64, 128, 1024, K, 1.0 / 16, 14, 14, 0, 2, S()),
         (g_roi.numel() * 4 + feat.numel() * 4) / MB, "bin gradients read + the map's gradient written once (zero fill excluded)"),
        ("avgpool_fwd_kernel", lambda: _lib.call("dd_avgpool_forward", P(res5), P(pooled), K, 49, 2048, S()),
         (res5.numel() * 4 + pooled.numel() * 4) / MB, "[512,7,7,2048] read, [512,2048] written"),
        ("avgpool_relu_bwd_kernel", lambda: _lib.call("dd_avgpool_relu_backward", P(g_pool), P(res5), P(g_res5), K, 49, 2048, S()),
         (2 * res5.numel() * 4 + pooled.numel() * 4) / MB, "res5 output (mask) + pooled gradient read, [512,7,7,2048] written"),
        ("maxpool3x3s2_kernel", lambda: o.maxpool3x3s2(stem_out), (stem_out.numel() * 4 * 1.25) / MB,
         "[2,512,1024,64] read, [2,256,512,64] written"),
        ("stem (pad + conv7x7 tcgen05)", lambda: o.stem_conv7x7s2(img, w7, sc, bi), (img.numel() * 4 + 2 * 512 * 1024 * 64 * 4) / MB,
         "image read, [2,512,1024,64] written (the padded NHWC4 copy is an internal 118 MB round trip)"),
        ("grl scale_kernel", lambda: _lib.call("dd_grl_backward_dev", P(gfeat), P(wdev), P(gout), gfeat.numel(), 0, S()),
         2 * gfeat.numel() * 4 / MB, "feature-map gradient read + written"),
        ("rpn_topk_decode_kernel", lambda: o.rpn_topk_decode(logits, deltas, anchors, 12000, 2048, 1024, 0.0),
         (logits.numel() * 4 + 2 * 12000 * 40) / MB, "latency-bound: 2 x 122 880 logits -> 12 000 sorted, decoded boxes"),
        ("nms (mask + scan, batch of 2)", lambda: o.nms_sorted_batched(boxes, valid, 0.7, 2000), 2 * 12000 * 16 / MB,
         "latency-bound: 2 x 12 000 sorted boxes -> <= 2000 survivors"),
        ("balanced_sample_kernel", lambda: o.balanced_sample(labels, None, keys, 256, 128), 2 * 122880 * 4 / MB,
         "latency-bound: 122 880 labels + keys -> 256 sampled indices"),
    ]
    o.set_default_impl(o.IMPL_TCGEN05)
    if once:
        for name, fn, mb, what in cases:
            fn()
        torch.cuda.synchronize()
        return
    print("# HBM- / latency-bound kernels at the benchmarked shapes, each timed alone (CUDA events, 20 launches, warm L2 for "
          "inputs that fit the 126 MB L2); peak = %.0f GB/s (MEASURED_PEAKS.json)" % peak)
    print("%-34s %9s %9s %9s %6s  %s" % ("kernel", "us", "MB", "GB/s", "frac", "algorithmic traffic"))
    for name, fn, mb, what in cases:
        us = timed(fn)
        gbs = mb / us * 1e3 / 1e3 * 1e3 / 1e3      # MB/us = TB/s -> GB/s
        gbs = mb * 1e6 / (us * 1e-6) / 1e9
        print("%-34s %9.1f %9.1f %9.0f %6.2f  %s" % (name, us, mb, gbs, gbs / peak, what))


if __name__ == "__main__":
    main()
