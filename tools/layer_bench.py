"""Developer probe (GPU): time every distinct dense layer of one config-2/3 training step (2 x 1024 x 2048,
512 ROIs) through the C-ABI conv entry points — forward, dgrad, wgrad — and print ms, TFLOP/s and the
algorithmic HBM GB/s of each, plus the per-step total weighted by how often the shape occurs.

  python tools/layer_bench.py [simt|tc|x3|cudnn|cudnn32] [filter-substring]

`cudnn` / `cudnn32` time the SAME shapes through stock torch (F.conv2d, torch.nn.grad.conv2d_input / conv2d_weight,
channels_last, cuDNN with TF32 allowed / forbidden): the "reference on the same B200" per-layer baseline of SURVEY
§8(d).  cuDNN's numbers are the bare convolution — the reference runs FrozenBN, the residual add and the ReLU as
separate elementwise passes on top, ours are fused into the epilogue and included in the time.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dadetect_b200 import ops as o

# name, N, H, W, Cin, Cout, k, stride, pad, occurrences per step, trainable (needs dgrad/wgrad), residual, relu
L = [
    ("res2.b0.conv1   1x1  64->64  ", 2, 256, 512, 64, 64, 1, 1, 0, 1, False, False, True),
    ("res2.bX.conv2   3x3  64->64  ", 2, 256, 512, 64, 64, 3, 1, 1, 3, False, False, True),
    ("res2.bX.conv3   1x1  64->256 ", 2, 256, 512, 64, 256, 1, 1, 0, 3, False, True, True),
    ("res2.b0.down    1x1  64->256 ", 2, 256, 512, 64, 256, 1, 1, 0, 1, False, False, False),
    ("res2.b12.conv1  1x1 256->64  ", 2, 256, 512, 256, 64, 1, 1, 0, 2, False, False, True),
    ("res3.b0.conv1   1x1s2 256->128", 2, 256, 512, 256, 128, 1, 2, 0, 1, True, False, True),
    ("res3.bX.conv2   3x3 128->128 ", 2, 128, 256, 128, 128, 3, 1, 1, 4, True, False, True),
    ("res3.bX.conv3   1x1 128->512 ", 2, 128, 256, 128, 512, 1, 1, 0, 4, True, True, True),
    ("res3.b0.down    1x1s2 256->512", 2, 256, 512, 256, 512, 1, 2, 0, 1, True, False, False),
    ("res3.b123.conv1 1x1 512->128 ", 2, 128, 256, 512, 128, 1, 1, 0, 3, True, False, True),
    ("res4.b0.conv1   1x1s2 512->256", 2, 128, 256, 512, 256, 1, 2, 0, 1, True, False, True),
    ("res4.bX.conv2   3x3 256->256 ", 2, 64, 128, 256, 256, 3, 1, 1, 6, True, False, True),
    ("res4.bX.conv3   1x1 256->1024", 2, 64, 128, 256, 1024, 1, 1, 0, 6, True, True, True),
    ("res4.b0.down    1x1s2 512->1024", 2, 128, 256, 512, 1024, 1, 2, 0, 1, True, False, False),
    ("res4.b1-5.conv1 1x1 1024->256", 2, 64, 128, 1024, 256, 1, 1, 0, 5, True, False, True),
    ("rpn.conv        3x3 1024->1024", 2, 64, 128, 1024, 1024, 3, 1, 1, 1, True, False, True),
    ("rpn.cls         1x1 1024->15 ", 2, 64, 128, 1024, 15, 1, 1, 0, 1, True, False, False),
    ("rpn.bbox        1x1 1024->60 ", 2, 64, 128, 1024, 60, 1, 1, 0, 1, True, False, False),
    ("res5.b0.conv1   1x1 1024->512", 512, 7, 7, 1024, 512, 1, 1, 0, 1, True, False, True),
    ("res5.bX.conv2   3x3 512->512 ", 512, 7, 7, 512, 512, 3, 1, 1, 3, True, False, True),
    ("res5.bX.conv3   1x1 512->2048", 512, 7, 7, 512, 2048, 1, 1, 0, 3, True, True, True),
    ("res5.b0.down    1x1 1024->2048", 512, 7, 7, 1024, 2048, 1, 1, 0, 1, True, False, False),
    ("res5.b12.conv1  1x1 2048->512", 512, 7, 7, 2048, 512, 1, 1, 0, 2, True, False, True),
    ("daimg.conv1     1x1 1024->512", 2, 64, 128, 1024, 512, 1, 1, 0, 1, True, False, True),
    ("dains.fc1       fc 2048->1024", 512, 1, 1, 2048, 1024, 1, 1, 0, 1, True, False, True),
    ("dains.fc2       fc 1024->1024", 512, 1, 1, 1024, 1024, 1, 1, 0, 1, True, False, True),
]


def timed(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    arm = sys.argv[1] if len(sys.argv) > 1 else "tc"
    impl = {"simt": o.IMPL_SIMT, "x3": o.IMPL_TCGEN05_X3}.get(arm, o.IMPL_TCGEN05)
    cudnn = arm.startswith("cudnn")
    if cudnn:
        torch.backends.cudnn.allow_tf32 = arm == "cudnn"
        torch.backends.cuda.matmul.allow_tf32 = arm == "cudnn"
        torch.backends.cudnn.benchmark = True
    print("# arm:", arm)
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    dev = "cuda"
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    totf = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    print("%-32s %5s | %8s %7s %7s | %8s %7s %7s | %8s %7s" % ("layer", "count", "fwd ms", "TF/s", "GB/s", "dgrad ms",
                                                                 "TF/s", "GB/s", "wgrad ms", "TF/s"))
    for (name, n, h, w, cin, cout, k, stride, pad, cnt, train, has_res, relu) in L:
        if filt and filt not in name:
            continue
        oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        x = torch.randn(n, h, w, cin, device=dev)
        wt = torch.randn(cout, k, k, cin, device=dev) / (cin * k * k) ** 0.5
        sc = torch.rand(cout, device=dev) + 0.5
        bi = torch.randn(cout, device=dev)
        res = torch.randn(n, oh, ow, cout, device=dev) if has_res else None
        gy = torch.randn(n, oh, ow, cout, device=dev)
        flop = 2.0 * n * oh * ow * cout * cin * k * k
        by_f = 4.0 * (n * oh * ow * cin * (1 if stride == 1 else 1) + n * oh * ow * cout * (2 if has_res else 1) + wt.numel())
        by_d = 4.0 * (n * oh * ow * cout + n * h * w * cin + wt.numel())
        if cudnn:
            xc = x.permute(0, 3, 1, 2)                       # logical NCHW over channels_last storage
            wc = wt.permute(0, 3, 1, 2)
            gc = gy.permute(0, 3, 1, 2)
            t_f = timed(lambda: torch.nn.functional.conv2d(xc, wc, None, stride, pad))
        else:
            t_f = timed(lambda: o.conv2d_forward_raw(x, wt, sc, bi, res, k, k, stride, pad, relu, impl=impl))
        line = "%-32s %5d | %8.3f %7.1f %7.0f" % (name, cnt, t_f, flop / t_f / 1e9, by_f / t_f / 1e6)
        tot["fwd"] += t_f * cnt
        totf["fwd"] += flop * cnt
        if train:
            if cudnn:
                t_d = timed(lambda: torch.nn.grad.conv2d_input(xc.shape, wc, gc, stride, pad))
                t_w = timed(lambda: torch.nn.grad.conv2d_weight(xc, wc.shape, gc, stride, pad))
            else:
                t_d = timed(lambda: o.conv2d_dgrad_raw(gy, wt, sc, (n, h, w, cin), k, k, stride, pad, impl=impl))
                t_w = timed(lambda: o.conv2d_wgrad_raw(gy, x, sc, cout, k, k, stride, pad, impl=impl))
            line += " | %8.3f %7.1f %7.0f | %8.3f %7.1f" % (t_d, flop / t_d / 1e9, by_d / t_d / 1e6, t_w, flop / t_w / 1e9)
            tot["dgrad"] += t_d * cnt
            tot["wgrad"] += t_w * cnt
            totf["dgrad"] += flop * cnt
            totf["wgrad"] += flop * cnt
        print(line, flush=True)
        del x, wt, res, gy
    for kk in ("fwd", "dgrad", "wgrad"):
        if tot[kk] > 0:
            print("TOTAL %-6s %8.3f ms/step  %7.1f TFLOP/s average" % (kk, tot[kk], totf[kk] / tot[kk] / 1e9))
    print("TOTAL dense %8.3f ms/step" % sum(tot.values()))


if __name__ == "__main__":
    main()
