"""Developer probe (GPU): tcgen05 conv arm vs torch CPU fp32 on a list of shapes; prints error statistics."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dadetect_b200 import ops as o

CASES = [
    (2, 16, 24, 64, 64, 1, 1, 0),
    (1, 16, 16, 32, 64, 1, 1, 0),
    (2, 16, 24, 64, 128, 1, 2, 0),
    (1, 15, 23, 32, 48, 3, 1, 1),
    (3, 7, 7, 128, 256, 3, 1, 1),
    (2, 8, 12, 256, 15, 1, 1, 0),
    (2, 8, 12, 64, 1, 1, 1, 0),
    (37, 1, 1, 2048, 1024, 1, 1, 0),
    (1, 64, 128, 1024, 1024, 3, 1, 1),
    (4, 7, 7, 512, 512, 3, 1, 1),
]
dev = "cuda"
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
only = [int(a) for a in sys.argv[1:]]
for idx, case in enumerate(CASES):
    if only and idx not in only:
        continue
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = torch.randn(cout, generator=g) * 0.1
    xr = x.clone().requires_grad_(True)
    y0 = F.conv2d(xr, wt, stride=stride, padding=pad) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    res = torch.randn(y0.shape, generator=g)
    want = F.relu(y0 + res)
    go = torch.randn(want.shape, generator=g)
    (gx_want,) = torch.autograd.grad(want, xr, go)
    xd, wd = nhwc(x).to(dev), wt.permute(0, 2, 3, 1).contiguous().to(dev)
    sd, bd, rd = scale.to(dev), bias.to(dev), nhwc(res).to(dev)
    torch.cuda.synchronize()
    got = o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=o.IMPL_TCGEN05)
    torch.cuda.synchronize()
    err = (got.permute(0, 3, 1, 2).cpu() - want)
    print("case", idx, case, "FWD max|err| %.3e  rms(err)/rms(want) %.3e" % (err.abs().max(), err.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()), flush=True)
    gpre = o.relu_backward_raw(nhwc(go).to(dev), o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=o.IMPL_SIMT))
    gx = o.conv2d_dgrad_raw(gpre, wd, sd, tuple(xd.shape), k, k, stride, pad, impl=o.IMPL_TCGEN05)
    torch.cuda.synchronize()
    err = (gx.permute(0, 3, 1, 2).cpu() - gx_want)
    print("          DGRAD max|err| %.3e  rms(err)/rms(want) %.3e" % (err.abs().max(), err.pow(2).mean().sqrt() / gx_want.pow(2).mean().sqrt()), flush=True)
    if n * h * w * cout * cin * k * k > 1e9:
        for impl, name in ((o.IMPL_SIMT, "simt"), (o.IMPL_TCGEN05, "tcgen05")):
            for _ in range(2):
                o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                o.conv2d_forward_raw(xd, wd, sd, bd, rd, k, k, stride, pad, True, impl=impl)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
            print("          %s fwd %.3f ms  %.1f TFLOP/s" % (name, ms, 2.0 * n * oh * ow * cout * cin * k * k / ms / 1e9), flush=True)
