"""Operator layer: torch tensors in, C-ABI kernel launches on the current CUDA stream, torch tensors out.

torch is plumbing here (device memory, streams, autograd bookkeeping); every arithmetic kernel is one
of ours from libdadetect_b200.so.  Activations are NHWC (``[N, H, W, C]`` contiguous); convolution
weights keep the reference's logical ``[Cout, Cin, KH, KW]`` shape (state-dict compatible) but are
stored channels_last, i.e. physically OHWI, which is the layout the kernels consume.

Each autograd Function fuses what the reference runs as several ATen kernels:
  ConvBnAct       Conv2d + FrozenBatchNorm2d (+ residual add) (+ ReLU)   resnet.py:294-314, batch_norm.py:19-24
  RoIAlign        layers/roi_align.py:11-44 -> _C.roi_align_forward/backward
  GradientScalar  layers/gradient_scalar_layer.py:4-24
  *Loss           rpn/loss.py:132-141, box_head/loss.py:200-219, da_heads/loss.py:140-222, consistency_loss.py
"""
import ctypes

import torch

from . import _lib

IMPL_SIMT = 0
IMPL_TCGEN05 = 1
IMPL_TCGEN05_X3 = 2          # 3xTF32 on the tensor cores: fp32-grade forward / data-gradient products
IMPL_TCGEN05_MIXED = 3       # host-level policy: forward products 3xTF32 (fp32-grade losses and hard decisions),
                             # data- and weight-gradient products plain TF32 (what cuDNN computes by default)
_default_impl = IMPL_SIMT
TC_IMPLS = (IMPL_TCGEN05, IMPL_TCGEN05_X3, IMPL_TCGEN05_MIXED)
NUM_SMS = 148                # B200


def fwd_impl(impl=None):
    """The kernel arm (a DD_IMPL_* value of the C ABI) that forward products run on under `impl`."""
    impl = _default_impl if impl is None else impl
    return IMPL_TCGEN05_X3 if impl == IMPL_TCGEN05_MIXED else impl


def bwd_impl(impl=None):
    """The kernel arm that data- and weight-gradient products run on under `impl`."""
    impl = _default_impl if impl is None else impl
    return IMPL_TCGEN05 if impl == IMPL_TCGEN05_MIXED else impl


def set_default_impl(impl):
    """Select the dense-tier arm used by conv2d/linear: IMPL_SIMT (fp32 FMA), IMPL_TCGEN05 (TF32 tensor cores),
    IMPL_TCGEN05_X3 (3xTF32 tensor cores, fp32-grade) or IMPL_TCGEN05_MIXED (3xTF32 forward, TF32 backward)."""
    global _default_impl
    _default_impl = int(impl)


def get_default_impl():
    return _default_impl


_direct_wgrad = False


def set_direct_weight_grad(flag):
    """When on, the weight-gradient kernels ACCUMULATE straight into `param.grad` (which must exist, be zeroed
    before backward and keep its address — FlatSGDTrainer's flat gradient buffer) and autograd receives no
    weight gradient: no per-parameter AccumulateGrad add kernels.  Off: gradients are returned to autograd."""
    global _direct_wgrad
    _direct_wgrad = bool(flag)


_wgrad_side = None


def set_wgrad_side(main_stream, side_stream, wgrad_sms):
    """Run the direct-mode weight-gradient kernels launched from `main_stream` on `side_stream`, limited to `wgrad_sms`
    SMs, while the data-gradient kernels of the fused stages keep the other SMs (None, ... = off).  The weight and data
    gradient of a layer are independent, and the small layers of res3 / res4 leave most of a 148-SM launch idle (their
    split-K items are short and end in a burst of reductions): two half-width kernels side by side finish sooner than
    two full-width ones in a row.  The caller joins with wgrad_side_join() before it reads the gradients."""
    global _wgrad_side
    _wgrad_side = None if main_stream is None else dict(main=main_stream, side=side_stream, w=int(wgrad_sms),
                                                        d=NUM_SMS - int(wgrad_sms), used=False)


def wgrad_side_join():
    cfg = _wgrad_side
    if cfg is not None and cfg["used"]:
        cfg["main"].wait_stream(cfg["side"])
        cfg["used"] = False


def _wgrad_side_active():
    # TF32 backward arms only (tcgen05, mixed): the SIMT weight-gradient kernels share one split-K scratch buffer per
    # device, and on the all-3xTF32 arm the weight gradients are twice the work of the data gradients — an even split
    # of the SMs slows it down (27.4 -> 33.3 ms)
    cfg = _wgrad_side
    return cfg is not None and bwd_impl() == IMPL_TCGEN05 and torch.cuda.current_stream() == cfg["main"]


def _wgrad_into(param, gy, x, scale, cout, kh, kw, stride, pad):
    """Weight gradient of one conv: returned for autograd, or (direct mode) accumulated into param.grad."""
    if _direct_wgrad and param.is_leaf and param.grad is not None:
        g = param.grad
        phys = g.permute(0, 2, 3, 1) if g.dim() == 4 else g
        if phys.is_contiguous() and phys.data_ptr() % 16 == 0:
            if gy.is_cuda and _wgrad_side_active():
                cfg = _wgrad_side
                cfg["side"].wait_stream(cfg["main"])          # gy and x are complete
                gy.record_stream(cfg["side"])
                x.record_stream(cfg["side"])
                with torch.cuda.stream(cfg["side"]), sm_budget(cfg["w"]):
                    conv2d_wgrad_raw(gy, x, scale, cout, kh, kw, stride, pad, out=phys, accumulate=True)
                cfg["used"] = True
                return None
            conv2d_wgrad_raw(gy, x, scale, cout, kh, kw, stride, pad, out=phys, accumulate=True)
            return None
    return grad_like_weight(conv2d_wgrad_raw(gy, x, scale, cout, kh, kw, stride, pad), param)


_milestone_cb = None


def set_grad_milestone_callback(fn):
    """fn(tag) is called during backward as soon as the gradient of a tensor marked with grad_milestone(x, tag) is
    complete, i.e. when every kernel of the layers downstream of x has been launched (None = off).  FlatSGDTrainer
    uses it to start the gradient all-reduce of a finished parameter segment while backward is still running."""
    global _milestone_cb
    _milestone_cb = fn


def grad_milestone(x, tag):
    if _milestone_cb is not None and torch.is_tensor(x) and x.requires_grad:
        x.register_hook(lambda g, t=tag: (_milestone_cb(t), None)[1])
    return x


class sm_budget(object):
    """Context manager: the persistent dense kernels launched inside occupy at most `sms` SMs (baked into the launches
    of a stream capture).  For dense work that runs beside few-CTA latency-bound kernels on another stream."""

    def __init__(self, sms):
        self.sms = int(sms)

    def __enter__(self):
        self.prev = _lib.load().dd_set_sm_budget(self.sms)
        return self

    def __exit__(self, *exc):
        _lib.load().dd_set_sm_budget(self.prev)
        return False


class _InjectGrad(torch.autograd.Function):
    """value(loss) with the gradient of `holder` handed to `x` in backward: the join of an early backward pass.
    A loss that depends on `x` only through `cut = x.detach().requires_grad_()` has been back-propagated already
    (cut.grad is complete); this node re-attaches that gradient to the graph of x.  Exact for an incoming gradient
    of 1 (losses summed with unit weights, as the reference loop does: engine/trainer.py:228-231)."""

    @staticmethod
    def forward(ctx, loss, x, holder):
        ctx.holder = holder
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, g):
        gx = ctx.holder.grad
        ctx.holder = None
        return None, gx, None


def inject_grad(loss, x, holder):
    return _InjectGrad.apply(loss, x, holder)


def tcgen05_available():
    return bool(_lib.load().dd_tcgen05_built())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype=torch.float32, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("dadetect_b200: {} must be a CUDA tensor — there is no CPU path".format(name))
    if t.dtype != dtype:
        raise RuntimeError("dadetect_b200: {} must be {}, got {}".format(name, dtype, t.dtype))
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0 and t.numel() > 0:
        t = t.clone()
    return t


_workspaces = {}
_retired = []          # outgrown scratch buffers: a captured CUDA graph may still hold their addresses


def _workspace(nbytes, device, tag="ws"):
    """Grow-only scratch buffer per (device, tag); reused across calls on the same stream.  A buffer that is
    outgrown is retired, not freed: kernels recorded into a CUDA graph keep using the address they were
    captured with."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired.append(buf)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def weight_ohwi(w):
    """[Cout,Cin,KH,KW] (any strides) or [Cout,Cin] -> contiguous physical OHWI view (no copy when the
    parameter is already channels_last)."""
    if w.dim() == 2:
        return w if w.is_contiguous() else w.contiguous()
    return w.permute(0, 2, 3, 1).contiguous()


def grad_like_weight(g_ohwi, w):
    """physical OHWI gradient -> tensor with w's logical shape (channels_last strides for 4-D)."""
    if w.dim() == 2:
        return g_ohwi.view_as(w)
    co, ci, kh, kw = w.shape
    return g_ohwi.view(co, kh, kw, ci).permute(0, 3, 1, 2)


class WeightPrepPlan(object):
    """Per-step preparation of the dense tier's weight operands, batched.  Every 3xTF32 forward conv splits its weight
    into hi / lo planes and every tensor-core data gradient builds the flipped, BN-scaled transpose W' of its weight —
    ~60 + ~80 small launches per training step, each between two large persistent kernels.  A trainer records the
    (weight, shape) pairs of its first step (`recording`), and from then on prepares ALL of them with a handful of
    batched launches at the start of the step (run()); the conv calls of that step find their operand in the plan by
    the weight's address and skip their own preparation.  Only tensors that live in storage the trainer declared stable
    (parameters, buffers) are recorded, and the plan keeps them alive, so an address can never come to mean another
    tensor.  Outside `weight_prep(plan)` nothing is looked up: the operands are only valid for the step they were
    prepared in (the optimiser changes the weights)."""

    def __init__(self, stable_storages):
        self.stable = set(stable_storages)
        self.fwd, self.dgrad = {}, {}          # key -> [tensors kept alive..., dims, workspace]
        self.recording = True

    def is_stable(self, t):
        return t is None or t.untyped_storage().data_ptr() in self.stable

    def finalize(self):
        self.recording = False

    EARLY_BYTES = 8 << 20        # forward operands prepared on the step's own stream before anything else runs

    def _split(self, ents):
        n = len(ents)
        if n == 0:
            return
        vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
        _lib.call("dd_conv2d_forward_prepare_batch", n, vp(*[e["w"].data_ptr() for e in ents]),
                  vp(*[e["ws"].data_ptr() for e in ents]), ip(*[e["dims"][0] for e in ents]),
                  ip(*[e["dims"][1] for e in ents]), ip(*[e["dims"][2] for e in ents]),
                  ip(*[e["dims"][3] for e in ents]), IMPL_TCGEN05_X3, _stream())

    def run(self):
        """Prepare every recorded operand.  The forward operands of the first layers (in the order the step uses them,
        EARLY_BYTES worth: stem, res2, res3) on the current stream; everything else — 430 MB of split planes for res4,
        the RPN and the heads, and the data-gradient transposes that only the backward pass reads — on a side stream
        beside the memory-bound start of the trunk.  A consumer stream waits for the side stream the first time it
        looks one of those operands up (`_join`)."""
        main = torch.cuda.current_stream()
        ents = list(self.fwd.values())                   # insertion order = order of first use
        early, late, acc = [], [], 0
        for e in ents:
            acc += e["ws"].numel() * 4
            (early if acc <= self.EARLY_BYTES else late).append(e)
            e["late"] = acc > self.EARLY_BYTES
        self._split(early)
        self.joined = set()
        if not late and not self.dgrad:
            self.side = None
            return
        if getattr(self, "side", None) is None:
            self.side = torch.cuda.Stream(device=ents[0]["ws"].device if ents else None)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self._split(late)
            for impl in sorted({e["impl"] for e in self.dgrad.values()}):
                ds = [e for e in self.dgrad.values() if e["impl"] == impl]
                n = len(ds)
                vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
                _lib.call("dd_conv2d_dgrad_prepare_batch", n, vp(*[e["w"].data_ptr() for e in ds]),
                          vp(*[None if e["scale"] is None else e["scale"].data_ptr() for e in ds]),
                          vp(*[e["ws"].data_ptr() for e in ds]), ip(*[e["dims"][0] for e in ds]),
                          ip(*[e["dims"][1] for e in ds]), ip(*[e["dims"][2] for e in ds]),
                          ip(*[e["dims"][3] for e in ds]), impl, _stream())

    def _join(self):
        """The current stream is about to read an operand prepared on the side stream."""
        side = getattr(self, "side", None)
        if side is None:
            return
        cur = torch.cuda.current_stream()
        if cur.cuda_stream not in self.joined:
            cur.wait_stream(side)
            self.joined.add(cur.cuda_stream)


_prep_plan = None


class weight_prep(object):
    """Context manager around one training step: `plan` is consulted (and, while it is recording, extended) by the
    conv calls inside; a finished plan has its operands prepared on entry."""

    def __init__(self, plan):
        self.plan = plan

    def __enter__(self):
        global _prep_plan
        self.prev = _prep_plan
        _prep_plan = self.plan
        if self.plan is not None and not self.plan.recording:
            self.plan.run()
        return self.plan

    def __exit__(self, *exc):
        global _prep_plan
        _prep_plan = self.prev
        if self.plan is not None and self.plan.recording and exc[0] is None:
            self.plan.finalize()
        return False


def _prep_fwd(w_ohwi, cin, cout, kh, kw):
    """The prepared hi / lo planes of a forward weight for this step, or None."""
    plan = _prep_plan
    if plan is None:
        return None
    key = (w_ohwi.data_ptr(), cin, cout, kh, kw)
    ent = plan.fwd.get(key)
    if ent is not None:
        if plan.recording:
            return None
        if ent.get("late"):
            plan._join()
        return ent["ws"]
    if plan.recording and plan.is_stable(w_ohwi):
        nbytes = _lib.load().dd_conv2d_forward_workspace_bytes(cin, cout, kh, kw, IMPL_TCGEN05_X3)
        plan.fwd[key] = dict(w=w_ohwi, dims=(cin, cout, kh, kw),
                             ws=torch.empty(max(int(nbytes) // 4, 4), dtype=torch.float32, device=w_ohwi.device))
    return None


def _prep_dgrad(w_ohwi, scale, cin, cout, kh, kw, impl):
    """The prepared dgrad weights W' of a layer for this step, or None."""
    plan = _prep_plan
    if plan is None or impl == IMPL_SIMT:
        return None
    key = (w_ohwi.data_ptr(), 0 if scale is None else scale.data_ptr(), cin, cout, kh, kw, impl)
    ent = plan.dgrad.get(key)
    if ent is not None:
        if plan.recording:
            return None
        plan._join()
        return ent["ws"]
    if plan.recording and plan.is_stable(w_ohwi):          # (the plan keeps `scale`, a cached derived tensor, alive)
        plan.dgrad[key] = dict(w=w_ohwi, scale=scale, dims=(cin, cout, kh, kw), impl=impl,
                               ws=dgrad_workspace(cin, cout, kh, kw, w_ohwi.device))
    return None


# ----------------------------------------------------------------------------------------- raw kernels
def conv2d_forward_raw(x, w_ohwi, scale, bias, residual, kh, kw, stride, pad, relu, impl=None):
    n, h, wd, cin = x.shape
    cout = w_ohwi.shape[0]
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (wd + 2 * pad - kw) // stride + 1
    y = torch.empty((n, oh, ow, cout), dtype=torch.float32, device=x.device)
    impl = fwd_impl(impl)
    ws, entry = None, "dd_conv2d_forward"
    if impl == IMPL_TCGEN05_X3:
        ws = _prep_fwd(w_ohwi, cin, cout, kh, kw)          # prepared for this step by the trainer's plan?
        if ws is not None:
            entry = "dd_conv2d_forward_prepared"
        else:
            nbytes = _lib.load().dd_conv2d_forward_workspace_bytes(cin, cout, kh, kw, impl)
            ws = torch.empty(max(int(nbytes) // 4, 4), dtype=torch.float32, device=x.device)
    _lib.call(entry, _ptr(x), _ptr(w_ohwi), _ptr(scale), _ptr(bias), _ptr(residual), _ptr(y),
              n, h, wd, cin, cout, kh, kw, stride, pad, 1 if relu else 0, impl, _ptr(ws), _stream())
    return y


def conv2d_dgrad_raw(gy, w_ohwi, scale, x_shape, kh, kw, stride, pad, addend=None, mask_act=None, impl=None,
                     prepared_ws=None):
    """prepared_ws: a float buffer that already holds this layer's dgrad weights W' (see dgrad_workspace)."""
    n, h, wd, cin = x_shape
    cout = w_ohwi.shape[0]
    gx = torch.empty(x_shape, dtype=torch.float32, device=gy.device)
    if prepared_ws is None:
        prepared_ws = _prep_dgrad(w_ohwi, scale, cin, cout, kh, kw, bwd_impl(impl))
    ws = prepared_ws if prepared_ws is not None else dgrad_workspace(cin, cout, kh, kw, gy.device)
    _lib.call("dd_conv2d_dgrad", _ptr(gy), _ptr(w_ohwi), _ptr(scale), _ptr(addend), _ptr(mask_act), _ptr(gx),
              n, h, wd, cin, cout, kh, kw, stride, pad, bwd_impl(impl), _ptr(ws),
              1 if prepared_ws is not None else 0, _stream())
    return gx


def dgrad_workspace(cin, cout, kh, kw, device):
    """Scratch for the tcgen05 arm's prepared dgrad weights (stream-ordered reuse through torch's allocator)."""
    nbytes = _lib.load().dd_conv2d_dgrad_workspace_bytes(cin, cout, kh, kw)
    return torch.empty(max(int(nbytes) // 4, 4), dtype=torch.float32, device=device)


def dgrad_prepare_batch(layers, device, impl=None):
    """Prepare the dgrad weights W' of several layers with one launch per 16 layers (a ResNet stage does all of its
    layers up front).  layers: list of (w_ohwi, scale | None); returns one prepared workspace per layer, or a list of
    None on the SIMT arm (which consumes the weights as they are)."""
    impl = bwd_impl(impl)
    n = len(layers)
    if impl == IMPL_SIMT or n == 0:
        return [None] * n
    dims = []
    for w, _ in layers:
        if w.dim() == 4:
            dims.append((w.shape[3], w.shape[0], w.shape[1], w.shape[2]))      # OHWI -> (Cin, Cout, KH, KW)
        else:
            dims.append((w.shape[1], w.shape[0], 1, 1))
    planned = [_prep_dgrad(w, sc, ci, co, kh, kw, impl) for (w, sc), (ci, co, kh, kw) in zip(layers, dims)]
    if all(p is not None for p in planned):            # prepared for this step by the trainer's plan
        return planned
    wss = [dgrad_workspace(ci, co, kh, kw, device) for ci, co, kh, kw in dims]
    vp, ip = ctypes.c_void_p * n, ctypes.c_int * n
    _lib.call("dd_conv2d_dgrad_prepare_batch", n, vp(*[w.data_ptr() for w, _ in layers]),
              vp(*[None if s is None else s.data_ptr() for _, s in layers]), vp(*[t.data_ptr() for t in wss]),
              ip(*[d[0] for d in dims]), ip(*[d[1] for d in dims]), ip(*[d[2] for d in dims]),
              ip(*[d[3] for d in dims]), impl, _stream())
    return wss


def conv2d_wgrad_raw(gy, x, scale, cout, kh, kw, stride, pad, out=None, accumulate=False, impl=None):
    n, h, wd, cin = x.shape
    if out is None:
        out = torch.empty((cout, kh, kw, cin), dtype=torch.float32, device=x.device)
    nbytes = _lib.load().dd_conv2d_wgrad_workspace_bytes(n, h, wd, cin, cout, kh, kw, stride, pad)
    ws = _workspace(nbytes, x.device, "wgrad")
    _lib.call("dd_conv2d_wgrad", _ptr(gy), _ptr(x), _ptr(scale), _ptr(out), n, h, wd, cin, cout, kh, kw, stride, pad,
              1 if accumulate else 0, bwd_impl(impl), _ptr(ws), _stream())
    return out


def bias_grad_raw(gy2d, c):
    rows = gy2d.numel() // c
    gb = torch.empty((c,), dtype=torch.float32, device=gy2d.device)
    _lib.call("dd_bias_grad", _ptr(gy2d), _ptr(gb), rows, c, 0, _stream())
    return gb


def relu_backward_raw(g, act):
    out = torch.empty_like(g)
    _lib.call("dd_relu_backward", _ptr(g), _ptr(act), _ptr(out), g.numel(), _stream())
    return out


def nchw_to_nhwc(x):
    x = _chk(x, name="input")
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    _lib.call("dd_nchw_to_nhwc", _ptr(x), _ptr(y), n, c, h, w, _stream())
    return y


def nhwc_to_nchw(x):
    x = _chk(x, name="input")
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _lib.call("dd_nhwc_to_nchw", _ptr(x), _ptr(y), n, c, h, w, _stream())
    return y


def stem_conv7x7s2(x_nchw, weight, scale, bias, relu=True):
    """BaseStem conv (7x7/2, 3 -> Cout<=64) + FrozenBN + ReLU on the tensor cores, straight from the
    reference's NCHW image (the layout conversion is part of the operand staging); returns NHWC.  The stem is
    frozen in every DA config (FREEZE_CONV_BODY_AT = 2), so this op is forward-only."""
    x = _chk(x_nchw, name="images")
    n, c, h, w = x.shape
    cout = weight.shape[0]
    if c != 3 or tuple(weight.shape[1:]) != (3, 7, 7):
        raise RuntimeError("dadetect_b200: stem_conv7x7s2 expects [N,3,H,W] images and a [Cout,3,7,7] weight")
    y = torch.empty((n, h // 2, w // 2, cout), dtype=torch.float32, device=x.device)
    ws = _workspace(_lib.load().dd_stem_workspace_bytes(n, h, w, cout), x.device, "stem")
    _lib.call("dd_stem_conv7x7s2_forward", _ptr(x), _ptr(weight_ohwi(weight.detach())), _ptr(scale), _ptr(bias),
              _ptr(y), n, h, w, cout, 1 if relu else 0, fwd_impl(), _ptr(ws), _stream())
    return y


def stem_tc_supported(x_nchw, weight):
    return (_default_impl in TC_IMPLS and x_nchw.shape[1] == 3 and x_nchw.shape[2] % 2 == 0
            and x_nchw.shape[3] % 2 == 0 and tuple(weight.shape[1:]) == (3, 7, 7) and weight.shape[0] % 4 == 0
            and weight.shape[0] <= 64 and not weight.requires_grad)


def maxpool3x3s2(x):
    n, h, w, c = x.shape
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, oh, ow, c), dtype=torch.float32, device=x.device)
    _lib.call("dd_maxpool3x3s2", _ptr(x), _ptr(y), n, h, w, c, _stream())
    return y


# ----------------------------------------------------------------------------------------- dense layers
class _ConvBnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, scale, bias, residual, stride, pad, relu):
        x = _chk(x, name="x")
        w = weight_ohwi(weight)
        if weight.dim() == 4:
            kh, kw = weight.shape[2], weight.shape[3]
        else:
            kh = kw = 1
        residual = _chk(residual, name="residual")
        y = conv2d_forward_raw(x, w, scale, bias, residual, kh, kw, stride, pad, relu)
        ctx.save_for_backward(x, weight, scale, y if relu else None)
        ctx.cfg = (kh, kw, stride, pad, relu, bias is not None, residual is not None)
        ctx.weight_ref = weight
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, scale, y = ctx.saved_tensors
        kh, kw, stride, pad, relu, has_bias, has_res = ctx.cfg
        gy = _chk(gy, name="grad")
        g = relu_backward_raw(gy, y) if relu else gy
        w = weight_ohwi(weight)
        gx = gw = gb = gres = None
        if ctx.needs_input_grad[0]:
            if g.is_cuda and _wgrad_side_active():
                with sm_budget(_wgrad_side["d"]):
                    gx = conv2d_dgrad_raw(g, w, scale, tuple(x.shape), kh, kw, stride, pad)
            else:
                gx = conv2d_dgrad_raw(g, w, scale, tuple(x.shape), kh, kw, stride, pad)
        if ctx.needs_input_grad[1]:
            gw = _wgrad_into(ctx.weight_ref, g, x, scale, w.shape[0], kh, kw, stride, pad)
        if has_bias and ctx.needs_input_grad[3]:
            gb = bias_grad_raw(g, g.shape[-1])
        if has_res and ctx.needs_input_grad[4]:
            gres = g
        return gx, gw, None, gb, gres, None, None, None


def conv_bn_act(x, weight, scale=None, bias=None, residual=None, stride=1, pad=0, relu=False):
    """act(conv(x, weight) * scale + bias + residual); x NHWC, weight [Cout,Cin,KH,KW]."""
    return _ConvBnAct.apply(x, weight, scale, bias, residual, stride, pad, relu)


class _BottleneckStage(torch.autograd.Function):
    """A whole ResNet stage (make_stage, resnet.py:197-224: a run of BottleneckWithFixedBatchNorm blocks) as ONE
    autograd node.  Forward = the same fused conv+FrozenBN(+residual)+ReLU kernels as conv_bn_act.  Backward is
    written out explicitly so that every ReLU mask and every residual fan-in add rides in a dgrad epilogue:

        g2 = dgrad3(g_out)            * (y2 > 0)
        g1 = dgrad2(g2)               * (y1 > 0)
        gx = (dgrad1(g1) + g_out)     * (x > 0)      identity block   (x is the previous block's ReLU output, so
        gx = (dgrad1(g1) + dgradD(g_out)) * (x > 0)  downsample block  gx is already that block's masked g_out)

    i.e. no relu_backward and no gradient-add kernels inside a stage (the reference's autograd graph runs
    threshold_backward, the BN-scale multiply and the add as separate passes over every activation).

    flags: input_is_relu  — the stage input is a ReLU output whose ONLY consumer is this stage, and its producer
                            accepts an already-masked gradient (then gx is returned masked by x > 0);
           grad_premasked — the incoming gradient is already masked by (output > 0).
    """

    @staticmethod
    def forward(ctx, x, meta, *tensors):
        strides, has_down, input_is_relu, grad_premasked, pool_output = meta
        x = _chk(x, name="x")
        need_graph = x.requires_grad or any(t.requires_grad for t in tensors)
        saved, per_block = [], []
        k = 0
        cur = x
        for bi, stride in enumerate(strides):
            w1, s1, b1, w2, s2, b2, w3, s3, b3 = tensors[k:k + 9]
            k += 9
            wd = sd = bd = None
            if has_down[bi]:
                wd, sd, bd = tensors[k:k + 3]
                k += 3
            y1 = conv2d_forward_raw(cur, weight_ohwi(w1), s1, b1, None, 1, 1, stride, 0, True)
            y2 = conv2d_forward_raw(y1, weight_ohwi(w2), s2, b2, None, 3, 3, 1, 1, True)
            identity = cur
            if wd is not None:
                identity = conv2d_forward_raw(cur, weight_ohwi(wd), sd, bd, None, 1, 1, stride, 0, False)
            out = conv2d_forward_raw(y2, weight_ohwi(w3), s3, b3, identity, 1, 1, 1, 0, True)
            if need_graph:
                saved.extend([cur, y1, y2])
            cur = out
        if need_graph:
            saved.append(cur)
            ctx.save_for_backward(*saved, *tensors)
            ctx.n_saved = len(saved)
            ctx.meta = meta
            ctx.param_refs = tensors
        if pool_output:                       # nn.AvgPool2d over the whole map: [K,h,w,C] -> [K,C]
            k, h, w, c = cur.shape
            pooled = torch.empty((k, c), dtype=torch.float32, device=cur.device)
            _lib.call("dd_avgpool_forward", _ptr(cur), _ptr(pooled), k, h * w, c, _stream())
            return pooled
        return cur

    @staticmethod
    def backward(ctx, g):
        if g.is_cuda and _wgrad_side_active():             # the data-gradient kernels leave the side stream its SMs
            with sm_budget(_wgrad_side["d"]):
                return _BottleneckStage._backward(ctx, g)
        return _BottleneckStage._backward(ctx, g)

    @staticmethod
    def _backward(ctx, g):
        strides, has_down, input_is_relu, grad_premasked, pool_output = ctx.meta
        all_saved = ctx.saved_tensors
        acts, tensors = all_saved[:ctx.n_saved], all_saved[ctx.n_saved:]
        y_last = acts[-1]
        g = _chk(g, name="grad")
        if pool_output:                       # pooling backward and the last ReLU mask in one pass
            k, h, w, c = y_last.shape
            g_out = torch.empty_like(y_last)
            _lib.call("dd_avgpool_relu_backward", _ptr(g), _ptr(y_last), _ptr(g_out), k, h * w, c, _stream())
        else:
            g_out = g if grad_premasked else relu_backward_raw(g, y_last)
        grads = [None] * len(tensors)
        # tensor offsets of each block
        offs, k = [], 0
        for bi in range(len(strides)):
            offs.append(k)
            k += 12 if has_down[bi] else 9
        need_x = ctx.needs_input_grad[0]
        gx = None
        # the dgrad weights of every layer of the stage, prepared by one launch (per 16 layers)
        ohwi, todo = {}, []
        for bi in range(len(strides)):
            o = offs[bi]
            for j in (0, 3, 6) + ((9,) if has_down[bi] else ()):
                ohwi[o + j] = weight_ohwi(tensors[o + j])
                if j in (3, 6) or bi > 0 or need_x:
                    todo.append(o + j)
        prep = dict(zip(todo, dgrad_prepare_batch([(ohwi[i], tensors[i + 1]) for i in todo], g_out.device)))
        for bi in reversed(range(len(strides))):
            o = offs[bi]
            w1, s1, _, w2, s2, _, w3, s3, _ = tensors[o:o + 9]
            x_in, y1, y2 = acts[3 * bi:3 * bi + 3]
            stride = strides[bi]
            first = bi == 0
            w1o, w2o, w3o = ohwi[o], ohwi[o + 3], ohwi[o + 6]
            refs = ctx.param_refs
            if ctx.needs_input_grad[2 + o + 6]:
                grads[o + 6] = _wgrad_into(refs[o + 6], g_out, y2, s3, w3o.shape[0], 1, 1, 1, 0)
            g2 = conv2d_dgrad_raw(g_out, w3o, s3, tuple(y2.shape), 1, 1, 1, 0, mask_act=y2, prepared_ws=prep[o + 6])
            if ctx.needs_input_grad[2 + o + 3]:
                grads[o + 3] = _wgrad_into(refs[o + 3], g2, y1, s2, w2o.shape[0], 3, 3, 1, 1)
            g1 = conv2d_dgrad_raw(g2, w2o, s2, tuple(y1.shape), 3, 3, 1, 1, mask_act=y1, prepared_ws=prep[o + 3])
            del g2
            if ctx.needs_input_grad[2 + o]:
                grads[o] = _wgrad_into(refs[o], g1, x_in, s1, w1o.shape[0], 1, 1, stride, 0)
            wd = sd = None
            if has_down[bi]:
                wd, sd = tensors[o + 9], tensors[o + 10]
                if ctx.needs_input_grad[2 + o + 9]:
                    grads[o + 9] = _wgrad_into(refs[o + 9], g_out, x_in, sd, ohwi[o + 9].shape[0], 1, 1, stride, 0)
            if first and not need_x:
                break
            # data gradient of the block input: conv1 path + residual path, masked by the producer's ReLU
            mask = x_in if (not first or input_is_relu) else None
            if wd is not None:
                t = conv2d_dgrad_raw(g_out, ohwi[o + 9], sd, tuple(x_in.shape), 1, 1, stride, 0,
                                     prepared_ws=prep[o + 9])
            else:
                t = g_out
            g_out = conv2d_dgrad_raw(g1, w1o, s1, tuple(x_in.shape), 1, 1, stride, 0, addend=t, mask_act=mask,
                                     prepared_ws=prep[o])
            del g1, t
            if first:
                gx = g_out
        return (gx, None) + tuple(grads)


def bottleneck_stage(x, blocks, strides, input_is_relu=False, grad_premasked=False, pool_output=False):
    """blocks: list of dicts with w1,s1,b1,w2,s2,b2,w3,s3,b3 and optionally wd,sd,bd (tensors).
    pool_output: return the spatial average [K,C] of the stage output (the box head's AvgPool2d(7))."""
    tensors, has_down = [], []
    for b in blocks:
        tensors.extend([b["w1"], b["s1"], b["b1"], b["w2"], b["s2"], b["b2"], b["w3"], b["s3"], b["b3"]])
        has_down.append("wd" in b)
        if "wd" in b:
            tensors.extend([b["wd"], b["sd"], b["bd"]])
    meta = (tuple(int(s) for s in strides), tuple(has_down), bool(input_is_relu), bool(grad_premasked),
            bool(pool_output))
    return _BottleneckStage.apply(x, meta, *tensors)


def linear(x, weight, bias=None, relu=False):
    """F.linear (+ReLU) on [R, Cin] through the same implicit-GEMM kernels (a 1x1 conv on R 1x1 'images')."""
    r, cin = x.shape
    y = _ConvBnAct.apply(x.view(r, 1, 1, cin), weight, None, bias, None, 1, 0, relu)
    return y.view(r, weight.shape[0])


def fused_heads(x, weights, biases):
    """Several narrow output heads that read the same input as ONE tensor-core GEMM: the heads' rows are stacked and
    zero-padded to a multiple of 32 output channels (the K block of the data-gradient GEMM and the M granularity of the
    weight-gradient GEMM), so that forward, dgrad and wgrad all run on tcgen05 — a 1-, 9-, 15- or 36-channel head by
    itself falls to the SIMT kernels in the backward pass.  The pad rows are zero and receive zero gradient.
    x: [..., Cin] (NHWC map or [R, Cin] rows); weights: list of [Cout_i, Cin] or [Cout_i, Cin, 1, 1]; returns one
    contiguous tensor per head.  On the SIMT arm the heads run one by one."""
    lead, cin = x.shape[:-1], x.shape[-1]
    if _default_impl not in TC_IMPLS:
        outs = []
        for w, b in zip(weights, biases):
            y = _ConvBnAct.apply(x.reshape(-1, 1, 1, cin), w.reshape(w.shape[0], cin), None, b, None, 1, 0, False)
            outs.append(y.view(*lead, w.shape[0]))
        return outs
    sizes = [w.shape[0] for w in weights]
    pad = (-sum(sizes)) % 32
    rows = [w.reshape(w.shape[0], cin) for w in weights]
    bs = list(biases)
    if pad:
        rows.append(rows[0].new_zeros(pad, cin))
        bs.append(bs[0].new_zeros(pad))
    wcat, bcat = torch.cat(rows, dim=0), torch.cat(bs)
    x4 = x if x.dim() == 4 else x.reshape(-1, 1, 1, cin)
    y = _ConvBnAct.apply(x4, wcat, None, bcat, None, 1, 0, False).view(*lead, wcat.shape[0])
    outs, o = [], 0
    for n in sizes:
        outs.append(y[..., o:o + n].contiguous())
        o += n
    return outs


class _AvgPoolHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x, name="x")
        k, h, w, c = x.shape
        y = torch.empty((k, c), dtype=torch.float32, device=x.device)
        _lib.call("dd_avgpool_forward", _ptr(x), _ptr(y), k, h * w, c, _stream())
        ctx.shape = (k, h, w, c)
        return y

    @staticmethod
    def backward(ctx, gy):
        k, h, w, c = ctx.shape
        gy = _chk(gy, name="grad")
        gx = torch.empty(ctx.shape, dtype=torch.float32, device=gy.device)
        _lib.call("dd_avgpool_backward", _ptr(gy), _ptr(gx), k, h * w, c, _stream())
        return gx


def avgpool_hw(x):
    """nn.AvgPool2d(7) on a [K,7,7,C] NHWC ROI feature -> [K,C]."""
    return _AvgPoolHW.apply(x)


class _RoIAlign(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, spatial_scale, pooled, sampling_ratio, bin_step):
        feat = _chk(feat, name="feat")
        rois = _chk(rois, name="rois")
        n, h, w, c = feat.shape
        k = rois.shape[0]
        o = (pooled + bin_step - 1) // bin_step
        out = torch.empty((k, o, o, c), dtype=torch.float32, device=feat.device)
        _lib.call("dd_roi_align_forward", _ptr(feat), _ptr(rois), _ptr(out), n, h, w, c, k, float(spatial_scale),
                  pooled, pooled, sampling_ratio, bin_step, _stream())
        ctx.save_for_backward(rois)
        ctx.meta = (n, h, w, c, k, float(spatial_scale), pooled, sampling_ratio, bin_step)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        n, h, w, c, k, scale, pooled, sr, bin_step = ctx.meta
        g = _chk(g, name="grad")
        gfeat = torch.zeros((n, h, w, c), dtype=torch.float32, device=g.device)
        _lib.call("dd_roi_align_backward", _ptr(g), _ptr(rois), _ptr(gfeat), n, h, w, c, k, scale, pooled, pooled, sr,
                  bin_step, _stream())
        return gfeat, None, None, None, None, None


def roi_align(feat, rois, spatial_scale, pooled, sampling_ratio, bin_step=1):
    """ROIAlign on an NHWC map; bin_step=2 returns only the even bins (see include/dadetect_b200.h)."""
    return _RoIAlign.apply(feat, rois, spatial_scale, pooled, sampling_ratio, bin_step)


class _RoIAlignLevels(torch.autograd.Function):
    """Multi-level Pooler.forward (poolers.py:104-121): LevelMapper on the device, then one launch per level over
    all ROIs (a ROI is pooled only from its own level's map) — no nonzero / index_put, no host read."""

    @staticmethod
    def forward(ctx, rois, scales, pooled, sampling_ratio, k_min, k_max, *feats):
        rois = _chk(rois, name="rois")
        feats = [_chk(f, name="feat") for f in feats]
        k = rois.shape[0]
        c = feats[0].shape[3]
        levels = torch.empty((max(k, 1),), dtype=torch.int32, device=rois.device)
        _lib.call("dd_fpn_level_map", _ptr(rois), k, int(k_min), int(k_max), 224.0, 4, 1e-6, _ptr(levels), _stream())
        out = torch.empty((k, pooled, pooled, c), dtype=torch.float32, device=rois.device)
        for lvl, (f, sc) in enumerate(zip(feats, scales)):
            n, h, w, _ = f.shape
            _lib.call("dd_roi_align_level_forward", _ptr(f), _ptr(rois), _ptr(levels), lvl, _ptr(out), n, h, w, c, k,
                      float(sc), pooled, pooled, int(sampling_ratio), _stream())
        ctx.save_for_backward(rois, levels)
        ctx.meta = ([tuple(f.shape) for f in feats], tuple(float(sc) for sc in scales), pooled, int(sampling_ratio))
        ctx.mark_non_differentiable(levels)
        return out, levels

    @staticmethod
    def backward(ctx, g, _glevels):
        rois, levels = ctx.saved_tensors
        shapes, scales, pooled, sr = ctx.meta
        g = _chk(g, name="grad")
        k = rois.shape[0]
        grads = []
        for lvl, (shape, sc) in enumerate(zip(shapes, scales)):
            if not ctx.needs_input_grad[6 + lvl]:
                grads.append(None)
                continue
            n, h, w, c = shape
            gf = torch.zeros(shape, dtype=torch.float32, device=g.device)
            _lib.call("dd_roi_align_level_backward", _ptr(g), _ptr(rois), _ptr(levels), lvl, _ptr(gf), n, h, w, c, k,
                      sc, pooled, pooled, sr, _stream())
            grads.append(gf)
        return (None, None, None, None, None, None) + tuple(grads)


def roi_align_levels(feats, rois, scales, pooled, sampling_ratio):
    """feats: list of NHWC maps (finest first), scales their spatial scales -> (out [K,pooled,pooled,C], levels int32
    [K]); level range from the scales as in Pooler.__init__ (poolers.py:71-73)."""
    import math
    k_min, k_max = -math.log2(scales[0]), -math.log2(scales[-1])
    return _RoIAlignLevels.apply(rois, tuple(scales), int(pooled), int(sampling_ratio), int(k_min), int(k_max), *feats)


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x, name="x")
        n, h, w, c = x.shape
        y = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.float32, device=x.device)
        _lib.call("dd_upsample2x_forward", _ptr(x), _ptr(y), n, h, w, c, _stream())
        ctx.shape = (n, h, w, c)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, h, w, c = ctx.shape
        gy = _chk(gy, name="grad")
        gx = torch.empty(ctx.shape, dtype=torch.float32, device=gy.device)
        _lib.call("dd_upsample2x_backward", _ptr(gy), _ptr(gx), n, h, w, c, _stream())
        return gx


def upsample2x(x):
    """F.interpolate(x, scale_factor=2, mode="nearest") on NHWC (fpn.py:62)."""
    return _Upsample2x.apply(x)


class _Subsample2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x, name="x")
        n, h, w, c = x.shape
        y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=torch.float32, device=x.device)
        _lib.call("dd_subsample2_forward", _ptr(x), _ptr(y), n, h, w, c, _stream())
        ctx.shape = (n, h, w, c)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, h, w, c = ctx.shape
        gy = _chk(gy, name="grad")
        gx = torch.empty(ctx.shape, dtype=torch.float32, device=gy.device)
        _lib.call("dd_subsample2_backward", _ptr(gy), _ptr(gx), n, h, w, c, _stream())
        return gx


def subsample2(x):
    """LastLevelMaxPool: F.max_pool2d(x, 1, 2, 0) on NHWC (fpn.py:80-82)."""
    return _Subsample2.apply(x)


class _GradientScalar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        ctx.weight = float(weight)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = _chk(g, name="grad")
        out = torch.empty_like(g)
        _lib.call("dd_grl_backward", _ptr(g), ctx.weight, _ptr(out), g.numel(), 0, _stream())
        return out, None


def gradient_scalar(x, weight):
    return _GradientScalar.apply(x, weight)


class _GradientScalarDev(torch.autograd.Function):
    """GRL whose weight lives in a device float[1] that may be written AFTER the forward pass."""

    @staticmethod
    def forward(ctx, x, weight_dev):
        ctx.weight_dev = weight_dev
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = _chk(g, name="grad")
        out = torch.empty_like(g)
        _lib.call("dd_grl_backward_dev", _ptr(g), _ptr(ctx.weight_dev), _ptr(out), g.numel(), 0, _stream())
        return out, None


def gradient_scalar_dev(x, weight_dev):
    return _GradientScalarDev.apply(x, weight_dev)


def adv_grl_weight(loss, bce, lam, lam_adv, threshold, out=None):
    """Device-side AdvGRL weight from a device loss scalar (no host sync)."""
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=loss.device)
    _lib.call("dd_adv_grl_weight", _ptr(loss.detach().view(1)), float(bce), float(lam), float(lam_adv),
              float(threshold), _ptr(out), _stream())
    return out


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep):
        x = _chk(x, name="x")
        keep = _chk(keep, name="keep")
        out = torch.empty_like(x)
        _lib.call("dd_dropout_apply", _ptr(x), _ptr(keep), _ptr(out), x.numel(), _stream())
        ctx.save_for_backward(keep)
        return out

    @staticmethod
    def backward(ctx, g):
        (keep,) = ctx.saved_tensors
        g = _chk(g, name="grad")
        out = torch.empty_like(g)
        _lib.call("dd_dropout_apply", _ptr(g), _ptr(keep), _ptr(out), g.numel(), _stream())
        return out, None


def dropout_with_mask(x, keep):
    """F.dropout(p=0.5, training=True) with a caller-supplied {0,1} keep mask."""
    return _Dropout.apply(x, keep)


# ----------------------------------------------------------------------------------------- losses
class _ScaledGradLoss(torch.autograd.Function):
    """Common tail: forward computed (loss, grads) in one fused kernel; backward multiplies the stored
    gradients by the upstream scalar without reading it on the host."""

    @staticmethod
    def forward(ctx, loss, n_inputs, *tensors):
        inputs, grads = tensors[:n_inputs], tensors[n_inputs:]
        ctx.save_for_backward(*grads)
        ctx.n = n_inputs
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        outs = [None, None]
        for i, gr in enumerate(ctx.saved_tensors):
            outs.append(gr * g if ctx.needs_input_grad[2 + i] else None)
        outs.extend([None] * ctx.n)
        return tuple(outs)


def _finish(loss, inputs, grads):
    need = any(t.requires_grad for t in inputs)
    if not need:
        return loss.view(())
    return _ScaledGradLoss.apply(loss, len(inputs), *inputs, *grads)


def bce_with_logits_mean(x, targets=None, seg_labels=None, seg_len=0):
    """mean BCE-with-logits; targets float tensor of x's shape, or per-segment uint8 labels."""
    xc = _chk(x, name="logits").view(-1)
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(xc)
    t = _chk(targets, name="targets")
    sl = _chk(seg_labels, torch.uint8, "seg_labels")
    _lib.call("dd_bce_logits_mean", _ptr(xc), _ptr(t), _ptr(sl), int(seg_len), xc.numel(), _ptr(loss), _ptr(grad),
              _stream())
    return _finish(loss, (x,), (grad.view_as(x),))


def softmax_ce_mean(logits, labels, row_mask):
    lg = _chk(logits, name="logits")
    rows, c = lg.shape
    loss = torch.empty(1, dtype=torch.float32, device=lg.device)
    grad = torch.empty_like(lg)
    _lib.call("dd_softmax_ce_mean", _ptr(lg), _ptr(_chk(labels, torch.int64, "labels")),
              _ptr(_chk(row_mask, torch.uint8, "row_mask")), rows, c, _ptr(loss), _ptr(grad), _stream())
    return _finish(loss, (logits,), (grad,))


def box_reg_loss(box_reg, reg_targets, labels, row_mask):
    br = _chk(box_reg, name="box_regression")
    rows, c4 = br.shape
    loss = torch.empty(1, dtype=torch.float32, device=br.device)
    grad = torch.empty_like(br)
    _lib.call("dd_box_reg_loss", _ptr(br), _ptr(_chk(reg_targets, name="regression_targets")),
              _ptr(_chk(labels, torch.int64, "labels")), _ptr(_chk(row_mask, torch.uint8, "row_mask")), rows, c4 // 4,
              _ptr(loss), _ptr(grad), _stream())
    return _finish(loss, (box_reg,), (grad,))


def smooth_l1_sum(x, t, beta, divisor):
    xc = _chk(x, name="x")
    tc = _chk(t, name="t")
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(xc)
    _lib.call("dd_smooth_l1_sum", _ptr(xc), _ptr(tc), xc.numel(), float(beta), float(divisor), _ptr(loss), _ptr(grad),
              _stream())
    return _finish(loss, (x,), (grad,))


def consistency_loss(img_logits, ins_logits, n_src, row_valid=None):
    """img_logits [2, hw] (any trailing shape), ins_logits [K] or [K,1]; sigmoid folded into the kernel.
    row_valid (uint8 [K], optional): padding ROIs of the fixed-capacity layout are skipped."""
    il = _chk(img_logits, name="img_logits")
    sl = _chk(ins_logits, name="ins_logits")
    if il.shape[0] != 2:
        raise AssertionError("only batch size=2 is supported for consistency loss now, received batch size: {}".format(
            il.shape[0]))
    hw = il.numel() // 2
    k = sl.numel()
    loss = torch.empty(1, dtype=torch.float32, device=il.device)
    gi, gs = torch.empty_like(il), torch.empty_like(sl)
    ws = torch.empty(4, dtype=torch.float32, device=il.device)
    _lib.call("dd_consistency_loss", _ptr(il), hw, _ptr(sl), k, int(n_src), _ptr(_chk(row_valid, torch.uint8, "row_valid")),
              _ptr(loss), _ptr(gi), _ptr(gs), _ptr(ws), _stream())
    return _finish(loss, (img_logits, ins_logits), (gi, gs))


def adaptive_margin_update(state, prev_loss, margin_cfg, lr, max_margin, out=None):
    """The adaptive triplet margin (da_heads/loss.py:182-200) as device state: `state` double[1] (0 = uninitialised),
    `prev_loss` the previous step's triplet loss (device float[1] or None) -> float[1] margin for this step."""
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=state.device)
    _lib.call("dd_adaptive_margin_update", _ptr(_chk(state, torch.float64, "state")),
              _ptr(_chk(prev_loss, name="prev_loss")), float(margin_cfg), float(lr), float(max_margin), _ptr(out),
              _stream())
    return out


def triplet_margin_loss(a, p, n, margin, rows, d, inner):
    """nn.TripletMarginLoss(margin, p=2) with the distance over `d` elements strided by `inner`.
    margin: a Python float, or a device float[1] tensor (adaptive margin kept on the device)."""
    ac, pc, nc = _chk(a, name="anchor"), _chk(p, name="positive"), _chk(n, name="negative")
    loss = torch.empty(1, dtype=torch.float32, device=ac.device)
    need = a.requires_grad or p.requires_grad or n.requires_grad
    ga = torch.empty_like(ac) if need else None
    gp = torch.empty_like(pc) if need else None
    gn = torch.empty_like(nc) if need else None
    m_dev = _chk(margin, name="margin") if torch.is_tensor(margin) else None
    _lib.call("dd_triplet_margin_loss", _ptr(ac), _ptr(pc), _ptr(nc), int(rows), int(d), int(inner),
              0.0 if m_dev is not None else float(margin), _ptr(m_dev), _ptr(loss), _ptr(ga), _ptr(gp), _ptr(gn),
              _stream())
    if not need:
        return loss.view(())
    return _ScaledGradLoss.apply(loss, 3, a, p, n, ga, gp, gn)


# ----------------------------------------------------------------------------------------- detection ops
def anchor_grid(cell_anchors, fh, fw, stride, img_w, img_h, straddle):
    a = cell_anchors.shape[0]
    anchors = torch.empty((fh * fw * a, 4), dtype=torch.float32, device=cell_anchors.device)
    vis = torch.empty((fh * fw * a,), dtype=torch.uint8, device=cell_anchors.device)
    _lib.call("dd_anchor_grid", _ptr(_chk(cell_anchors, name="cell_anchors")), a, fh, fw, int(stride), int(img_w),
              int(img_h), int(straddle), _ptr(anchors), _ptr(vis), _stream())
    return anchors, vis


def rpn_topk_decode(logits, deltas, anchors, k, img_w, img_h, min_size):
    """logits [N,FH,FW,A], deltas [N,FH,FW,4A] NHWC -> boxes [N,k,4], scores [N,k], idx [N,k], valid [N]."""
    n, fh, fw, a = logits.shape
    dev = logits.device
    boxes = torch.empty((n, k, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((n, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    valid = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.call("dd_rpn_topk_decode", _ptr(_chk(logits, name="logits")), _ptr(_chk(deltas, name="deltas")),
              _ptr(_chk(anchors, name="anchors")), n, fh, fw, a, int(k), int(img_w), int(img_h), float(min_size),
              _ptr(boxes), _ptr(scores), _ptr(idx), _ptr(valid), None, _stream())
    return boxes, scores, idx, valid


def nms_sorted(boxes_sorted, thresh, max_keep=0):
    """boxes sorted by descending score -> (keep positions int64 [n] buffer, device count int32[1])."""
    b = _chk(boxes_sorted, name="boxes")
    n = b.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=b.device)
    count = torch.zeros((1,), dtype=torch.int32, device=b.device)
    ws = _workspace(_lib.load().dd_nms_workspace_bytes(n), b.device, "nms")
    _lib.call("dd_nms_sorted", _ptr(b), n, float(thresh), int(max_keep), _ptr(keep), _ptr(count), _ptr(ws), _stream())
    return keep, count


def nms_sorted_batched(boxes_sorted, valid, thresh, max_keep):
    """boxes [N, cap, 4] sorted by descending score per image, valid int32 [N] on the device ->
    (keep positions int64 [N, max_keep], counts int32 [N]); two launches for the whole batch, no host sync."""
    b = _chk(boxes_sorted, name="boxes")
    v = _chk(valid, torch.int32, "valid")
    n_img, cap = b.shape[0], b.shape[1]
    keep = torch.empty((n_img, max_keep), dtype=torch.int64, device=b.device)
    count = torch.empty((n_img,), dtype=torch.int32, device=b.device)
    ws = _workspace(_lib.load().dd_nms_batched_workspace_bytes(n_img, cap), b.device, "nms_batched")
    _lib.call("dd_nms_sorted_batched", _ptr(b), _ptr(v), n_img, cap, float(thresh), int(max_keep), _ptr(keep),
              int(max_keep), _ptr(count), _ptr(ws), _stream())
    return keep, count


def nms(boxes, scores, thresh):
    """`_C.nms` semantics (csrc/nms.h:10-28, GPU variant): kept original indices, ascending, int64."""
    b = _chk(boxes, name="boxes")
    s = _chk(scores, name="scores")
    n = b.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=b.device)
    keep = torch.empty((n,), dtype=torch.int64, device=b.device)
    count = torch.zeros((1,), dtype=torch.int32, device=b.device)
    ws = _workspace(_lib.load().dd_nms_workspace_bytes(n), b.device, "nms")
    _lib.call("dd_nms", _ptr(b), _ptr(s), n, float(thresh), _ptr(keep), _ptr(count), _ptr(ws), _stream())
    return keep[: int(count.item())]


def match(gt, pred, high, low, allow_low_quality, m_dev=None):
    """Fused boxlist_iou + Matcher: int64 [N] with -1 (below low) / -2 (between) sentinels.
    m_dev (device int32[1], optional): gt is padded to its row capacity and only the first *m_dev rows exist."""
    g = _chk(gt, name="gt")
    p = _chk(pred, name="pred")
    m, n = g.shape[0], p.shape[0]
    if m == 0:
        raise ValueError("No ground-truth boxes available for one of the images during training")
    if n == 0:
        raise ValueError("No proposal boxes available for one of the images during training")
    matches = torch.empty((n,), dtype=torch.int64, device=p.device)
    vals = torch.empty((n,), dtype=torch.float32, device=p.device)
    best = torch.empty((m,), dtype=torch.float32, device=p.device)
    _lib.call("dd_match", _ptr(g), m, _ptr(_chk(m_dev, torch.int32, "m_dev")), _ptr(p), n, float(high), float(low),
              1 if allow_low_quality else 0,
              _ptr(matches), _ptr(vals), _ptr(best), _stream())
    return matches, vals


def box_encode(gt, pred, matches, weights, wrap_negative=False, m_dev=None):
    g = _chk(gt, name="gt")
    p = _chk(pred, name="pred")
    out = torch.empty((p.shape[0], 4), dtype=torch.float32, device=p.device)
    wx, wy, ww, wh = (float(v) for v in weights)
    _lib.call("dd_box_encode", _ptr(g), g.shape[0], _ptr(_chk(m_dev, torch.int32, "m_dev")), _ptr(p),
              _ptr(_chk(matches, torch.int64, "matches")), p.shape[0],
              wx, wy, ww, wh, 1 if wrap_negative else 0, _ptr(out), _stream())
    return out


def rpn_anchor_labels(matches, visibility):
    """int32 [N] RPN labels (1 / 0 / -1 = ignored) from the Matcher result and the uint8 visibility mask."""
    m = _chk(matches, torch.int64, "matches")
    out = torch.empty(m.shape, dtype=torch.int32, device=m.device)
    _lib.call("dd_rpn_anchor_labels", _ptr(m), _ptr(_chk(visibility, torch.uint8, "visibility")), m.numel(), _ptr(out),
              _stream())
    return out


class _RpnSampledLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, deltas, anchors, sel, counts, labels, matches, gt_cat, gt_offsets, src_img, beta):
        lg = _chk(logits, name="logits")
        dl = _chk(deltas, name="deltas")
        s, b = sel.shape
        a = anchors.shape[0]
        losses = torch.empty(2, dtype=torch.float32, device=lg.device)
        g_lg, g_dl = torch.zeros_like(lg), torch.zeros_like(dl)
        _lib.call("dd_rpn_sampled_losses", _ptr(lg), _ptr(dl), _ptr(_chk(anchors, name="anchors")), a, s, b,
                  _ptr(_chk(sel, torch.int64, "sel")), _ptr(_chk(counts, torch.int32, "counts")),
                  _ptr(_chk(labels, torch.int32, "labels")), _ptr(_chk(matches, torch.int64, "matches")),
                  _ptr(_chk(gt_cat, name="gt")), _ptr(_chk(gt_offsets, torch.int32, "gt_offsets")),
                  _ptr(_chk(src_img, torch.int32, "src_img")), float(beta), _ptr(losses), _ptr(g_lg), _ptr(g_dl),
                  _stream())
        ctx.save_for_backward(g_lg, g_dl)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_obj, g_box):
        g_lg, g_dl = ctx.saved_tensors
        return (g_lg * g_obj if ctx.needs_input_grad[0] else None, g_dl * g_box if ctx.needs_input_grad[1] else None,
                None, None, None, None, None, None, None, None, None)


def rpn_sampled_losses(logits, deltas, anchors, sel, counts, labels, matches, gt_cat, gt_offsets, src_img, beta):
    """(objectness loss, box regression loss) of RPNLossComputation (rpn/loss.py:118-141) over the anchors sampled by
    balanced_sample for the S source images, forward + gradient in one launch.  logits [n, h, w, a] / deltas
    [n, h, w, 4a] (source images first); sel int64 [S, B]; counts int32 [S, 2]; labels int32 / matches int64 [S, A];
    gt_cat [G, 4] with gt_offsets int32 [n + 1]; src_img int32 [S]."""
    return _RpnSampledLosses.apply(logits, deltas, anchors, sel, counts, labels, matches, gt_cat, gt_offsets, src_img,
                                   beta)


def roi_labels(matches, gt_labels, is_source, n_prop, out=None):
    """int32 [cap] box-head labels of one image's proposal buffer (class / 0 / -1 ignored); n_prop: device int32 [1]."""
    m = _chk(matches, torch.int64, "matches")
    if out is None:
        out = torch.empty(m.shape, dtype=torch.int32, device=m.device)
    _lib.call("dd_roi_labels", _ptr(m), _ptr(_chk(gt_labels, torch.int64, "gt_labels")), 1 if is_source else 0,
              _ptr(_chk(n_prop, torch.int32, "n_prop")), m.numel(), _ptr(out), _stream())
    return out


def roi_gather_sampled(boxes, objectness, sel, counts, labels, matches, gt_cat, gt_offsets, gt_counts, is_source,
                       weights):
    """The sampled ROIs of a batch in one launch -> dict(rois [K,5], labels int64 [K], regression_targets [K,4],
    domain_labels bool [K], valid bool [K], objectness [K]) with K = images x batch."""
    n_img, cap = objectness.shape
    b = sel.shape[1]
    dev = boxes.device
    k = n_img * b
    rois = torch.empty((k, 5), dtype=torch.float32, device=dev)
    lab = torch.empty((k,), dtype=torch.int64, device=dev)
    reg = torch.empty((k, 4), dtype=torch.float32, device=dev)
    dom = torch.empty((k,), dtype=torch.uint8, device=dev)
    val = torch.empty((k,), dtype=torch.uint8, device=dev)
    obj = torch.empty((k,), dtype=torch.float32, device=dev)
    wx, wy, ww, wh = (float(v) for v in weights)
    _lib.call("dd_roi_gather_sampled", _ptr(_chk(boxes, name="boxes")), _ptr(_chk(objectness, name="objectness")),
              _ptr(_chk(sel, torch.int64, "sel")), _ptr(_chk(counts, torch.int32, "counts")),
              _ptr(_chk(labels, torch.int32, "labels")), _ptr(_chk(matches, torch.int64, "matches")),
              _ptr(_chk(gt_cat, name="gt")), _ptr(_chk(gt_offsets, torch.int32, "gt_offsets")),
              _ptr(_chk(gt_counts, torch.int32, "gt_counts")), _ptr(_chk(is_source, torch.uint8, "is_source")),
              n_img, cap, b, wx, wy, ww, wh, _ptr(rois), _ptr(lab), _ptr(reg), _ptr(dom), _ptr(val), _ptr(obj), _stream())
    return dict(rois=rois, labels=lab, regression_targets=reg, domain_labels=dom.view(torch.bool),
                valid=val.view(torch.bool), objectness=obj)


def box_decode(codes, boxes, weights):
    c = _chk(codes, name="codes")
    b = _chk(boxes, name="boxes")
    r, k4 = c.shape
    out = torch.empty_like(c)
    wx, wy, ww, wh = (float(v) for v in weights)
    _lib.call("dd_box_decode", _ptr(c), _ptr(b), r, k4 // 4, wx, wy, ww, wh, _ptr(out), _stream())
    return out


def proposals_gather(boxes, scores, keep, keep_count, gt_cat, gt_offsets, append_gt, cap, gt_counts=None):
    """boxes [N,k,4], scores [N,k], keep int64 [N,post], keep_count int32 [N], gt_cat [G,4], gt_offsets int32 [N+1],
    append_gt uint8 [N] -> (proposals [N,cap,4], objectness [N,cap], count int32 [N]); no host read.
    gt_counts (device int32 [N], optional): live GT rows per image when the GT boxes are padded to a capacity."""
    n, k = scores.shape
    post = keep.shape[1]
    dev = boxes.device
    out_b = torch.empty((n, cap, 4), dtype=torch.float32, device=dev)
    out_s = torch.empty((n, cap), dtype=torch.float32, device=dev)
    out_c = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.call("dd_proposals_gather", _ptr(_chk(boxes, name="boxes")), _ptr(_chk(scores, name="scores")),
              _ptr(_chk(keep, torch.int64, "keep")), _ptr(_chk(keep_count, torch.int32, "keep_count")),
              _ptr(_chk(gt_cat, name="gt")), _ptr(_chk(gt_offsets, torch.int32, "gt_offsets")),
              _ptr(_chk(gt_counts, torch.int32, "gt_counts")),
              _ptr(_chk(append_gt, torch.uint8, "append_gt")), n, k, post, int(cap), _ptr(out_b), _ptr(out_s),
              _ptr(out_c), _stream())
    return out_b, out_s, out_c


def balanced_sample(labels, n_dev, keys, batch, max_pos):
    """labels int32 [images, n_cap], n_dev int32 [images] or None, keys float [images, n_cap] ->
    (sel_idx int64 [images, batch] ascending, counts int32 [images, 2] = {positives, total})."""
    lab = _chk(labels, torch.int32, "labels")
    images, n_cap = lab.shape
    sel = torch.empty((images, batch), dtype=torch.int64, device=lab.device)
    cnt = torch.empty((images, 2), dtype=torch.int32, device=lab.device)
    _lib.call("dd_balanced_sample", _ptr(lab), _ptr(_chk(n_dev, torch.int32, "n_dev")), _ptr(_chk(keys, name="keys")),
              images, n_cap, int(batch), int(max_pos), _ptr(sel), _ptr(cnt), _stream())
    return sel, cnt


def sgd_momentum_dev_(p, g, buf, lr_dev, lr_factor, momentum, wd, grad_scale):
    _lib.call("dd_sgd_momentum_dev", _ptr(p), _ptr(g), _ptr(buf), p.numel(), _ptr(lr_dev), float(lr_factor),
              float(momentum), float(wd), float(grad_scale), _stream())


def sgd_momentum_(p, g, buf, lr, momentum, wd, grad_scale, first_step):
    _lib.call("dd_sgd_momentum", _ptr(p), _ptr(g), _ptr(buf), p.numel(), float(lr), float(momentum), float(wd),
              float(grad_scale), 1 if first_step else 0, _stream())
