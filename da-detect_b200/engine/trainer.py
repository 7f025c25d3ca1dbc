"""One training iteration of the reference's do_da_train loop (engine/trainer.py:196-242) with the
B200-native optimiser/communication tail:

  loss_dict = model(images, targets); losses = sum(loss_dict.values())
  optimizer.zero_grad(); losses.backward(); [DDP all-reduce]; optimizer.step()

Differences from the reference (SURVEY §5 "Distributed communication backend", §8f-1):
  * all trainable gradients live in ONE flat fp32 buffer (each p.grad is a view), so the data-parallel
    exchange is a single `dist.all_reduce` (NCCL over NVLink/NVSwitch) per step instead of DDP's bucket
    sequence; the 1/world average is folded into the SGD kernel's grad_scale;
  * parameters are views of one flat buffer as well, ordered [weights..., biases...], so the whole SGD
    update (momentum 0.9, weight decay on weights only, bias lr x BIAS_LR_FACTOR — solver/build.py:7-20)
    is two launches of one fused kernel instead of ~170 per-tensor param-group updates;
  * parameters no loss of the active configuration depends on (e.g. `da_heads.*` in triplet mode, a DA head
    whose loss weights are all 0) stay OUTSIDE the flat buffers and are never touched: under the reference's
    autograd their .grad stays None and torch.optim.SGD skips them — no weight decay, no momentum
    (`unused_parameter_names`).
"""
import math
import os
from bisect import bisect_right

import torch
import torch.distributed as dist

from .. import ops
from ..utils.sections import section


class WarmupMultiStepLR(object):
    """solver/lr_scheduler.py:10-52 (host-side scalar schedule)."""

    def __init__(self, base_lr, milestones, gamma=0.1, warmup_factor=1.0 / 3, warmup_iters=500,
                 warmup_method="linear"):
        if list(milestones) != sorted(milestones):
            raise ValueError("Milestones should be a list of increasing integers. Got {}".format(milestones))
        if warmup_method not in ("constant", "linear"):
            raise ValueError("Only 'constant' or 'linear' warmup_method accepted got {}".format(warmup_method))
        self.base_lr, self.milestones, self.gamma = base_lr, list(milestones), gamma
        self.warmup_factor, self.warmup_iters, self.warmup_method = warmup_factor, warmup_iters, warmup_method

    def lr_at(self, it):
        f = 1.0
        if it < self.warmup_iters:
            if self.warmup_method == "constant":
                f = self.warmup_factor
            else:
                alpha = float(it) / self.warmup_iters
                f = self.warmup_factor * (1 - alpha) + alpha
        return self.base_lr * f * self.gamma ** bisect_right(self.milestones, it)


class WarmupCosineLR(object):
    """The schedule tools/train_net_triplet.py:67-81 builds with timm's CosineLRScheduler
    (t_initial = MAX_ITER, lr_min, linear warm-up from warmup_lr_init over warmup_t updates,
    cycle_limit 1, t_in_epochs False)."""

    def __init__(self, base_lr, max_iter, lr_min, warmup_lr, warmup_iters):
        self.base_lr, self.max_iter, self.lr_min = base_lr, max_iter, lr_min
        self.warmup_lr, self.warmup_iters = warmup_lr, warmup_iters

    def lr_at(self, it):
        if it < self.warmup_iters:
            return self.warmup_lr + it * (self.base_lr - self.warmup_lr) / self.warmup_iters
        if it >= self.max_iter:
            return self.lr_min
        return self.lr_min + 0.5 * (self.base_lr - self.lr_min) * (1 + math.cos(math.pi * it / self.max_iter))


def unused_parameter_names(model):
    """Names of the trainable parameters that no loss term of the model's configuration depends on.  The reference
    builds `da_heads` even in triplet mode (generalized_rcnn.py:53) and evaluates heads whose losses are then
    dropped for a zero weight (da_heads.py:417-436): those parameters never receive a gradient, so SGD never
    updates them (torch.optim.SGD skips p.grad is None)."""
    unused = set()

    def heads(mod, prefix):
        img = mod.img_weight > 0 or mod.cst_weight > 0
        ins = mod.ins_weight > 0 or mod.cst_weight > 0
        for n, _ in mod.named_parameters():
            if (n.startswith("imghead.") and not img) or (n.startswith("inshead.") and not ins):
                unused.add(prefix + n)

    da, tri = getattr(model, "da_heads", None), getattr(model, "da_heads_triplet", None)
    if da:
        if tri:
            unused.update("da_heads." + n for n, _ in da.named_parameters())
        else:
            heads(da, "da_heads.")
    if tri:
        heads(tri, "da_heads_triplet.")
    return unused


class FlatSGDTrainer(object):
    gt_capacity = 128            # GT boxes per image the step graphs are captured for (grows by doubling)

    def __init__(self, model, cfg, schedule=None, world_size=None):
        S = cfg.SOLVER
        self.model = model
        self.momentum = S.MOMENTUM
        self.base_lr, self.bias_lr_factor = S.BASE_LR, S.BIAS_LR_FACTOR
        self.wd, self.wd_bias = S.WEIGHT_DECAY, S.WEIGHT_DECAY_BIAS
        self.schedule = schedule
        self.iteration = 0
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.unused = unused_parameter_names(model)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n not in self.unused]
        weights = [(n, p) for n, p in named if "bias" not in n]      # solver/build.py:14: `"bias" in key`
        biases = [(n, p) for n, p in named if "bias" in n]
        self.order = weights + biases
        pad4 = lambda n: (n + 3) // 4 * 4                      # every segment starts 16-byte aligned
        self.n_weight = sum(pad4(p.numel()) for _, p in weights)
        total = sum(pad4(p.numel()) for _, p in self.order)
        dev = self.order[0][1].device
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_buf = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for _, p in self.order:
                n = p.numel()
                seg_p, seg_g = self.flat_param[off:off + n], self.flat_grad[off:off + n]
                if p.dim() == 4:       # keep the physical OHWI (channels_last) layout inside the flat buffers
                    co, ci, kh, kw = p.shape
                    vp = seg_p.view(co, kh, kw, ci).permute(0, 3, 1, 2)
                    vg = seg_g.view(co, kh, kw, ci).permute(0, 3, 1, 2)
                else:
                    vp, vg = seg_p.view_as(p), seg_g.view_as(p)
                vp.copy_(p)
                p.data = vp
                p.grad = vg
                off += pad4(n)
        self.total = total
        # Gradient exchange in reverse-order segments (world > 1): backward finishes the heads (RPN, box head, DA
        # heads) first, then res4, then res3; each segment of the flat buffer is all-reduced on a communication
        # stream as soon as its last weight-gradient kernel has been launched (ops.grad_milestone), overlapping the
        # exchange with the rest of backward.  Segments: [heads | backbone.body.layer3 | the rest + all biases].
        offs, o = {}, 0
        for n, p in self.order:
            offs[n] = o
            o += pad4(p.numel())
        wnames = [n for n, _ in weights]
        first_head = next((offs[n] for n in wnames if not n.startswith("backbone.")), self.n_weight)
        l3 = [offs[n] for n in wnames if n.startswith("backbone.body.layer3.")]
        l3_begin = min(l3) if l3 else first_head
        self.segments = {"out:body": (first_head, self.n_weight), "in:layer3": (l3_begin, first_head)}
        self.tail_segments = [(0, l3_begin), (self.n_weight, total)]
        # Off by default: measured on 8 x B200 (profiles/r02_scale_8gpu.txt) the overlapped exchange is SLOWER than one
        # all_reduce after backward (17.98 vs 17.79 ms/step).  NCCL's NVLS kernels occupy 24 SMs while they run, and the
        # dense kernels are persistent one-CTA-per-SM grids that need 227 KB of shared memory per SM: every kernel that
        # overlaps the exchange gets a second wave on the blocked SMs.  DD_OVERLAP_EXCHANGE=1 switches it on.
        self.overlap_exchange = os.environ.get("DD_OVERLAP_EXCHANGE", "0") == "1"
        # SMs left to the collective's kernels while a segment exchange may be running beside backward: the persistent
        # dense kernels launched from the first milestone on are limited to the rest (ops.sm_budget), so that none of
        # their CTAs queues behind NCCL's for an SM.  Pair it with NCCL_MAX_CTAS of the same value (bench.py does).
        self.exchange_ctas = int(os.environ.get("DD_EXCHANGE_CTAS", "0"))
        self._budget_prev = None
        self.comm_stream = None
        self._reduced = set()
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_graphs = None          # signature -> captured whole-step CUDA graph (enable_step_graph)
        self.graph_launches = 0          # kernels of ours replayed through step graphs so far
        self.graph_stream = None
        self.graph_pool = None           # one memory pool shared by all step graphs (their replays never overlap)
        self.max_step_graphs = 16        # signatures kept; the least recently used graph is dropped beyond that
        self.early_backward = False
        # the weight operands of the whole step (3xTF32 hi / lo planes, dgrad transposes) prepared by a few batched
        # launches at the start of the step instead of one small launch in front of every conv (ops.WeightPrepPlan)
        self.batch_weight_prep = False
        self._prep_plan = None
        # weight-gradient kernels of the main backward pass on a side stream with this many SMs (0 = off), the fused
        # stages' data-gradient kernels on the rest (ops.set_wgrad_side)
        self.wgrad_side_sms = 0
        self._wgrad_stream = None

    def enable_step_graph(self, flag=True):
        """Capture zero_grad + forward + backward + all-reduce + SGD of one iteration into ONE CUDA graph per
        batch signature (image shape, GT boxes per image, source flags) and replay it afterwards: the step has
        no host reads (model.enable_static_shapes), the learning rate is read from device memory, and new
        batches are copied into the graph's static input buffers."""
        self.step_graphs = {} if flag else None
        ops.set_direct_weight_grad(flag)     # wgrad kernels accumulate into the flat gradient buffer directly
        if flag:
            self.model.enable_static_shapes(True)
            self.model.enable_cuda_graphs(False)
            self.enable_early_backward(os.environ.get("DD_EARLY_BACKWARD", "1") != "0")
            self.batch_weight_prep = os.environ.get("DD_BATCH_WEIGHT_PREP", "1") != "0"
            # measured on configs[1] (16.47 ms without): 48 SMs 16.60, 64 SMs 16.08, 74 SMs 16.24, 96 SMs 17.64
            self.wgrad_side_sms = int(os.environ.get("DD_WGRAD_SIDE_SMS", "64"))

    def enable_early_backward(self, flag=True):
        """Back-propagate the RPN losses during the forward pass, beside the latency-bound proposal chain
        (modeling/rpn.py::_forward_static_early).  The gradients are then zeroed BEFORE the forward pass; the loss
        weights are 1 (`losses = sum(loss_dict.values())`, engine/trainer.py:228-231)."""
        self.early_backward = bool(flag)
        self.model.early_backward = self.early_backward

    def zero_grad(self):
        self.flat_grad.zero_()

    def lr(self):
        return self.schedule.lr_at(self.iteration) if self.schedule is not None else self.base_lr

    def _exchange(self, a, b, side):
        """all_reduce of flat_grad[a:b]; side=True: on the communication stream, ordered after everything launched so
        far on the current stream (inside a step-graph capture the fork and the join are graph edges)."""
        if b <= a:
            return
        seg = self.flat_grad[a:b]
        if side and seg.is_cuda:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(device=seg.device)
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(seg)
        else:
            dist.all_reduce(seg)

    def _on_milestone(self, tag):
        rng = self.segments.get(tag)
        if rng is not None and tag not in self._reduced:
            self._reduced.add(tag)
            self._exchange(rng[0], rng[1], side=True)
            if self.exchange_ctas > 0 and self._budget_prev is None and self.flat_grad.is_cuda:
                from .. import _lib
                self._budget_prev = _lib.load().dd_set_sm_budget(ops.NUM_SMS - self.exchange_ctas)

    def begin_backward(self):
        """Arm the overlapped exchange for the backward pass that follows (no-op for one rank)."""
        self._reduced = set()
        if self.world > 1 and self.overlap_exchange:
            ops.set_grad_milestone_callback(self._on_milestone)

    def all_reduce(self):
        """Finish the gradient exchange: whatever segment was not reduced during backward, then join."""
        if self.world <= 1:
            return
        ops.set_grad_milestone_callback(None)
        if self._budget_prev is not None:
            from .. import _lib
            _lib.load().dd_set_sm_budget(self._budget_prev)
            self._budget_prev = None
        if not self._reduced:
            dist.all_reduce(self.flat_grad)              # one call over the whole buffer (overlap off / no milestones)
            return
        for tag, (a, b) in self.segments.items():
            if tag not in self._reduced:
                self._exchange(a, b, side=False)
        for a, b in self.tail_segments:
            self._exchange(a, b, side=False)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._reduced = set()

    def optimizer_step(self):
        lr = self.lr()
        first = self.iteration == 0
        scale = 1.0 / self.world
        nw = self.n_weight
        ops.sgd_momentum_(self.flat_param[:nw], self.flat_grad[:nw], self.flat_buf[:nw], lr, self.momentum, self.wd,
                          scale, first)
        if self.total > nw:
            ops.sgd_momentum_(self.flat_param[nw:], self.flat_grad[nw:], self.flat_buf[nw:], lr * self.bias_lr_factor,
                              self.momentum, self.wd_bias, scale, first)
        self.iteration += 1

    def _optimizer_step_dev(self):
        scale = 1.0 / self.world
        nw = self.n_weight
        ops.sgd_momentum_dev_(self.flat_param[:nw], self.flat_grad[:nw], self.flat_buf[:nw], self.lr_dev, 1.0,
                              self.momentum, self.wd, scale)
        if self.total > nw:
            ops.sgd_momentum_dev_(self.flat_param[nw:], self.flat_grad[nw:], self.flat_buf[nw:], self.lr_dev,
                                  self.bias_lr_factor, self.momentum, self.wd_bias, scale)

    def _graph_step(self, images, targets):
        from .. import _lib
        from ..structures import BoxList, cache_source_flags, is_source_image
        from ..structures.image_list import ImageList
        tensors = images.tensors if isinstance(images, ImageList) else images
        if not torch.is_tensor(tensors) or not self.model.roi_heads:
            return None
        if not getattr(self.model, "static_shapes", False):
            return None                  # host-driven control flow (FPN, enable_static_shapes(False)) reads sizes
        # the un-padded (h, w) of every image: anchor visibility and proposal clipping depend on them
        sizes = tuple((int(h), int(w)) for h, w in (images.image_sizes if isinstance(images, ImageList)
                                                    else [tensors.shape[-2:]] * tensors.shape[0]))
        cache_source_flags(targets)
        counts = [len(t) for t in targets]
        if min(counts) == 0:             # matcher.py:53-62 raises on an image without ground truth
            raise ValueError("No ground-truth boxes available for one of the images during training")
        # GT boxes live in fixed-capacity buffers with device-side counts: ONE graph per image shape serves every
        # batch, whatever the number of boxes per image (real batches differ in it almost every step)
        cap = int(self.gt_capacity)
        while cap < max(counts):
            cap *= 2
        key = (tuple(tensors.shape), sizes, cap) + tuple((bool(is_source_image(t)), tuple(t.size)) for t in targets)
        ent = self.step_graphs.pop(key, None)
        if ent is not None:
            self.step_graphs[key] = ent          # most recently used last
        if ent is None:
            while len(self.step_graphs) >= self.max_step_graphs:
                self.step_graphs.pop(next(iter(self.step_graphs)))
            dev = tensors.device
            counts_dev = torch.zeros(len(targets), dtype=torch.int32, device=dev)
            st_targets = []
            for i, t in enumerate(targets):
                b = BoxList(torch.zeros((cap, 4), dtype=torch.float32, device=dev), t.size, mode="xyxy")
                b.add_field("labels", torch.zeros((cap,), dtype=t.get_field("labels").dtype, device=dev))
                b._is_source_image = bool(is_source_image(t))
                b._gt_count_dev = counts_dev[i:i + 1]
                st_targets.append(b)
            ent = dict(images=torch.empty_like(tensors), sizes=[torch.Size(s) for s in sizes], targets=st_targets,
                       counts=counts_dev, graph=None, loss_keys=None, loss_vec=None, calls=0, launches=0)
            self.step_graphs[key] = ent
        ent["images"].copy_(tensors, non_blocking=True)
        for st, t, n in zip(ent["targets"], targets, counts):
            st.bbox[:n].copy_(t.convert("xyxy").bbox, non_blocking=True)
            st.get_field("labels")[:n].copy_(t.get_field("labels"), non_blocking=True)
        ent["counts"].copy_(torch.tensor(counts, dtype=torch.int32), non_blocking=True)
        self.lr_dev.fill_(self.lr())
        ent["calls"] += 1
        batch = ImageList(ent["images"], ent["sizes"])
        # Eager warm-up and capture run on ONE dedicated stream, and nothing returned keeps the autograd graph
        # alive: a stale AccumulateGrad node bound to another stream would invalidate the capture.
        if self.graph_stream is None:
            # high priority: the step's main stream carries its critical path (proposal chain, box head); the dense
            # kernels of the early backward passes on the side streams must not queue ahead of it for the SMs
            self.graph_stream = torch.cuda.Stream(priority=int(os.environ.get("DD_MAIN_PRIORITY", "-1")))
        cur = torch.cuda.current_stream()
        if ent["calls"] == 1:
            # first sight of a signature: one eager step (lazy workspaces, cached constants, kernel attributes)
            self.graph_stream.wait_stream(cur)
            with torch.cuda.stream(self.graph_stream):
                ld = self._eager_step(batch, ent["targets"], dev_lr=True)
                loss_dict = {k: v.detach().clone() for k, v in ld.items()}
                del ld
            cur.wait_stream(self.graph_stream)
        else:
            if ent["graph"] is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                if self.graph_pool is None:
                    self.graph_pool = torch.cuda.graph_pool_handle()
                before = _lib.launch_count()
                with torch.cuda.graph(g, pool=self.graph_pool, stream=self.graph_stream):
                    ld = self._eager_step(batch, ent["targets"], dev_lr=True)
                    ent["loss_keys"] = list(ld.keys())
                    ent["loss_vec"] = torch.stack([v.detach().reshape(()) for v in ld.values()])
                    del ld
                ent["launches"] = _lib.launch_count() - before
                ent["graph"] = g
            ent["graph"].replay()
            self.graph_launches += ent["launches"]
            # a copy: the graph's own output buffer is overwritten by the next replay of this signature
            vec = ent["loss_vec"].clone()
            loss_dict = {k: vec[i] for i, k in enumerate(ent["loss_keys"])}
        self.iteration += 1
        return loss_dict

    def _eager_step(self, images, targets, dev_lr=False):
        if self.batch_weight_prep and self.flat_param.is_cuda:
            if self._prep_plan is None:       # first step: record which weights the dense tier prepares
                stable = [t.untyped_storage().data_ptr() for t in list(self.model.parameters()) + list(self.model.buffers())]
                self._prep_plan = ops.WeightPrepPlan(stable + [self.flat_param.untyped_storage().data_ptr()])
            with ops.weight_prep(self._prep_plan):
                return self._eager_step_body(images, targets, dev_lr)
        return self._eager_step_body(images, targets, dev_lr)

    def _eager_step_body(self, images, targets, dev_lr=False):
        self.begin_backward()                 # (the milestone hooks are registered during the forward pass)
        if self.early_backward:
            self.zero_grad()
        loss_dict = self.model(images, targets)
        losses = sum(loss_dict.values())
        if not self.early_backward:
            self.zero_grad()
        if self.wgrad_side_sms > 0 and self.flat_grad.is_cuda and ops._direct_wgrad:
            if self._wgrad_stream is None:
                self._wgrad_stream = torch.cuda.Stream(device=self.flat_grad.device)
            ops.set_wgrad_side(torch.cuda.current_stream(), self._wgrad_stream, self.wgrad_side_sms)
        try:
            losses.backward()
            ops.wgrad_side_join()
        finally:
            ops.set_wgrad_side(None, None, 0)
        self.all_reduce()
        if dev_lr:
            self._optimizer_step_dev()
        else:
            self.optimizer_step()
        return loss_dict

    def step(self, images, targets):
        """images: ImageList/tensor on the device; returns the (unreduced) loss dict of this rank."""
        if self.step_graphs is not None and self.model.training:
            out = self._graph_step(images, targets)
            if out is not None:
                return out
        with section("forward"):
            self.begin_backward()
            if self.early_backward:
                self.zero_grad()
            loss_dict = self.model(images, targets)
            losses = sum(loss_dict.values())
        with section("backward"):
            if not self.early_backward:
                self.zero_grad()
            losses.backward()
        with section("allreduce+sgd"):
            self.all_reduce()
            self.optimizer_step()
        return loss_dict
