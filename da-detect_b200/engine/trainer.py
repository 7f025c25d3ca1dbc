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
  * parameters that receive no gradient in a configuration (e.g. `da_heads.*` in triplet mode) keep a
    zero slot, which is what DDP(find_unused_parameters) semantics would produce.
"""
import math
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


class FlatSGDTrainer(object):
    def __init__(self, model, cfg, schedule=None, world_size=None):
        S = cfg.SOLVER
        self.model = model
        self.momentum = S.MOMENTUM
        self.base_lr, self.bias_lr_factor = S.BASE_LR, S.BIAS_LR_FACTOR
        self.wd, self.wd_bias = S.WEIGHT_DECAY, S.WEIGHT_DECAY_BIAS
        self.schedule = schedule
        self.iteration = 0
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
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

    def zero_grad(self):
        self.flat_grad.zero_()

    def lr(self):
        return self.schedule.lr_at(self.iteration) if self.schedule is not None else self.base_lr

    def all_reduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad)

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

    def step(self, images, targets):
        """images: ImageList/tensor on the device; returns the (unreduced) loss dict of this rank."""
        with section("forward"):
            loss_dict = self.model(images, targets)
            losses = sum(loss_dict.values())
        with section("backward"):
            self.zero_grad()
            losses.backward()
        with section("allreduce+sgd"):
            self.all_reduce()
            self.optimizer_step()
        return loss_dict
