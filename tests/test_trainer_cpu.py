"""CPU-only: host-side logic of the data-parallel tail (engine/trainer.py) with world_size 2 over gloo —
flat parameter/gradient buffers, [weights..., biases...] ordering, one all_reduce of the flat gradient whose
1/world average is applied by the optimiser's grad_scale — and the LR schedules."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(4, 8, 3, bias=True)
        self.fc = nn.Linear(8, 3)
        self.frozen = nn.Linear(2, 2)
        for p in self.frozen.parameters():
            p.requires_grad = False


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dadetect_b200.config import get_cfg_defaults
    from dadetect_b200.engine import FlatSGDTrainer
    torch.manual_seed(0)
    model = _Toy()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    tr = FlatSGDTrainer(model, get_cfg_defaults())
    assert tr.world == world
    # parameters are views of the flat buffer and kept their values; weights precede biases
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    names = [n for n, _ in tr.order]
    assert names == ["conv.weight", "fc.weight", "conv.bias", "fc.bias"]
    assert tr.n_weight == 8 * 4 * 9 + 3 * 8
    off = 0
    for n, p in tr.order:
        assert p.data_ptr() == tr.flat_param.data_ptr() + 4 * off, n
        assert p.grad.data_ptr() == tr.flat_grad.data_ptr() + 4 * off, n
        off += (p.numel() + 3) // 4 * 4
    # rank-dependent gradients written through p.grad land in the flat buffer; one all_reduce sums them
    tr.zero_grad()
    x = torch.randn(2, 4, 5, 5, generator=torch.Generator().manual_seed(100 + rank))
    y = model.fc(model.conv(x).mean(dim=(2, 3))).sum()
    y.backward()
    local = tr.flat_grad.clone()
    tr.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(tr.flat_grad, sum(gathered), atol=1e-6)
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_flat_buffers_and_all_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get() == "ok"


def test_lr_schedules():
    from dadetect_b200.engine import WarmupCosineLR, WarmupMultiStepLR
    s = WarmupMultiStepLR(0.001, (50000,), 0.1, 1.0 / 3, 500, "linear")
    assert s.lr_at(0) == pytest.approx(0.001 / 3)
    assert s.lr_at(250) == pytest.approx(0.001 * (1 / 3 * 0.5 + 0.5))
    assert s.lr_at(500) == pytest.approx(0.001)
    assert s.lr_at(50000) == pytest.approx(0.0001)
    with pytest.raises(ValueError):
        WarmupMultiStepLR(0.001, (5, 3))
    c = WarmupCosineLR(0.001, 170000, 1e-6, 1e-4, 33200)
    assert c.lr_at(0) == pytest.approx(1e-4)
    assert c.lr_at(33200) < 0.001 and c.lr_at(33200) > 0.0009
    assert c.lr_at(170000) == pytest.approx(1e-6)


@pytest.mark.parametrize("name", ["da_img_only", "da_img_ins_cst", "triplet_aligned_advgrl", "triplet_yaml_default"])
def test_parameters_outside_the_flat_buffers_are_the_ones_the_reference_never_updates(name):
    """torch.optim.SGD skips parameters whose .grad is None (no weight decay, no momentum).  The set the trainer
    leaves outside its flat buffers must be exactly the set the REAL reference left without a gradient
    (tests/golden/scenario_*.pt `params_without_grad`, recorded from the reference's own backward)."""
    from dadetect_b200.config import get_cfg_defaults
    from dadetect_b200.engine.trainer import unused_parameter_names
    from dadetect_b200.modeling import build_detection_model
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fx = torch.load(os.path.join(root, "tests", "golden", "scenario_{}.pt".format(name)), weights_only=False)
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(root, "configs", fx["yaml"]))
    cfg.merge_from_list(list(fx["opts"]))
    model = build_detection_model(cfg)
    assert sorted(unused_parameter_names(model)) == sorted(fx["params_without_grad"])


class _ToyBody(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer2 = nn.Linear(6, 6)
        self.layer3 = nn.Linear(6, 6)


class _ToyBackbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = _ToyBody()


class _ToyDetector(nn.Module):
    """Parameter names shaped like the detector's (backbone.body.layer2/3, heads after the backbone) and the same
    milestone marks as ResNetC4.forward, so that the trainer's segment bookkeeping is exercised on the CPU."""

    def __init__(self):
        super().__init__()
        self.backbone = _ToyBackbone()
        self.rpn = nn.Linear(6, 4)
        self.roi_heads = nn.Linear(6, 3)

    def forward(self, x):
        from dadetect_b200 import ops
        x = torch.relu(self.backbone.body.layer2(x))
        x = ops.grad_milestone(x, "in:layer3")
        x = torch.relu(self.backbone.body.layer3(x))
        x = ops.grad_milestone(x, "out:body")
        return self.rpn(x).sum() + self.roi_heads(x).pow(2).sum()


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dadetect_b200.config import get_cfg_defaults
    from dadetect_b200.engine import FlatSGDTrainer
    torch.manual_seed(0)
    model = _ToyDetector()
    tr = FlatSGDTrainer(model, get_cfg_defaults())
    names = [n for n, _ in tr.order]
    assert names[:4] == ["backbone.body.layer2.weight", "backbone.body.layer3.weight", "rpn.weight", "roi_heads.weight"]
    assert tr.segments == {"out:body": (72, 72 + 24 + 20), "in:layer3": (36, 72)}, tr.segments
    assert tr.tail_segments == [(0, 36), (tr.n_weight, tr.total)]
    x = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 + rank))
    results = []
    for overlap in (True, False):
        tr.overlap_exchange = overlap
        seen = []
        orig = tr._exchange
        tr._exchange = lambda a, b, side, _o=orig, _s=seen: (_s.append((a, b, side)), _o(a, b, side))[1]
        tr.begin_backward()
        loss = model(x)
        tr.zero_grad()
        loss.backward()
        tr.all_reduce()
        tr._exchange = orig
        results.append(tr.flat_grad.clone())
        if overlap:      # heads first, then layer3 (both during backward), then the rest and the biases
            assert seen == [(72, 116, True), (36, 72, True), (0, 36, False), (tr.n_weight, tr.total, False)], seen
        else:
            assert seen == []
    assert torch.equal(results[0], results[1])
    # and it is the sum over ranks of the local gradients
    tr.overlap_exchange = False
    loss = model(x)
    tr.zero_grad()
    loss.backward()
    local = tr.flat_grad.clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(results[0], sum(gathered), atol=1e-6)
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_overlapped_segment_exchange_world2():
    """The gradient exchange in reverse-order segments started from backward milestones (heads, then layer3, then
    the rest + biases) gives exactly the single all_reduce of the whole flat buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get() == "ok"
