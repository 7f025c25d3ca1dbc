"""Random draws on the training path, made injectable so that a parity test can replay the exact
draws of a CPU oracle run: ``randperm(n)`` (reference balanced_positive_negative_sampler.py:57-58)
and ``dropout_keep(shape)`` (da_heads.py:63,65).  The default draws on the device with torch's
CUDA generator, like the reference does on GPU."""
import torch


class RandomSource(object):
    def randperm(self, n, device):
        return torch.randperm(n, device=device)

    def dropout_keep(self, shape, device):
        return torch.empty(shape, dtype=torch.float32, device=device).bernoulli_(0.5)


class ReplaySource(RandomSource):
    """Replays recorded CPU draws (lists of tensors) in order."""

    def __init__(self, perms, masks):
        self.perms, self.masks = list(perms), list(masks)

    def randperm(self, n, device):
        p = self.perms.pop(0)
        assert p.numel() == n, "randperm replay out of step: want {} have {}".format(n, p.numel())
        return p.to(device)

    def dropout_keep(self, shape, device):
        m = self.masks.pop(0)
        assert tuple(m.shape) == tuple(shape), (tuple(m.shape), tuple(shape))
        return m.to(device=device, dtype=torch.float32).contiguous()
