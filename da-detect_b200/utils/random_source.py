"""Random draws on the training path, made injectable so that a parity test can replay the exact
draws of a CPU oracle run: ``randperm(n)`` (reference balanced_positive_negative_sampler.py:57-58)
and ``dropout_keep(shape)`` (da_heads.py:63,65).  The default draws on the device with torch's
CUDA generator, like the reference does on GPU."""
import torch


class RandomSource(object):
    replay = False

    def randperm(self, n, device):
        return torch.randperm(n, device=device)

    def sample_keys(self, labels):
        """Random keys for the device-side fg/bg sampler (ops.balanced_sample): one uniform draw per candidate,
        the smallest keys of each class are selected — the distribution of `positive[randperm(n)[:k]]`
        (balanced_positive_negative_sampler.py:57-63) with a fixed-shape draw and no host read."""
        return torch.rand(labels.shape, dtype=torch.float32, device=labels.device)

    def consume_da_draws(self, counts):
        """subsample_for_da (box_head/loss.py:132-163) draws two more permutations per image whose result is
        not used (it selects every already-sampled proposal); only a replay has to account for them."""

    def dropout_keep(self, shape, device, row_valid=None):
        return torch.empty(shape, dtype=torch.float32, device=device).bernoulli_(0.5)


class ReplaySource(RandomSource):
    """Replays recorded CPU draws (lists of tensors) in order."""

    def __init__(self, perms, masks):
        self.perms, self.masks = list(perms), list(masks)

    replay = True

    def randperm(self, n, device):
        p = self.perms.pop(0)
        assert p.numel() == n, "randperm replay out of step: want {} have {}".format(n, p.numel())
        return p.to(device)

    def sample_keys(self, labels):
        """Keys that make the device sampler reproduce the recorded draws: the candidate the reference would
        pick i-th (positive[perm[i]]) gets key i.  Test-only: reads sizes on the host."""
        keys = torch.full(labels.shape, 3.0e7, dtype=torch.float32, device=labels.device)
        for cond in (labels >= 1, labels == 0):                    # same order as the reference: pos, then neg
            idx = torch.nonzero(cond).squeeze(1)
            perm = self.randperm(idx.numel(), labels.device)
            keys[idx[perm]] = torch.arange(idx.numel(), dtype=torch.float32, device=labels.device)
        return keys

    def consume_da_draws(self, counts):
        for n in counts.tolist():
            self.randperm(0, counts.device)
            self.randperm(int(n), counts.device)

    def dropout_keep(self, shape, device, row_valid=None):
        """row_valid: the recorded mask covers only the rows that exist; it is scattered into the fixed-capacity
        row layout (padding rows keep an all-ones mask; they carry no loss)."""
        m = self.masks.pop(0).to(device=device, dtype=torch.float32)
        if row_valid is not None:
            full = torch.ones(shape, dtype=torch.float32, device=device)
            idx = torch.nonzero(row_valid.bool()).squeeze(1)
            assert m.shape[0] == idx.numel() and tuple(m.shape[1:]) == tuple(shape[1:]), (tuple(m.shape), tuple(shape))
            full[idx] = m
            return full
        assert tuple(m.shape) == tuple(shape), (tuple(m.shape), tuple(shape))
        return m.contiguous()


class HashSource(RandomSource):
    """Deterministic, capture-safe pseudo-random draws (a pure function of the flat element index — NOT of the
    shape, so that a candidate keeps its key when the buffer it sits in is padded to another capacity): lets a test
    compare a captured whole-step graph with the eager step on identical random choices."""

    def _u(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        i = torch.arange(n, dtype=torch.float32, device=device)
        return torch.frac(torch.sin(i * 12.9898 + 78.233) * 43758.5453).abs().reshape(shape)

    def sample_keys(self, labels):
        return self._u(tuple(labels.shape), labels.device)

    def dropout_keep(self, shape, device, row_valid=None):
        return (self._u(tuple(shape), device) < 0.5).to(torch.float32)
