"""CPU checks of the host logic behind the device-side fg/bg sampler: the "smallest random key wins" formulation
(dadetect_b200.utils.random_source + ops.balanced_sample) selects exactly what the reference's
BalancedPositiveNegativeSampler selects with `positive[randperm(n)[:k]]`
(maskrcnn_benchmark/modeling/balanced_positive_negative_sampler.py:27-76) when the keys are the ranks in the same
permutations.  No kernel is launched here; the kernel itself is compared with this rule in tests/test_gpu_ops.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import da_frcnn_ref as orc                                    # noqa: E402
from dadetect_b200.utils.random_source import ReplaySource   # noqa: E402


def _select_by_keys(labels, keys, batch, max_pos):
    """What ops.balanced_sample computes (ties -> lower index), in plain torch."""
    pos = torch.nonzero(labels >= 1).squeeze(1)
    neg = torch.nonzero(labels == 0).squeeze(1)
    num_pos = min(pos.numel(), max_pos)
    num_neg = min(neg.numel(), batch - num_pos)
    pick = lambda idx, k: idx[torch.sort(keys[idx], stable=True)[1][:k]]
    return torch.sort(torch.cat([pick(pos, num_pos), pick(neg, num_neg)]))[0], num_pos


def test_replayed_keys_reproduce_the_reference_sampler():
    g = torch.Generator().manual_seed(0)
    for n, p_pos, p_neg, batch, frac in [(5000, 0.01, 0.7, 256, 0.5), (300, 0.3, 0.5, 256, 0.25), (64, 0.0, 0.9, 16, 0.5),
                                         (2000, 0.2, 0.0, 256, 0.25)]:
        u = torch.rand(n, generator=g)
        labels = torch.where(u < p_pos, 2, torch.where(u < p_pos + p_neg, 0, -1)).to(torch.int32)
        # the oracle's sampler (a line-by-line restatement of the reference's), with its draws recorded
        rec = orc.RecordingHooks()
        torch.manual_seed(n)
        pos_m, neg_m = orc.balanced_sampler([labels.to(torch.float32)], batch, frac, rec)
        want = torch.nonzero(pos_m[0] | neg_m[0]).squeeze(1)
        # the key formulation fed with the same permutations
        keys = ReplaySource(rec.perms, []).sample_keys(labels)
        got, num_pos = _select_by_keys(labels, keys, batch, int(batch * frac))
        assert torch.equal(got, want)
        assert num_pos == int(pos_m[0].sum())


def test_uniform_keys_pick_a_uniform_subset():
    """Production keys are i.i.d. uniform: every positive is equally likely to be among the k chosen."""
    g = torch.Generator().manual_seed(1)
    labels = torch.zeros(200, dtype=torch.int32)
    labels[:40] = 1
    hits = torch.zeros(40)
    trials = 3000
    for _ in range(trials):
        keys = torch.rand(200, generator=g)
        sel, _ = _select_by_keys(labels, keys, 64, 10)
        hits[sel[sel < 40]] += 1
    freq = hits / trials                         # expected 10 / 40 = 0.25 each
    assert float((freq - 0.25).abs().max()) < 0.04
