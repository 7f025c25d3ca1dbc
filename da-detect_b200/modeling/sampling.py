"""Label assignment glue shared by the RPN and box-head losses.

The heavy parts are kernels (fused IoU+Matcher, box encode); what remains here is the
reference's per-image fg/bg subsampling with ``randperm`` (balanced_positive_negative_sampler.py:27-76),
kept in the same call order so the RNG stream lines up with the reference (SURVEY §7 "RNG-stream parity").
"""
import torch

BELOW_LOW_THRESHOLD = -1
BETWEEN_THRESHOLDS = -2


def balanced_sample(labels_per_image, batch_size_per_image, positive_fraction, rng):
    """labels: list of tensors with -1 (ignore) / 0 (negative) / >=1 (positive).
    Returns two lists of bool masks (selected positives, selected negatives)."""
    pos_masks, neg_masks = [], []
    for lab in labels_per_image:
        positive = torch.nonzero(lab >= 1).squeeze(1)
        negative = torch.nonzero(lab == 0).squeeze(1)
        num_pos = min(positive.numel(), int(batch_size_per_image * positive_fraction))
        num_neg = min(negative.numel(), batch_size_per_image - num_pos)
        perm1 = rng.randperm(positive.numel(), lab.device)[:num_pos]
        perm2 = rng.randperm(negative.numel(), lab.device)[:num_neg]
        pm = torch.zeros_like(lab, dtype=torch.bool)
        nm = torch.zeros_like(lab, dtype=torch.bool)
        pm[positive[perm1]] = True
        nm[negative[perm2]] = True
        pos_masks.append(pm)
        neg_masks.append(nm)
    return pos_masks, neg_masks
