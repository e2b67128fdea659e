"""Uncertainty calibration error (reference utils/uce.py:9-40): the uncertainty range [min, max] is cut into `n_bins`
equal-width bins (lo, hi]; UCE = sum over occupied bins of |mean uncertainty - mean error| * (share of the pixels)."""
import torch


def uceloss(errors, uncert, n_bins=15, outlier=0.0, range=None):
    """Returns (uce[1], per-bin mean error, per-bin mean uncertainty, per-bin share) — the per-bin vectors list the bins
    whose share exceeds `outlier`, the share vector lists all bins (as the reference does)."""
    lo, hi = (uncert.min().item(), uncert.max().item()) if range is None else range
    edges = torch.linspace(lo, hi, n_bins + 1, device=errors.device)
    # membership of every element in every bin at once: (n_bins, N), lower edge exclusive, upper edge inclusive
    member = (uncert.reshape(1, -1) > edges[:-1].reshape(-1, 1)) & (uncert.reshape(1, -1) <= edges[1:].reshape(-1, 1))
    count = member.sum(1)
    share = count.float() / uncert.numel()
    keep = share > outlier
    safe = count.clamp_min(1).float()
    mean_err = (member * errors.reshape(1, -1).float()).sum(1) / safe
    mean_unc = (member * uncert.reshape(1, -1)).sum(1) / safe
    uce = (torch.abs(mean_unc - mean_err) * share)[keep].sum().reshape(1)
    return uce, mean_err[keep], mean_unc[keep], share
