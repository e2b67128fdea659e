"""Uncertainty calibration error (reference utils/uce.py:9-40): 15 equal-width uncertainty bins (lo, hi]."""
import torch


def uceloss(errors, uncert, n_bins=15, outlier=0.0, range=None):
    device = errors.device
    lo, hi = (uncert.min().item(), uncert.max().item()) if range is None else range
    edges = torch.linspace(lo, hi, n_bins + 1, device=device)
    uce = torch.zeros(1, device=device)
    errs, uncs, props = [], [], []
    for a, b in zip(edges[:-1], edges[1:]):
        in_bin = uncert.gt(a.item()) * uncert.le(b.item())
        prop = in_bin.float().mean()
        props.append(prop)
        if prop.item() > outlier:
            e = errors[in_bin].float().mean()
            u = uncert[in_bin].mean()
            uce += torch.abs(u - e) * prop
            errs.append(e)
            uncs.append(u)
    return uce, torch.tensor(errs, device=device), torch.tensor(uncs, device=device), torch.tensor(props, device=device)
