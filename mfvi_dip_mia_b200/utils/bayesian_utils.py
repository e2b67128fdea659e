"""Losses of the MFVI runners (reference utils/bayesian_utils.py:29-39) on the library's fused NLL kernel."""
from .. import functional as Fn


def gaussian_nll(mu, neg_logvar, target, reduction='mean'):
    """mean/sum of exp(clamp(s,-20,20)) * (target - mu)^2 - clamp(s)  with s = neg_logvar (no 1/2 factor).
    mu, neg_logvar: (N,1,H,W) (N = MC samples); target broadcasts as one (H,W) image."""
    return Fn.GaussianNllFn.apply(mu, neg_logvar, target, None, 0, reduction)


def gaussian_nll_inpainting(mu, neg_logvar, target, mask, reduction='mean'):
    """Masked variant: the loss map is multiplied by `mask` and averaged over ALL elements.
    mu (N,3,H,W) already sigmoid-ed by the caller (bayesian_optimization.py:3034), neg_logvar (N,1,H,W)."""
    return Fn.GaussianNllFn.apply(mu, neg_logvar, target, mask, 3, reduction)
