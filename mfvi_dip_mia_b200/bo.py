"""Bayesian-optimisation outer loop over (posterior temperature, prior sigma) — reference
bayesian_optimization.py:3545-3887 (`ExactGPModel`, `train_gp`, `expected_improvement`, `find_candidates`,
`normalize_X`, `bo`) without its gpytorch / skimage / matplotlib dependencies.

The trials themselves (the expensive part) are `runners.eval_trials`: one process per candidate, one candidate per
GPU, no communication.  What is here is the tiny host-side model that proposes the next candidates:
an exact GP in float64 on <= a few hundred points — constant mean with a N(15, 4) prior, scaled RBF kernel
(lengthscale initialised to 0.3), Gaussian noise with a Gamma(0.01, 100) prior and the > 1e-4 bound, hyper-parameters
fitted by 2000 Adam steps (lr 0.05) on the exact marginal log-likelihood — the model the reference builds from gpytorch
parts (:3546-3601), restated with torch.linalg.  Expected improvement on the 100x100 log grid, up to 4 local maxima +
the global one, each polished by L-BFGS through the sigmoid reparameterisation of [0,1] (:3604-3683).

gpytorch is not installed in the build container, so this restatement is checked against closed-form GP algebra and
scipy (tests/test_host_cpu.py), not against gpytorch outputs: "parity unpinned" for the GP hyper-parameter fit.
"""
from __future__ import annotations

import itertools
import math
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_SOFTPLUS_INV = lambda v: math.log(math.expm1(v))


class ExactGPModel(torch.nn.Module):
    """ConstantMean(prior N(15,4)) + ScaleKernel(RBFKernel) + GaussianLikelihood(noise prior Gamma(0.01, 100), noise >
    1e-4), all parameters softplus-constrained like gpytorch's defaults (reference :3546-3560, :3565-3567)."""

    def __init__(self, train_x: torch.Tensor, train_y: torch.Tensor):
        super().__init__()
        self.train_x, self.train_y = train_x.double(), train_y.double()
        z = lambda v: torch.nn.Parameter(torch.tensor(v, dtype=torch.float64, device=train_x.device))
        self.constant = z(0.0)
        self.raw_lengthscale = z(_SOFTPLUS_INV(3e-1))         # covar_module.base_kernel.lengthscale = 3e-1
        self.raw_outputscale = z(0.0)
        self.raw_noise = z(0.0)
        self._cache = None

    lengthscale = property(lambda self: F.softplus(self.raw_lengthscale))
    outputscale = property(lambda self: F.softplus(self.raw_outputscale))
    noise = property(lambda self: F.softplus(self.raw_noise) + 1e-4)

    def kernel(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        d2 = torch.cdist(a / self.lengthscale, b / self.lengthscale).pow(2)
        return self.outputscale * torch.exp(-0.5 * d2)

    def marginal_log_likelihood(self) -> torch.Tensor:
        """gpytorch.mlls.ExactMarginalLogLikelihood: (log N(y | c, K + noise I) + log-priors) / n."""
        n = self.train_y.numel()
        K = self.kernel(self.train_x, self.train_x) + self.noise * torch.eye(n, dtype=torch.float64, device=self.train_x.device)
        Lc = torch.linalg.cholesky(K)
        r = (self.train_y - self.constant).unsqueeze(1)
        alpha = torch.cholesky_solve(r, Lc)
        ll = -0.5 * (r * alpha).sum() - Lc.diagonal().log().sum() - 0.5 * n * math.log(2 * math.pi)
        ll = ll + torch.distributions.Normal(15.0, 4.0).log_prob(self.constant)
        ll = ll + torch.distributions.Gamma(0.01, 100.0).log_prob(self.noise)
        return ll / n

    def eval(self):
        self._cache = None
        return super().eval()

    def _posterior_cache(self):
        if self._cache is None:
            n = self.train_y.numel()
            K = self.kernel(self.train_x, self.train_x) + self.noise * torch.eye(n, dtype=torch.float64, device=self.train_x.device)
            Lc = torch.linalg.cholesky(K)
            alpha = torch.cholesky_solve((self.train_y - self.constant).unsqueeze(1), Lc)
            self._cache = (Lc, alpha)
        return self._cache

    def predict(self, X: torch.Tensor):
        """Posterior of the latent function at X (what `gp(X)` returns in eval mode): (mean, variance)."""
        X = X.double()
        if self.training:
            self._cache = None
        Lc, alpha = self._posterior_cache()
        Ks = self.kernel(X, self.train_x)
        mean = self.constant + (Ks @ alpha).squeeze(1)
        v = torch.linalg.solve_triangular(Lc, Ks.t(), upper=False)
        var = self.outputscale - v.pow(2).sum(0)
        return mean, var

    def confidence_region(self, X: torch.Tensor):
        mean, var = self.predict(X)
        sd = var.clamp_min(0).sqrt()
        return mean - 2 * sd, mean + 2 * sd


def train_gp(X_train: torch.Tensor, Y_train: torch.Tensor, iter_max: int = 2000, verbose: bool = False) -> ExactGPModel:
    """Adam (lr 0.05) on -MLL for iter_max iterations (reference :3563-3601).  Returns the model in eval mode."""
    gp = ExactGPModel(X_train, Y_train)
    gp.train()
    opt = torch.optim.Adam(gp.parameters(), lr=0.05)
    for i in range(iter_max):
        opt.zero_grad()
        loss = -gp.marginal_log_likelihood()
        loss.backward()
        if verbose and i % 100 == 0:
            print(f"Iter {i + 1:4d}/{iter_max} - Loss: {loss.item():.4f}   lengthscale: {gp.lengthscale.item():.3f}   "
                  f"noise: {gp.noise.item():.4f}")
        opt.step()
    gp.eval()
    return gp


def expected_improvement(gp: ExactGPModel, X: torch.Tensor, X_train: torch.Tensor) -> torch.Tensor:
    """EI over the best posterior mean at the observed points (reference :3604-3633); shape (m, 1)."""
    mu, var = gp.predict(X)
    mu_sample, _ = gp.predict(X_train)
    sigma = var.clamp_min(1e-9).sqrt().reshape(-1, 1)
    u = (mu - mu_sample.max()).reshape(-1, 1) / sigma
    normal = torch.distributions.Normal(torch.zeros_like(u), torch.ones_like(u))
    ei = sigma * (torch.exp(normal.log_prob(u)) + u * normal.cdf(u))
    return ei.clamp_min(0)


def upper_confidence_bound(gp: ExactGPModel, X: torch.Tensor, kappa: float = 2):
    mu, var = gp.predict(X)
    return mu + kappa * var.clamp_min(0).sqrt()


def acquisition_fun(gp, X, X_train, acq_fn, *args):
    assert acq_fn in ['ei', 'ucb']
    return expected_improvement(gp, X, X_train) if acq_fn == 'ei' else upper_confidence_bound(gp, X, *args)


def peak_local_max(image: np.ndarray, min_distance: int = 1, threshold_rel: Optional[float] = None,
                   num_peaks: Optional[int] = None) -> np.ndarray:
    """Coordinates of local maxima, strongest first — skimage.feature.peak_local_max semantics as the reference uses
    them (:3653): maximum filter of size 2*min_distance+1, peaks above threshold_rel*max, border of min_distance
    excluded."""
    from scipy.ndimage import maximum_filter
    size = 2 * min_distance + 1
    mask = image == maximum_filter(image, size=size, mode="nearest")
    if threshold_rel is not None:
        mask &= image > threshold_rel * image.max()
    if min_distance > 0:
        border = np.zeros_like(mask)
        border[min_distance:-min_distance, min_distance:-min_distance] = True
        mask &= border
    coords = np.argwhere(mask)
    order = np.argsort(-image[mask], kind="stable")
    coords = coords[order]
    return coords[:num_peaks] if num_peaks is not None else coords


def find_candidates(gp: ExactGPModel, X_: torch.Tensor, samples: torch.Tensor, acq_fn: str = 'ei', grid: int = 100):
    """Reference :3652-3683: acquisition on the grid, <= 4 local maxima + the global one, L-BFGS polish of the first 4
    through the sigmoid map onto [0,1]^2.  Returns (candidates [list of (1,2) tensors], acquisition values, grid acq)."""
    with torch.no_grad():
        acq = acquisition_fun(gp, X_, samples, acq_fn)
    acq = acq.cpu().numpy().reshape(grid, grid)
    peaks = peak_local_max(acq, min_distance=5, threshold_rel=0.1, num_peaks=4)
    global_max = np.array(np.unravel_index(np.argmax(acq, axis=None), acq.shape)).reshape(1, -1)
    peaks = np.unique(np.append(peaks, global_max, axis=0), axis=0)
    flat = np.ravel_multi_index(peaks.transpose(), acq.shape)
    X_init = X_[torch.as_tensor(flat, device=X_.device)]
    candidates, values = [], []
    for i in range(len(X_init[:4])):
        x0 = X_init[i].unsqueeze(0).double().clamp(1e-6, 1 - 1e-6)
        u = torch.logit(x0).clone().detach().requires_grad_(True)      # transform_to(interval(0,1)).inv
        minimizer = torch.optim.LBFGS([u], line_search_fn='strong_wolfe')

        def closure():
            minimizer.zero_grad()
            y = -acquisition_fun(gp, torch.sigmoid(u), samples, acq_fn).sum()
            y.backward()
            return y

        minimizer.step(closure)
        X = torch.sigmoid(u).detach()
        with torch.no_grad():
            values.append(acquisition_fun(gp, X, samples, acq_fn).item())
        candidates.append(X.cpu())
    return candidates, values, acq


def normalize_X(X_unnorm: torch.Tensor, x1_logbounds, x2_logbounds) -> torch.Tensor:
    """log10, then each axis mapped from its log-bounds onto [0,1] (reference :3686-3694)."""
    X = X_unnorm.clone().log10()
    X[:, 0] = (X[:, 0] - x1_logbounds[0]) / (x1_logbounds[1] - x1_logbounds[0])
    X[:, 1] = (X[:, 1] - x2_logbounds[0]) / (x2_logbounds[1] - x2_logbounds[0])
    return X


def unnormalize_X(X_norm: torch.Tensor, x1_logbounds, x2_logbounds) -> torch.Tensor:
    X = X_norm.clone()
    X[:, 0] = X[:, 0] * (x1_logbounds[1] - x1_logbounds[0]) + x1_logbounds[0]
    X[:, 1] = X[:, 1] * (x2_logbounds[1] - x2_logbounds[0]) + x2_logbounds[0]
    return torch.pow(10, X)


def bo(trial_fn: Callable[..., float], bo_params: Dict[str, Dict[str, Sequence[float]]], run_params: Dict, *,
       rounds: int = 20, gp_iters: int = 2000, start_method: str = "spawn", verbose: bool = True, persistent: bool = True):
    """The reference's `bo()` (:3726-3887) with the trial runner passed in: every round evaluates the current candidates
    (one process per candidate, devices round-robin, NaN results dropped), refits the GP on ALL observations, proposes
    the next candidates and writes `<bo_results_path>/<round>_fig_data.npz` with the reference's keys.
    bo_params: {"<name1>": {"logbounds": [lo, hi], "candidates": [...]}, "<name2>": {...}} (bo_configs/*.json);
    run_params: keyword arguments of the trial runner plus "bo_results_path" and "devices".  Returns (X, Y).
    persistent (default): the trials of ALL rounds run on one worker process per device (runners.TrialPool), so that a round
    costs the trials' kernels instead of a process start + CUDA context per trial; False = one process per trial, as the
    reference starts them."""
    from .runners import TrialPool, _finite, eval_trials
    run_params = dict(run_params)
    out_path = run_params.pop("bo_results_path")
    devices = list(run_params.pop("devices"))
    os.makedirs(out_path, exist_ok=True)
    (p1_logbounds, p2_logbounds) = [v["logbounds"] for v in bo_params.values()]
    X_lr = torch.logspace(*p1_logbounds, 100, dtype=torch.double)
    X_wd = torch.logspace(*p2_logbounds, 100, dtype=torch.double)
    XX_lr, XX_wd = torch.meshgrid(X_lr, X_wd, indexing="ij")
    X_ = torch.stack([XX_lr.reshape(-1), XX_wd.reshape(-1)]).transpose(1, 0)
    X_test = normalize_X(X_, p1_logbounds, p2_logbounds)
    candidates = list(itertools.product(*[v["candidates"] for v in bo_params.values()]))
    X: List = []
    Y: List[float] = []
    pool = TrialPool(devices, trial_fn, run_params, start_method=start_method) if persistent else None
    try:
        for r in range(rounds):
            if pool is not None:
                cl = [tuple(c) for c in candidates]
                cands_run, y_run = _finite(cl, pool.run(cl))
            else:
                cands_run, y_run = eval_trials(candidates, devices, trial_fn, run_params, start_method=start_method)
            if verbose:
                names = list(bo_params.keys())
                print(f"\n{names[0]}      {names[1]}       psnr")
                for c, y in zip(cands_run, y_run):
                    print(f"{c[0]:.6g}  {c[1]:.6g}  {y:.6f}")
            X += [tuple(c) for c in cands_run]
            Y += list(y_run)
            X_train = normalize_X(torch.tensor(np.array(X), dtype=torch.double), p1_logbounds, p2_logbounds)
            Y_train = torch.tensor(np.array(Y), dtype=torch.double)
            gp = train_gp(X_train, Y_train, iter_max=gp_iters, verbose=False)
            cands, exp_imp, acq = find_candidates(gp, X_test, X_train)
            cands = torch.unique(torch.cat(cands), dim=0)
            cand_np = unnormalize_X(cands, p1_logbounds, p2_logbounds).numpy()
            with torch.no_grad():
                mean, _ = gp.predict(X_test)
                lo, hi = gp.confidence_region(X_test)
            np.savez(os.path.join(out_path, f"{r}_fig_data.npz"), XX_lr=XX_lr.numpy(), XX_wd=XX_wd.numpy(),
                     pred=mean.reshape(100, 100).numpy(), observed_X=np.array(X), observed_Y=np.array(Y),
                     expected_improvement=np.array(exp_imp), confidence=(hi - lo).reshape(100, 100).numpy(),
                     acq=acq.reshape(100, 100), candidates=cand_np)
            candidates = [tuple(float(v) for v in c) for c in cand_np]
    finally:
        if pool is not None:
            pool.close()
    return X, Y
