"""Trial runner and trial fan-out around the MFVI-DIP training step.

`run_den_mfvi` is the denoising runner of the reference (bayesian_optimization.py:1240-1444) on the fast path: the
hot loop (:1360-1372) is `MfviDipTrainer.step()`, and the per-iteration bookkeeping (:1374-1416) stays ON THE DEVICE —
exp(-s), the 0.99/0.01 exponential moving average of the output, clipping, the ring buffers of the last 25 outputs
(epistemic = variance of the predicted means, aleatoric = mean of the predicted variances), PSNR / SSIM — and is read
back only every `show_every` iterations instead of ~8 host synchronisations per iteration.  The return value is the
reference's BO objective: PSNR of the EMA-smoothed output against the ground truth at the last iteration (:1444).

`eval_trials` is the fan-out of `f()` / `bo()` / `eval()` (bayesian_optimization.py:3709-3775, eval_result.py:19-47):
one OS process per (temp, sigma) candidate, devices assigned round-robin, results `(candidate, objective)` returned
through a queue, NaN results dropped.  Trials never communicate ("replicas only").
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .utils.common_utils import peak_signal_noise_ratio, structural_similarity


def run_den_mfvi(img_gt: np.ndarray, *, temp: float, sigma: float, lr: float = 1e-3, num_iter: int = 100,
                 mc_samples: int = 1, p_sigma: float = 0.1, seed: int = 1, device="cuda:0", input_depth: int = 16,
                 reg_noise_std: float = 0.1, exp_weight: float = 0.99, mc_ring: int = 25, show_every: int = 100,
                 math_mode: Optional[int] = None, rank: int = 0, world_size: int = 1, process_group=None,
                 return_history: bool = False):
    """img_gt: (1,H,W) ground-truth image in [0,1] with H, W multiples of 32.  Returns psnr_gt_sm of the last
    iteration (and, with return_history, a dict of the per-`show_every` metrics and the final uncertainty maps)."""
    from . import MfviDipTrainer, SkipSpec, _lib as L
    from .utils.common_utils import get_noise
    dev = torch.device(device)
    np.random.seed(seed)
    torch.manual_seed(seed)
    gt = torch.as_tensor(img_gt, dtype=torch.float32)[None].to(dev)                  # (1,1,H,W)
    noisy = np.clip(img_gt + np.random.normal(scale=p_sigma, size=img_gt.shape), 0, 1).astype(np.float32)
    noisy_t = torch.from_numpy(noisy)[None].to(dev)
    H, W = img_gt.shape[-2:]
    net_input = get_noise(input_depth, 'noise', (H, W))
    tr = MfviDipTrainer(SkipSpec(input_depth, 2), "den", net_input, temp=temp, sigma=sigma, lr=lr, mc_samples=mc_samples,
                        seed=seed, reg_noise_std=reg_noise_std, device=dev, target=noisy_t,
                        math_mode=L.MATH_TF32 if math_mode is None else math_mode, rank=rank, world_size=world_size,
                        process_group=process_group)
    n_steps = num_iter + 1                       # the reference runs num_iter + 1 iterations (:1287)
    out_avg = None
    ring_epi = torch.zeros(mc_ring, 1, H, W, device=dev)
    ring_ale = torch.zeros(mc_ring, 1, H, W, device=dev)
    hist: Dict[str, List[float]] = {"it": [], "psnr_noisy": [], "psnr_gt": [], "psnr_gt_sm": [], "ssim_gt_sm": [], "loss": []}
    for i in range(n_steps):
        tr.step()
        out = tr.eng.out_nchw()                                   # (S,2,H,W): ch0 = mean, ch1 = -log var
        mean = out[:, :1].mean(0, keepdim=True)
        var = torch.exp(-out[:, 1:]).mean(0, keepdim=True)        # out[:,1:] = exp(-s)   (:1375)
        cur = torch.cat([mean, var], 1)
        out_avg = cur.clone() if out_avg is None else out_avg * exp_weight + cur * (1 - exp_weight)   # (:1378-1381)
        ring_epi[i % mc_ring] = mean[0].clamp(0, 1)
        ring_ale[i % mc_ring] = var[0].clamp(0, 1)
        if i % show_every == 0 or i == n_steps - 1:
            sm = out_avg[:, :1].clamp(0, 1)
            nll, kl, loss = tr.loss_terms()
            hist["it"].append(i)
            hist["loss"].append(loss)
            hist["psnr_noisy"].append(peak_signal_noise_ratio(noisy_t, mean.clamp(0, 1)))
            hist["psnr_gt"].append(peak_signal_noise_ratio(gt, mean.clamp(0, 1)))
            hist["psnr_gt_sm"].append(peak_signal_noise_ratio(gt, sm))
            hist["ssim_gt_sm"].append(structural_similarity(gt, sm))
    psnr_gt_sm = hist["psnr_gt_sm"][-1]
    if not return_history:
        return psnr_gt_sm
    n_valid = min(n_steps, mc_ring)
    hist["epistemic"] = ring_epi[:n_valid].var(0, unbiased=True).cpu() if n_valid > 1 else torch.zeros(1, H, W)
    hist["aleatoric"] = ring_ale[:n_valid].mean(0).cpu()
    hist["recon"] = out_avg[0, :1].clamp(0, 1).cpu()
    return psnr_gt_sm, hist


def _trial_entry(fn, kwargs, candidate, device, queue):
    """Body of one trial process (reference f(), bayesian_optimization.py:3709-3724)."""
    try:
        val = fn(temp=candidate[0], sigma=candidate[1], device=device, **kwargs)
    except Exception as e:  # a crashed trial is reported as NaN and dropped, like a diverged one
        print(f"[trial {candidate} on {device}] failed: {type(e).__name__}: {e}", flush=True)
        val = float("nan")
    queue.put((tuple(candidate), float(val)))


def eval_trials(candidates: Iterable[Sequence[float]], devices: Sequence[str], fn: Callable[..., float],
                fn_kwargs: Optional[dict] = None, *, max_parallel: Optional[int] = None,
                start_method: str = "spawn") -> Tuple[List[Tuple[float, float]], List[float]]:
    """Run `fn(temp=, sigma=, device=, **fn_kwargs)` once per candidate, one process per trial, devices assigned
    round-robin; at most `max_parallel` (default: one per device) trials run at a time.  Returns (candidates, values)
    with NaN results removed (bayesian_optimization.py:3772-3781)."""
    import queue as queue_mod
    import torch.multiprocessing as mp
    ctx = mp.get_context(start_method)
    cands = [tuple(c) for c in candidates]
    max_parallel = max_parallel or len(devices)
    queue = ctx.Queue()
    dev_cycle = itertools.cycle(devices)
    results: Dict[Tuple[float, float], float] = {}
    pending = list(cands)
    running = {}                                    # process -> candidate
    while pending or running:
        while pending and len(running) < max_parallel:
            c = pending.pop(0)
            pr = ctx.Process(target=_trial_entry, args=(fn, fn_kwargs or {}, c, next(dev_cycle), queue))
            pr.start()
            running[pr] = c
        try:
            c, v = queue.get(timeout=0.5)
            results[c] = v
        except queue_mod.Empty:
            pass
        for pr in [q for q in running if not q.is_alive()]:
            pr.join()
            c = running.pop(pr)
            try:                                     # its result may still sit in the queue
                while c not in results:
                    c2, v2 = queue.get(timeout=0.5)
                    results[c2] = v2
            except queue_mod.Empty:
                results.setdefault(c, float("nan"))  # the process died without reporting: dropped like a NaN trial
    X, Y = [], []
    for c in cands:
        v = results.get(c, float("nan"))
        if not math.isnan(v):
            X.append(c)
            Y.append(v)
    return X, Y


def log_grid(bounds_log10: Sequence[Sequence[float]], n: int) -> List[Tuple[float, float]]:
    """n x n (temp, sigma) candidates on a log10 grid inside `bounds_log10` = [[lo_t, hi_t], [lo_s, hi_s]]
    (bo_configs/bo_mfvi.json: logbounds [-10, 0]^2)."""
    t = np.logspace(bounds_log10[0][0], bounds_log10[0][1], n)
    s = np.logspace(bounds_log10[1][0], bounds_log10[1][1], n)
    return [(float(a), float(b)) for a in t for b in s]
