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



def _on_device(fn):
    from ._lib import on_device
    return on_device(fn)


class DeviceBookkeeping:
    """Per-iteration bookkeeping of the runners (reference bayesian_optimization.py:1374-1416; SR :2190-2222, inpainting
    :3041-3069, CT :583-610) as ONE kernel per iteration, registered as a post-step hook of the trainer so that it is replayed
    inside the step's CUDA graph; the iteration index comes from the trainer's device-side step counter.  Nothing is read back
    until `metrics()`.

    `channels` = image channels of the network output (1; 3 for inpainting), `sigmoid` squashes them (inpainting),
    `aleatoric` = the output carries s = -log sigma^2 behind them (False for CT), `mask` (H,W) multiplies both images of the
    ground-truth comparisons (inpainting).  gt / noisy: (channels,H,W)."""

    N_ACC = 8

    def __init__(self, trainer, gt=None, noisy=None, exp_weight: float = 0.99, ring: int = 25, *, channels: int = 1,
                 sigmoid: bool = False, aleatoric: bool = True, mask=None):
        from . import _lib as L
        self.L, self.tr = L, trainer
        self.device = trainer.eng.device
        e = trainer.eng
        self.S, self.H, self.W, _ = e.out.shape
        dev = e.device
        self.Cm, self.flags = int(channels), (1 if sigmoid else 0) | (2 if aleatoric else 0)
        self.aleatoric = aleatoric
        f = lambda t, c: None if t is None else torch.as_tensor(t, dtype=torch.float32).reshape(c, self.H, self.W).to(dev).contiguous()
        self.gt, self.noisy, self.mask = f(gt, self.Cm), f(noisy, self.Cm), f(mask, 1)
        self.exp_weight, self.ring = float(exp_weight), int(ring)
        self.out_avg = torch.zeros(self.Cm + (1 if aleatoric else 0), self.H, self.W, device=dev)
        self.ring_epi = torch.zeros(self.Cm, max(ring, 1), self.H, self.W, device=dev)
        self.ring_ale = torch.zeros(max(ring, 1), self.H, self.W, device=dev)
        self.acc = torch.zeros(self.N_ACC, dtype=torch.float64, device=dev)     # [0..4] squared errors, [5..6] SSIM sums
        trainer.post_step_hooks.append(self.record)

    @_on_device
    def record(self):
        """Enqueue the bookkeeping of the step that was just computed (called by the trainer before the step counter
        advances; asynchronous)."""
        L, e = self.L, self.tr.eng
        L.call("mfvi_fill_f32", self.acc.data_ptr(), 2 * self.N_ACC, 0.0)
        L.call("mfvi_bookkeep_step_ex", L.view(e.out), self.S, self.H, self.W, self.Cm, self.flags, self.exp_weight,
               L.ptr(self.gt), L.ptr(self.noisy), L.ptr(self.mask), self.out_avg.data_ptr(), self.ring_epi.data_ptr(),
               self.ring_ale.data_ptr(), self.ring, self.tr.step_dev.data_ptr(), 0, self.acc.data_ptr(),
               meta={"bytes": 4.0 * self.H * self.W * ((self.Cm + 1) * self.S + 8 * self.Cm)})

    @_on_device
    def metrics(self, ssim: bool = True) -> Dict[str, float]:
        """PSNR / MSE / SSIM of the last recorded iteration (one synchronising read of 8 doubles)."""
        L = self.L
        n = self.Cm * self.H * self.W
        if ssim and self.gt is not None:
            L.call("mfvi_fill_f32", self.acc[5:].data_ptr(), 2 * (self.N_ACC - 5), 0.0)     # mfvi_ssim accumulates
            a_img, b_img = self.gt, self.out_avg[:self.Cm]
            if self.mask is not None:            # inpainting compares img*mask with clip(out_avg)*mask (:3065, :3068)
                a_img, b_img = (a_img * self.mask).contiguous(), (b_img.clamp(0, 1) * self.mask).contiguous()
            for c in range(self.Cm):             # the window is per channel; equal-sized planes => mean over channels
                L.call("mfvi_ssim", a_img[c].data_ptr(), b_img[c].data_ptr(), self.H, self.W, 1, self.acc[5:].data_ptr())
        a = self.acc.cpu().tolist()
        psnr = lambda sse: 10.0 * math.log10(n / sse) if sse > 0 else float("inf")
        m = {"mse_corrupted": a[3] / n, "mse_gt": a[4] / n}
        if self.noisy is not None:
            m["psnr_noisy"] = psnr(a[0])
        if self.gt is not None:
            m.update(psnr_gt=psnr(a[1]), psnr_gt_sm=psnr(a[2]))
            if ssim:
                m["ssim_gt_sm"] = a[5] / n
        return m

    @_on_device
    def uncertainty(self, n_valid: Optional[int] = None):
        """(epistemic, aleatoric, err2) maps from the ring buffers (:1410-1411), each (channels,H,W) — (H,W) for one channel;
        err2 is None without a ground truth, aleatoric is zero for a net without the s channel."""
        L = self.L
        n = min(self.tr.steps_done, self.ring) if n_valid is None else n_valid
        epi = torch.empty(self.Cm, self.H, self.W, device=self.out_avg.device)
        ale = torch.zeros_like(epi)
        err2 = torch.empty_like(epi) if self.gt is not None else None
        for c in range(self.Cm):
            L.call("mfvi_ring_uncertainty", self.ring_epi[c].data_ptr(), self.ring_ale.data_ptr(), max(n, 1), self.H, self.W,
                   None if self.gt is None else self.gt[c].data_ptr(), epi[c].data_ptr(), ale[c].data_ptr(),
                   None if err2 is None else err2[c].data_ptr())
        if not self.aleatoric:
            ale.zero_()
        if self.Cm == 1:
            return epi[0], ale[0], (None if err2 is None else err2[0])
        return epi, ale, err2

    def uce(self, n_bins: int = 15) -> float:
        """Uncertainty calibration error of (epistemic + aleatoric) against the squared error of the ring means
        (reference utils/uce.py, recipe eval_denoising.ipynb:467-482)."""
        from .utils.uce import uceloss
        epi, ale, err2 = self.uncertainty()
        return float(uceloss(err2.reshape(-1), (epi + ale).reshape(-1), n_bins=n_bins)[0])


def _run_loop(tr, bk, num_iter: int, show_every: int, return_history: bool):
    """The runners' loop (:1359-1422): num_iter + 1 iterations (:1287) of hot loop + device bookkeeping, metrics read back every
    `show_every`; returns the BO objective psnr_gt_sm of the last iteration (:1444)."""
    n_steps = num_iter + 1
    hist: Dict[str, list] = {"it": [], "psnr_noisy": [], "psnr_gt": [], "psnr_gt_sm": [], "ssim_gt_sm": [], "loss": []}
    for i in range(n_steps):
        tr.step()                                 # hot loop + bookkeeping kernel, no host synchronisation
        if i % show_every == 0 or i == n_steps - 1:
            m = bk.metrics()
            hist["it"].append(i)
            hist["loss"].append(tr.loss_terms()[2])
            for k in ("psnr_noisy", "psnr_gt", "psnr_gt_sm", "ssim_gt_sm"):
                hist[k].append(m.get(k, float("nan")))
    psnr_gt_sm = hist["psnr_gt_sm"][-1]
    if not return_history:
        return psnr_gt_sm
    epi, ale, _ = bk.uncertainty()
    hist["epistemic"], hist["aleatoric"] = epi.cpu(), ale.cpu()
    hist["uce"] = bk.uce()
    hist["recon"] = bk.out_avg[:bk.Cm].clamp(0, 1).cpu()
    return psnr_gt_sm, hist


def _math(L, math_mode):
    return L.MATH_TF32 if math_mode is None else math_mode


def run_den_mfvi(img_gt: np.ndarray, *, temp: float, sigma: float, lr: float = 1e-3, num_iter: int = 100,
                 mc_samples: int = 1, p_sigma: float = 0.1, seed: int = 1, device="cuda:0", input_depth: int = 16,
                 reg_noise_std: float = 0.1, exp_weight: float = 0.99, mc_ring: int = 25, show_every: int = 100,
                 math_mode: Optional[int] = None, rank: int = 0, world_size: int = 1, process_group=None,
                 return_history: bool = False, spec=None, img_noisy: Optional[np.ndarray] = None):
    """Denoising runner (reference run_den_mfvi, bayesian_optimization.py:1240-1444).  img_gt: (1,H,W) ground-truth image in
    [0,1] with H, W multiples of 32.  Returns psnr_gt_sm of the last iteration (and, with return_history, a dict of the
    per-`show_every` metrics and the final uncertainty maps).  `spec` (a SkipSpec) overrides the runner's 5-scale net;
    `img_noisy` overrides the seeded noisy observation.  `math_mode` defaults to the tf32 tensor-core mode (the separately
    stated reduced-precision mode); pass _lib.MATH_FP32 for reference arithmetic."""
    from . import MfviDipTrainer, SkipSpec, _lib as L
    from .utils.common_utils import get_noise
    dev = torch.device(device)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if img_noisy is None:
        img_noisy = np.clip(img_gt + np.random.normal(scale=p_sigma, size=img_gt.shape), 0, 1).astype(np.float32)
    H, W = img_gt.shape[-2:]
    spec = spec or SkipSpec(input_depth, 2)
    net_input = get_noise(spec.num_input_channels, 'noise', (H, W))
    tr = MfviDipTrainer(spec, "den", net_input, temp=temp, sigma=sigma, lr=lr, mc_samples=mc_samples,
                        seed=seed, reg_noise_std=reg_noise_std, device=dev, target=torch.from_numpy(img_noisy)[None],
                        math_mode=_math(L, math_mode), rank=rank, world_size=world_size, process_group=process_group)
    bk = DeviceBookkeeping(tr, gt=img_gt, noisy=img_noisy, exp_weight=exp_weight, ring=mc_ring)
    return _run_loop(tr, bk, num_iter, show_every, return_history)


def run_sr_mfvi(img_hr: np.ndarray, *, temp: float, sigma: float, factor: int = 4, lr: float = 1e-3, num_iter: int = 100,
                mc_samples: int = 1, seed: int = 2, device="cuda:0", input_depth: int = 32, reg_noise_std: float = 0.1,
                exp_weight: float = 0.99, mc_ring: int = 25, show_every: int = 100, math_mode: Optional[int] = None, rank: int = 0,
                world_size: int = 1, process_group=None, return_history: bool = False, spec=None):
    """4x super-resolution runner (reference run_sr_mfvi, bayesian_optimization.py:2048-2263): the low-resolution observation
    is every `factor`-th pixel of img_hr (nearest, :2095-2102), the NLL is taken on the equally sub-sampled output (:2182-2185),
    bookkeeping and the objective psnr_gt_sm on the full-resolution output (:2190-2222)."""
    from . import MfviDipTrainer, SkipSpec, _lib as L
    from .utils.common_utils import get_noise
    dev = torch.device(device)
    np.random.seed(seed)
    torch.manual_seed(seed)
    H, W = img_hr.shape[-2:]
    img_lr = np.ascontiguousarray(img_hr[..., ::factor, ::factor])
    spec = spec or SkipSpec(input_depth, 2)
    net_input = get_noise(spec.num_input_channels, 'noise', (H, W))
    tr = MfviDipTrainer(spec, "sr", net_input, temp=temp, sigma=sigma, lr=lr, mc_samples=mc_samples, seed=seed,
                        reg_noise_std=reg_noise_std, device=dev, target=torch.from_numpy(img_lr)[None], sr_factor=factor,
                        math_mode=_math(L, math_mode), rank=rank, world_size=world_size, process_group=process_group)
    bk = DeviceBookkeeping(tr, gt=img_hr, noisy=None, exp_weight=exp_weight, ring=mc_ring)
    return _run_loop(tr, bk, num_iter, show_every, return_history)


def run_inp_mfvi(img: np.ndarray, mask: np.ndarray, *, temp: float, sigma: float, lr: float = 2e-3, num_iter: int = 100,
                 mc_samples: int = 1, seed: int = 2, device="cuda:0", input_depth: int = 16, reg_noise_std: float = 0.1,
                 exp_weight: float = 0.99, mc_ring: int = 25, show_every: int = 100, math_mode: Optional[int] = None,
                 rank: int = 0, world_size: int = 1, process_group=None, return_history: bool = False, spec=None):
    """Inpainting runner (reference run_inp_mfvi, bayesian_optimization.py:2892-3114): img (3,H,W), mask (1,H,W) rounded to
    {0,1} (:3024); 6-scale net with 5x5 down filters, no skip branches, no 1x1 up convs, nearest upsampling, 4 outputs
    (:2970-2998); masked NLL on sigmoid(out[:, :3]) (:3033-3036); metrics on img*mask vs out*mask (:3064-3069)."""
    from . import MfviDipTrainer, SkipSpec, _lib as L
    from .utils.common_utils import get_noise
    dev = torch.device(device)
    np.random.seed(seed)
    torch.manual_seed(seed)
    H, W = img.shape[-2:]
    mask = np.round(mask).astype(np.float32)
    spec = spec or SkipSpec(input_depth, 4, (16, 32, 64, 128, 128, 128), (16, 32, 64, 128, 128, 128), (0,) * 6, 5, 3, 1, False,
                            False, "nearest")
    net_input = get_noise(spec.num_input_channels, 'noise', (H, W))
    tr = MfviDipTrainer(spec, "inp", net_input, temp=temp, sigma=sigma, lr=lr, mc_samples=mc_samples, seed=seed,
                        reg_noise_std=reg_noise_std, device=dev, target=torch.from_numpy(img)[None],
                        mask=torch.from_numpy(mask)[None], math_mode=_math(L, math_mode), rank=rank, world_size=world_size,
                        process_group=process_group)
    bk = DeviceBookkeeping(tr, gt=img, noisy=img, exp_weight=exp_weight, ring=mc_ring, channels=3, sigmoid=True, mask=mask)
    return _run_loop(tr, bk, num_iter, show_every, return_history)


def run_ct_mfvi(img_gt: np.ndarray, *, temp: float, sigma: float, theta_deg=None, lr: float = 1e-3, num_iter: int = 100,
                mc_samples: int = 1, seed: int = 1, device="cuda:0", input_depth: int = 16, reg_noise_std: float = 0.1,
                exp_weight: float = 0.99, mc_ring: int = 25, show_every: int = 100, math_mode: Optional[int] = None, rank: int = 0,
                world_size: int = 1, process_group=None, return_history: bool = False, spec=None):
    """Sparse-view CT runner (reference run_ct_mfvi, bayesian_optimization.py:442-648): img_gt (1,H,H); the observation is its
    sinogram under FastRadonTransform at `theta_deg` (default arange(0,180,4) = 45 angles, :545-547); one output channel, plain
    MSE in sinogram space (:533,576), AdamW skipped on a non-finite loss (:581-582)."""
    from . import MfviDipTrainer, SkipSpec, _lib as L
    from .radon import FastRadonTransform
    from .utils.common_utils import get_noise
    dev = torch.device(device)
    np.random.seed(seed)
    torch.manual_seed(seed)
    H, W = img_gt.shape[-2:]
    theta = torch.arange(0, 180., step=4.) if theta_deg is None else torch.as_tensor(theta_deg, dtype=torch.float32)
    img_t = torch.from_numpy(np.ascontiguousarray(img_gt))[None].to(dev)
    sino = FastRadonTransform(tuple(img_t.shape), theta).to(dev)(img_t).detach()
    spec = spec or SkipSpec(input_depth, 1)
    net_input = get_noise(spec.num_input_channels, 'noise', (H, W))
    tr = MfviDipTrainer(spec, "ct", net_input, temp=temp, sigma=sigma, lr=lr, mc_samples=mc_samples, seed=seed,
                        reg_noise_std=reg_noise_std, device=dev, theta_deg=theta, sino=sino, math_mode=_math(L, math_mode),
                        rank=rank, world_size=world_size, process_group=process_group)
    bk = DeviceBookkeeping(tr, gt=img_gt, noisy=img_gt, exp_weight=exp_weight, ring=mc_ring, aleatoric=False)
    return _run_loop(tr, bk, num_iter, show_every, return_history)


RUNNERS = {"den": run_den_mfvi, "sr": run_sr_mfvi, "inp": run_inp_mfvi, "ct": run_ct_mfvi}
_TASK_ALIASES = {"denoising": "den", "inpainting": "inp", "super-resolution": "sr", "ct": "ct", "den": "den", "sr": "sr", "inp": "inp"}


def f(task: str, bayes: str, candidate: Sequence[float], device, params: dict) -> float:
    """The reference's trial wrapper (bayesian_optimization.py:3709-3724): maps (task, bayes) to run_<task>_<bayes> and the
    candidate to (temp, sigma).  Only bayes='mfvi' exists here (the other methods are the paper's baselines, out of scope)."""
    if bayes != "mfvi":
        raise ValueError(f"bayes={bayes!r}: only the MFVI runners are implemented")
    run = RUNNERS[_TASK_ALIASES[task]]
    return run(temp=candidate[0], sigma=candidate[1], device=device, **params)


def _trial_value(fn, kwargs, candidate, device) -> float:
    try:
        return float(fn(temp=candidate[0], sigma=candidate[1], device=device, **kwargs))
    except Exception as e:  # a crashed trial is reported as NaN and dropped, like a diverged one
        print(f"[trial {candidate} on {device}] failed: {type(e).__name__}: {e}", flush=True)
        return float("nan")


def _trial_entry(fn, kwargs, candidate, device, queue):
    """Body of one trial process (reference f(), bayesian_optimization.py:3709-3724)."""
    queue.put((tuple(candidate), _trial_value(fn, kwargs, candidate, device)))


def _worker_entry(fn, kwargs, device, conn):
    """Body of one persistent worker: trials of its device, one after the other, until the None sentinel (or the parent's end
    of the pipe closes).  The CUDA context, the loaded library and whatever `fn` caches per process survive from trial to
    trial.  Results go back synchronously over the worker's OWN pipe: a worker that dies cannot leave a lock of a shared queue
    behind."""
    while True:
        try:
            cand = conn.recv()
        except EOFError:
            return
        if cand is None:
            return
        conn.send((tuple(cand), _trial_value(fn, kwargs, cand, device)))


class TrialPool:
    """Persistent trial workers: one process per device slot, alive from construction to close(); `run(candidates)` feeds each
    worker its candidates one after the other and returns {candidate: value} (NaN for a trial that raised or whose worker
    died; a dead worker is replaced).  Keeps `import torch`, the CUDA context and the loaded library across the trials of one
    sweep AND across the rounds of `bo()`.  Use as a context manager."""

    def __init__(self, devices: Sequence[str], fn: Callable[..., float], fn_kwargs: Optional[dict] = None, *,
                 n_workers: Optional[int] = None, start_method: str = "spawn"):
        import torch.multiprocessing as mp
        self._ctx = mp.get_context(start_method)
        self._fn, self._kwargs = fn, dict(fn_kwargs or {})
        self._devices = list(devices)
        self._n = n_workers or len(self._devices)
        self._workers: Dict[int, list] = {}           # wid -> [process, parent end of its pipe, candidate in flight]

    def _spawn(self, wid: int):
        device = self._devices[wid % len(self._devices)]
        parent, child = self._ctx.Pipe(duplex=True)
        pr = self._ctx.Process(target=_worker_entry, args=(self._fn, self._kwargs, device, child))
        pr.start()
        child.close()                                 # the worker holds the only other end: its death reads as EOF here
        self._workers[wid] = [pr, parent, None]

    def run(self, candidates: Iterable[Sequence[float]]) -> Dict[Tuple[float, float], float]:
        from multiprocessing.connection import wait
        cands = [tuple(c) for c in candidates]
        results: Dict[Tuple[float, float], float] = {}
        pending = list(cands)
        workers = self._workers

        def feed(wid):
            w = workers[wid]
            w[2] = pending.pop(0) if pending else None
            if w[2] is not None:
                w[1].send(w[2])

        def replace(wid):                             # the worker died inside its trial: NaN, fresh process for the rest
            w = workers[wid]
            results[w[2]] = float("nan")
            w[1].close()
            w[0].join()
            self._spawn(wid)
            feed(wid)

        for wid in range(min(self._n, len(cands))):
            if wid not in workers or not workers[wid][0].is_alive():
                self._spawn(wid)
            feed(wid)
        while True:
            busy = {w[1]: wid for wid, w in workers.items() if w[2] is not None}
            if not busy:
                break
            for conn in wait(list(busy), timeout=0.5):
                wid = busy[conn]
                try:
                    c, v = conn.recv()
                except (EOFError, OSError):
                    replace(wid)
                    continue
                results[c] = v
                feed(wid)
        return results

    def close(self):
        for w in self._workers.values():
            try:
                w[1].send(None)
            except (OSError, ValueError):
                pass
        for w in self._workers.values():
            w[0].join(timeout=60)
            if w[0].is_alive():
                w[0].terminate()
            w[1].close()
        self._workers = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _finite(cands, results):
    """(candidates, values) in candidate order with the NaN results removed (bayesian_optimization.py:3778-3781)."""
    X, Y = [], []
    for c in cands:
        v = results.get(c, float("nan"))
        if not math.isnan(v):
            X.append(c)
            Y.append(v)
    return X, Y


def eval_trials(candidates: Iterable[Sequence[float]], devices: Sequence[str], fn: Callable[..., float],
                fn_kwargs: Optional[dict] = None, *, max_parallel: Optional[int] = None,
                start_method: str = "spawn", persistent: bool = False) -> Tuple[List[Tuple[float, float]], List[float]]:
    """Run `fn(temp=, sigma=, device=, **fn_kwargs)` once per candidate, devices assigned round-robin; at most `max_parallel`
    (default: one per device) trials run at a time.  Returns (candidates, values) with NaN results removed
    (bayesian_optimization.py:3772-3781).
    persistent=False: one OS process per trial, as the reference starts them.  persistent=True: one worker process per
    device slot that runs its trials back to back -- the process start, `import torch`, the CUDA context and the library load
    (seconds, against a 300-iteration trial's fraction of a second of kernels) are paid once per device instead of once per
    trial; a trial that raises is NaN, a worker that dies takes only its current trial with it and is replaced."""
    import queue as queue_mod
    import torch.multiprocessing as mp
    ctx = mp.get_context(start_method)
    cands = [tuple(c) for c in candidates]
    max_parallel = max_parallel or len(devices)
    if persistent:
        with TrialPool(devices, fn, fn_kwargs, n_workers=max_parallel, start_method=start_method) as pool:
            results = pool.run(cands)
    else:
        queue = ctx.Queue()
        dev_cycle = itertools.cycle(devices)
        results = {}
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
    return _finite(cands, results)


def log_grid(bounds_log10: Sequence[Sequence[float]], n: int) -> List[Tuple[float, float]]:
    """n x n (temp, sigma) candidates on a log10 grid inside `bounds_log10` = [[lo_t, hi_t], [lo_s, hi_s]]
    (bo_configs/bo_mfvi.json: logbounds [-10, 0]^2)."""
    t = np.logspace(bounds_log10[0][0], bounds_log10[0][1], n)
    s = np.logspace(bounds_log10[1][0], bounds_log10[1][1], n)
    return [(float(a), float(b)) for a in t for b in s]
