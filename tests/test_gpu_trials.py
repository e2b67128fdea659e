"""Trial fan-out on real devices (reference f() / bo() / eval(), bayesian_optimization.py:3709-3781,
eval_result.py:19-47): eval_trials starts one OS process per (temp, sigma) candidate, round-robin over the CUDA devices of the
box, each running the real denoising runner on the engine; the objective comes back through the queue and a diverged (NaN)
trial is dropped."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _den_trial(temp, sigma, device, size=64, num_iter=40, poison=False):
    """One trial = the denoising runner on a small phantom (module level: the spawn start method pickles it by name)."""
    from mfvi_dip_mia_b200.runners import run_den_mfvi
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom
    if poison and temp > 1e-3:
        return float("nan")                      # stands in for a diverged run (the reference drops NaN objectives, :3778-3781)
    return run_den_mfvi(ellipse_phantom(size), temp=temp, sigma=sigma, lr=1e-2, num_iter=num_iter, device=device, show_every=10 ** 9)


def test_trials_run_one_process_per_candidate_on_the_gpus():
    from mfvi_dip_mia_b200.runners import eval_trials
    n_dev = torch.cuda.device_count()
    devices = [f"cuda:{i}" for i in range(min(n_dev, 8))]
    cands = [(5.6e-7, 1.5e-5), (1e-8, 1e-4), (1e-2, 1e-3)]           # the last one is reported as NaN and must be dropped
    X, Y = eval_trials(cands, devices, _den_trial, {"poison": True}, max_parallel=max(2, len(devices)))
    assert X == cands[:2], X
    assert all(math.isfinite(v) and 5.0 < v < 40.0 for v in Y), Y     # PSNR (dB) of the smoothed reconstruction
    # the same candidate in this process gives the same objective up to the run-to-run drift of a seeded trajectory (fp32
    # atomics order + chaotic optimisation: ~0.15 dB after 40 steps at lr 1e-2)
    same = _den_trial(*cands[0], devices[0])
    assert abs(same - Y[0]) < 0.5, (same, Y[0])


def test_trials_on_persistent_workers_match_per_trial_processes():
    """persistent=True: one worker process per device runs its trials back to back (context, library and imports paid once);
    the objectives are those of the one-process-per-trial fan-out, the NaN trial is dropped the same way."""
    from mfvi_dip_mia_b200.runners import eval_trials
    n_dev = torch.cuda.device_count()
    devices = [f"cuda:{i}" for i in range(min(n_dev, 8))]
    cands = [(5.6e-7, 1.5e-5), (1e-8, 1e-4), (1e-2, 1e-3), (1e-7, 1e-2), (3e-6, 1e-3)]
    X, Y = eval_trials(cands, devices, _den_trial, {"poison": True}, persistent=True)
    assert X == [c for c in cands if c[0] <= 1e-3], X
    assert all(math.isfinite(v) and 5.0 < v < 40.0 for v in Y), Y
    X1, Y1 = eval_trials(cands[:2], devices, _den_trial, {"poison": True}, max_parallel=max(2, len(devices)))
    assert X1 == cands[:2] and all(abs(a - b) < 0.5 for a, b in zip(Y[:2], Y1)), (Y, Y1)


def test_one_process_drives_two_devices():
    """Engines on cuda:0 and cuda:1 in ONE process: every entry point makes its device current, and the kernels' opt-in to
    large dynamic shared memory is taken per device (cudaFuncSetAttribute applies to the current device only).  The same
    seeded step on the two devices agrees to summation order (fp32 / double atomics, which the BatchNorm chain amplifies in the
    gradient — measured 2.4e-3 relative L2 after three steps): loss terms to 1e-5, gradient to 1e-2, parameters to 1e-4."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import numpy as np
    from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec, _lib as L
    from mfvi_dip_mia_b200.utils.common_utils import get_noise
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom
    img = ellipse_phantom(64)
    res = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        np.random.seed(3)
        torch.manual_seed(3)
        noisy = np.clip(img + np.random.normal(scale=0.1, size=img.shape), 0, 1).astype(np.float32)
        spec = SkipSpec(16, 2)
        tr = MfviDipTrainer(spec, "den", get_noise(spec.num_input_channels, 'noise', (64, 64)), temp=1e-6, sigma=0.1, lr=1e-3,
                            mc_samples=2, seed=3, reg_noise_std=0.1, device=torch.device(dev), target=torch.from_numpy(noisy)[None],
                            math_mode=L.MATH_TF32)
        tr.step()
        torch.cuda.synchronize(torch.device(dev))
        nll, kl, _ = tr.loss_terms()
        res.append((float(nll), float(kl), tr.eng.grad.detach().double().cpu(), tr.eng.theta.detach().double().cpu()))
    for other in (res[1], res[2]):
        assert abs(other[0] - res[0][0]) <= 1e-5 * abs(res[0][0]) and abs(other[1] - res[0][1]) <= 1e-5 * abs(res[0][1]), (other[:2], res[0][:2])
        for k, bar in ((2, 1e-2), (3, 1e-4)):
            rel = float((other[k] - res[0][k]).norm() / res[0][k].norm())
            assert rel < bar, (k, rel)
