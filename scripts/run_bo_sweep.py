"""BASELINE config 5 (bo_configs/bo_mfvi.json): a sweep of independent (temp, sigma) trials over the GPUs of the box — replicas
only, no communication.  Default: one persistent worker process per GPU that runs its trials back to
back; --per-trial-process starts one OS process per trial as the reference does (process start + CUDA context per trial).
    python scripts/run_bo_sweep.py [--grid 8] [--num-iter 300] [--size 256] [--per-trial-process] [--out gpurun_out/r02_bo_sweep.json]
Prints one JSON line: trials/hour, wall time, trials per device, the best candidate."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def trial(temp, sigma, device, size, num_iter):
    import torch
    from mfvi_dip_mia_b200.runners import run_den_mfvi
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom
    t0 = time.time()
    v = run_den_mfvi(ellipse_phantom(size), temp=temp, sigma=sigma, lr=2e-3, num_iter=num_iter, seed=1, device=device,
                     show_every=10 ** 9)
    torch.cuda.synchronize()
    print(f"[trial temp={temp:.2e} sigma={sigma:.2e} {device}] psnr_gt_sm {v:.3f} dB in {time.time() - t0:.1f} s", flush=True)
    return v


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=8)
    ap.add_argument("--num-iter", type=int, default=300)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--out", default="gpurun_out/r02_bo_sweep.json")
    ap.add_argument("--per-trial-process", action="store_true")
    a = ap.parse_args()
    import torch
    from mfvi_dip_mia_b200.runners import eval_trials, log_grid
    n_dev = torch.cuda.device_count()
    devices = [f"cuda:{i}" for i in range(n_dev)]
    cands = log_grid([[-10, 0], [-10, 0]], a.grid)                 # bo_configs/bo_mfvi.json: logbounds [-10, 0]^2
    t0 = time.time()
    X, Y = eval_trials(cands, devices, trial, {"size": a.size, "num_iter": a.num_iter}, persistent=not a.per_trial_process)
    wall = time.time() - t0
    best = max(range(len(Y)), key=lambda i: Y[i]) if Y else None
    line = {"config": "bo_mfvi sweep", "fan_out": "process per trial" if a.per_trial_process else "persistent worker per GPU", "trials": len(cands), "finished": len(Y), "dropped_nan": len(cands) - len(Y), "n_gpus": n_dev,
            "num_iter": a.num_iter, "size": a.size, "wall_s": wall, "trials_per_hour": 3600.0 * len(cands) / wall,
            "steps_per_s_aggregate": len(cands) * (a.num_iter + 1) / wall,
            "best": None if best is None else {"temp": X[best][0], "sigma": X[best][1], "psnr_gt_sm": Y[best]}}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(dict(line, X=X, Y=Y), f)
    print(json.dumps(line))
