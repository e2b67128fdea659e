"""Generates tests/golden/trajectory_den64.npz: the distribution of the FINAL PSNR / SSIM / UCE of the denoising
runner's loop (reference bayesian_optimization.py:1360-1416, UCE recipe eval_denoising.ipynb:467-482) run with the
IMPORTED REFERENCE (read-only /root/reference: its MeanFieldVI, skip net, gaussian_nll, metrics, torch's own RNG) over
K seeds on the 64x64 ellipse phantom.  Runs only in the build container.

    python tests/golden/make_trajectory_golden.py [K] [n_it]

Why a distribution: the optimisation is chaotic — the same reference arithmetic in fp32 vs fp64, or with eps perturbed
by 2e-6, ends 0.6-0.8 dB apart after 2400 steps (measured) — so a single trajectory cannot be
matched to 0.1 dB by anything, including the reference itself.  The GPU test compares ensemble means.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

H = W = 64
TEMP, SIGMA, LR = 5.656911698337764e-07, 1.4616642493692077e-05, 1e-2
RING, EXPW, REG = 25, 0.99, 0.1
CHECK = (300, 600, 900, 1200)


def one_seed(seed, n_it):
    torch.set_num_threads(2)
    import make_golden as G                      # imports the reference (stubs for matplotlib / skimage)
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    R, O = G.R, G.O
    cfg = O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear")
    torch.manual_seed(seed)
    np.random.seed(seed)
    gt = torch.from_numpy(ellipse_phantom(H))[None]
    tgt = torch.from_numpy(noisy(ellipse_phantom(H), 0.1, 1))[None]
    net = G.build_ref_net(cfg, float(np.sqrt(TEMP) * SIGMA))
    net.train()
    x_saved = torch.rand(1, cfg.num_input_channels, H, W) * 0.1      # get_noise 'u', var 1/10 (common_utils.py:134-162)
    noise = x_saved.clone()
    opt = torch.optim.AdamW(net.parameters(), lr=LR, weight_decay=0)
    out_avg = None
    ring_epi = torch.zeros(RING, 1, H, W)
    ring_ale = torch.zeros(RING, 1, H, W)
    res = {}
    for i in range(n_it):
        opt.zero_grad()
        x = x_saved + noise.normal_() * REG
        out = net(x)
        nll = R["nll"](out[:, :1], out[:, 1:], tgt)
        loss = nll + TEMP * net.kl()
        loss.backward()
        opt.step()
        with torch.no_grad():
            out[:, 1:] = torch.exp(-out[:, 1:])
            out_avg = out.detach() if out_avg is None else out_avg * EXPW + out.detach() * (1 - EXPW)
            ring_epi[i % RING] = out.detach()[0, :1].clip(0, 1)
            ring_ale[i % RING] = out.detach()[0, 1:].clip(0, 1)
            if (i + 1) in CHECK or i == n_it - 1:
                sm = out_avg[:, :1].clip(0, 1)
                unc = ring_epi.var(0) + ring_ale.mean(0)
                err2 = ((ring_epi - gt) ** 2).mean(0)
                uce = float(R["uce"](err2.reshape(-1), unc.reshape(-1), n_bins=15)[0])
                res[i + 1] = (float(R["psnr"](gt, sm)), float(R["ssim"](gt, sm)), uce)
    return res


def _worker(args):
    return one_seed(*args)


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
    t0 = time.time()
    with mp.get_context("spawn").Pool(4) as pool:
        runs = pool.map(_worker, [(100 + k, n_it) for k in range(K)])
    its = sorted(runs[0])
    arr = np.array([[r[i] for i in its] for r in runs])              # (K, checkpoints, 3)
    np.savez(os.path.join(HERE, "trajectory_den64.npz"), its=np.array(its), metrics=arr, seeds=np.arange(100, 100 + K),
             hyper=np.array([TEMP, SIGMA, LR, RING, EXPW, REG, H]))
    for j, i in enumerate(its):
        print(i, "mean", arr[:, j].mean(0), "std", arr[:, j].std(0, ddof=1))
    print("wall", time.time() - t0)
