#!/usr/bin/env python
"""Does MFVI-DIP training survive bf16 conv operands?  A CPU study ahead of the GPU bring-up of the bf16-operand mode
 — test infrastructure (it runs the oracle), not part of the package or of bench.py.

The denoising loop of tests/golden/make_trajectory_golden.py (64x64 phantom, 3-scale net, lr 1e-2, 25-deep rings, EMA 0.99)
is run with the CPU oracle twice per seed on IDENTICAL random streams: once in fp32, once with every convolution emulating the
bf16 mode exactly as the engine defines it — x, w and dy rounded to bf16 (nearest-even), products and sums in fp32, outputs,
input gradients, statistics and parameters in fp32.  Reported: PSNR / SSIM / UCE of the EMA output at the checkpoints, as
mean over seeds, and the PAIRED difference bf16 - fp32 with its standard error (the optimisation is chaotic, so single
trajectories differ by ~0.5 dB; the ensemble mean is what north_star's 0.1 dB / 0.005 bar can be held to).

    python tests/studies/bf16_quality_study.py [K seeds = 8] [n_it = 1200] [workers = 4] [size = 64] [net = small | metric] [arm = bf16 | bf16s] [lr = 1e-2]

`arm = bf16s` ("bf16 storage") goes one step further than the mode that is written: the convolution OUTPUTS and the input
gradients the data-gradient kernels write are rounded to bf16 as well, i.e. every activation and activation gradient that
touches HBM is bf16 and only accumulators, statistics, parameters and their gradients are fp32 — the variant that would halve the
traffic of the elementwise kernels too.

`net = metric` is the 5-scale, 16-channel-input network of the metric configuration (test_configs/mfvi_den.json) instead of the
3-scale network of the trajectory fixture.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

H = W = 64
TEMP, SIGMA, LR = 5.656911698337764e-07, 1.4616642493692077e-05, 1e-2
RING, EXPW, REG = 25, 0.99, 0.1
CHECK = (300, 600, 900, 1200)


class Bf16OperandConv(torch.autograd.Function):
    """conv2d with bf16 operands and fp32 accumulation, forward and both gradients (dy is a bf16 operand of dgrad / wgrad).
    round_io: the stored results (y, dx) are rounded to bf16 too."""
    round_io = False

    @staticmethod
    def forward(ctx, x, w, b, stride):
        xb, wb = x.bfloat16().float(), w.bfloat16().float()
        ctx.save_for_backward(xb, wb)
        ctx.stride, ctx.has_b = stride, b is not None
        y = F.conv2d(xb, wb, b, stride=stride)
        return y.bfloat16().float() if Bf16OperandConv.round_io else y

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        dyb = dy.bfloat16().float()
        dx = torch.nn.grad.conv2d_input(xb.shape, wb, dyb, stride=ctx.stride)
        if Bf16OperandConv.round_io:
            dx = dx.bfloat16().float()
        dw = torch.nn.grad.conv2d_weight(xb, wb.shape, dyb, stride=ctx.stride)
        return dx, dw, (dy.sum((0, 2, 3)) if ctx.has_b else None), None


def one_run(seed, n_it, bf16, size=64, net="small", lr=LR):
    global H, W, LR
    H = W = size
    LR = lr
    torch.set_num_threads(2)
    from oracle import mfvi_oracle as O
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    from mfvi_dip_mia_b200.utils.uce import uceloss
    if bf16:
        Bf16OperandConv.round_io = bf16 == 2
        def conv2d_rt(x, W_mu, W_rho, bias_mu, bias_rho, eps_w, eps_b, stride=1, padding=0, training=True):
            assert padding == 0 and training
            w = O.rsample(W_mu, O.softplus(W_rho), eps_w)
            b = O.rsample(bias_mu, O.softplus(bias_rho), eps_b) if bias_mu is not None else None
            return Bf16OperandConv.apply(x, w, b, stride)
        O.conv2d_rt = conv2d_rt
    cfg = O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear") if net == "small" else O.SkipCfg(16, 2)
    g = torch.Generator().manual_seed(seed)
    gt = torch.from_numpy(ellipse_phantom(H))[None]
    tgt = torch.from_numpy(noisy(ellipse_phantom(H), 0.1, 1))[None]
    lay = O.skip_layout(cfg)
    sd = {}
    for c in lay.convs_in_exec_order():
        shp = (c.cout, c.cin, c.k, c.k)
        sd[c.key + ".W_mu"] = (0.1 * torch.randn(shp, generator=g)).requires_grad_(True)
        sd[c.key + ".W_rho"] = (-3 + 0.1 * torch.randn(shp, generator=g)).requires_grad_(True)
        sd[c.key + ".bias_mu"] = (0.1 * torch.randn(c.cout, generator=g)).requires_grad_(True)
        sd[c.key + ".bias_rho"] = (-3 + 0.1 * torch.randn(c.cout, generator=g)).requires_grad_(True)
    for sc in lay.scales:
        for b, ch in ((sc.skip_bn, sc.skip_conv.cout if sc.skip_conv else 0), (sc.d1_bn, sc.d1.cout), (sc.d2_bn, sc.d2.cout),
                      (sc.cat_bn, sc.up.cin), (sc.up_bn, sc.up.cout), (sc.up1_bn, sc.up1.cout if sc.up1 else 0)):
            if b is not None:
                sd[b + ".weight"] = torch.ones(ch, requires_grad=True)
                sd[b + ".bias"] = torch.zeros(ch, requires_grad=True)
    saved = torch.rand(1, cfg.num_input_channels, H, W, generator=g) * 0.1
    opt = torch.optim.AdamW(list(sd.values()), lr=LR, weight_decay=0)
    prior = O.prior_scale(TEMP, SIGMA)
    out_avg = None
    ring_epi, ring_ale = torch.zeros(RING, 1, H, W), torch.zeros(RING, 1, H, W)
    res = {}
    for i in range(n_it):
        opt.zero_grad()
        x = saved + REG * torch.randn(saved.shape, generator=g)
        eps = {}
        for c in lay.convs_in_exec_order():
            eps[c.key + ".W"] = torch.randn(sd[c.key + ".W_mu"].shape, generator=g)
            eps[c.key + ".b"] = torch.randn(c.cout, generator=g)
        loss, _, _, outs = O.mfvi_loss(sd, cfg, x, [eps], task="den", temp=TEMP, prior_sigma_plus_eps=prior, target=tgt)
        loss.backward()
        opt.step()
        with torch.no_grad():
            out = outs[0].detach().clone()
            out[:, 1:] = torch.exp(-out[:, 1:])
            out_avg = out if out_avg is None else out_avg * EXPW + out * (1 - EXPW)
            ring_epi[i % RING] = out[0, :1].clip(0, 1)
            ring_ale[i % RING] = out[0, 1:].clip(0, 1)
            if (i + 1) in CHECK or i == n_it - 1:
                sm = out_avg[:, :1].clip(0, 1)
                unc = ring_epi.var(0) + ring_ale.mean(0)
                err2 = ((ring_epi - gt) ** 2).mean(0)
                res[i + 1] = (float(O.psnr(gt, sm)), float(O.ssim(gt, sm)), float(uceloss(err2.reshape(-1), unc.reshape(-1), n_bins=15)[0]))
    return res


def _worker(a):
    return one_run(*a)


if __name__ == "__main__":
    import torch.multiprocessing as mp
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    size = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    net = sys.argv[5] if len(sys.argv) > 5 else "small"
    arm = sys.argv[6] if len(sys.argv) > 6 else "bf16"
    lr = float(sys.argv[7]) if len(sys.argv) > 7 else LR
    t0 = time.time()
    jobs = [(100 + k, n_it, b, size, net, lr) for k in range(K) for b in (0, 2 if arm == "bf16s" else 1)]
    with mp.get_context("spawn").Pool(workers) as pool:
        runs = pool.map(_worker, jobs)
    its = sorted(runs[0])
    arr = np.array([[r[i] for i in its] for r in runs]).reshape(K, 2, len(its), 3)        # (seed, arm, checkpoint, metric)
    print(f"# {K} seeds x {n_it} iterations, lr {lr:g}, {size}x{size} denoising, {net} net, oracle fp32 vs emulated {arm} "
          f"({'bf16 conv operands AND bf16 conv outputs / input gradients' if arm == 'bf16s' else 'bf16 conv operands'}); "
          f"wall {time.time() - t0:.0f} s")
    print("# it   arm    PSNR dB   SSIM     UCE      | paired difference bf16 - fp32 (mean +- standard error)")
    for j, i in enumerate(its):
        d = arr[:, 1, j] - arr[:, 0, j]
        se = d.std(0, ddof=1) / np.sqrt(K) if K > 1 else np.zeros(3)
        for a, name in ((0, "fp32"), (1, arm)):
            m = arr[:, a, j].mean(0)
            tail = f" | {d.mean(0)[0]:+.3f}+-{se[0]:.3f} dB  {d.mean(0)[1]:+.4f}+-{se[1]:.4f}  {d.mean(0)[2]:+.4f}+-{se[2]:.4f}" if a else ""
            print(f"{i:5d}  {name}  {m[0]:8.3f}  {m[1]:.4f}  {m[2]:.4f}{tail}")
