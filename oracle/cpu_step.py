"""ORACLE — test infrastructure only.  The reference training step (bayesian_optimization.py:1361-1372) restated on
the CPU with PyTorch fp32, used by bench.py's `cpu_baseline` leg and `--impl reference` arm and by tests.  MC>1 is
the restatement of SURVEY §4: S sequential forwards on the same net_input, NLL averaged, one backward, AdamW.
"""
from __future__ import annotations

import math
import time

import torch

from . import mfvi_oracle as O


class OracleStepper:
    def __init__(self, cfg: O.SkipCfg, H: int, W: int, *, mc_samples: int, temp: float, sigma: float, lr: float,
                 seed: int = 1, task: str = "den", threads: int | None = None, device="cpu", head: dict | None = None):
        """`head`: the task's data (target / mask / theta_deg + sino) as oracle.mfvi_loss takes it; default = a random
        denoising target.  `device`: "cpu" (the reference's CPU path) or a CUDA device (the same eager PyTorch ops there — the
        GPU bar of SURVEY section 2a; TF32 is switched off by the caller)."""
        if threads:
            torch.set_num_threads(threads)
        self.cfg, self.S, self.temp, self.task = cfg, mc_samples, temp, task
        self.device = torch.device(device)
        self.prior = O.prior_scale(temp, sigma)
        g = torch.Generator().manual_seed(seed)
        self.g = g
        lay = O.skip_layout(cfg)
        self.lay = lay
        sd = {}
        for c in lay.convs_in_exec_order():
            shp = (c.cout, c.cin, c.k, c.k)
            sd[c.key + ".W_mu"] = (0.1 * torch.randn(shp, generator=g)).requires_grad_(True)
            sd[c.key + ".W_rho"] = (-3 + 0.1 * torch.randn(shp, generator=g)).requires_grad_(True)
            sd[c.key + ".bias_mu"] = (0.1 * torch.randn(c.cout, generator=g)).requires_grad_(True)
            sd[c.key + ".bias_rho"] = (-3 + 0.1 * torch.randn(c.cout, generator=g)).requires_grad_(True)
        for sc in lay.scales:
            for b, ch in ((sc.skip_bn, sc.skip_conv.cout if sc.skip_conv else 0), (sc.d1_bn, sc.d1.cout),
                          (sc.d2_bn, sc.d2.cout), (sc.cat_bn, sc.up.cin), (sc.up_bn, sc.up.cout),
                          (sc.up1_bn, sc.up1.cout if sc.up1 else 0)):
                if b is not None:
                    sd[b + ".weight"] = torch.ones(ch, requires_grad=True)
                    sd[b + ".bias"] = torch.zeros(ch, requires_grad=True)
        if self.device.type != "cpu":
            sd = {k: v.detach().to(self.device).requires_grad_(True) for k, v in sd.items()}
        self.sd = sd
        self.saved = (torch.rand(1, cfg.num_input_channels, H, W, generator=g) * 0.1).to(self.device)
        if head is None:
            head = {"target": torch.rand(1, 1, H, W, generator=g)}
        self.head = {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in head.items()}
        self.gd = torch.Generator(device=self.device).manual_seed(seed) if self.device.type != "cpu" else g
        self.opt = torch.optim.AdamW(list(sd.values()), lr=lr, weight_decay=0)

    def step(self, mc_samples: int | None = None) -> float:
        S = mc_samples or self.S
        self.opt.zero_grad()
        rn = lambda shape: torch.randn(shape, generator=self.gd, device=self.device)
        x = self.saved + 0.1 * rn(self.saved.shape)
        eps = []
        for _ in range(S):
            e = {}
            for c in self.lay.convs_in_exec_order():
                e[c.key + ".W"] = rn(self.sd[c.key + ".W_mu"].shape)
                e[c.key + ".b"] = rn((c.cout,))
            eps.append(e)
        loss, nll, kl, _ = O.mfvi_loss(self.sd, self.cfg, x, eps, task=self.task, temp=self.temp,
                                       prior_sigma_plus_eps=self.prior, **self.head)
        loss.backward()
        self.opt.step()
        return float(loss.detach())


def time_steps(stepper: OracleStepper, steps: int, warmup: int, mc_samples: int | None = None):
    for _ in range(warmup):
        stepper.step(mc_samples)
    if stepper.device.type != "cpu":
        torch.cuda.synchronize(stepper.device)
    t0 = time.perf_counter()
    for _ in range(steps):
        stepper.step(mc_samples)
    if stepper.device.type != "cpu":
        torch.cuda.synchronize(stepper.device)
    return (time.perf_counter() - t0) / max(steps, 1)
