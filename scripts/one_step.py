"""Two eager MFVI-DIP steps (metric shape: 256^2, MC=8, tf32) — a short driver for `ncu --set full -k regex:<kernel>`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_problem, TEMP, SIGMA, LR
from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec, _lib as L

x, t = synthetic_problem(256)
tr = MfviDipTrainer(SkipSpec(), "den", x, temp=TEMP, sigma=SIGMA, lr=LR, mc_samples=8, seed=1, device="cuda:0", target=t,
                    math_mode=L.MATH_TF32, use_graph=False)
for _ in range(2):
    tr.step()
torch.cuda.synchronize()
print("ok", tr.loss_terms())
