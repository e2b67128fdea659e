"""Where one BO trial's wall time goes (config 5): imports, CUDA context, trainer/plan build, graph capture, loop.
    python scripts/trial_setup_profile.py [num_iter] [size]"""
import os, sys, time
t00 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
t_torch = time.time()
import numpy as np
from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec, _lib as L
from mfvi_dip_mia_b200.runners import DeviceBookkeeping, _run_loop
from mfvi_dip_mia_b200.utils.common_utils import get_noise
from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom
t_pkg = time.time()
num_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 300
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.zeros(1, device="cuda:0"); torch.cuda.synchronize()
t_ctx = time.time()
print(f"import torch {t_torch - t00:.2f} s | package + library {t_pkg - t_torch:.2f} s | CUDA context {t_ctx - t_pkg:.2f} s")
for rep in range(3):
    t0 = time.time()
    img = ellipse_phantom(size)
    np.random.seed(1); torch.manual_seed(1)
    noisy = np.clip(img + np.random.normal(scale=0.1, size=img.shape), 0, 1).astype(np.float32)
    spec = SkipSpec(16, 2)
    net_input = get_noise(spec.num_input_channels, 'noise', (size, size))
    t1 = time.time()
    tr = MfviDipTrainer(spec, "den", net_input, temp=1e-6, sigma=0.1, lr=2e-3, mc_samples=1, seed=1, reg_noise_std=0.1,
                        device=torch.device("cuda:0"), target=torch.from_numpy(noisy)[None], math_mode=L.MATH_TF32)
    torch.cuda.synchronize(); t2 = time.time()
    bk = DeviceBookkeeping(tr, gt=img, noisy=noisy, exp_weight=0.99, ring=25)
    torch.cuda.synchronize(); t3 = time.time()
    tr.step(); torch.cuda.synchronize(); t4 = time.time()
    tr.step(); torch.cuda.synchronize(); t5 = time.time()
    tr.step(); torch.cuda.synchronize(); t6 = time.time()
    v = _run_loop(tr, bk, num_iter, 10 ** 9, False)
    torch.cuda.synchronize(); t7 = time.time()
    print(f"[rep {rep}] data {t1 - t0:.2f} | trainer {t2 - t1:.2f} | bookkeeping {t3 - t2:.2f} | step1 {t4 - t3:.3f} step2 {t5 - t4:.3f} "
          f"step3 {t6 - t5:.3f} | loop of {num_iter + 1} {t7 - t6:.2f} s | total {t7 - t0:.2f} s  psnr {v:.2f}")
    del tr, bk
