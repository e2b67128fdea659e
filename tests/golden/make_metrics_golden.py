"""Generates tests/golden/metrics.npz: PSNR / SSIM / UCE of the IMPORTED REFERENCE (utils/common_utils.py:297-353,
utils/uce.py:9-40) on seeded random images — the pin for oracle.psnr / ssim / uce, which in turn check the device-side
bookkeeping kernels (tests/test_gpu_parity.py).  Build container only.

    python tests/golden/make_metrics_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (imports the reference)

R = G.R


def main():
    g = torch.Generator().manual_seed(31)
    arrs = {}
    for i, (H, W) in enumerate([(64, 64), (40, 72), (33, 17)]):
        a = torch.rand(1, 1, H, W, generator=g)
        b = (a + 0.1 * torch.randn(1, 1, H, W, generator=g)).clamp(0, 1)
        arrs[f"m{i}/a"], arrs[f"m{i}/b"] = a, b
        arrs[f"m{i}/psnr"] = np.float64(R["psnr"](a, b))
        arrs[f"m{i}/ssim"] = np.float64(R["ssim"](a, b))
        err = (a - b).reshape(-1) ** 2
        unc = (0.01 * torch.rand(H * W, generator=g) + 0.5 * err)
        u, _, _, _ = R["uce"](err, unc, n_bins=15)
        arrs[f"m{i}/err"], arrs[f"m{i}/unc"] = err, unc
        arrs[f"m{i}/uce"] = np.float64(float(u))
    G.save("metrics.npz", **arrs)


if __name__ == "__main__":
    main()
