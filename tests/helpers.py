"""Shared helpers for the test-suite (fixture loading, error metrics)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def group(d, prefix, as_torch=True, dtype=None):
    out = {}
    for k, v in d.items():
        if k.startswith(prefix):
            t = torch.from_numpy(np.asarray(v)) if as_torch else v
            if as_torch and dtype is not None and t.is_floating_point():
                t = t.to(dtype)
            out[k[len(prefix):]] = t
    return out


def keys_of(d):
    return json.loads(str(d["keys"]))


def rel_err(a, b):
    """max|a-b| / max|b| — the normalised max error used for every fp parity bar in this repo."""
    a = torch.as_tensor(a).detach().double().reshape(-1)
    b = torch.as_tensor(b).detach().double().reshape(-1)
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def grad_errs(ours, ref, floor_frac=1e-3, bias_frac=0.05):
    """Per-tensor normalised max error for a dict of gradients: max|a-b| / den_k.
    den_k = max|ref_k|, floored at floor_frac * (largest max|ref| over all tensors).
    Conv biases (`bias_mu`/`bias_rho`) that feed a BatchNorm have a mathematically ZERO data gradient
    (BN removes the mean): their reference values are the tiny T*dKL term plus fp32 cancellation noise
    of a sum over all pixels, so their denominator is additionally floored at bias_frac * max|grad of
    the same layer's W_mu / W_rho| (two fp32 CPU implementations of the same math already differ
    by ~1e-5 of the layer's weight-gradient scale there)."""
    gmax = max(float(torch.as_tensor(v).abs().max()) for v in ref.values())
    out = {}
    for k, b in ref.items():
        a = torch.as_tensor(ours[k]).detach().double().reshape(-1)
        b = torch.as_tensor(b).double().reshape(-1)
        den = max(float(b.abs().max()), floor_frac * gmax)
        if k.endswith("bias_mu") or k.endswith("bias_rho"):
            wk = k.replace("bias_mu", "W_mu").replace("bias_rho", "W_rho")
            if wk in ref:
                den = max(den, bias_frac * float(torch.as_tensor(ref[wk]).abs().max()))
        out[k] = float((a - b).abs().max()) / den
    return out


def fake_trial(temp, sigma, device, offset=0.0):
    """Stand-in for a trial runner in the CPU test of the trial fan-out: objective = a smooth function of the candidate;
    one candidate diverges (NaN) and one crashes."""
    import math
    if temp == 3.0:
        return float("nan")
    if temp == 4.0:
        raise RuntimeError("boom")
    if temp == 6.0:
        import os
        os._exit(3)                       # the process dies without reporting
    return offset + math.log10(temp) - 2.0 * math.log10(sigma) + (0.0 if device == "cpu:0" else 0.5)


def pid_trial(temp, sigma, device):
    """Stand-in trial whose value is the id of the process that ran it (worker reuse in the TrialPool test)."""
    import os
    return float(os.getpid())


def quadratic_trial(temp, sigma, device, peak=(1e-6, 1e-3)):
    """Analytic stand-in for a trial runner in the BO test: a smooth PSNR-like bowl in log10 space, maximum at `peak`."""
    import math
    return 30.0 - 0.6 * (math.log10(temp) - math.log10(peak[0])) ** 2 - 0.9 * (math.log10(sigma) - math.log10(peak[1])) ** 2
