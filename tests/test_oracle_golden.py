"""Pins oracle/ (the CPU restatement) against fixtures produced by the imported reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mfvi_oracle as O
from oracle import philox
from tests.helpers import GOLDEN, grad_errs, group, keys_of, load_npz, rel_err

SMALL = {
    "den": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "sr": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "ct": O.SkipCfg(4, 1, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "inp": O.SkipCfg(4, 4, (8, 16, 16), (8, 16, 16), (0, 0, 0), 5, 3, 1, False, False, "nearest"),
}
FULL = {
    "den": O.SkipCfg(16, 2), "sr": O.SkipCfg(32, 2), "ct": O.SkipCfg(16, 1),
    "inp": O.SkipCfg(16, 4, (16, 32, 64, 128, 128, 128), (16, 32, 64, 128, 128, 128), (0,) * 6, 5, 3, 1, False, False,
                     "nearest"),
}


def test_philox_known_answers():
    for c, k, exp in philox.KAT:
        out = philox.philox4x32_10(*[np.array([x], dtype=np.uint32) for x in c], k[0], k[1])
        assert tuple(int(o[0]) for o in out) == exp
    z = philox.philox_normal(1 << 18, 99, 3, 1, 7)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1) < 0.01


@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_layout_matches_reference_keys(task):
    with open(os.path.join(GOLDEN, "full_net_keys.json")) as f:
        ref = json.load(f)[task]
    lay = O.skip_layout(FULL[task])
    ours = set()
    for c in lay.convs_in_exec_order():
        ours |= {c.key + s for s in (".W_mu", ".W_rho", ".bias_mu", ".bias_rho")}
    for sc in lay.scales:
        for b in (sc.skip_bn, sc.d1_bn, sc.d2_bn, sc.cat_bn, sc.up_bn, sc.up1_bn):
            if b is not None:
                ours |= {b + s for s in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")}
    assert ours == set(ref["keys"])
    shapes = dict(zip(ref["keys"], ref["shapes"]))
    for c in lay.convs_in_exec_order():
        assert shapes[c.key + ".W_mu"] == [c.cout, c.cin, c.k, c.k]


@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_small_step_matches_reference(task):
    d = load_npz(f"skipnet_small_{task}.npz")
    cfg = SMALL[task]
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in group(d, "sd/").items()}
    S = int(d["S"])
    eps = [group(d, f"eps{s}/") for s in range(S)]
    ex = group(d, "extra/")
    temp, sigma = float(d["temp"]), float(d["sigma"])
    kw = dict(task=task, temp=temp, prior_sigma_plus_eps=O.prior_scale(temp, sigma))
    if task in ("den", "sr"):
        kw["target"] = ex["target"]
    if task == "inp":
        kw.update(target=ex["target"], mask=ex["mask"])
    if task == "ct":
        kw.update(theta_deg=ex["theta"], sino=ex["sino"])
    loss, nll, kl, outs = O.mfvi_loss(sd, cfg, torch.from_numpy(d["net_input"]), eps, **kw)
    loss.backward()
    assert rel_err(nll, d["nll"]) < 2e-6
    assert rel_err(kl, d["kl"]) < 2e-6
    assert rel_err(loss, d["loss"]) < 2e-6
    for s in range(S):
        assert rel_err(outs[s], d[f"out{s}"]) < 2e-5
    grads = group(d, "grad/")
    assert len(grads) > 0
    errs = grad_errs({k: sd[k].grad for k in grads}, grads)
    worst = max(errs, key=errs.get)
    assert errs[worst] < 5e-4, (worst, errs[worst])


def test_conv_linear_layers_match_reference():
    d = load_npz("layers.npz")
    for name in ["c1x1", "c3s1", "c3s2", "c5s2", "c5s1", "c3s1_odd"]:
        g = group(d, name + "/")
        st = int(g["meta"][3])
        leaves = {k: g[k].clone().requires_grad_(True) for k in ["x", "W_mu", "W_rho", "bias_mu", "bias_rho"]}
        y = O.conv2d_rt(leaves["x"], leaves["W_mu"], leaves["W_rho"], leaves["bias_mu"], leaves["bias_rho"], g["eps_w"],
                        g["eps_b"], stride=st)
        assert rel_err(y, g["y"]) < 1e-6
        y.backward(g["dy"])
        for k in ["x", "W_mu", "W_rho", "bias_mu", "bias_rho"]:
            assert rel_err(leaves[k].grad, g["d" + k]) < 1e-5, (name, k)
        y_eval = O.conv2d_rt(g["x"], g["W_mu"], g["W_rho"], g["bias_mu"], g["bias_rho"], None, None, stride=st,
                             training=False)
        assert rel_err(y_eval, g["y_eval"]) < 1e-6
    g = group(d, "lin/")
    y = O.linear_rt(g["x"], g["W_mu"], g["W_rho"], g["bias_mu"], g["bias_rho"], g["eps_w"], g["eps_b"])
    assert rel_err(y, g["y"]) < 1e-6


def test_kl_matches_reference():
    d = load_npz("layers.npz")
    for i in range(4):
        g = group(d, f"kl{i}/")
        temp, sigma, rev = [float(x) for x in g["meta"]]
        leaves = [g[k].clone().requires_grad_(True) for k in ["W_mu", "W_rho", "bias_mu", "bias_rho"]]
        kl = O.kl_layers(leaves, 0.0, O.prior_scale(temp, sigma), "reverse" if rev else "forward")
        assert rel_err(kl, g["kl"]) < 2e-6, i
        kl.backward()
        for leaf, k in zip(leaves, ["W_mu", "W_rho", "bias_mu", "bias_rho"]):
            assert rel_err(leaf.grad, g["d" + k]) < 1e-5, (i, k)


def test_nll_radon_metrics_match_reference():
    d = load_npz("layers.npz")
    g = group(d, "nll/")
    assert rel_err(O.gaussian_nll(g["mu"], g["s"], g["t"]), g["v"]) < 1e-6
    g = group(d, "nlli/")
    assert rel_err(O.gaussian_nll_inpainting(torch.sigmoid(g["mu"]), g["s"], g["t"], g["m"]), g["v"]) < 1e-6
    for name in ["r32", "r48", "r40c2"]:
        g = group(d, name + "/")
        img = g["img"].clone().requires_grad_(True)
        sino = O.radon_forward(img, g["theta"])
        assert sino.shape == g["sino"].shape
        assert rel_err(sino, g["sino"]) < 1e-5, name
        sino.backward(g["dsino"])
        assert rel_err(img.grad, g["dimg"]) < 1e-5, name
    g = group(d, "met/")
    assert abs(O.psnr(g["a"], g["b"]) - float(g["psnr"])) < 1e-4
    assert abs(O.ssim(g["a"], g["b"]) - float(g["ssim"])) < 1e-5
    assert abs(O.uce(g["err"], g["unc"]) - float(g["uce"])) < 1e-7


@pytest.mark.parametrize("i", [0, 1, 2])
def test_metrics_match_reference(i):
    """PSNR / SSIM / UCE (the trajectory-parity metrics and the device-side bookkeeping's checkers) against the imported
    reference's utils/common_utils.py:297-353 and utils/uce.py (tests/golden/metrics.npz); the package's host-side
    restatements (mfvi_dip_mia_b200/utils) are held to the same vectors."""
    from mfvi_dip_mia_b200.utils.common_utils import peak_signal_noise_ratio, structural_similarity
    from mfvi_dip_mia_b200.utils.uce import uceloss
    g = group(load_npz("metrics.npz"), f"m{i}/")
    a, b = g["a"], g["b"]
    assert abs(O.psnr(a, b) - float(g["psnr"])) < 1e-4
    assert abs(O.ssim(a, b) - float(g["ssim"])) < 1e-6
    assert abs(O.uce(g["err"], g["unc"], 15) - float(g["uce"])) < 1e-7
    assert abs(peak_signal_noise_ratio(a, b) - float(g["psnr"])) < 1e-4
    assert abs(structural_similarity(a, b) - float(g["ssim"])) < 1e-6
    assert abs(float(uceloss(g["err"], g["unc"], n_bins=15)[0]) - float(g["uce"])) < 1e-7


@pytest.mark.parametrize("name", ["l1x1", "l3s1p1", "l3s2", "l5s1nb", "lin"])
def test_lrt_layers_match_reference(name):
    """Local-reparameterisation layers (row f3): oracle.conv2d_lrt / linear_lrt against the imported reference
    (tests/golden/lrt_layers.npz: Conv2dLRT / LinearLRT with the output-space eps injected), forward, backward, eval."""
    g = group(load_npz("lrt_layers.npz"), name + "/")
    has_bias = "bias_mu" in g
    leaves = {k: g[k].clone().requires_grad_(True) for k in ["x", "W_mu", "W_rho"] + (["bias_mu", "bias_rho"] if has_bias else [])}
    bm, br = (leaves["bias_mu"], leaves["bias_rho"]) if has_bias else (None, None)
    if name == "lin":
        y = O.linear_lrt(leaves["x"], leaves["W_mu"], leaves["W_rho"], bm, br, g["eps"])
        y_eval = O.linear_lrt(g["x"], g["W_mu"], g["W_rho"], g.get("bias_mu"), g.get("bias_rho"), None, training=False)
    else:
        cin, cout, k, st, pad = [int(v) for v in g["meta"][:5]]
        y = O.conv2d_lrt(leaves["x"], leaves["W_mu"], leaves["W_rho"], bm, br, g["eps"], stride=st, padding=pad)
        y_eval = O.conv2d_lrt(g["x"], g["W_mu"], g["W_rho"], g.get("bias_mu"), g.get("bias_rho"), None, stride=st, padding=pad,
                              training=False)
    assert rel_err(y, g["y"]) < 1e-6 and rel_err(y_eval, g["y_eval"]) < 1e-6
    y.backward(g["dy"])
    for k in leaves:
        assert rel_err(leaves[k].grad, g["d" + k]) < 1e-5, k
