"""GPU parity tests: the CUDA path (through the C ABI) against the oracle / the golden fixtures generated from the
imported reference (tests/golden/make_golden.py).  Tolerance: north_star's rtol 1e-3 in fp32, measured as the
normalised max error max|a-b|/max|b| (tests/helpers.py); most checks are far tighter and say so.
"""
import json

import numpy as np
import pytest
import torch

from oracle import mfvi_oracle as O
from oracle import philox
from tests.helpers import grad_errs, group, load_npz, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-3

SMALL = {
    "den": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "sr": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "ct": O.SkipCfg(4, 1, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "inp": O.SkipCfg(4, 4, (8, 16, 16), (8, 16, 16), (0, 0, 0), 5, 3, 1, False, False, "nearest"),
}


def spec_of(cfg):
    from mfvi_dip_mia_b200 import SkipSpec
    return SkipSpec(cfg.num_input_channels, cfg.num_output_channels, tuple(cfg.down), tuple(cfg.up), tuple(cfg.skip),
                    cfg.filter_down, cfg.filter_up, cfg.filter_skip, cfg.need1x1_up, cfg.need_sigmoid, cfg.upsample_mode)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


# ----------------------------------------------------------------------------------------------- RNG
def test_philox_bit_exact_and_normals(dev):
    from mfvi_dip_mia_b200 import _lib as L
    n = 4099
    raw = torch.empty(n, dtype=torch.int32, device=dev)
    L.call("mfvi_philox_raw_fill", raw.data_ptr(), n, L.key(0x1234567890ABCDEF, 7, 3), 5)
    ref = philox.philox_raw(n, 0x1234567890ABCDEF, 5, 3, 7).reshape(-1)[:n]
    assert np.array_equal(raw.cpu().numpy().view(np.uint32), ref)           # integer part: bit exact
    z = torch.empty(n, dtype=torch.float32, device=dev)
    L.call("mfvi_philox_normal_fill", z.data_ptr(), n, L.key(99, 1, 2), 0)
    zr = philox.philox_normal(n, 99, 0, 2, 1)
    assert np.abs(z.cpu().numpy() - zr).max() < 2e-5                         # fp32 log/sincos vs fp64
    # device-side step counter adds to the step word
    ctr = torch.tensor([4], dtype=torch.int32, device=dev)
    z2 = torch.empty(n, dtype=torch.float32, device=dev)
    L.call("mfvi_philox_normal_fill", z2.data_ptr(), n, L.key(99, 1, 2, ctr), 0)
    assert np.abs(z2.cpu().numpy() - philox.philox_normal(n, 99, 0, 2, 5)).max() < 2e-5


# ----------------------------------------------------------------------------------------------- layers
@pytest.mark.parametrize("name", ["c1x1", "c3s1", "c3s2", "c5s2", "c5s1", "c3s1_odd"])
def test_conv2drt_matches_reference(dev, name):
    from mfvi_dip_mia_b200.BayTorch.modules import Conv2dRT
    g = group(load_npz("layers.npz"), name + "/")
    cin, cout, k, st, H, W = [int(v) for v in g["meta"]]
    layer = Conv2dRT(cin, cout, k, stride=st, prior={"mu": 0.0, "sigma": 1e-8}).to(dev)
    with torch.no_grad():
        for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
            getattr(layer, pn).copy_(g[pn])
    x = g["x"].to(dev).requires_grad_(True)
    layer.inject_eps(g["eps_w"], g["eps_b"])
    y = layer(x)
    assert rel_err(y.cpu(), g["y"]) < 1e-5
    y.backward(g["dy"].to(dev))
    assert rel_err(x.grad.cpu(), g["dx"]) < 1e-5
    for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
        assert rel_err(getattr(layer, pn).grad.cpu(), g["d" + pn]) < 2e-5, pn
    layer.eval()
    assert rel_err(layer(g["x"].to(dev)).cpu(), g["y_eval"]) < 1e-5


@pytest.mark.parametrize("name", ["l1x1", "l3s1p1", "l3s2", "l5s1nb"])
def test_conv2dlrt_matches_reference(dev, name):
    """Local reparameterisation (row f3): Conv2dLRT forward / backward / eval / KL against the imported reference
    (tests/golden/lrt_layers.npz, output-space eps injected)."""
    from mfvi_dip_mia_b200.BayTorch.modules import Conv2dLRT
    g = group(load_npz("lrt_layers.npz"), name + "/")
    cin, cout, k, st, pad, bias, N, H, W = [int(v) for v in g["meta"]]
    layer = Conv2dLRT(cin, cout, k, bias=bool(bias), stride=st, padding=pad, prior={"mu": 0.0, "sigma": 0.05}).to(dev)
    pnames = ["W_mu", "W_rho"] + (["bias_mu", "bias_rho"] if bias else [])
    with torch.no_grad():
        for pn in pnames:
            getattr(layer, pn).copy_(g[pn])
    x = g["x"].to(dev).requires_grad_(True)
    layer.inject_eps(g["eps"])
    y = layer(x)
    assert y.shape == g["y"].shape and rel_err(y.cpu(), g["y"]) < 1e-5
    y.backward(g["dy"].to(dev))
    assert rel_err(x.grad.cpu(), g["dx"]) < 2e-5
    for pn in pnames:
        assert rel_err(getattr(layer, pn).grad.cpu(), g["d" + pn]) < 5e-5, pn
    layer.eval()
    assert rel_err(layer(g["x"].to(dev)).cpu(), g["y_eval"]) < 1e-5
    assert rel_err(layer._kl.cpu(), g["kl"]) < 1e-5
    layer.train()                                      # fresh Philox eps: two forwards differ, mean stays close to eval
    a, b = layer(g["x"].to(dev)), layer(g["x"].to(dev))
    assert not torch.equal(a, b) and torch.isfinite(a).all()


def test_linearlrt_matches_reference(dev):
    from mfvi_dip_mia_b200.BayTorch.modules import LinearLRT
    g = group(load_npz("lrt_layers.npz"), "lin/")
    lin = LinearLRT(20, 12, prior={"mu": 0.0, "sigma": 0.05}).to(dev)
    with torch.no_grad():
        for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
            getattr(lin, pn).copy_(g[pn])
    x = g["x"].to(dev).requires_grad_(True)
    lin.inject_eps(g["eps"])
    y = lin(x)
    assert rel_err(y.cpu(), g["y"]) < 1e-5
    y.backward(g["dy"].to(dev))
    assert rel_err(x.grad.cpu(), g["dx"]) < 2e-5
    for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
        assert rel_err(getattr(lin, pn).grad.cpu(), g["d" + pn]) < 5e-5, pn
    lin.eval()
    assert rel_err(lin(g["x"].to(dev)).cpu(), g["y_eval"]) < 1e-5


def test_meanfieldvi_local_reparam_trains(dev):
    """MeanFieldVI's default reparam='local' on the skip net: module-by-module execution on the library's kernels,
    finite loss and gradients for every parameter, KL equal to the weight-space model's KL at equal parameters."""
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    from mfvi_dip_mia_b200.models import get_net
    from mfvi_dip_mia_b200.utils.bayesian_utils import gaussian_nll
    torch.manual_seed(0)
    net = get_net(4, "skip", "reflection", upsample_mode="bilinear", n_channels=2, skip_n33d=8, skip_n33u=8, skip_n11=2,
                  num_scales=2, need_sigmoid=False)
    mf = MeanFieldVI(net, prior={"mu": 0.0, "sigma": 0.05}, device=dev)
    x = torch.rand(1, 4, 32, 32, device=dev) * 0.1
    t = torch.rand(1, 1, 32, 32, device=dev)
    out = mf(x)
    loss = gaussian_nll(out[:, :1], out[:, 1:], t) + 1e-6 * mf.kl()
    loss.backward()
    assert torch.isfinite(loss).all()
    for k, p in mf.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_linearrt_matches_reference(dev):
    from mfvi_dip_mia_b200.BayTorch.modules import LinearRT
    g = group(load_npz("layers.npz"), "lin/")
    layer = LinearRT(20, 12, prior={"mu": 0.0, "sigma": 1e-8}).to(dev)
    with torch.no_grad():
        for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
            getattr(layer, pn).copy_(g[pn])
    x = g["x"].to(dev).requires_grad_(True)
    layer.inject_eps(g["eps_w"], g["eps_b"])
    y = layer(x)
    assert rel_err(y.cpu(), g["y"]) < 1e-5
    y.backward(g["dy"].to(dev))
    assert rel_err(x.grad.cpu(), g["dx"]) < 1e-5
    for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
        assert rel_err(getattr(layer, pn).grad.cpu(), g["d" + pn]) < 2e-5, pn


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_layer_kl_matches_reference(dev, i):
    from mfvi_dip_mia_b200.BayTorch.modules import Conv2dRT
    g = group(load_npz("layers.npz"), f"kl{i}/")
    temp, sigma, rev = [float(x) for x in g["meta"]]
    layer = Conv2dRT(6, 5, 3, prior={"mu": 0.0, "sigma": np.sqrt(temp) * sigma},
                     kl_type="reverse" if rev else "forward").to(dev)
    with torch.no_grad():
        for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
            getattr(layer, pn).copy_(g[pn])
    kl = layer._kl
    assert rel_err(kl.cpu(), g["kl"]) < 1e-5
    kl.backward()
    for pn in ["W_mu", "W_rho", "bias_mu", "bias_rho"]:
        assert rel_err(getattr(layer, pn).grad.cpu(), g["d" + pn]) < 1e-4, pn


def test_nll_matches_reference(dev):
    from mfvi_dip_mia_b200.utils.bayesian_utils import gaussian_nll, gaussian_nll_inpainting
    d = load_npz("layers.npz")
    g = group(d, "nll/")
    mu, s = g["mu"].to(dev).requires_grad_(True), g["s"].to(dev).requires_grad_(True)
    v = gaussian_nll(mu, s, g["t"].to(dev))
    assert rel_err(v.cpu(), g["v"]) < 1e-5
    v.backward()
    assert rel_err(mu.grad.cpu(), g["dmu"]) < 1e-5 and rel_err(s.grad.cpu(), g["ds"]) < 1e-5
    assert rel_err(gaussian_nll(mu, s, g["t"].to(dev), reduction="sum").cpu(), g["v"] * mu.numel()) < 1e-5
    g = group(d, "nlli/")
    mu3, s1 = g["mu"].to(dev).requires_grad_(True), g["s"].to(dev).requires_grad_(True)
    v = gaussian_nll_inpainting(mu3.sigmoid(), s1, g["t"].to(dev), g["m"].to(dev))
    assert rel_err(v.cpu(), g["v"]) < 1e-5
    v.backward()
    assert rel_err(mu3.grad.cpu(), g["dmu"]) < 1e-5 and rel_err(s1.grad.cpu(), g["ds"]) < 1e-5


@pytest.mark.parametrize("name", ["r32", "r48", "r40c2"])
def test_radon_matches_reference(dev, name):
    from mfvi_dip_mia_b200.radon import FastRadonTransform
    g = group(load_npz("layers.npz"), name + "/")
    img = g["img"].to(dev).requires_grad_(True)
    rad = FastRadonTransform(tuple(img.shape), g["theta"]).to(dev)
    sino = rad(img)
    assert tuple(sino.shape) == tuple(g["sino"].shape)
    assert rel_err(sino.cpu(), g["sino"]) < 2e-5
    sino.backward(g["dsino"].to(dev))
    assert rel_err(img.grad.cpu(), g["dimg"]) < 2e-5


# ----------------------------------------------------------------------------------------------- whole step
def _fixture(task):
    d = load_npz(f"skipnet_small_{task}.npz")
    S = int(d["S"])
    sd = group(d, "sd/")
    eps = [group(d, f"eps{s}/") for s in range(S)]
    ex = group(d, "extra/")
    grads = group(d, "grad/")
    return d, S, sd, eps, ex, grads


def _head_kwargs(task, ex):
    if task in ("den", "sr"):
        return dict(target=ex["target"])
    if task == "inp":
        return dict(target=ex["target"], mask=ex["mask"])
    return dict(theta_deg=ex["theta"], sino=ex["sino"])


@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_engine_step_matches_reference(dev, task):
    """Engine + loss head + reparam/KL kernel == S sequential reference forwards + loss.backward()."""
    from mfvi_dip_mia_b200 import SkipEngine, _lib as L
    from mfvi_dip_mia_b200.engine import KL, NLL
    from mfvi_dip_mia_b200.trainer import LossHead
    d, S, sd, eps, ex, grads = _fixture(task)
    x = torch.from_numpy(d["net_input"])
    eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev)
    eng.load_params(sd, prefix="net.")
    eng.pack_eps(eps, prefix="net.")
    head = LossHead(eng, task, **_head_kwargs(task, ex))
    temp, sigma = float(d["temp"]), float(d["sigma"])
    eng.zero_accumulators()
    eng.set_input(x[0].permute(1, 2, 0).contiguous().to(dev), None, 0.0, L.key(0))
    eng.sample_weights(L.key(0))
    eng.forward()
    out = eng.out_nchw().cpu()
    for s in range(S):
        assert rel_err(out[s:s + 1], d[f"out{s}"]) < 1e-4, s
    head.run()
    eng.backward()
    eng.reparam_kl(L.key(0), prior_mu=0.0, prior_sigma_plus_eps=O.prior_scale(temp, sigma), direction=0, kscale=temp)
    a = eng.arena[:2].cpu()
    assert rel_err(a[NLL], d["nll"]) < 1e-5
    assert rel_err(a[KL], d["kl"]) < 1e-5
    ours = {"net." + k: v.cpu() for k, v in eng.param_views("grad").items()}
    errs = grad_errs({k: ours[k] for k in grads}, grads)
    worst = max(errs, key=errs.get)
    assert errs[worst] < RTOL, (worst, errs[worst])


@pytest.mark.parametrize("task", ["den", "inp"])
def test_meanfieldvi_dropin_matches_reference(dev, task):
    """The reference's own call sequence (get_net/skip -> MeanFieldVI -> nll + temp*kl -> backward) on our classes."""
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    from mfvi_dip_mia_b200.models.skip import skip
    from mfvi_dip_mia_b200.utils.bayesian_utils import gaussian_nll, gaussian_nll_inpainting
    d, S, sd, eps, ex, grads = _fixture(task)
    cfg = SMALL[task]
    temp, sigma = float(d["temp"]), float(d["sigma"])
    net = skip(cfg.num_input_channels, cfg.num_output_channels, num_channels_down=list(cfg.down),
               num_channels_up=list(cfg.up), num_channels_skip=list(cfg.skip), filter_size_down=cfg.filter_down,
               filter_size_up=cfg.filter_up, filter_skip_size=cfg.filter_skip, need_sigmoid=False, need_bias=True,
               pad="reflection", upsample_mode=cfg.upsample_mode, need1x1_up=cfg.need1x1_up,
               dropout_mode_down="None", dropout_mode_up="None", dropout_mode_skip="None", dropout_mode_output="None")
    net = MeanFieldVI(net, prior={"mu": 0.0, "sigma": np.sqrt(temp) * sigma}, replace_layers="all", reparam="",
                      device=dev, mc_samples=S)
    assert set(net.state_dict().keys()) == set(json.loads(str(d["keys"])))
    net.load_state_dict({k: v for k, v in sd.items()})
    x = torch.from_numpy(d["net_input"]).to(dev)
    net.prepare(x)
    net.inject_eps(eps, prefix="net.")      # eps keys in the fixture carry the 'net.' prefix of MeanFieldVI.net
    optimizer = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0)
    optimizer.zero_grad()
    out = net(x)
    assert tuple(out.shape) == (S, cfg.num_output_channels, x.shape[2], x.shape[3])
    if task == "den":
        nll = gaussian_nll(out[:, :1], out[:, 1:], ex["target"].to(dev))
    else:
        nll = gaussian_nll_inpainting(out[:, :3].sigmoid(), out[:, 3:], ex["target"].to(dev), ex["mask"].to(dev))
    kl = net.kl()
    assert tuple(kl.shape) == (1,)
    loss = nll + temp * kl
    loss.backward()
    assert rel_err(nll.cpu(), d["nll"]) < 1e-5 and rel_err(kl.cpu(), d["kl"]) < 1e-5
    assert rel_err(loss.cpu(), d["loss"]) < 1e-5
    ours = {k: p.grad.cpu() for k, p in net.named_parameters()}
    errs = grad_errs({k: ours[k] for k in grads}, grads)
    worst = max(errs, key=errs.get)
    assert errs[worst] < RTOL, (worst, errs[worst])
    optimizer.step()      # the optimiser updates the flat storage through the views
    assert torch.isfinite(net._engine.theta).all()


def test_den256_full_size_matches_reference(dev):
    """Metric-shape net (256^2, 26 convs), S=2, eps from OUR Philox stream definition regenerated by the oracle and
    injected — compared with summaries of the imported reference's step (den256_summary.npz)."""
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L
    from mfvi_dip_mia_b200.engine import KL, NLL
    from mfvi_dip_mia_b200.trainer import LossHead
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    d = load_npz("den256_summary.npz")
    S, seed = int(d["S"]), int(d["philox_seed"])
    temp, sigma = float(d["temp"]), float(d["sigma"])
    eng = SkipEngine(SkipSpec(), 256, 256, S, dev)
    # parameters: the reference net was built under torch.manual_seed(1) — re-create it through OUR classes, which
    # draw the initial values from the same global RNG in the same order (VIModule.reset_parameters)
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    from mfvi_dip_mia_b200.models import get_net
    torch.manual_seed(int(d["init_seed"]))
    net = get_net(16, "skip", "reflection", skip_n33d=[16, 32, 64, 128, 128], skip_n33u=[16, 32, 64, 128, 128],
                  skip_n11=4, num_scales=5, n_channels=2, upsample_mode="bilinear")
    net = MeanFieldVI(net, prior={"mu": 0.0, "sigma": np.sqrt(temp) * sigma}, replace_layers="all", reparam="")
    names = json.loads(str(d["grad_names"]))
    pn = dict(net.named_parameters())
    got_norms = np.array([float(pn[k].double().norm()) for k in names])
    assert np.allclose(got_norms, d["param_norms"], rtol=1e-6), "initialisation order differs from the reference"
    eng.load_params(net.net.state_dict())
    eps = []
    for s in range(S):
        e = {}
        for li, c in enumerate(eng.lay.convs):
            e[c.key + ".W"] = torch.from_numpy(philox.philox_normal(c.w_numel, seed, 2 * li, s, 0)).reshape(c.cout, c.cin, c.k, c.k)
            e[c.key + ".b"] = torch.from_numpy(philox.philox_normal(c.cout, seed, 2 * li + 1, s, 0))
        eps.append(e)
    eng.pack_eps(eps)
    target = torch.from_numpy(noisy(ellipse_phantom(256), 0.1, 1))[None]
    g = torch.Generator().manual_seed(int(d["input_seed"]))
    net_input = torch.rand(1, 16, 256, 256, generator=g) * 0.1
    head = LossHead(eng, "den", target=target)
    eng.zero_accumulators()
    eng.set_input(net_input[0].permute(1, 2, 0).contiguous().to(dev), None, 0.0, L.key(0))
    eng.sample_weights(L.key(0))
    eng.forward()
    head.run()
    eng.backward()
    eng.reparam_kl(L.key(0), prior_mu=0.0, prior_sigma_plus_eps=O.prior_scale(temp, sigma), direction=0, kscale=temp)
    out = eng.out_nchw().cpu()
    for s in range(S):
        assert rel_err(out[s:s + 1, :, ::8, ::8], d[f"out{s}_sub"]) < RTOL
        assert np.allclose(out[s].double().mean(dim=(1, 2)).numpy(), d[f"out{s}_mean"], rtol=1e-3, atol=1e-5)
    a = eng.arena[:2].cpu()
    assert rel_err(a[NLL], d["nll"]) < 1e-4 and rel_err(a[KL], d["kl"]) < 1e-5
    assert rel_err(a[NLL] + temp * a[KL], d["loss"]) < 1e-4
    gv = {"net." + k: v.cpu() for k, v in eng.param_views("grad").items()}
    norms = np.array([float(gv[k].double().norm()) for k in names])
    big = d["grad_norms"] > 1e-3 * d["grad_norms"].max()
    assert np.allclose(norms[big], d["grad_norms"][big], rtol=2e-3), np.abs(norms[big] / d["grad_norms"][big] - 1).max()
    gref = group(d, "grad/")
    for k, v in gref.items():
        ours = gv[k].reshape(-1)[:4096]
        den = max(float(v.abs().max()), 1e-3 * float(d["grad_absmax"].max()))
        assert float((ours - v).abs().max()) / den < RTOL * 2, k


# ----------------------------------------------------------------------------------------------- trainer
def _oracle_eps_from_stream(eng, seed, step, S, sample0=0):
    """eps the library's weight stream produces (flat storage index), mapped to reference-shaped tensors."""
    out = []
    for s in range(S):
        flat = torch.from_numpy(philox.philox_normal(eng.lay.P, seed, 0, sample0 + s, step))
        e = {}
        for c in eng.lay.convs:
            e["net." + c.key + ".W"] = flat[c.w_off:c.w_off + c.w_numel].view(c.k, c.k, c.cout, c.cin).permute(2, 3, 0, 1).contiguous()
            e["net." + c.key + ".b"] = flat[c.b_off:c.b_off + c.cout].clone()
        out.append(e)
    return out


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainer_steps_match_oracle(dev, use_graph):
    """4 optimiser steps of MfviDipTrainer (in-kernel Philox eps + input jitter, AdamW), eager and as a CUDA graph.
    AdamW's sign-like early updates amplify fp32 summation noise chaotically, so each step is checked on its own:
    the oracle is evaluated at the GPU's parameters before the step on the same regenerated streams (loss terms,
    every gradient), and the parameter update is checked against the oracle's AdamW applied to the GPU gradient."""
    from mfvi_dip_mia_b200 import MfviDipTrainer
    task = "den"
    d, S, sd, _, ex, _ = _fixture(task)
    cfg = SMALL[task]
    temp, sigma, lr, seed = float(d["temp"]), float(d["sigma"]), 1e-2, 4242
    x = torch.from_numpy(d["net_input"])
    tr = MfviDipTrainer(spec_of(cfg), task, x, temp=temp, sigma=sigma, lr=lr, mc_samples=S, seed=seed, device=dev,
                        target=ex["target"], use_graph=use_graph)
    tr.eng.load_params(sd, prefix="net.")
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    H, W, Cn = x.shape[2], x.shape[3], x.shape[1]
    rm0 = tr.eng.running_mean.clone()
    for it in range(4):
        theta0, m0, v0 = tr.eng.theta.clone().double().cpu(), tr.m.clone().double().cpu(), tr.v.clone().double().cpu()
        p_old = {"net." + k: t.detach().clone().cpu() for k, t in tr.eng.param_views("theta").items()}
        tr.step()
        nll_g, kl_g, loss_g = tr.loss_terms()
        z = torch.from_numpy(philox.philox_normal(Cn * H * W, seed, 1, 0, it)).reshape(1, Cn, H, W)
        leaves = {k: p_old[k].clone().requires_grad_(True) for k in names}
        full = dict(sd)
        full.update(leaves)
        eps = _oracle_eps_from_stream(tr.eng, seed, it, S)
        loss, nll, kl, _ = O.mfvi_loss(full, cfg, x + 0.1 * z, eps, task=task, temp=temp,
                                       prior_sigma_plus_eps=O.prior_scale(temp, sigma), target=ex["target"])
        loss.backward()
        assert rel_err(nll_g, nll) < 1e-4, (it, nll_g, float(nll))
        assert rel_err(kl_g, kl) < 1e-5, it
        ours = {"net." + k: t.cpu() for k, t in tr.eng.param_views("grad").items()}
        errs = grad_errs({k: ours[k] for k in names}, {k: leaves[k].grad for k in names})
        worst = max(errs, key=errs.get)
        # 90 % of the tensors within rtol 1e-3; the rest are the badly conditioned BatchNorm biases (sums that cancel to ~0, where
        # the reference's own fp32 run is up to 2e-2 .. 6e-2 of the tensor's scale away from its fp64 run — see
        # tests/test_gpu_fullsize_parity.py — and where our own result moves by ~2e-2 from run to run with the order of the atomics)
        assert float(np.percentile(list(errs.values()), 90)) < RTOL, (it, float(np.percentile(list(errs.values()), 90)))
        assert errs[worst] < 6e-2, (it, worst, errs[worst])
        # AdamW kernel == torch.optim.AdamW arithmetic on the gradient the GPU produced
        p_ref, _, _ = O.adamw_step(theta0, tr.eng.grad.double().cpu(), m0, v0, it + 1, lr)
        assert float((tr.eng.theta.double().cpu() - p_ref).abs().max()) < 1e-6 + 1e-4 * lr, it
    assert tr.steps_done == 4
    assert torch.isfinite(tr.eng.running_var).all() and not torch.equal(tr.eng.running_mean, rm0)


# ----------------------------------------------------------------------------------------------- bookkeeping (row f1)
def _host_bookkeeping(gt, ring):
    """The reference's per-iteration bookkeeping (bayesian_optimization.py:1374-1416) restated with torch on the host."""
    state = {"avg": None, "means": [], "vars": []}

    def update(out):                                             # out (S,2,H,W): mean over MC samples first
        cur = torch.cat([out[:, :1].mean(0, keepdim=True), torch.exp(-out[:, 1:]).mean(0, keepdim=True)], 1)
        state["avg"] = cur.clone() if state["avg"] is None else state["avg"] * 0.99 + cur * 0.01
        state["means"] = (state["means"] + [cur[:, :1].clamp(0, 1)])[-ring:]
        state["vars"] = (state["vars"] + [cur[:, 1:].clamp(0, 1)])[-ring:]
        state["cur"] = cur

    def metrics():
        sm = state["avg"][:, :1].clamp(0, 1)
        means = torch.cat(state["means"])
        unc = means.var(0, unbiased=True) + torch.cat(state["vars"]).mean(0)
        err2 = ((means - gt) ** 2).mean(0)
        return O.psnr(gt, sm), O.ssim(gt, sm), O.uce(err2.reshape(-1), unc.reshape(-1), 15)
    return update, metrics, state


def test_bookkeeping_kernels_match_host_restatement(dev):
    """mfvi_bookkeep_step / mfvi_ssim / mfvi_ring_uncertainty against the torch restatement of the reference's
    bookkeeping on random network outputs: 31 iterations (ring of 7 wraps 4 times), S=2, non-square image that is not a
    multiple of the SSIM tile."""
    from mfvi_dip_mia_b200 import _lib as L
    S, H, W, ring = 2, 40, 72, 7
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(1, 1, H, W, generator=g)
    noisy_t = (gt + 0.1 * torch.randn(1, 1, H, W, generator=g)).clamp(0, 1)
    upd, met, st = _host_bookkeeping(gt, ring)
    out_avg = torch.zeros(2, H, W, device=dev)
    r_epi, r_ale = torch.zeros(ring, H, W, device=dev), torch.zeros(ring, H, W, device=dev)
    acc = torch.zeros(8, dtype=torch.float64, device=dev)
    it_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    gt_d, no_d = gt.reshape(H, W).to(dev), noisy_t.reshape(H, W).to(dev)
    for it in range(31):
        out = torch.randn(S, 2, H, W, generator=g) * 0.5 + 0.4
        upd(out)
        o = torch.zeros(S, H, W, 4, device=dev)                  # channel pitch 4 like the engine's output buffer
        o[..., :2] = out.permute(0, 2, 3, 1).to(dev)
        acc.zero_()
        L.call("mfvi_bookkeep_step", L.view(o), S, H, W, 0.99, gt_d.data_ptr(), no_d.data_ptr(), out_avg.data_ptr(),
               r_epi.data_ptr(), r_ale.data_ptr(), ring, it_dev.data_ptr(), 0, acc.data_ptr())
        L.call("mfvi_counter_add", it_dev.data_ptr(), 1)
    assert rel_err(out_avg.cpu(), st["avg"][0]) < 1e-5
    a = acc.cpu().numpy()
    n = H * W
    cur = st["cur"]
    assert abs(10 * np.log10(n / a[2]) - O.psnr(gt, st["avg"][:, :1].clamp(0, 1))) < 1e-3
    assert abs(10 * np.log10(n / a[1]) - O.psnr(gt, cur[:, :1].clamp(0, 1))) < 1e-3
    assert abs(10 * np.log10(n / a[0]) - O.psnr(noisy_t, cur[:, :1].clamp(0, 1))) < 1e-3
    assert abs(a[3] / n - float(((st["avg"][:, :1] - noisy_t) ** 2).mean())) < 1e-6
    assert abs(a[4] / n - float(((st["avg"][:, :1] - gt) ** 2).mean())) < 1e-6
    L.call("mfvi_ssim", gt_d.data_ptr(), out_avg.data_ptr(), H, W, 1, acc[5:].data_ptr())
    assert abs(float(acc[5]) / n - O.ssim(gt, st["avg"][:, :1].clamp(0, 1))) < 1e-5
    epi, ale, err2 = (torch.empty(H, W, device=dev) for _ in range(3))
    L.call("mfvi_ring_uncertainty", r_epi.data_ptr(), r_ale.data_ptr(), ring, H, W, gt_d.data_ptr(), epi.data_ptr(),
           ale.data_ptr(), err2.data_ptr())
    means, vars_ = torch.cat(st["means"]), torch.cat(st["vars"])
    assert rel_err(epi.cpu(), means.var(0, unbiased=True)[0]) < 1e-4
    assert rel_err(ale.cpu(), vars_.mean(0)[0]) < 1e-5
    assert rel_err(err2.cpu(), ((means - gt) ** 2).mean(0)[0]) < 1e-5


# ----------------------------------------------------------------------------------------------- trajectory metrics
_TRAJ = dict(H=64, temp=5.656911698337764e-07, sigma=1.4616642493692077e-05, lr=1e-2, ring=25)


def test_den_short_trajectory_metrics_match_oracle(dev):
    """PSNR / SSIM / UCE after 25 optimiser steps (one full ring) of the denoising runner (device-side bookkeeping inside
    the step graph) against the CPU oracle replaying the same Philox streams — north_star's 0.1 dB / 0.005 bar, pointwise.
    A few tens of steps is the horizon over which one trajectory is comparable at all (at 40 steps PSNR / SSIM still agree
    to 0.014 dB / 0.002 but the UCE, a binned statistic of the tiny ring variance, is already 0.007 apart): the optimisation is chaotic (the SAME reference
    arithmetic in fp32 vs fp64, or with eps perturbed by 2e-6, is 0.1 dB apart after 100 steps, 0.4 dB after 300 and
    0.6-0.8 dB after 2400; measured with the oracle) — the long horizon is covered by the
    ensemble test below."""
    from mfvi_dip_mia_b200 import MfviDipTrainer, _lib as L
    from mfvi_dip_mia_b200.runners import DeviceBookkeeping
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    cfg = SMALL["den"]
    H = W = _TRAJ["H"]
    n_it, ring, seed = 25, _TRAJ["ring"], 11
    temp, sigma, lr = _TRAJ["temp"], _TRAJ["sigma"], _TRAJ["lr"]
    gt = torch.from_numpy(ellipse_phantom(H))[None]
    tgt = torch.from_numpy(noisy(ellipse_phantom(H), 0.1, 1))[None]
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, cfg.num_input_channels, H, W, generator=g) * 0.1
    tr = MfviDipTrainer(spec_of(cfg), "den", x, temp=temp, sigma=sigma, lr=lr, mc_samples=1, seed=seed, device=dev,
                        target=tgt, math_mode=L.MATH_FP32, use_graph=True)
    bk = DeviceBookkeeping(tr, gt=gt, noisy=tgt, ring=ring)
    sd0 = {"net." + k: v.detach().cpu().clone() for k, v in tr.eng.param_views("theta").items()}
    for _ in range(n_it):
        tr.step()
    m = bk.metrics()
    pg, sg, ug = m["psnr_gt_sm"], m["ssim_gt_sm"], bk.uce()
    names = [k for k, v in sd0.items() if v.is_floating_point() and "running" not in k]
    Cn = cfg.num_input_channels
    leaves = {k: sd0[k].clone().requires_grad_(True) for k in names}
    full = dict(sd0)
    full.update(leaves)
    opt = torch.optim.AdamW(list(leaves.values()), lr=lr, weight_decay=0)
    upd, met, _ = _host_bookkeeping(gt, ring)
    for it in range(n_it):
        opt.zero_grad()
        z = torch.from_numpy(philox.philox_normal(Cn * H * W, seed, 1, 0, it)).reshape(1, Cn, H, W)
        eps = _oracle_eps_from_stream(tr.eng, seed, it, 1)
        loss, _, _, outs = O.mfvi_loss(full, cfg, x + 0.1 * z, eps, task="den", temp=temp,
                                       prior_sigma_plus_eps=O.prior_scale(temp, sigma), target=tgt)
        loss.backward()
        opt.step()
        upd(outs[0].detach())
    po, so, uo = met()
    print(f"[trajectory {n_it} steps] PSNR {pg:.4f} / {po:.4f} dB  SSIM {sg:.5f} / {so:.5f}  UCE {ug:.5f} / {uo:.5f}  (gpu / oracle)")
    assert abs(pg - po) < 0.1 and abs(sg - so) < 0.005 and abs(ug - uo) < 0.005


@pytest.mark.parametrize("math", ["tf32", "fp32"])
def test_den_final_metrics_match_reference_ensemble(dev, math):
    """Final PSNR / SSIM / UCE of the denoising runner after 1200 iterations on the 64x64 phantom, as a DISTRIBUTION over
    seeds, against the same distribution produced by the imported reference itself (tests/golden/trajectory_den64.npz,
    64 seeds, generated by tests/golden/make_trajectory_golden.py: the reference's MeanFieldVI + skip net + torch RNG).
    Single trajectories are chaotic (see the test above), so the bar is on the ensemble means: north_star's 0.1 dB /
    0.005 plus 3 standard errors of the difference of the two means."""
    from mfvi_dip_mia_b200 import _lib as L
    from mfvi_dip_mia_b200.runners import run_den_mfvi
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    d = load_npz("trajectory_den64.npz")
    its = [int(v) for v in d["its"]]
    ref = np.asarray(d["metrics"])[:, its.index(1200)]            # (K_ref, 3)
    H = _TRAJ["H"]
    gt = ellipse_phantom(H)
    ny = noisy(gt, 0.1, 1)
    K = 32
    ours = []
    for k in range(K):
        _, h = run_den_mfvi(gt, temp=_TRAJ["temp"], sigma=_TRAJ["sigma"], lr=_TRAJ["lr"], num_iter=1199, mc_samples=1,
                            seed=1000 + k, device=dev, mc_ring=_TRAJ["ring"], show_every=10 ** 9,
                            math_mode=L.MATH_TF32 if math == "tf32" else L.MATH_FP32, return_history=True,
                            spec=spec_of(SMALL["den"]), img_noisy=ny)
        ours.append((h["psnr_gt_sm"][-1], h["ssim_gt_sm"][-1], h["uce"]))
    ours = np.array(ours)
    mo, mr = ours.mean(0), ref.mean(0)
    se = np.sqrt(ours.var(0, ddof=1) / len(ours) + ref.var(0, ddof=1) / len(ref))
    print(f"[ensemble {math}] ours  PSNR {mo[0]:.3f} SSIM {mo[1]:.4f} UCE {mo[2]:.4f}  (std {ours.std(0, ddof=1)})")
    print(f"[ensemble {math}] ref   PSNR {mr[0]:.3f} SSIM {mr[1]:.4f} UCE {mr[2]:.4f}  (std {ref.std(0, ddof=1)})  se {se}")
    tol = np.array([0.1, 0.005, 0.005])
    assert np.all(np.abs(mo - mr) < tol + 3 * se), (mo, mr, se)


# ----------------------------------------------------------------------------------------------- paired trajectories
_PAIRED = dict(n_it=400, seeds=(31, 32, 33, 34, 35, 36, 37, 38), lr=1e-3)
_paired_oracle_cache = {}


def _paired_oracle_run(dev, seed, sd0, x, gt, tgt, n_it, lay_eng):
    """The reference arithmetic (oracle port, PyTorch eager fp32 on the GPU, TF32 off) driven by the SAME Philox streams the
    engine draws (weights eps keyed by step / sample, input jitter): returns (psnr_gt_sm, ssim_gt_sm, uce) after n_it steps."""
    if seed in _paired_oracle_cache:
        return _paired_oracle_cache[seed]
    cfg = SMALL["den"]
    H, W, Cn = x.shape[2], x.shape[3], x.shape[1]
    temp, sigma, lr = _TRAJ["temp"], _TRAJ["sigma"], _PAIRED["lr"]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        names = [k for k, v in sd0.items() if v.is_floating_point() and "running" not in k]
        leaves = {k: sd0[k].clone().to(dev).requires_grad_(True) for k in names}
        full = {k: v.to(dev) for k, v in sd0.items()}
        full.update(leaves)
        opt = torch.optim.AdamW(list(leaves.values()), lr=lr, weight_decay=0)
        upd, met, _ = _host_bookkeeping(gt, _TRAJ["ring"])
        xd, tg = x.to(dev), tgt.to(dev)
        for it in range(n_it):
            opt.zero_grad()
            z = torch.from_numpy(philox.philox_normal(Cn * H * W, seed, 1, 0, it)).reshape(1, Cn, H, W).to(dev)
            eps = [{k: v.to(dev) for k, v in e.items()} for e in _oracle_eps_from_stream(lay_eng, seed, it, 1)]
            loss, _, _, outs = O.mfvi_loss(full, cfg, xd + 0.1 * z, eps, task="den", temp=temp,
                                           prior_sigma_plus_eps=O.prior_scale(temp, sigma), target=tg)
            loss.backward()
            opt.step()
            upd(outs[0].detach().cpu())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    _paired_oracle_cache[seed] = met()
    return _paired_oracle_cache[seed]


@pytest.mark.parametrize("math", ["tf32", "fp32"])
def test_den_paired_trajectories_match_reference_arithmetic(dev, math):
    """north_star's final-metric bar (PSNR / SSIM / UCE within 0.1 dB / 0.005 / 0.005) as a PAIRED comparison: for every seed the
    engine and the reference arithmetic (oracle port in eager fp32 on the same GPU) run 400 optimiser steps on identical
    Philox streams at the denoising configuration's own learning rate (test_configs/mfvi_den.json: 1e-3).  The optimisation is
    chaotic, so single pairs drift apart (measured on a B200: up to 0.1 dB in fp32, 0.27 dB in tf32 after 400 steps, 0.4 - 0.7 dB
    after 1000 in BOTH modes — two fp32 implementations of the same arithmetic included), but the pairing removes the
    seed-to-seed spread (0.8 dB) from the comparison: the standard error of the mean difference is 0.02 - 0.05 dB instead of
    the 0.17 dB of the unpaired ensemble test above.  Bar: |mean difference| < 0.1 dB / 0.005 / 0.005 + 2 standard errors, and
    every single pair within 5x the bar."""
    from mfvi_dip_mia_b200 import MfviDipTrainer, _lib as L
    from mfvi_dip_mia_b200.runners import DeviceBookkeeping
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    cfg = SMALL["den"]
    H = W = _TRAJ["H"]
    n_it, lr = _PAIRED["n_it"], _PAIRED["lr"]
    gt = torch.from_numpy(ellipse_phantom(H))[None]
    tgt = torch.from_numpy(noisy(ellipse_phantom(H), 0.1, 1))[None]
    diffs = []
    for seed in _PAIRED["seeds"]:
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(1, cfg.num_input_channels, H, W, generator=g) * 0.1
        tr = MfviDipTrainer(spec_of(cfg), "den", x, temp=_TRAJ["temp"], sigma=_TRAJ["sigma"], lr=lr, mc_samples=1, seed=seed,
                            device=dev, target=tgt, math_mode=L.MATH_TF32 if math == "tf32" else L.MATH_FP32, use_graph=True)
        bk = DeviceBookkeeping(tr, gt=gt, noisy=tgt, ring=_TRAJ["ring"])
        sd0 = {"net." + k: v.detach().cpu().clone() for k, v in tr.eng.param_views("theta").items()}
        for _ in range(n_it):
            tr.step()
        m = bk.metrics()
        ours = np.array([m["psnr_gt_sm"], m["ssim_gt_sm"], bk.uce()])
        ref = np.array(_paired_oracle_run(dev, seed, sd0, x, gt, tgt, n_it, tr.eng))
        diffs.append(ours - ref)
        print(f"[paired {math} seed {seed}] PSNR {ours[0]:.4f} / {ref[0]:.4f}  SSIM {ours[1]:.5f} / {ref[1]:.5f}  "
              f"UCE {ours[2]:.5f} / {ref[2]:.5f}  (engine / reference arithmetic)")
    diffs = np.array(diffs)
    tol = np.array([0.1, 0.005, 0.005])
    mean, se = diffs.mean(0), diffs.std(0, ddof=1) / np.sqrt(len(diffs))
    print(f"[paired {math}] mean difference {mean}  standard error {se}  worst pair {np.abs(diffs).max(0)}")
    assert np.all(np.abs(mean) < tol + 2 * se), (mean, se)
    assert np.all(np.abs(diffs).max(0) < 5 * tol), np.abs(diffs).max(0)


@pytest.mark.parametrize("variant", ["inp", "ct"])
def test_bookkeeping_kernel_other_runners(dev, variant):
    """mfvi_bookkeep_step_ex for the inpainting (3 sigmoid channels + s, masked gt comparisons: reference
    bayesian_optimization.py:3041-3069) and CT (one channel, no aleatoric channel: :583-610) runners against the reference's
    per-iteration formulas written with torch, over 9 iterations (ring of 4 wraps twice), S=2."""
    from mfvi_dip_mia_b200 import _lib as L
    S, H, W, ring, expw = 2, 24, 40, 4, 0.99
    Cm, sig, ale = (3, True, True) if variant == "inp" else (1, False, False)
    g = torch.Generator().manual_seed(9)
    gt = torch.rand(Cm, H, W, generator=g)
    noisy_t = gt.clone() if variant == "inp" else (gt + 0.05 * torch.randn(Cm, H, W, generator=g))
    mask = (torch.rand(1, H, W, generator=g) < 0.6).float() if variant == "inp" else None
    Ctot = Cm + (1 if ale else 0)
    out_avg = torch.zeros(Ctot, H, W, device=dev)
    r_epi, r_ale = torch.zeros(Cm, ring, H, W, device=dev), torch.zeros(ring, H, W, device=dev)
    acc = torch.zeros(8, dtype=torch.float64, device=dev)
    it_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    gt_d, no_d = gt.to(dev).contiguous(), noisy_t.to(dev).contiguous()
    mk_d = None if mask is None else mask.to(dev).contiguous()
    avg_ref, ring_e, ring_a = None, torch.zeros(Cm, ring, H, W), torch.zeros(ring, H, W)
    for it in range(9):
        out = torch.randn(S, H, W, Ctot, generator=g) * 0.7 + 0.3
        o = torch.zeros(S, H, W, 4, device=dev)                          # channel pitch 4 like the engine's output buffer
        o[..., :Ctot] = out.to(dev)
        acc.zero_()
        L.call("mfvi_bookkeep_step_ex", L.view(o), S, H, W, Cm, (1 if sig else 0) | (2 if ale else 0), expw, gt_d.data_ptr(),
               no_d.data_ptr(), L.ptr(mk_d), out_avg.data_ptr(), r_epi.data_ptr(), r_ale.data_ptr(), ring, it_dev.data_ptr(), 0,
               acc.data_ptr())
        L.call("mfvi_counter_add", it_dev.data_ptr(), 1)
        img = out[..., :Cm]
        cur = (torch.sigmoid(img) if sig else img).mean(0).permute(2, 0, 1)
        parts = [cur] + ([torch.exp(-out[..., Cm]).mean(0)[None]] if ale else [])
        cur_all = torch.cat(parts)
        avg_ref = cur_all.clone() if avg_ref is None else avg_ref * expw + cur_all * (1 - expw)
        ring_e[:, it % ring] = cur.clamp(0, 1)
        if ale:
            ring_a[it % ring] = cur_all[Cm].clamp(0, 1)
    assert rel_err(out_avg.cpu(), avg_ref) < 1e-5
    assert rel_err(r_epi.cpu(), ring_e) < 1e-6 and (not ale or rel_err(r_ale.cpu(), ring_a) < 1e-5)
    a = acc.cpu().numpy()
    mk = mask if mask is not None else torch.ones(1, H, W)
    mc, am, amc = cur.clamp(0, 1), avg_ref[:Cm], avg_ref[:Cm].clamp(0, 1)
    ref = [((noisy_t - mc) ** 2).sum(), ((gt * mk - mc * mk) ** 2).sum(), ((gt * mk - amc * mk) ** 2).sum(),
           ((noisy_t - am) ** 2).sum(), ((gt - am) ** 2).sum()]
    for k in range(5):
        assert abs(a[k] - float(ref[k])) < 1e-4 * max(1.0, float(ref[k])), (k, a[k], float(ref[k]))
