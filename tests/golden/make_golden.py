"""Generates tests/golden/*.npz by running the IMPORTED REFERENCE (read-only, /root/reference) on
seeded synthetic inputs.  Runs only in the build container (the GPU box has no /root/reference);
the committed .npz files are what tests consume.

    python tests/golden/make_golden.py

eps is injected by replacing VIModule.rsample (BayTorch/modules/module.py:82-85) with a function
that pops pre-drawn eps in call order; everything else is the reference's unmodified code.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "skimage", "skimage.metrics", "skimage.feature",
              "seaborn"]:
        try:
            __import__(n)
        except Exception:
            sys.modules[n] = types.ModuleType(n)
    sm = sys.modules["skimage.metrics"]
    if not hasattr(sm, "peak_signal_noise_ratio"):
        sm.peak_signal_noise_ratio = lambda *a, **k: None
        sm.structural_similarity = lambda *a, **k: None
    sys.path.insert(0, "/root/reference")
    import BayTorch.modules.module as ref_module
    from BayTorch.freq_to_bayes import MeanFieldVI
    from BayTorch.modules import Conv2dRT, LinearRT
    from models import get_net
    from models.skip import skip
    from radon import FastRadonTransform
    from utils.bayesian_utils import gaussian_nll, gaussian_nll_inpainting
    from utils.common_utils import peak_signal_noise_ratio, structural_similarity
    from utils.uce import uceloss
    return dict(module=ref_module, MeanFieldVI=MeanFieldVI, Conv2dRT=Conv2dRT, LinearRT=LinearRT, get_net=get_net,
                skip=skip, Radon=FastRadonTransform, nll=gaussian_nll, nll_inp=gaussian_nll_inpainting,
                psnr=peak_signal_noise_ratio, ssim=structural_similarity, uce=uceloss)


R = import_reference()
from oracle import mfvi_oracle as O          # noqa: E402
from oracle.philox import philox_normal      # noqa: E402


class EpsInjector:
    """Replaces VIModule.rsample; records shapes, returns mu + eps*sigma with eps from a queue."""

    def __init__(self):
        self.queue = []
        self.seen = []

    def __enter__(self):
        self.orig = R["module"].VIModule.rsample
        inj = self

        def rsample(mu, sigma):
            eps = inj.queue.pop(0)
            assert eps.shape == mu.shape, (eps.shape, mu.shape)
            inj.seen.append(tuple(mu.shape))
            return mu + eps * sigma

        R["module"].VIModule.rsample = staticmethod(rsample)
        return self

    def __exit__(self, *a):
        R["module"].VIModule.rsample = self.orig


def build_ref_net(cfg: O.SkipCfg, prior_sigma):
    net = R["skip"](cfg.num_input_channels, cfg.num_output_channels,
                    num_channels_down=list(cfg.down), num_channels_up=list(cfg.up), num_channels_skip=list(cfg.skip),
                    filter_size_down=cfg.filter_down, filter_size_up=cfg.filter_up, filter_skip_size=cfg.filter_skip,
                    need_sigmoid=cfg.need_sigmoid, need_bias=True, pad="reflection", upsample_mode=cfg.upsample_mode,
                    need1x1_up=cfg.need1x1_up, dropout_mode_down="None", dropout_mode_up="None",
                    dropout_mode_skip="None", dropout_mode_output="None")
    return R["MeanFieldVI"](net, prior={"mu": 0.0, "sigma": prior_sigma}, replace_layers="all", reparam="")


def make_eps(lay: O.SkipLayout, sd, S, seed, step=0, use_philox=False):
    """eps dicts keyed '<convkey>.W' / '.b' per sample."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for s in range(S):
        d = {}
        for li, c in enumerate(lay.convs_in_exec_order()):
            for j, (suffix, pk) in enumerate(((".W", ".W_mu"), (".b", ".bias_mu"))):
                shape = sd[c.key + pk].shape
                if use_philox:
                    z = philox_normal(int(np.prod(shape)), seed, 2 * li + j, s, step)
                    d[c.key + suffix] = torch.from_numpy(z).reshape(shape)
                else:
                    d[c.key + suffix] = torch.randn(shape, generator=g)
        out.append(d)
    return out


def run_ref_step(net, cfg, lay, net_input, eps_list, task, temp, extra):
    """S sequential reference forwards + loss.backward(), exactly as SURVEY §4 prescribes."""
    net.zero_grad()
    nlls, outs = [], []
    with EpsInjector() as inj:
        for eps in eps_list:
            for c in lay.convs_in_exec_order():
                inj.queue += [eps[c.key + ".W"], eps[c.key + ".b"]]
            out = net(net_input)
            assert not inj.queue
            if task == "den":
                nll = R["nll"](out[:, :1], out[:, 1:], extra["target"])
            elif task == "sr":
                lr = torch.nn.functional.interpolate(out, scale_factor=1 / extra["factor"], mode="nearest",
                                                     recompute_scale_factor=False)
                nll = R["nll"](lr[:, :1], lr[:, 1:], extra["target"])
            elif task == "inp":
                nll = R["nll_inp"](out[:, :3].sigmoid(), out[:, 3:], extra["target"], extra["mask"])
            elif task == "ct":
                nll = torch.nn.functional.mse_loss(extra["radon"](out), extra["sino"])
            outs.append(out.detach().clone())
            nlls.append(nll)
    nll_mean = torch.stack(nlls).mean()
    kl = net.kl()
    loss = nll_mean + temp * kl
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    return loss.detach(), nll_mean.detach(), kl.detach(), outs, grads


def save(name, **arrs):
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        flat[k] = v
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **flat)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(flat)} arrays")


SMALL = {
    "den": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "sr": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "ct": O.SkipCfg(4, 1, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "inp": O.SkipCfg(4, 4, (8, 16, 16), (8, 16, 16), (0, 0, 0), 5, 3, 1, False, False, "nearest"),
}
FULL = {
    "den": O.SkipCfg(16, 2),
    "sr": O.SkipCfg(32, 2),
    "ct": O.SkipCfg(16, 1),
    "inp": O.SkipCfg(16, 4, (16, 32, 64, 128, 128, 128), (16, 32, 64, 128, 128, 128), (0,) * 6, 5, 3, 1, False, False,
                     "nearest"),
}


def small_task(task, seed, S=2, hw=32):
    cfg = SMALL[task]
    temp, sigma = 5.656911698337764e-07, 1.4616642493692077e-05
    torch.manual_seed(seed)
    net = build_ref_net(cfg, np.sqrt(temp) * sigma)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    lay = O.skip_layout(cfg)
    g = torch.Generator().manual_seed(seed + 100)
    # perturb BN affine so that gamma/beta gradients are exercised away from the (1,0) init
    with torch.no_grad():
        for k, p in net.named_parameters():
            if "BatchNorm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net_input = torch.rand(1, cfg.num_input_channels, hw, hw, generator=g) * 0.1
    extra, meta = {}, {}
    if task == "den":
        extra["target"] = torch.rand(1, 1, hw, hw, generator=g)
    elif task == "sr":
        extra["factor"] = 4
        extra["target"] = torch.rand(1, 1, hw // 4, hw // 4, generator=g)
    elif task == "inp":
        extra["target"] = torch.rand(1, 3, hw, hw, generator=g)
        extra["mask"] = (torch.rand(1, 1, hw, hw, generator=g) < 0.5).float()
    elif task == "ct":
        theta = torch.arange(0, 180., step=30.)
        extra["radon"] = R["Radon"]((1, 1, hw, hw), theta)
        extra["sino"] = extra["radon"](torch.rand(1, 1, hw, hw, generator=g)).detach()
        meta["theta"] = theta
    eps_list = make_eps(lay, sd, S, seed + 7)
    loss, nll, kl, outs, grads = run_ref_step(net, cfg, lay, net_input, eps_list, task, temp, extra)
    arrs = {"loss": loss, "nll": nll, "kl": kl, "temp": np.float64(temp), "sigma": np.float64(sigma),
            "net_input": net_input, "S": np.int64(S), "keys": np.array(json.dumps(list(sd.keys())))}
    for i, o in enumerate(outs):
        arrs[f"out{i}"] = o
    for k, v in sd.items():
        arrs["sd/" + k] = v
    for k, v in grads.items():
        arrs["grad/" + k] = v
    for s, eps in enumerate(eps_list):
        for k, v in eps.items():
            arrs[f"eps{s}/" + k] = v
    for k, v in extra.items():
        if isinstance(v, torch.Tensor):
            arrs["extra/" + k] = v
    for k, v in meta.items():
        arrs["extra/" + k] = v
    save(f"skipnet_small_{task}.npz", **arrs)


def layer_fixtures():
    arrs = {}
    g = torch.Generator().manual_seed(11)
    shapes = [  # (name, cin, cout, k, stride, H, W)
        ("c1x1", 16, 4, 1, 1, 12, 12), ("c3s1", 12, 16, 3, 1, 10, 14), ("c3s2", 16, 24, 3, 2, 14, 18),
        ("c5s2", 8, 8, 5, 2, 16, 16), ("c5s1", 4, 8, 5, 1, 9, 9), ("c3s1_odd", 36, 16, 3, 1, 9, 11),
    ]
    for name, cin, cout, k, st, H, W in shapes:
        layer = R["Conv2dRT"](cin, cout, k, stride=st, prior={"mu": 0.0, "sigma": 1e-8})
        with torch.no_grad():
            for p in layer.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.1 if p.dim() > 1 else 0.1) - (3.0 if False else 0.0))
            layer.W_rho.copy_(-3 + 0.1 * torch.randn(layer.W_rho.shape, generator=g))
            layer.bias_rho.copy_(-3 + 0.1 * torch.randn(layer.bias_rho.shape, generator=g))
        x = torch.randn(1, cin, H, W, generator=g, requires_grad=True)
        eps_w = torch.randn(layer.W_mu.shape, generator=g)
        eps_b = torch.randn(layer.bias_mu.shape, generator=g)
        with EpsInjector() as inj:
            inj.queue += [eps_w, eps_b]
            y = layer(x)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        arrs.update({f"{name}/x": x, f"{name}/W_mu": layer.W_mu, f"{name}/W_rho": layer.W_rho,
                     f"{name}/bias_mu": layer.bias_mu, f"{name}/bias_rho": layer.bias_rho, f"{name}/eps_w": eps_w,
                     f"{name}/eps_b": eps_b, f"{name}/y": y, f"{name}/dy": dy, f"{name}/dx": x.grad,
                     f"{name}/dW_mu": layer.W_mu.grad, f"{name}/dW_rho": layer.W_rho.grad,
                     f"{name}/dbias_mu": layer.bias_mu.grad, f"{name}/dbias_rho": layer.bias_rho.grad,
                     f"{name}/meta": np.array([cin, cout, k, st, H, W])})
        # eval mode uses the mean weights (reparam_layers.py:33-35)
        layer.eval()
        arrs[f"{name}/y_eval"] = layer(x.detach())
    # LinearRT
    lin = R["LinearRT"](20, 12, prior={"mu": 0.0, "sigma": 1e-8})
    x = torch.randn(5, 20, generator=g, requires_grad=True)
    eps_w = torch.randn(lin.W_mu.shape, generator=g)
    eps_b = torch.randn(lin.bias_mu.shape, generator=g)
    with EpsInjector() as inj:
        inj.queue += [eps_w, eps_b]
        y = lin(x)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    arrs.update({"lin/x": x, "lin/W_mu": lin.W_mu, "lin/W_rho": lin.W_rho, "lin/bias_mu": lin.bias_mu,
                 "lin/bias_rho": lin.bias_rho, "lin/eps_w": eps_w, "lin/eps_b": eps_b, "lin/y": y, "lin/dy": dy,
                 "lin/dx": x.grad, "lin/dW_mu": lin.W_mu.grad, "lin/dW_rho": lin.W_rho.grad,
                 "lin/dbias_mu": lin.bias_mu.grad, "lin/dbias_rho": lin.bias_rho.grad})
    # KL: layer._kl for several priors, both kl types, with analytic grads through autograd
    for i, (temp, sigma, kl_type) in enumerate([(5.656911698337764e-07, 1.4616642493692077e-05, "reverse"),
                                                (1e-12, 6.506e-4, "reverse"), (1e-2, 0.5, "reverse"),
                                                (1e-2, 0.5, "forward")]):
        layer = R["Conv2dRT"](6, 5, 3, prior={"mu": 0.0, "sigma": np.sqrt(temp) * sigma}, kl_type=kl_type)
        with torch.no_grad():
            layer.W_mu.copy_(0.1 * torch.randn(layer.W_mu.shape, generator=g))
            layer.W_rho.copy_(-3 + 0.5 * torch.randn(layer.W_rho.shape, generator=g))
            layer.bias_mu.copy_(0.1 * torch.randn(layer.bias_mu.shape, generator=g))
            layer.bias_rho.copy_(-3 + 0.5 * torch.randn(layer.bias_rho.shape, generator=g))
        kl = layer._kl
        kl.backward()
        arrs.update({f"kl{i}/W_mu": layer.W_mu, f"kl{i}/W_rho": layer.W_rho, f"kl{i}/bias_mu": layer.bias_mu,
                     f"kl{i}/bias_rho": layer.bias_rho, f"kl{i}/kl": kl, f"kl{i}/dW_mu": layer.W_mu.grad,
                     f"kl{i}/dW_rho": layer.W_rho.grad, f"kl{i}/dbias_mu": layer.bias_mu.grad,
                     f"kl{i}/dbias_rho": layer.bias_rho.grad,
                     f"kl{i}/meta": np.array([temp, sigma, 1.0 if kl_type == "reverse" else 0.0])})
    # NLL
    mu = torch.rand(1, 1, 16, 16, generator=g, requires_grad=True)
    s = (8 * torch.randn(1, 1, 16, 16, generator=g)).requires_grad_(True)   # exercises the +-20 clamp
    t = torch.rand(1, 1, 16, 16, generator=g)
    v = R["nll"](mu, s, t)
    v.backward()
    arrs.update({"nll/mu": mu, "nll/s": s, "nll/t": t, "nll/v": v, "nll/dmu": mu.grad, "nll/ds": s.grad})
    mu3 = torch.randn(1, 3, 16, 16, generator=g, requires_grad=True)
    s1 = (8 * torch.randn(1, 1, 16, 16, generator=g)).requires_grad_(True)
    t3 = torch.rand(1, 3, 16, 16, generator=g)
    m = (torch.rand(1, 1, 16, 16, generator=g) < 0.5).float()
    v = R["nll_inp"](mu3.sigmoid(), s1, t3, m)
    v.backward()
    arrs.update({"nlli/mu": mu3, "nlli/s": s1, "nlli/t": t3, "nlli/m": m, "nlli/v": v, "nlli/dmu": mu3.grad,
                 "nlli/ds": s1.grad})
    # radon
    for name, n, theta in [("r32", 32, torch.arange(0, 180., step=36.)), ("r48", 48, torch.tensor([0., 10., 45., 90., 135.5])),
                           ("r40c2", 40, torch.arange(0, 180., step=60.))]:
        C = 2 if name.endswith("c2") else 1
        rad = R["Radon"]((1, C, n, n), theta)
        img = torch.rand(1, C, n, n, generator=g, requires_grad=True)
        sino = rad(img)
        ds = torch.randn(sino.shape, generator=g)
        sino.backward(ds)
        arrs.update({f"{name}/img": img, f"{name}/theta": theta, f"{name}/sino": sino, f"{name}/dsino": ds,
                     f"{name}/dimg": img.grad})
    # metrics
    a = torch.rand(1, 1, 40, 40, generator=g)
    b = (a + 0.1 * torch.randn(1, 1, 40, 40, generator=g)).clip(0, 1)
    arrs.update({"met/a": a, "met/b": b, "met/psnr": np.float64(R["psnr"](a, b)), "met/ssim": np.float64(R["ssim"](a, b))})
    err = torch.rand(40, 40, generator=g) * 0.01
    unc = torch.rand(40, 40, generator=g) * 0.02
    arrs.update({"met/err": err, "met/unc": unc, "met/uce": R["uce"](err, unc)[0]})
    save("layers.npz", **arrs)


def full_keys():
    """State-dict key lists + parameter counts of the 4 full-size nets (key/shape parity of our builder)."""
    out = {}
    for task, cfg in FULL.items():
        torch.manual_seed(1)
        net = build_ref_net(cfg, 1e-8)
        sd = net.state_dict()
        out[task] = {"keys": list(sd.keys()), "shapes": [list(v.shape) for v in sd.values()],
                     "n_params": int(sum(p.numel() for p in net.parameters())),
                     "param_sum": float(sum(p.double().sum() for p in net.parameters())),
                     "param_abs_sum": float(sum(p.double().abs().sum() for p in net.parameters()))}
    with open(os.path.join(HERE, "full_net_keys.json"), "w") as f:
        json.dump(out, f)
    print("wrote full_net_keys.json", {k: v["n_params"] for k, v in out.items()})


def full_den256(S=2):
    """Full-size 256^2 denoise net (the metric shape), philox eps, summary outputs only."""
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    cfg = FULL["den"]
    temp, sigma = 5.656911698337764e-07, 1.4616642493692077e-05
    torch.manual_seed(1)
    net = build_ref_net(cfg, np.sqrt(temp) * sigma)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    lay = O.skip_layout(cfg)
    target = torch.from_numpy(noisy(ellipse_phantom(256), 0.1, 1))[None]
    g = torch.Generator().manual_seed(5)
    net_input = torch.rand(1, 16, 256, 256, generator=g) * 0.1
    eps_list = make_eps(lay, sd, S, seed=1234, step=0, use_philox=True)
    loss, nll, kl, outs, grads = run_ref_step(net, cfg, lay, net_input, eps_list, "den", temp, {"target": target})
    arrs = {"loss": loss, "nll": nll, "kl": kl, "temp": np.float64(temp), "sigma": np.float64(sigma), "S": np.int64(S),
            "philox_seed": np.int64(1234), "input_seed": np.int64(5), "init_seed": np.int64(1)}
    for i, o in enumerate(outs):
        arrs[f"out{i}_sub"] = o[:, :, ::8, ::8]
        arrs[f"out{i}_mean"] = o.double().mean(dim=(0, 2, 3))
    names = list(grads.keys())
    arrs["grad_names"] = np.array(json.dumps(names))
    arrs["grad_norms"] = np.array([float(grads[k].double().norm()) for k in names])
    arrs["grad_absmax"] = np.array([float(grads[k].abs().max()) for k in names])
    arrs["param_norms"] = np.array([float(dict(net.named_parameters())[k].double().norm()) for k in names])
    for k in names:
        if "Conv2d_up_9." in k or "Conv2d_deeper_10." in k or "Conv2d_up_11." in k or "BatchNorm2d_up_9." in k:
            arrs["grad/" + k] = grads[k].reshape(-1)[:4096]
    save("den256_summary.npz", **arrs)


def full_summary(task, size, S, fname, seed=1234):
    """Full-size net of a BASELINE config (den 256^2 / sr, inp, ct 512^2), philox eps, summary outputs only:
    sub-sampled outputs, loss terms, the norm / abs-max of EVERY gradient tensor and the first 1024 elements of each.  The
    synthetic inputs are the ones bench.py --config <task> uses (mfvi_dip_mia_b200.utils.phantoms)."""
    from mfvi_dip_mia_b200.utils import phantoms as ph
    cfg = FULL[task]
    temp, sigma = {"den": (5.656911698337764e-07, 1.4616642493692077e-05), "sr": (4.3817e-07, 4.9e-08),
                   "inp": (1e-12, 6.506e-4), "ct": (2.2e-10, 1.7e-7)}[task]
    torch.manual_seed(1)
    net = build_ref_net(cfg, np.sqrt(temp) * sigma)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    lay = O.skip_layout(cfg)
    g = torch.Generator().manual_seed(5)
    net_input = torch.rand(1, cfg.num_input_channels, size, size, generator=g) * 0.1
    extra = {}
    if task == "den":
        extra["target"] = torch.from_numpy(ph.noisy(ph.ellipse_phantom(size), 0.1, 1))[None]
    elif task == "sr":
        extra["factor"] = 4
        extra["target"] = torch.from_numpy(ph.ellipse_phantom(size))[None][:, :, ::4, ::4].contiguous()
    elif task == "inp":
        extra["target"] = torch.from_numpy(ph.rgb_phantom(size))[None]
        extra["mask"] = torch.from_numpy(ph.random_mask(size, 2))[None]
    elif task == "ct":
        theta = torch.arange(0, 180., step=2.)                       # 90 angles (BASELINE config 4)
        extra["radon"] = R["Radon"]((1, 1, size, size), theta)
        extra["sino"] = extra["radon"](torch.from_numpy(ph.shepp_logan(size))[None]).detach()
    eps_list = make_eps(lay, sd, S, seed=seed, step=0, use_philox=True)
    loss, nll, kl, outs, grads = run_ref_step(net, cfg, lay, net_input, eps_list, task, temp, extra)
    arrs = {"loss": loss, "nll": nll, "kl": kl, "temp": np.float64(temp), "sigma": np.float64(sigma), "S": np.int64(S),
            "philox_seed": np.int64(seed), "input_seed": np.int64(5), "init_seed": np.int64(1), "size": np.int64(size),
            "task": np.array(task)}
    for i, o in enumerate(outs):
        arrs[f"out{i}_sub"] = o[:, :, ::8, ::8]
        arrs[f"out{i}_mean"] = o.double().mean(dim=(0, 2, 3))
    names = list(grads.keys())
    arrs["grad_names"] = np.array(json.dumps(names))
    arrs["grad_norms"] = np.array([float(grads[k].double().norm()) for k in names])
    arrs["grad_absmax"] = np.array([float(grads[k].abs().max()) for k in names])
    arrs["param_norms"] = np.array([float(dict(net.named_parameters())[k].double().norm()) for k in names])
    for k in names:
        arrs["grad/" + k] = grads[k].reshape(-1)[:1024]
    # the same step by the reference in float64: the yardstick for both sides.  Per-sample BatchNorm over the small maps of the
    # deep scales is badly conditioned, so the reference's OWN fp32 gradients are up to a few 1e-2 of a tensor's scale away from
    # its fp64 run on the smallest tensors — `ref_err32` records that distance per tensor (same denominators as the test uses).
    net.double()
    extra64 = {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in extra.items()}
    if task == "ct":
        extra64["radon"] = R["Radon"]((1, 1, size, size), theta).double()
    eps64 = [{k: v.double() for k, v in e.items()} for e in eps_list]
    loss64, nll64, kl64, outs64, grads64 = run_ref_step(net, cfg, lay, net_input.double(), eps64, task, temp, extra64)
    gmax = max(float(grads64[k].abs().max()) for k in names)
    ref_err = []
    for k in names:
        g64 = grads64[k].reshape(-1)[:1024]
        arrs["grad64/" + k] = g64
        den = max(float(grads64[k].abs().max()), 1e-3 * gmax)
        if k.endswith("bias_mu") or k.endswith("bias_rho"):
            wk = k.replace("bias_mu", "W_mu").replace("bias_rho", "W_rho")
            den = max(den, 0.05 * float(grads64[wk].abs().max()))
        ref_err.append(float((grads[k].reshape(-1)[:1024].double() - g64).abs().max()) / den)
    arrs["ref_err32"] = np.array(ref_err)
    arrs["grad64_absmax"] = np.array([float(grads64[k].abs().max()) for k in names])
    arrs["nll64"], arrs["kl64"] = nll64, kl64
    for i, o in enumerate(outs64):
        arrs[f"out{i}_sub64"] = o[:, :, ::8, ::8]
    print(f"  {fname}: reference fp32 vs fp64: worst tensor {max(ref_err):.2e} ({names[int(np.argmax(ref_err))].rsplit('.', 2)[-2:]}), "
          f"median {float(np.median(ref_err)):.2e}")
    save(fname, **arrs)


if __name__ == "__main__":
    what = sys.argv[1:] or ["small", "layers", "keys", "den256"]
    if "small" in what:
        for i, t in enumerate(["den", "sr", "ct", "inp"]):
            small_task(t, seed=20 + i)
    if "layers" in what:
        layer_fixtures()
    if "keys" in what:
        full_keys()
    if "den256" in what:
        full_den256()
    if "full" in what:          # BASELINE configs at full size: the metric shape at MC=8, configs 2-4 at 512^2 with one sample
        full_summary("den", 256, 8, "full_den256_s8.npz")
        full_summary("sr", 512, 1, "full_sr512.npz")
        full_summary("ct", 512, 1, "full_ct512.npz")
        full_summary("inp", 512, 1, "full_inp512.npz")
