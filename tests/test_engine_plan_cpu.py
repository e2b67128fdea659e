"""The engine's kernel plan, interpreted on the CPU (tests/plan_interpreter.py), against the reference fixtures.

A SkipEngine built with plan_only=True on the CPU holds the real plan — buffers, views, op lists — and cannot run it; the
interpreter executes every op with PyTorch as include/mfvi_dip.h defines it.  What is checked is the host side of the engine
(which buffer feeds which kernel, shapes, strides, paddings, accumulate flags, sample sharing of the input, the reparam chain
around the plan), for the exact-fp32 plan to 1e-4 of the imported reference."""
import pytest
import torch
import torch.nn.functional as F

from oracle import mfvi_oracle as O
from tests.helpers import grad_errs, group, load_npz, rel_err
from tests.plan_interpreter import PlanInterpreter

SMALL = {
    "den": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "sr": O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "ct": O.SkipCfg(4, 1, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "inp": O.SkipCfg(4, 4, (8, 16, 16), (8, 16, 16), (0, 0, 0), 5, 3, 1, False, False, "nearest"),
}


def _head(task, ex, S):
    """(S,C,H,W) network outputs -> mean over the samples of the task's data loss (oracle.mfvi_loss)."""
    per = {
        "den": lambda o: O.gaussian_nll(o[:, :1], o[:, 1:], ex["target"]),
        "sr": lambda o: (lambda lr: O.gaussian_nll(lr[:, :1], lr[:, 1:], ex["target"]))(O.sr_downsample_nearest(o, 4)),
        "inp": lambda o: O.gaussian_nll_inpainting(torch.sigmoid(o[:, :3]), o[:, 3:], ex["target"], ex["mask"]),
        "ct": lambda o: F.mse_loss(O.radon_forward(o, ex["theta"]), ex["sino"]),
    }[task]
    return lambda out: torch.stack([per(out[s:s + 1]) for s in range(S)]).mean()


def _run(task, math):
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec
    d = load_npz(f"skipnet_small_{task}.npz")
    S, sd, ex, grads = int(d["S"]), group(d, "sd/"), group(d, "extra/"), group(d, "grad/")
    eps = [group(d, f"eps{s}/") for s in range(S)]
    cfg = SMALL[task]
    spec = SkipSpec(cfg.num_input_channels, cfg.num_output_channels, tuple(cfg.down), tuple(cfg.up), tuple(cfg.skip),
                    cfg.filter_down, cfg.filter_up, cfg.filter_skip, cfg.need1x1_up, cfg.need_sigmoid, cfg.upsample_mode)
    x = torch.from_numpy(d["net_input"])
    eng = SkipEngine(spec, x.shape[2], x.shape[3], S, "cpu", math=math, plan_only=True)
    eng.load_params(sd, prefix="net.")
    eng.pack_eps(eps, prefix="net.")
    it = PlanInterpreter(eng)
    nll = it.step(x[0], _head(task, ex, S))
    # the reparameterisation chain and the tempered KL around the plan (mfvi_kl_reparam_fwd_bwd, include/mfvi_dip.h)
    temp, sigma = float(d["temp"]), float(d["sigma"])
    P = eng.lay.P
    mu, rho = eng.mu.clone().requires_grad_(True), eng.rho.clone().requires_grad_(True)
    kl = O.kl_elementwise(mu, rho, 0.0, O.prior_scale(temp, sigma)).sum()
    kl.backward()
    eng.g_mu.copy_(eng.dw[:, :P].sum(0) + temp * mu.grad)
    eng.g_rho.copy_(torch.sigmoid(eng.rho) * (eng.eps[:, :P] * eng.dw[:, :P]).sum(0) + temp * rho.grad)
    ours = {"net." + k: v.clone() for k, v in eng.param_views("grad").items()}
    out = eng.out.permute(0, 3, 1, 2)
    return d, S, out, nll, float(kl), ours, grads


@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_fp32_plan_reproduces_the_reference_step(task):
    from mfvi_dip_mia_b200 import _lib as L
    d, S, out, nll, kl, ours, grads = _run(task, L.MATH_FP32)
    for s in range(S):
        assert rel_err(out[s:s + 1], d[f"out{s}"]) < 1e-4, s
    assert rel_err(nll, d["nll"]) < 1e-5 and rel_err(kl, d["kl"]) < 1e-5
    errs = grad_errs({k: ours[k] for k in grads}, grads)
    worst = max(errs, key=errs.get)
    assert errs[worst] < 1e-3, (worst, errs[worst])


@pytest.mark.parametrize("math_name", ["fp32"])
def test_metric_network_plan_against_the_oracle(math_name):
    """The 5-scale, 16-channel-input network of the metric (bilinear upsampling, skip branches, 1x1 up convs) at 128x128 with
    random parameters: interpreted plan vs the oracle's autograd (no fixture: the oracle itself is pinned to the reference by
    tests/test_oracle_golden.py).  Per-sample BatchNorm over the 4x4 maps of the deepest scale is badly conditioned — the fp32
    oracle itself is 1e-3 away from its fp64 run on the smallest gradients — so fp64 is the yardstick and the fp32 plan has to
    stay within 4x the fp32 oracle's own error (+1e-3)."""
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L
    S, H = 1, 128
    g = torch.Generator().manual_seed(11)
    eng = SkipEngine(SkipSpec(), H, H, S, "cpu", math=L.MATH_FP32, plan_only=True)
    eng.mu.copy_(0.1 * torch.randn(eng.lay.P, generator=g))
    eng.rho.copy_(-3.0 + 0.1 * torch.randn(eng.lay.P, generator=g))
    eng.gamma.copy_(1.0 + 0.1 * torch.randn(eng.lay.Q, generator=g))
    eng.beta.copy_(0.1 * torch.randn(eng.lay.Q, generator=g))
    eps = []
    for _ in range(S):
        e = {}
        for c in eng.lay.convs:
            e["net." + c.key + ".W"] = torch.randn(c.cout, c.cin, c.k, c.k, generator=g)
            e["net." + c.key + ".b"] = torch.randn(c.cout, generator=g)
        eps.append(e)
    eng.pack_eps(eps, prefix="net.")
    x = torch.rand(1, 16, H, H, generator=g) * 0.1
    target = torch.rand(1, 1, H, H, generator=g)
    it = PlanInterpreter(eng)
    nll = it.step(x[0], lambda out: torch.stack([O.gaussian_nll(out[s:s + 1, :1], out[s:s + 1, 1:], target) for s in range(S)]).mean())
    P = eng.lay.P
    eng.g_mu.copy_(eng.dw[:, :P].sum(0))                         # data term only: the KL chain is not part of the plan
    eng.g_rho.copy_(torch.sigmoid(eng.rho) * (eng.eps[:, :P] * eng.dw[:, :P]).sum(0))
    ours = {"net." + k: v for k, v in eng.param_views("grad").items()}

    def oracle(dtype):
        sd = {"net." + k: v.detach().clone().to(dtype).requires_grad_("running" not in k) for k, v in eng.param_views("theta").items()}
        ee = [{k: v.to(dtype) for k, v in e.items()} for e in eps]
        _, nll_o, _, outs = O.mfvi_loss(sd, O.SkipCfg(16, 2), x.to(dtype), ee, task="den", temp=1.0, prior_sigma_plus_eps=1.0,
                                        target=target.to(dtype))
        nll_o.backward()
        return float(nll_o), [o.detach() for o in outs], {k: v.grad for k, v in sd.items() if v.grad is not None}

    nll64, outs64, g64 = oracle(torch.float64)
    out = eng.out.permute(0, 3, 1, 2)
    e_out = max(rel_err(out[s:s + 1], outs64[s]) for s in range(S))
    assert len(g64) == len(ours)
    if math_name == "fp32":
        _, _, g32 = oracle(torch.float32)
        assert e_out < 1e-4 and rel_err(nll, nll64) < 1e-5
        e_plan, e_o32 = grad_errs(ours, g64), grad_errs(g32, g64)
        for k in e_plan:
            assert e_plan[k] < 4 * e_o32[k] + 1e-3, (k, e_plan[k], e_o32[k])
