"""Bring-up tests of the bf16-operand convolutions (tcgen05 kind::f16; stage A of DESIGN.md section 8).  NOT part of the
`-m gpu` suite: these kernels are not on the product path yet and have not run on a GPU.  Run with
`MFVI_TEST_NEXT=1 python -m pytest tests -m gpu_next -x -q`.

Recipe: the inputs are rounded to bf16 FIRST and the exact-fp32 CUDA-core kernels run on the rounded values, so the two
results differ only by the fp32 summation order (products of two bf16 numbers are exact in fp32): the bar is 2e-5 of the
output's max — a layout or descriptor bug cannot hide behind a reduced-precision tolerance."""
import ctypes as C

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu_next
TOL = 2e-5

# (cin, cout, k, H, W[, stride]) as in test_gpu_tc.py; channel counts as the nets have them (36, 68, 132: concat layers)
SHAPES = [
    (64, 64, 3, 32, 32), (128, 128, 1, 16, 16), (32, 32, 1, 64, 64), (16, 16, 3, 64, 64), (36, 16, 3, 64, 64),
    (68, 32, 3, 32, 48), (132, 128, 3, 16, 16), (132, 64, 3, 64, 64), (16, 16, 5, 24, 24), (20, 24, 3, 37, 29),
    (16, 4, 1, 32, 32), (16, 2, 1, 64, 64), (36, 16, 3, 256, 256),
    (16, 16, 3, 128, 128, 2), (64, 128, 3, 16, 16, 2), (32, 64, 3, 31, 33, 2), (16, 16, 5, 32, 32, 2),
]


def _pitch8(c):
    return (c + 7) // 8 * 8


def _case(shape, S=2, seed=0):
    cin, cout, k, H, W = shape[:5]
    stride = shape[5] if len(shape) > 5 else 1
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    Hin, Win = (H - 1) * stride + k, (W - 1) * stride + k
    if stride == 2:
        Hin, Win = Hin + 1, Win + 1
    bf = lambda t: t.to(torch.bfloat16)
    # bf16 storage with 16-byte aligned pixels / weight rows (channel pitch rounded up to 8); the padding channels hold
    # garbage on purpose: the tensor maps must not read them
    xb = bf(torch.randn(S, Hin, Win, _pitch8(cin), device=dev, generator=g))
    dyb = bf(torch.randn(S, H, W, _pitch8(cout), device=dev, generator=g))
    wb = bf(torch.randn(S, k * k, cout, _pitch8(cin), device=dev, generator=g) * 0.1)
    bias = torch.randn(S, cout, device=dev, generator=g)
    return dict(cin=cin, cout=cout, k=k, H=H, W=W, stride=stride, Hin=Hin, Win=Win, S=S, xb=xb, dyb=dyb, wb=wb, bias=bias)


def _reference(c):
    """Exact-fp32 CUDA-core kernels on the bf16-rounded values."""
    from mfvi_dip_mia_b200 import _lib as L
    S, cin, cout, k = c["S"], c["cin"], c["cout"], c["k"]
    x = c["xb"][..., :cin].float().contiguous()
    dy = c["dyb"][..., :cout].float().contiguous()
    w = torch.cat([c["wb"][..., :cin].float().reshape(S, -1), c["bias"]], 1).contiguous()
    y = torch.zeros(S, c["H"], c["W"], cout, device=x.device)
    dx = torch.zeros(S, c["Hin"], c["Win"], cin, device=x.device)
    stats = torch.zeros(S, cout, 2, dtype=torch.float64, device=x.device)
    d = L.ConvDesc(S, cin, cout, k, k, c["stride"], c["Hin"], c["Win"], c["H"], c["W"], L.MATH_FP32)
    P, boff = w.shape[1], k * k * cout * cin
    L.call("mfvi_conv2d_fwd", C.byref(d), L.view(x), w.data_ptr(), w.data_ptr() + 4 * boff, P, L.view(y), stats.data_ptr())
    L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dy), w.data_ptr(), P, L.view(dx), 0)
    torch.cuda.synchronize()
    return y, dx, stats


def _wgrad(c, bf16):
    """dw (+ bias gradient) of the case: fp32 CUDA-core kernel on the rounded values, or the bf16 tensor-core kernel."""
    from mfvi_dip_mia_b200 import _lib as L
    S, cin, cout, k = c["S"], c["cin"], c["cout"], c["k"]
    dev = c["xb"].device
    P, boff = k * k * cout * cin + cout, k * k * cout * cin
    dw = torch.zeros(S, (P + 3) // 4 * 4, device=dev)
    dy32 = c["dyb"][..., :cout].float().contiguous()
    d = L.ConvDesc(S, cin, cout, k, k, c["stride"], c["Hin"], c["Win"], c["H"], c["W"], L.MATH_TF32 if bf16 else L.MATH_FP32)
    if bf16:
        L.call("mfvi_conv2d_wgrad_bf16", C.byref(d), L.view(c["xb"][..., :cin]), L.view(c["dyb"][..., :cout]), dw.data_ptr(),
               dw.stride(0), L.view(dy32), dw.data_ptr() + 4 * boff)
    else:
        x32 = c["xb"][..., :cin].float().contiguous()
        L.call("mfvi_conv2d_wgrad", C.byref(d), L.view(x32), L.view(dy32), dw.data_ptr(), dw.data_ptr() + 4 * boff, dw.stride(0))
    torch.cuda.synchronize()
    return dw[:, :P]


def _bf16(c, accumulate=False):
    from mfvi_dip_mia_b200 import _lib as L
    S, cin, cout, k = c["S"], c["cin"], c["cout"], c["k"]
    dev = c["xb"].device
    y = torch.zeros(S, c["H"], c["W"], cout, device=dev)
    dx = torch.full((S, c["Hin"], c["Win"], cin), 1.0 if accumulate else 0.0, device=dev)
    stats = torch.zeros(S, cout, 2, dtype=torch.float64, device=dev)
    d = L.ConvDesc(S, cin, cout, k, k, c["stride"], c["Hin"], c["Win"], c["H"], c["W"], L.MATH_TF32)
    wb, cp = c["wb"], _pitch8(cin)
    L.call("mfvi_conv2d_fwd_bf16", C.byref(d), L.view(c["xb"][..., :cin]), wb.data_ptr(), cp, wb.stride(0), c["bias"].data_ptr(),
           c["bias"].stride(0), L.view(y), stats.data_ptr())
    L.call("mfvi_conv2d_dgrad_bf16", C.byref(d), L.view(c["dyb"][..., :cout]), wb.data_ptr(), cp, wb.stride(0), L.view(dx),
           1 if accumulate else 0)
    torch.cuda.synchronize()
    return y, dx, stats


@pytest.mark.parametrize("shape", SHAPES)
def test_bf16_conv_equals_fp32_kernels_on_rounded_inputs(shape):
    c = _case(shape)
    ref = _reference(c)
    got = _bf16(c)
    for n, a, b in zip(["y", "dx", "bn stats"], got, ref):
        assert torch.isfinite(a).all(), (shape, n)
        assert rel_err(a, b) < TOL, (shape, n, rel_err(a, b))


@pytest.mark.parametrize("shape", SHAPES)
def test_bf16_wgrad_equals_fp32_kernel_on_rounded_inputs(shape):
    """Sums over up to 65536 pixels in a different order (split-K atomics on both sides): 1e-4 of the largest entry."""
    c = _case(shape)
    ref, got = _wgrad(c, False), _wgrad(c, True)
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 1e-4, (shape, rel_err(got, ref))


def test_bf16_dgrad_accumulates():
    c = _case((36, 16, 3, 64, 64))
    _, dx0, _ = _bf16(c)
    _, dx1, _ = _bf16(c, accumulate=True)
    assert rel_err(dx1 - 1.0, dx0) < TOL


def test_bf16_conv_rejects_unaligned_views():
    from mfvi_dip_mia_b200 import _lib as L
    c = _case((16, 16, 3, 16, 16))
    bad = c["xb"][..., 1:13]                                  # pixels no longer 16-byte aligned
    d = L.ConvDesc(c["S"], 12, 16, 3, 3, 1, c["Hin"], c["Win"], c["H"], c["W"], L.MATH_TF32)
    y = torch.zeros(c["S"], c["H"], c["W"], 16, device=bad.device)
    with pytest.raises(L.MfviError, match="16-byte aligned"):
        L.call("mfvi_conv2d_fwd_bf16", C.byref(d), L.view(bad), c["wb"].data_ptr(), 16, c["wb"].stride(0), None, 0, L.view(y), None)


# ---------------------------------------------------------------------------------------------------------------- stage C
def _padded_bf16(S, H, W, Cn, dev):
    """bf16 NHWC buffer with the channel pitch rounded up to 8 (poisoned), and its first Cn channels as the view."""
    full = torch.full((S, H, W, _pitch8(Cn)), float("nan"), dtype=torch.bfloat16, device=dev)
    return full, full[..., :Cn]


@pytest.mark.parametrize("Cn,H,W,pad,act", [(16, 32, 32, 1, 1), (36, 17, 23, 1, 0), (128, 8, 8, 0, 1), (20, 16, 16, 2, 1), (2, 9, 9, 1, 1)])
def test_bn_act_pad_fwd_bf16_is_the_rounded_fp32_kernel(Cn, H, W, pad, act):
    from mfvi_dip_mia_b200 import _lib as L
    dev, S = torch.device("cuda:0"), 3
    g = torch.Generator(device=dev).manual_seed(1)
    y = torch.randn(S, H, W, Cn, device=dev, generator=g)
    sums = torch.stack([y.double().sum((1, 2)), (y.double() ** 2).sum((1, 2))], -1).contiguous()      # [S][C][2]
    gamma, beta = torch.rand(Cn, device=dev, generator=g) + 0.5, torch.randn(Cn, device=dev, generator=g)
    x32 = torch.zeros(S, H + 2 * pad, W + 2 * pad, Cn, device=dev)
    full, x16 = _padded_bf16(S, H + 2 * pad, W + 2 * pad, Cn, dev)
    args = (L.view(y), S, H, W, Cn, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, pad)
    L.call("mfvi_bn_act_pad_fwd", *args, L.view(x32))
    L.call("mfvi_bn_act_pad_fwd_bf16", *args, L.view(x16))
    torch.cuda.synchronize()
    assert torch.equal(x16, x32.to(torch.bfloat16))
    assert torch.isnan(full[..., Cn:]).all()                   # the pitch padding is never written


@pytest.mark.parametrize("Cn,H,W", [(16, 32, 32), (36, 17, 23), (4, 64, 64), (128, 8, 8)])
def test_bn_bwd_apply_bf16_is_the_rounded_fp32_kernel(Cn, H, W):
    from mfvi_dip_mia_b200 import _lib as L
    dev, S = torch.device("cuda:0"), 2
    gen = torch.Generator(device=dev).manual_seed(2)
    y, g = torch.randn(S, H, W, Cn, device=dev, generator=gen), torch.randn(S, H, W, Cn, device=dev, generator=gen)
    sums = torch.stack([y.double().sum((1, 2)), (y.double() ** 2).sum((1, 2))], -1).contiguous()
    mean = sums[..., 0] / (H * W)
    invstd = 1.0 / torch.sqrt(sums[..., 1] / (H * W) - mean ** 2 + 1e-5)
    xhat = (y.double() - mean[:, None, None]) * invstd[:, None, None]
    red = torch.stack([g.double().sum((1, 2)), (g.double() * xhat).sum((1, 2))], -1).contiguous()
    gamma = torch.rand(Cn, device=dev, generator=gen) + 0.5
    dgam, dbet = torch.zeros(Cn, device=dev), torch.zeros(Cn, device=dev)
    dy32 = torch.zeros_like(g)
    _, dy16 = _padded_bf16(S, H, W, Cn, dev)
    args = (L.view(g), L.view(y), S, H, W, Cn, sums.data_ptr(), red.data_ptr(), gamma.data_ptr())
    L.call("mfvi_bn_bwd_apply", *args, L.view(dy32), dgam.data_ptr(), dbet.data_ptr())
    L.call("mfvi_bn_bwd_apply_bf16", *args, L.view(dy16), dgam.data_ptr(), dbet.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(dy16, dy32.to(torch.bfloat16))


def test_view_and_weight_conversions_round_like_torch():
    from mfvi_dip_mia_b200 import _lib as L
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    src = torch.randn(2, 10, 12, 20, device=dev, generator=gen)[:, 1:9, 2:11, :18]        # a strided interior view
    full, dst = _padded_bf16(2, 8, 9, 18, dev)
    L.call("mfvi_view_f32_to_bf16", L.view(src), 2, 8, 9, 18, L.view(dst))
    # two layers: 3x3 36->16 and 1x1 16->4, fp32 blocks back to back, bf16 blocks with rows of 40 / 16 channels
    S, layers = 3, [(9 * 16, 36), (4, 16)]
    P = sum(r * c for r, c in layers)
    P16 = sum(r * _pitch8(c) for r, c in layers)
    w = torch.randn(S, P + 5, device=dev, generator=gen)
    w16 = torch.full((S, P16 + 8), float("nan"), dtype=torch.bfloat16, device=dev)
    w_off, w16_off, o, o16 = [], [], 0, 0
    for r, c in layers:
        w_off.append(o); w16_off.append(o16)
        o += r * c; o16 += r * _pitch8(c)
    arr = lambda v, t: (t * len(v))(*v)
    L.call("mfvi_pack_weights_bf16", w.data_ptr(), w.stride(0), S, len(layers), arr(w_off, C.c_longlong), arr(w16_off, C.c_longlong),
           arr([r for r, _ in layers], C.c_int), arr([c for _, c in layers], C.c_int), w16.data_ptr(), w16.stride(0))
    torch.cuda.synchronize()
    assert torch.equal(dst, src.to(torch.bfloat16)) and torch.isnan(full[..., 18:]).all()
    for (r, c), a, b in zip(layers, w_off, w16_off):
        got = w16[:, b:b + r * _pitch8(c)].view(S, r, _pitch8(c))
        assert torch.equal(got[..., :c], w[:, a:a + r * c].view(S, r, c).to(torch.bfloat16))
        assert (got[..., c:] == 0).all()
    assert torch.isnan(w16[:, P16:]).all()


# ---------------------------------------------------------------------------------------------------------------- stage D
@pytest.mark.parametrize("task", ["den", "inp", "sr", "ct"])
def test_bf16_engine_step_equals_its_cpu_interpretation(task):
    """Whole step in the bf16-operand mode on the GPU against the SAME plan interpreted on the CPU with the kernels' rounding
    points (tests/plan_interpreter.py; tests/test_engine_plan_cpu.py holds that interpretation to the reference within bf16
    accuracy).  Both sides round the same values to bf16, so they differ only where an fp32 summation-order difference flips
    a rounding: the bars are far below the bf16 error itself (output 2-5e-2, gradient 3-18 %), and a kernel bug cannot hide
    behind a reduced-precision tolerance."""
    from mfvi_dip_mia_b200 import SkipEngine, _lib as L
    from mfvi_dip_mia_b200.engine import NLL
    from mfvi_dip_mia_b200.trainer import LossHead
    from oracle import mfvi_oracle as O
    from tests.test_engine_plan_cpu import _run
    from tests.test_gpu_parity import SMALL, _fixture, _head_kwargs, spec_of
    dev = torch.device("cuda:0")
    d, S, sd, eps, ex, grads = _fixture(task)
    x = torch.from_numpy(d["net_input"])
    eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev, math=L.MATH_BF16)
    eng.load_params(sd, prefix="net.")
    eng.pack_eps(eps, prefix="net.")
    head = LossHead(eng, task, **_head_kwargs(task, ex))
    temp, sigma = float(d["temp"]), float(d["sigma"])
    eng.zero_accumulators()
    eng.set_input(x[0].permute(1, 2, 0).contiguous().to(dev), None, 0.0, L.key(0))
    eng.sample_weights(L.key(0))
    eng.forward()
    head.run()
    eng.backward()
    eng.reparam_kl(L.key(0), prior_mu=0.0, prior_sigma_plus_eps=O.prior_scale(temp, sigma), direction=0, kscale=temp)
    out = eng.out_nchw().cpu()
    nll = float(eng.arena[:2].cpu()[NLL])
    ours = {"net." + k: v.cpu() for k, v in eng.param_views("grad").items()}
    _, _, out_i, nll_i, _, interp, _ = _run(task, L.MATH_BF16)              # the CPU interpretation of the same plan
    e_out = rel_err(out, out_i)
    va = torch.cat([ours[k].double().reshape(-1) for k in grads])
    vb = torch.cat([interp[k].double().reshape(-1) for k in grads])
    e_l2, cos = float((va - vb).norm() / vb.norm()), float((va @ vb) / (va.norm() * vb.norm()))
    print(f"bf16 {task} GPU vs CPU interpretation: out {e_out:.2e}  nll {rel_err(nll, nll_i):.2e}  grad relL2 {e_l2:.2e}  cos {cos:.6f}")
    assert torch.isfinite(va).all()
    assert e_out < 5e-3 and rel_err(nll, nll_i) < 1e-3 and e_l2 < 2e-2 and cos > 0.9995


def test_bf16_trainer_runs_graph_replayed_steps():
    """The graph-captured trainer step in bf16 mode: finite, and the loss falls on a small denoising problem."""
    from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec, _lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    spec = SkipSpec(8, 2, (8, 16, 16), (8, 16, 16), (4, 4, 4), 3, 3, 1, True, False, "bilinear")
    x = torch.rand(1, 8, 64, 64, generator=g) * 0.1
    target = torch.rand(1, 1, 64, 64, generator=g)
    tr = MfviDipTrainer(spec, "den", x, temp=5.6e-7, sigma=1.5e-5, lr=1e-2, mc_samples=4, seed=3, device=dev, target=target,
                        math_mode=L.MATH_BF16)
    losses = []
    for i in range(60):
        tr.step()
        if i % 20 == 19:
            losses.append(tr.loss_terms()[0])
    assert all(map(lambda v: v == v and abs(v) < 1e6, losses)) and losses[-1] < losses[0]


# ------------------------------------------------------------------------------------------ fused BatchNorm backward (fp32)
@pytest.mark.parametrize("Cn,H,W,pad,act", [(16, 32, 32, 1, 1), (36, 17, 23, 1, 0), (128, 8, 8, 0, 1), (20, 16, 16, 2, 1), (2, 9, 9, 1, 1)])
def test_fused_bn_backward_equals_the_standard_pair(Cn, H, W, pad, act):
    """mfvi_pad_act_bwd_reduce + mfvi_bn_bwd_apply_from_dxp against mfvi_pad_act_bwd + mfvi_bn_bwd_apply on the same inputs:
    the same loads and the same arithmetic, so the results agree to fp32 rounding (1e-6 of the max), and the bf16 variant is the
    rounded fp32 result up to one bf16 ulp."""
    from mfvi_dip_mia_b200 import _lib as L
    dev, S = torch.device("cuda:0"), 3
    gen = torch.Generator(device=dev).manual_seed(4)
    y = torch.randn(S, H, W, Cn, device=dev, generator=gen)
    dxp = torch.randn(S, H + 2 * pad, W + 2 * pad, Cn, device=dev, generator=gen)
    sums = torch.stack([y.double().sum((1, 2)), (y.double() ** 2).sum((1, 2))], -1).contiguous()
    gamma, beta = torch.rand(Cn, device=dev, generator=gen) + 0.5, torch.randn(Cn, device=dev, generator=gen)
    z = lambda: torch.zeros(Cn, device=dev)
    # standard pair
    g, red0, dy0, dg0, db0 = torch.zeros_like(y), torch.zeros(S, Cn, 2, dtype=torch.float64, device=dev), torch.zeros_like(y), z(), z()
    L.call("mfvi_pad_act_bwd", L.view(dxp), S, H, W, Cn, pad, L.view(y), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act,
           L.view(g), red0.data_ptr())
    L.call("mfvi_bn_bwd_apply", L.view(g), L.view(y), S, H, W, Cn, sums.data_ptr(), red0.data_ptr(), gamma.data_ptr(), L.view(dy0),
           dg0.data_ptr(), db0.data_ptr())
    # fused pair
    red1, dy1, dg1, db1 = torch.zeros_like(red0), torch.zeros_like(y), z(), z()
    _, dy16 = _padded_bf16(S, H, W, Cn, dev)
    L.call("mfvi_pad_act_bwd_reduce", L.view(dxp), S, H, W, Cn, pad, L.view(y), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
           act, red1.data_ptr())
    tail = (L.view(dxp), L.view(y), S, H, W, Cn, pad, sums.data_ptr(), red1.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act)
    L.call("mfvi_bn_bwd_apply_from_dxp", *tail, L.view(dy1), dg1.data_ptr(), db1.data_ptr())
    L.call("mfvi_bn_bwd_apply_from_dxp_bf16", *tail, L.view(dy16), dg1.data_ptr(), db1.data_ptr())
    torch.cuda.synchronize()
    assert rel_err(red1, red0) < 1e-12 or torch.allclose(red1, red0, rtol=1e-9, atol=1e-9)
    assert rel_err(dy1, dy0) < 1e-6 and rel_err(dg1, dg0) < 1e-6 and rel_err(db1, db0) < 1e-6
    assert rel_err(dy16.float(), dy0) < 2 ** -8


# ---------------------------------------------------------------------------------------------------------- acceptance
def test_bf16_final_metrics_match_reference_ensemble():
    """The mode's acceptance test, the bf16 twin of test_gpu_parity.py::test_den_final_metrics_match_reference_ensemble: final
    PSNR / SSIM / UCE of the denoising runner after 1200 iterations over 32 seeds against the imported reference's own
    distribution (tests/golden/trajectory_den64.npz) — north_star's 0.1 dB / 0.005 bar on the ensemble means plus three
    standard errors.  The CPU study (tests/studies/bf16_quality_study.py) predicts a shift of +0.09 +- 0.17 dB."""
    import numpy as np
    from mfvi_dip_mia_b200 import _lib as L
    from mfvi_dip_mia_b200.runners import run_den_mfvi
    from mfvi_dip_mia_b200.utils.phantoms import ellipse_phantom, noisy
    from tests.helpers import load_npz
    from tests.test_gpu_parity import SMALL, _TRAJ, spec_of
    d = load_npz("trajectory_den64.npz")
    its = [int(v) for v in d["its"]]
    ref = np.asarray(d["metrics"])[:, its.index(1200)]
    gt = ellipse_phantom(_TRAJ["H"])
    ny = noisy(gt, 0.1, 1)
    ours = []
    for k in range(32):
        _, h = run_den_mfvi(gt, temp=_TRAJ["temp"], sigma=_TRAJ["sigma"], lr=_TRAJ["lr"], num_iter=1199, mc_samples=1,
                            seed=1000 + k, device=torch.device("cuda:0"), mc_ring=_TRAJ["ring"], show_every=10 ** 9,
                            math_mode=L.MATH_BF16, return_history=True, spec=spec_of(SMALL["den"]), img_noisy=ny)
        ours.append((h["psnr_gt_sm"][-1], h["ssim_gt_sm"][-1], h["uce"]))
    ours = np.array(ours)
    mo, mr = ours.mean(0), ref.mean(0)
    se = np.sqrt(ours.var(0, ddof=1) / len(ours) + ref.var(0, ddof=1) / len(ref))
    print(f"[ensemble bf16] ours PSNR {mo[0]:.3f} SSIM {mo[1]:.4f} UCE {mo[2]:.4f}   ref PSNR {mr[0]:.3f} SSIM {mr[1]:.4f} UCE {mr[2]:.4f}   se {se}")
    assert np.all(np.abs(mo - mr) < np.array([0.1, 0.005, 0.005]) + 3 * se), (mo, mr, se)
