"""The persistent multi-stage kernel (csrc/mega.cu, include/mfvi_dip.h mfvi_mega_*) against the stand-alone kernels.

1. Every convolution pass as a ONE-stage program against the exact-fp32 CUDA-core kernel of the library on the layer shapes
   of the coarse scales: the 3xTF32 mode (desc.math = fp32) to 2e-5 (fp32 accuracy on the tensor cores), the tf32 mode to
   the tf32 bar of tests/test_gpu_tc.py.
2. Every elementwise op as a one-stage program against its stand-alone kernel: the program runs the SAME body over virtual
   block indices, so outputs are bit-identical and the atomically accumulated statistics agree to double rounding.
3. The whole engine with the coarse scales fused (mega_from = 1, 2: the down path of that scale and everything below it as one
   launch per direction, one thread-block cluster per MC sample) against the reference fixtures in the exact-fp32 mode (the
   same bars as the op-by-op plan: 1e-3 per tensor), and against the op-by-op plan in the tf32 mode.
"""
import ctypes as C

import pytest
import torch

from tests.helpers import grad_errs, rel_err

pytestmark = pytest.mark.gpu

# (cin, cout, k, Hout, Wout, stride)
CONV_SHAPES = [(128, 128, 3, 8, 8, 1), (128, 128, 3, 8, 8, 2), (132, 128, 3, 16, 16, 1), (128, 128, 1, 16, 16, 1), (128, 4, 1, 16, 16, 1),
               (64, 128, 3, 16, 16, 2), (64, 64, 3, 32, 32, 1), (132, 128, 3, 32, 32, 1), (18, 16, 3, 8, 8, 1), (16, 16, 5, 12, 20, 2),
               (10, 6, 3, 9, 7, 1), (16, 2, 1, 24, 24, 1)]


def _one_stage(S, name, *args):
    """Record ONE op as a program and run it for S samples (one 8-CTA cluster per sample)."""
    from mfvi_dip_mia_b200 import _lib as L
    m = L.record_program([(name, args, {})], torch.device("cuda:0"))
    assert m["n_stages"] >= 1
    L.call("mfvi_mega_run", m["prog"].data_ptr(), m["n_stages"], S, None)
    torch.cuda.synchronize()
    return m


def _standalone(S, name, *args):
    from mfvi_dip_mia_b200 import _lib as L
    L.call(name, *args)


@pytest.mark.parametrize("math", ["fp32", "tf32"])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_mega_conv_stage_matches_fp32_kernels(shape, math):
    from mfvi_dip_mia_b200 import _lib as L
    cin, cout, k, H, W, stride = shape
    S = 3
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    Hin, Win = (H - 1) * stride + k, (W - 1) * stride + k
    if stride == 2:
        Hin, Win = Hin + 1, Win + 1
    cp = lambda c: (c + 3) // 4 * 4                       # channel pitch like the engine's buffers
    x = torch.randn(S, Hin, Win, cp(cin), device=dev, generator=g)[..., :cin]
    P = k * k * cout * cin + cout
    Pp = (P + 3) // 4 * 4
    w = torch.randn(S, Pp, device=dev, generator=g) * 0.1
    dy = torch.randn(S, H, W, cp(cout), device=dev, generator=g)[..., :cout]
    boff = k * k * cout * cin
    tol = 2e-5 if math == "fp32" else 3e-3
    d_ref = L.ConvDesc(S, cin, cout, k, k, stride, Hin, Win, H, W, L.MATH_FP32)
    d = L.ConvDesc(S, cin, cout, k, k, stride, Hin, Win, H, W, L.MATH_FP32 if math == "fp32" else L.MATH_TF32)

    def outputs():
        return (torch.zeros(S, H, W, cp(cout), device=dev)[..., :cout], torch.zeros(S, Hin, Win, cp(cin), device=dev)[..., :cin],
                torch.zeros_like(w), torch.zeros(S, cout, 2, dtype=torch.float64, device=dev))
    y0, dx0, dw0, st0 = outputs()
    # reference: the exact-fp32 CUDA-core kernels (conv_simt.cu), called directly
    L.call("mfvi_conv2d_fwd_simt", C.byref(d_ref), L.view(x), w.data_ptr(), w.data_ptr() + 4 * boff, Pp, L.view(y0), st0.data_ptr())
    L.call("mfvi_conv2d_dgrad_simt", C.byref(d_ref), L.view(dy), w.data_ptr(), Pp, L.view(dx0), 0)
    L.call("mfvi_conv2d_wgrad_simt", C.byref(d_ref), L.view(x), L.view(dy), dw0.data_ptr(), None, Pp)
    # the same tiles as stand-alone launches (mfvi_conv2d_*_mma; the fp32 mode dispatches to them with MFVI_FP32_MMA=1)
    if math == "fp32":
        y2, dx2_, dw2, st2 = outputs()
        L.call("mfvi_conv2d_fwd_mma", C.byref(d), L.view(x), w.data_ptr(), w.data_ptr() + 4 * boff, Pp, L.view(y2), st2.data_ptr())
        L.call("mfvi_conv2d_dgrad_mma", C.byref(d), L.view(dy), w.data_ptr(), Pp, L.view(dx2_), 0)
        L.call("mfvi_conv2d_wgrad_mma", C.byref(d), L.view(x), L.view(dy), dw2.data_ptr(), dw2.data_ptr() + 4 * boff, Pp)
        dw_ref = dw0.clone()
        dw_ref[:, boff:boff + cout] = dy.sum((1, 2))
        for n, a, b in (("y", y2, y0), ("stats", st2, st0), ("dx", dx2_, dx0), ("dw+dbias", dw2, dw_ref)):
            assert rel_err(a, b) < tol, ("stand-alone", shape, n, rel_err(a, b))
    y1, dx1, dw1, st1 = outputs()
    _one_stage(S, "mfvi_conv2d_fwd", C.byref(d), L.view(x), w.data_ptr(), w.data_ptr() + 4 * boff, Pp, L.view(y1), st1.data_ptr())
    _one_stage(S, "mfvi_conv2d_dgrad", C.byref(d), L.view(dy), w.data_ptr(), Pp, L.view(dx1), 0)
    _one_stage(S, "mfvi_conv2d_wgrad", C.byref(d), L.view(x), L.view(dy), dw1.data_ptr(), None, Pp)
    for n, a, b in (("y", y1, y0), ("stats", st1, st0), ("dx", dx1, dx0), ("dw", dw1, dw0)):
        assert torch.isfinite(a).all(), n
        assert rel_err(a, b) < tol, (shape, math, n, rel_err(a, b))
    # accumulate flag of the data gradient
    dx2 = dx0.clone()
    _one_stage(S, "mfvi_conv2d_dgrad", C.byref(d), L.view(dy), w.data_ptr(), Pp, L.view(dx2), 1)
    assert rel_err(dx2, 2 * dx0) < tol


@pytest.mark.parametrize("C_,H,W,pad", [(128, 8, 8, 1), (132, 16, 16, 1), (64, 32, 32, 1), (18, 8, 12, 1), (16, 16, 16, 2), (4, 16, 16, 0)])
def test_mega_elementwise_stages_are_the_standalone_kernels(C_, H, W, pad):
    from mfvi_dip_mia_b200 import _lib as L
    S = 3
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2)
    cp = (C_ + 3) // 4 * 4
    y = torch.randn(S, H, W, cp, device=dev, generator=g)[..., :C_]
    sums = torch.stack([y.double().sum((1, 2)), (y.double() ** 2).sum((1, 2))], -1).contiguous()      # (S,C,2)
    gamma = 1 + 0.1 * torch.randn(C_, device=dev, generator=g)
    beta = 0.1 * torch.randn(C_, device=dev, generator=g)
    # bn_act_pad_fwd
    xp = [torch.zeros(S, H + 2 * pad, W + 2 * pad, cp, device=dev)[..., :C_] for _ in range(2)]
    args = lambda o: (L.view(y), S, H, W, C_, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, pad, L.view(o))
    L.call("mfvi_bn_act_pad_fwd", *args(xp[0]))
    _one_stage(S, "mfvi_bn_act_pad_fwd", *args(xp[1]))
    assert torch.equal(xp[0], xp[1])
    # pad_act_bwd + bn_bwd_apply
    dxp = torch.randn(S, H + 2 * pad, W + 2 * pad, cp, device=dev, generator=g)[..., :C_]
    outs = []
    for run in (_standalone, _one_stage):
        gbuf = torch.zeros(S, H, W, cp, device=dev)[..., :C_]
        red = torch.zeros(S, C_, 2, dtype=torch.float64, device=dev)
        dg, db = torch.zeros(C_, device=dev), torch.zeros(C_, device=dev)
        run(S, "mfvi_pad_act_bwd", L.view(dxp), S, H, W, C_, pad, L.view(y), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1,
            L.view(gbuf), red.data_ptr())
        run(S, "mfvi_bn_bwd_apply", L.view(gbuf), L.view(y), S, H, W, C_, sums.data_ptr(), red.data_ptr(), gamma.data_ptr(),
            L.view(gbuf), dg.data_ptr(), db.data_ptr())
        if run is _one_stage:        # a program leaves the affine-parameter gradients (a sum over all samples) to this kernel
            assert float(dg.abs().sum()) == 0.0
            i64 = lambda v: torch.tensor([v], dtype=torch.int64, device=dev)
            tabs = (i64(red.data_ptr()), i64(dg.data_ptr()), i64(db.data_ptr()), torch.tensor([C_], dtype=torch.int32, device=dev))
            L.call("mfvi_bn_param_grads", tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(), tabs[3].data_ptr(), 1, S)
        torch.cuda.synchronize()
        outs.append((gbuf, red, dg, db))
    assert rel_err(outs[1][1], outs[0][1]) < 1e-6                    # another virtual grid: fp32 partial sums grouped differently
    for a, b in zip(outs[1][0:1] + outs[1][2:], outs[0][0:1] + outs[0][2:]):
        assert rel_err(a, b) < 1e-6


@pytest.mark.parametrize("Cs,Cd,H,W,mode", [(4, 128, 16, 16, 0), (2, 16, 8, 12, 0), (0, 128, 16, 16, 1), (4, 64, 32, 32, 0)])
def test_mega_concat_stages_are_the_standalone_kernels(Cs, Cd, H, W, mode):
    from mfvi_dip_mia_b200 import _lib as L
    S = 2
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    cp = lambda c: max((c + 3) // 4 * 4, 4)
    ys = torch.randn(S, H, W, cp(Cs), device=dev, generator=g)[..., :Cs] if Cs else None
    yd = torch.randn(S, H // 2, W // 2, cp(Cd), device=dev, generator=g)[..., :Cd]
    stat = lambda t: torch.stack([t.double().sum((1, 2)), (t.double() ** 2).sum((1, 2))], -1).contiguous()
    ss, sd = (stat(ys) if Cs else None), stat(yd)
    gs_, bs_ = (1 + 0.1 * torch.randn(max(Cs, 1), device=dev, generator=g)), 0.1 * torch.randn(max(Cs, 1), device=dev, generator=g)
    gd_, bd_ = (1 + 0.1 * torch.randn(Cd, device=dev, generator=g)), 0.1 * torch.randn(Cd, device=dev, generator=g)
    null = L.View(None, 0, 0, 0)
    sargs = (L.view(ys), Cs, ss.data_ptr(), gs_.data_ptr(), bs_.data_ptr()) if Cs else (null, 0, None, None, None)
    res = []
    for run in (_standalone, _one_stage):
        A = torch.zeros(S, H, W, cp(Cs + Cd), device=dev)[..., :Cs + Cd]
        sA = torch.zeros(S, Cs + Cd, 2, dtype=torch.float64, device=dev)
        run(S, "mfvi_cat_up_fwd", *sargs, L.view(yd), Cd, sd.data_ptr(), gd_.data_ptr(), bd_.data_ptr(), S, H, W, mode, L.view(A),
            sA.data_ptr())
        torch.cuda.synchronize()
        res.append((A, sA))
    assert torch.equal(res[0][0], res[1][0]) and rel_err(res[1][1], res[0][1]) < 1e-6
    dA = torch.randn(S, H, W, cp(Cs + Cd), device=dev, generator=g)[..., :Cs + Cd]
    res = []
    for run in (_standalone, _one_stage):
        gs = torch.zeros(S, H, W, cp(Cs), device=dev)[..., :Cs] if Cs else None
        gd = torch.zeros(S, H // 2, W // 2, cp(Cd), device=dev)[..., :Cd]
        rs = torch.zeros(S, max(Cs, 1), 2, dtype=torch.float64, device=dev)
        rd = torch.zeros(S, Cd, 2, dtype=torch.float64, device=dev)
        cat = (L.view(dA), S, H, W, mode, sargs[0], Cs, *sargs[2:], L.view(gs) if Cs else null, rs.data_ptr() if Cs else None,
               L.view(yd), Cd, sd.data_ptr(), gd_.data_ptr(), bd_.data_ptr(), L.view(gd), rd.data_ptr())
        run(S, "mfvi_cat_up_bwd", *cat, 2)
        if Cs:
            run(S, "mfvi_cat_up_bwd", *cat, 1)
        torch.cuda.synchronize()
        res.append((gd, rd, gs, rs))
    assert torch.equal(res[0][0], res[1][0]) and rel_err(res[1][1], res[0][1]) < 1e-6
    if Cs:
        assert torch.equal(res[0][2], res[1][2]) and rel_err(res[1][3], res[0][3]) < 1e-6


def _step(task, mega_from, math, S_override=None):
    from mfvi_dip_mia_b200 import SkipEngine, _lib as L
    from mfvi_dip_mia_b200.engine import KL, NLL
    from mfvi_dip_mia_b200.trainer import LossHead
    from oracle import mfvi_oracle as O
    from tests.test_gpu_parity import SMALL, _fixture, _head_kwargs, spec_of
    dev = torch.device("cuda:0")
    d, S, sd, eps, ex, grads = _fixture(task)
    x = torch.from_numpy(d["net_input"])
    eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev, math=math, mega_from=mega_from)
    assert (eng.mega_from is not None) == (mega_from >= 0)
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
    torch.cuda.synchronize()
    names = [op[0] for op in eng.fwd_ops + eng.bwd_ops]
    return d, grads, eng.out_nchw().cpu(), eng.arena[:2].cpu(), {"net." + k: v.cpu().clone() for k, v in eng.param_views("grad").items()}, names


@pytest.mark.parametrize("mega_from", [1, 2])
@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_fused_fp32_plan_matches_reference(task, mega_from):
    """Exact-fp32 mode with the coarse scales fused (3xTF32 convolutions on the tensor cores inside the persistent kernel): the
    same bars against the reference fixture as the op-by-op fp32 plan (north_star's rtol 1e-3)."""
    from mfvi_dip_mia_b200 import _lib as L
    from mfvi_dip_mia_b200.engine import KL, NLL
    d, grads, out, a, ours, names = _step(task, mega_from, L.MATH_FP32)
    assert names.count("mfvi_mega_run") == 2 and names.count("mfvi_bn_param_grads") == 1, names
    S = int(d["S"])
    for s in range(S):
        assert rel_err(out[s:s + 1], d[f"out{s}"]) < 1e-4, s
    assert rel_err(a[NLL], d["nll"]) < 1e-5 and rel_err(a[KL], d["kl"]) < 1e-5
    errs = grad_errs({k: ours[k] for k in grads}, grads)
    worst = max(errs, key=errs.get)
    print(f"[mega fp32 {task} from scale {mega_from}] worst tensor {errs[worst]:.2e} ({worst.rsplit('.', 2)[-2]})")
    assert errs[worst] < 1e-3, (worst, errs[worst])


@pytest.mark.parametrize("task", ["den", "inp"])
def test_fused_tf32_plan_matches_unfused(task):
    """tf32 mode: fused coarse scales (mma.sync tf32, operands rounded to nearest) against the op-by-op plan (tcgen05 tf32,
    operands truncated): two tf32 evaluations of the same step, within the tf32 whole-step bars of tests/test_gpu_tc.py."""
    from mfvi_dip_mia_b200 import _lib as L
    from tests.test_gpu_tc import TF32_STEP_BARS
    d, grads, out0, a0, g0, _ = _step(task, -1, L.MATH_TF32)
    _, _, out1, a1, g1, names = _step(task, 1, L.MATH_TF32)
    assert names.count("mfvi_mega_run") == 2
    b_out, b_nll, b_l2 = TF32_STEP_BARS[0], TF32_STEP_BARS[1], TF32_STEP_BARS[2]
    assert rel_err(out1, out0) < b_out and rel_err(a1, a0) < b_nll
    va = torch.cat([g1[k].double().reshape(-1) for k in grads])
    vb = torch.cat([g0[k].double().reshape(-1) for k in grads])
    assert float((va - vb).norm() / vb.norm()) < b_l2
