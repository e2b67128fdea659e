"""tcgen05 (kind::tf32, fp32 accumulate) convolution kernels against the exact-fp32 CUDA-core kernels of the same
library, on the layer shapes of the MFVI-DIP nets.  TF32 keeps a 10-bit mantissa for the operands, so the bar here is
the separately stated reduced-precision tolerance (north_star: "bf16 operands with fp32 accumulate stated
separately"): normalised max error < 3e-3 (observed ~5e-4); the fp32 path is held to 1e-3 in test_gpu_parity.py."""
import ctypes as C

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TF32_TOL = 3e-3

# (cin, cout, k, H, W[, stride]) — H, W = output size; input is ((H-1)*stride+k, (W-1)*stride+k)
SHAPES = [
    (132, 128, 3, 16, 16), (36, 16, 3, 64, 64), (16, 4, 1, 32, 32), (128, 128, 1, 16, 16), (16, 2, 1, 64, 64),
    (68, 32, 3, 32, 48), (16, 16, 5, 24, 24), (128, 128, 3, 8, 8), (64, 64, 3, 32, 32), (32, 32, 1, 128, 128),
    (36, 16, 3, 256, 256), (68, 32, 3, 128, 128), (16, 16, 1, 256, 256), (132, 64, 3, 64, 64), (128, 4, 1, 16, 16),
    (16, 4, 1, 256, 256), (132, 128, 3, 32, 32), (64, 64, 1, 64, 64), (20, 24, 3, 37, 29), (16, 16, 5, 130, 70),
    (16, 16, 3, 128, 128, 2), (64, 128, 3, 16, 16, 2), (128, 128, 3, 8, 8, 2), (16, 16, 5, 32, 32, 2), (32, 64, 3, 31, 33, 2),
    (16, 32, 1, 16, 16, 2),
]


def _run(shape, math, S=2, seed=0, broadcast_x=False):
    from mfvi_dip_mia_b200 import _lib as L
    cin, cout, k, H, W = shape[:5]
    stride = shape[5] if len(shape) > 5 else 1
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    Hin, Win = (H - 1) * stride + k, (W - 1) * stride + k
    if stride == 2:          # the nets' stride-2 convs read an even-sized padded image (one unused trailing row/col)
        Hin, Win = Hin + 1, Win + 1
    x = torch.randn(1 if broadcast_x else S, Hin, Win, cin, device=dev, generator=g)
    w = torch.randn(S, k * k * cout * cin + cout, device=dev, generator=g) * 0.1
    dy = torch.randn(S, H, W, cout, device=dev, generator=g)
    y = torch.zeros(S, H, W, cout, device=dev)
    dx = torch.zeros(S, Hin, Win, cin, device=dev)
    dw = torch.zeros_like(w)
    stats = torch.zeros(S, cout, 2, dtype=torch.float64, device=dev)
    d = L.ConvDesc(S, cin, cout, k, k, stride, Hin, Win, H, W, math)
    P = w.shape[1]
    boff = k * k * cout * cin
    L.call("mfvi_conv2d_fwd", C.byref(d), L.view(x), w.data_ptr(), w.data_ptr() + 4 * boff, P, L.view(y), stats.data_ptr())
    L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dy), w.data_ptr(), P, L.view(dx), 0)
    L.call("mfvi_conv2d_wgrad", C.byref(d), L.view(x), L.view(dy), dw.data_ptr(), dw.data_ptr() + 4 * boff, P)
    torch.cuda.synchronize()
    return y, dx, dw, stats


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_conv_matches_fp32_kernels(shape):
    from mfvi_dip_mia_b200 import _lib as L
    ref = _run(shape, L.MATH_FP32)
    got = _run(shape, L.MATH_TF32)
    names = ["y", "dx", "dw(+dbias)", "bn stats"]
    for n, a, b in zip(names, got, ref):
        assert torch.isfinite(a).all(), n
        assert rel_err(a, b) < TF32_TOL, (shape, n, rel_err(a, b))


@pytest.mark.parametrize("shape,S", [((16, 16, 3, 64, 64), 2), ((16, 16, 3, 128, 128, 2), 8), ((32, 16, 3, 40, 24, 2), 2),
                                     ((16, 4, 1, 64, 64), 8), ((16, 8, 3, 32, 32), 4), ((16, 16, 5, 32, 32, 2), 6)])
def test_tc_conv_broadcast_input(shape, S):
    """One input image shared by all samples (sample stride 0: the first layers of the net).  Covers the sample-folded
    weight-gradient mode (samples stacked along M when S*Cout <= 128) and shapes that do not qualify for it."""
    from mfvi_dip_mia_b200 import _lib as L
    ref = _run(shape, L.MATH_FP32, S=S, broadcast_x=True)
    got = _run(shape, L.MATH_TF32, S=S, broadcast_x=True)
    for n, (a, b) in zip(["y", "dx", "dw", "stats"], zip(got, ref)):
        if n == "dx":
            continue                      # the gradient of a broadcast input is not defined per sample (never requested)
        assert rel_err(a, b) < TF32_TOL, (shape, S, n, rel_err(a, b))


# whole-step bars of the tf32 mode on the four small task nets: ~2x the errors measured on a B200 (printed by the test,
# profiles/r02_parity_errors.txt: output 2.5e-3 .. 4.1e-3, nll 0.8e-4 .. 2.3e-4, gradient relL2 0.8e-2 .. 1.7e-2, cosine >= 0.99986,
# worst tensor 0.03 .. 0.24).  (output, nll, whole-gradient relative L2, cosine, median / p90 / max per-tensor normalised error)
TF32_STEP_BARS = (8e-3, 5e-4, 3.5e-2, 0.9995, 3e-2, 1.5e-1, 0.5)


@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_tf32_engine_step_close_to_reference(task):
    """Whole step in TF32 mode against the reference fixture, all four tasks.  This is the separately stated
    reduced-precision mode (10-bit operand mantissa, truncated by the tensor core; fp32 accumulate).  Individual small
    tensors (BatchNorm biases, 1x1 skip convs) carry the largest relative error through cancellation in the BN backward;
    the fp32 mode (test_gpu_parity.py) is the one held to rtol 1e-3."""
    from mfvi_dip_mia_b200 import SkipEngine, _lib as L
    from mfvi_dip_mia_b200.engine import KL, NLL
    from mfvi_dip_mia_b200.trainer import LossHead
    from oracle import mfvi_oracle as O
    from tests.helpers import grad_errs
    from tests.test_gpu_parity import SMALL, _fixture, _head_kwargs, spec_of
    dev = torch.device("cuda:0")
    d, S, sd, eps, ex, grads = _fixture(task)
    x = torch.from_numpy(d["net_input"])
    eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev, math=L.MATH_TF32)
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
    b_out, b_nll, b_l2, b_cos, b_med, b_p90, b_max = TF32_STEP_BARS
    e_out = max(rel_err(out[s:s + 1], d[f"out{s}"]) for s in range(S))
    a = eng.arena[:2].cpu()
    e_nll = rel_err(a[NLL], d["nll"])
    ours = {"net." + k: v.cpu() for k, v in eng.param_views("grad").items()}
    va = torch.cat([ours[k].double().reshape(-1) for k in grads])
    vb = torch.cat([grads[k].double().reshape(-1) for k in grads])
    e_l2 = float((va - vb).norm() / vb.norm())
    cos = float((va @ vb) / (va.norm() * vb.norm()))
    errs = grad_errs({k: ours[k] for k in grads}, grads)
    worst = max(errs, key=errs.get)
    import numpy as np
    ev = np.array(list(errs.values()))
    med, p90 = float(np.median(ev)), float(np.percentile(ev, 90))
    print(f"[parity small {task} tf32] out {e_out:.2e}  nll {e_nll:.2e}  grad relL2 {e_l2:.2e}  cos {cos:.6f}  "
          f"per-tensor median/p90/max {med:.2e}/{p90:.2e}/{errs[worst]:.2e} ({worst.rsplit('.', 2)[-2]}.{worst.rsplit('.', 1)[-1]})")
    assert e_out < b_out and e_nll < b_nll, (e_out, e_nll)
    assert e_l2 < b_l2 and cos > b_cos, (e_l2, cos)
    assert med < b_med and p90 < b_p90 and errs[worst] < b_max, (med, p90, worst, errs[worst])
