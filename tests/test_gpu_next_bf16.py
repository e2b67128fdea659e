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
