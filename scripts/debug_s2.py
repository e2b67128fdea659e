import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfvi_dip_mia_b200 import _lib as L
dev = torch.device("cuda:0")
cin, cout, k, H, W, stride = 16, 16, 3, 16, 16, 2
S = 1
Hin, Win = (H - 1) * stride + k + 1, (W - 1) * stride + k + 1
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(S, Hin, Win, cin, device=dev, generator=g)
dy = torch.randn(S, H, W, cout, device=dev, generator=g)
P = k * k * cout * cin + cout
def run(math, w):
    y = torch.zeros(S, H, W, cout, device=dev); dx = torch.zeros(S, Hin, Win, cin, device=dev)
    d = L.ConvDesc(S, cin, cout, k, k, stride, Hin, Win, H, W, math)
    L.call("mfvi_conv2d_fwd", C.byref(d), L.view(x), w.data_ptr(), None, P, L.view(y), None)
    L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dy), w.data_ptr(), P, L.view(dx), 0)
    torch.cuda.synchronize()
    return y, dx
for t in range(k * k):
    w = torch.zeros(S, P, device=dev)
    w[:, t * cout * cin:(t + 1) * cout * cin] = torch.randn(cout * cin, device=dev, generator=g)
    y0, dx0 = run(L.MATH_FP32, w); y1, dx1 = run(L.MATH_TF32, w)
    ey = float((y1 - y0).abs().max() / y0.abs().max()); ed = float((dx1 - dx0).abs().max() / dx0.abs().max())
    # which rows/cols of y are wrong
    bad = ((y1 - y0).abs().amax(dim=(0, 3)) > 1e-2 * y0.abs().max())
    print(f"tap {t} (r={t//k}, s={t%k}): y err {ey:.2e} bad rows {bad.any(1).nonzero().flatten().tolist()[:6]} bad cols {bad.any(0).nonzero().flatten().tolist()[:6]}  dx err {ed:.2e}")
    if ed > 1e-2:
        badx = ((dx1 - dx0).abs().amax(dim=(0, 3)) > 1e-2 * dx0.abs().max())
        print("    dx bad rows", badx.any(1).nonzero().flatten().tolist()[:10], "bad cols", badx.any(0).nonzero().flatten().tolist()[:10])
