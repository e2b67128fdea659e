"""Cost of one link of a dependent kernel chain inside a CUDA graph (programmatic dependent launch on), per op type, on the
smallest layer shapes of the metric net: captures N back-to-back calls of one op and replays them.
    python scripts/link_cost.py [S]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mfvi_dip_mia_b200 import _lib as L  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
N = 100


def chain(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:58s} {e0.elapsed_time(e1) * 1e3 / (10 * N):7.2f} us per link")


ctr = torch.zeros(1, dtype=torch.int32, device=dev)
chain("counter_add (1 thread: the empty-kernel floor)", lambda: L.call("mfvi_counter_add", ctr.data_ptr(), 1))
buf = torch.zeros(1 << 14, device=dev)
chain("fill 64 KB", lambda: L.call("mfvi_fill_f32", buf.data_ptr(), buf.numel(), 0.0))
for (Cn, H) in ((128, 8), (128, 16), (64, 32), (32, 64), (16, 128), (16, 256)):
    y = torch.randn(S, H, H, Cn, device=dev)
    xp = torch.zeros(S, H + 2, H + 2, Cn, device=dev)
    sums = torch.stack([y.double().sum((1, 2)), (y.double() ** 2).sum((1, 2))], -1).contiguous()
    gam, bet = torch.ones(Cn, device=dev), torch.zeros(Cn, device=dev)
    chain(f"bn_act_pad_fwd C={Cn} {H}x{H} S={S}", lambda: L.call("mfvi_bn_act_pad_fwd", L.view(y), S, H, H, Cn, sums.data_ptr(),
                                                                  gam.data_ptr(), bet.data_ptr(), 1, 1, L.view(xp)))
    g = torch.zeros(S, H, H, Cn, device=dev)
    red = torch.zeros(S, Cn, 2, dtype=torch.float64, device=dev)
    dg, db = torch.zeros(Cn, device=dev), torch.zeros(Cn, device=dev)
    chain(f"pad_act_bwd    C={Cn} {H}x{H} S={S}", lambda: L.call("mfvi_pad_act_bwd", L.view(xp), S, H, H, Cn, 1, L.view(y), sums.data_ptr(),
                                                                  gam.data_ptr(), bet.data_ptr(), 1, L.view(g), red.data_ptr()))
    chain(f"bn_bwd_apply   C={Cn} {H}x{H} S={S}", lambda: L.call("mfvi_bn_bwd_apply", L.view(g), L.view(y), S, H, H, Cn, sums.data_ptr(),
                                                                  red.data_ptr(), gam.data_ptr(), L.view(g), dg.data_ptr(), db.data_ptr()))
    k = 3
    cout = Cn
    P = k * k * cout * Cn + cout
    Pp = (P + 3) // 4 * 4
    w = torch.randn(S, Pp, device=dev) * 0.05
    yo = torch.zeros(S, H, H, cout, device=dev)
    st = torch.zeros(S, cout, 2, dtype=torch.float64, device=dev)
    d = L.ConvDesc(S, Cn, cout, k, k, 1, H + 2, H + 2, H, H, L.MATH_TF32)
    chain(f"conv fwd  tf32 {Cn}->{cout} k3 {H}x{H} S={S}", lambda: L.call("mfvi_conv2d_fwd", C.byref(d), L.view(xp), w.data_ptr(),
                                                                          w.data_ptr() + 4 * (P - cout), Pp, L.view(yo), st.data_ptr()))
    dxp = torch.zeros_like(xp)
    chain(f"conv dgrad tf32 {Cn}->{cout} k3 {H}x{H} S={S}", lambda: L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(yo), w.data_ptr(), Pp,
                                                                           L.view(dxp), 0))

# the concat / x2-upsample pair of every scale of the metric net (skip branch 4 channels, bilinear)
for (Cd, H) in ((128, 16), (128, 32), (128, 64), (64, 128), (32, 256)):
    Cs, h2 = 4, H // 2
    ys = torch.randn(S, H, H, Cs, device=dev)
    yd = torch.randn(S, h2, h2, Cd, device=dev)
    mk = lambda t: torch.stack([t.double().sum((1, 2)), (t.double() ** 2).sum((1, 2))], -1).contiguous()
    ss, sd = mk(ys), mk(yd)
    gs_, bs_ = torch.ones(Cs, device=dev), torch.zeros(Cs, device=dev)
    gd_, bd_ = torch.ones(Cd, device=dev), torch.zeros(Cd, device=dev)
    A = torch.zeros(S, H + 2, H + 2, Cs + Cd, device=dev)
    Ai = A[:, 1:-1, 1:-1, :]
    sA = torch.zeros(S, Cs + Cd, 2, dtype=torch.float64, device=dev)
    chain(f"cat_up_fwd  Cs=4 Cd={Cd} {H}x{H} S={S}", lambda: L.call("mfvi_cat_up_fwd", L.view(ys), Cs, ss.data_ptr(), gs_.data_ptr(),
                                                                     bs_.data_ptr(), L.view(yd), Cd, sd.data_ptr(), gd_.data_ptr(),
                                                                     bd_.data_ptr(), S, H, H, 0, L.view(Ai), sA.data_ptr()))
    dA = torch.randn(S, H, H, Cs + Cd, device=dev)
    g_s, g_d = torch.zeros_like(ys), torch.zeros_like(yd)
    rs = torch.zeros(S, Cs, 2, dtype=torch.float64, device=dev)
    rd = torch.zeros(S, Cd, 2, dtype=torch.float64, device=dev)
    for part, nm in ((1, "skip"), (2, "up  ")):
        chain(f"cat_up_bwd {nm} Cs=4 Cd={Cd} {H}x{H} S={S}", lambda: L.call(
            "mfvi_cat_up_bwd", L.view(dA), S, H, H, 0, L.view(ys), Cs, ss.data_ptr(), gs_.data_ptr(), bs_.data_ptr(), L.view(g_s),
            rs.data_ptr(), L.view(yd), Cd, sd.data_ptr(), gd_.data_ptr(), bd_.data_ptr(), L.view(g_d), rd.data_ptr(), part))
