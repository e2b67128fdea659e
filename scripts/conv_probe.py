"""Time (or profile under ncu) single conv launches of the library on chosen layer shapes.
usage: python scripts/conv_probe.py [fwd|dgrad|wgrad|all] cin cout k H W [stride] [S] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfvi_dip_mia_b200 import _lib as L

which = sys.argv[1]
cin, cout, k, H, W = (int(a) for a in sys.argv[2:7])
stride = int(sys.argv[7]) if len(sys.argv) > 7 else 1
S = int(sys.argv[8]) if len(sys.argv) > 8 else 8
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 20
dev = torch.device("cuda:0")
Hin, Win = (H - 1) * stride + k, (W - 1) * stride + k
if stride == 2:
    Hin, Win = Hin + 1, Win + 1
x = torch.randn(S, Hin, Win, cin, device=dev)
w = torch.randn(S, k * k * cout * cin + cout, device=dev) * 0.1
dy = torch.randn(S, H, W, cout, device=dev)
y = torch.zeros(S, H, W, cout, device=dev)
dx = torch.zeros(S, Hin, Win, cin, device=dev)
dw = torch.zeros_like(w)
stats = torch.zeros(S, cout, 2, dtype=torch.float64, device=dev)
d = L.ConvDesc(S, cin, cout, k, k, stride, Hin, Win, H, W, L.MATH_TF32)
P = w.shape[1]
boff = k * k * cout * cin
flush = torch.empty(256 * 2 ** 20 // 4, device=dev)

ops = {
    "fwd": lambda: L.call("mfvi_conv2d_fwd", C.byref(d), L.view(x), w.data_ptr(), w.data_ptr() + 4 * boff, P, L.view(y), stats.data_ptr()),
    "dgrad": lambda: L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dy), w.data_ptr(), P, L.view(dx), 0),
    "wgrad": lambda: L.call("mfvi_conv2d_wgrad", C.byref(d), L.view(x), L.view(dy), dw.data_ptr(), None if os.environ.get("PROBE_NOBIAS") else dw.data_ptr() + 4 * boff, P),
}
flops = 2.0 * S * H * W * cout * cin * k * k
for name, fn in ops.items():
    if which not in (name, "all"):
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    byt = 4.0 * (x.numel() + y.numel() + w.numel())
    print(f"{name:6s} {cin}->{cout} k{k} s{stride} {H}x{W} S={S}: median {med*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us  "
          f"{flops/med/1e9:7.1f} TF/s  {byt/med/1e6:7.0f} GB/s(alg)")

if os.environ.get("TC2_TIMELINE"):
    dbg = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
    os.environ["MFVI_TC2_DBG"] = hex(dbg.data_ptr())
    ops[which if which != "all" else "fwd"]()
    torch.cuda.synchronize()
    t = dbg.cpu().view(-1, 8)
    t0 = int(t[0, 0])
    names = ["A issue", "mma acc_empty ok", "mma A full", "mma committed", "epi wait", "epi acc_full", "epi done", "epi flushed"]
    print("timeline of CTA 0 (cycles since first A issue):", names)
    for i in range(t.shape[0]):
        if int(t[i, 0]) == 0:
            break
        print(i, [int(v) - t0 if int(v) else None for v in t[i]])
    print("mma warp totals: wait_a %d  wait_b %d  issue %d cycles for %d MMAs" % tuple(int(v) for v in t[63, :4]))
