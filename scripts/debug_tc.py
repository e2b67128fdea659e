"""Debug helper: compares tcgen05 conv kernels with the fp32 CUDA-core kernels and prints error summaries."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_tc import _run, SHAPES
from mfvi_dip_mia_b200 import _lib as L

def summarize(n, a, b):
    a, b = a.double(), b.double()
    den = float(b.abs().max())
    err = float((a - b).abs().max()) / (den or 1)
    nz = float((a == 0).double().mean())
    ratio = float((a * b).sum() / ((b * b).sum() + 1e-30))
    print(f"   {n:12s} relerr {err:9.3e}  zeros {nz:5.2f}  proj-ratio {ratio:8.4f}  |ref|max {den:9.3e}", flush=True)

shapes = SHAPES if len(sys.argv) < 2 else [SHAPES[int(i)] for i in sys.argv[1:]]
for sh in shapes:
    print(sh, flush=True)
    ref = _run(sh, L.MATH_FP32)
    got = _run(sh, L.MATH_TF32)
    for n, a, b in zip(["y", "dx", "dw", "stats"], got, ref):
        summarize(n, a, b)
    # per-tap error of dw
    cin, cout, k, H, W = sh[:5]
    dwa = got[2][:, :k*k*cout*cin].reshape(-1, k*k, cout, cin); dwb = ref[2][:, :k*k*cout*cin].reshape(-1, k*k, cout, cin)
    print("   dw per tap:", [f"{float((dwa[:,t]-dwb[:,t]).abs().max()/dwb.abs().max()):.1e}" for t in range(k*k)])
    dxa, dxb = got[1], ref[1]
    print("   dx per cin block:", [f"{float((dxa[...,c:c+32]-dxb[...,c:c+32]).abs().max()/dxb.abs().max()):.1e}" for c in range(0, cin, 32)])
