"""In-graph cost of every op of the real step plan: each C-ABI op of the engine's forward / backward plan is captured N times
back to back (same arguments, programmatic dependent launch on) into a CUDA graph and replayed; time / N = what one link of
that op costs inside the step's graph with warm L2 (the upper bound of its share of the critical path when it sits on the
main lane).  Lanes are printed so side-lane ops (weight gradients, skip branch) can be told apart.
    python scripts/plan_link_cost.py [tf32|fp32] [mc] [config] [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mfvi_dip_mia_b200 import _lib as L  # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else "tf32"
mc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
config = sys.argv[3] if len(sys.argv) > 3 else "den"
size = int(sys.argv[4]) if len(sys.argv) > 4 else (256 if config == "den" else 512)
args = type("A", (), dict(config=config, size=size, mc=mc))()
tr = bench.build_trainer(args, L.MATH_TF32 if math == "tf32" else L.MATH_FP32, torch.device("cuda:0"), 0, 1)
for _ in range(3):
    tr.step()
torch.cuda.synchronize()
eng = tr.eng
N = 50


def chain(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * N)


tot = {}
rows = []
for phase, ops in (("fwd", eng.fwd_ops), ("bwd", eng.bwd_ops)):
    for name, a, meta in ops:
        if name.startswith("__"):
            continue
        lane = "wgrad" if name == "mfvi_conv2d_wgrad" else meta.get("lane", "main")
        us = chain(lambda: L.call(name, *a))
        rows.append((phase, lane, name, meta.get("layer", ""), meta.get("shape", ""), us, meta.get("flops", 0), meta.get("bytes", 0)))
        k = (lane, name)
        tot[k] = tot.get(k, 0.0) + us
print(f"# {config} {size}^2 MC={mc} {math}: in-graph link cost of every plan op (us), {len(rows)} ops")
for phase, lane, name, layer, shape, us, fl, by in rows:
    print(f"{phase} {lane:5s} {name:24s} {layer:22s} {shape:28s} {us:8.2f} us  {fl / us / 1e6 if fl else 0:7.1f} TF/s  {by / us / 1e3 if by else 0:7.0f} GB/s")
print("# totals per (lane, op)")
for (lane, name), us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{lane:5s} {name:24s} {us:9.1f} us")
print(f"# main lane total {sum(v for (l, _), v in tot.items() if l == 'main'):.1f} us; wgrad lane {sum(v for (l, _), v in tot.items() if l == 'wgrad'):.1f} us; "
      f"skip lane {sum(v for (l, _), v in tot.items() if l == 'skip'):.1f} us")
