"""Per-launch device time of one eager MFVI-DIP step (CUDA events around every C-ABI call).
usage: python scripts/layer_table.py [fp32|tf32] [mc] [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mfvi_dip_mia_b200 import _lib as L

math = sys.argv[1] if len(sys.argv) > 1 else "tf32"
mc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
size = int(sys.argv[3]) if len(sys.argv) > 3 else 256
args = type("A", (), dict(config="den", size=size, mc=mc))()
tr = bench.build_trainer(args, L.MATH_TF32 if math == "tf32" else L.MATH_FP32, torch.device("cuda:0"), 0, 1)
tr.use_graph = False
for _ in range(3):
    tr.step()
torch.cuda.synchronize()
acc = {}
N = 5
for _ in range(N):
    L.timeline = []
    tr._step_eager()
    torch.cuda.synchronize()
    tl, L.timeline = L.timeline, None
    for i, (name, e0, e1, meta) in enumerate(tl):
        a = acc.setdefault(i, [name, meta, 0.0])
        a[2] += e0.elapsed_time(e1) / N
tot = sum(a[2] for a in acc.values())
print(f"total {tot:.3f} ms over {len(acc)} launches ({math}, MC={mc}, {size}^2)")
for i, (name, meta, ms) in acc.items():
    meta = meta or {}
    fl = meta.get("flops", 0); by = meta.get("bytes", 0)
    extra = f"{meta.get('layer',''):22s} {meta.get('shape',''):30s}"
    print(f"{i:4d} {name:26s} {extra} {ms*1e3:8.1f} us  {fl/ms/1e9 if fl else 0:7.1f} TF/s  {by/ms/1e6 if by else 0:7.0f} GB/s")
