"""Per-stage timeline of the persistent multi-stage kernel (csrc/mega.cu): CTA 0's %globaltimer at every stage start.
    MFVI_MEGA_PROFILE=1 MFVI_MEGA_FROM=2 python scripts/mega_profile.py [mc] [size]"""
import os
import sys

os.environ.setdefault("MFVI_MEGA_PROFILE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mfvi_dip_mia_b200 import _lib as L  # noqa: E402

mc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
args = type("A", (), dict(config="den", size=size, mc=mc))()
tr = bench.build_trainer(args, L.MATH_TF32, torch.device("cuda:0"), 0, 1)
for _ in range(6):
    tr.step()
torch.cuda.synchronize()
for m in tr.eng._mega:
    t = m["prof"].cpu().tolist()
    total = (t[-1] - t[0]) / 1e3
    print(f"--- program of {m['n_stages']} stages: {total:.1f} us")
    for i, name in enumerate(m["names"]):
        print(f"  {i:3d} {(t[i + 1] - t[i]) / 1e3:8.2f} us  {name}")
