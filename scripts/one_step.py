"""Two eager steps of a BASELINE configuration (default: the metric shape 256^2, MC=8, tf32) — a short driver for
`ncu --set full -k regex:<kernel>`.   usage: python scripts/one_step.py [den|sr|inp|ct] [tf32|fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mfvi_dip_mia_b200 import _lib as L

config = sys.argv[1] if len(sys.argv) > 1 else "den"
math = sys.argv[2] if len(sys.argv) > 2 else "tf32"
args = type("A", (), dict(config=config, size=bench.CONFIGS[config]["size"], mc=bench.CONFIGS[config]["mc"]))()
tr = bench.build_trainer(args, L.MATH_TF32 if math == "tf32" else L.MATH_FP32, torch.device("cuda:0"), 0, 1)
tr.use_graph = False
for _ in range(2):
    tr.step()
torch.cuda.synchronize()
print("ok", tr.loss_terms())
