#!/usr/bin/env python
"""Print the convolution dispatch table of a task network: which kernel family libmfvidip runs for every forward / data-
gradient / weight-gradient launch and with which tile plan.  Host-only (mfvi_conv2d_plan on a plan-only engine): runs without
a GPU.   python scripts/dispatch_table.py [den|sr|ct|inp] [S] [fp32|tf32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L  # noqa: E402

NETS = {"den": (SkipSpec(), 256), "sr": (SkipSpec(32, 2), 512), "ct": (SkipSpec(16, 1), 512),
        "inp": (SkipSpec(16, 4, (16, 32, 64, 128, 128, 128), (16, 32, 64, 128, 128, 128), (0,) * 6, 5, 3, 1, False, False,
                         "nearest"), 512)}


def main():
    task = sys.argv[1] if len(sys.argv) > 1 else "den"
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    math = L.MATH_FP32 if (len(sys.argv) > 3 and sys.argv[3] == "fp32") else L.MATH_TF32
    spec, H = NETS[task]
    rows = SkipEngine(spec, H, H, S, "meta", math=math).conv_dispatch_table()
    print(f"# {task} net {H}x{H}, S={S}, {'tf32' if math == L.MATH_TF32 else 'fp32'}: {len(rows)} convolution launches")
    for r in rows:
        plan = " ".join(f"{k}={v}" for k, v in r["plan"].items())
        print(f"{r['op']:5s} {r['layer']:17s} {r['shape']:31s} {r['family']:9s} grid={r['grid'][0]}x{r['grid'][1]}x{r['grid'][2]} "
              f"block={r['block']} smem={r['smem_bytes']} {plan}")


if __name__ == "__main__":
    main()
