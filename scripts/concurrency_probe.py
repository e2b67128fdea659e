"""How much does running the MC samples as G independent groups on G streams help?  G trainers with S/G samples each,
their step graphs launched on G streams concurrently, against one trainer with S samples.
usage: python scripts/concurrency_probe.py [S] [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_problem, TEMP, SIGMA, LR
from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec, _lib as L

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
x, t = synthetic_problem(size)
dev = torch.device("cuda:0")


def make(s):
    return MfviDipTrainer(SkipSpec(), "den", x, temp=TEMP, sigma=SIGMA, lr=LR, mc_samples=s, seed=1, device=dev, target=t,
                          math_mode=L.MATH_TF32, use_graph=True)


def time_group(trs, steps=40):
    streams = [torch.cuda.Stream() for _ in trs]
    for tr, st in zip(trs, streams):
        with torch.cuda.stream(st):
            for _ in range(4):
                tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    for _ in range(steps):
        for tr, st in zip(trs, streams):
            with torch.cuda.stream(st):
                tr.step()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for G in (1, 2, 4):
    trs = [make(S // G) for _ in range(G)]
    ms = time_group(trs)
    print(f"S={S} as {G} group(s) of {S // G} on {G} stream(s): {ms:.3f} ms per full step  ({1e3 / ms:.1f} steps/s)")
    del trs
