import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfvi_dip_mia_b200 import SkipEngine, _lib as L
from mfvi_dip_mia_b200.trainer import LossHead
from tests.test_gpu_parity import SMALL, _fixture, _head_kwargs, spec_of
dev = torch.device("cuda:0")
task = sys.argv[1] if len(sys.argv) > 1 else "ct"
d, S, sd, eps, ex, grads = _fixture(task)
x = torch.from_numpy(d["net_input"])
eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev)
eng.load_params(sd, prefix="net."); eng.pack_eps(eps, prefix="net.")
head = LossHead(eng, task, **_head_kwargs(task, ex))
xin = x[0].permute(1, 2, 0).contiguous().to(dev)
# name the buffers by the op that writes them
writer = {}
for name, args, meta in eng.fwd_ops + eng.bwd_ops:
    for a in args:
        if isinstance(a, L.View) and a.ptr:
            writer.setdefault(a.ptr, []).append(name)
snaps = []
for rep in range(6):
    eng.zero_accumulators()
    eng.set_input(xin, None, 0.0, L.key(0))
    eng.sample_weights(L.key(0)); eng.forward(); head.run(); eng.backward()
    torch.cuda.synchronize()
    snap = [b.clone() for b in eng._bufs] + [eng.out.clone(), eng.dout.clone(), eng.arena.clone(), eng.dw.clone()]
    if task == "ct":
        snap += [head.sino.clone(), head.dsino.clone()]
    snaps.append(snap)
names = [f"buf{i}:{tuple(b.shape)}:{'/'.join(sorted(set(writer.get(b.data_ptr(), []))))}" for i, b in enumerate(eng._bufs)] + ["out", "dout", "arena", "dw", "sino", "dsino"]
for rep in range(1, 6):
    diffs = []
    for n, a, b in zip(names, snaps[0], snaps[rep]):
        if not torch.equal(a, b):
            e = float((a.double() - b.double()).abs().max() / (a.double().abs().max() + 1e-30))
            diffs.append((n, f"{e:.1e}"))
    print("rep", rep, "differs in", len(diffs), diffs[:12], flush=True)
