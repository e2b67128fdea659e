import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfvi_dip_mia_b200 import SkipEngine, _lib as L
from mfvi_dip_mia_b200.trainer import LossHead
from oracle import mfvi_oracle as O
from tests.test_gpu_parity import SMALL, _fixture, _head_kwargs, spec_of
dev = torch.device("cuda:0")
for task in ["ct", "den", "inp"]:
    d, S, sd, eps, ex, grads = _fixture(task)
    x = torch.from_numpy(d["net_input"])
    eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev)
    eng.load_params(sd, prefix="net."); eng.pack_eps(eps, prefix="net.")
    head = LossHead(eng, task, **_head_kwargs(task, ex))
    for b in eng._bufs:
        b.fill_(float("nan"))
    eng.dout.fill_(float("nan"))
    if task == "ct":
        head.sino.fill_(float("nan")); head.dsino.fill_(float("nan"))
    temp, sigma = float(d["temp"]), float(d["sigma"])
    eng.zero_accumulators()
    eng.set_input(x[0].permute(1, 2, 0).contiguous().to(dev), None, 0.0, L.key(0))
    eng.sample_weights(L.key(0)); eng.forward(); head.run(); eng.backward()
    eng.reparam_kl(L.key(0), prior_mu=0.0, prior_sigma_plus_eps=O.prior_scale(temp, sigma), direction=0, kscale=temp)
    torch.cuda.synchronize()
    print(task, "nan in grad:", bool(torch.isnan(eng.grad).any()), "nan in arena:", bool(torch.isnan(eng.arena).any()))
    if torch.isnan(eng.arena).any():
        for b in eng.lay.bns:
            n = S * b.C * 2
            print("   ", b.key.split(".")[-1], "sums nan", bool(torch.isnan(eng.arena[b.sums_off:b.sums_off+n]).any()), "red nan", bool(torch.isnan(eng.arena[b.red_off:b.red_off+n]).any()))
