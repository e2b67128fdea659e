import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfvi_dip_mia_b200 import SkipEngine, _lib as L
from mfvi_dip_mia_b200.engine import KL, NLL
from mfvi_dip_mia_b200.trainer import LossHead
from oracle import mfvi_oracle as O
from tests.helpers import grad_errs, rel_err
from tests.test_gpu_parity import SMALL, _fixture, _head_kwargs, spec_of
dev = torch.device("cuda:0")
math = L.MATH_TF32 if (len(sys.argv) > 1 and sys.argv[1] == "tf32") else L.MATH_FP32
for task in ["den", "sr", "ct", "inp"]:
    d, S, sd, eps, ex, grads = _fixture(task)
    x = torch.from_numpy(d["net_input"])
    for rep in range(3):
        eng = SkipEngine(spec_of(SMALL[task]), x.shape[2], x.shape[3], S, dev, math=math)
        eng.load_params(sd, prefix="net."); eng.pack_eps(eps, prefix="net.")
        head = LossHead(eng, task, **_head_kwargs(task, ex))
        temp, sigma = float(d["temp"]), float(d["sigma"])
        eng.zero_accumulators()
        eng.set_input(x[0].permute(1, 2, 0).contiguous().to(dev), None, 0.0, L.key(0))
        eng.sample_weights(L.key(0)); eng.forward(); head.run(); eng.backward()
        eng.reparam_kl(L.key(0), prior_mu=0.0, prior_sigma_plus_eps=O.prior_scale(temp, sigma), direction=0, kscale=temp)
        out = eng.out_nchw().cpu()
        oe = max(rel_err(out[s:s+1], d[f"out{s}"]) for s in range(S))
        ours = {"net." + k: v.cpu() for k, v in eng.param_views("grad").items()}
        errs = grad_errs({k: ours[k] for k in grads}, grads)
        va = torch.cat([ours[k].double().reshape(-1) for k in grads]); vb = torch.cat([grads[k].double().reshape(-1) for k in grads])
        print("   global rel L2 err %.3e  cosine %.6f" % (float((va - vb).norm() / vb.norm()), float((va @ vb) / (va.norm() * vb.norm()))))
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(task, rep, f"out err {oe:.2e}", [(k.split('.')[-2][-12:] + '.' + k.split('.')[-1], f"{v:.1e}") for k, v in top], flush=True)

