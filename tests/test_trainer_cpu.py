"""MfviDipTrainer's host logic on the CPU: the trainer is built plan-only, and tests/plan_interpreter.py executes every
libmfvidip call it makes.  What is under test is everything around the kernels — Philox keys and the device-side step counter,
the input jitter stream, KL / reparameterisation scaling, AdamW bias correction, MC-sample sharding and the per-step gradient
all-reduce — against an independent loop written with the oracle and torch.optim.AdamW.  The kernels themselves are
tests/test_gpu_*.py's business."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import mfvi_oracle as O
from oracle import philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEMP, SIGMA, LR, SEED = 5.656911698337764e-07, 1.4616642493692077e-05, 1e-2, 7
CFG = O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear")


def _problem():
    g = torch.Generator().manual_seed(0)
    return torch.rand(1, 4, 32, 32, generator=g) * 0.1, torch.rand(1, 1, 32, 32, generator=g)


def _trainer(S, rank=0, world=1):
    from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec
    x, tgt = _problem()
    spec = SkipSpec(CFG.num_input_channels, CFG.num_output_channels, tuple(CFG.down), tuple(CFG.up), tuple(CFG.skip),
                    CFG.filter_down, CFG.filter_up, CFG.filter_skip, CFG.need1x1_up, CFG.need_sigmoid, CFG.upsample_mode)
    return MfviDipTrainer(spec, "den", x, temp=TEMP, sigma=SIGMA, lr=LR, mc_samples=S, seed=SEED, device="cpu", target=tgt,
                          rank=rank, world_size=world, plan_only=True)


def _oracle_loop(tr, S, n_steps):
    """The same optimisation written independently: oracle forward / autograd, torch's AdamW, eps and jitter drawn from the
    Philox streams by (seed, stream, GLOBAL sample id, step)."""
    x, tgt = _problem()
    lay = tr.eng.lay
    sd = {"net." + k: v.detach().clone().requires_grad_("running" not in k) for k, v in tr.eng.param_views("theta").items()}
    opt = torch.optim.AdamW([v for v in sd.values() if v.requires_grad], lr=LR, weight_decay=0)
    for step in range(n_steps):
        opt.zero_grad()
        z = torch.from_numpy(philox.philox_normal(x.numel(), SEED, 1, 0, step)).reshape(x.shape)
        eps = []
        for s in range(S):
            flat = torch.from_numpy(philox.philox_normal(lay.P, SEED, 0, s, step))
            eps.append({**{"net." + c.key + ".W": flat[c.w_off:c.w_off + c.w_numel].view(c.k, c.k, c.cout, c.cin).permute(2, 3, 0, 1)
                           for c in lay.convs},
                        **{"net." + c.key + ".b": flat[c.b_off:c.b_off + c.cout] for c in lay.convs}})
        loss, _, _, _ = O.mfvi_loss(sd, CFG, x + 0.1 * z, eps, task="den", temp=TEMP,
                                    prior_sigma_plus_eps=O.prior_scale(TEMP, SIGMA), target=tgt)
        loss.backward()
        opt.step()
    return sd, float(loss.detach())


def test_interpreted_trainer_follows_the_oracle_optimisation():
    from tests.plan_interpreter import TrainerInterpreter
    S, n_steps = 2, 3
    tr = _trainer(S)
    theta0 = tr.eng.theta.clone()
    sd, loss_ref = _oracle_loop(tr, S, n_steps)                      # reads the initial parameters: before the trainer moves them
    with TrainerInterpreter(tr):
        for _ in range(n_steps):
            tr.step()
        loss = tr.loss_terms()[2]
    assert tr.steps_done == n_steps
    assert abs(loss - loss_ref) < 1e-3 * abs(loss_ref)                # the third step's loss, after two Adam updates
    assert (tr.eng.theta - theta0).abs().max() > 0.5 * LR             # Adam moves every parameter by ~lr per step
    ours = tr.eng.param_views("theta")
    sq = n = 0.0
    for k, v in sd.items():
        if v.requires_grad:
            sq += float((ours[k[len("net."):]] - v.detach()).double().pow(2).sum())
            n += v.numel()
    # parameters agree to a small fraction of the distance they travelled (Adam's normalisation turns the fp32 noise of the
    # smallest gradients into update noise, so the yardstick is the step length, not the parameter value)
    rms = (sq / n) ** 0.5 / (LR * n_steps)
    assert rms < 2e-2, rms


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from tests.plan_interpreter import TrainerInterpreter
from tests.test_trainer_cpu import _trainer, LR
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
S, n_steps = 4, 2
single = _trainer(S)                                   # every rank also runs the unsharded trainer for comparison
with TrainerInterpreter(single):
    for _ in range(n_steps):
        single.step()
sharded = _trainer(S, rank, world)                     # rank r owns global samples [r*S/world, (r+1)*S/world)
assert sharded.S == S // world and sharded.sample0 == rank * (S // world)
with TrainerInterpreter(sharded):
    for _ in range(n_steps):
        sharded.step()
err = float((sharded.eng.theta - single.eng.theta).abs().max()) / (LR * n_steps)
mine = sharded.eng.theta.clone()
other = mine.clone()
dist.broadcast(other, src=0)
print(f"rank {rank} vs single-process {err:.3e}  replicas differ by {float((mine - other).abs().max()):.1e}")
assert err < 2e-2, err                                 # fraction of the distance travelled (fp32 summation order only)
assert torch.equal(mine, other)                        # identical AdamW on identical averaged gradients: no drift between ranks
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharded_trainer_equals_single_process(tmp_path):
    """world_size 2 on CPU (gloo): the real MfviDipTrainer per rank (MC samples sharded, eps keyed by global sample id, one
    all-reduce of the flat gradient per step, identical AdamW on every rank) ends where the single-process trainer ends, and
    the two replicas stay bit-identical."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 31500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=dict(os.environ, OMP_NUM_THREADS="2"), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 vs single-process" in r.stdout and "rank 1 vs single-process" in r.stdout
