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


TASK_CFG = {
    "den": CFG,
    "sr": CFG,
    "ct": O.SkipCfg(4, 1, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear"),
    "inp": O.SkipCfg(4, 4, (8, 16, 16), (8, 16, 16), (0, 0, 0), 5, 3, 1, False, False, "nearest"),
}


def _problem(task="den"):
    """(net input, head kwargs as MfviDipTrainer / oracle.mfvi_loss take them)"""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 4, 32, 32, generator=g) * 0.1
    if task == "den":
        return x, dict(target=torch.rand(1, 1, 32, 32, generator=g))
    if task == "sr":
        return x, dict(target=torch.rand(1, 1, 8, 8, generator=g))
    if task == "inp":
        return x, dict(target=torch.rand(1, 3, 32, 32, generator=g), mask=(torch.rand(1, 1, 32, 32, generator=g) > 0.5).float())
    theta = torch.arange(0., 180., 20.)
    return x, dict(theta_deg=theta, sino=O.radon_forward(torch.rand(1, 1, 32, 32, generator=g), theta))


def _trainer(S, rank=0, world=1, task="den"):
    from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec
    x, head = _problem(task)
    cfg = TASK_CFG[task]
    spec = SkipSpec(cfg.num_input_channels, cfg.num_output_channels, tuple(cfg.down), tuple(cfg.up), tuple(cfg.skip),
                    cfg.filter_down, cfg.filter_up, cfg.filter_skip, cfg.need1x1_up, cfg.need_sigmoid, cfg.upsample_mode)
    return MfviDipTrainer(spec, task, x, temp=TEMP, sigma=SIGMA, lr=LR, mc_samples=S, seed=SEED, device="cpu",
                          rank=rank, world_size=world, plan_only=True, **head)


def _oracle_loop(tr, S, n_steps, task="den"):
    """The same optimisation written independently: oracle forward / autograd, torch's AdamW, eps and jitter drawn from the
    Philox streams by (seed, stream, GLOBAL sample id, step)."""
    x, head = _problem(task)
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
        loss, _, _, _ = O.mfvi_loss(sd, TASK_CFG[task], x + 0.1 * z, eps, task=task, temp=TEMP,
                                    prior_sigma_plus_eps=O.prior_scale(TEMP, SIGMA), **head)
        loss.backward()
        if step == 0:
            first_grads = {k: v.grad.clone() for k, v in sd.items() if v.grad is not None}
        opt.step()
    return sd, float(loss.detach()), first_grads


@pytest.mark.parametrize("task", ["den", "sr", "inp", "ct"])
def test_interpreted_trainer_follows_the_oracle_optimisation(task):
    from tests.plan_interpreter import TrainerInterpreter
    S, n_steps = 2, 3
    tr = _trainer(S, task=task)
    theta0 = tr.eng.theta.clone()
    sd, loss_ref, first_grads = _oracle_loop(tr, S, n_steps, task)                      # reads the initial parameters: before the trainer moves them
    from tests.helpers import grad_errs
    with TrainerInterpreter(tr):
        for i in range(n_steps):
            tr.step()
            if i == 0:          # the gradient of the first step (data term over the samples + T * KL), before dynamics amplify anything
                errs = grad_errs({"net." + k: v.clone() for k, v in tr.eng.param_views("grad").items()}, first_grads)
        loss = tr.loss_terms()[2]
    assert max(errs.values()) < 1e-4, max(errs, key=errs.get)
    assert tr.steps_done == n_steps
    assert abs(loss - loss_ref) < 2e-3 * abs(loss_ref)                # the third step's loss, after two Adam updates
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
    assert rms < 5e-2, rms


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


def test_device_bookkeeping_matches_the_reference_loop():
    """runners.DeviceBookkeeping (reference bayesian_optimization.py:1374-1416) attached to an interpreted trainer: the
    bookkeeping launch is a post-step hook that must see the output of the step just computed and the iteration index of the
    device counter BEFORE it advances.  After 30 iterations (more than one turn of the 25-deep rings) the EMA output, the PSNR
    values, SSIM, the uncertainty maps and the UCE equal the reference's per-iteration formulas applied to the recorded outputs."""
    import math
    from mfvi_dip_mia_b200.runners import DeviceBookkeeping
    from mfvi_dip_mia_b200.utils.uce import uceloss
    from tests.plan_interpreter import TrainerInterpreter
    S, n_it, ring, expw = 2, 30, 25, 0.99
    tr = _trainer(S)
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(32, 32, generator=g)
    noisy = (gt + 0.1 * torch.randn(32, 32, generator=g)).clamp(0, 1)
    bk = DeviceBookkeeping(tr, gt=gt.numpy(), noisy=noisy.numpy(), exp_weight=expw, ring=ring)
    outs = []
    with TrainerInterpreter(tr):
        for _ in range(n_it):
            tr.step()
            outs.append(tr.eng.out.clone())                     # (S,H,W,2) of the step just taken
        m = bk.metrics()
        epi, ale, err2 = bk.uncertainty()
        uce = bk.uce()
    # the reference loop on the recorded outputs
    out_avg, r_epi, r_ale = None, torch.zeros(ring, 32, 32), torch.zeros(ring, 32, 32)
    for i, o in enumerate(outs):
        cur = torch.stack([o[..., 0].mean(0), torch.exp(-o[..., 1]).mean(0)])
        out_avg = cur if out_avg is None else out_avg * expw + cur * (1 - expw)
        r_epi[i % ring], r_ale[i % ring] = cur[0].clamp(0, 1), cur[1].clamp(0, 1)
    psnr = lambda a, b: 10 * math.log10(1.0 / float(((a - b) ** 2).mean()))
    last = torch.stack([outs[-1][..., 0].mean(0), torch.exp(-outs[-1][..., 1]).mean(0)])
    assert abs(m["psnr_noisy"] - psnr(noisy, last[0].clamp(0, 1))) < 1e-4
    assert abs(m["psnr_gt"] - psnr(gt, last[0].clamp(0, 1))) < 1e-4
    assert abs(m["psnr_gt_sm"] - psnr(gt, out_avg[0].clamp(0, 1))) < 1e-4
    assert abs(m["ssim_gt_sm"] - O.ssim(gt[None, None], out_avg[0].clamp(0, 1)[None, None])) < 1e-5
    assert torch.allclose(bk.out_avg, out_avg, atol=1e-6)
    assert torch.allclose(epi, r_epi.var(0), atol=1e-7) and torch.allclose(ale, r_ale.mean(0), atol=1e-6)
    ref_err2 = ((r_epi - gt) ** 2).mean(0)
    assert torch.allclose(err2, ref_err2, atol=1e-6)
    assert abs(uce - float(uceloss(ref_err2.reshape(-1), (r_epi.var(0) + r_ale.mean(0)).reshape(-1), n_bins=15)[0])) < 1e-6


_NAN_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from tests.plan_interpreter import TrainerInterpreter
from tests.test_trainer_cpu import _trainer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
tr = _trainer(4, rank, world, task="ct")               # the CT trainer switches the NaN guard on by default
assert tr.nan_guard
theta0 = tr.eng.theta.clone()
clean = tr.head.sino_t.clone()
with TrainerInterpreter(tr):
    if rank == 1:
        tr.head.sino_t[0] = float("nan")               # ONE rank sees a non-finite data loss
    tr.step()
    skipped = torch.equal(tr.eng.theta, theta0) and int(tr.adam_dev) == 0 and tr.steps_done == 1
    tr.head.sino_t.copy_(clean)
    tr.step()                                          # a clean step: Adam's first update (bias correction of step 1)
moved = float((tr.eng.theta - theta0).abs().max())
mine = tr.eng.theta.clone()
other = mine.clone()
dist.broadcast(other, src=0)
print(f"rank {rank} skipped={skipped} adam_steps={int(tr.adam_dev)} moved={moved:.2e} identical={torch.equal(mine, other)}")
assert skipped, "a NaN loss on one rank must skip the update on EVERY rank"
assert int(tr.adam_dev) == 1 and tr.steps_done == 2 and torch.isfinite(tr.eng.theta).all()
assert 0.5e-2 < moved < 1.5e-2                         # first Adam update moves every parameter by ~lr (1e-2): bias correction of step 1
assert torch.equal(mine, other)
dist.destroy_process_group()
'''


def test_two_rank_nan_guard_is_rank_consistent(tmp_path):
    """CT runner's NaN guard (reference bayesian_optimization.py:577-582) under MC-sample sharding: the tested scalar is
    nll + temp*kl carried through the gradient all-reduce, so a non-finite loss on ONE rank skips AdamW on BOTH (no rank applies
    NaN gradients, replicas stay bit-identical), and the optimiser's own step count does not advance on the skipped update."""
    script = tmp_path / "nan_worker.py"
    script.write_text(_NAN_WORKER)
    port = 33600 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=dict(os.environ, OMP_NUM_THREADS="2"), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 skipped=True" in r.stdout and "rank 1 skipped=True" in r.stdout
