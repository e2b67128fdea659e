"""CPU-only checks (no GPU, no compute calls into the library): the C-ABI library loads and exports every symbol
include/mfvi_dip.h declares, the host-side mirror of the reference interface builds the reference's module tree /
state-dict keys, the plugin surface refuses to run without CUDA (no silent fallback), and the N>1 sharding maths holds
under a world_size-2 gloo group."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, load_npz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mfvi_dip_mia_b200 import _lib
    with open(os.path.join(ROOT, "include", "mfvi_dip.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(mfvi_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for sym in sorted(declared):
        assert hasattr(_lib.lib, sym), f"libmfvidip.so lacks {sym}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib.mfvi_abi_version() == _lib.ABI_VERSION
    # the struct mirrors must match the C layout (x86-64 SysV)
    import ctypes as C
    assert C.sizeof(_lib.View) == 24 and C.sizeof(_lib.ConvDesc) == 44 and C.sizeof(_lib.PhiloxKey) == 24


def _ref_keys():
    with open(os.path.join(GOLDEN, "full_net_keys.json")) as f:
        return json.load(f)


def _build(task):
    from mfvi_dip_mia_b200.models import get_net
    from mfvi_dip_mia_b200.models.skip import skip
    if task == "inp":      # bayesian_optimization.py:2970-2998 of the reference
        return skip(16, 4, num_channels_down=[16, 32, 64, 128, 128, 128], num_channels_up=[16, 32, 64, 128, 128, 128],
                    num_channels_skip=[0] * 6, filter_size_down=5, filter_size_up=3, filter_skip_size=1,
                    need_sigmoid=False, need_bias=True, pad="reflection", upsample_mode="nearest", need1x1_up=False,
                    dropout_mode_down="None", dropout_mode_up="None", dropout_mode_skip="None", dropout_mode_output="None")
    depth, out = {"den": (16, 2), "sr": (32, 2), "ct": (16, 1)}[task]
    return get_net(depth, "skip", "reflection", skip_n33d=[16, 32, 64, 128, 128], skip_n33u=[16, 32, 64, 128, 128],
                   skip_n11=4, num_scales=5, n_channels=out, upsample_mode="bilinear")


@pytest.mark.parametrize("task", ["den", "sr", "ct", "inp"])
def test_meanfieldvi_state_dict_matches_reference(task):
    from mfvi_dip_mia_b200 import build_layout
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    ref = _ref_keys()[task]
    torch.manual_seed(1)
    net = MeanFieldVI(_build(task), prior={"mu": 0.0, "sigma": 1e-8}, replace_layers="all", reparam="")
    sd = net.state_dict()
    assert list(sd.keys()) == ref["keys"]                       # same keys in the same order
    assert [list(v.shape) for v in sd.values()] == ref["shapes"]
    assert sum(p.numel() for p in net.parameters()) == ref["n_params"]
    # identical RNG consumption => identical initial values (reference built under torch.manual_seed(1))
    assert abs(float(sum(p.double().sum() for p in net.parameters())) - ref["param_sum"]) < 1e-6 * ref["param_abs_sum"]
    lay = build_layout(net.net._skip_spec)
    assert 2 * lay.P + 2 * lay.Q == ref["n_params"]
    assert {f"net.{c.key}.W_mu" for c in lay.convs} <= set(ref["keys"])
    assert {f"net.{b.key}.weight" for b in lay.bns} <= set(ref["keys"])


def test_replace_layers_substring_semantics():
    """freq_to_bayes.py:55: the key of the LEAF module must contain `replace_layers` ('up' -> 11 of 26)."""
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    from mfvi_dip_mia_b200.BayTorch.modules import Conv2dRT
    counts = {}
    for mode in ["all", "up", "down", "none"]:
        net = MeanFieldVI(_build("den"), prior={"mu": 0.0, "sigma": 0.1}, replace_layers=mode, reparam="")
        counts[mode] = sum(isinstance(m, Conv2dRT) for m in net.modules())
    assert counts == {"all": 26, "up": 11, "down": 0, "none": 0}
    # the reference's default reparam='local' builds local-reparameterisation layers (freq_to_bayes.py:22-25)
    from mfvi_dip_mia_b200.BayTorch.modules import Conv2dLRT
    local = MeanFieldVI(_build("den"))
    assert sum(isinstance(m, Conv2dLRT) for m in local.modules()) == 26 and not local.fused
    assert set(local.state_dict()) == set(MeanFieldVI(_build("den"), reparam="").state_dict())


def test_no_cpu_fallback():
    from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec, _lib
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    from mfvi_dip_mia_b200.BayTorch.modules import Conv2dRT
    from mfvi_dip_mia_b200.radon import FastRadonTransform
    from mfvi_dip_mia_b200.utils.bayesian_utils import gaussian_nll
    net = MeanFieldVI(_build("den"), prior={"mu": 0.0, "sigma": 0.1}, replace_layers="all", reparam="")
    with pytest.raises(_lib.MfviError):
        net(torch.zeros(1, 16, 32, 32))
    with pytest.raises(_lib.MfviError):
        Conv2dRT(4, 4, 3)(torch.zeros(1, 4, 8, 8))
    with pytest.raises(_lib.MfviError):
        gaussian_nll(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))
    with pytest.raises(_lib.MfviError):
        FastRadonTransform((1, 1, 8, 8), torch.arange(0., 180., 45.))(torch.zeros(1, 1, 8, 8))
    with pytest.raises(_lib.MfviError):
        MfviDipTrainer(SkipSpec(), "den", torch.zeros(1, 16, 32, 32), temp=1e-6, sigma=1e-4, lr=1e-3, device="cpu",
                       target=torch.zeros(1, 1, 32, 32))


def test_shard_samples():
    from mfvi_dip_mia_b200.sharding import shard_samples
    assert [shard_samples(8, r, 4) for r in range(4)] == [(2, 0), (2, 2), (2, 4), (2, 6)]
    assert shard_samples(8, 0, 1) == (8, 0)
    with pytest.raises(ValueError):
        shard_samples(8, 0, 3)
    with pytest.raises(ValueError):
        shard_samples(8, 4, 4)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from oracle import mfvi_oracle as O
from mfvi_dip_mia_b200.sharding import allreduce_mean_, shard_samples
from tests.helpers import group, load_npz
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
d = load_npz("skipnet_small_den.npz")
cfg = O.SkipCfg(4, 2, (8, 16, 16), (8, 16, 16), (2, 2, 2), 3, 3, 1, True, False, "bilinear")
S = int(d["S"]); temp, sigma = float(d["temp"]), float(d["sigma"])
sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in group(d, "sd/").items()}
s_local, s0 = shard_samples(S, rank, world)
eps = [group(d, f"eps{s}/") for s in range(s0, s0 + s_local)]          # eps keyed by GLOBAL sample id
loss, nll, kl, _ = O.mfvi_loss(sd, cfg, torch.from_numpy(d["net_input"]), eps, task="den", temp=temp,
                               prior_sigma_plus_eps=O.prior_scale(temp, sigma), target=group(d, "extra/")["target"])
loss.backward()                                                          # local: mean over OWN samples + T*KL
names = sorted(group(d, "grad/").keys())
flat = torch.cat([sd[k].grad.reshape(-1) for k in names])
allreduce_mean_(flat)                                                    # the step's only collective
ref = torch.cat([group(d, "grad/")[k].reshape(-1) for k in names])
err = float((flat - ref).abs().max() / ref.abs().max())
print(f"rank {rank} err {err:.3e}")
assert err < 1e-4, err
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharded_gradient_equals_single_process(tmp_path):
    """world_size 2 on CPU (gloo): each rank evaluates its own MC-sample shard with the oracle, the gradients are
    averaged with the trainer's collective helper, and every rank ends with the single-process S=2 reference gradient."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 err" in r.stdout and "rank 1 err" in r.stdout


def test_bench_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--size", "64", "--mc", "2"], capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ["metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert k in line, k
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def test_trial_fanout_round_robin_and_nan_filter():
    """eval_trials: one process per candidate, devices round-robin, NaN / crashed trials dropped
    (reference bayesian_optimization.py:3756-3781)."""
    from mfvi_dip_mia_b200.runners import eval_trials, log_grid
    from tests.helpers import fake_trial
    cands = [(1.0, 0.1), (2.0, 0.1), (3.0, 0.1), (4.0, 0.1), (5.0, 0.01)]
    X, Y = eval_trials(cands, ["cpu:0", "cpu:1"], fake_trial, {"offset": 1.0}, start_method="fork")
    assert X == [(1.0, 0.1), (2.0, 0.1), (5.0, 0.01)]
    import math
    exp = [1.0 + math.log10(t) - 2 * math.log10(s) + (0.0 if i % 2 == 0 else 0.5) for i, (t, s) in zip([0, 1, 4], X)]
    assert all(abs(a - b) < 1e-12 for a, b in zip(Y, exp))
    g = log_grid([[-10, 0], [-10, 0]], 8)
    assert len(g) == 64 and abs(g[0][0] - 1e-10) < 1e-20 and abs(g[-1][1] - 1.0) < 1e-12


def test_trial_fanout_persistent_workers():
    """eval_trials(persistent=True): one worker per device runs its trials back to back; a NaN trial and a raising trial are
    dropped, a trial that kills its worker is dropped and the worker replaced, the rest of the candidates still finish."""
    import math
    from mfvi_dip_mia_b200.runners import eval_trials
    from tests.helpers import fake_trial
    cands = [(1.0, 0.1), (2.0, 0.1), (3.0, 0.1), (4.0, 0.1), (6.0, 0.1), (5.0, 0.01), (7.0, 0.1), (8.0, 0.1), (9.0, 0.1)]
    X, Y = eval_trials(cands, ["cpu:0", "cpu:1"], fake_trial, {"offset": 1.0}, start_method="fork", persistent=True)
    assert X == [c for c in cands if c[0] not in (3.0, 4.0, 6.0)]
    for (t, s), y in zip(X, Y):          # which device ran a trial depends on timing: the value is one of the two
        base = 1.0 + math.log10(t) - 2 * math.log10(s)
        assert min(abs(y - base), abs(y - base - 0.5)) < 1e-12
    X1, Y1 = eval_trials(cands[:1], ["cpu:0", "cpu:1"], fake_trial, {}, start_method="fork", persistent=True)
    assert X1 == cands[:1] and abs(Y1[0] - (math.log10(1.0) + 2.0)) < 1e-12


def test_trial_pool_keeps_its_workers_across_runs():
    """TrialPool (what bo() runs its rounds on): the worker processes of the first run serve the second one too."""
    from mfvi_dip_mia_b200.runners import TrialPool
    from tests.helpers import pid_trial
    import os
    with TrialPool(["cpu:0", "cpu:1"], pid_trial, start_method="fork") as pool:
        r1 = pool.run([(float(i + 1), 0.1) for i in range(6)])
        r2 = pool.run([(float(i + 10), 0.1) for i in range(5)])
        r3 = pool.run([(99.0, 0.1)])                                    # fewer candidates than workers
    pids1, pids2 = set(r1.values()), set(r2.values())
    assert len(r1) == 6 and len(r2) == 5 and len(r3) == 1
    assert 1 <= len(pids1) <= 2 and pids2 <= pids1 and set(r3.values()) <= pids1 and float(os.getpid()) not in pids1


# ----------------------------------------------------------------------------------------------- BO outer loop (row f4)
def test_gp_posterior_and_ei_match_closed_form():
    """ExactGPModel (constant mean + scaled RBF + Gaussian noise) against plain numpy GP algebra at the fitted
    hyper-parameters, and expected_improvement against scipy's normal pdf / cdf (reference :3604-3633)."""
    from scipy.stats import norm
    from mfvi_dip_mia_b200.bo import expected_improvement, train_gp
    g = torch.Generator().manual_seed(0)
    Xt = torch.rand(12, 2, generator=g, dtype=torch.float64)
    Yt = 20 + 5 * torch.sin(3 * Xt[:, 0]) * torch.cos(2 * Xt[:, 1]) + 0.05 * torch.randn(12, generator=g, dtype=torch.float64)
    gp = train_gp(Xt, Yt, iter_max=300)
    Xs = torch.rand(40, 2, generator=g, dtype=torch.float64)
    mean, var = gp.predict(Xs)
    ls, os_, nz, c = (float(v) for v in (gp.lengthscale, gp.outputscale, gp.noise, gp.constant))
    assert nz > 1e-4 and ls > 0 and os_ > 0
    k = lambda a, b: os_ * np.exp(-0.5 * ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1) / ls ** 2)
    A, B = Xt.numpy(), Xs.numpy()
    K = k(A, A) + nz * np.eye(12)
    m_ref = c + k(B, A) @ np.linalg.solve(K, Yt.numpy() - c)
    v_ref = os_ - np.einsum("ij,ji->i", k(B, A), np.linalg.solve(K, k(A, B)))
    assert np.allclose(mean.detach().numpy(), m_ref, atol=1e-8) and np.allclose(var.detach().numpy(), v_ref, atol=1e-8)
    # the fit explains the data and the MLL improved over the initial hyper-parameters
    assert float((gp.predict(Xt)[0] - Yt).abs().max()) < 1.0
    ei = expected_improvement(gp, Xs, Xt).detach().numpy().reshape(-1)
    best = float(gp.predict(Xt)[0].max())
    sd = np.sqrt(np.maximum(v_ref, 1e-9))
    u = (m_ref - best) / sd
    assert np.allclose(ei, np.maximum(sd * (norm.pdf(u) + u * norm.cdf(u)), 0), atol=1e-8)


def test_normalize_roundtrip_and_peak_local_max():
    from mfvi_dip_mia_b200.bo import normalize_X, peak_local_max, unnormalize_X
    X = torch.tensor([[1e-10, 1.0], [1e-5, 1e-5], [3.3e-7, 2e-2]], dtype=torch.float64)
    Xn = normalize_X(X, [-10, 0], [-10, 0])
    assert torch.allclose(Xn[0], torch.tensor([0.0, 1.0], dtype=torch.float64)) and torch.allclose(Xn[1], torch.tensor([0.5, 0.5], dtype=torch.float64))
    assert torch.allclose(unnormalize_X(Xn, [-10, 0], [-10, 0]), X, rtol=1e-12)
    yy, xx = np.mgrid[0:100, 0:100]
    img = (np.exp(-((yy - 30) ** 2 + (xx - 40) ** 2) / 50.0) + 0.6 * np.exp(-((yy - 70) ** 2 + (xx - 80) ** 2) / 30.0)
           + 0.05 * np.exp(-((yy - 10) ** 2 + (xx - 90) ** 2) / 20.0) + 2.0 * np.exp(-((yy - 2) ** 2 + (xx - 2) ** 2) / 8.0))
    pk = peak_local_max(img, min_distance=5, threshold_rel=0.1, num_peaks=4)
    # strongest first; the 0.05 bump is below 10 % of the maximum; the corner peak lies inside the excluded border
    assert pk.tolist() == [[30, 40], [70, 80]]


def test_bo_loop_closes_on_analytic_objective(tmp_path):
    """bo(): trial fan-out -> GP fit -> EI candidates -> next round, on an analytic objective with a known optimum; the
    per-round npz carries the reference's keys (bayesian_optimization.py:3876-3887)."""
    from mfvi_dip_mia_b200.bo import bo
    from tests.helpers import quadratic_trial
    bo_params = {"temp": {"logbounds": [-10, 0], "candidates": [1e-9, 1e-5, 1e-1]},
                 "sigma": {"logbounds": [-10, 0], "candidates": [1e-8, 1e-4, 1e-1]}}
    X, Y = bo(quadratic_trial, bo_params, {"bo_results_path": str(tmp_path), "devices": ["cpu:0", "cpu:1"]}, rounds=3,
              gp_iters=300, start_method="fork", verbose=False)
    assert len(X) == len(Y) and len(X) > 9                       # 9 initial candidates + proposals of 2 rounds
    z = np.load(tmp_path / "2_fig_data.npz")
    for k in ["XX_lr", "XX_wd", "pred", "observed_X", "observed_Y", "expected_improvement", "confidence", "acq", "candidates"]:
        assert k in z.files, k
    assert z["pred"].shape == (100, 100) and z["acq"].shape == (100, 100) and z["candidates"].shape[1] == 2
    assert (z["candidates"] >= 1e-10).all() and (z["candidates"] <= 1.0).all()
    # proposals move towards the optimum (1e-6, 1e-3): the best observation improves on the initial grid's best
    assert max(Y) > max(Y[:9]) - 1e-9 and max(Y) > 29.0


def test_initialisation_draw_order_matches_reference():
    """VIModule.reset_parameters / MeanFieldVI conversion draw the initial values from torch's global RNG in the reference's
    order (W_mu, W_rho, bias_mu, bias_rho per converted layer, module.py:56-62): under the fixture's seed every parameter
    norm equals the imported reference's (tests/golden/den256_summary.npz)."""
    from mfvi_dip_mia_b200.BayTorch import MeanFieldVI
    from mfvi_dip_mia_b200.models import get_net
    d = load_npz("den256_summary.npz")
    temp, sigma = float(d["temp"]), float(d["sigma"])
    torch.manual_seed(int(d["init_seed"]))
    net = get_net(16, "skip", "reflection", skip_n33d=[16, 32, 64, 128, 128], skip_n33u=[16, 32, 64, 128, 128],
                  skip_n11=4, num_scales=5, n_channels=2, upsample_mode="bilinear")
    net = MeanFieldVI(net, prior={"mu": 0.0, "sigma": np.sqrt(temp) * sigma}, replace_layers="all", reparam="")
    names = json.loads(str(d["grad_names"]))
    pn = dict(net.named_parameters())
    got = np.array([float(pn[k].detach().double().norm()) for k in names])
    assert np.allclose(got, d["param_norms"], rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# Convolution dispatch and tile plans, asked from the library's own host code in planning-only mode (no GPU needed)
_TASK_NETS = {
    "den": (dict(), 256),                                                              # test_configs/mfvi_den.json (metric shape)
    "sr": (dict(num_input_channels=32), 512),                                          # test_configs/mfvi_sr.json
    "ct": (dict(num_output_channels=1), 512),                                          # test_configs/mfvi_ct.json
    "inp": (dict(num_output_channels=4, down=(16, 32, 64, 128, 128, 128), up=(16, 32, 64, 128, 128, 128), skip=(0,) * 6,
                 filter_down=5, need1x1_up=False, upsample_mode="nearest"), 512),      # test_configs/mfvi_inp.json
}


def _cdiv(a, b):
    return -(-a // b)


@pytest.mark.parametrize("S", [1, 2, 8])
@pytest.mark.parametrize("task", sorted(_TASK_NETS))
def test_conv_dispatch_table_of_the_task_networks(task, S):
    """Every convolution launch (forward, data gradient, weight gradient) of the four task networks at their full sizes,
    for the per-GPU sample counts of 8-, 4- and 1-GPU runs: under MFVI_MATH_TF32 none falls back to the fp32 CUDA-core
    kernels, and every tile plan respects the hardware limits (227 KB shared memory, 512 TMEM columns) and covers its
    output exactly.  The plans come from mfvi_conv2d_plan, i.e. from the same host code that launches the kernels."""
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L
    kw, H = _TASK_NETS[task]
    eng = SkipEngine(SkipSpec(**kw), H, H, S, "meta", math=L.MATH_TF32)
    rows = eng.conv_dispatch_table()
    n_conv = len(eng.lay.convs)
    assert sum(r["op"] == "fwd" for r in rows) == n_conv and sum(r["op"] == "wgrad" for r in rows) == n_conv
    # the convolutions that read the network input have no data gradient (skip_1 / deeper_1)
    assert sum(r["op"] == "dgrad" for r in rows) == n_conv - (2 if eng.lay.scales[0].skip_conv is not None else 1)
    for r in rows:
        where = f"{task} S={S} {r['op']} {r['layer']} {r['shape']}"
        assert r["family"] in ("pointwise", "halo", "alias", "tc"), f"{where}: fell back to {r['family']}"
        assert r["smem_bytes"] <= 227 * 1024 and 32 <= r["block"] <= 1024 and min(r["grid"]) >= 1, where
        # the last layer has a bias gradient: a second launch, except in the pointwise kernel, which sums dy in the same pass
        last_wgrad = r["op"] == "wgrad" and r["layer"] == eng.lay.final.key.rsplit(".", 1)[-1]
        # and the alias weight gradient takes a 64-channel layer as two launches over 32-channel slices of dy
        alias_parts = max(1, r["Cout"] // 32) if (r["op"] == "wgrad" and r["family"] == "alias") else 1
        assert r["launches"] == alias_parts + (1 if (last_wgrad and r["family"] != "pointwise") else 0), where
        p = r["plan"]
        assert p.get("tmem_cols", 32) in (32, 64, 128, 256, 512), where
        if r["family"] == "halo":
            dgrad = r["op"] == "dgrad"
            Mh, Mw = (r["Hin"], r["Win"]) if dgrad else (r["Hout"], r["Wout"])
            if p["cls"] == 4:                                       # stride-2 dgrad: four output-parity classes of dx
                Mh, Mw = (Mh + 1) // 2, (Mw + 1) // 2
            assert p["n_mt"] * 128 >= (p["TH"] - 1) * p["Pw"] + p["TW"], where        # the M tiles cover the flat tile
            assert p["acc_stages"] * p["n_mt"] * p["BN"] <= p["tmem_cols"], where
            assert p["BN"] * p["nb"] >= (r["Cin"] if dgrad else r["Cout"]), where
            assert p["tiles"] == r["S"] * p["nb"] * p["cls"] * _cdiv(Mh, p["TH"]) * _cdiv(Mw, p["TW"]), where
            assert r["grid"][0] <= min(p["tiles"], 2 * 148), where                    # persistent CTAs, at most 2 per SM
    # exact-fp32 parity mode: every convolution runs on the CUDA-core kernels
    rows32 = SkipEngine(SkipSpec(**kw), H, H, S, "meta", math=L.MATH_FP32).conv_dispatch_table()
    assert len(rows32) == len(rows) and {r["family"] for r in rows32} == {"simt"}


def test_weight_gradient_families_of_the_metric_net():
    """Round-2 dispatch of the weight gradients: the 1x1 layers with <= 4 output channels take the
    pointwise streaming kernel (weight and bias gradient in ONE launch, also for the last layer), and 132->64 at 64^2 runs
    on the alias kernel as two 32-channel launches with 8 samples per GPU but stays on the per-tap kernel with one."""
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L
    kw, H = _TASK_NETS["den"]
    by_s = {}
    for S in (8, 1):
        eng = SkipEngine(SkipSpec(**kw), H, H, S, "meta", math=L.MATH_TF32)
        by_s[S] = {r["layer"]: r for r in eng.conv_dispatch_table() if r["op"] == "wgrad"}
    for S, rows in by_s.items():
        for name, r in rows.items():
            if r["K"] == 1 and r["Cout"] <= 4 and r["Cin"] <= 32:
                assert r["family"] == "pointwise" and r["launches"] == 1, (S, name, r["family"], r["launches"])
        assert rows["Conv2d_up_11"]["family"] == "pointwise" and rows["Conv2d_skip_1"]["family"] == "pointwise"
    assert by_s[8]["Conv2d_up_5"]["family"] == "alias" and by_s[8]["Conv2d_up_5"]["launches"] == 2
    assert by_s[1]["Conv2d_up_5"]["family"] == "tc" and by_s[1]["Conv2d_up_5"]["launches"] == 1
    assert by_s[8]["Conv2d_up_7"]["family"] == "alias" and by_s[8]["Conv2d_up_7"]["launches"] == 1


def test_plan_only_engine_cannot_execute():
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L
    eng = SkipEngine(SkipSpec(), 64, 64, 2, "meta")
    with pytest.raises(L.MfviError, match="plan only"):
        eng.forward()


@pytest.mark.parametrize("H,W", [(64, 96), (160, 64), (224, 352), (352, 288)])
def test_conv_plans_on_non_square_images(H, W):
    """The runners crop to multiples of 32, not to squares: tile plans keep their invariants on ragged sizes, only weight
    gradients over fewer than 16 pixels may leave the tensor cores (tf32)."""
    from mfvi_dip_mia_b200 import SkipEngine, SkipSpec, _lib as L
    for S in (1, 3):
        for r in SkipEngine(SkipSpec(), H, W, S, "meta", math=L.MATH_TF32).conv_dispatch_table():
            where = f"{H}x{W} S={S} {r['op']} {r['layer']} {r['shape']}"
            if r["family"] == "simt":
                assert r["op"] == "wgrad" and r["Hout"] * r["Wout"] < 16, where
                continue
            assert r["smem_bytes"] <= 227 * 1024, where
            if r["family"] == "halo":
                p = r["plan"]
                Mh, Mw = (r["Hin"], r["Win"]) if r["op"] == "dgrad" else (r["Hout"], r["Wout"])
                if p["cls"] == 4:
                    Mh, Mw = (Mh + 1) // 2, (Mw + 1) // 2
                assert p["n_mt"] * 128 >= (p["TH"] - 1) * p["Pw"] + p["TW"], where
                assert p["acc_stages"] * p["n_mt"] * p["BN"] <= p["tmem_cols"] <= 512, where
                assert p["tiles"] == r["S"] * p["nb"] * p["cls"] * _cdiv(Mh, p["TH"]) * _cdiv(Mw, p["TW"]), where

