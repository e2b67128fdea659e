#!/usr/bin/env python
"""MFVI-DIP ELBO steps/sec benchmark (BASELINE.json metric: 256x256 denoise net, MC=8).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config den|sr|inp|ct] [--math tf32|fp32]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full optimiser step of the hot path (input jitter, S-sample forward, data loss, KL, backward,
[all-reduce], AdamW) on synthetic data of the named shape.  `value` = steps/s with everything resident in HBM
(CUDA-graph replay, CUDA-event timing, max over ranks); `e2e` = the same through MfviDipTrainer.step_from_host with
pinned HOST buffers (H2D of net input + target and D2H of the loss inside the timed region).  MC samples are split
across GPUs (total work fixed => "strong" scaling).

--config selects the BASELINE.json configuration (default `den` = the metric configuration):
    den  test_configs/mfvi_den.json   256x256 denoising, MC = 8          (config 1 shape at the metric's MC)
    sr   test_configs/mfvi_sr.json    512x512 4x super-resolution, MC = 1 (config 2)
    inp  test_configs/mfvi_inp.json   512x512 inpainting, MC = 8          (config 3; 8 GPUs: one sample per GPU)
    ct   test_configs/mfvi_ct.json    512x512 sparse-view CT, 90 angles, MC = 1 (config 4; adds a `radon` object)
--math: tf32 = tcgen05 kind::tf32 tensor-core convolutions (the separately stated reduced-precision mode, tested to the bars in
tests/test_gpu_fullsize_parity.py); fp32 = CUDA-core convolutions, the mode held to north_star's rtol 1e-3.  At N=1 the JSON
line carries BOTH in `modes`.  `--impl reference` times the CPU restatement of the reference step (oracle/, PyTorch fp32 on all
host threads; /root/reference itself does not exist on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "mfvi_dip_elbo_steps_per_sec", "steps/s"

# hyper-parameters of test_configs/mfvi_{den,sr,inp,ct}.json (temp, sigma, lr, input_depth) and the benchmark shape
CONFIGS = {
    "den": dict(temp=5.656911698337764e-07, sigma=1.4616642493692077e-05, lr=1e-3, depth=16, size=256, mc=8,
                name="mfvi_den {n}x{n} 5-scale skip net"),
    "sr": dict(temp=4.3817e-07, sigma=4.9e-08, lr=1e-3, depth=32, size=512, mc=1, name="mfvi_sr {n}x{n} 4x SR 5-scale skip net"),
    "inp": dict(temp=1e-12, sigma=6.506e-4, lr=2e-3, depth=16, size=512, mc=8,
                name="mfvi_inp {n}x{n} inpainting 6-scale 5x5 net"),
    "ct": dict(temp=2.2e-10, sigma=1.7e-7, lr=1e-3, depth=16, size=512, mc=1,
               name="mfvi_ct {n}x{n} sparse-view CT (90 angles) 5-scale skip net"),
}
TESTED_TO = {
    "tf32": "tests/test_gpu_fullsize_parity.py BARS['tf32'] vs the imported reference (tcgen05 kind::tf32 operands, fp32 "
            "accumulate; the separately stated reduced-precision mode)",
    "fp32": "tests/test_gpu_fullsize_parity.py BARS['fp32'] vs the imported reference: north_star rtol 1e-3 (fp32 CUDA-core convs)",
}


def workload_name(cfg, size, mc):
    """The SAME string in both arms (the driver compares them)."""
    return f"{cfg['name'].format(n=size)}, MC={mc}, AdamW"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_problem(config, size, seed=1):
    """(net_input (1,C,H,W), head kwargs) of a BASELINE configuration on the synthetic phantoms of SURVEY.md section 8d.
    For CT the head carries the phantom (`image`); its sinogram is produced by the caller's projector (ours on the GPU,
    the oracle's on the CPU)."""
    import torch
    from mfvi_dip_mia_b200.utils import phantoms as ph
    cfg = CONFIGS[config]
    g = torch.Generator().manual_seed(seed)
    net_input = torch.rand(1, cfg["depth"], size, size, generator=g) * 0.1             # get_noise('noise', 'u', 1/10)
    if config == "den":
        head = dict(target=torch.from_numpy(ph.noisy(ph.ellipse_phantom(size), 0.1, seed))[None])
    elif config == "sr":
        head = dict(target=torch.from_numpy(ph.ellipse_phantom(size))[None][:, :, ::4, ::4].contiguous())
    elif config == "inp":
        head = dict(target=torch.from_numpy(ph.rgb_phantom(size))[None], mask=torch.from_numpy(ph.random_mask(size, 2))[None])
    else:
        head = dict(theta_deg=torch.arange(0, 180., step=2.), image=torch.from_numpy(ph.shepp_logan(size))[None])
    return net_input, head


def spec_kwargs(config):
    d = CONFIGS[config]["depth"]
    if config == "inp":
        return dict(num_input_channels=d, num_output_channels=4, down=(16, 32, 64, 128, 128, 128), up=(16, 32, 64, 128, 128, 128),
                    skip=(0,) * 6, filter_down=5, filter_up=3, filter_skip=1, need1x1_up=False, upsample_mode="nearest")
    return dict(num_input_channels=d, num_output_channels=1 if config == "ct" else 2)


def oracle_cfg(config):
    from oracle import mfvi_oracle as O
    k = spec_kwargs(config)
    if config == "inp":
        return O.SkipCfg(k["num_input_channels"], 4, k["down"], k["up"], k["skip"], 5, 3, 1, False, False, "nearest")
    return O.SkipCfg(k["num_input_channels"], k["num_output_channels"])


def oracle_stepper(config, size, mc, device="cpu"):
    """The reference step restated with PyTorch (oracle/cpu_step.py) on `device` for a BASELINE configuration."""
    from oracle import mfvi_oracle as O
    from oracle.cpu_step import OracleStepper
    cfg = CONFIGS[config]
    _, head = synthetic_problem(config, size)
    if config == "ct":
        img = head.pop("image")
        head["sino"] = O.radon_forward(img, head["theta_deg"])
    return OracleStepper(oracle_cfg(config), size, size, mc_samples=mc, temp=cfg["temp"], sigma=cfg["sigma"], lr=cfg["lr"], seed=1,
                         task=config, device=device, head=head)


def time_oracle(config, size, mc, device, budget_s, n_steps, warmup):
    """Bounded sample: probe one sample-forward/backward, evaluate `s_eval` of the mc samples per step so that the run fits the
    budget, scale linearly.  Returns (seconds per full step, s_eval)."""
    from oracle.cpu_step import time_steps
    st = oracle_stepper(config, size, mc, device)
    t0 = time.perf_counter()
    st.step(1)
    if st.device.type != "cpu":
        import torch
        torch.cuda.synchronize()
    probe = time.perf_counter() - t0
    s_eval = max(1, min(mc, int(budget_s / max(probe * (n_steps + warmup), 1e-9))))
    sec = time_steps(st, n_steps, warmup, s_eval)
    return sec * mc / s_eval, s_eval


PORT_NOTE = ("oracle port of the reference step (the same ATen kernels for conv/BN/pad/upsample/AdamW; its closed-form kl() is "
             "CHEAPER than the reference's 1951-op torch.distributions kl(), so this baseline is if anything faster than the "
             "reference itself)")


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CONFIGS[args.config]
    sec_full, s_eval = time_oracle(args.config, args.size, args.mc, "cpu", 240.0, args.steps, args.warmup)
    v = 1.0 / sec_full
    sample = (f"{args.steps} steps x {s_eval} of {args.mc} MC samples per step (time scaled by {args.mc}/{s_eval}), "
              f"{args.size}x{args.size}, {PORT_NOTE}, torch {torch.__version__} CPU fp32")
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg, args.size, args.mc)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def kernel_breakdown(tr, iters=3):
    """Per-kernel device time of one eager step (CUDA events around every C-ABI call)."""
    import torch
    from mfvi_dip_mia_b200 import _lib as L
    agg = {}
    for _ in range(iters):
        L.timeline = []
        tr._step_eager()
        torch.cuda.synchronize()
        tl, L.timeline = L.timeline, None
        for name, e0, e1, meta in tl:
            a = agg.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0, "samples": 0.0})
            a["ms"] += e0.elapsed_time(e1)
            a["n"] += 1
            if meta:
                a["flops"] += meta.get("flops", 0.0)
                a["bytes"] += meta.get("bytes", 0.0)
                a["samples"] += meta.get("samples", 0.0)
    for a in agg.values():
        for k in a:
            a[k] /= iters
    return agg


def family_in_graph_ms(tr, family, reps=20, replays=3):
    """Device time of every launch of one kernel family of the step plan, measured INSIDE a CUDA graph: each op of the family is
    captured `reps` times back to back (its own arguments, programmatic dependent launch on, as in the step's graph) and the
    replay is timed with CUDA events; returns (sum over the family's launches of time / reps in ms, number of launches).
    The eager per-launch events of kernel_breakdown also contain the host's launch latency between the first event and the
    kernel (a few us per launch), which a graph replay does not have."""
    import torch
    from mfvi_dip_mia_b200 import _lib as L
    total, n = 0.0, 0
    for name, a, meta in tr.eng.fwd_ops + tr.eng.bwd_ops:
        if name != family:
            continue
        for _ in range(2):
            L.call(name, *a)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                L.call(name, *a)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1) / (reps * replays)
        n += 1
    return total, n


def build_trainer(args, math_mode, dev, rank, world):
    from mfvi_dip_mia_b200 import MfviDipTrainer, SkipSpec
    from mfvi_dip_mia_b200.radon import FastRadonTransform
    cfg = CONFIGS[args.config]
    net_input, head = synthetic_problem(args.config, args.size)
    if args.config == "ct":
        img = head.pop("image").to(dev)
        head["sino"] = FastRadonTransform(tuple(img.shape), head["theta_deg"]).to(dev)(img).detach()
    return MfviDipTrainer(SkipSpec(**spec_kwargs(args.config)), args.config, net_input, temp=cfg["temp"], sigma=cfg["sigma"],
                          lr=cfg["lr"], mc_samples=args.mc, seed=1, device=dev, rank=rank, world_size=world, math_mode=math_mode,
                          use_graph=True, **head)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mfvi_dip_mia_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: native libraries (NCCL's version banner) write to fd 1 as well, so fd 1 is
    # pointed at stderr for the duration of the run and the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    # A collective that never completes (a rank died, mismatched call counts) would otherwise hang until the caller's
    # timeout: every rank gives up on its own — after 10 minutes alone (the run includes the CPU baseline leg), after 5 minutes
    # in a multi-rank run (no CPU leg; a healthy 8-GPU run takes well under a minute) — rank 0 reporting why on the JSON channel.
    state = {"printed": False}
    limit_s = 600.0 if world == 1 else 300.0

    def _give_up():
        if rank == 0 and not state["printed"]:
            os.write(json_fd, (json.dumps({"metric": METRIC, "error": f"bench.py watchdog: no result after {limit_s:.0f} s "
                                           "(hung collective or device?)", "n_gpus": world}) + "\n").encode())
        os._exit(0 if state["printed"] else 3)
    watchdog = threading.Timer(limit_s, _give_up)
    watchdog.daemon = True
    watchdog.start()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    cfg = CONFIGS[args.config]
    math_of = {"tf32": L.MATH_TF32, "fp32": L.MATH_FP32}
    tr = build_trainer(args, math_of[args.math], dev, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(args.warmup, 3)
    for _ in range(warm + 5):   # the first 2 steps run eagerly, the third captures the graph: 5 more replays settle NCCL / clocks
        tr.step()
    torch.cuda.synchronize()
    if args.profile:             # under ncu: nothing but the timed steps after the warm-up (a number printed here is not a bench value)
        ms_total = timed(tr.step, args.steps)
        if rank == 0:
            os.write(json_fd, (json.dumps({"metric": METRIC, "profile_run": True, "ms_per_step": ms_total / args.steps,
                                           "launches_per_step": tr.launches_per_step,
                                           "config": {"workload": workload_name(cfg, args.size, args.mc)}}) + "\n").encode())
        state["printed"] = True
        return
    with ClockSampler(local) as clk:
        ms_total = timed(tr.step, args.steps)
        # the loss of optimiser step number warm + steps: the same step index on every GPU count (eps is keyed by the global
        # sample id, parameters start from the same seed), so the N = 1 / 2 / 4 / 8 lines show G-invariance of the trajectory
        nll_k, kl_k, loss_k = tr.loss_terms()
        if world > 1:       # the data term is a mean over the samples: average the per-rank means
            t = torch.tensor([nll_k], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            nll_k = float(t.item()) / world
            loss_k = nll_k + cfg["temp"] * kl_k
        step_index = tr.steps_done
        # K steps can be shorter than nvidia-smi's sampling period: keep the SAME load running (untimed) for ~1 s so that the
        # clock record describes the loaded state.  The number of extra steps is derived from ms_total, which is already the
        # max over ranks and therefore IDENTICAL on every rank — every rank must issue the same number of all-reduces (a
        # per-rank "until my sampler has 3 rows" loop deadlocks NCCL as soon as two ranks disagree; seen at 8 GPUs).
        extra = int(min(4000, max(25, 1000.0 / max(ms_total / args.steps, 1e-3))))
        for _ in range(extra):
            tr.step()
        torch.cuda.synchronize()
    clocks = clk.summary()
    clocks["window"] = f"timed region + {extra} untimed steps of the same load"
    launches = tr.launches_per_step * args.steps
    value = args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, loss out, through the public trainer call
    # Host-driven loop, pipelined one step deep (two sets of pinned host buffers): every step copies ITS net input and
    # target host->device and its [kl, nll] device->host; the host reads the loss of step i-1 while step i runs.
    bufs = [tr.host_buffers() for _ in range(2)]
    for h_in, h_tgt, _ in bufs:
        h_in.copy_(tr.saved.cpu())
        h_tgt.copy_(tr._target_tensor().cpu())
    io = [0, 0]
    pending = []
    losses = []

    def e2e_step():
        h_in, h_tgt, h_res = bufs[len(losses) % 2]
        io[0], io[1], done = tr.step_from_host(h_in, h_tgt, h_res)
        pending.append((done, h_res))
        if len(pending) == 2:                          # read the previous step's loss while this one runs
            pending[0][0].synchronize()
            losses.append(float(pending[0][1][1]))
            pending.pop(0)
        else:
            losses.append(None)
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e = {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1],
           "pipeline_depth": 1}

    # ---- roofline of the dominant kernel family (eager pass with CUDA events around every C-ABI launch)
    pk = peaks()
    agg = kernel_breakdown(tr)
    step_ms_eager = sum(a["ms"] for a in agg.values())
    dom = max(agg, key=lambda k: agg[k]["ms"])
    a = agg[dom]
    # measured DRAM bytes per step of every kernel family: ncu dram__bytes_read.sum + dram__bytes_write.sum of THIS round's
    # build (scripts/traffic_from_ncu.py writes profiles/r02_dram_traffic.json from the committed launch list); null otherwise
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if os.path.exists(tp) and args.config == "den" and args.size == 256 and args.mc == 8 and world == 1 and args.math == "tf32":
        with open(tp) as f:
            traffic = json.load(f).get("bytes_per_step", {}).get(dom)
    # duration of the family's launches as they run in the timed region: inside a CUDA graph (family_in_graph_ms); the eager
    # per-launch events above also hold the host's launch latency and are kept next to it
    fam_ms, fam_n = family_in_graph_ms(tr, dom)
    timing = "in-graph: every launch of the family captured 20x back to back and replayed, CUDA events around the replay"
    if fam_n != int(round(a["n"])) or fam_ms <= 0.0:          # a family outside the engine plan (trainer-level op): eager events
        fam_ms, timing = a["ms"], "eager: CUDA events around every C-ABI launch"
    step_ms_graph = ms_total / args.steps
    if a["flops"] > 0:
        # the convolutions run tcgen05 kind::tf32: half the dense bf16 rate MEASURED_PEAKS.json reports
        peak = {"tf32": pk["tensor"] / 2.0, "fp32": 75.0}[args.math]
        ach = a["flops"] / (fam_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic,
                "peak_source": {"tf32": pk["src"] + " bf16 sustained / 2 (kind::tf32)", "fp32": "fp32 CUDA-core nominal"}[args.math],
                "algorithmic_gbs": a["bytes"] / (fam_ms * 1e-3) / 1e9, "hbm_peak_gbs": pk["hbm"],
                "launches_per_step": a["n"], "ms_per_step": fam_ms, "share_of_step": fam_ms / step_ms_graph,
                "timing": timing, "ms_per_step_eager_events": a["ms"], "share_of_eager_step": a["ms"] / step_ms_eager}
    else:
        ach = a["bytes"] / (fam_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": traffic, "peak_source": pk["src"], "launches_per_step": a["n"], "ms_per_step": fam_ms,
                "share_of_step": fam_ms / step_ms_graph, "timing": timing, "ms_per_step_eager_events": a["ms"],
                "share_of_eager_step": a["ms"] / step_ms_eager}
    kernels = {}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(v["ms"], 4), "n": v["n"], "share": round(v["ms"] / step_ms_eager, 4)}
        if v["flops"]:
            e["tflops"] = round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3)
        if v["bytes"]:
            e["gbs"] = round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)
        kernels[k] = e

    # ---- CT: the radon projector on its own (samples/s and GB/s under both byte definitions)
    radon = None
    if args.config == "ct":
        radon = {}
        H = args.size
        T = tr.head.T
        for name in ("mfvi_radon_fwd", "mfvi_radon_bwd"):
            v = agg[name]
            sec = v["ms"] * 1e-3
            ref_bytes = tr.S * (T * H * H * 8.0 + 2.0 * T * H * H * 4.0)     # reference: grid read + (T,C,H,W) intermediate write+read
            radon[name] = {"ms": round(v["ms"], 4), "bilinear_samples_per_s": v["samples"] / sec,
                           "algorithmic_gbs": v["bytes"] / sec / 1e9, "frac_of_hbm_peak": v["bytes"] / sec / 1e9 / pk["hbm"],
                           "gbs_under_reference_traffic_model": ref_bytes / sec / 1e9,
                           "frac_of_hbm_peak_reference_model": ref_bytes / sec / 1e9 / pk["hbm"]}
        radon["note"] = (f"{T} angles, {H}x{H}, S={tr.S}: algorithmic bytes = image + sinogram once each "
                         "(gather/issue-bound at this size); reference model = the bytes radon/radon.py moves")

    # ---- the other arithmetic mode on the same workload (N=1 only): fp32 next to tf32
    modes = None
    if world == 1 and not args.no_modes:
        modes = {args.math: {"value": value, "unit": UNIT, "ms_per_step": ms_total / args.steps, "tested_to": TESTED_TO[args.math]}}
        other = "fp32" if args.math == "tf32" else "tf32"
        del bufs
        tr2 = build_trainer(args, math_of[other], dev, rank, world)
        for _ in range(warm):
            tr2.step()
        n2 = max(5, args.steps // 3)
        ms2 = timed(tr2.step, n2)
        modes[other] = {"value": n2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / n2, "steps": n2, "tested_to": TESTED_TO[other]}
        del tr2
        torch.cuda.empty_cache()

    if rank != 0:
        return
    # ---- the reference step as eager PyTorch on this GPU (cuDNN fp32, TF32 off): the GPU bar of SURVEY section 2a
    gpu_eager = None
    if world == 1 and not args.no_cpu:
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            sec, s_eval = time_oracle(args.config, args.size, args.mc, dev, 10.0, 3, 2)
            gpu_eager = {"value": 1.0 / sec, "unit": UNIT, "ms_per_step": sec * 1e3,
                         "what": f"{PORT_NOTE} run as eager PyTorch {torch.__version__} on this GPU (cuDNN fp32, allow_tf32=False), "
                                 f"3 steps x {s_eval} of {args.mc} MC samples (time scaled), wall clock with synchronise"}
        except Exception as e:                           # never let the comparison arm break the bench line
            gpu_eager = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    # ---- CPU baseline on this box's host cores (bounded sample), N=1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_cpu = 3
        sec, s_eval = time_oracle(args.config, args.size, args.mc, "cpu", 20.0, n_cpu, 1)
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} steps x {s_eval} of {args.mc} MC samples (time scaled by {args.mc}/{s_eval}), {PORT_NOTE}, "
                         f"torch CPU fp32"}
    ws_mb = sum(t.numel() for t in tr.eng._bufs) * 4 / 2 ** 20
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"tf32": "tf32", "fp32": "f32"}[args.math], "data": "synthetic",
            "config": {"workload": workload_name(cfg, args.size, args.mc), "mc_split": f"{args.mc} MC samples over {world} GPU(s)",
                       "mc_per_gpu": tr.S, "l2": f"per-step activation working set {ws_mb:.0f} MiB > 126 MB L2 (no flush needed)",
                       "cuda_graph": True, "loss_at_step": {"step": step_index, "nll": nll_k, "kl": kl_k, "loss": loss_k}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launches_per_step": tr.launches_per_step, "roofline": roof,
            "kernels": kernels}
    if modes is not None:
        line["modes"] = modes
    if radon is not None:
        line["radon"] = radon
    if gpu_eager is not None:
        line["gpu_eager_baseline"] = gpu_eager
    if cpu is not None:
        line["cpu_baseline"] = cpu
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    state["printed"] = True            # the watchdog stays armed for the final rendezvous, but no longer reports an error


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="den", choices=sorted(CONFIGS))
    ap.add_argument("--mc", type=int, default=None, help="MC samples per step (default: the configuration's)")
    ap.add_argument("--size", type=int, default=None, help="image size (default: the configuration's)")
    ap.add_argument("--math", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--no-modes", action="store_true", help="skip timing the other arithmetic mode")
    ap.add_argument("--profile", action="store_true", help="profiling run (ncu): warm-up + the K timed steps only, minimal JSON line")
    args = ap.parse_args()
    args.mc = args.mc or CONFIGS[args.config]["mc"]
    args.size = args.size or CONFIGS[args.config]["size"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            # Captured CUDA graphs hold NCCL kernels: tearing the communicator down under them can hang at exit.
            # All ranks rendezvous once more, then leave without running the NCCL/CUDA destructors.
            import torch
            import torch.distributed as dist
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier()
                torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)


if __name__ == "__main__":
    main()
