"""The MFVI-DIP training step on flat buffers: the fast path behind the runners.

One `MfviDipTrainer.step()` is exactly the reference hot loop (bayesian_optimization.py:1361-1372, task variants
:2182-2188, :3033-3038, :576-582):

    zero_grad -> net_input = saved + N(0,1)*reg_noise_std -> out = net(net_input)   [S weight samples at once]
    -> nll (task head) -> kl -> loss = mean_s nll_s + temp*kl -> backward -> [all-reduce] -> AdamW

executed as a static list of libmfvidip kernels on one stream and replayed as a CUDA graph.  MC samples are
sharded over ranks (rank r owns global samples [r*S/G, (r+1)*S/G)); eps is keyed by the GLOBAL sample id, so
results do not depend on the number of GPUs beyond fp32 summation order; one NCCL all-reduce (average) of the
flat gradient per step is the only communication.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib as L
from .engine import KL, NLL, SkipEngine, SkipSpec
from .sharding import allreduce_mean_, shard_samples


class LossHead:
    """Data term of the ELBO and its gradient w.r.t. the network output (engine.out -> engine.dout)."""

    def __init__(self, eng: SkipEngine, task: str, *, target=None, mask=None, theta_deg=None, sino=None, sr_factor=4):
        self.eng, self.task = eng, task
        dev = eng.device
        S, H, W, Cn = eng.out.shape
        f = lambda t: None if t is None else torch.as_tensor(t).to(dev, torch.float32).contiguous()
        if task in ("den", "sr"):
            sub = 1 if task == "den" else int(sr_factor)
            t = f(target).reshape(-1)
            assert t.numel() == (H // sub) * (W // sub), "target must be (H/sub, W/sub)"
            self.target, self.sub = t, sub
            assert Cn >= 2
        elif task == "inp":
            t = f(target)
            assert Cn == 4 and t.numel() == 3 * H * W
            # reference layout (1,3,H,W) -> NHWC (H,W,3)
            self.target = t.reshape(3, H, W).permute(1, 2, 0).contiguous()
            self.mask = f(mask).reshape(H, W).contiguous()
        elif task == "ct":
            th = torch.as_tensor(theta_deg, dtype=torch.float64)
            self.theta = torch.deg2rad(th).to(dev, torch.float32).contiguous()
            self.T = self.theta.numel()
            self.sino_t = f(sino).reshape(-1)
            assert self.sino_t.numel() == Cn * self.T * W
            self.sino = torch.empty(S, Cn, self.T, W, dtype=torch.float32, device=dev)
            self.dsino = torch.empty_like(self.sino)
        else:
            raise ValueError(f"unknown task {task!r}")

    @property
    def device(self):
        return self.eng.device

    @L.on_device
    def run(self):
        e = self.eng
        S, H, W, Cn = e.out.shape
        loss_ptr = e._aptr(NLL)
        if self.task in ("den", "sr"):
            L.call("mfvi_gauss_nll_fwd_bwd", 0, L.view(e.out), S, H, W, Cn, self.sub, self.target.data_ptr(), None,
                   loss_ptr, L.view(e.dout), meta={"bytes": 4.0 * (2 * e.out.numel() + self.target.numel())})
        elif self.task == "inp":
            L.call("mfvi_gauss_nll_fwd_bwd", 1, L.view(e.out), S, H, W, Cn, 1, self.target.data_ptr(),
                   self.mask.data_ptr(), loss_ptr, L.view(e.dout),
                   meta={"bytes": 4.0 * (2 * e.out.numel() + self.target.numel() + self.mask.numel())})
        else:
            n = Cn * self.T * W
            rb = {"bytes": 4.0 * (e.out.numel() + self.sino.numel()), "samples": float(S * Cn * self.T * H * W)}
            L.call("mfvi_radon_fwd", L.view(e.out), S, Cn, H, W, self.theta.data_ptr(), self.T, self.sino.data_ptr(), meta=rb)
            L.call("mfvi_mse_fwd_bwd", self.sino.data_ptr(), n, self.sino_t.data_ptr(), n, S, loss_ptr,
                   self.dsino.data_ptr(), meta={"bytes": 4.0 * (2 * self.sino.numel() + n)})
            L.call("mfvi_radon_bwd", self.dsino.data_ptr(), S, Cn, H, W, self.theta.data_ptr(), self.T, L.view(e.dout), meta=rb)


class MfviDipTrainer:
    def __init__(self, spec: SkipSpec, task: str, net_input: torch.Tensor, *, temp: float, sigma: float, lr: float,
                 mc_samples: int = 1, seed: int = 0, reg_noise_std: float = 0.1, device="cuda", math_mode: int = L.MATH_FP32,
                 target=None, mask=None, theta_deg=None, sino=None, sr_factor: int = 4,
                 rank: int = 0, world_size: int = 1, process_group=None, use_graph: bool = True,
                 nan_guard: Optional[bool] = None, betas=(0.9, 0.999), adam_eps: float = 1e-8, weight_decay: float = 0.0,
                 kl_type: str = "reverse", prior_mu: float = 0.0, plan_only: bool = False):
        device = torch.device(device)
        assert net_input.dim() == 4 and net_input.shape[0] == 1, "net_input must be (1,C,H,W) like the reference's"
        _, Cin, H, W = net_input.shape
        assert Cin == spec.num_input_channels
        self.S_global = mc_samples
        self.S, self.sample0 = shard_samples(mc_samples, rank, world_size)
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        self.device = device
        self.temp, self.sigma, self.lr = float(temp), float(sigma), float(lr)
        self.prior_mu = float(prior_mu)
        # prior scale: sqrt(temp)*sigma handed to VIModule, which adds 1e-6 (bayesian_optimization.py:1335-1336, module.py:38)
        self.prior_sigma_plus_eps = math.sqrt(self.temp) * self.sigma + 1e-6
        self.direction = 0 if kl_type == "reverse" else 1
        self.betas, self.adam_eps, self.weight_decay = betas, adam_eps, weight_decay
        self.reg_noise_std = float(reg_noise_std)
        self.seed = int(seed)
        # plan_only: build every buffer and the kernel plan on any device without being able to launch it (the test-suite
        # interprets such a trainer on the CPU: tests/plan_interpreter.py)
        self.eng = SkipEngine(spec, H, W, self.S, device, math=math_mode, plan_only=plan_only)
        self.head = LossHead(self.eng, task, target=target, mask=mask, theta_deg=theta_deg, sino=sino, sr_factor=sr_factor)
        self.nan_guard = (task == "ct") if nan_guard is None else nan_guard   # CT runner skips the update on NaN loss
        e = self.eng
        self.saved = net_input[0].to(device, torch.float32).permute(1, 2, 0).contiguous()   # NHWC (H,W,C)
        self.noise = None                       # injected jitter normals (tests)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=device)     # device-side step counter (Philox key, bookkeeping)
        # NaN guard (CT runner): AdamW's own step count, which — like torch's — does not advance on a skipped update
        self.adam_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.m = torch.zeros_like(e.theta)
        self.v = torch.zeros_like(e.theta)
        self.losses = torch.zeros(2, dtype=torch.float64, device=device)
        self.init_parameters(self.seed)
        self.post_step_hooks = []               # callables enqueued after AdamW, before the step counter advances
        self.use_graph = use_graph and not plan_only
        self._graph = None
        self._warm = 0

    # ------------------------------------------------------------------
    @L.on_device
    def init_parameters(self, seed: int, mu=(0.0, 0.1), rho=(-3.0, 0.1)):
        """VIModule.reset_parameters (module.py:56-62): mu ~ N(0,0.1), rho ~ N(-3,0.1); BN gamma=1, beta=0.
        Same seed on every rank => identical replicas without a broadcast."""
        e = self.eng
        g = torch.Generator(device=e.device).manual_seed(seed)
        e.theta.zero_()
        e.mu.normal_(mu[0], mu[1], generator=g)
        e.rho.normal_(rho[0], rho[1], generator=g)
        e.gamma.fill_(1.0)
        e.running_mean.zero_()
        e.running_var.fill_(1.0)
        self.m.zero_()
        self.v.zero_()
        self.step_dev.zero_()
        self.adam_dev.zero_()

    def _keys(self):
        kw = L.key(self.seed, 0, self.sample0, self.step_dev)
        kj = L.key(self.seed, 0, 0, self.step_dev)
        return kw, kj

    @L.on_device
    def _step_eager(self):
        e = self.eng
        kw, kj = self._keys()
        e.zero_accumulators()
        e.set_input(self.saved, self.noise, self.reg_noise_std, kj)
        e.sample_weights(kw)
        e.forward()
        self.head.run()
        e.backward()
        e.reparam_kl(kw, prior_mu=self.prior_mu, prior_sigma_plus_eps=self.prior_sigma_plus_eps, direction=self.direction,
                     kscale=self.temp)
        # NaN guard (reference :577-582 tests nll + temp*kl): this rank's loss goes into the tail slot of the gradient buffer,
        # so after the all-reduce every rank tests the SAME value (NaN / Inf on any rank poisons the mean) and either all
        # ranks apply the update or none does
        flag = e.grad.data_ptr() + 4 * e.n_theta_pad
        if self.nan_guard:
            L.call("mfvi_loss_flag", e._aptr(KL), e._aptr(NLL), self.temp, flag)
        if self.world_size > 1:
            allreduce_mean_(e.grad_buf if self.nan_guard else e.grad, self.pg)
        e.update_running_stats()
        L.call("mfvi_adamw_step", e.theta.data_ptr(), e.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
               e.n_theta_pad, self.lr, self.betas[0], self.betas[1], self.adam_eps, self.weight_decay, 1,
               (self.adam_dev if self.nan_guard else self.step_dev).data_ptr(), flag if self.nan_guard else None,
               meta={"bytes": 28.0 * e.n_theta_pad})
        if self.nan_guard:
            L.call("mfvi_counter_add_if_finite", self.adam_dev.data_ptr(), 1, flag)
        for hook in self.post_step_hooks:      # e.g. runners.DeviceBookkeeping.record (captured into the graph)
            hook()
        L.call("mfvi_counter_add", self.step_dev.data_ptr(), 1)

    @L.on_device
    def step(self):
        """One optimiser step, asynchronous on the current stream."""
        if not self.use_graph:
            return self._step_eager()
        if self._graph is None:
            if self._warm < 2:           # let lazy initialisation (func attributes, NCCL) happen outside the capture
                self._warm += 1
                return self._step_eager()
            torch.cuda.synchronize()
            before = L.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager()
            self.launches_per_step = L.launch_count - before
            self._graph = g
            # the capture itself did not execute: fall through and replay once
        self._graph.replay()

    # ------------------------------------------------------------------ host-buffer entry (the e2e path of bench.py)
    def _target_tensor(self):
        tgt = getattr(self.head, "target", None)
        return self.head.sino_t if tgt is None else tgt

    @L.on_device
    def host_buffers(self):
        """Pinned host staging buffers: (net_input NHWC (H,W,C), target as the head stores it, result[2] doubles)."""
        tgt = self._target_tensor()
        return (torch.empty_like(self.saved, device="cpu").pin_memory(), torch.empty_like(tgt, device="cpu").pin_memory(),
                torch.zeros(2, dtype=torch.float64).pin_memory())

    @L.on_device
    def step_from_host(self, net_input_host, target_host, result_host):
        """One step with HOST inputs: H2D copy of the net input and the target, the step, D2H of [kl, nll].
        Asynchronous and pipelined one step deep: the H2D copies go to one of two device staging slots on a copy stream
        (so they overlap the previous step's kernels), the compute stream moves the slot into the step's fixed input
        buffers (device-to-device) before replaying the graph, and the D2H of the losses follows the graph.  Returns
        (h2d_bytes, d2h_bytes, done_event); `done_event.synchronize()` before reading result_host or reusing the slot's
        host buffers."""
        tgt = self._target_tensor()
        if not hasattr(self, "_stage"):
            self._copy_stream = torch.cuda.Stream(device=self.eng.device)
            self._stage = [(torch.empty_like(self.saved), torch.empty_like(tgt)) for _ in range(2)]
            self._ev_h2d = [torch.cuda.Event() for _ in range(2)]
            self._ev_used = [None, None]
            self._slot = 0
        k = self._slot
        self._slot ^= 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            if self._ev_used[k] is not None:
                self._copy_stream.wait_event(self._ev_used[k])     # the step that last read this slot has consumed it
            self._stage[k][0].copy_(net_input_host, non_blocking=True)
            self._stage[k][1].copy_(target_host, non_blocking=True)
            self._ev_h2d[k].record(self._copy_stream)
        main.wait_event(self._ev_h2d[k])
        self.saved.copy_(self._stage[k][0], non_blocking=True)
        tgt.copy_(self._stage[k][1], non_blocking=True)
        used = torch.cuda.Event()
        used.record(main)
        self._ev_used[k] = used
        self.step()
        result_host.copy_(self.eng.arena[:2], non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        return self.saved.numel() * 4 + tgt.numel() * 4, 16, done

    # ------------------------------------------------------------------
    @L.on_device
    def loss_terms(self):
        """(nll, kl, loss) of the last step as Python floats (synchronises)."""
        a = self.eng.arena[:2].cpu()
        kl, nll = float(a[KL]), float(a[NLL])
        return nll, kl, nll + self.temp * kl

    @property
    def steps_done(self) -> int:
        return int(self.step_dev.item())
