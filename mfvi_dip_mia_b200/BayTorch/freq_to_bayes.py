"""MeanFieldVI (reference BayTorch/freq_to_bayes.py:7-89): converts the Conv2d/Linear leaves of a network into
mean-field Gaussian layers and exposes `.forward(x)` and `.kl()`.

When the wrapped network is a skip() hour-glass whose every convolution was converted (replace_layers='all',
reparam='' — what all MFVI runners of the reference do, bayesian_optimization.py:1338-1342) and it lives on a
CUDA device, forward/backward/kl run on the fused sm_100a engine:
  * all parameters are re-homed into one flat buffer (module Parameters become views, so `net.parameters()`,
    `state_dict()` and any torch optimiser keep working),
  * `forward(x)` evaluates `mc_samples` weight samples at once and returns (mc_samples, C, H, W) — for the
    reference's MC=1 that is the familiar (1, C, H, W),
  * gradients of ALL parameters are produced by the engine during `loss.backward()` and appear in `p.grad`.
Any other network falls back to module-by-module execution where each Conv2dRT/LinearRT still runs the
library's sampled-weight kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from ..engine import SkipEngine
from .modules import Conv2dLRT, Conv2dRT, LinearLRT, LinearRT, VIModule


@L.device_guarded
class _FusedForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, owner):
        eng = owner._engine_for(x)
        owner._generation += 1
        ctx.owner, ctx.eng, ctx.generation = owner, eng, owner._generation
        xh = x.detach().to(torch.float32).contiguous()
        nhwc = torch.empty(x.shape[2], x.shape[3], x.shape[1], dtype=torch.float32, device=x.device)
        L.call("mfvi_nchw_to_nhwc", xh.data_ptr(), nhwc.data_ptr(), 1, x.shape[1], x.shape[2], x.shape[3])
        eng.zero_accumulators()
        eng.set_input(nhwc, None, 0.0, L.key(0))
        ctx.key = L.key(owner.seed, owner._forward_calls, owner.sample0)
        owner._forward_calls += 1
        eng.sample_weights(ctx.key)
        eng.forward()
        eng.update_running_stats()
        return eng.out_nchw()

    @staticmethod
    def backward(ctx, dout):
        owner, eng = ctx.owner, ctx.eng
        if ctx.generation != owner._generation:
            raise L.MfviError("MeanFieldVI: backward through a stale forward — the fused engine keeps the activations "
                              "of the latest forward only")
        d = dout.to(torch.float32).contiguous()
        S, Cn, H, W = d.shape
        dn = torch.empty(S, H, W, Cn, dtype=torch.float32, device=d.device)
        L.call("mfvi_nchw_to_nhwc", d.data_ptr(), dn.data_ptr(), S, Cn, H, W)
        eng.dout.copy_(dn)                      # dout keeps a 4-float channel pitch
        owner._attach_grads()
        Pp, Q = eng.lay.P_pad, eng.lay.Q
        bn_before = eng.grad[2 * Pp:2 * Pp + 2 * Q].clone()
        eng.backward()                        # writes BN grads, accumulates per-sample dw
        eng.grad[2 * Pp:2 * Pp + 2 * Q] += bn_before
        eng.reparam_kl(ctx.key, prior_mu=0.0, prior_sigma_plus_eps=1.0, direction=0, kscale=0.0, data_term=True,
                       accumulate=True, want_kl=False)
        return None, torch.zeros_like(owner._anchor), None


@L.device_guarded
class _FusedKl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, owner):
        eng = owner._engine
        ctx.owner = owner
        acc = torch.zeros(1, dtype=torch.float64, device=eng.device)
        L.call("mfvi_kl_reparam_fwd_bwd", eng.mu.data_ptr(), eng.rho.data_ptr(), eng.lay.P, owner._prior_mu,
               owner._prior_scale, owner._direction, 0.0, None, None, 0, 0, None, 0, L.key(0), 0.0, acc.data_ptr(),
               None, None, 0)
        return acc.to(torch.float32)          # shape [1] like the reference's FloatTensor([0.0]) accumulator

    @staticmethod
    def backward(ctx, g):
        owner = ctx.owner
        eng = owner._engine
        owner._attach_grads()
        gs = g.detach().to(torch.float32).reshape(-1)[:1].contiguous()
        L.call("mfvi_kl_reparam_fwd_bwd", eng.mu.data_ptr(), eng.rho.data_ptr(), eng.lay.P, owner._prior_mu,
               owner._prior_scale, owner._direction, 1.0, gs.data_ptr(), None, 0, 0, None, 0, L.key(0), 0.0, None,
               eng.g_mu.data_ptr(), eng.g_rho.data_ptr(), 1)
        return torch.zeros_like(owner._anchor), None


class MeanFieldVI(nn.Module):
    def __init__(self, net, prior=None, posteriors=None, kl_type='reverse', reparam='local', replace_layers='all',
                 device=torch.device('cpu'), mc_samples: int = 1, seed=None, sample0: int = 0, math=L.MATH_FP32):
        super().__init__()
        self.net = net
        self.device = torch.device(device)
        # reparam='local' (the reference's default, freq_to_bayes.py:22-25) builds local-reparameterisation layers; they
        # run module by module on the library's kernels.  The fused whole-network engine serves reparam='' (weight-space
        # sampling), which is what every MFVI runner passes (bayesian_optimization.py:1342).
        self._local = reparam == 'local'
        self._conv2d, self._linear = (Conv2dLRT, LinearLRT) if self._local else (Conv2dRT, LinearRT)
        assert replace_layers in ['up', 'down', 'all', 'none']
        self._replace_layers = '' if replace_layers == 'all' else replace_layers
        self.mc_samples, self.sample0, self.math = int(mc_samples), int(sample0), math
        self.seed = int(torch.initial_seed() if seed is None else seed)
        n_conv_before = sum(isinstance(m, (nn.Conv2d, nn.Linear)) for m in net.modules())
        self._replace_deterministic_modules(self.net, prior, posteriors, kl_type)
        self.net = net.to(self.device)
        n_left = sum(isinstance(m, (nn.Conv2d, nn.Linear)) for m in net.modules())
        vi = [m for m in net.modules() if isinstance(m, VIModule)]
        self._spec = getattr(net, "_skip_spec", None) if (n_left == 0 and n_conv_before > 0 and not self._local) else None
        if vi:
            self._prior_mu, self._prior_scale = vi[0].prior_loc, vi[0].prior_scale
            self._direction = 0 if vi[0].kl_type == 'reverse' else 1
        self._engine = None
        self._generation = 0
        self._forward_calls = 0
        self._anchor = None

    # ------------------------------------------------------------------ conversion (freq_to_bayes.py:50-89)
    def _replace_deterministic_modules(self, module, prior, posteriors, kl_type):
        for key, child in module._modules.items():
            if len(child._modules):
                self._replace_deterministic_modules(child, prior, posteriors, kl_type)
            elif self._replace_layers in key:
                if isinstance(child, nn.Linear):
                    # like the reference (freq_to_bayes.py:56-61): Linear layers keep the layer's DEFAULT prior / posterior
                    # initialisation / kl_type — only the convolutions receive the ones passed in
                    module._modules[key] = self._linear(child.in_features, child.out_features, torch.is_tensor(child.bias))
                elif isinstance(child, nn.Conv2d):
                    module._modules[key] = self._conv2d(
                        in_channels=child.in_channels, out_channels=child.out_channels, kernel_size=child.kernel_size,
                        bias=torch.is_tensor(child.bias), stride=child.stride, padding=child.padding,
                        dilation=child.dilation, groups=child.groups, prior=prior, posteriors=posteriors,
                        kl_type=kl_type)
                elif isinstance(child, nn.Conv3d):
                    raise NotImplementedError("Conv3d layers are outside the MFVI-DIP hot path (no runner builds them)")

    # ------------------------------------------------------------------ fused engine plumbing
    @property
    def fused(self) -> bool:
        return self._spec is not None and self.device.type == "cuda"

    def _engine_for(self, x) -> SkipEngine:
        H, W = x.shape[2], x.shape[3]
        if x.shape[0] != 1:
            raise L.MfviError("MeanFieldVI (fused): the input batch must be 1, as in every reference runner; MC samples "
                              "are requested with mc_samples=")
        eng = self._engine
        if eng is not None and (eng.H, eng.W) == (H, W):
            return eng
        new = SkipEngine(self._spec, H, W, self.mc_samples, x.device, math=self.math)
        if eng is not None:                       # keep parameters when the image size changes
            new.theta.copy_(eng.theta)
            new.running_mean.copy_(eng.running_mean)
            new.running_var.copy_(eng.running_var)
        else:
            sd = {k: v for k, v in self.net.state_dict().items()}
            new.load_params(sd)
        self._engine = new
        self._rehome_parameters()
        return new

    def _rehome_parameters(self):
        """Point every module Parameter / BN buffer at its view of the engine's flat storage."""
        eng = self._engine
        views = eng.param_views("theta")
        named = dict(self.net.named_parameters())
        named.update(dict(self.net.named_buffers()))
        for k, v in views.items():
            named[k].data = v
        self._grad_views = {k: v for k, v in eng.param_views("grad").items()}
        self._params = {k: p for k, p in self.net.named_parameters()}
        if self._anchor is None:
            self._anchor = torch.zeros(1, device=eng.device, requires_grad=True)

    def _attach_grads(self):
        """Make p.grad a view of the flat gradient buffer; a missing grad (after zero_grad(set_to_none=True))
        means the flat buffer starts from zero."""
        first = next(iter(self._params.values()))
        if first.grad is not None and first.grad.data_ptr() == self._grad_views[next(iter(self._params))].data_ptr():
            return
        self._engine.grad.zero_()
        for k, p in self._params.items():
            p.grad = self._grad_views[k]

    def prepare(self, x):
        """Build the fused engine for inputs shaped like `x` ahead of the first forward (so that an optimiser
        created afterwards sees the final parameter storage). Optional."""
        if self.fused:
            self._engine_for(x)
        return self

    def inject_eps(self, eps_per_sample, prefix=""):
        """Parity-test hook: use these eps ('<convkey>.W' / '<convkey>.b' per MC sample) in every following forward
        instead of the Philox stream."""
        assert self._engine is not None, "call prepare(x) first"
        self._engine.pack_eps(eps_per_sample, prefix)

    # ------------------------------------------------------------------ reference surface
    def forward(self, x):
        if self._spec is not None and x.device.type != "cuda":
            raise L.MfviError(f"MeanFieldVI: input on {x.device}; this implementation runs on CUDA only")
        if not self.fused or not self.training:
            # eval mode (never used by a reference runner): module by module, so that nn.BatchNorm2d normalises with its
            # RUNNING statistics and RTLayer uses w = mu exactly as the reference's net.eval() does (reparam_layers.py:33-35);
            # the modules' parameters and buffers are views of the engine's flat storage, so they are current
            return self.net(x)
        L.require_cuda(x, "MeanFieldVI.forward")
        self._engine_for(x)
        return _FusedForward.apply(x, self._anchor, self)

    def kl(self):
        if self.fused and self._engine is not None:
            return _FusedKl.apply(self._anchor, self)
        kl = torch.zeros(1, dtype=torch.float32, device=self.device)
        for layer in self.modules():
            if isinstance(layer, VIModule):
                kl = kl + layer._kl
        return kl
