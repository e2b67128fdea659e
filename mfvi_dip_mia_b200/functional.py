"""torch.autograd.Function wrappers over the libmfvidip C ABI for the stand-alone (module-by-module) use of the
BayTorch surface: one sampled-weight conv / linear layer, the Gaussian-posterior KL, the heteroscedastic NLL and
the radon projector.  The whole-network fast path lives in engine.py / trainer.py; these are the same kernels
behind the reference's per-module call signatures.  Tensors must live on a CUDA device — there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_rng_calls = 0


def philox_normal(shape, device, seed: int, sample: int = 0, step: int = 0, stream_id: int = L.STREAM_WEIGHTS):
    """Standard normals from the library's Philox4x32-10 + Box-Muller stream (replaces torch.randn_like in
    VIModule.rsample, reference BayTorch/modules/module.py:82-85)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    L.require_cuda(out, "philox_normal")
    if out.numel():
        L.call("mfvi_philox_normal_fill", out.data_ptr(), out.numel(), L.key(seed, step, sample), stream_id)
    return out


def fresh_eps_like(t: torch.Tensor):
    """eps for a stand-alone layer call: a new Philox stream per call, seeded from torch's global seed so that
    torch.manual_seed() makes runs reproducible."""
    global _rng_calls
    _rng_calls += 1
    return philox_normal(t.shape, t.device, torch.initial_seed(), sample=0, step=_rng_calls & 0xFFFFFFFF,
                         stream_id=2 + (_rng_calls >> 32))


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    N, Cn, H, W = x.shape
    x = x.contiguous()
    y = torch.empty(N, H, W, Cn, dtype=torch.float32, device=x.device)
    L.call("mfvi_nchw_to_nhwc", x.data_ptr(), y.data_ptr(), N, Cn, H, W)
    return y


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    N, H, W, Cn = x.shape
    x = x.contiguous()
    y = torch.empty(N, Cn, H, W, dtype=torch.float32, device=x.device)
    L.call("mfvi_nhwc_to_nchw", x.data_ptr(), y.data_ptr(), N, Cn, H, W)
    return y


def sample_flat(mu, rho, eps):
    """w = mu + softplus(rho) * eps, elementwise over contiguous tensors of one shape."""
    mu, rho, eps = mu.contiguous(), rho.contiguous(), eps.contiguous()
    w = torch.empty_like(mu)
    n = mu.numel()
    L.call("mfvi_sample_weights", mu.data_ptr(), rho.data_ptr(), n, 1, eps.data_ptr(), n, L.key(0), w.data_ptr(), n)
    return w


def reparam_grads(mu, rho, eps, dw):
    """(dL/dmu, dL/drho) from dL/dw through w = mu + softplus(rho)*eps:  dmu = dw, drho = dw*eps*sigmoid(rho)."""
    mu, rho, eps, dw = mu.contiguous(), rho.contiguous(), eps.contiguous(), dw.contiguous()
    gm, gr = torch.empty_like(mu), torch.empty_like(mu)
    n = mu.numel()
    L.call("mfvi_kl_reparam_fwd_bwd", mu.data_ptr(), rho.data_ptr(), n, 0.0, 1.0, 0, 0.0, None, dw.data_ptr(), n, 1,
           eps.data_ptr(), n, L.key(0), 1.0, None, gm.data_ptr(), gr.data_ptr(), 0)
    return gm, gr


@L.device_guarded
class SampledConv2dFn(torch.autograd.Function):
    """RTLayer.forward with layer_fn=conv2d (reference reparam_layers.py:26-37, conv.py:6-38): one weight sample
    shared by the whole batch.  x (N,Cin,H,W); W_* (Cout,Cin,KH,KW); zero `padding`; dilation=groups=1."""

    @staticmethod
    def forward(ctx, x, W_mu, W_rho, b_mu, b_rho, eps_w, eps_b, stride, padding, training, math):
        L.require_cuda(x, "Conv2dRT")
        N, Cin, H, W = x.shape
        Cout, _, KH, KW = W_mu.shape
        has_bias = b_mu is not None
        if training:
            w = sample_flat(W_mu, W_rho, eps_w)
            b = sample_flat(b_mu, b_rho, eps_b) if has_bias else None
        else:
            w, b = W_mu.detach().contiguous(), (b_mu.detach().contiguous() if has_bias else None)
        wt = w.permute(2, 3, 0, 1).contiguous()            # tap-major storage [KH][KW][Cout][Cin]
        xh = to_nhwc(x.detach().float())
        if padding:
            xp = torch.zeros(N, H + 2 * padding, W + 2 * padding, Cin, dtype=torch.float32, device=x.device)
            xp[:, padding:padding + H, padding:padding + W] = xh
            xh = xp
        Hin, Win = xh.shape[1], xh.shape[2]
        Ho, Wo = (Hin - KH) // stride + 1, (Win - KW) // stride + 1
        d = L.ConvDesc(N, Cin, Cout, KH, KW, stride, Hin, Win, Ho, Wo, math)
        y = torch.empty(N, Ho, Wo, Cout, dtype=torch.float32, device=x.device)
        vx, vy = L.view(xh), L.view(y)
        if N == 1:
            vx.sstride = 0
        L.call("mfvi_conv2d_fwd", C.byref(d), vx, wt.data_ptr(), L.ptr(b), 0, vy, None)
        ctx.save_for_backward(xh, wt, W_mu, W_rho, b_mu, b_rho, eps_w, eps_b)
        ctx.geom = (d, padding, training, (N, Cin, H, W))
        return to_nchw(y)

    @staticmethod
    def backward(ctx, dy):
        xh, wt, W_mu, W_rho, b_mu, b_rho, eps_w, eps_b = ctx.saved_tensors
        d, padding, training, (N, Cin, H, W) = ctx.geom
        dyh = to_nhwc(dy.float())
        dwt = torch.zeros_like(wt)
        has_bias = b_mu is not None
        db = torch.zeros(d.Cout, dtype=torch.float32, device=dy.device) if has_bias else None
        L.call("mfvi_conv2d_wgrad", C.byref(d), L.view(xh), L.view(dyh), dwt.data_ptr(), L.ptr(db), 0)
        dx = None
        if ctx.needs_input_grad[0]:
            dxh = torch.empty_like(xh)
            L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dyh), wt.data_ptr(), 0, L.view(dxh), 0)
            if padding:
                dxh = dxh[:, padding:padding + H, padding:padding + W].contiguous()
            dx = to_nchw(dxh)
        dw = dwt.permute(2, 3, 0, 1).contiguous()          # back to (Cout,Cin,KH,KW)
        if training:
            gWm, gWr = reparam_grads(W_mu, W_rho, eps_w, dw)
            gbm, gbr = reparam_grads(b_mu, b_rho, eps_b, db) if has_bias else (None, None)
        else:
            gWm, gWr, gbm, gbr = dw, None, db, None
        return dx, gWm, gWr, gbm, gbr, None, None, None, None, None, None


def _pad_nhwc(xh, padding):
    if not padding:
        return xh
    N, H, W, Cin = xh.shape
    xp = torch.zeros(N, H + 2 * padding, W + 2 * padding, Cin, dtype=torch.float32, device=xh.device)
    xp[:, padding:padding + H, padding:padding + W] = xh
    return xp


@L.device_guarded
class LrtConv2dFn(torch.autograd.Function):
    """LRTLayer.forward with layer_fn=conv2d (reference reparam_layers.py:39-72, conv.py:75-107):
        act_mu = conv(x, W_mu, b_mu);  act_var = conv(x^2, softplus(W_rho)^2, softplus(b_rho)^2)
        out = act_mu + sqrt(1e-16 + act_var) * eps            (training)      out = act_mu   (eval)
    with eps of the OUTPUT's shape.  Both convolutions and every elementwise piece run on libmfvidip kernels."""

    @staticmethod
    def forward(ctx, x, W_mu, W_rho, b_mu, b_rho, eps, stride, padding, training, math):
        L.require_cuda(x, "Conv2dLRT")
        N, Cin, H, W = x.shape
        Cout, _, KH, KW = W_mu.shape
        has_bias = b_mu is not None
        dev = x.device
        xh = _pad_nhwc(to_nhwc(x.detach().float()), padding)
        Hin, Win = xh.shape[1], xh.shape[2]
        Ho, Wo = (Hin - KH) // stride + 1, (Win - KW) // stride + 1
        d = L.ConvDesc(N, Cin, Cout, KH, KW, stride, Hin, Win, Ho, Wo, math)
        wt = W_mu.detach().permute(2, 3, 0, 1).contiguous()           # tap-major [KH][KW][Cout][Cin]
        bm = b_mu.detach().contiguous() if has_bias else None
        act_mu = torch.empty(N, Ho, Wo, Cout, dtype=torch.float32, device=dev)
        L.call("mfvi_conv2d_fwd", C.byref(d), L.view(xh), wt.data_ptr(), L.ptr(bm), 0, L.view(act_mu), None)
        saved = [xh, wt, W_rho, b_rho]
        if training:
            rt = W_rho.detach().permute(2, 3, 0, 1).contiguous()
            s2 = torch.empty_like(rt)
            L.call("mfvi_softplus_sq_fwd", rt.data_ptr(), rt.numel(), s2.data_ptr())
            b2 = None
            if has_bias:
                b2 = torch.empty_like(bm)
                L.call("mfvi_softplus_sq_fwd", b_rho.detach().contiguous().data_ptr(), b2.numel(), b2.data_ptr())
            x2 = torch.empty_like(xh)
            L.call("mfvi_square_fwd", xh.data_ptr(), xh.numel(), x2.data_ptr())
            var = torch.empty_like(act_mu)
            L.call("mfvi_conv2d_fwd", C.byref(d), L.view(x2), s2.data_ptr(), L.ptr(b2), 0, L.view(var), None)
            eh = to_nhwc(eps.to(dev, torch.float32))
            out = torch.empty_like(act_mu)
            L.call("mfvi_lrt_noise_fwd", act_mu.data_ptr(), var.data_ptr(), eh.data_ptr(), out.numel(), out.data_ptr())
            saved += [rt, s2, x2, var, eh]
        else:
            out = act_mu
        ctx.save_for_backward(*saved)
        ctx.geom = (d, padding, training, has_bias, (N, Cin, H, W))
        return to_nchw(out)

    @staticmethod
    def backward(ctx, dy):
        d, padding, training, has_bias, (N, Cin, H, W) = ctx.geom
        xh, wt, W_rho, b_rho = ctx.saved_tensors[:4]
        dev = dy.device
        dyh = to_nhwc(dy.float())
        dwt = torch.zeros_like(wt)
        db = torch.zeros(d.Cout, dtype=torch.float32, device=dev) if has_bias else None
        L.call("mfvi_conv2d_wgrad", C.byref(d), L.view(xh), L.view(dyh), dwt.data_ptr(), L.ptr(db), 0)
        need_dx = ctx.needs_input_grad[0]
        dxh = None
        if need_dx:
            dxh = torch.empty_like(xh)
            L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dyh), wt.data_ptr(), 0, L.view(dxh), 0)
        gWr = gbr = None
        if training:
            rt, s2, x2, var, eh = ctx.saved_tensors[4:]
            dvar = torch.empty_like(var)
            L.call("mfvi_lrt_noise_bwd", dyh.data_ptr(), var.data_ptr(), eh.data_ptr(), dvar.numel(), dvar.data_ptr())
            ds2 = torch.zeros_like(s2)
            db2 = torch.zeros(d.Cout, dtype=torch.float32, device=dev) if has_bias else None
            L.call("mfvi_conv2d_wgrad", C.byref(d), L.view(x2), L.view(dvar), ds2.data_ptr(), L.ptr(db2), 0)
            grt = torch.empty_like(rt)
            L.call("mfvi_softplus_sq_bwd", rt.data_ptr(), ds2.data_ptr(), rt.numel(), grt.data_ptr(), 0)
            gWr = grt.permute(2, 3, 0, 1).contiguous()
            if has_bias:
                gbr = torch.empty_like(db2)
                L.call("mfvi_softplus_sq_bwd", b_rho.detach().contiguous().data_ptr(), db2.data_ptr(), db2.numel(),
                       gbr.data_ptr(), 0)
            if need_dx:
                dx2 = torch.empty_like(xh)
                L.call("mfvi_conv2d_dgrad", C.byref(d), L.view(dvar), s2.data_ptr(), 0, L.view(dx2), 0)
                L.call("mfvi_square_bwd", xh.data_ptr(), dx2.data_ptr(), xh.numel(), dxh.data_ptr(), 1)
        dx = None
        if need_dx:
            if padding:
                dxh = dxh[:, padding:padding + H, padding:padding + W].contiguous()
            dx = to_nchw(dxh)
        gWm = dwt.permute(2, 3, 0, 1).contiguous()
        return dx, gWm, gWr, db, gbr, None, None, None, None, None


@L.device_guarded
class KlFn(torch.autograd.Function):
    """VIModule._kl (reference module.py:64-80) for one (mu, rho) pair: sum of closed-form Gaussian KLs,
    returned as a 0-dim fp32 tensor."""

    @staticmethod
    def forward(ctx, mu, rho, prior_mu, prior_sigma_plus_eps, direction):
        L.require_cuda(mu, "VIModule._kl")
        m, r = mu.detach().contiguous(), rho.detach().contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=mu.device)
        L.call("mfvi_kl_reparam_fwd_bwd", m.data_ptr(), r.data_ptr(), m.numel(), float(prior_mu),
               float(prior_sigma_plus_eps), direction, 0.0, None, None, 0, 0, None, 0, L.key(0), 0.0, acc.data_ptr(),
               None, None, 0)
        ctx.save_for_backward(m, r)
        ctx.cfg = (float(prior_mu), float(prior_sigma_plus_eps), direction)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        m, r = ctx.saved_tensors
        pm, ps, direction = ctx.cfg
        gm, gr = torch.empty_like(m), torch.empty_like(r)
        gs = g.detach().to(torch.float32).reshape(1).contiguous()
        L.call("mfvi_kl_reparam_fwd_bwd", m.data_ptr(), r.data_ptr(), m.numel(), pm, ps, direction, 1.0, gs.data_ptr(),
               None, 0, 0, None, 0, L.key(0), 0.0, None, gm.data_ptr(), gr.data_ptr(), 0)
        return gm, gr, None, None, None


@L.device_guarded
class GaussianNllFn(torch.autograd.Function):
    """utils/bayesian_utils.py:29-39 of the reference.  mode 0: mu, s (N,1,H,W); mode 1 (inpainting): `mu` holds the
    PRE-sigmoid colour channels (N,3,H,W), s (N,1,H,W), mask (1,1,H,W)."""

    @staticmethod
    def forward(ctx, mu, s, target, mask, mode, reduction):
        L.require_cuda(mu, "gaussian_nll")
        N, Cm, H, W = mu.shape
        out = torch.cat([mu.detach().float(), s.detach().float()], dim=1).permute(0, 2, 3, 1).contiguous()
        dout = torch.empty_like(out)
        acc = torch.zeros(1, dtype=torch.float64, device=mu.device)
        if mode == 0:
            t = target.to(mu.device, torch.float32).reshape(-1).contiguous()
            assert Cm == 1 and t.numel() == H * W, "gaussian_nll: target must broadcast as one (H,W) image"
            L.call("mfvi_gauss_nll_fwd_bwd", 0, L.view(out), N, H, W, 2, 1, t.data_ptr(), None, acc.data_ptr(), L.view(dout))
            count = N * H * W
        else:
            t = target.to(mu.device, torch.float32).reshape(3, H, W).permute(1, 2, 0).contiguous()
            m = mask.to(mu.device, torch.float32).reshape(H, W).contiguous()
            L.call("mfvi_gauss_nll_fwd_bwd", mode, L.view(out), N, H, W, 4, 1, t.data_ptr(), m.data_ptr(), acc.data_ptr(),
                   L.view(dout))
            count = N * H * W * 3
        scale = 1.0 if reduction == "mean" else float(count)
        ctx.save_for_backward(dout)
        ctx.cfg = (Cm, scale)
        return (acc * scale).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        Cm, scale = ctx.cfg
        d = (dout * (g.to(torch.float32) * scale)).permute(0, 3, 1, 2)
        return d[:, :Cm].contiguous(), d[:, Cm:].contiguous(), None, None, None, None


@L.device_guarded
class RadonFn(torch.autograd.Function):
    """FastRadonTransform.forward (reference radon/radon.py:48-55): image (1,C,H,W) -> sinogram (1,C,T,W)."""

    @staticmethod
    def forward(ctx, image, theta_rad):
        L.require_cuda(image, "FastRadonTransform")
        B, Cn, H, W = image.shape
        if B != 1:
            raise L.MfviError("FastRadonTransform: batch must be 1 (the reference expands the batch axis to the angles)")
        img = image.detach().float().contiguous()
        T = theta_rad.numel()
        sino = torch.empty(1, Cn, T, W, dtype=torch.float32, device=image.device)
        # NCHW with one "channel" per plane: sample = c, channel count 1
        v = L.View(img.data_ptr(), H * W, W, 1)
        L.call("mfvi_radon_fwd", v, Cn, 1, H, W, theta_rad.data_ptr(), T, sino.data_ptr())
        ctx.save_for_backward(theta_rad)
        ctx.geom = (Cn, H, W, T)
        return sino

    @staticmethod
    def backward(ctx, dsino):
        (theta_rad,) = ctx.saved_tensors
        Cn, H, W, T = ctx.geom
        ds = dsino.float().contiguous()
        dimg = torch.empty(1, Cn, H, W, dtype=torch.float32, device=dsino.device)
        v = L.View(dimg.data_ptr(), H * W, W, 1)
        L.call("mfvi_radon_bwd", ds.data_ptr(), Cn, 1, H, W, theta_rad.data_ptr(), T, v)
        return dimg, None
