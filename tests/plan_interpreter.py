"""CPU interpretation of a SkipEngine kernel plan — TEST INFRASTRUCTURE, never imported by the package.

`SkipEngine(..., device="cpu", plan_only=True)` builds the real plan (buffers, views, the forward / backward op lists with
their raw pointers) but cannot execute it.  This module executes such a plan op by op with plain PyTorch on the CPU, each op
restated from its contract in include/mfvi_dip.h, resolving every raw pointer back to the engine's buffers.  It checks the
HOST side of the engine — which buffer feeds which kernel, shapes, strides, paddings, accumulate flags, the bf16 operand
plumbing — without a GPU.  It says nothing about the CUDA kernels themselves (tests/test_gpu_*.py do).

bf16 ops are interpreted as the kernels define them: operands hold bf16 values, products and sums are fp32, results are
rounded to nearest-even where a kernel stores bf16.
"""
import torch
import torch.nn.functional as F

BN_EPS, SLOPE = 1e-5, 0.2


class PlanInterpreter:
    def __init__(self, eng):
        assert eng.plan_only and eng.device.type == "cpu"
        self.eng = eng
        self.stores = []
        seen = set()
        tensors = [eng.theta, eng.grad, eng.zbuf, eng.w, eng.dout, eng.running_mean, eng.running_var] + list(eng._bufs)
        if eng.w16 is not None:
            tensors.append(eng.w16)
        if eng.eps is not None:
            tensors.append(eng.eps)
        for t in tensors:
            st = t.untyped_storage()
            if st.data_ptr() not in seen:
                seen.add(st.data_ptr())
                self.stores.append((st.data_ptr(), st.data_ptr() + st.nbytes(), st))

    # ------------------------------------------------------------------ pointer resolution
    def _strided(self, ptr, size, stride, dtype):
        esz = torch.empty(0, dtype=dtype).element_size()
        for lo, hi, st in self.stores:
            if lo <= ptr < hi:
                off = ptr - lo
                assert off % esz == 0
                t = torch.empty(0, dtype=dtype).set_(st, off // esz, size, stride)
                last = sum((n - 1) * s for n, s in zip(size, stride)) if all(n > 0 for n in size) else 0
                assert ptr + (last + 1) * esz <= hi, "view reaches beyond its buffer"
                return t
        raise AssertionError(f"pointer {ptr:#x} is not inside any engine buffer")

    def view(self, v, S, H, W, Cn, dtype=torch.float32):
        """(S,H,W,Cn) tensor of an MfviView (sample stride 0 = broadcast)."""
        return self._strided(v.ptr, (S, H, W, Cn), (v.sstride, v.hstride, v.wstride, 1), dtype)

    def vec(self, ptr, n, dtype=torch.float32):
        return self._strided(ptr, (n,), (1,), dtype)

    @staticmethod
    def _dtype_of(name):
        return torch.bfloat16 if name.endswith("_bf16") else torch.float32

    # ------------------------------------------------------------------ BatchNorm helpers
    def _bn(self, sums_ptr, gamma_ptr, beta_ptr, S, Cn, count):
        """per-sample (mean, invstd, scale, shift), each (S,1,1,C)"""
        if sums_ptr is None:
            mean, invstd = torch.zeros(S, Cn, dtype=torch.float64), torch.ones(S, Cn, dtype=torch.float64)
        else:
            sums = self.vec(sums_ptr, S * Cn * 2, torch.float64).view(S, Cn, 2)
            mean = sums[..., 0] / count
            var = (sums[..., 1] / count - mean * mean).clamp_min(0.0)
            invstd = torch.rsqrt(var + BN_EPS)
        mean, invstd = mean.float(), invstd.float()
        gamma = self.vec(gamma_ptr, Cn) if gamma_ptr is not None else torch.ones(Cn)
        beta = self.vec(beta_ptr, Cn) if beta_ptr is not None else torch.zeros(Cn)
        sc = gamma * invstd
        sh = beta - mean * sc
        r = lambda t: t.reshape(S, 1, 1, Cn)
        return r(mean), r(invstd), r(sc), r(sh)

    def _add_red(self, ptr, a, b):
        """red[S][C][2] += (sum a, sum b) over the pixels"""
        S, _, _, Cn = a.shape
        red = self.vec(ptr, S * Cn * 2, torch.float64).view(S, Cn, 2)
        red[..., 0] += a.double().sum((1, 2))
        red[..., 1] += b.double().sum((1, 2))

    @staticmethod
    def _up2x(t_nhwc, mode):
        t = t_nhwc.permute(0, 3, 1, 2)
        u = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False) if mode == 0 else \
            F.interpolate(t, scale_factor=2, mode="nearest")
        return u.permute(0, 2, 3, 1)

    # ------------------------------------------------------------------ convolution helpers
    def _weights(self, name, d, args):
        """per-sample (Cout,Cin,KH,KW) fp32 weights of a conv op"""
        taps = d.KH * d.KW
        out = []
        for s in range(d.S):
            if name.endswith("_bf16"):
                wptr, cpitch, wss = args
                blk = self.vec(wptr + 2 * s * wss, taps * d.Cout * cpitch, torch.bfloat16).view(d.KH, d.KW, d.Cout, cpitch)
                blk = blk[..., :d.Cin].float()
            else:
                wptr, wss = args
                blk = self.vec(wptr + 4 * s * wss, taps * d.Cout * d.Cin).view(d.KH, d.KW, d.Cout, d.Cin)
            out.append(blk.permute(2, 3, 0, 1).contiguous())
        return out

    # ------------------------------------------------------------------ the ops
    def op_conv_fwd(self, name, args):
        if name.endswith("_bf16"):
            d, x, wptr, cpitch, wss, bptr, bss, y, stats = args
            wargs = (wptr, cpitch, wss)
        else:
            d, x, wptr, bptr, wss, y, stats = args
            wargs, bss = (wptr, wss), wss
        d = d._obj
        X = self.view(x, d.S, d.Hin, d.Win, d.Cin, self._dtype_of(name)).float()
        Y = self.view(y, d.S, d.Hout, d.Wout, d.Cout)
        for s, w in enumerate(self._weights(name, d, wargs)):
            b = self.vec(bptr + 4 * s * bss, d.Cout) if bptr is not None else None
            Y[s] = F.conv2d(X[s].permute(2, 0, 1)[None], w, b, stride=d.stride)[0].permute(1, 2, 0)
        if stats is not None:
            self._add_red(stats, Y, Y * Y)

    def op_conv_dgrad(self, name, args):
        if name.endswith("_bf16"):
            d, dy, wptr, cpitch, wss, dx, acc = args
            wargs = (wptr, cpitch, wss)
        else:
            d, dy, wptr, wss, dx, acc = args
            wargs = (wptr, wss)
        d = d._obj
        DY = self.view(dy, d.S, d.Hout, d.Wout, d.Cout, self._dtype_of(name)).float()
        DX = self.view(dx, d.S, d.Hin, d.Win, d.Cin)
        for s, w in enumerate(self._weights(name, d, wargs)):
            g = F.conv_transpose2d(DY[s].permute(2, 0, 1)[None], w, stride=d.stride)[0].permute(1, 2, 0)
            full = torch.zeros(d.Hin, d.Win, d.Cin)            # stride 2: an even-sized input has one unused trailing row / column
            full[:g.shape[0], :g.shape[1]] = g
            DX[s] = DX[s] + full if acc else full

    def op_conv_wgrad(self, name, args):
        if name.endswith("_bf16"):
            d, x, dy, dwptr, wss, dyf, dbptr = args
        else:
            d, x, dy, dwptr, dbptr, wss = args
            dyf = dy
        d = d._obj
        dt = self._dtype_of(name)
        X = self.view(x, d.S, d.Hin, d.Win, d.Cin, dt).float()
        DY = self.view(dy, d.S, d.Hout, d.Wout, d.Cout, dt).float()
        for s in range(d.S):
            xs = X[s].permute(2, 0, 1)[None]
            # the kernels read only the pixels the convolution touches
            used_h, used_w = (d.Hout - 1) * d.stride + d.KH, (d.Wout - 1) * d.stride + d.KW
            gw = torch.nn.grad.conv2d_weight(xs[..., :used_h, :used_w], (d.Cout, d.Cin, d.KH, d.KW), DY[s].permute(2, 0, 1)[None],
                                             stride=d.stride)
            self.vec(dwptr + 4 * s * wss, d.KH * d.KW * d.Cout * d.Cin).add_(gw.permute(2, 3, 0, 1).reshape(-1))
        if dbptr is not None:
            DYF = self.view(dyf, d.S, d.Hout, d.Wout, d.Cout)
            for s in range(d.S):
                self.vec(dbptr + 4 * s * wss, d.Cout).add_(DYF[s].sum((0, 1)))

    def op_bn_act_pad_fwd(self, name, args):
        y, S, H, W, Cn, sums, gamma, beta, act, pad, xp = args
        Y = self.view(y, S, H, W, Cn)
        _, _, sc, sh = self._bn(sums, gamma, beta, S, Cn, H * W)
        Z = Y * sc + sh
        if act:
            Z = torch.where(Z > 0, Z, SLOPE * Z)
        if pad:
            Z = F.pad(Z.permute(0, 3, 1, 2), (pad,) * 4, mode="reflect").permute(0, 2, 3, 1)
        self.view(xp, S, H + 2 * pad, W + 2 * pad, Cn, self._dtype_of(name)).copy_(Z)

    def op_cat_up_fwd(self, name, args):
        ys, Cs, sums_s, gamma_s, beta_s, yd, Cd, sums_d, gamma_d, beta_d, S, H, W, mode, A, sumsA = args
        parts = []
        if Cs:
            _, _, sc, sh = self._bn(sums_s, gamma_s, beta_s, S, Cs, H * W)
            z = self.view(ys, S, H, W, Cs) * sc + sh
            parts.append(torch.where(z > 0, z, SLOPE * z))
        _, _, sc, sh = self._bn(sums_d, gamma_d, beta_d, S, Cd, (H // 2) * (W // 2))
        z = self.view(yd, S, H // 2, W // 2, Cd) * sc + sh
        parts.append(self._up2x(torch.where(z > 0, z, SLOPE * z), mode))
        out = torch.cat(parts, 3)
        self.view(A, S, H, W, Cs + Cd).copy_(out)
        self._add_red(sumsA, out, out * out)

    def op_pad_act_bwd(self, name, args):
        dxp, S, H, W, Cn, pad, y, sums, gamma, beta, act, g, red = args
        DXP = self.view(dxp, S, H + 2 * pad, W + 2 * pad, Cn)
        Y = self.view(y, S, H, W, Cn)
        mean, invstd, sc, sh = self._bn(sums, gamma, beta, S, Cn, H * W)
        if pad:       # adjoint of the reflection pad
            probe = torch.zeros(S, Cn, H, W, requires_grad=True)
            F.pad(probe, (pad,) * 4, mode="reflect").backward(DXP.permute(0, 3, 1, 2).contiguous())
            G = probe.grad.permute(0, 2, 3, 1)
        else:
            G = DXP.clone()
        if act:
            G = torch.where(Y * sc + sh > 0, G, SLOPE * G)
        self.view(g, S, H, W, Cn).copy_(G)
        self._add_red(red, G, G * ((Y - mean) * invstd))

    def op_bn_bwd_apply(self, name, args):
        g, y, S, H, W, Cn, sums, red, gamma, dy, dgamma, dbeta = args
        G, Y = self.view(g, S, H, W, Cn).clone(), self.view(y, S, H, W, Cn)
        mean, invstd, sc, _ = self._bn(sums, gamma, None, S, Cn, H * W)
        R = self.vec(red, S * Cn * 2, torch.float64).view(S, Cn, 2)
        m1 = (R[..., 0] / (H * W)).float().reshape(S, 1, 1, Cn)
        m2 = (R[..., 1] / (H * W)).float().reshape(S, 1, 1, Cn)
        out = sc * (G - m1 - (Y - mean) * invstd * m2)
        self.view(dy, S, H, W, Cn, self._dtype_of(name)).copy_(out)
        if dgamma is not None:
            self.vec(dgamma, Cn).copy_(R[..., 1].sum(0).float())
            self.vec(dbeta, Cn).copy_(R[..., 0].sum(0).float())

    def op_cat_up_bwd(self, name, args):
        (dA, S, H, W, mode, ys, Cs, sums_s, gamma_s, beta_s, gs, red_s, yd, Cd, sums_d, gamma_d, beta_d, gd, red_d, part) = args
        DA = self.view(dA, S, H, W, Cs + Cd)
        if Cs and part != 2:
            mean, invstd, sc, sh = self._bn(sums_s, gamma_s, beta_s, S, Cs, H * W)
            Ys = self.view(ys, S, H, W, Cs)
            G = torch.where(Ys * sc + sh > 0, DA[..., :Cs], SLOPE * DA[..., :Cs])
            self.view(gs, S, H, W, Cs).copy_(G)
            self._add_red(red_s, G, G * ((Ys - mean) * invstd))
        if part != 1:
            h2, w2 = H // 2, W // 2
            mean, invstd, sc, sh = self._bn(sums_d, gamma_d, beta_d, S, Cd, h2 * w2)
            Yd = self.view(yd, S, h2, w2, Cd)
            probe = torch.zeros(S, h2, w2, Cd, requires_grad=True)
            self._up2x(probe, mode).backward(DA[..., Cs:].contiguous())
            G = torch.where(Yd * sc + sh > 0, probe.grad, SLOPE * probe.grad)
            self.view(gd, S, h2, w2, Cd).copy_(G)
            self._add_red(red_d, G, G * ((Yd - mean) * invstd))

    def op_fill(self, name, args):
        ptr, n, val = args
        self.vec(ptr, n).fill_(val)

    def op_view_to_bf16(self, name, args):
        src, S, H, W, Cn, dst = args
        self.view(dst, S, H, W, Cn, torch.bfloat16).copy_(self.view(src, S, H, W, Cn))

    OPS = {"mfvi_conv2d_fwd": op_conv_fwd, "mfvi_conv2d_fwd_bf16": op_conv_fwd, "mfvi_conv2d_dgrad": op_conv_dgrad,
           "mfvi_conv2d_dgrad_bf16": op_conv_dgrad, "mfvi_conv2d_wgrad": op_conv_wgrad, "mfvi_conv2d_wgrad_bf16": op_conv_wgrad,
           "mfvi_bn_act_pad_fwd": op_bn_act_pad_fwd, "mfvi_bn_act_pad_fwd_bf16": op_bn_act_pad_fwd, "mfvi_cat_up_fwd": op_cat_up_fwd,
           "mfvi_pad_act_bwd": op_pad_act_bwd, "mfvi_bn_bwd_apply": op_bn_bwd_apply, "mfvi_bn_bwd_apply_bf16": op_bn_bwd_apply,
           "mfvi_cat_up_bwd": op_cat_up_bwd, "mfvi_fill_f32": op_fill, "mfvi_view_f32_to_bf16": op_view_to_bf16}

    def run(self, ops):
        """Executes an op list in order (lanes only express concurrency: program order is a valid schedule)."""
        for name, args, _ in ops:
            if name == "__join__":
                continue
            self.OPS[name](self, name, args)

    # ------------------------------------------------------------------ the pieces around the op lists (trainer.py / tests)
    def set_input(self, x_chw):
        """x0 = reflect_pad(net input) (no jitter), as SkipEngine.set_input with std 0."""
        e = self.eng
        p = e.pad0
        x = x_chw[None] if p == 0 else F.pad(x_chw[None], (p,) * 4, mode="reflect")
        e.x0.copy_(x.permute(0, 2, 3, 1))

    def sample_weights(self):
        """w_s = mu + softplus(rho) * eps_s from the injected eps (mfvi_sample_weights), plus the bf16 repack."""
        e = self.eng
        P = e.lay.P
        e.w[:, :P] = e.mu[None] + F.softplus(e.rho)[None] * e.eps[:, :P]
        if e.w16 is not None:
            e.w16.zero_()
            for c in e.lay.convs:
                src = e.w[:, c.w_off:c.w_off + c.w_numel].view(e.S, c.k * c.k * c.cout, c.cin)
                e.w16[:, c.w16_off:c.w16_off + c.k * c.k * c.cout * c.cpitch].view(e.S, -1, c.cpitch)[..., :c.cin] = src

    def step(self, x_chw, nll_of_out):
        """forward, loss head (`nll_of_out`: (S,C,H,W) -> scalar mean-over-samples data loss), backward.  Returns the loss."""
        e = self.eng
        e.zbuf.zero_()
        self.set_input(x_chw)
        self.sample_weights()
        self.run(e.fwd_ops)
        out = e.out.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        nll = nll_of_out(out)
        nll.backward()
        e.dout.copy_(out.grad.permute(0, 2, 3, 1))
        self.run(e.bwd_ops)
        return float(nll.detach())
