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
        tensors = [eng.theta, eng.grad_buf, eng.zbuf, eng.w, eng.dout, eng.running_mean, eng.running_var] + list(eng._bufs)
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
        # a buffer allocated after the interpreter was built (e.g. the outputs of DeviceBookkeeping.uncertainty()): look it up
        # among the live CPU tensors once, then resolve again
        import gc
        for obj in gc.get_objects():
            if isinstance(obj, torch.Tensor) and obj.device.type == "cpu" and obj.numel() > 0:
                st = obj.untyped_storage()
                if st.data_ptr() <= ptr < st.data_ptr() + st.nbytes() and all(lo != st.data_ptr() for lo, _, _ in self.stores):
                    self.stores.append((st.data_ptr(), st.data_ptr() + st.nbytes(), st))
                    return self._strided(ptr, size, stride, dtype)
        raise AssertionError(f"pointer {ptr:#x} is not inside any live buffer")

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

    # the experimental pair without the intermediate gradient buffer (csrc/elementwise_fused.cu): same contract, g recomputed
    def _folded_g(self, dxp, S, H, W, Cn, pad, y, sums, gamma, beta, act):
        DXP = self.view(dxp, S, H + 2 * pad, W + 2 * pad, Cn)
        Y = self.view(y, S, H, W, Cn)
        mean, invstd, sc, sh = self._bn(sums, gamma, beta, S, Cn, H * W)
        if pad:
            probe = torch.zeros(S, Cn, H, W, requires_grad=True)
            F.pad(probe, (pad,) * 4, mode="reflect").backward(DXP.permute(0, 3, 1, 2).contiguous())
            G = probe.grad.permute(0, 2, 3, 1)
        else:
            G = DXP.clone()
        if act:
            G = torch.where(Y * sc + sh > 0, G, SLOPE * G)
        return G, (Y - mean) * invstd, sc

    def op_pad_act_bwd_reduce(self, name, args):
        dxp, S, H, W, Cn, pad, y, sums, gamma, beta, act, red = args
        G, xhat, _ = self._folded_g(dxp, S, H, W, Cn, pad, y, sums, gamma, beta, act)
        self._add_red(red, G, G * xhat)

    def op_bn_bwd_apply_from_dxp(self, name, args):
        dxp, y, S, H, W, Cn, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta = args
        G, xhat, sc = self._folded_g(dxp, S, H, W, Cn, pad, y, sums, gamma, beta, act)
        R = self.vec(red, S * Cn * 2, torch.float64).view(S, Cn, 2)
        m1 = (R[..., 0] / (H * W)).float().reshape(S, 1, 1, Cn)
        m2 = (R[..., 1] / (H * W)).float().reshape(S, 1, 1, Cn)
        self.view(dy, S, H, W, Cn, self._dtype_of(name)).copy_(sc * (G - m1 - xhat * m2))
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

    OPS = {"mfvi_conv2d_fwd": op_conv_fwd, "mfvi_conv2d_dgrad": op_conv_dgrad, "mfvi_conv2d_wgrad": op_conv_wgrad,
           "mfvi_bn_act_pad_fwd": op_bn_act_pad_fwd, "mfvi_cat_up_fwd": op_cat_up_fwd,
           "mfvi_pad_act_bwd": op_pad_act_bwd, "mfvi_bn_bwd_apply": op_bn_bwd_apply,
           "mfvi_cat_up_bwd": op_cat_up_bwd, "mfvi_fill_f32": op_fill}

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


class TrainerInterpreter(PlanInterpreter):
    """Interprets a whole `MfviDipTrainer(plan_only=True, device="cpu")`: inside `with TrainerInterpreter(tr):` every libmfvidip
    call the trainer makes (`_lib.call`) and every op list its engine runs is executed here on the CPU, so `tr.step()` works —
    host logic under test: Philox keys and the device step counter, MC-sample sharding and the gradient all-reduce (gloo), the
    KL / reparameterisation scaling, AdamW, post-step hooks, the four task heads."""

    def __init__(self, tr):
        super().__init__(tr.eng)
        self.tr = tr
        e = tr.eng
        for t in (tr.saved, tr.step_dev, tr.m, tr.v, tr.losses, e._bn_ch_off, e._bn_sums_off, e._bn_C, e._bn_count):
            self.register(t)
        self.register_object(tr.head)

    def register(self, t):
        if t is None:
            return
        st = t.untyped_storage()
        if all(lo != st.data_ptr() for lo, _, _ in self.stores):
            self.stores.append((st.data_ptr(), st.data_ptr() + st.nbytes(), st))

    def __enter__(self):
        from mfvi_dip_mia_b200 import _lib as L
        from mfvi_dip_mia_b200.engine import SkipEngine
        self._saved = (L.call, SkipEngine._run)
        interp = self

        def call(name, *args, stream=None, meta=None):
            L.launch_count += 1
            interp.dispatch(name, args)

        def _run(engine, op_list):
            assert engine is interp.eng
            interp.run(op_list)
        L.call, SkipEngine._run = call, _run
        return self

    def __exit__(self, *exc):
        from mfvi_dip_mia_b200 import _lib as L
        from mfvi_dip_mia_b200.engine import SkipEngine
        L.call, SkipEngine._run = self._saved

    def register_object(self, obj):
        """registers every tensor attribute of `obj` (e.g. a runners.DeviceBookkeeping, or tensors it allocates on demand)"""
        for v in vars(obj).values():
            if isinstance(v, torch.Tensor):
                self.register(v)

    def dispatch(self, name, args):
        fn = self.OPS.get(name) or self.TRAINER_OPS.get(name)
        if fn is None:
            raise NotImplementedError(f"{name} is not interpreted")
        for hook in self.tr.post_step_hooks:
            if hasattr(hook, "__self__"):
                self.register_object(hook.__self__)
        if self.eng.eps is not None:
            self.register(self.eng.eps)
        if self.tr.noise is not None:
            self.register(self.tr.noise)
        fn(self, name, args)

    # ------------------------------------------------------------------ Philox streams (oracle/philox.py restates the kernels')
    def _step_of(self, key):
        return int(key.step) + (int(self.vec(key.step_dev, 1, torch.int32)[0]) if key.step_dev else 0)

    def _normals(self, n, key, stream, sample):
        from oracle import philox
        return torch.from_numpy(philox.philox_normal(n, int(key.seed), stream, int(key.sample0) + sample, self._step_of(key)))

    # ------------------------------------------------------------------ trainer-level ops
    def op_input_jitter_pad(self, name, args):
        saved, noise, H, W, Cn, std, pad, key, xp = args
        x = self.vec(saved, H * W * Cn).view(H, W, Cn)
        if noise is not None:
            z = self.vec(noise, H * W * Cn).view(H, W, Cn)
        else:                               # the reference's noise tensor is NCHW: flat index (c*H + h)*W + w
            z = self._normals(Cn * H * W, key, 1, 0).view(Cn, H, W).permute(1, 2, 0)
        y = (x + std * z).permute(2, 0, 1)[None]
        if pad:
            y = F.pad(y, (pad,) * 4, mode="reflect")
        self.view(xp, 1, H + 2 * pad, W + 2 * pad, Cn).copy_(y.permute(0, 2, 3, 1))

    def _eps_rows(self, eps, eps_ss, n, S, key):
        if eps is not None:
            return torch.stack([self.vec(eps + 4 * s * eps_ss, n) for s in range(S)])
        return torch.stack([self._normals(n, key, 0, s) for s in range(S)])

    def op_sample_weights(self, name, args):
        mu, rho, n, S, eps, eps_ss, key, w_out, w_ss = args
        m, sg = self.vec(mu, n), F.softplus(self.vec(rho, n))
        E = self._eps_rows(eps, eps_ss, n, S, key)
        for s in range(S):
            self.vec(w_out + 4 * s * w_ss, n).copy_(m + sg * E[s])

    def op_gauss_nll(self, name, args):
        from oracle import mfvi_oracle as O
        mode, out, S, H, W, Cn, sub, target, mask, loss_out, dout = args
        o = self.view(out, S, H, W, Cn).permute(0, 3, 1, 2).clone().requires_grad_(True)
        if mode == 0:
            t = self.vec(target, (H // sub) * (W // sub)).view(1, 1, H // sub, W // sub)
            per = [O.gaussian_nll(q[:, :1], q[:, 1:2], t) for q in (O.sr_downsample_nearest(o[s:s + 1], sub) if sub > 1 else o[s:s + 1]
                                                                      for s in range(S))]
        elif mode == 1:
            t = self.vec(target, H * W * 3).view(H, W, 3).permute(2, 0, 1)[None]
            mk = self.vec(mask, H * W).view(1, 1, H, W)
            per = [O.gaussian_nll_inpainting(torch.sigmoid(o[s:s + 1, :3]), o[s:s + 1, 3:], t, mk) for s in range(S)]
        else:
            raise NotImplementedError(f"nll mode {mode}")
        loss = torch.stack(per).mean()
        loss.backward()
        self.vec(loss_out, 1, torch.float64).add_(float(loss.detach()))
        self.view(dout, S, H, W, Cn).copy_(o.grad.permute(0, 2, 3, 1))

    def op_kl_reparam(self, name, args):
        from oracle import mfvi_oracle as O
        (mu, rho, n, pm, ps, direction, kscale, kscale_dev, dw, dw_ss, S, eps, eps_ss, key, gscale, kl_out, gmu, grho, acc) = args
        if kscale_dev is not None:
            kscale = kscale * float(self.vec(kscale_dev, 1)[0])
        m = self.vec(mu, n).clone().requires_grad_(True)
        r = self.vec(rho, n).clone().requires_grad_(True)
        kl = O.kl_elementwise(m, r, pm, ps, "reverse" if direction == 0 else "forward").sum()
        kl.backward()
        if kl_out is not None:
            self.vec(kl_out, 1, torch.float64).add_(float(kl.detach()))
        if gmu is None:
            return
        g_m, g_r = kscale * m.grad, kscale * r.grad
        if dw is not None and S > 0:
            D = torch.stack([self.vec(dw + 4 * s * dw_ss, n) for s in range(S)])
            E = self._eps_rows(eps, eps_ss, n, S, key)
            g_m = g_m + gscale * D.sum(0)
            g_r = g_r + gscale * (D * E).sum(0) * torch.sigmoid(r.detach())
        GM, GR = self.vec(gmu, n), self.vec(grho, n)
        GM.copy_(GM + g_m if acc else g_m)
        GR.copy_(GR + g_r if acc else g_r)

    def op_bn_running_update(self, name, args):
        arena, ch_off, sums_off, Cs, counts, n_bn, S, mom, rmean, rvar = args
        ch, so = self.vec(ch_off, n_bn, torch.int32), self.vec(sums_off, n_bn, torch.int64)
        Cv, cnt = self.vec(Cs, n_bn, torch.int32), self.vec(counts, n_bn, torch.int32)
        for b in range(n_bn):
            Cn, N = int(Cv[b]), int(cnt[b])
            sums = self.vec(arena + 8 * int(so[b]), S * Cn * 2, torch.float64).view(S, Cn, 2)
            rm, rv = self.vec(rmean + 4 * int(ch[b]), Cn), self.vec(rvar + 4 * int(ch[b]), Cn)
            for s in range(S):
                mean = sums[s, :, 0] / N
                var = (sums[s, :, 1] / N - mean * mean).clamp_min(0.0) * (N / max(N - 1, 1))
                rm.copy_((1 - mom) * rm + mom * mean.float())
                rv.copy_((1 - mom) * rv + mom * var.float())

    def op_adamw(self, name, args):
        p, g, m, v, n, lr, b1, b2, eps, wd, step, step_dev, skip_ptr = args
        if skip_ptr is not None:
            l = float(self.vec(skip_ptr, 1)[0])
            if l != l or abs(l) > 3.0e38:
                return
        t = step + (int(self.vec(step_dev, 1, torch.int32)[0]) if step_dev is not None else 0)
        P, G, M, V = (self.vec(q, n) for q in (p, g, m, v))
        bc1, bc2s = 1.0 - b1 ** t, (1.0 - b2 ** t) ** 0.5
        P.mul_(1.0 - lr * wd)
        M.copy_(b1 * M + (1 - b1) * G)
        V.copy_(b2 * V + (1 - b2) * G * G)
        P.sub_((lr / bc1) * M / (V.sqrt() / bc2s + eps))

    def op_counter_add(self, name, args):
        ptr, inc = args
        self.vec(ptr, 1, torch.int32).add_(int(inc))

    def op_counter_add_if_finite(self, name, args):
        ptr, inc, flag = args
        l = float(self.vec(flag, 1)[0])
        if l == l and abs(l) <= 3.0e38:
            self.vec(ptr, 1, torch.int32).add_(int(inc))

    def op_loss_flag(self, name, args):
        kl, nll, temp, flag = args
        self.vec(flag, 1)[0] = float(self.vec(nll, 1, torch.float64)[0] + temp * self.vec(kl, 1, torch.float64)[0])

    # ------------------------------------------------------------------ CT head: radon projector + sinogram MSE
    def _theta_deg(self, theta_rad, T):
        return torch.rad2deg(self.vec(theta_rad, T).double()).float()

    def op_radon_fwd(self, name, args):
        from oracle import mfvi_oracle as O
        img, S, Cn, H, W, theta, T, sino = args
        X = self.view(img, S, H, W, Cn).permute(0, 3, 1, 2)
        out = self.vec(sino, S * Cn * T * W).view(S, Cn, T, W)
        for s in range(S):
            out[s] = O.radon_forward(X[s:s + 1].contiguous(), self._theta_deg(theta, T))[0]

    def op_radon_bwd(self, name, args):
        from oracle import mfvi_oracle as O
        dsino, S, Cn, H, W, theta, T, dimg = args
        D = self.vec(dsino, S * Cn * T * W).view(S, Cn, T, W)
        G = self.view(dimg, S, H, W, Cn)
        for s in range(S):
            probe = torch.zeros(1, Cn, H, W, requires_grad=True)
            O.radon_forward(probe, self._theta_deg(theta, T)).backward(D[s:s + 1])
            G[s] = probe.grad[0].permute(1, 2, 0)

    def op_mse(self, name, args):
        a, a_ss, b, n, S, loss_out, da = args
        B = self.vec(b, n)
        for s in range(S):
            diff = self.vec(a + 4 * s * a_ss, n) - B
            if loss_out is not None:
                self.vec(loss_out, 1, torch.float64).add_(float((diff.double() ** 2).sum()) / (n * S))
            if da is not None:
                self.vec(da + 4 * s * a_ss, n).copy_(2.0 * diff / (n * S))

    # ------------------------------------------------------------------ runner bookkeeping (csrc/bookkeeping.cu)
    def op_bookkeep_step(self, name, args):
        out, S, H, W, expw, gt, noisy, out_avg, ring_epi, ring_ale, R, iter_dev, iter_off, acc = args
        return self.op_bookkeep_step_ex(name, (out, S, H, W, 1, 2, expw, gt, noisy, None, out_avg, ring_epi, ring_ale, R, iter_dev,
                                               iter_off, acc))

    def op_bookkeep_step_ex(self, name, args):
        out, S, H, W, Cm, flags, expw, gt, noisy, mask, out_avg, ring_epi, ring_ale, R, iter_dev, iter_off, acc = args
        it = iter_off + (int(self.vec(iter_dev, 1, torch.int32)[0]) if iter_dev is not None else 0)
        sig, ale = bool(flags & 1), bool(flags & 2)
        O_ = self.view(out, S, H, W, Cm + (1 if ale else 0))
        img = O_[..., :Cm]
        m = (torch.sigmoid(img) if sig else img).mean(0).permute(2, 0, 1)                # (Cm,H,W)
        AVG = self.vec(out_avg, (Cm + (1 if ale else 0)) * H * W).view(-1, H, W)
        AVG[:Cm] = m if it == 0 else AVG[:Cm] * expw + m * (1 - expw)
        mc = m.clamp(0, 1)
        if ale:
            v = torch.exp(-O_[..., Cm]).mean(0)
            AVG[Cm] = v if it == 0 else AVG[Cm] * expw + v * (1 - expw)
        if R > 0:
            self.vec(ring_epi, Cm * R * H * W).view(Cm, R, H, W)[:, it % R] = mc
            if ale:
                self.vec(ring_ale, R * H * W).view(R, H, W)[it % R] = v.clamp(0, 1)
        A = self.vec(acc, 5, torch.float64)
        am, amc = AVG[:Cm], AVG[:Cm].clamp(0, 1)
        mk = self.vec(mask, H * W).view(1, H, W) if mask is not None else 1.0
        if noisy is not None:
            t = self.vec(noisy, Cm * H * W).view(Cm, H, W)
            A[0] += ((t - mc) ** 2).double().sum()
            A[3] += ((t - am) ** 2).double().sum()
        if gt is not None:
            t = self.vec(gt, Cm * H * W).view(Cm, H, W)
            A[1] += ((t * mk - mc * mk) ** 2).double().sum()
            A[2] += ((t * mk - amc * mk) ** 2).double().sum()
            A[4] += ((t - am) ** 2).double().sum()

    def op_ssim(self, name, args):
        from oracle import mfvi_oracle as O
        a, b, H, W, clip_b, out_sum = args
        A, B = self.vec(a, H * W).view(1, 1, H, W), self.vec(b, H * W).view(1, 1, H, W)
        if clip_b:
            B = B.clamp(0, 1)
        self.vec(out_sum, 1, torch.float64).add_(float(O.ssim(A, B)) * H * W)

    def op_ring_uncertainty(self, name, args):
        ring_epi, ring_ale, n, H, W, gt, epi, ale, err2 = args
        RE = self.vec(ring_epi, n * H * W).view(n, H, W)
        RA = self.vec(ring_ale, n * H * W).view(n, H, W)
        self.vec(epi, H * W).view(H, W).copy_(RE.var(0) if n > 1 else torch.zeros(H, W))
        self.vec(ale, H * W).view(H, W).copy_(RA.mean(0))
        if err2 is not None:
            t = self.vec(gt, H * W).view(H, W) if gt is not None else torch.zeros(H, W)
            self.vec(err2, H * W).view(H, W).copy_(((RE - t) ** 2).mean(0))

    TRAINER_OPS = {"mfvi_radon_fwd": op_radon_fwd, "mfvi_radon_bwd": op_radon_bwd, "mfvi_mse_fwd_bwd": op_mse,
                   "mfvi_bookkeep_step": op_bookkeep_step, "mfvi_bookkeep_step_ex": op_bookkeep_step_ex, "mfvi_ssim": op_ssim, "mfvi_ring_uncertainty": op_ring_uncertainty,
                   "mfvi_input_jitter_pad": op_input_jitter_pad, "mfvi_sample_weights": op_sample_weights,
                   "mfvi_gauss_nll_fwd_bwd": op_gauss_nll,
                   "mfvi_kl_reparam_fwd_bwd": op_kl_reparam, "mfvi_bn_running_update": op_bn_running_update,
                   "mfvi_adamw_step": op_adamw, "mfvi_counter_add": op_counter_add,
                   "mfvi_counter_add_if_finite": op_counter_add_if_finite, "mfvi_loss_flag": op_loss_flag}
