"""Execution engine of the MFVI-DIP hour-glass network: the host-side plan that strings the libmfvidip kernels
into the forward and backward of `models/skip.py` (reference models/skip.py:58-132, models/common.py:23-135)
for S Monte-Carlo weight samples at once.

Data layout in HBM (all fp32):
  * `theta`  = [ mu (P) | rho (P) | bn_gamma (Q) | bn_beta (Q) ]   — every trainable scalar of the net, one buffer
    (one AdamW launch, one all-reduce).  Inside mu/rho all weight blocks come first, layer after layer in
    eps-draw order, each stored tap-major [KH][KW][Cout][Cin]; the bias vectors follow.
  * `grad`   = same layout as theta.
  * `w`, `dw`, `eps` = [S][P] sampled weights / per-sample weight gradients / injected eps (storage layout).
  * activations: NHWC (S,H,W,C); every conv input is a reflection-padded buffer written by the fused
    BN+LeakyReLU+pad kernel, so convs are plain "valid" convolutions over a padded view.
  * `arena` (float64): [kl, nll, …] scalars followed by per-BatchNorm (sum, sumsq) and backward (sum g, sum g*xhat)
    tables, [S][C][2] each.

The plan is a static list of (C-ABI symbol, ctypes args): it is built once, executes with no host-side
tensor work and is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L


@dataclass
class SkipSpec:
    """Arguments of skip() that determine the graph (reference models/skip.py:5-14)."""
    num_input_channels: int = 16
    num_output_channels: int = 2
    down: Sequence[int] = (16, 32, 64, 128, 128)
    up: Sequence[int] = (16, 32, 64, 128, 128)
    skip: Sequence[int] = (4, 4, 4, 4, 4)
    filter_down: int = 3
    filter_up: int = 3
    filter_skip: int = 1
    need1x1_up: bool = True
    need_sigmoid: bool = False
    upsample_mode: str = "bilinear"


@dataclass
class ConvLayer:
    key: str          # state-dict prefix of the Conv2dRT module
    cin: int
    cout: int
    k: int
    stride: int
    index: int = -1   # position in eps-draw order
    w_off: int = 0    # offset of the [k][k][cout][cin] block inside mu / rho
    b_off: int = 0    # offset of the bias vector inside mu / rho
    @property
    def w_numel(self):
        return self.cout * self.cin * self.k * self.k


@dataclass
class BnLayer:
    key: str
    C: int
    index: int = -1
    ch_off: int = 0     # offset inside gamma / beta / running_mean / running_var
    sums_off: int = 0   # offset (in doubles) of the [S][C][2] (sum, sumsq) table inside the arena
    red_off: int = 0    # offset (in doubles) of the backward reduction table
    count: int = 0      # H*W the statistics are taken over


@dataclass
class ScaleLayers:
    skip_conv: Optional[ConvLayer]
    skip_bn: Optional[BnLayer]
    d1: ConvLayer
    d1_bn: BnLayer
    d2: ConvLayer
    d2_bn: BnLayer
    cat_bn: BnLayer
    up: ConvLayer
    up_bn: BnLayer
    up1: Optional[ConvLayer]
    up1_bn: Optional[BnLayer]


@dataclass
class SkipLayout:
    """Module keys of the hour-glass net as produced by skip() + rename_modules
    (reference models/skip.py:56-132, utils/common_utils.py:248-262)."""
    spec: SkipSpec
    scales: List[ScaleLayers] = field(default_factory=list)
    final: ConvLayer = None
    convs: List[ConvLayer] = field(default_factory=list)   # eps-draw (execution) order
    bns: List[BnLayer] = field(default_factory=list)       # module-tree order of first use
    P: int = 0        # number of (mu, rho) pairs
    P_pad: int = 0    # P rounded up to a multiple of 4
    Q: int = 0        # number of BatchNorm channels


def build_layout(spec: SkipSpec, root: str = "") -> SkipLayout:
    n = len(spec.down)
    assert len(spec.up) == n and len(spec.skip) == n
    lay = SkipLayout(spec)
    n_skip, n_deep = 1, 1
    n_up = 2 * n if spec.need1x1_up else n
    prefix = root
    cin = spec.num_input_channels
    for i in range(n):
        k = n_up - 1
        if spec.skip[i] != 0:
            deep_prefix = f"{prefix}Concat_up_{k}.1."
            sp = f"{prefix}Concat_up_{k}.0."
            skip_conv = ConvLayer(f"{sp}Sequential_skip_{n_skip}.Conv2d_skip_{n_skip}", cin, spec.skip[i], spec.filter_skip, 1)
            skip_bn = BnLayer(f"{sp}BatchNorm2d_skip_{n_skip}", spec.skip[i])
            n_skip += 1
            up_seq = f"Sequential_up_{k}"
        else:
            # without a skip branch the deeper Sequential itself is named Sequential_up_k, the conv block collides
            deep_prefix = f"{prefix}Sequential_up_{k}."
            skip_conv = skip_bn = None
            up_seq = f"Sequential_up_{k}_1"
        d1 = ConvLayer(f"{deep_prefix}Sequential_deeper_{n_deep}.Conv2d_deeper_{n_deep}", cin, spec.down[i], spec.filter_down, 2)
        d1_bn = BnLayer(f"{deep_prefix}BatchNorm2d_deeper_{n_deep}", spec.down[i])
        n_deep += 1
        d2 = ConvLayer(f"{deep_prefix}Sequential_deeper_{n_deep}.Conv2d_deeper_{n_deep}", spec.down[i], spec.down[i], spec.filter_down, 1)
        d2_bn = BnLayer(f"{deep_prefix}BatchNorm2d_deeper_{n_deep}", spec.down[i])
        n_deep += 1
        c_deep = spec.up[i + 1] if i < n - 1 else spec.down[i]
        cat_bn = BnLayer(f"{prefix}BatchNorm2d_up_{k}", spec.skip[i] + c_deep)
        up = ConvLayer(f"{prefix}{up_seq}.Conv2d_up_{k}", spec.skip[i] + c_deep, spec.up[i], spec.filter_up, 1)
        up_bn = BnLayer(f"{prefix}BatchNorm2d_up_{k}_1", spec.up[i])
        if spec.need1x1_up:
            up1 = ConvLayer(f"{prefix}Sequential_up_{k + 1}.Conv2d_up_{k + 1}", spec.up[i], spec.up[i], 1, 1)
            up1_bn = BnLayer(f"{prefix}BatchNorm2d_up_{k + 1}", spec.up[i])
            n_up -= 1
        else:
            up1 = up1_bn = None
        n_up -= 1
        lay.scales.append(ScaleLayers(skip_conv, skip_bn, d1, d1_bn, d2, d2_bn, cat_bn, up, up_bn, up1, up1_bn))
        cin = spec.down[i]
        prefix = deep_prefix + "7."
    root_children = 5 + (3 if spec.need1x1_up else 0)
    final_iter = 2 * n + 1 if spec.need1x1_up else n + 1
    lay.final = ConvLayer(f"{root}{root_children + 1}.Conv2d_up_{final_iter}", spec.up[0], spec.num_output_channels, 1, 1)

    # eps-draw order: skip branch of a Concat runs before the deeper branch (models/common.py:25-26)
    pre, post = [], []
    for sc in lay.scales:
        if sc.skip_conv is not None:
            pre.append(sc.skip_conv)
        pre += [sc.d1, sc.d2]
    for sc in reversed(lay.scales):
        post.append(sc.up)
        if sc.up1 is not None:
            post.append(sc.up1)
    lay.convs = pre + post + [lay.final]
    off = 0
    for i, c in enumerate(lay.convs):
        c.index = i
        c.w_off = off
        off += c.w_numel
    for c in lay.convs:
        c.b_off = off
        off += c.cout
    lay.P = off
    lay.P_pad = (off + 3) // 4 * 4
    q = 0
    for sc in lay.scales:
        for b in (sc.skip_bn, sc.d1_bn, sc.d2_bn, sc.cat_bn, sc.up_bn, sc.up1_bn):
            if b is not None:
                b.index = len(lay.bns)
                b.ch_off = q
                q += b.C
                lay.bns.append(b)
    lay.Q = q
    return lay


# ------------------------------------------------------------------------------------------------------------
N_SCALARS = 8          # doubles reserved at the head of the arena: [0]=kl, [1]=data loss, rest spare
KL, NLL = 0, 1


class SkipEngine:
    """Forward/backward plan of one hour-glass net at a fixed (H, W, S)."""

    def __init__(self, spec: SkipSpec, H: int, W: int, S: int, device, *, math: int = L.MATH_FP32,
                 layout: Optional[SkipLayout] = None, need_input_grad: bool = False, plan_only: bool = False,
                 mega_from: Optional[int] = None):
        device = torch.device(device)
        # A plan-only engine builds the kernel plan (buffers, views, op lists) and can never execute it: device "meta" (shapes,
        # strides and alignments, no memory) for conv_dispatch_table() on a machine without a GPU; plan_only=True on any device
        # (real buffers) for the test-suite's CPU interpretation of the plan (tests/plan_interpreter.py)
        self.plan_only = plan_only or device.type == "meta"
        if device.type != "cuda" and not self.plan_only:
            raise L.MfviError(f"SkipEngine needs a CUDA device, got {device}: there is no CPU fallback")
        n = len(spec.down)
        if H % (1 << n) or W % (1 << n):
            raise L.MfviError(f"image size {H}x{W} must be a multiple of 2^{n} (the runners crop to multiples of 32; "
                              "centre-crop concat of odd sizes is not implemented)")
        if spec.need_sigmoid:
            raise L.MfviError("need_sigmoid=True is not supported by the fused engine (no MFVI runner uses it)")
        if spec.upsample_mode not in ("bilinear", "nearest"):
            raise L.MfviError(f"upsample_mode={spec.upsample_mode!r} is not supported")
        self.spec, self.H, self.W, self.S, self.device, self.math = spec, H, W, S, device, math
        self.lay = layout or build_layout(spec)
        lay = self.lay
        P, Pp, Q = lay.P, lay.P_pad, lay.Q
        f32 = dict(dtype=torch.float32, device=device)
        # ---- parameters, gradients, optimiser state
        self.n_theta = 2 * Pp + 2 * Q
        self.n_theta_pad = (self.n_theta + 3) // 4 * 4
        self.theta = torch.zeros(self.n_theta_pad, **f32)
        # 4 floats beyond the parameters: [0] = the step's loss (NaN-guard flag), all-reduced together with the gradient
        self.grad_buf = torch.zeros(self.n_theta_pad + 4, **f32)
        self.grad = self.grad_buf[:self.n_theta_pad]
        self.mu, self.rho = self.theta[:P], self.theta[Pp:Pp + P]
        self.gamma, self.beta = self.theta[2 * Pp:2 * Pp + Q], self.theta[2 * Pp + Q:2 * Pp + 2 * Q]
        self.g_mu, self.g_rho = self.grad[:P], self.grad[Pp:Pp + P]
        self.g_gamma, self.g_beta = self.grad[2 * Pp:2 * Pp + Q], self.grad[2 * Pp + Q:2 * Pp + 2 * Q]
        self.gamma.fill_(1.0)
        self.running_mean = torch.zeros(Q, **f32)
        self.running_var = torch.ones(Q, **f32)
        # ---- arena layout
        n_dbl = N_SCALARS
        for b in lay.bns:
            b.sums_off = n_dbl
            n_dbl += S * b.C * 2
            b.red_off = n_dbl
            n_dbl += S * b.C * 2
        self.n_arena = n_dbl
        # dw and the arena share one buffer so that a single fill zeroes every atomic accumulator of the step
        self.zbuf = torch.zeros(S * Pp + 2 * n_dbl, **f32)
        self.dw = self.zbuf[:S * Pp].view(S, Pp)
        self.arena = self.zbuf[S * Pp:].view(torch.float64)
        self.w = torch.zeros(S, Pp, **f32)
        self.eps = None                      # [S][Pp] injected eps (allocated on demand)
        self.inject_eps = False
        self._bufs: List[torch.Tensor] = []
        self.fwd_ops: List[Tuple[str, tuple, dict]] = []
        self.bwd_ops: List[Tuple[str, tuple, dict]] = []
        self.need_input_grad = need_input_grad
        # weight-gradient kernels only feed dw (consumed after the whole backward), so they run on a side stream
        # concurrently with the dgrad / elementwise chain (fork-join by events; becomes a parallel branch of the graph)
        # The 1x1 skip-branch convolutions (forward) and their backward chain (BN backward, dgrad into the scale's input
        # gradient) hang off the long down/up chain, so they run on a third stream ("skip" lane) and join where the
        # concatenation needs them.  Ops carry their lane in meta["lane"]; meta["after"] names the lane that produced
        # their inputs; ("__join__", (), {"lane": X}) makes the main stream wait for lane X.
        self.overlap_wgrad = True
        self.overlap_skip = os.environ.get("MFVI_SKIP_LANE", "1") != "0"
        # Persistent per-sample kernel (csrc/mega.cu): the down path of scale `mega_from`, everything below it and its upsample +
        # concat run as ONE launch per direction instead of one launch per op (the scale's own skip and up convolutions, which
        # work at twice the resolution, stay outside).  None = the library default (by map size, _default_mega_from); -1 or
        # MFVI_MEGA=0 = off.
        env = os.environ.get("MFVI_MEGA_FROM")
        if mega_from is None and env is not None:
            mega_from = int(env)
        if mega_from is None:
            mega_from = self._default_mega_from()
        if self.plan_only or os.environ.get("MFVI_MEGA", "1") == "0" or mega_from < 1 or mega_from >= len(spec.down):
            mega_from = None
        self.mega_from = mega_from
        self._mega = []                      # recorded programs (kept alive)
        self._side = None if self.plan_only else torch.cuda.Stream(device=device)
        self._side2 = None if self.plan_only else torch.cuda.Stream(device=device)
        self._build_plan()
        # per-BN tables for the running-stat update
        self._bn_ch_off = torch.tensor([b.ch_off for b in lay.bns], dtype=torch.int32, device=device)
        self._bn_sums_off = torch.tensor([b.sums_off for b in lay.bns], dtype=torch.int64, device=device)
        self._bn_C = torch.tensor([b.C for b in lay.bns], dtype=torch.int32, device=device)
        self._bn_count = torch.tensor([b.count for b in lay.bns], dtype=torch.int32, device=device)

    # ---------------------------------------------------------------- helpers
    def _default_mega_from(self) -> int:
        """First scale whose down-path maps are small enough for the persistent per-sample kernel: (H / 2^(i+1)) * (W / 2^(i+1))
        pixels per sample <= MFVI_MEGA_PIXELS (default 0 = off until measured)."""
        limit = int(os.environ.get("MFVI_MEGA_PIXELS", "0"))
        for i in range(1, len(self.spec.down)):
            if (self.H >> (i + 1)) * (self.W >> (i + 1)) <= limit:
                return i
        return -1

    def _fuse(self, ops, n0: int, n1: int, tag: str):
        """Replace ops[n0:n1] by one mfvi_mega_run op.  Lane joins of the slice are kept in front of it (the program runs on the
        main lane), weight gradients stay separate launches on their lane behind it, and the gradients of the BatchNorm affine
        parameters (a sum over all samples, which the per-sample program cannot form) follow as one mfvi_bn_param_grads."""
        sub = ops[n0:n1]
        inner = [op for op in sub if op[0] not in ("__join__", "mfvi_conv2d_wgrad")]
        if not inner:
            return
        joins = [op for op in sub if op[0] == "__join__"]
        later = [op for op in sub if op[0] == "mfvi_conv2d_wgrad"]
        m = L.record_program(inner, self.device)
        self._mega.append(m)
        meta = {"flops": sum(o[2].get("flops", 0.0) for o in inner), "bytes": sum(o[2].get("bytes", 0.0) for o in inner),
                "layer": f"mega_{tag}", "shape": f"{m['n_stages']} stages", "sub_ops": inner}
        new = joins + [("mfvi_mega_run", (m["prog"].data_ptr(), m["n_stages"], self.S, L.ptr(m["prof"])), meta)]
        bns = [op[1] for op in inner if op[0] == "mfvi_bn_bwd_apply" and op[1][10] is not None]
        if bns:
            i64 = lambda vals: torch.tensor(vals, dtype=torch.int64, device=self.device)
            tabs = (i64([a[7] for a in bns]), i64([a[10] for a in bns]), i64([a[11] for a in bns]),
                    torch.tensor([a[5] for a in bns], dtype=torch.int32, device=self.device))
            m["bn_tables"] = tabs
            new.append(("mfvi_bn_param_grads", (tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(), tabs[3].data_ptr(),
                                                len(bns), self.S), {"bytes": 16.0 * self.S * sum(a[5] for a in bns)}))
        ops[n0:n1] = new + later

    def _buf(self, H, W, Cn, S=None):
        t = torch.empty(self.S if S is None else S, H, W, Cn, dtype=torch.float32, device=self.device)
        self._bufs.append(t)
        return t

    def _aptr(self, off):
        return self.arena.data_ptr() + 8 * off

    def _bn_args(self, b: BnLayer):
        """(sums ptr, gamma ptr, beta ptr) of a BatchNorm."""
        return self._aptr(b.sums_off), self.gamma.data_ptr() + 4 * b.ch_off, self.beta.data_ptr() + 4 * b.ch_off

    def _interior(self, t, pad):
        return t if pad == 0 else t[:, pad:t.shape[1] - pad, pad:t.shape[2] - pad, :]

    def _desc(self, c: ConvLayer, Hin, Win):
        Ho, Wo = (Hin - c.k) // c.stride + 1, (Win - c.k) // c.stride + 1
        return L.ConvDesc(self.S, c.cin, c.cout, c.k, c.k, c.stride, Hin, Win, Ho, Wo, self.math), Ho, Wo

    def _conv_fwd(self, c: ConvLayer, x: torch.Tensor, bn: Optional[BnLayer]):
        """x: padded input view (S or 1, Hin, Win, cin) -> raw conv output y (S,Ho,Wo,cout)."""
        d, Ho, Wo = self._desc(c, x.shape[1], x.shape[2])
        # channel pitch rounded up to 4 floats: a 2-channel output keeps 16-byte aligned pixels, so TMA can read its
        # gradient (dgrad / wgrad of the final conv run on the tensor-core kernels too)
        y = self._buf(Ho, Wo, (c.cout + 3) // 4 * 4)[..., :c.cout]
        if bn is not None:
            bn.count = Ho * Wo
        self.fwd_ops.append(("mfvi_conv2d_fwd", (
            C.byref(d), L.view(x), self.w.data_ptr() + 4 * c.w_off, self.w.data_ptr() + 4 * c.b_off, self.lay.P_pad,
            L.view(y), None if bn is None else self._aptr(bn.sums_off)), self._conv_meta(c, d, x, y)))
        self._keep.append(d)
        return y, d

    def _conv_meta(self, c: ConvLayer, d, x, y):
        """Algorithmic work of one conv launch (fwd, dgrad and wgrad all perform the same MACs):
        flops = 2*S*Hout*Wout*Cout*Cin*KH*KW; bytes = input read once + sampled weights read once + output written once."""
        flops = 2.0 * self.S * d.Hout * d.Wout * c.cout * c.cin * c.k * c.k
        nbytes = x.element_size() * x.numel() + 4.0 * self.S * c.w_numel + 4.0 * self.S * c.cout \
            + y.element_size() * y.numel()
        return {"flops": flops, "bytes": nbytes, "layer": c.key.rsplit(".", 1)[-1],
                "shape": f"{c.cin}->{c.cout} k{c.k} s{c.stride} out{d.Hout}x{d.Wout}"}

    @staticmethod
    def _ew_meta(*tensors):
        """Elementwise kernels: every operand tensor is read or written exactly once."""
        return {"bytes": float(sum(t.element_size() * t.numel() for t in tensors if t is not None))}

    def _bn_act_pad(self, y, bn: BnLayer, sums_ptr, gamma_ptr, beta_ptr, act, pad):
        S, H, W, Cn = y.shape
        xp = self._buf(H + 2 * pad, W + 2 * pad, Cn)
        self.fwd_ops.append(("mfvi_bn_act_pad_fwd", (L.view(y), S, H, W, Cn, sums_ptr, gamma_ptr, beta_ptr, act, pad,
                                                     L.view(xp)), self._ew_meta(y, xp)))
        return xp

    # backward of  x = pad(act(bn(y)))  followed by the BN statistics backward:  dxp -> dy (returned)
    def _bn_act_pad_bwd(self, ops, dxp, y, bn: BnLayer, act, pad):
        S, H, W, Cn = y.shape
        sums, gamma, beta = self._bn_args(bn)
        red = self._aptr(bn.red_off)
        g = self._buf(H, W, Cn)
        ops.append(("mfvi_pad_act_bwd", (L.view(dxp), S, H, W, Cn, pad, L.view(y), sums, gamma, beta, act, L.view(g), red),
                    self._ew_meta(dxp, y, g)))
        ops.append(("mfvi_bn_bwd_apply", (L.view(g), L.view(y), S, H, W, Cn, sums, red, gamma, L.view(g),
                                          self.g_gamma.data_ptr() + 4 * bn.ch_off, self.g_beta.data_ptr() + 4 * bn.ch_off),
                    self._ew_meta(g, y, g)))
        return g

    def _dbias_ptr(self, c: ConvLayer, bn_follows: bool):
        """A conv bias that feeds a train-mode BatchNorm has an identically zero data gradient (BN subtracts the
        per-sample channel mean, so sum_pixels dL/dy = gamma*invstd*(sum g - sum g - mean(g*xhat)*sum xhat) = 0; the
        reference's autograd returns rounding noise ~1e-9 there).  The engine leaves dw[s][bias] at its zero fill for those
        layers instead of reducing dy again; only the final conv (no BN behind it) computes a bias gradient."""
        return None if bn_follows else self.dw.data_ptr() + 4 * c.b_off

    def _conv_bwd(self, ops, c: ConvLayer, d, x, dy, need_dx=True, bn_follows=True):
        """wgrad into dw[s] (+bias), dgrad into a fresh padded-input-sized buffer (returned)."""
        meta = self._conv_meta(c, d, x, dy)
        ops.append(("mfvi_conv2d_wgrad", (C.byref(d), L.view(x), L.view(dy), self.dw.data_ptr() + 4 * c.w_off,
                                          self._dbias_ptr(c, bn_follows), self.lay.P_pad), meta))
        if not need_dx:
            return None
        dx = self._buf(x.shape[1], x.shape[2], c.cin)
        ops.append(("mfvi_conv2d_dgrad", (C.byref(d), L.view(dy), self.w.data_ptr() + 4 * c.w_off, self.lay.P_pad,
                                          L.view(dx), 0), meta))
        return dx

    # ---------------------------------------------------------------- plan
    @L.on_device
    def _build_plan(self):
        spec, lay, S = self.spec, self.lay, self.S
        self._keep = []
        mode = 0 if spec.upsample_mode == "bilinear" else 1
        pd, pu, ps = (spec.filter_down - 1) // 2, (spec.filter_up - 1) // 2, (spec.filter_skip - 1) // 2
        n = len(lay.scales)

        def in_pad(i):   # padding of the buffer feeding scale i (skip conv and first down conv read it)
            return max(pd, ps if lay.scales[i].skip_conv is not None else 0)

        # padded network input: ONE image broadcast to all samples
        self.pad0 = in_pad(0)
        self.x0 = self._buf(self.H + 2 * self.pad0, self.W + 2 * self.pad0, spec.num_input_channels, S=1)
        self.dx0 = None

        def run_scale(i, T, Tpad):
            """T: padded (by Tpad) activated input of scale i.  Returns (z, z_bn, backward closure)."""
            sc = lay.scales[i]
            ys = d_s = None
            if sc.skip_conv is not None:
                ys, d_s = self._conv_fwd(sc.skip_conv, self._interior(T, Tpad - ps), sc.skip_bn)
                self.fwd_ops[-1][2].update(lane="skip", after="main")
            n_fuse0 = len(self.fwd_ops)
            x_d1 = self._interior(T, Tpad - pd)
            y1, d_1 = self._conv_fwd(sc.d1, x_d1, sc.d1_bn)
            X2 = self._bn_act_pad(y1, sc.d1_bn, *self._bn_args(sc.d1_bn), 1, pd)
            y2, d_2 = self._conv_fwd(sc.d2, X2, sc.d2_bn)
            if i < n - 1:
                npad = in_pad(i + 1)
                Tn = self._bn_act_pad(y2, sc.d2_bn, *self._bn_args(sc.d2_bn), 1, npad)
                z, z_bn, inner_bwd = run_scale(i + 1, Tn, npad)
            else:
                z, z_bn, inner_bwd = y2, sc.d2_bn, None
            Hs, Ws = 2 * z.shape[1], 2 * z.shape[2]
            Cs = sc.skip_conv.cout if sc.skip_conv is not None else 0
            Cd = z.shape[3]
            A = self._buf(Hs, Ws, Cs + Cd)
            sc.cat_bn.count = Hs * Ws
            null_view = L.View(None, 0, 0, 0)
            sb = self._bn_args(sc.skip_bn) if Cs else (None, None, None)
            zb = self._bn_args(z_bn)
            if Cs:
                self.fwd_ops.append(("__join__", (), {"lane": "skip"}))
            self.fwd_ops.append(("mfvi_cat_up_fwd", (L.view(ys) if Cs else null_view, Cs, *sb, L.view(z), Cd, *zb,
                                                     S, Hs, Ws, mode, L.view(A), self._aptr(sc.cat_bn.sums_off)),
                                 self._ew_meta(ys, z, A)))
            if self.mega_from is not None and i == self.mega_from:
                self._fuse(self.fwd_ops, n_fuse0, len(self.fwd_ops), "fwd")
            XA = self._bn_act_pad(A, sc.cat_bn, *self._bn_args(sc.cat_bn), 0, pu)
            yu, d_u = self._conv_fwd(sc.up, XA, sc.up_bn)
            if sc.up1 is not None:
                X1 = self._bn_act_pad(yu, sc.up_bn, *self._bn_args(sc.up_bn), 1, 0)
                yu1, d_u1 = self._conv_fwd(sc.up1, X1, sc.up1_bn)
                z_out, z_out_bn = yu1, sc.up1_bn
            else:
                z_out, z_out_bn = yu, sc.up_bn

            def backward(ops, dz_out):
                """dz_out: gradient w.r.t. the raw conv output z_out.  Returns dT (padded) or None for scale 0."""
                if sc.up1 is not None:
                    dX1 = self._conv_bwd(ops, sc.up1, d_u1, X1, dz_out)
                    dyu = self._bn_act_pad_bwd(ops, dX1, yu, sc.up_bn, 1, 0)
                else:
                    dyu = dz_out
                dXA = self._conv_bwd(ops, sc.up, d_u, XA, dyu)
                dA = self._bn_act_pad_bwd(ops, dXA, A, sc.cat_bn, 0, pu)
                gd = self._buf(z.shape[1], z.shape[2], Cd)
                gs = self._buf(Hs, Ws, Cs) if Cs else None
                cat_args = (L.view(dA), S, Hs, Ws, mode, L.view(ys) if Cs else null_view, Cs, *sb,
                            L.view(gs) if Cs else null_view, self._aptr(sc.skip_bn.red_off) if Cs else None,
                            L.view(z), Cd, *zb, L.view(gd), self._aptr(z_bn.red_off))
                # The skip branch's half goes to the skip lane first (it only needs dA): LeakyReLU + BN backward, weight gradient,
                # then its dgrad into the (zero-filled) input gradient of this scale, which the first down conv's dgrad accumulates
                # onto after the join.  The upsampled branch then continues the main chain.
                need_dT = i > 0 or self.need_input_grad
                dT = self._buf(T.shape[1], T.shape[2], T.shape[3]) if need_dT else None
                if Cs:
                    ops.append(("mfvi_cat_up_bwd", cat_args + (1,), dict(self._ew_meta(dA[..., :Cs], ys, gs), lane="skip", after="main")))
                    x_s = self._interior(T, Tpad - ps)
                    ms = self._conv_meta(sc.skip_conv, d_s, x_s, gs)
                    ops.append(("mfvi_bn_bwd_apply", (
                        L.view(gs), L.view(ys), S, Hs, Ws, Cs, sb[0], self._aptr(sc.skip_bn.red_off), sb[1], L.view(gs),
                        self.g_gamma.data_ptr() + 4 * sc.skip_bn.ch_off, self.g_beta.data_ptr() + 4 * sc.skip_bn.ch_off),
                        dict(self._ew_meta(gs, ys, gs), lane="skip", after="skip")))
                    ops.append(("mfvi_conv2d_wgrad", (C.byref(d_s), L.view(x_s), L.view(gs), self.dw.data_ptr() + 4 * sc.skip_conv.w_off,
                                                      self._dbias_ptr(sc.skip_conv, True), lay.P_pad), dict(ms, after="skip")))
                    if need_dT:
                        ops.append(("mfvi_fill_f32", (dT.data_ptr(), dT.numel(), 0.0), dict(self._ew_meta(dT), lane="skip", after="skip")))
                        ops.append(("mfvi_conv2d_dgrad", (C.byref(d_s), L.view(gs), self.w.data_ptr() + 4 * sc.skip_conv.w_off,
                                                          lay.P_pad, L.view(self._interior(dT, Tpad - ps)), 1),
                                    dict(ms, lane="skip", after="skip")))
                n_fuse0 = len(ops)
                ops.append(("mfvi_cat_up_bwd", cat_args + (2,), self._ew_meta(dA[..., Cs:], z, gd)))
                ops.append(("mfvi_bn_bwd_apply", (
                    L.view(gd), L.view(z), S, z.shape[1], z.shape[2], Cd, zb[0], self._aptr(z_bn.red_off), zb[1], L.view(gd),
                    self.g_gamma.data_ptr() + 4 * z_bn.ch_off, self.g_beta.data_ptr() + 4 * z_bn.ch_off),
                    self._ew_meta(gd, z, gd)))
                if inner_bwd is not None:
                    dTn = inner_bwd(ops, gd)
                    dy2 = self._bn_act_pad_bwd(ops, dTn, y2, sc.d2_bn, 1, Tn_pad)
                else:
                    dy2 = gd
                dX2 = self._conv_bwd(ops, sc.d2, d_2, X2, dy2)
                dy1 = self._bn_act_pad_bwd(ops, dX2, y1, sc.d1_bn, 1, pd)
                m1 = self._conv_meta(sc.d1, d_1, x_d1, dy1)
                ops.append(("mfvi_conv2d_wgrad", (C.byref(d_1), L.view(x_d1), L.view(dy1), self.dw.data_ptr() + 4 * sc.d1.w_off,
                                                  self._dbias_ptr(sc.d1, True), lay.P_pad), m1))
                if need_dT:
                    dT_d1 = self._interior(dT, Tpad - pd)
                    if Cs:
                        ops.append(("__join__", (), {"lane": "skip"}))      # dT holds the skip branch's contribution
                    elif Tpad != pd:
                        ops.append(("mfvi_fill_f32", (dT.data_ptr(), dT.numel(), 0.0), self._ew_meta(dT)))
                    ops.append(("mfvi_conv2d_dgrad", (C.byref(d_1), L.view(dy1), self.w.data_ptr() + 4 * sc.d1.w_off, lay.P_pad,
                                                      L.view(dT_d1), 1 if (Cs or Tpad != pd) else 0), m1))
                if self.mega_from is not None and i == self.mega_from:
                    self._fuse(ops, n_fuse0, len(ops), "bwd")
                return dT

            Tn_pad = in_pad(i + 1) if i < n - 1 else 0
            return z_out, z_out_bn, backward

        x_in = self.x0
        z0, z0_bn, bwd0 = run_scale(0, x_in, self.pad0)
        XF = self._bn_act_pad(z0, z0_bn, *self._bn_args(z0_bn), 1, 0)
        self.out, d_f = self._conv_fwd(lay.final, XF, None)
        self.dout = torch.zeros(self.out.shape[:3] + ((lay.final.cout + 3) // 4 * 4,), dtype=torch.float32,
                                device=self.device)[..., :lay.final.cout]
        ops = self.bwd_ops
        dXF = self._conv_bwd(ops, lay.final, d_f, XF, self.dout, bn_follows=False)
        dz0 = self._bn_act_pad_bwd(ops, dXF, z0, z0_bn, 1, 0)
        self.dx0 = bwd0(ops, dz0)

    # ---------------------------------------------------------------- execution
    @L.on_device
    def zero_accumulators(self):
        L.call("mfvi_fill_f32", self.zbuf.data_ptr(), self.zbuf.numel(), 0.0, meta={"bytes": 4.0 * self.zbuf.numel()})

    @L.on_device
    def set_input(self, x_nhwc: torch.Tensor, noise: Optional[torch.Tensor], std: float, key: L.PhiloxKey):
        """x0 = reflect_pad(saved + std * N(0,1))  (reference bayesian_optimization.py:1363-1364).
        x_nhwc: (H,W,C) contiguous.  noise (optional, (H,W,C)) injects the normals."""
        L.call("mfvi_input_jitter_pad", x_nhwc.data_ptr(), L.ptr(noise), self.H, self.W, self.spec.num_input_channels,
               float(std), self.pad0, key, L.view(self.x0))

    def ensure_eps(self):
        if self.eps is None:
            self.eps = torch.zeros(self.S, self.lay.P_pad, dtype=torch.float32, device=self.device)
        return self.eps

    @L.on_device
    def sample_weights(self, key: L.PhiloxKey):
        """w_s = mu + softplus(rho) * eps_s for every layer at once (reference module.py:82-85)."""
        inj = self.inject_eps
        L.call("mfvi_sample_weights", self.mu.data_ptr(), self.rho.data_ptr(), self.lay.P, self.S,
               self.eps.data_ptr() if inj else None, self.lay.P_pad, key, self.w.data_ptr(), self.lay.P_pad,
               meta={"bytes": 4.0 * self.lay.P * (2 + self.S)})
    @L.on_device
    def use_mean_weights(self):
        """Eval mode of RTLayer (reparam_layers.py:33-35): w = mu for every sample."""
        self.w[:, :self.lay.P].copy_(self.mu.unsqueeze(0).expand(self.S, -1))

    def _run(self, op_list):
        """Launch an op list: ops tagged lane="skip" go to the skip stream, weight-gradient kernels to the wgrad stream
        (each waits for the lane that produced its inputs), everything else to the current stream; every lane used is
        joined back before returning.  With the per-kernel timeline on, everything runs on one stream."""
        if self.plan_only:
            raise L.MfviError("this SkipEngine is plan only (device 'meta' or plan_only=True): it cannot execute")
        serial = L.timeline is not None
        main = torch.cuda.current_stream(self.device)
        streams = {"main": main, "wgrad": self._side, "skip": self._side2}
        dirty = set()

        def join(lane):
            ev = torch.cuda.Event()
            ev.record(streams[lane])
            main.wait_event(ev)
            dirty.discard(lane)

        for name, args, meta in op_list:
            lane = meta.get("lane", "main")
            if name == "mfvi_conv2d_wgrad" and self.overlap_wgrad:
                lane = "wgrad"
            if lane == "skip" and not self.overlap_skip:
                lane = "main"
            if name == "__join__":
                if meta["lane"] in dirty:
                    join(meta["lane"])
                continue
            if serial or lane == "main":
                after = meta.get("after", "main")
                if not serial and after in dirty:          # a main-stream op consuming a side lane's result
                    join(after)
                L.call(name, *args, meta=meta)
                continue
            after = meta.get("after", "main")
            if after != lane and (after == "main" or after in dirty):
                ev = torch.cuda.Event()
                ev.record(streams[after])
                streams[lane].wait_event(ev)
            L.call(name, *args, stream=streams[lane].cuda_stream, meta=meta)
            dirty.add(lane)
        for lane in list(dirty):
            join(lane)

    @L.on_device
    def forward(self):
        self._run(self.fwd_ops)

    @L.on_device
    def backward(self):
        """Consumes self.dout; fills dw[s], BN gamma/beta grads (and dx0 when requested)."""
        self._run(self.bwd_ops)

    @L.on_device
    def reparam_kl(self, key: L.PhiloxKey, *, prior_mu: float, prior_sigma_plus_eps: float, direction: int,
                   kscale: float, kscale_dev=None, data_term: bool = True, gscale: float = 1.0, accumulate: bool = False,
                   want_grad: bool = True, want_kl: bool = True):
        """grad_mu/rho (=|+=) gscale * reparam-chain(dw) + kscale * dKL ; arena[KL] += KL."""
        inj = self.inject_eps
        L.call("mfvi_kl_reparam_fwd_bwd", self.mu.data_ptr(), self.rho.data_ptr(), self.lay.P, float(prior_mu),
               float(prior_sigma_plus_eps), direction, float(kscale), L.ptr(kscale_dev),
               self.dw.data_ptr() if data_term else None, self.lay.P_pad, self.S if data_term else 0,
               self.eps.data_ptr() if inj else None, self.lay.P_pad, key, float(gscale),
               self._aptr(KL) if want_kl else None, self.g_mu.data_ptr() if want_grad else None,
               self.g_rho.data_ptr() if want_grad else None, 1 if accumulate else 0,
               meta={"bytes": 4.0 * self.lay.P * (2 + (self.S if data_term else 0) + (2 if want_grad else 0))})

    @L.on_device
    def update_running_stats(self, momentum: float = 0.1):
        L.call("mfvi_bn_running_update", self.arena.data_ptr(), self._bn_ch_off.data_ptr(), self._bn_sums_off.data_ptr(),
               self._bn_C.data_ptr(), self._bn_count.data_ptr(), len(self.lay.bns), self.S, float(momentum),
               self.running_mean.data_ptr(), self.running_var.data_ptr())

    def conv_dispatch_table(self) -> List[dict]:
        """One row per convolution launch of the plan (forward, data gradient, weight gradient): the kernel family
        libmfvidip would run for it and its tile plan, asked from the library's own dispatch code in planning-only mode
        (mfvi_conv2d_plan).  Works on a plan-only (device 'meta') engine, i.e. without a GPU."""
        rows = []
        for name, args, meta in self.fwd_ops + self.bwd_ops:
            if name == "mfvi_conv2d_fwd":
                d, x, _, b, wss, y, _ = args
                info = L.conv_plan(d._obj, L.PASS_FWD, x, y, wss, 0, b is not None)
            elif name == "mfvi_conv2d_dgrad":
                d, dy, _, wss, dx, acc = args
                info = L.conv_plan(d._obj, L.PASS_DGRAD, dy, dx, wss, acc)
            elif name == "mfvi_conv2d_wgrad":
                d, x, dy, _, db, wss = args
                info = L.conv_plan(d._obj, L.PASS_WGRAD, x, dy, wss, 0, db is not None)
            else:
                continue
            g = d._obj
            rows.append(dict(info, op=name[len("mfvi_conv2d_"):], layer=meta["layer"], shape=meta["shape"], S=g.S, Cin=g.Cin,
                             Cout=g.Cout, K=g.KH, stride=g.stride, Hin=g.Hin, Win=g.Win, Hout=g.Hout, Wout=g.Wout))
        return rows

    # ---------------------------------------------------------------- parameter views (reference shapes)
    def param_views(self, src: str = "theta") -> Dict[str, torch.Tensor]:
        """Reference-named, reference-shaped views into theta (src='theta') or grad (src='grad')."""
        mu, rho, gamma, beta = ((self.mu, self.rho, self.gamma, self.beta) if src == "theta"
                                else (self.g_mu, self.g_rho, self.g_gamma, self.g_beta))
        out = {}
        for c in self.lay.convs:
            for nm, flat in (("mu", mu), ("rho", rho)):
                out[f"{c.key}.W_{nm}"] = flat[c.w_off:c.w_off + c.w_numel].view(c.k, c.k, c.cout, c.cin).permute(2, 3, 0, 1)
                out[f"{c.key}.bias_{nm}"] = flat[c.b_off:c.b_off + c.cout]
        for b in self.lay.bns:
            out[f"{b.key}.weight"] = gamma[b.ch_off:b.ch_off + b.C]
            out[f"{b.key}.bias"] = beta[b.ch_off:b.ch_off + b.C]
            if src == "theta":
                out[f"{b.key}.running_mean"] = self.running_mean[b.ch_off:b.ch_off + b.C]
                out[f"{b.key}.running_var"] = self.running_var[b.ch_off:b.ch_off + b.C]
        return out

    @L.on_device
    def load_params(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        """Copy a reference state_dict (keys optionally prefixed, e.g. 'net.') into the flat buffers."""
        views = self.param_views()
        for k, v in views.items():
            src = sd.get(prefix + k)
            if src is None:
                raise KeyError(f"state dict lacks {prefix + k}")
            v.copy_(torch.as_tensor(src).to(self.device, torch.float32))

    @L.on_device
    def pack_eps(self, eps: Sequence[Dict[str, torch.Tensor]], prefix: str = ""):
        """Injected eps: per sample a dict '<convkey>.W' (Cout,Cin,k,k) / '<convkey>.b' (Cout,) -> storage layout."""
        E = self.ensure_eps()
        assert len(eps) == self.S
        for s, d in enumerate(eps):
            for c in self.lay.convs:
                ew = torch.as_tensor(d[prefix + c.key + ".W"]).to(self.device, torch.float32)
                E[s, c.w_off:c.w_off + c.w_numel].view(c.k, c.k, c.cout, c.cin).copy_(ew.permute(2, 3, 0, 1))
                E[s, c.b_off:c.b_off + c.cout].copy_(torch.as_tensor(d[prefix + c.key + ".b"]).to(self.device, torch.float32))
        self.inject_eps = True

    @L.on_device
    def out_nchw(self) -> torch.Tensor:
        S, H, W, Cn = self.out.shape
        o = torch.empty(S, Cn, H, W, dtype=torch.float32, device=self.device)
        L.call("mfvi_nhwc_to_nchw", self.out.contiguous().data_ptr(), o.data_ptr(), S, Cn, H, W)
        return o
