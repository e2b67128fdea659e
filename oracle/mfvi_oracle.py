"""ORACLE — test infrastructure only.  CPU restatement (PyTorch CPU / float32 or float64) of
the MFVI-DIP training step of Cardio-AI/mfvi-dip-mia.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product
package never does.

Every function cites the reference file:line it follows (paths relative to the reference root).
The arithmetic primitives (conv2d, batch-norm statistics, bilinear interpolation) live in the
reference's third-party dependency PyTorch (environment.yml pins torch==1.9.0); they are restated
here through the same primitive calls or, for the radon projector and the KL, as explicit formulas.

Parity pin: the reference ships no tests or golden vectors, so this restatement is
pinned against outputs of the *imported reference itself*, generated in the build container by
`tests/golden/make_golden.py` and committed as `tests/golden/*.npz` (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# a2  sampling + layer function
# --------------------------------------------------------------------------------------
def softplus(rho: torch.Tensor) -> torch.Tensor:
    """sigma = softplus(rho), beta=1, threshold=20 (torch default; BayTorch/modules/module.py:3,70)."""
    return F.softplus(rho)


def rsample(mu: torch.Tensor, sigma: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """w = mu + eps * sigma  (BayTorch/modules/module.py:82-85, eps injected instead of randn_like)."""
    return mu + eps * sigma


def conv2d_rt(x, W_mu, W_rho, bias_mu, bias_rho, eps_w, eps_b, stride=1, padding=0, training=True):
    """RTLayer.forward with layer_fn = conv2d (BayTorch/modules/reparam_layers.py:26-37, conv.py:6-38)."""
    if training:
        w = rsample(W_mu, softplus(W_rho), eps_w)
        b = rsample(bias_mu, softplus(bias_rho), eps_b) if bias_mu is not None else None
    else:
        w, b = W_mu, bias_mu
    return F.conv2d(x, w, b, stride=stride, padding=padding, dilation=1, groups=1)


def linear_rt(x, W_mu, W_rho, bias_mu, bias_rho, eps_w, eps_b, training=True):
    """RTLayer.forward with layer_fn = linear (BayTorch/modules/linear.py:5-27)."""
    if training:
        w = rsample(W_mu, softplus(W_rho), eps_w)
        b = rsample(bias_mu, softplus(bias_rho), eps_b) if bias_mu is not None else None
    else:
        w, b = W_mu, bias_mu
    return F.linear(x, w, b)


def conv2d_lrt(x, W_mu, W_rho, bias_mu, bias_rho, eps, stride=1, padding=0, training=True):
    """LRTLayer.forward with layer_fn = conv2d (BayTorch/modules/reparam_layers.py:59-72, conv.py:75-107): the OUTPUT is
    sampled, act_mu + sqrt(1e-16 + conv(x^2, softplus(W_rho)^2, softplus(bias_rho)^2)) * eps, eps of the output's shape."""
    act_mu = F.conv2d(x, W_mu, bias_mu, stride=stride, padding=padding)
    if not training:
        return act_mu
    bias_var = softplus(bias_rho) ** 2 if bias_mu is not None else None
    act_std = torch.sqrt(1e-16 + F.conv2d(x ** 2, softplus(W_rho) ** 2, bias_var, stride=stride, padding=padding))
    return rsample(act_mu, act_std, eps)


def linear_lrt(x, W_mu, W_rho, bias_mu, bias_rho, eps, training=True):
    """LRTLayer.forward with layer_fn = linear (BayTorch/modules/linear.py:29-50)."""
    act_mu = F.linear(x, W_mu, bias_mu)
    if not training:
        return act_mu
    bias_var = softplus(bias_rho) ** 2 if bias_mu is not None else None
    act_std = torch.sqrt(1e-16 + F.linear(x ** 2, softplus(W_rho) ** 2, bias_var))
    return rsample(act_mu, act_std, eps)


# --------------------------------------------------------------------------------------
# a7  KL
# --------------------------------------------------------------------------------------
def prior_scale(temp: float, sigma: float) -> float:
    """prior sigma handed to VIModule = sqrt(temp)*sigma (bayesian_optimization.py:1335-1336);
    VIModule adds 1e-6 (BayTorch/modules/module.py:38)."""
    return math.sqrt(temp) * sigma + 1e-6


def kl_elementwise(mu, rho, prior_mu: float, prior_sigma_plus_eps: float, kl_type: str = "reverse"):
    """Per-element KL of VIModule._kl (BayTorch/modules/module.py:64-80).
    kl_type='reverse' (the default, never overridden) calls kl_divergence(prior, posterior):
        KL(N(m_p,s_p) || N(mu,sig)) = 0.5*(r + t - 1 - log r), r=(s_p/sig)^2, t=((m_p-mu)/sig)^2
    (torch.distributions.kl._kl_normal_normal).  Any other kl_type gives KL(posterior || prior)."""
    sig = softplus(rho)
    sp = torch.as_tensor(prior_sigma_plus_eps, dtype=mu.dtype)
    mp = torch.as_tensor(prior_mu, dtype=mu.dtype)
    if kl_type == "reverse":
        r = (sp / sig) ** 2
        t = ((mp - mu) / sig) ** 2
    else:
        r = (sig / sp) ** 2
        t = ((mu - mp) / sp) ** 2
    return 0.5 * (r + t - 1.0 - torch.log(r))


def kl_layers(params: Sequence[torch.Tensor], prior_mu, prior_sigma_plus_eps, kl_type="reverse"):
    """MeanFieldVI.kl (BayTorch/freq_to_bayes.py:43-48): sum over all layers of W and bias KL.
    `params` = flat sequence (mu0, rho0, mu1, rho1, ...). Returns shape-[1] tensor like the reference."""
    total = torch.zeros(1, dtype=params[0].dtype, device=params[0].device)
    for mu, rho in zip(params[0::2], params[1::2]):
        total = total + kl_elementwise(mu, rho, prior_mu, prior_sigma_plus_eps, kl_type).sum()
    return total


# --------------------------------------------------------------------------------------
# a6  losses
# --------------------------------------------------------------------------------------
def gaussian_nll(mu, neg_logvar, target, reduction="mean"):
    """utils/bayesian_utils.py:29-32."""
    s = torch.clamp(neg_logvar, min=-20, max=20)
    loss = torch.exp(s) * (target - mu) ** 2 - s
    return loss.mean() if reduction == "mean" else loss.sum()


def gaussian_nll_inpainting(mu, neg_logvar, target, mask, reduction="mean"):
    """utils/bayesian_utils.py:35-39 (mask multiplies the loss; mean over ALL elements)."""
    s = torch.clamp(neg_logvar, min=-20, max=20)
    loss = (torch.exp(s) * (target - mu) ** 2 - s) * mask
    return loss.mean() if reduction == "mean" else loss.sum()


def sr_downsample_nearest(x, factor: int):
    """F.interpolate(scale_factor=1/factor, mode='nearest', recompute_scale_factor=False)
    (bayesian_optimization.py:2095-2099) == pixels 0, factor, 2*factor, ..."""
    return x[..., ::factor, ::factor]


# --------------------------------------------------------------------------------------
# a5  skip-net elementwise pieces
# --------------------------------------------------------------------------------------
def bn_train(x, weight, bias, eps=1e-5):
    """nn.BatchNorm2d in training mode (models/common.py:96-97): batch statistics over (N,H,W),
    biased variance.  The runners always call it with N=1."""
    m = x.mean(dim=(0, 2, 3), keepdim=True)
    v = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    return (x - m) / torch.sqrt(v + eps) * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def lrelu(x):
    """nn.LeakyReLU(0.2) (models/common.py:83)."""
    return F.leaky_relu(x, 0.2)


def reflect_pad(x, p: int):
    """nn.ReflectionPad2d(p) (models/common.py:117-121)."""
    return F.pad(x, (p, p, p, p), mode="reflect") if p > 0 else x


def upsample2x(x, mode: str):
    """nn.Upsample(scale_factor=2, mode=...) (models/skip.py:102); align_corners=False default."""
    if mode == "bilinear":
        return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return F.interpolate(x, scale_factor=2, mode="nearest")


def center_crop_cat(tensors: List[torch.Tensor]):
    """Concat.forward (models/common.py:23-43): centre-crop to min H/W, cat on channels."""
    h = min(t.shape[2] for t in tensors)
    w = min(t.shape[3] for t in tensors)
    out = []
    for t in tensors:
        d2 = (t.shape[2] - h) // 2
        d3 = (t.shape[3] - w) // 2
        out.append(t[:, :, d2:d2 + h, d3:d3 + w])
    return torch.cat(out, dim=1)


# --------------------------------------------------------------------------------------
# skip-net structure (models/skip.py:5-134) as data
# --------------------------------------------------------------------------------------
@dataclass
class SkipCfg:
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
class ConvRef:
    key: str            # state-dict prefix of the Conv2dRT ('...Conv2d_up_9')
    cin: int
    cout: int
    k: int
    stride: int


@dataclass
class ScaleRef:
    skip_conv: Optional[ConvRef]
    skip_bn: Optional[str]
    d1: ConvRef
    d1_bn: str
    d2: ConvRef
    d2_bn: str
    cat_bn: str
    up: ConvRef
    up_bn: str
    up1: Optional[ConvRef]
    up1_bn: Optional[str]


@dataclass
class SkipLayout:
    scales: List[ScaleRef] = field(default_factory=list)
    final: ConvRef = None

    def convs_in_exec_order(self) -> List[ConvRef]:
        """Order in which RTLayer.forward draws eps (skip branch of a Concat first)."""
        pre, post = [], []
        for sc in self.scales:
            if sc.skip_conv is not None:
                pre.append(sc.skip_conv)
            pre += [sc.d1, sc.d2]
        for sc in reversed(self.scales):
            post.append(sc.up)
            if sc.up1 is not None:
                post.append(sc.up1)
        return pre + post + [self.final]


def skip_layout(cfg: SkipCfg, root: str = "net.") -> SkipLayout:
    """State-dict key layout produced by skip() + rename_modules (models/skip.py:56-132,
    utils/common_utils.py:248-262).  Verified against the imported reference's state_dict keys
    in tests/test_oracle_golden.py."""
    n = len(cfg.down)
    it_skip, it_deep = 1, 1
    it_up = 2 * n if cfg.need1x1_up else n
    lay = SkipLayout()
    P = root
    cin = cfg.num_input_channels
    for i in range(n):
        k = it_up - 1
        has_skip = cfg.skip[i] != 0
        if has_skip:
            D = f"{P}Concat_up_{k}.1."
            skip_conv = ConvRef(f"{P}Concat_up_{k}.0.Sequential_skip_{it_skip}.Conv2d_skip_{it_skip}",
                                cin, cfg.skip[i], cfg.filter_skip, 1)
            skip_bn = f"{P}Concat_up_{k}.0.BatchNorm2d_skip_{it_skip}"
            it_skip += 1
            up_seq = f"Sequential_up_{k}"
        else:
            D = f"{P}Sequential_up_{k}."
            skip_conv, skip_bn = None, None
            up_seq = f"Sequential_up_{k}_1"
        d1 = ConvRef(f"{D}Sequential_deeper_{it_deep}.Conv2d_deeper_{it_deep}", cin, cfg.down[i], cfg.filter_down, 2)
        d1_bn = f"{D}BatchNorm2d_deeper_{it_deep}"
        it_deep += 1
        d2 = ConvRef(f"{D}Sequential_deeper_{it_deep}.Conv2d_deeper_{it_deep}", cfg.down[i], cfg.down[i], cfg.filter_down, 1)
        d2_bn = f"{D}BatchNorm2d_deeper_{it_deep}"
        it_deep += 1
        kk = cfg.up[i + 1] if i < n - 1 else cfg.down[i]
        cat_bn = f"{P}BatchNorm2d_up_{k}"
        up = ConvRef(f"{P}{up_seq}.Conv2d_up_{k}", cfg.skip[i] + kk, cfg.up[i], cfg.filter_up, 1)
        up_bn = f"{P}BatchNorm2d_up_{k}_1"
        if cfg.need1x1_up:
            up1 = ConvRef(f"{P}Sequential_up_{k + 1}.Conv2d_up_{k + 1}", cfg.up[i], cfg.up[i], 1, 1)
            up1_bn = f"{P}BatchNorm2d_up_{k + 1}"
            it_up -= 1
        else:
            up1, up1_bn = None, None
        it_up -= 1
        lay.scales.append(ScaleRef(skip_conv, skip_bn, d1, d1_bn, d2, d2_bn, cat_bn, up, up_bn, up1, up1_bn))
        cin = cfg.down[i]
        P = D + "7."
    n_children = (3 + 2) + (3 if cfg.need1x1_up else 0)
    final_iter = 2 * n + 1 if cfg.need1x1_up else n + 1
    lay.final = ConvRef(f"{root}{n_children + 1}.Conv2d_up_{final_iter}", cfg.up[0], cfg.num_output_channels, 1, 1)
    return lay


def skip_forward(sd: Dict[str, torch.Tensor], cfg: SkipCfg, x: torch.Tensor,
                 eps: Dict[str, torch.Tensor], root: str = "net.") -> torch.Tensor:
    """Forward of the hour-glass net for ONE MC sample, N=1 (models/skip.py:58-132 as executed by
    MeanFieldVI.forward, BayTorch/freq_to_bayes.py:40-41).  `sd` maps reference state-dict keys to
    tensors (leaf tensors requiring grad for the backward oracle); `eps` maps '<convkey>.W' /
    '<convkey>.b' to the injected standard normals."""
    lay = skip_layout(cfg, root)

    def conv(c: ConvRef, t):
        p = (c.k - 1) // 2
        t = reflect_pad(t, p)                                   # models/common.py:117-121
        return conv2d_rt(t, sd[c.key + ".W_mu"], sd[c.key + ".W_rho"], sd[c.key + ".bias_mu"],
                         sd[c.key + ".bias_rho"], eps[c.key + ".W"], eps[c.key + ".b"], stride=c.stride)

    def bn(key, t):
        return bn_train(t, sd[key + ".weight"], sd[key + ".bias"])

    def scale(i, t):
        sc = lay.scales[i]
        branches = []
        if sc.skip_conv is not None:                            # models/skip.py:70-75
            branches.append(lrelu(bn(sc.skip_bn, conv(sc.skip_conv, t))))
        d = lrelu(bn(sc.d1_bn, conv(sc.d1, t)))                 # models/skip.py:77-82
        d = lrelu(bn(sc.d2_bn, conv(sc.d2, d)))                 # models/skip.py:86-89
        if i < len(lay.scales) - 1:
            d = scale(i + 1, d)                                 # models/skip.py:99
        d = upsample2x(d, cfg.upsample_mode)                    # models/skip.py:102
        branches.append(d)
        u = center_crop_cat(branches) if len(branches) > 1 else d   # models/skip.py:63-66
        u = bn(sc.cat_bn, u)                                    # models/skip.py:68
        u = lrelu(bn(sc.up_bn, conv(sc.up, u)))                 # models/skip.py:104-108
        if sc.up1 is not None:
            u = lrelu(bn(sc.up1_bn, conv(sc.up1, u)))           # models/skip.py:112-117
        return u

    out = conv(lay.final, scale(0, x))                          # models/skip.py:129-130
    if cfg.need_sigmoid:
        out = torch.sigmoid(out)
    return out


def vi_param_pairs(sd: Dict[str, torch.Tensor], cfg: SkipCfg, root: str = "net."):
    """(mu, rho) pairs of every converted layer, W then bias (module.py:70-72)."""
    out = []
    for c in skip_layout(cfg, root).convs_in_exec_order():
        out += [sd[c.key + ".W_mu"], sd[c.key + ".W_rho"], sd[c.key + ".bias_mu"], sd[c.key + ".bias_rho"]]
    return out


# --------------------------------------------------------------------------------------
# a10  radon
# --------------------------------------------------------------------------------------
def radon_forward(image: torch.Tensor, theta_deg: torch.Tensor) -> torch.Tensor:
    """FastRadonTransform.forward (radon/radon.py:32-55) written out explicitly:
    affine_grid(align_corners=False) base coords x_j=(2j+1)/W-1, y_i=(2i+1)/H-1;
    grid = (cos*x - sin*y, sin*x + cos*y); grid_sample bilinear / zeros / align_corners=False;
    sum over rows i.  image (1,C,H,W) -> (1,C,T,W)."""
    assert image.shape[0] == 1 and image.shape[2] == image.shape[3]
    _, C, H, W = image.shape
    dt = image.dtype
    th = torch.deg2rad(theta_deg.to(device=image.device, dtype=dt))
    ts, tc = torch.sin(th), torch.cos(th)
    xs = (2 * torch.arange(W, dtype=dt, device=image.device) + 1) / W - 1
    ys = (2 * torch.arange(H, dtype=dt, device=image.device) + 1) / H - 1
    gx = tc[:, None, None] * xs[None, None, :] - ts[:, None, None] * ys[None, :, None]   # (T,H,W)
    gy = ts[:, None, None] * xs[None, None, :] + tc[:, None, None] * ys[None, :, None]
    ix = ((gx + 1) * W - 1) / 2
    iy = ((gy + 1) * H - 1) / 2
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fx = ix - x0
    fy = iy - y0
    img = image[0]                                               # (C,H,W)

    def tap(yy, xx, wgt):
        valid = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        xi = xx.clamp(0, W - 1).long()
        yi = yy.clamp(0, H - 1).long()
        v = img[:, yi, xi]                                       # (C,T,H,W)
        return v * (wgt * valid.to(dt))[None]

    s = (tap(y0, x0, (1 - fx) * (1 - fy)) + tap(y0, x0 + 1, fx * (1 - fy))
         + tap(y0 + 1, x0, (1 - fx) * fy) + tap(y0 + 1, x0 + 1, fx * fy))
    return s.sum(dim=2)[None]                                    # (1,C,T,W)


# --------------------------------------------------------------------------------------
# a8  the training step (loss side)
# --------------------------------------------------------------------------------------
def mfvi_loss(sd, cfg: SkipCfg, net_input, eps_per_sample: List[Dict[str, torch.Tensor]], *,
              task: str, temp: float, prior_sigma_plus_eps: float, target=None, mask=None,
              sr_factor: int = 4, theta_deg=None, sino=None, root="net."):
    """loss = mean_s NLL_s + temp*KL  (bayesian_optimization.py:1366-1370 and task variants
    :2182-2188, :3033-3038, :576-578), with the MC>1 restatement of SURVEY §4: S sequential
    forwards on the same net_input, eps_s injected, NLL averaged over samples.
    Returns (loss[1], nll_mean, kl[1], outs list)."""
    outs, nlls = [], []
    for eps in eps_per_sample:
        out = skip_forward(sd, cfg, net_input, eps, root)
        if task == "den":
            nll = gaussian_nll(out[:, :1], out[:, 1:], target)
        elif task == "sr":
            lr = sr_downsample_nearest(out, sr_factor)
            nll = gaussian_nll(lr[:, :1], lr[:, 1:], target)
        elif task == "inp":
            nll = gaussian_nll_inpainting(torch.sigmoid(out[:, :3]), out[:, 3:], target, mask)
        elif task == "ct":
            nll = F.mse_loss(radon_forward(out, theta_deg), sino)
        else:
            raise ValueError(task)
        outs.append(out)
        nlls.append(nll)
    nll_mean = torch.stack(nlls).mean()
    kl = kl_layers(vi_param_pairs(sd, cfg, root), 0.0, prior_sigma_plus_eps)
    loss = nll_mean + temp * kl
    return loss, nll_mean, kl, outs


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.AdamW single-tensor update (bayesian_optimization.py:1357,1372; wd=0)."""
    p = p * (1 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


# --------------------------------------------------------------------------------------
# metrics ("next" row f1, used for trajectory parity)
# --------------------------------------------------------------------------------------
def psnr(image_true, image_test):
    """utils/common_utils.py:297-305."""
    err = F.mse_loss(image_true, image_test)
    return float(10 * torch.log10(1 / err))


def ssim(image_true, image_test, window_size=11, sigma=1.5):
    """utils/common_utils.py:308-353 (11x11 Gaussian window, zero padding, C1=0.01^2, C2=0.03^2)."""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = g / g.sum()
    ch = image_true.shape[1]
    w2 = (g[:, None] @ g[None, :]).float()[None, None].expand(ch, 1, window_size, window_size).contiguous()
    w2 = w2.to(image_true.dtype)
    p = window_size // 2
    mu1 = F.conv2d(image_true, w2, padding=p, groups=ch)
    mu2 = F.conv2d(image_test, w2, padding=p, groups=ch)
    s1 = F.conv2d(image_true * image_true, w2, padding=p, groups=ch) - mu1 ** 2
    s2 = F.conv2d(image_test * image_test, w2, padding=p, groups=ch) - mu2 ** 2
    s12 = F.conv2d(image_true * image_test, w2, padding=p, groups=ch) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))
    return float(m.mean())


def uce(errors, uncert, n_bins=15):
    """utils/uce.py:9-40 (bin edges from uncert.min()..max(), bins (lo, hi])."""
    edges = torch.linspace(float(uncert.min()), float(uncert.max()), n_bins + 1)
    total = torch.zeros(1)
    for lo, hi in zip(edges[:-1], edges[1:]):
        in_bin = uncert.gt(float(lo)) * uncert.le(float(hi))
        prop = in_bin.float().mean()
        if float(prop) > 0.0:
            total += torch.abs(uncert[in_bin].mean() - errors[in_bin].float().mean()) * prop
    return float(total)
