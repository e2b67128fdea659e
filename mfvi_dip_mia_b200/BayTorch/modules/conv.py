"""Conv2dRT (reference BayTorch/modules/conv.py:6-38) and Conv2dLRT (:75-107)."""
import torch
from torch.nn.modules.utils import _pair

from ... import functional as Fn
from .reparam_layers import LRTLayer, RTLayer


def _single(v, what):
    a, b = _pair(v)
    if a != b:
        raise NotImplementedError(f"Conv2dRT: anisotropic {what}={v} is not supported by the sm_100a kernels")
    return int(a)


def _conv2d(layer, x, eps_w, eps_b):
    kw = layer.kwargs
    if _single(kw.get("dilation", 1), "dilation") != 1 or kw.get("groups", 1) != 1:
        raise NotImplementedError("Conv2dRT: dilation/groups != 1 are outside the MFVI-DIP hot path (no runner uses them)")
    return Fn.SampledConv2dFn.apply(x, layer.W_mu, layer.W_rho, layer.bias_mu, layer.bias_rho, eps_w, eps_b,
                                    _single(kw.get("stride", 1), "stride"), _single(kw.get("padding", 0), "padding"),
                                    layer.training, layer.math)


class Conv2dRT(RTLayer):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, stride=1, padding=0, dilation=1, groups=1,
                 prior=None, posteriors=None, kl_type="reverse"):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        weight_size = (out_channels, in_channels, self.kernel_size[0], self.kernel_size[1])
        bias_size = (out_channels) if bias else None
        super().__init__(layer_fn=_conv2d, weight_size=weight_size, bias_size=bias_size, prior=prior,
                         posteriors=posteriors, kl_type=kl_type, stride=stride, padding=padding, dilation=dilation,
                         groups=groups)


def _out_shape(x, cout, k, stride, padding):
    return (x.shape[0], cout, (x.shape[2] + 2 * padding - k[0]) // stride + 1, (x.shape[3] + 2 * padding - k[1]) // stride + 1)


def _conv2d_lrt(layer, x, eps):
    kw = layer.kwargs
    if _single(kw.get("dilation", 1), "dilation") != 1 or kw.get("groups", 1) != 1:
        raise NotImplementedError("Conv2dLRT: dilation/groups != 1 are outside the MFVI-DIP path")
    stride, padding = _single(kw.get("stride", 1), "stride"), _single(kw.get("padding", 0), "padding")
    if layer.training and eps is None:
        shape = _out_shape(x, layer.out_channels, layer.kernel_size, stride, padding)
        eps = Fn.fresh_eps_like(torch.empty(shape, device=x.device))
    return Fn.LrtConv2dFn.apply(x, layer.W_mu, layer.W_rho, layer.bias_mu, layer.bias_rho, eps, stride, padding,
                                layer.training, layer.math)


class Conv2dLRT(LRTLayer):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, stride=1, padding=0, dilation=1, groups=1,
                 prior=None, posteriors=None, kl_type="reverse"):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        weight_size = (out_channels, in_channels, self.kernel_size[0], self.kernel_size[1])
        bias_size = (out_channels) if bias else None
        super().__init__(layer_fn=_conv2d_lrt, weight_size=weight_size, bias_size=bias_size, prior=prior,
                         posteriors=posteriors, kl_type=kl_type, stride=stride, padding=padding, dilation=dilation,
                         groups=groups)
