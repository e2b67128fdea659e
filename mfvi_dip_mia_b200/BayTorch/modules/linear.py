"""LinearRT (reference BayTorch/modules/linear.py:5-27) and LinearLRT (:29-50): y = x W^T + b with one weight sample per forward.
Executed as a 1x1 sampled-weight convolution whose "pixels" are the batch rows."""
from ... import functional as Fn
from .reparam_layers import LRTLayer, RTLayer


def _linear(layer, x, eps_w, eps_b):
    lead = x.shape[:-1]
    rows = x.reshape(-1, x.shape[-1])
    # (rows, in) -> NCHW (1, in, rows, 1)
    xi = rows.t().reshape(1, rows.shape[1], rows.shape[0], 1)
    W_mu = layer.W_mu[:, :, None, None]
    W_rho = layer.W_rho[:, :, None, None]
    ew = eps_w[:, :, None, None] if eps_w is not None else None
    y = Fn.SampledConv2dFn.apply(xi, W_mu, W_rho, layer.bias_mu, layer.bias_rho, ew, eps_b, 1, 0, layer.training,
                                 layer.math)
    return y.reshape(y.shape[1], rows.shape[0]).t().reshape(*lead, y.shape[1])


class LinearRT(RTLayer):
    def __init__(self, in_features, out_features, bias=True, prior=None, posteriors=None, kl_type="reverse"):
        self.in_features = in_features
        self.out_featurs = out_features      # (sic) attribute name of the reference, linear.py:17
        self.out_features = out_features
        weight_size = (out_features, in_features)
        bias_size = (out_features) if bias else None
        super().__init__(layer_fn=_linear, weight_size=weight_size, bias_size=bias_size, prior=prior,
                         posteriors=posteriors, kl_type=kl_type)


def _linear_lrt(layer, x, eps):
    lead = x.shape[:-1]
    rows = x.reshape(-1, x.shape[-1])
    xi = rows.t().reshape(1, rows.shape[1], rows.shape[0], 1)             # (rows, in) -> NCHW (1, in, rows, 1)
    if layer.training and eps is None:
        eps = Fn.fresh_eps_like(x.new_empty(rows.shape[0], layer.out_features))
    ei = eps.reshape(rows.shape[0], -1).t().reshape(1, -1, rows.shape[0], 1) if eps is not None else None
    y = Fn.LrtConv2dFn.apply(xi, layer.W_mu[:, :, None, None], layer.W_rho[:, :, None, None], layer.bias_mu,
                             layer.bias_rho, ei, 1, 0, layer.training, layer.math)
    return y.reshape(y.shape[1], rows.shape[0]).t().reshape(*lead, y.shape[1])


class LinearLRT(LRTLayer):
    def __init__(self, in_features, out_features, bias=True, prior=None, posteriors=None, kl_type="reverse"):
        self.in_features = in_features
        self.out_featurs = out_features      # (sic) attribute name of the reference, linear.py:40
        self.out_features = out_features
        weight_size = (out_features, in_features)
        bias_size = (out_features) if bias else None
        super().__init__(layer_fn=_linear_lrt, weight_size=weight_size, bias_size=bias_size, prior=prior,
                         posteriors=posteriors, kl_type=kl_type)
