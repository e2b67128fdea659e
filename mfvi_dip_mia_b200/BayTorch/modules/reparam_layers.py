"""RTLayer / LRTLayer: weight-space and local (output-space) reparameterisation layers (reference
BayTorch/modules/reparam_layers.py:6-37 and :39-72)."""
from ... import _lib as L
from ... import functional as Fn
from .module import VIModule


class RTLayer(VIModule):
    def __init__(self, layer_fn, weight_size, bias_size=None, prior=None, posteriors=None, kl_type="reverse",
                 _version="old", **kwargs):
        super().__init__(layer_fn=layer_fn, weight_size=weight_size, bias_size=bias_size, prior=prior,
                         posteriors=posteriors, kl_type=kl_type)
        self.kwargs = kwargs
        self.math = L.MATH_FP32
        self._injected_eps = None      # (eps_w, eps_b) consumed by the next forward (parity tests)

    def inject_eps(self, eps_w, eps_b=None):
        self._injected_eps = (eps_w, eps_b)

    def _draw_eps(self):
        if self._injected_eps is not None:
            ew, eb = self._injected_eps
            self._injected_eps = None
            dev = self.W_mu.device
            return ew.to(dev), (eb.to(dev) if eb is not None else None)
        ew = Fn.fresh_eps_like(self.W_mu)
        eb = Fn.fresh_eps_like(self.bias_mu) if self.bias_mu is not None else None
        return ew, eb

    def forward(self, x):
        ew, eb = self._draw_eps() if self.training else (None, None)
        return self.layer_fn(self, x, ew, eb)


class LRTLayer(VIModule):
    """Local reparameterisation: the layer output is sampled, act_mu + sqrt(1e-16 + act_var) * eps with eps of the
    output's shape (training); the mean activation in eval mode."""

    def __init__(self, layer_fn, weight_size, bias_size=None, prior=None, posteriors=None, kl_type="reverse", **kwargs):
        super().__init__(layer_fn=layer_fn, weight_size=weight_size, bias_size=bias_size, prior=prior,
                         posteriors=posteriors, kl_type=kl_type)
        self.kwargs = kwargs
        self.math = L.MATH_FP32
        self._injected_eps = None      # output-shaped eps consumed by the next training forward (parity tests)

    def inject_eps(self, eps):
        self._injected_eps = eps

    def forward(self, x):
        eps, self._injected_eps = self._injected_eps, None
        return self.layer_fn(self, x, eps)
