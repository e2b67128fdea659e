"""VIModule: parameters, prior and KL of one mean-field Gaussian layer (reference BayTorch/modules/module.py:9-85).
The KL is evaluated by the library's flat warp-shuffle kernel instead of torch.distributions."""
import torch
from torch.nn import Module, Parameter

from ... import functional as Fn

_DEFAULT_PRIOR = {"mu": 0, "sigma": 0.1}
_DEFAULT_POSTERIORS = {"mu": (0, 0.1), "rho": (-3.0, 0.1)}


class VIModule(Module):
    """Holds (W_mu, W_rho[, bias_mu, bias_rho]); sigma = softplus(rho).  Initial values are drawn from torch's global RNG
    in the reference's order — W_mu, W_rho, bias_mu, bias_rho — so equal seeds give equal initialisations."""

    def __init__(self, layer_fn, weight_size, bias_size=None, prior=None, posteriors=None, kl_type="reverse"):
        super().__init__()
        prior = dict(_DEFAULT_PRIOR if prior is None else prior)
        posteriors = dict(_DEFAULT_POSTERIORS if posteriors is None else posteriors)
        if "pi" in prior:
            raise NotImplementedError("scale-mixture priors ('pi') are outside the MFVI-DIP hot path "
                                      "(no runner config sets them; reference BayTorch/distributions)")
        self.layer_fn, self.kl_type = layer_fn, kl_type
        # reference: Normal(mu, sigma + 1e-6); kept as Python floats (double), like the reference's 0-dim tensor
        self.prior = {k: float(prior[k]) for k in ("mu", "sigma")}
        self.prior_loc, self.prior_scale = self.prior["mu"], self.prior["sigma"] + 1e-6
        self.posterior_mu_initial, self.posterior_rho_initial = posteriors["mu"], posteriors["rho"]
        for stem, size in (("W", weight_size), ("bias", bias_size)):
            for kind in ("mu", "rho"):
                if size is None:
                    self.register_parameter(f"{stem}_{kind}", None)
                else:
                    setattr(self, f"{stem}_{kind}", Parameter(torch.empty(size)))
        self.reset_parameters()

    def _pairs(self):
        """(mu, rho) of the weight and, when present, of the bias."""
        yield self.W_mu, self.W_rho
        if self.bias_mu is not None:
            yield self.bias_mu, self.bias_rho

    def reset_parameters(self):
        for mu, rho in self._pairs():
            mu.data.normal_(*self.posterior_mu_initial)
            rho.data.normal_(*self.posterior_rho_initial)

    @property
    def _kl(self):
        direction = 0 if self.kl_type == "reverse" else 1
        return sum(Fn.KlFn.apply(mu, rho, self.prior_loc, self.prior_scale, direction) for mu, rho in self._pairs())

    @staticmethod
    def rsample(mu, sigma):
        """mu + eps*sigma with eps from the library's Philox stream (kept for API compatibility; the layers
        themselves sample inside the fused kernels)."""
        return mu + Fn.fresh_eps_like(mu) * sigma
