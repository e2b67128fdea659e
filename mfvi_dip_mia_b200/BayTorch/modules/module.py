"""VIModule: parameters, prior and KL of one mean-field Gaussian layer (reference BayTorch/modules/module.py:9-85).
The KL is evaluated by the library's flat warp-shuffle kernel instead of torch.distributions."""
import torch
from torch.nn import Module, Parameter

from ... import functional as Fn


class VIModule(Module):
    def __init__(self, layer_fn, weight_size, bias_size=None, prior=None, posteriors=None, kl_type="reverse"):
        super().__init__()
        self.layer_fn = layer_fn
        if prior is None:
            prior = {"mu": 0, "sigma": 0.1}
        if posteriors is None:
            posteriors = {"mu": (0, 0.1), "rho": (-3.0, 0.1)}
        if "pi" in prior:
            raise NotImplementedError("scale-mixture priors ('pi') are outside the MFVI-DIP hot path "
                                      "(no runner config sets them; reference BayTorch/distributions)")
        # reference: Normal(mu, sigma + 1e-6); the scale is kept in double like the reference's 0-dim tensor
        self.prior = {"mu": float(prior["mu"]), "sigma": float(prior["sigma"])}
        self.prior_loc = float(prior["mu"])
        self.prior_scale = float(prior["sigma"]) + 1e-6
        self.kl_type = kl_type
        self.posterior_mu_initial = posteriors["mu"]
        self.posterior_rho_initial = posteriors["rho"]
        self.W_mu = Parameter(torch.empty(weight_size))
        self.W_rho = Parameter(torch.empty(weight_size))
        if bias_size is not None:
            self.bias_mu = Parameter(torch.empty(bias_size))
            self.bias_rho = Parameter(torch.empty(bias_size))
        else:
            self.register_parameter("bias_mu", None)
            self.register_parameter("bias_rho", None)
        self.reset_parameters()

    def reset_parameters(self):
        self.W_mu.data.normal_(*self.posterior_mu_initial)
        self.W_rho.data.normal_(*self.posterior_rho_initial)
        if self.bias_mu is not None:
            self.bias_mu.data.normal_(*self.posterior_mu_initial)
            self.bias_rho.data.normal_(*self.posterior_rho_initial)

    @property
    def _kl(self):
        direction = 0 if self.kl_type == "reverse" else 1
        kl = Fn.KlFn.apply(self.W_mu, self.W_rho, self.prior_loc, self.prior_scale, direction)
        if self.bias_mu is not None:
            kl = kl + Fn.KlFn.apply(self.bias_mu, self.bias_rho, self.prior_loc, self.prior_scale, direction)
        return kl

    @staticmethod
    def rsample(mu, sigma):
        """mu + eps*sigma with eps from the library's Philox stream (kept for API compatibility; the layers
        themselves sample inside the fused kernels)."""
        eps = Fn.fresh_eps_like(mu)
        return mu + eps * sigma
