"""Generates tests/golden/lrt_layers.npz: Conv2dLRT / LinearLRT (local reparameterisation, reference
BayTorch/modules/reparam_layers.py:39-72, conv.py:75-107, linear.py:29-50) forward / backward run with the IMPORTED
REFERENCE, the output-space eps injected through VIModule.rsample.  Build container only.

    python tests/golden/make_lrt_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (imports the reference)

from BayTorch.modules import Conv2dLRT, LinearLRT  # noqa: E402


def fill(layer, g):
    with torch.no_grad():
        layer.W_mu.copy_(0.1 * torch.randn(layer.W_mu.shape, generator=g))
        layer.W_rho.copy_(-3 + 0.5 * torch.randn(layer.W_rho.shape, generator=g))
        if layer.bias_mu is not None:
            layer.bias_mu.copy_(0.1 * torch.randn(layer.bias_mu.shape, generator=g))
            layer.bias_rho.copy_(-3 + 0.5 * torch.randn(layer.bias_rho.shape, generator=g))


def record(arrs, name, layer, x, g, meta):
    with G.EpsInjector() as inj:
        y_shape = layer.eval()(x.detach()).shape
        layer.train()
        eps = torch.randn(y_shape, generator=g)
        inj.queue.append(eps)
        y = layer(x)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    arrs.update({f"{name}/x": x, f"{name}/W_mu": layer.W_mu, f"{name}/W_rho": layer.W_rho, f"{name}/eps": eps,
                 f"{name}/y": y, f"{name}/dy": dy, f"{name}/dx": x.grad, f"{name}/dW_mu": layer.W_mu.grad,
                 f"{name}/dW_rho": layer.W_rho.grad, f"{name}/meta": np.array(meta)})
    if layer.bias_mu is not None:
        arrs.update({f"{name}/bias_mu": layer.bias_mu, f"{name}/bias_rho": layer.bias_rho,
                     f"{name}/dbias_mu": layer.bias_mu.grad, f"{name}/dbias_rho": layer.bias_rho.grad})
    layer.eval()
    arrs[f"{name}/y_eval"] = layer(x.detach())
    arrs[f"{name}/kl"] = layer._kl.detach()


def main():
    arrs = {}
    g = torch.Generator().manual_seed(23)
    # (name, cin, cout, k, stride, padding, bias, N, H, W)
    shapes = [("l1x1", 16, 8, 1, 1, 0, 1, 1, 12, 12), ("l3s1p1", 12, 16, 3, 1, 1, 1, 2, 10, 14),
              ("l3s2", 16, 24, 3, 2, 0, 1, 1, 15, 19), ("l5s1nb", 4, 8, 5, 1, 2, 0, 1, 9, 9)]
    for name, cin, cout, k, st, pad, bias, N, H, W in shapes:
        layer = Conv2dLRT(cin, cout, k, bias=bool(bias), stride=st, padding=pad, prior={"mu": 0.0, "sigma": 0.05})
        fill(layer, g)
        x = torch.randn(N, cin, H, W, generator=g, requires_grad=True)
        record(arrs, name, layer, x, g, [cin, cout, k, st, pad, bias, N, H, W])
    lin = LinearLRT(20, 12, prior={"mu": 0.0, "sigma": 0.05})
    fill(lin, g)
    x = torch.randn(5, 20, generator=g, requires_grad=True)
    record(arrs, "lin", lin, x, g, [20, 12])
    G.save("lrt_layers.npz", **arrs)


if __name__ == "__main__":
    main()
