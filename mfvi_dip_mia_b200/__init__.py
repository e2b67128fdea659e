"""mfvi_dip_mia_b200 — B200-native (sm_100a) implementation of the MFVI-DIP training step.

Public surface (mirrors the reference's Python plugin API for this path):
    BayTorch.MeanFieldVI, BayTorch.modules.{Conv2dRT, LinearRT}
    models.{get_net, skip}
    utils.bayesian_utils.{gaussian_nll, gaussian_nll_inpainting}
    radon.FastRadonTransform
    trainer.MfviDipTrainer          (flat-buffer fast path used by the runners)
Importing the package loads csrc/libmfvidip.so and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from .engine import SkipEngine, SkipSpec, build_layout  # noqa: F401
from .trainer import MfviDipTrainer  # noqa: F401

__version__ = "0.1.0"
