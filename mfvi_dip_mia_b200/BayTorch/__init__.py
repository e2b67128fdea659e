"""Drop-in mirror of the reference's BayTorch package for the MFVI hot path (reference BayTorch/__init__.py)."""
from .freq_to_bayes import MeanFieldVI

__all__ = ["MeanFieldVI"]
