from .radon import FastRadonTransform

__all__ = ["FastRadonTransform"]
