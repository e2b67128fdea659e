from .conv import Conv2dRT
from .linear import LinearRT
from .module import VIModule
from .reparam_layers import RTLayer

__all__ = ["Conv2dRT", "LinearRT", "VIModule", "RTLayer"]
