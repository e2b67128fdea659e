from .conv import Conv2dLRT, Conv2dRT
from .linear import LinearLRT, LinearRT
from .module import VIModule
from .reparam_layers import LRTLayer, RTLayer

__all__ = ["Conv2dRT", "Conv2dLRT", "LinearRT", "LinearLRT", "VIModule", "RTLayer", "LRTLayer"]
