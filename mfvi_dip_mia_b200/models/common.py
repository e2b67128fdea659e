"""Building blocks of the hour-glass net (reference models/common.py:15-135), limited to what the MFVI runners
instantiate: Concat, reflection/zero padded conv blocks with 'stride' downsampling, BatchNorm2d, LeakyReLU(0.2)."""
from collections import OrderedDict

import torch
import torch.nn as nn


def push(seq: nn.Sequential, module: nn.Module):
    """Append under the 1-based positional key the reference's `nn.Module.add` monkey-patch uses
    (models/common.py:8-12) — without patching torch."""
    seq.add_module(str(len(seq) + 1), module)


class Concat(nn.Module):
    """Runs every child on the same input, centre-crops to the smallest H/W and concatenates along `dim`
    (reference models/common.py:15-46)."""

    def __init__(self, dim, *branches):
        super().__init__()
        self.dim = dim
        for idx, module in enumerate(branches):
            self.add_module(str(idx), module)

    def forward(self, input):
        outs = [m(input) for m in self._modules.values()]
        h = min(o.shape[2] for o in outs)
        w = min(o.shape[3] for o in outs)
        cropped = []
        for o in outs:
            dh, dw = (o.shape[2] - h) // 2, (o.shape[3] - w) // 2
            cropped.append(o if (dh == 0 and dw == 0 and o.shape[2] == h and o.shape[3] == w)
                           else o[:, :, dh:dh + h, dw:dw + w])
        return torch.cat(cropped, dim=self.dim)

    def __len__(self):
        return len(self._modules)


def act(act_fun='LeakyReLU'):
    if isinstance(act_fun, str):
        if act_fun == 'LeakyReLU':
            return nn.LeakyReLU(0.2, inplace=True)
        if act_fun == 'none':
            return nn.Sequential()
        raise NotImplementedError(f"act_fun={act_fun!r}: the MFVI runners only use LeakyReLU")
    return act_fun()


def bn(num_features):
    return nn.BatchNorm2d(num_features)


def conv(in_f, out_f, kernel_size, stride=1, bias=True, pad='zero', downsample_mode='stride', dropout_mode=None,
         dropout_p=0.2, iterator=1, string='deeper'):
    """[ReflectionPad2d] -> Conv2d block with children named '<Class>_<string>_<iterator>'
    (reference models/common.py:100-135)."""
    if stride != 1 and downsample_mode != 'stride':
        raise NotImplementedError(f"downsample_mode={downsample_mode!r}: every MFVI runner uses 'stride'")
    if dropout_mode in ('1d', '2d'):
        raise NotImplementedError("dropout inside conv blocks belongs to the MC-dropout baseline, not to MFVI")
    to_pad = int((kernel_size - 1) / 2)
    layers = []
    if pad == 'reflection':
        layers.append(nn.ReflectionPad2d(to_pad))
        to_pad = 0
    layers.append(nn.Conv2d(in_f, out_f, kernel_size, stride, padding=to_pad, bias=bias))
    return nn.Sequential(OrderedDict((f'{m._get_name()}_{string}_{iterator}', m) for m in layers))
