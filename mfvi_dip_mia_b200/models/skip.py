"""skip(): the hour-glass encoder-decoder with skip connections (reference models/skip.py:5-134).

The returned nn.Sequential has the reference's module tree and state-dict keys (rename_modules,
utils/common_utils.py:248-262) and additionally carries `_skip_spec`, the graph description the fused
sm_100a engine executes once the net is wrapped in MeanFieldVI."""
import torch.nn as nn

from ..engine import SkipSpec
from .common import Concat, act, bn, conv, push


def _tag_children(seq: nn.Sequential, string: str, number: int) -> int:
    """Give every still-positional child (single-character key) the name '<Class>_<string>_<number>'; a second
    child of the same class gets the suffix '_1'.  Mirrors rename_modules of the reference."""
    renamed = {}
    for key, module in list(seq._modules.items()):
        if len(key) == 1:
            name = f'{module._get_name()}_{string}_{number}'
            renamed[name if name not in renamed else name + '_1'] = module
        else:
            renamed[key] = module
    seq._modules.clear()
    seq._modules.update(renamed)
    return number + 1


def skip(num_input_channels=2, num_output_channels=3,
         num_channels_down=[16, 32, 64, 128, 128], num_channels_up=[16, 32, 64, 128, 128],
         num_channels_skip=[4, 4, 4, 4, 4],
         filter_size_down=3, filter_size_up=3, filter_skip_size=1, need_sigmoid=True, need_bias=True,
         pad='zero', upsample_mode='nearest', downsample_mode='stride', act_fun='LeakyReLU', need1x1_up=True,
         dropout_mode_down='2d', dropout_p_down=0.5, dropout_mode_up='2d', dropout_p_up=0.5,
         dropout_mode_skip='None', dropout_p_skip=0.5, dropout_mode_output='None', dropout_p_output=0.5):
    n = len(num_channels_down)
    assert n == len(num_channels_up) == len(num_channels_skip)
    per_scale = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * n
    upsample_mode, downsample_mode = per_scale(upsample_mode), per_scale(downsample_mode)
    filter_size_down, filter_size_up = per_scale(filter_size_down), per_scale(filter_size_up)

    model = nn.Sequential()
    level = model
    cin = num_input_channels
    it_skip, it_deep = 1, 1
    it_up = 2 * n if need1x1_up else n
    for i in range(n):
        deeper, side = nn.Sequential(), nn.Sequential()
        has_skip = num_channels_skip[i] != 0
        push(level, Concat(1, side, deeper) if has_skip else deeper)
        c_deep = num_channels_up[i + 1] if i < n - 1 else num_channels_down[i]
        push(level, bn(num_channels_skip[i] + c_deep))
        if has_skip:
            push(side, conv(cin, num_channels_skip[i], filter_skip_size, bias=need_bias, pad=pad,
                            dropout_mode=dropout_mode_skip, dropout_p=dropout_p_skip, iterator=it_skip, string='skip'))
            push(side, bn(num_channels_skip[i]))
            push(side, act(act_fun))
            it_skip = _tag_children(side, 'skip', it_skip)
        push(deeper, conv(cin, num_channels_down[i], filter_size_down[i], 2, bias=need_bias, pad=pad,
                          downsample_mode=downsample_mode[i], dropout_mode=dropout_mode_down, dropout_p=dropout_p_down,
                          iterator=it_deep, string='deeper'))
        push(deeper, bn(num_channels_down[i]))
        push(deeper, act(act_fun))
        it_deep = _tag_children(deeper, 'deeper', it_deep)
        push(deeper, conv(num_channels_down[i], num_channels_down[i], filter_size_down[i], bias=need_bias, pad=pad,
                          dropout_mode=dropout_mode_down, dropout_p=dropout_p_down, iterator=it_deep, string='deeper'))
        push(deeper, bn(num_channels_down[i]))
        push(deeper, act(act_fun))
        it_deep = _tag_children(deeper, 'deeper', it_deep)
        inner = nn.Sequential()
        if i < n - 1:
            push(deeper, inner)
        push(deeper, nn.Upsample(scale_factor=2, mode=upsample_mode[i]))
        push(level, conv(num_channels_skip[i] + c_deep, num_channels_up[i], filter_size_up[i], 1, bias=need_bias,
                         pad=pad, dropout_mode=dropout_mode_up, dropout_p=dropout_p_up, iterator=it_up - 1, string='up'))
        push(level, bn(num_channels_up[i]))
        push(level, act(act_fun))
        _tag_children(level, 'up', it_up - 1)
        if need1x1_up:
            push(level, conv(num_channels_up[i], num_channels_up[i], 1, bias=need_bias, pad=pad,
                             dropout_mode=dropout_mode_up, dropout_p=dropout_p_up, iterator=it_up, string='up'))
            push(level, bn(num_channels_up[i]))
            push(level, act(act_fun))
            _tag_children(level, 'up', it_up)
            it_up -= 1
        it_up -= 1
        cin = num_channels_down[i]
        level = inner
    final_iter = 2 * n + 1 if need1x1_up else n + 1
    push(model, conv(num_channels_up[0], num_output_channels, 1, bias=need_bias, pad=pad,
                     dropout_mode=dropout_mode_output, dropout_p=dropout_p_output, iterator=final_iter, string='up'))
    if need_sigmoid:
        push(model, nn.Sigmoid())

    fused_ok = (pad == 'reflection' and act_fun == 'LeakyReLU' and need_bias and len(set(filter_size_down)) == 1
                and len(set(filter_size_up)) == 1 and len(set(upsample_mode)) == 1
                and all(m == 'stride' for m in downsample_mode))
    if fused_ok:
        model._skip_spec = SkipSpec(num_input_channels, num_output_channels, tuple(num_channels_down),
                                    tuple(num_channels_up), tuple(num_channels_skip), filter_size_down[0],
                                    filter_size_up[0], filter_skip_size, need1x1_up, need_sigmoid, upsample_mode[0])
    return model
