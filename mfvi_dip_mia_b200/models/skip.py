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


_DROPOUT_DEFAULTS = {"dropout_mode_down": '2d', "dropout_p_down": 0.5, "dropout_mode_up": '2d', "dropout_p_up": 0.5,
                     "dropout_mode_skip": 'None', "dropout_p_skip": 0.5, "dropout_mode_output": 'None', "dropout_p_output": 0.5}


def skip(num_input_channels=2, num_output_channels=3,
         num_channels_down=[16, 32, 64, 128, 128], num_channels_up=[16, 32, 64, 128, 128],
         num_channels_skip=[4, 4, 4, 4, 4],
         filter_size_down=3, filter_size_up=3, filter_skip_size=1, need_sigmoid=True, need_bias=True,
         pad='zero', upsample_mode='nearest', downsample_mode='stride', act_fun='LeakyReLU', need1x1_up=True, **dropout):
    """Keyword-compatible with the reference's skip(): the eight dropout_mode_* / dropout_p_* keywords (same defaults) arrive
    in **dropout and are passed to the conv blocks, which reject the MC-dropout modes."""
    unknown = set(dropout) - set(_DROPOUT_DEFAULTS)
    if unknown:
        raise TypeError(f"skip() got unexpected keyword arguments {sorted(unknown)}")
    drop = {**_DROPOUT_DEFAULTS, **dropout}
    n = len(num_channels_down)
    assert n == len(num_channels_up) == len(num_channels_skip)
    per_scale = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * n
    upsample_mode, downsample_mode = per_scale(upsample_mode), per_scale(downsample_mode)
    filter_size_down, filter_size_up = per_scale(filter_size_down), per_scale(filter_size_up)

    def block(seq, cin, cout, k, stride, where, tag, it, down_mode='stride', with_act=True):
        """conv -> bn -> act appended to `seq`, then the positional children get their reference names."""
        push(seq, conv(cin, cout, k, stride, bias=need_bias, pad=pad, downsample_mode=down_mode,
                       dropout_mode=drop[f"dropout_mode_{where}"], dropout_p=drop[f"dropout_p_{where}"], iterator=it,
                       string=tag))
        push(seq, bn(cout))
        if with_act:
            push(seq, act(act_fun))
        return _tag_children(seq, tag, it)

    model = nn.Sequential()
    level = model
    cin = num_input_channels
    it_skip, it_deep = 1, 1
    it_up = 2 * n if need1x1_up else n
    for i in range(n):
        deeper, side = nn.Sequential(), nn.Sequential()
        c_skip, c_down, c_up = num_channels_skip[i], num_channels_down[i], num_channels_up[i]
        push(level, Concat(1, side, deeper) if c_skip != 0 else deeper)
        c_deep = num_channels_up[i + 1] if i < n - 1 else c_down
        push(level, bn(c_skip + c_deep))
        if c_skip != 0:
            it_skip = block(side, cin, c_skip, filter_skip_size, 1, 'skip', 'skip', it_skip)
        it_deep = block(deeper, cin, c_down, filter_size_down[i], 2, 'down', 'deeper', it_deep, downsample_mode[i])
        it_deep = block(deeper, c_down, c_down, filter_size_down[i], 1, 'down', 'deeper', it_deep)
        inner = nn.Sequential()
        if i < n - 1:
            push(deeper, inner)
        push(deeper, nn.Upsample(scale_factor=2, mode=upsample_mode[i]))
        block(level, c_skip + c_deep, c_up, filter_size_up[i], 1, 'up', 'up', it_up - 1)
        if need1x1_up:
            block(level, c_up, c_up, 1, 1, 'up', 'up', it_up)
            it_up -= 1
        it_up -= 1
        cin = c_down
        level = inner
    final_iter = 2 * n + 1 if need1x1_up else n + 1
    push(model, conv(num_channels_up[0], num_output_channels, 1, bias=need_bias, pad=pad,
                     dropout_mode=drop["dropout_mode_output"], dropout_p=drop["dropout_p_output"], iterator=final_iter,
                     string='up'))
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
