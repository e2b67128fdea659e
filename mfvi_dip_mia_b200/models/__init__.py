"""get_net (reference models/__init__.py:4-27): only NET_TYPE='skip' exists in the reference.  The reference's eight
`dropout_mode_*` / `dropout_p_*` keywords are accepted and handed on unchanged (MFVI runs leave them at 'None')."""
from .skip import skip


def _per_scale(value, num_scales):
    return [value] * num_scales if isinstance(value, int) else value


def get_net(input_depth, NET_TYPE, pad, upsample_mode, n_channels=3, act_fun='LeakyReLU', need_sigmoid=False,
            skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, downsample_mode='stride', **dropout):
    if NET_TYPE != 'skip':
        raise AssertionError(f"NET_TYPE={NET_TYPE!r}: the reference only builds 'skip'")
    widths = {name: _per_scale(v, num_scales) for name, v in (("num_channels_down", skip_n33d),
                                                              ("num_channels_up", skip_n33u),
                                                              ("num_channels_skip", skip_n11))}
    for where in ("down", "up", "skip", "output"):          # the reference's defaults here are 'None' / 0.5
        dropout.setdefault(f"dropout_mode_{where}", 'None')
        dropout.setdefault(f"dropout_p_{where}", 0.5)
    return skip(input_depth, n_channels, upsample_mode=upsample_mode, downsample_mode=downsample_mode, pad=pad,
                need_sigmoid=need_sigmoid, need_bias=True, act_fun=act_fun, **widths, **dropout)
