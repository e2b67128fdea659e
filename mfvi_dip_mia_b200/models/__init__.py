"""get_net (reference models/__init__.py:4-27): only NET_TYPE='skip' exists in the reference."""
from .skip import skip


def get_net(input_depth, NET_TYPE, pad, upsample_mode, n_channels=3, act_fun='LeakyReLU', need_sigmoid=False,
            skip_n33d=128, skip_n33u=128, skip_n11=4, num_scales=5, downsample_mode='stride',
            dropout_mode_down='None', dropout_p_down=0.5, dropout_mode_up='None', dropout_p_up=0.5,
            dropout_mode_skip='None', dropout_p_skip=0.5, dropout_mode_output='None', dropout_p_output=0.5):
    if NET_TYPE != 'skip':
        raise AssertionError(f"NET_TYPE={NET_TYPE!r}: the reference only builds 'skip'")
    as_list = lambda v: [v] * num_scales if isinstance(v, int) else v
    return skip(input_depth, n_channels, num_channels_down=as_list(skip_n33d), num_channels_up=as_list(skip_n33u),
                num_channels_skip=as_list(skip_n11), upsample_mode=upsample_mode, downsample_mode=downsample_mode,
                need_sigmoid=need_sigmoid, need_bias=True, pad=pad, act_fun=act_fun,
                dropout_mode_down=dropout_mode_down, dropout_p_down=dropout_p_down, dropout_mode_up=dropout_mode_up,
                dropout_p_up=dropout_p_up, dropout_mode_skip=dropout_mode_skip, dropout_p_skip=dropout_p_skip,
                dropout_mode_output=dropout_mode_output, dropout_p_output=dropout_p_output)
