"""Setup helpers of the runners (reference utils/common_utils.py): input noise, parameter lists, array/tensor
conversion and the on-device PSNR / SSIM metrics used for per-iteration bookkeeping."""
import numpy as np
import torch
import torch.nn.functional as F


def fill_noise(x, noise_type):
    if noise_type == 'u':
        x.uniform_()
    elif noise_type == 'n':
        x.normal_()
    else:
        raise AssertionError(noise_type)


def get_noise(input_depth, method, spatial_size, noise_type='u', var=1. / 10, library='torch',
              data_format='channels_first'):
    """(1, input_depth, H, W) tensor of U[0,1)*var (or N(0,1)*var) — reference utils/common_utils.py:134-176."""
    if isinstance(spatial_size, int):
        spatial_size = (spatial_size, spatial_size)
    if method != 'noise':
        raise NotImplementedError("get_noise: only method='noise' is used by the MFVI runners")
    shape = ([1, input_depth, spatial_size[0], spatial_size[1]] if data_format == 'channels_first'
             else [1, spatial_size[0], spatial_size[1], input_depth])
    net_input = torch.zeros(shape)
    fill_noise(net_input, noise_type)
    net_input *= var
    return net_input


def get_params(opt_over, net, net_input, downsampler=None):
    """Parameters to optimise over ('net', 'down', 'input'; reference utils/common_utils.py:29-53)."""
    params = []
    for opt in opt_over.split(','):
        if opt == 'net':
            params += [x for x in net.parameters()]
        elif opt == 'down':
            assert downsampler is not None
            params = [x for x in downsampler.parameters()]
        elif opt == 'input':
            net_input.requires_grad = True
            params += [net_input]
        else:
            raise AssertionError('what is it?')
    return params


def np_to_torch(img_np):
    return torch.from_numpy(img_np)[None, :]


def torch_to_np(img_var):
    return img_var.detach().cpu().numpy()[0]


def crop_to_multiple(img_np, d=32):
    """Centre-crop a (C,H,W) array so that H and W are divisible by d (crop_image of the reference, on arrays)."""
    _, h, w = img_np.shape
    nh, nw = h - h % d, w - w % d
    t, l = (h - nh) // 2, (w - nw) // 2
    return img_np[:, t:t + nh, l:l + nw]


def peak_signal_noise_ratio(image_true, image_test):
    """10*log10(1/mse) for images in [0,1] (reference utils/common_utils.py:297-305)."""
    err = F.mse_loss(image_true, image_test)
    return (10 * torch.log10(1 / err)).item()


def structural_similarity(image_true, image_test, window_size=11, size_average=True, sigma=1.5):
    """11x11-Gaussian SSIM with zero padding, C1=0.01^2, C2=0.03^2 (reference utils/common_utils.py:308-353)."""
    g = torch.tensor([np.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    g = g / g.sum()
    ch = image_true.shape[1]
    win = (g[:, None] @ g[None, :]).expand(ch, 1, window_size, window_size).contiguous().to(image_true)
    p = window_size // 2
    mu1 = F.conv2d(image_true, win, padding=p, groups=ch)
    mu2 = F.conv2d(image_test, win, padding=p, groups=ch)
    s1 = F.conv2d(image_true * image_true, win, padding=p, groups=ch) - mu1 * mu1
    s2 = F.conv2d(image_test * image_test, win, padding=p, groups=ch) - mu2 * mu2
    s12 = F.conv2d(image_true * image_test, win, padding=p, groups=ch) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return ssim_map.mean().item() if size_average else ssim_map.mean(1).mean(1).mean(1).item()
