"""FastRadonTransform (reference radon/radon.py:4-55) on the library's analytic gather-reduce kernel: no
(T,H,W,2) sampling grid is materialised — the reference's `grid`, `trans`, `z` buffers are therefore absent,
`theta`, `ts`, `tc` are kept."""
import torch

from .. import functional as Fn


class FastRadonTransform(torch.nn.Module):
    def __init__(self, image_size, theta=None):
        super().__init__()
        assert image_size[-2] == image_size[-1]
        if theta is None:
            theta = torch.deg2rad(torch.arange(180.))
        else:
            theta = torch.deg2rad(theta)
        theta = theta.to(torch.float32)
        self.image_size = tuple(image_size)
        self.register_buffer("theta", theta)
        self.register_buffer("ts", torch.sin(theta))
        self.register_buffer("tc", torch.cos(theta))

    def forward(self, image):
        """image (1,C,H,W) -> sinogram (1,C,T,W), rows summed (dim 2) like the reference."""
        return Fn.RadonFn.apply(image, self.theta.contiguous())
