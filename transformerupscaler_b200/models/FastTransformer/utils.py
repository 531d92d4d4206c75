"""Parameter holders mirroring the reference's models/FastTransformer/utils.py (default_conv :6-10,
BasicConv :13-40, Upsampler :43-98): same module tree => same state_dict keys."""
import torch.nn as nn


def default_conv(in_channels, out_channels, kernel_size, bias=True, groups=1):
    return nn.Conv2d(in_channels, out_channels, kernel_size, padding=(kernel_size // 2), bias=bias, groups=groups)


class BasicConv(nn.Module):
    """conv (no bias) + ReLU; only the configuration the model uses (64->3, k3 s1 p1) is supported."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=1, relu=True, bias=False):
        super().__init__()
        self.out_channels, self.in_channels = out_planes, in_planes
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)
        self.bn = None
        self.relu = nn.ReLU(inplace=True) if relu else None


class Upsampler(nn.Module):
    """Sub-pixel upsamplers for the fixed scales {2,3,4,6}: conv(n -> r^2 n) + PixelShuffle(r); 4 = two x2 stages."""

    def __init__(self, conv, n_feats, valid_scales=(2, 3, 4, 6), bias=True):
        super().__init__()
        self.upsamplers = nn.ModuleDict()
        for scale in valid_scales:
            blocks = []
            if scale in (2, 4):
                for _ in range(scale // 2):
                    blocks += [conv(n_feats, 4 * n_feats, 3, bias), nn.PixelShuffle(2)]
            elif scale in (3, 6):
                blocks += [conv(n_feats, scale * scale * n_feats, 3, bias), nn.PixelShuffle(scale)]
            else:
                raise NotImplementedError(f"Scale={scale} not supported")
            self.upsamplers[str(scale)] = nn.Sequential(*blocks)
