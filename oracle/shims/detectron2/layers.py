"""Import-only stand-ins (Painter/models_painter.py:18); dead code for the stock configs
(residual_block_indexes=[])."""
import torch.nn as nn


class CNNBlockBase(nn.Module):
    def __init__(self, in_channels, out_channels, stride):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride


class Conv2d(nn.Conv2d):
    def __init__(self, *a, norm=None, activation=None, **k):
        super().__init__(*a, **k)
        self.norm, self.activation = norm, activation


def get_norm(norm, out_channels):
    raise NotImplementedError("detectron2 shim: get_norm is not on the Painter hot path")
