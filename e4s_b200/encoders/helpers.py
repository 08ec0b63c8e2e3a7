"""Building blocks of the RGI encoder and of the ArcFace backbone (mirror of src/models/encoders/helpers.py:
Flatten :10-12, l2_norm :15-18, Bottleneck/get_block/get_blocks :21-53, SEModule :56-72, bottleneck_IR :75-94,
bottleneck_IR_SE :97-119, bottleneck_IR_SE_Ours :122-144).

Module/parameter names match the reference so its checkpoints load (`res_layer.1.weight`, `res_layer.5.fc1.weight`,
`shortcut_layer.0.weight` ...).
"""
from collections import namedtuple

import torch
from torch import nn


class Bottleneck(namedtuple("Block", ["in_channel", "depth", "stride"])):
    """(in_channel, depth, stride) of one residual unit."""


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def get_blocks(num_layers):
    units = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    if num_layers not in units:
        raise ValueError("Invalid number of layers: {}. Must be one of [50, 100, 152]".format(num_layers))
    n = units[num_layers]
    return [get_block(64, 64, n[0]), get_block(64, 128, n[1]), get_block(128, 256, n[2]), get_block(256, 512, n[3])]


class Flatten(nn.Module):
    def forward(self, input):
        return input.reshape(input.size(0), -1)


def l2_norm(input, axis=1):
    return torch.div(input, torch.norm(input, 2, axis, True))


class SEModule(nn.Module):
    """Squeeze-and-excitation gate: x * sigmoid(fc2(relu(fc1(mean_hw(x)))))."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        gate = self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))
        return x * gate


class bottleneck_IR_SE_Ours(nn.Module):
    """IR-SE residual unit with InstanceNorm instead of BatchNorm (helpers.py:122-144)."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False),
                                                nn.InstanceNorm2d(depth))
        self.res_layer = nn.Sequential(
            nn.InstanceNorm2d(in_channel),
            nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False),
            nn.PReLU(depth),
            nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False),
            nn.InstanceNorm2d(depth),
            SEModule(depth, 16),
        )

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class bottleneck_IR(nn.Module):
    """IR residual unit (BatchNorm), helpers.py:75-94."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False), nn.BatchNorm2d(depth))
        self.res_layer = nn.Sequential(
            nn.BatchNorm2d(in_channel), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False), nn.PReLU(depth),
            nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False), nn.BatchNorm2d(depth))

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class bottleneck_IR_SE(nn.Module):
    """IR-SE residual unit (BatchNorm + squeeze-excitation), helpers.py:97-119: the ArcFace backbone of the identity loss."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False), nn.BatchNorm2d(depth))
        self.res_layer = nn.Sequential(
            nn.BatchNorm2d(in_channel), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False), nn.PReLU(depth),
            nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False), nn.BatchNorm2d(depth), SEModule(depth, 16))

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)
