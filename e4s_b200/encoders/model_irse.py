"""ArcFace IR / IR-SE backbone used by the identity loss: mirror of src/models/encoders/model_irse.py:9-69 (same module
tree and state-dict keys: ``input_layer.{0,1,2}``, ``body.{i}.{shortcut_layer,res_layer}.*``, ``output_layer.{0,3,4}``)."""
import torch
from torch import nn

from .helpers import Flatten, bottleneck_IR, bottleneck_IR_SE, get_blocks, l2_norm


class Backbone(nn.Module):
    def __init__(self, input_size, num_layers, mode="ir", drop_ratio=0.4, affine=True):
        super().__init__()
        assert input_size in [112, 224], "input_size should be 112 or 224"
        assert num_layers in [50, 100, 152], "num_layers should be 50, 100 or 152"
        assert mode in ["ir", "ir_se"], "mode should be ir or ir_se"
        unit_module = bottleneck_IR if mode == "ir" else bottleneck_IR_SE
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, (3, 3), 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        side = 7 if input_size == 112 else 14
        self.output_layer = nn.Sequential(nn.BatchNorm2d(512), nn.Dropout(drop_ratio), Flatten(),
                                          nn.Linear(512 * side * side, 512), nn.BatchNorm1d(512, affine=affine))
        modules = []
        for block in get_blocks(num_layers):
            for bottleneck in block:
                modules.append(unit_module(bottleneck.in_channel, bottleneck.depth, bottleneck.stride))
        self.body = nn.Sequential(*modules)

    def forward(self, x, multi_scale=False):
        x = self.input_layer(x)
        taps = []
        if multi_scale:                                   # features after units 2, 6, 20, 23 (model_irse.py:49-59)
            for i, unit in enumerate(self.body):
                x = unit(x)
                if i in (2, 6, 20, 23):
                    taps.append(l2_norm(x.reshape(x.size(0), -1)))
        else:
            x = self.body(x)
        x = self.output_layer(x)
        return taps + [l2_norm(x)]
