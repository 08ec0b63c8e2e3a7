"""RGI (region-wise) encoder: mirror of FSEncoder_PSP, src/models/encoders/psp_encoders.py:238-309.

The conv stack keeps the reference's module tree (checkpoint drop-in).  The per-region pooling
``get_per_comp_styleCode`` (:264-283) - a B x ncls Python loop with a host sync and a masked_select per
region in the reference - is ONE kernel over a uint8 label map here (``e4s_region_mean_f32``).
"""
import torch
from torch import nn

from .helpers import get_block, bottleneck_IR_SE_Ours
from .. import kernels as K
from ..stylegan2.modconv import LabelPyramid


class FSEncoder_PSP(nn.Module):
    def __init__(self, mode="ir_se", opts=None):
        super().__init__()
        assert mode in ["ir_se"], "the E4S RGI encoder is the ir_se variant (networks.py:48)"
        blocks = [get_block(64, 128, 3), get_block(128, 256, 4), get_block(256, 512, 14), get_block(512, 512, 3)]
        self.n_styles = 11
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, (3, 3), 1, 1, bias=False), nn.InstanceNorm2d(64), nn.PReLU(64))
        self.body = nn.Sequential(*[bottleneck_IR_SE_Ours(u.in_channel, u.depth, u.stride) for blk in blocks for u in blk])

    def get_per_comp_styleCode(self, style_feats, segmap):
        """style_feats [B,C,h,w]; segmap one-hot [B,ncls,H,W] (or LabelPyramid) -> [B,ncls,C] region means."""
        regions = LabelPyramid.from_mask(segmap)
        h, w = style_feats.shape[2:]
        codes, _area = K.region_mean(K.to_pixel_major(style_feats), regions.at(h, w), regions.ncls)
        return codes

    def forward(self, x, segmap):
        regions = LabelPyramid.from_mask(segmap)
        x = self.input_layer(x)
        taps = {}
        for i, unit in enumerate(self.body):
            x = unit(x)
            if i in (6, 20, 23):
                taps[i] = x
        codes = torch.cat([self.get_per_comp_styleCode(taps[i], regions) for i in (6, 20, 23)], dim=2)
        return codes, torch.zeros_like(x)
