"""Face-parsing feature loss: mirror of src/criteria/face_parsing/{face_parsing_loss.py:20-78, unet.py:6-92,
model_utils.py:177-221}.  ``unet`` keeps the full module tree (decoder included) so the shipped checkpoint
``pretrained_ckpts/auxiliray/model.pth`` loads with strict=True; the loss only runs the encoder half (unet.py:71-92)."""
import os

import torch
import torch.nn.functional as F
from torch import nn

from ..encoders.helpers import l2_norm


class unetConv2(nn.Module):
    def __init__(self, in_size, out_size, is_batchnorm):
        super().__init__()
        if is_batchnorm:
            self.conv1 = nn.Sequential(nn.Conv2d(in_size, out_size, 3, 1, 1), nn.BatchNorm2d(out_size), nn.ReLU())
            self.conv2 = nn.Sequential(nn.Conv2d(out_size, out_size, 3, 1, 1), nn.BatchNorm2d(out_size), nn.ReLU())
        else:
            self.conv1 = nn.Sequential(nn.Conv2d(in_size, out_size, 3, 1, 1), nn.ReLU())
            self.conv2 = nn.Sequential(nn.Conv2d(out_size, out_size, 3, 1, 1), nn.ReLU())

    def forward(self, inputs):
        return self.conv2(self.conv1(inputs))


class unetUp(nn.Module):
    def __init__(self, in_size, out_size, is_deconv, is_batchnorm):
        super().__init__()
        self.conv = unetConv2(in_size, out_size, is_batchnorm)
        self.up = nn.ConvTranspose2d(in_size, out_size, kernel_size=2, stride=2) if is_deconv else nn.UpsamplingBilinear2d(scale_factor=2)

    def forward(self, inputs1, inputs2):
        outputs2 = self.up(inputs2)
        offset = outputs2.size()[2] - inputs1.size()[2]
        outputs1 = F.pad(inputs1, 2 * [offset // 2, offset // 2])
        return self.conv(torch.cat([outputs1, outputs2], 1))


class unet(nn.Module):
    def __init__(self, feature_scale=4, n_classes=19, is_deconv=True, in_channels=3, is_batchnorm=True):
        super().__init__()
        filters = [int(x / feature_scale) for x in [64, 128, 256, 512, 1024]]
        self.conv1 = unetConv2(in_channels, filters[0], is_batchnorm)
        self.maxpool1 = nn.MaxPool2d(kernel_size=2)
        self.conv2 = unetConv2(filters[0], filters[1], is_batchnorm)
        self.maxpool2 = nn.MaxPool2d(kernel_size=2)
        self.conv3 = unetConv2(filters[1], filters[2], is_batchnorm)
        self.maxpool3 = nn.MaxPool2d(kernel_size=2)
        self.conv4 = unetConv2(filters[2], filters[3], is_batchnorm)
        self.maxpool4 = nn.MaxPool2d(kernel_size=2)
        self.center = unetConv2(filters[3], filters[4], is_batchnorm)
        self.up_concat4 = unetUp(filters[4], filters[3], is_deconv, is_batchnorm)
        self.up_concat3 = unetUp(filters[3], filters[2], is_deconv, is_batchnorm)
        self.up_concat2 = unetUp(filters[2], filters[1], is_deconv, is_batchnorm)
        self.up_concat1 = unetUp(filters[1], filters[0], is_deconv, is_batchnorm)
        self.final = nn.Conv2d(filters[0], n_classes, 1)

    def _encode(self, inputs):
        conv1 = self.conv1(inputs)
        conv2 = self.conv2(self.maxpool1(conv1))
        conv3 = self.conv3(self.maxpool2(conv2))
        conv4 = self.conv4(self.maxpool3(conv3))
        center = self.center(self.maxpool4(conv4))
        return conv1, conv2, conv3, conv4, center

    def forward(self, inputs):
        conv1, conv2, conv3, conv4, center = self._encode(inputs)
        up4 = self.up_concat4(conv4, center)
        up3 = self.up_concat3(conv3, up4)
        up2 = self.up_concat2(conv2, up3)
        up1 = self.up_concat1(conv1, up2)
        return self.final(up1)

    def extract_feats(self, inputs):
        bs = inputs.size(0)
        return [l2_norm(f.reshape(bs, -1)) for f in self._encode(inputs)]


class FaceParsingLoss(nn.Module):
    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.face_pool = nn.AdaptiveAvgPool2d((512, 512))
        self.G = unet()
        path = getattr(opts, "face_parsing_model_path", None)
        if path and os.path.exists(path):
            self.G.load_state_dict(torch.load(path, map_location="cpu"))
        self.G.eval()
        self.set_requires_grad(False)

    def set_requires_grad(self, flag=True):
        for p in self.parameters():
            p.requires_grad = flag

    def extract_feats(self, x):
        x = self.face_pool(x) if x.shape[2] != 512 else x       # resize to 512 if needed
        return self.G.extract_feats(x)

    @staticmethod
    def loss_from_feats(y_hat_feats_ms, y_feats_ms):
        loss_all = 0
        for y_hat_feats, y_feats in zip(y_hat_feats_ms, y_feats_ms):
            loss_all = loss_all + (1 - (y_hat_feats * y_feats).sum(1)).mean()
        return loss_all

    def forward(self, y_hat, y):
        y_feats_ms = [f.detach() for f in self.extract_feats(y)]
        y_hat_feats_ms = self.extract_feats(y_hat)
        loss_all = self.loss_from_feats(y_hat_feats_ms, y_feats_ms)
        sim_improvement_all = 0.0
        for y_hat_feats, y_feats in zip(y_hat_feats_ms, y_feats_ms):
            sim_improvement_all += float(((y_hat_feats * y_feats).sum(1) - (y_feats * y_feats).sum(1)).mean())
        return loss_all, sim_improvement_all
