"""ArcFace identity loss: mirror of src/criteria/id_loss.py:6-57.

``IDLoss(opts)`` reads ``opts.ir_se50_path`` (loaded when the file exists - the checkpoint cannot be downloaded here, so
benchmarks load seeded stand-ins afterwards) and ``opts.id_loss_multiscale``.  ``forward(y_hat, y)`` returns the
reference's 3-tuple; ``extract_feats`` / ``loss_from_feats`` are the cached-target path of `InversionLoss`.
"""
import os

import torch
from torch import nn

from ..encoders.model_irse import Backbone


class IDLoss(nn.Module):
    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.face_pool_1 = nn.AdaptiveAvgPool2d((256, 256))
        self.facenet = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode="ir_se")
        path = getattr(opts, "ir_se50_path", None)
        if path and os.path.exists(path):
            self.facenet.load_state_dict(torch.load(path, map_location="cpu"))
        self.face_pool_2 = nn.AdaptiveAvgPool2d((112, 112))
        self.facenet.eval()
        self.set_requires_grad(False)

    def set_requires_grad(self, flag=True):
        for p in self.parameters():
            p.requires_grad = flag

    def extract_feats(self, x):
        x = self.face_pool_1(x) if x.shape[2] != 256 else x     # (1) resize to 256 if needed
        x = x[:, :, 35:223, 32:220]                             # (2) crop the interesting region
        x = self.face_pool_2(x)                                 # (3) resize to 112 for the pre-trained model
        return self.facenet(x, multi_scale=getattr(self.opts, "id_loss_multiscale", True))

    @staticmethod
    def loss_from_feats(y_hat_feats_ms, y_feats_ms):
        """sum over scales of mean_i (1 - <y_hat_i, y_i>) (id_loss.py:41-55), as tensor ops (no per-sample host sync)."""
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
        return loss_all, sim_improvement_all, None
