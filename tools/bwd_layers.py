#!/usr/bin/env python
"""Per-launch times of one inversion step (one 1024x1024 face): which layers the backward spends its time in."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
from e4s_b200 import kernels as K
from e4s_b200.masks import labelMap2OneHot
from e4s_b200.optimization import invert

dev = torch.device("cuda", 0)
net = B.build_net(1024, 12, dev)
for prm in net.parameters():
    prm.requires_grad = False
labels = B.face_label_maps(1, 12, "faces", seed=200)
onehot = labelMap2OneHot(labels.to(dev), 12)
g = torch.Generator().manual_seed(300)
sv = 0.5 * torch.randn(1, 12, 1280, generator=g).to(dev)
with torch.no_grad():
    target, _, _ = net.gen_img(None, net.cal_style_codes(0.5 * torch.randn(1, 12, 1280, generator=g).to(dev)), onehot)
invert(net, target, onehot, style_vectors=sv, steps=3)
K.LaunchStats.reset(timing=True)
invert(net, target, onehot, style_vectors=sv, steps=1)
torch.cuda.synchronize()
tot = {}
for name, work, e0, e1 in K.LaunchStats.records:
    ms = e0.elapsed_time(e1)
    if "modconv3x3" in name:
        print(f"{name:28s} {ms:8.3f} ms")
    tot[name] = tot.get(name, 0.0) + ms
print({k: round(v, 3) for k, v in tot.items()})
