#!/usr/bin/env python
"""Diagnostic: run the RGI encoder conv stack unit by unit with two conv kernel generations and report where they diverge."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import e4s_oracle as O
from e4s_b200 import kernels as K
from e4s_b200.encoders.psp_encoders import FSEncoder_PSP

dev = "cuda:0"
enc = FSEncoder_PSP().eval()
st = O.synthetic_state({k: tuple(v.shape) for k, v in enc.state_dict().items()}, salt=5)
enc.load_state_dict(st)
enc = enc.to(dev)
x = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(1)).to(dev)


def run(mode):
    os.environ["E4S_B200_CONV"] = mode
    outs = []
    with torch.no_grad():
        b, c, h, w = x.shape
        xp = x.new_zeros((b, h, w, 32))
        xp[..., :c] = x.permute(0, 2, 3, 1)
        y = K.conv3x3_tcp(xp, enc._prepared("in", enc.input_layer[0].weight, pad_cin_to=32))
        outs.append(("conv0", y))
        s0, t0 = K.instnorm_affine(y)
        cur = K.norm_residual(y, s0, t0, 1.0, prelu=enc.input_layer[2].weight)
        for i, unit in enumerate(enc.body):
            cur = enc._unit(i, unit, cur)
            outs.append((f"unit{i} {tuple(cur.shape)}", cur))
    return outs


a = run(sys.argv[1] if len(sys.argv) > 1 else "tcp")
b = run(sys.argv[2] if len(sys.argv) > 2 else "tcr")
for (n, ta), (_, tb) in zip(a, b):
    err = float((ta - tb).abs().max() / ta.abs().max())
    print(f"{n:34s} rel diff {err:.3e}")
