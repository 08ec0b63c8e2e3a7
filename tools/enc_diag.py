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
        y = K.conv3x3_tc(xp, enc._prepared("in", enc.input_layer[0].weight, pad_cin_to=32))
        outs.append(("conv0", y))
        s0, t0 = K.instnorm_affine(y)
        cur = K.norm_residual(y, s0, t0, 1.0, prelu=enc.input_layer[2].weight)
        # unit 0 step by step
        unit = enc.body[0]
        conv1, prelu, conv2 = unit.res_layer[1], unit.res_layer[2], unit.res_layer[3]
        sx, tx = K.instnorm_affine(cur)
        r1 = K.conv3x3_tc(cur, enc._prepared("0.c1", conv1.weight), sx, tx, prelu.weight)
        outs.append(("u0.conv1 IN+PReLU 64->128", r1))
        r1n = K.conv3x3_tc(cur, enc._prepared("0.c1", conv1.weight))
        outs.append(("u0.conv1 plain 64->128", r1n))
        r2 = K.conv3x3_tc(r1n, enc._prepared("0.c2", conv2.weight), out_stride=2)
        outs.append(("u0.conv2 128->128 s2", r2))
        r2b = K.conv3x3_tc(r1n, enc._prepared("0.c2", conv2.weight), out_stride=1)
        outs.append(("u0.conv2 128->128 s1", r2b))
        sc = K.conv3x3_tc(cur, enc._prepared("0.sc", unit.shortcut_layer[0].weight), out_stride=2)
        outs.append(("u0.shortcut 64->128 s2", sc))
        for i, unit in enumerate(enc.body):
            cur = enc._unit(i, unit, cur)
            outs.append((f"unit{i} {tuple(cur.shape)}", cur))
            if i >= 3:
                break
    return outs


a = run(sys.argv[1] if len(sys.argv) > 1 else "tcp")
b = run(sys.argv[2] if len(sys.argv) > 2 else "tcr")
for (n, ta), (_, tb) in zip(a, b):
    err = float((ta - tb).abs().max() / ta.abs().max())
    print(f"{n:34s} rel diff {err:.3e}")
