#!/usr/bin/env python
"""A short, fixed sequence of hot-kernel launches for `ncu --set full` captures (one GPU, run under gpurun).

    ncu --set full --clock-control none --import-source on -k regex:'upfirdn2d_fir4|modconv3x3|torgb' -c 8 \
        -o gpurun_out/prof python tools/ncu_targets.py [--conv simt,tc]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from e4s_b200 import kernels as K
from e4s_b200.stylegan2.modconv import PreparedConv
from e4s_b200.stylegan2.op import upfirdn2d

ap = argparse.ArgumentParser()
ap.add_argument("--conv", default="simt,tc")
ap.add_argument("--batch", type=int, default=16)
args = ap.parse_args()
DEV, B = "cuda:0", args.batch
fir = torch.tensor([1., 3., 3., 1.])
fir = (torch.outer(fir, fir) / 64 * 4).to(DEV)

x = torch.randn(B, 32, 1025, 1025, device=DEV)
upfirdn2d(x, fir, pad=(1, 1))                      # the model's largest blur call
del x

for name, cin, cout, r, up in [("c11@256", 128, 128, 256, False), ("c7@64", 512, 512, 64, False), ("c12^512", 128, 64, 256, True),
                               ("c14^1024", 64, 32, 512, True), ("c15@1024", 32, 32, 1024, False)]:
    w = torch.randn(1, cout, cin, 3, 3, device=DEV)
    prep = PreparedConv().get(w, up, fir if up else None)
    xpm = torch.randn(B, r, r, cin, device=DEV)
    s = 1.0 + 0.1 * torch.randn(B, 1, cin, device=DEV)
    ro = 2 * r if up else r
    noise = torch.randn(B, 1, ro, ro, device=DEV)
    nw, bias = torch.tensor([0.1], device=DEV), torch.randn(cout, device=DEV)
    dm = K.demod(s, prep.wsq)
    label = None
    for mode in args.conv.split(","):
        if mode == "tc" and prep.w_hilo is not None:
            K.modconv3x3_tc_fwd(xpm, prep.w_hilo, s, dm, None, noise, nw, bias, up, True)
        elif mode == "tcp" and prep.w_hilo is not None:
            K.modconv3x3_tcp_fwd(xpm, prep.w_hilo, s, dm, None, noise, nw, bias, up, True)
        elif mode == "tcq" and prep.w_hilo is not None:
            K.modconv3x3_tcq_fwd(xpm, prep.w_hilo, s, dm, None, noise, nw, bias, up, True)
        elif mode == "tcr" and prep.w_hilo is not None:
            K.modconv3x3_tcr_fwd(xpm, prep.w_hilo, s, dm, label, noise, nw, bias, up, True)
        elif mode == "simt":
            K.modconv3x3_fwd(xpm, prep.wt, s, dm, None, noise, nw, bias, up, True)
    torch.cuda.synchronize()

w = torch.randn(1, 3, 32, 1, 1, device=DEV)
prep = PreparedConv().get(w, False, None)
xpm = torch.randn(B, 1024, 1024, 32, device=DEV)
K.torgb_fwd(xpm, prep.wrgb, torch.randn(B, 1, 32, device=DEV), None, torch.randn(3, device=DEV),
            torch.randn(B, 3, 512, 512, device=DEV), fir)
torch.cuda.synchronize()
print("done")
