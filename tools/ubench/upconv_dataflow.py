#!/usr/bin/env python
"""Executable specification (numpy, CPU) of the data flow DESIGN.md section 10 proposes for the up-sampling layers:

    conv_transpose2d(stride 2, 3x3) -> blur 4x4 pad (1,1)       (reference: src/models/stylegan2/model.py:287-300)

as  (1) ONE GEMM per pixel tile with no spatial taps,  Z[q, (k, o)] = sum_i x[q, i] * W[k][o][i]   for the nine taps k,
    (2) a horizontal combination of each pixel's per-tap products with those of its x-neighbours (warp shuffles in the kernel),
    (3) a vertical combination with the rows above / below (shared-memory exchange in the kernel),
on 8 x 16 patches of input pixels with a one-pixel halo (6 x 14 interior pixels -> 12 x 28 outputs per tile).

`python tools/ubench/upconv_dataflow.py` checks the whole flow against conv_transpose2d + upfirdn2d on random data;
tests/test_host_logic.py runs the same check.  The blur must be separable (the model's is outer([1,3,3,1])).
"""
from __future__ import annotations

import numpy as np

TH, TW = 8, 16            # patch of input pixels = the 128 rows of the MMA's M dimension
IH, IW = TH - 2, TW - 2   # interior pixels whose outputs the tile owns


def axis_coefficients(taps1d: np.ndarray) -> np.ndarray:
    """C[p][d][k]: weight of conv-transpose tap k of input pixel i + d - 1 in output 2 i + p, for one axis.

    u[2 q + k] += Z[k][q] (transposed convolution, stride 2), out[y] = sum_a bf[a] * u[y + a - 1] (blur, pad (1, 1), true
    convolution: bf = flipped taps).  With y = 2 i + p and q = i + d - 1:  a = 2 (d - 1) + k + 1 - p."""
    bf = np.asarray(taps1d, dtype=np.float64)[::-1]
    c = np.zeros((2, 3, 3))
    for p in range(2):
        for d in range(3):
            for k in range(3):
                a = 2 * (d - 1) + k + 1 - p
                if 0 <= a < len(bf):
                    c[p, d, k] = bf[a]
    return c


def tile_forward(x_tile: np.ndarray, w: np.ndarray, cy: np.ndarray, cx: np.ndarray) -> np.ndarray:
    """x_tile [TH, TW, Cin] (zero outside the image); w [3, 3, Cout, Cin] -> [2 IH, 2 IW, Cout] outputs of the interior."""
    cout = w.shape[2]
    # (1) the GEMM: M = 128 pixels, N = 9 * Cout, K = Cin
    z = np.einsum("yxi,abOi->yxabO", x_tile, w)                     # [TH, TW, ky, kx, Cout]
    # (2) horizontal: g[y, x, ky, px] = sum_{dx, kx} cx[px][dx][kx] * z[y, x + dx - 1, ky, kx]   (x-neighbours: lane +-1)
    g = np.zeros((TH, TW, 3, 2, cout))
    for px in range(2):
        for dx in range(3):
            for kx in range(3):
                if cx[px, dx, kx] == 0.0:
                    continue
                lo, hi = max(0, 1 - dx), min(TW, TW + 1 - dx)      # columns whose neighbour x + dx - 1 is inside the patch
                g[:, lo:hi, :, px] += cx[px, dx, kx] * z[:, lo + dx - 1:hi + dx - 1, :, kx]
    # (3) vertical: out[2 i + py, 2 j + px] = sum_{dy, ky} cy[py][dy][ky] * g[i + dy - 1, j, ky, px]   (rows: lane +-16)
    out = np.zeros((2 * IH, 2 * IW, cout))
    for py in range(2):
        for dy in range(3):
            for ky in range(3):
                if cy[py, dy, ky] == 0.0:
                    continue
                out[py::2] += cy[py, dy, ky] * g[dy:dy + IH, 1:1 + IW, ky].reshape(IH, 2 * IW, cout)      # [i, (j, px), o]
    return out


def upconv(x: np.ndarray, w: np.ndarray, taps1d=(1, 3, 3, 1), gain: float = 4.0) -> np.ndarray:
    """x [H, W, Cin], w [3, 3, Cout, Cin] (conv_transpose2d taps), blur = gain * outer(taps) / sum -> [2H, 2W, Cout]."""
    t = np.asarray(taps1d, dtype=np.float64)
    t1 = t / t.sum() * np.sqrt(gain)                                # separable: outer(t1, t1) = gain * outer(t, t) / sum^2
    cy = cx = axis_coefficients(t1)
    h, wd, _ = x.shape
    out = np.zeros((2 * h, 2 * wd, w.shape[2]))
    xp = np.pad(x, ((1, TH), (1, TW), (0, 0)))                      # zero outside the image (TMA out-of-bounds fill)
    for ty in range(0, h, IH):
        for tx in range(0, wd, IW):
            tile = xp[ty:ty + TH, tx:tx + TW]                       # patch origin (ty - 1, tx - 1) in image coordinates
            o = tile_forward(tile, w, cy, cx)
            hh, ww = min(2 * IH, 2 * h - 2 * ty), min(2 * IW, 2 * wd - 2 * tx)
            out[2 * ty:2 * ty + hh, 2 * tx:2 * tx + ww] = o[:hh, :ww]
    return out


def reference(x: np.ndarray, w: np.ndarray, taps1d=(1, 3, 3, 1), gain: float = 4.0) -> np.ndarray:
    import torch
    import torch.nn.functional as F
    t = torch.tensor(taps1d, dtype=torch.float64)
    fir = torch.outer(t, t) / t.sum() ** 2 * gain
    xt = torch.from_numpy(x).permute(2, 0, 1)[None]                 # [1, Cin, H, W]
    wt = torch.from_numpy(w).permute(3, 2, 0, 1)                    # conv_transpose2d weight [Cin, Cout, 3, 3]
    u = F.conv_transpose2d(xt, wt, stride=2, padding=0)             # [1, Cout, 2H+1, 2W+1]
    u = F.pad(u, [1, 1, 1, 1])
    y = F.conv2d(u.reshape(-1, 1, *u.shape[2:]), torch.flip(fir, [0, 1])[None, None])      # true convolution, pad (1, 1)
    return y.reshape(w.shape[2], *y.shape[2:]).permute(1, 2, 0).numpy()


def self_check(seed: int = 0) -> float:
    rng = np.random.default_rng(seed)
    worst = 0.0
    for h, wd, cin, cout, taps in [(6, 14, 5, 3, (1, 3, 3, 1)), (13, 31, 4, 6, (1, 3, 3, 1)), (7, 9, 3, 2, (0.5, 2.0, -1.0, 3.0))]:
        x = rng.standard_normal((h, wd, cin))
        w = rng.standard_normal((3, 3, cout, cin))
        a, b = upconv(x, w, taps), reference(x, w, taps)
        assert a.shape == b.shape == (2 * h, 2 * wd, cout)
        worst = max(worst, float(np.abs(a - b).max() / np.abs(b).max()))
    return worst


if __name__ == "__main__":
    e = self_check()
    print(f"tile data flow vs conv_transpose2d + blur: max-rel error {e:.2e}")
    assert e < 1e-12
