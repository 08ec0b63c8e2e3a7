#!/usr/bin/env python
"""Executable specification (torch CPU, float64) of the H-form of an up-sampling StyledConv, checked against
conv_transpose2d(stride 2) + upfirdn2d blur (reference model.py:287-300).

The polyphase form (e4s_b200/stylegan2/modconv.py:fold_upsample_kernels) spends 36 tap-MACs per input pixel: four output
parities x 3x3 taps.  The blur is separable, so its HORIZONTAL half can move behind the GEMM:

    T[py, kx][m, n'] = sum_dy V[py, kx][dy] . x[m + dy - 1, n']          V[py, kx][dy] = sum_ky Ay[py][dy, ky] W[ky, kx]
    out[2m + py, 2n + px] = sum_{dx, kx} Ax[px][dx, kx] T[py, kx][m, n + dx - 1]
    A*[p][d, k] = flipped_fir_1d[2 (d - 1) + k + 1 - p]     (0 outside 0..3)

The GEMM has N = 6 x Cout columns ((py, kx) groups) and K = 3 x Cin (three row taps, no column taps): 18 tap-MACs per input
pixel; the six non-zero (dx, kx) pairs per output parity are combined in the epilogue from the accumulators of the pixel
itself and of its left / right neighbours (lane +-1 of the same warp in the 8x16 patch layout).

`fold_vertical(W, fir1d_y)` and `combine_horizontal(T, fir1d_x)` below are what e4s_b200/stylegan2/modconv.py and
csrc/modconv_tch.cu implement.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def tap_matrix(fir1d_flipped, p):
    """A[p][d, k] = fir_flipped[2 (d - 1) + k + 1 - p] (zero outside 0..3); d = input offset index, k = convT tap."""
    a = torch.zeros(3, 3, dtype=torch.float64)
    for d in range(3):
        for k in range(3):
            i = 2 * (d - 1) + k + 1 - p
            if 0 <= i <= 3:
                a[d, k] = fir1d_flipped[i]
    return a


def fold_vertical(w, fir1d_y_flipped):
    """w [Cout, Cin, 3, 3] -> V [6 (py * 3 + kx), 3 (dy), Cout, Cin]."""
    cout, cin = w.shape[:2]
    v = torch.zeros(6, 3, cout, cin, dtype=torch.float64)
    for py in range(2):
        ay = tap_matrix(fir1d_y_flipped, py)
        for kx in range(3):
            for dy in range(3):
                for ky in range(3):
                    v[py * 3 + kx, dy] += ay[dy, ky] * w[:, :, ky, kx].double()
    return v


def gemm_rows(x, v):
    """x [Cin, H, W] -> T [6, Cout, H, W]: three row taps, zero padding above / below."""
    cin, h, w = x.shape
    xp = F.pad(x.double(), (0, 0, 1, 1))
    t = torch.zeros(6, v.shape[2], h, w, dtype=torch.float64)
    for g in range(6):
        for dy in range(3):
            t[g] += torch.einsum("oi,ihw->ohw", v[g, dy], xp[:, dy:dy + h, :])
    return t


def combine_horizontal(t, fir1d_x_flipped):
    """T [6, Cout, H, W] -> out [Cout, 2H, 2W]; T is zero outside the image (its operand rows are)."""
    _, cout, h, w = t.shape
    tp = F.pad(t, (1, 1))                                           # column n' = -1 .. W
    out = torch.zeros(cout, 2 * h, 2 * w, dtype=torch.float64)
    for py in range(2):
        for px in range(2):
            ax = tap_matrix(fir1d_x_flipped, px)
            acc = torch.zeros(cout, h, w, dtype=torch.float64)
            for dx in range(3):
                for kx in range(3):
                    if ax[dx, kx] != 0:
                        acc += ax[dx, kx] * tp[py * 3 + kx][:, :, dx:dx + w]
            out[:, py::2, px::2] = acc
    return out


def reference(x, w, fir2d):
    """conv_transpose2d(stride 2) -> upfirdn2d(blur, pad (1, 1)) exactly as model.py:287-300 (true convolution: flipped FIR)."""
    u = F.conv_transpose2d(x[None].double(), w.double().transpose(0, 1), stride=2)     # [1, Cout, 2H+1, 2W+1]
    cout = u.shape[1]
    up = F.pad(u, (1, 1, 1, 1))
    k = torch.flip(fir2d.double(), [0, 1])[None, None].repeat(cout, 1, 1, 1)
    return F.conv2d(up, k, groups=cout)[0]


def main():
    torch.manual_seed(0)
    worst = 0.0
    for fir1 in ([1., 3., 3., 1.], [1., 2., 4., 3.]):               # the model's FIR and an asymmetric one
        f = torch.tensor(fir1, dtype=torch.float64)
        f2 = torch.outer(f, f)
        f2 = f2 / f2.sum() * 4
        fy = f / f.sum() * 2                                         # separable halves: outer(fy, fx) == f2
        fx = f / f.sum() * 2
        assert torch.allclose(torch.outer(fy, fx), f2)
        for (cin, cout, h, w) in [(3, 2, 4, 5), (5, 4, 7, 3), (2, 3, 1, 1)]:
            x = torch.randn(cin, h, w)
            wt = torch.randn(cout, cin, 3, 3)
            v = fold_vertical(wt, torch.flip(fy, [0]))
            out = combine_horizontal(gemm_rows(x, v), torch.flip(fx, [0]))
            ref = reference(x, wt, f2)
            err = float((out - ref).abs().max() / ref.abs().max())
            worst = max(worst, err)
            print(f"fir {fir1} cin {cin} cout {cout} {h}x{w}: max rel err {err:.2e}")
    # the six terms per output parity that the kernel's epilogue hard-codes (left = n-1, own = n, right = n+1):
    fl = ["f0", "f1", "f2", "f3"]
    for px in range(2):
        terms = []
        for dx in range(3):
            for kx in range(3):
                i = 2 * (dx - 1) + kx + 1 - px
                if 0 <= i <= 3:
                    terms.append(f"{fl[i]}*T{kx}[{['n-1', 'n', 'n+1'][dx]}]")
        print(f"px={px}: " + " + ".join(terms))
    assert worst < 1e-12, worst
    print("OK")


if __name__ == "__main__":
    main()
