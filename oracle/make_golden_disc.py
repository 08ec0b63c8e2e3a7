#!/usr/bin/env python
"""Pins oracle/disc_oracle.py to the reference's Discriminator (src/models/stylegan2/model.py:740-799) and writes
tests/golden/disc_vectors.npz.  Build container only (needs /root/reference).  The reference module runs unmodified; its two
CUDA-only ops are routed to the reference's own CPU spellings exactly as in oracle/make_golden.py (import_reference)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from oracle import disc_oracle as DO, e4s_oracle as O  # noqa: E402
import make_golden as MG  # noqa: E402


def main():
    M, _, _ = MG.import_reference()
    out = {}
    worst = 0.0
    for size, batch in ((32, 4), (64, 8), (128, 2)):
        D = M.Discriminator(size).eval()
        # seeded PARAMETERS; the blur FIR buffers keep the values the reference registers
        D.load_state_dict(O.synthetic_state({k: tuple(v.shape) for k, v in D.named_parameters()}, salt=size + 1), strict=False)
        st = {k: v.detach().clone() for k, v in D.state_dict().items()}
        x = DO.synthetic_inputs(batch, size, seed=size)
        with torch.no_grad():
            ref = D(x)
            ora = DO.discriminator_forward(st, x, size)
        e = float((ref - ora).abs().max() / ref.abs().max())
        worst = max(worst, e)
        print(f"Discriminator({size}) batch {batch}: oracle vs reference {e:.2e}")
        assert e <= 2e-5, e
        out[f"d{size}/logits"] = ref.numpy()
    dst = os.path.join(ROOT, "tests", "golden", "disc_vectors.npz")
    np.savez_compressed(dst, **out)
    print(f"worst {worst:.2e}; wrote {dst}")


if __name__ == "__main__":
    main()
