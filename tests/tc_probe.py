"""Diagnostic (not a pytest file): runs the tcgen05 kernel on a ladder of cases in both operand-shift encodings
and prints the error against the SIMT kernel.  Used under gpurun when bringing the tensor-core path up."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from test_parity_gpu import _tc_case, TC_CASES, TCP_EXTRA, PRODUCTION_CASES

CASES = PRODUCTION_CASES + TC_CASES + TCP_EXTRA + [(2, 64, 128, 30, True, 5, "blobs"), (1, 160, 256, 28, False, 12, "iid"), (2, 256, 64, 16, True, 12, "iid")]

for mode in ("tcr",):
    for case in CASES:
        try:
            K, prep, x, args = _tc_case(*case, seed=sum(c for c in case if isinstance(c, int)))
            ref = K.modconv3x3_fwd(x, prep.wt, *args)
            out = K.modconv3x3_tcr_fwd(x, prep.w_hilo, *args)
            torch.cuda.synchronize()
            err = float((out - ref).abs().max() / ref.abs().max())
            bad = int(((out - ref).abs() > 1e-3 * ref.abs().max()).sum())
            print(f"kernel={mode} case={case}: max-rel err {err:.3e}, elements off {bad}/{out.numel()}", flush=True)
            if err > 1e-3:
                d = (out - ref).abs().amax(dim=3)[0]            # per pixel error map of sample 0
                rows = (d > 1e-3 * ref.abs().max()).nonzero()
                print("   first bad pixels:", rows[:12].tolist(), " last:", rows[-4:].tolist(), flush=True)
                ty = (rows[:, 0] // (16 if case[4] else 8)).unique().tolist()
                tx = (rows[:, 1] // (28 if case[4] else 14)).unique().tolist()
                print("   bad tile rows:", ty[:40], " bad tile cols:", tx[:40], flush=True)
        except Exception:
            traceback.print_exc()
            print(f"kernel={mode} case={case}: EXCEPTION", flush=True)
            sys.exit(1)          # a trapped kernel poisons the context
