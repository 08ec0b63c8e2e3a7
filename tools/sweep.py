#!/usr/bin/env python
"""BASELINE configs[4]: resolution / batch sweep of the synthesis step on one GPU (faces/s, device-timed).

    python tools/sweep.py [--sizes 256,512,1024] [--batches 1,2,4,8,16,32,64,128] [--ncls 12,19] [--out gpurun_out/sweep.json]

Each cell: 3 warm-up + 5 timed `Net3.gen_img` calls on synthetic codes and the example parsing masks, CUDA events on the
launching stream.  The N-GPU numbers of the sweep are bench.py under torchrun (faces shard with no collective)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench as B
from e4s_b200.masks import labelMap2OneHot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="256,512,1024")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128")
    ap.add_argument("--ncls", default="12,19")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    for ncls in [int(x) for x in args.ncls.split(",")]:
        for size in [int(x) for x in args.sizes.split(",")]:
            net = B.build_net(size, ncls, dev)
            for batch in [int(x) for x in args.batches.split(",")]:
                if ncls != 12 and batch not in (1, 16):
                    continue                                  # the class-count sweep point: two batch sizes are enough
                g = torch.Generator().manual_seed(batch)
                codes = torch.randn(batch, ncls, 18, 512, generator=g).to(dev)
                if ncls == 12:
                    labels = B.face_label_maps(batch, ncls, "faces", seed=1).to(dev)
                else:      # more regions than the example masks carry: split every region of the example masks pseudo-randomly
                    base = B.face_label_maps(batch, 12, "faces", seed=1).long()
                    yy = torch.arange(512).view(1, 1, 512, 1) // 64
                    labels = ((base + 12 * ((yy % 2) == 1)).clamp(max=ncls - 1)).to(torch.uint8).to(dev)
                onehot = labelMap2OneHot(labels, ncls)
                try:
                    with torch.no_grad():
                        for _ in range(3):
                            net.gen_img(None, codes, onehot)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(args.steps):
                            net.gen_img(None, codes, onehot)
                        e1.record()
                        torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.steps
                    row = {"size": size, "batch": batch, "ncls": ncls, "ms_per_step": round(ms, 3), "faces_per_s": round(batch / ms * 1e3, 1),
                           "algorithmic_tflops": round(B.ALGO_GFLOP_PER_FACE[size] * batch / ms, 1)}
                except torch.OutOfMemoryError:
                    row = {"size": size, "batch": batch, "ncls": ncls, "error": "out of memory"}
                    torch.cuda.empty_cache()
                rows.append(row)
                print(json.dumps(row), flush=True)
                del codes, onehot
            del net
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"rows": rows, "gpu": torch.cuda.get_device_name(0)}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
