"""Host-side streaming of the synthesis hot path: H2D of the next batch, the generator, and D2H of the previous batch's
images run on three CUDA streams, so a service pays max(copy, compute) per batch instead of their sum.

The reference's drivers (scripts/face_swap.py, scripts/optimization.py) keep everything on one stream and call
`.cpu()` per image; at 1024x1024 the 12.6 MB per face of fp32 output is ~15 % of a batch's time on a PCIe Gen5 link.

    pipe = SynthesisPipeline(net, ncls, depth=2)      # cuda_graph=True: the forward of a fixed batch shape replays as one graph
    for codes_host, labels_host in batches:           # pinned host tensors
        ticket = pipe.submit(codes_host, labels_host) # returns immediately
        ...
        images = pipe.result(ticket)                  # pinned host tensor [B,3,S,S]; valid until `depth` further submits
    pipe.drain()
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import kernels as K
from .stylegan2.modconv import LabelPyramid


class GraphedSynthesis:
    """``net.gen_img`` for ONE batch shape, captured once and replayed as a single CUDA graph: the ~90 launches of a
    1024x1024 forward run back to back, with no launch gaps and no host work per layer.

        synth = GraphedSynthesis(net, ncls, codes.shape, labels.shape)
        image = synth(codes, labels)        # device tensors; `image` is a static buffer the next call overwrites

    codes [B, ncls, n_latent, 512] fp32 (what ``Net3.cal_style_codes`` returns), labels [B, 1, H, W] uint8 class maps with
    values < ncls (``LabelPyramid``'s input; a float one-hot mask converts with ``LabelPyramid.from_mask(mask).base``, which
    also validates it).  Every replay draws fresh N(0,1) noise maps (torch's graph-safe Philox state advances per replay),
    like ``randomize_noise=True`` of the reference (model.py:333); ``randomize_noise=False`` uses the registered buffers."""

    def __init__(self, net, ncls: int, codes_shape, labels_shape, device: Optional[torch.device] = None,
                 randomize_noise: bool = True, warmup: int = 2):
        self.net, self.ncls = net, int(ncls)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.codes = torch.zeros(tuple(codes_shape), dtype=torch.float32, device=self.device)
        self.labels = torch.zeros(tuple(labels_shape), dtype=torch.uint8, device=self.device)
        assert self.labels.ndim == 4 and self.labels.shape[1] == 1, "labels: [B, 1, H, W] uint8"

        def forward():
            with torch.no_grad():
                regions = LabelPyramid(self.labels[:, 0], self.ncls)     # rebuilt inside the graph: its levels follow the labels
                return self.net.gen_img(None, self.codes, regions, randomize_noise=randomize_noise)[0]

        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):                          # weight preparation, allocator warm-up
                    forward()
            main.wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            before = K.LaunchStats.launches
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA
                self.image = forward()
            self.launches_per_replay = K.LaunchStats.launches - before   # C-ABI launches inside one replay
            K.LaunchStats.launches = before

    def __call__(self, codes: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if tuple(codes.shape) != tuple(self.codes.shape) or tuple(labels.shape) != tuple(self.labels.shape):
            raise RuntimeError(f"GraphedSynthesis was captured for codes {tuple(self.codes.shape)} / labels "
                               f"{tuple(self.labels.shape)}; got {tuple(codes.shape)} / {tuple(labels.shape)}")
        self.codes.copy_(codes, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        K.LaunchStats.launches += self.launches_per_replay
        return self.image


class SynthesisPipeline:
    def __init__(self, net, ncls: int, depth: int = 2, device: Optional[torch.device] = None, cuda_graph: bool = False):
        self.net, self.ncls, self.depth = net, ncls, depth
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.cuda_graph, self._graphed = cuda_graph, None
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._in: List[Optional[tuple]] = [None] * depth          # device staging buffers (codes, labels) per slot
        self._in_free: List[Optional[torch.cuda.Event]] = [None] * depth
        self._out: List[Optional[torch.Tensor]] = [None] * depth  # pinned host images per slot
        self._out_done: List[Optional[torch.cuda.Event]] = [None] * depth
        self._k = 0

    def submit(self, codes_host: torch.Tensor, labels_host: torch.Tensor) -> int:
        """codes_host [B,ncls,n_latent,512] fp32, labels_host [B,1,H,W] uint8 label maps; both pinned.  Returns a ticket."""
        s = self._k % self.depth
        self._k += 1
        main = torch.cuda.current_stream(self.device)
        if self._in[s] is None or self._in[s][0].shape != codes_host.shape or self._in[s][1].shape != labels_host.shape:
            self._in[s] = (torch.empty(codes_host.shape, dtype=codes_host.dtype, device=self.device),
                           torch.empty(labels_host.shape, dtype=labels_host.dtype, device=self.device))
            self._in_free[s] = None
        codes_dev, labels_dev = self._in[s]
        with torch.cuda.stream(self.h2d):
            if self._in_free[s] is not None:
                self.h2d.wait_event(self._in_free[s])             # the generator call that last read this slot is done
            codes_dev.copy_(codes_host, non_blocking=True)
            labels_dev.copy_(labels_host, non_blocking=True)
        main.wait_stream(self.h2d)
        if self.cuda_graph:
            if self._graphed is None or tuple(self._graphed.codes.shape) != tuple(codes_dev.shape) \
                    or tuple(self._graphed.labels.shape) != tuple(labels_dev.shape):
                self._graphed = GraphedSynthesis(self.net, self.ncls, codes_dev.shape, labels_dev.shape, self.device)
            img = self._graphed(codes_dev, labels_dev).clone()    # the static image is overwritten by the next replay
        else:
            with torch.no_grad():                                 # label maps go in as they are: no one-hot round trip
                img, _, _ = self.net.gen_img(None, codes_dev, LabelPyramid(labels_dev[:, 0], self.ncls))
        self._in_free[s] = torch.cuda.Event()
        self._in_free[s].record(main)
        if self._out_done[s] is not None:
            self._out_done[s].synchronize()                       # the host buffer of this slot is about to be rewritten
        if self._out[s] is None or self._out[s].shape != img.shape:
            self._out[s] = torch.empty(img.shape, dtype=img.dtype).pin_memory()
        self.d2h.wait_stream(main)
        with torch.cuda.stream(self.d2h):
            self._out[s].copy_(img, non_blocking=True)
        img.record_stream(self.d2h)                               # keep the allocator off `img` until the copy has run
        self._out_done[s] = torch.cuda.Event()
        self._out_done[s].record(self.d2h)
        return self._k - 1

    def result(self, ticket: int) -> torch.Tensor:
        if ticket < self._k - self.depth or ticket >= self._k:
            raise RuntimeError(f"ticket {ticket} is no longer (or not yet) held: {self.depth} results are kept")
        s = ticket % self.depth
        self._out_done[s].synchronize()
        return self._out[s]

    def drain(self) -> None:
        """Make the current stream wait for every copy in flight (so that an event recorded next covers them)."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.d2h)
        main.wait_stream(self.h2d)
