"""Per-face texture-vector optimisation (inversion) inner loop on the e4s_b200 kernels.

Mirror of the hot loop of ``Optimizer.invertion`` (scripts/optimization.py:209-232) and of its optimiser set-up
(``setup_W_optimizer`` :125-161): Adam (or sgd / sgdm / adamax) on a clone of the per-region texture vectors
[1, ncls, 1280]; every step maps them to W+ codes with the LocalMLPs, synthesises the face with fresh noise,
evaluates the loss against the target image and back-propagates through the generator.

Loss: the reference sums l2 (:98-101) with LPIPS, ArcFace-ID and a parsing-net loss (:92-116), each switchable by
its lambda.  Pass ``criterion=e4s_b200.criteria.InversionLoss(...)`` for that full loss (reference default weights
0.1 ID + 1.0 l2 + 0.8 LPIPS x3 scales + 0.1 parsing; the target image's features are cached once per inversion instead
of being recomputed every step).  Without a criterion the loop uses the l2 term plus optional callables
``extra_losses = [(weight, fn(recon, target) -> scalar)]``.
"""
from __future__ import annotations

from functools import partial
import contextlib
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

OPTIMIZERS = {"sgd": torch.optim.SGD, "adam": torch.optim.Adam, "sgdm": partial(torch.optim.SGD, momentum=0.9),
              "adamax": torch.optim.Adamax}


def setup_W_optimizer(W_init: torch.Tensor, opt_name: str = "adam", lr: float = 1e-2, noise_init=None):
    """Clone the initial texture vectors (and optionally noise maps) as leaves and build the optimiser."""
    latent = W_init.clone().detach().requires_grad_(True)
    params = [latent]
    noises = None
    if noise_init is not None:
        noises = [n.clone().detach().requires_grad_(True) for n in noise_init]
        params += noises
    opt = OPTIMIZERS[opt_name](params, lr=lr)
    return (opt, latent, noises) if noises is not None else (opt, latent)


def invert(net, target: torch.Tensor, onehot: torch.Tensor, style_vectors: Optional[torch.Tensor] = None,
           steps: int = 200, lr: float = 1e-2, opt_name: str = "adam", l2_lambda: float = 1.0,
           extra_losses: Sequence[Tuple[float, Callable]] = (), noise: Optional[List[torch.Tensor]] = None,
           callback: Optional[Callable] = None, cuda_graph: bool = False, stats: Optional[dict] = None,
           criterion=None):
    """Optimise the texture vectors of ONE batch of faces so that net.gen_img reproduces `target`.

    net: e4s_b200.networks.Net3 (eval, latent_avg set).  target [B,3,S,S]; onehot [B,ncls,Hm,Wm].
    style_vectors: initial [B,ncls,1280] (default: the encoder's, as scripts/optimization.py:178-180).
    noise: fixed noise list, or None for fresh noise every step (the reference's behaviour, :216).
    cuda_graph: capture one step and replay it (Adam only); stats (optional dict) then receives the device time of the
    replayed steps alone ("replay_ms_per_step", "replayed_steps").
    criterion: an e4s_b200.criteria.InversionLoss (scripts/optimization.py:88-122 with cached target features); replaces
    l2_lambda / extra_losses.
    Returns (latent [B,ncls,1280], final reconstruction, list of per-step loss values as 0-d tensors).
    """
    if style_vectors is None:
        with torch.no_grad():
            style_vectors, _ = net.get_style_vectors(target, onehot)
    if cuda_graph:
        if opt_name != "adam":
            raise ValueError(f"cuda_graph=True supports opt_name='adam' only (capturable optimiser), got {opt_name!r}")
        if callback is not None:
            raise ValueError("cuda_graph=True replays a captured step: a per-step Python callback cannot run inside it")
        return _invert_graphed(net, target, onehot, style_vectors, steps, lr, l2_lambda, extra_losses, noise, stats, criterion)
    if criterion is not None:
        criterion.set_target(target)
    opt, latent = setup_W_optimizer(style_vectors, opt_name, lr)
    history, recon = [], None
    for step in range(steps):
        opt.zero_grad(set_to_none=True)
        codes = net.cal_style_codes(latent)
        recon, _, _ = net.gen_img(None, codes, onehot, noise=noise) if noise is not None else net.gen_img(None, codes, onehot)
        loss = _loss(recon, target, l2_lambda, extra_losses, criterion)
        with _precision(criterion):
            loss.backward()
        opt.step()
        history.append(loss.detach())
        if callback is not None:
            callback(step, loss, recon, latent)
    return latent.detach(), recon.detach(), history


def _precision(criterion):
    """The cuDNN precision context of the criterion's networks (InversionLoss.exact) - their BACKWARD convolutions run inside
    loss.backward(), outside the criterion's own forward, and must see the same setting."""
    if criterion is not None and hasattr(criterion, "conv_precision"):
        return criterion.conv_precision()
    return contextlib.nullcontext()


def _loss(recon, target, l2_lambda, extra_losses, criterion):
    if criterion is not None:
        return criterion(recon)
    loss = l2_lambda * F.mse_loss(recon, target) if l2_lambda > 0 else recon.new_zeros(())
    for weight, fn in extra_losses:
        loss = loss + weight * fn(recon, target)
    return loss


def _invert_graphed(net, target, onehot, style_vectors, steps, lr, l2_lambda, extra_losses, noise, stats=None, criterion=None):
    """The same loop with one optimisation step (zero_grad, forward, loss, backward, Adam) captured in a CUDA graph and
    replayed: at one face per GPU the eager loop is bound by ~400 kernel launches per step, not by the kernels.
    Adam only (capturable); fresh noise comes from the graph-safe CUDA generator, so every replay draws new noise."""
    latent = style_vectors.clone().detach().requires_grad_(True)
    opt = torch.optim.Adam([latent], lr=lr, capturable=True)
    static_loss = torch.zeros((), device=latent.device)
    static_recon = torch.empty_like(target)
    if criterion is not None:
        criterion.set_target(target)

    def one_step():
        opt.zero_grad(set_to_none=False)
        codes = net.cal_style_codes(latent)
        recon, _, _ = net.gen_img(None, codes, onehot, noise=noise) if noise is not None else net.gen_img(None, codes, onehot)
        loss = _loss(recon, target, l2_lambda, extra_losses, criterion)
        with _precision(criterion):
            loss.backward()
        opt.step()
        static_loss.copy_(loss.detach())
        static_recon.copy_(recon.detach())

    history = []
    warm = min(3, steps)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # eager warm-up: weight preparation, mask validation, allocator
        for _ in range(warm):
            one_step()
            history.append(static_loss.clone())
    torch.cuda.current_stream().wait_stream(side)
    if steps > warm:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):                   # capture records the step, it does not run it
            one_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps - warm):                   # `warm` eager steps + these replays = `steps` updates
            graph.replay()
            history.append(static_loss.clone())
        e1.record()
        if stats is not None:
            e1.synchronize()
            stats["replayed_steps"] = steps - warm
            stats["replay_ms_per_step"] = e0.elapsed_time(e1) / (steps - warm)
    return latent.detach(), static_recon, history
