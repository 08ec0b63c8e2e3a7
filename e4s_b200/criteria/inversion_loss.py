"""The inversion loop's loss, ``Optimizer.calc_loss`` (scripts/optimization.py:88-122):

    loss = id_lambda * ID(recon, img) + l2_lambda * mse(recon, img)
         + lpips_lambda * sum_{i=0..2} LPIPS(pool(recon, 1024 / 2^i), pool(img, 1024 / 2^i)) + face_parsing_lambda * Parsing(recon, img)

with the reference's default weights (src/options/optim_options.py:44-48: 0.1 / 1.0 / 0.8 / 0.1).

What differs from the reference's execution (results are the same, tests/test_losses.py):
* the TARGET image's features - LPIPS activations at three scales, ArcFace features, parsing-net features - are computed
  once in ``set_target`` instead of in every step (the reference re-runs all three networks on the fixed image each step:
  lpips.py:30, id_loss.py:33, face_parsing_loss.py:55): half of the loss networks' forward work disappears;
* for 1024x1024 images the five adaptive poolings per image (LPIPS 512 / 256, ID 256, parsing 512) are ONE pyramid kernel
  (csrc/pool.cu) and one backward kernel;
* no host synchronisation (the reference calls float() on per-sample similarities in Python loops, id_loss.py:47-50), so the
  whole optimisation step - generator, loss networks, backward, Adam - can be captured in one CUDA graph.
"""
from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .. import kernels as K
from .face_parsing import FaceParsingLoss
from .id_loss import IDLoss
from .lpips import LPIPS


class conv_precision:
    """Context manager: cuDNN convolutions in full fp32 (exact=True: the parity setting) or TF32 (what torch and the reference
    use by default).  Sets both the legacy switch (torch.backends.cudnn.allow_tf32) and, where this torch has it, the newer
    torch.backends.cudnn.conv.fp32_precision - with only the legacy one, cuDNN still picked TF32-class algorithms for some
    layers (AlexNet's 5x5 convolution: 2e-3 against 1e-6, gpurun diagnostics of round 2)."""

    def __init__(self, exact: bool):
        self.exact = exact

    def __enter__(self):
        # new API (torch >= 2.9): per-operator switches; conv and rnn are set TOGETHER - torch raises on any legacy-flag read when
        # they differ ("mix of the legacy and new APIs")
        self.saved = []
        mods = [getattr(torch.backends.cudnn, n, None) for n in ("conv", "rnn")]
        mods = [m for m in mods if m is not None and hasattr(m, "fp32_precision")]
        if mods:
            for m in mods:
                self.saved.append((m, m.fp32_precision))
                m.fp32_precision = "ieee" if self.exact else "tf32"
        else:
            self.saved.append((None, torch.backends.cudnn.allow_tf32))
            torch.backends.cudnn.allow_tf32 = not self.exact
        return self

    def __exit__(self, *exc):
        for m, v in self.saved:
            if m is None:
                torch.backends.cudnn.allow_tf32 = v
            else:
                m.fp32_precision = v
        return False


class _Pyramid(torch.autograd.Function):
    """(x) -> (2x2 block means, 4x4 block means) with a fused backward."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        y2, y4 = K.avgpool_pyramid(x)
        return y2, y4

    @staticmethod
    def backward(ctx, g2, g4):
        return K.avgpool_pyramid_bwd(None, g2, g4, ctx.shape)


def pyramid(x: torch.Tensor, sizes: Sequence[int]):
    """[adaptive_avg_pool2d(x, s) for s in sizes]; one kernel when sizes == (H, H/2, H/4) on a CUDA tensor."""
    h, w = x.shape[2:]
    if (x.is_cuda and h == w and tuple(sizes) == (h, h // 2, h // 4) and h % 4 == 0 and w % 8 == 0
            and x.dtype == torch.float32):
        y2, y4 = _Pyramid.apply(x.contiguous())
        return [x, y2, y4]
    return [x if (h, w) == (s, s) else F.adaptive_avg_pool2d(x, (s, s)) for s in sizes]


class InversionLoss(nn.Module):
    def __init__(self, opts=None, id_lambda: float = 0.1, l2_lambda: float = 1.0, lpips_lambda: float = 0.8,
                 face_parsing_lambda: float = 0.1, lpips_sizes: Sequence[int] = (1024, 512, 256), exact: bool = True):
        """opts: the reference's option object (ir_se50_path, face_parsing_model_path, id_loss_multiscale; optional).
        exact=True keeps the loss networks' convolutions in full fp32 (TF32 off) - the parity setting."""
        super().__init__()
        opts = opts if opts is not None else SimpleNamespace(id_loss_multiscale=True)
        self.id_lambda, self.l2_lambda = id_lambda, l2_lambda
        self.lpips_lambda, self.face_parsing_lambda = lpips_lambda, face_parsing_lambda
        self.lpips_sizes, self.exact = tuple(lpips_sizes), exact
        self.lpips_loss = LPIPS(net_type="alex").eval() if lpips_lambda > 0 else None
        self.id_loss = IDLoss(opts).eval() if id_lambda > 0 else None
        self.face_parsing_loss = FaceParsingLoss(opts).eval() if face_parsing_lambda > 0 else None
        for p in self.parameters():
            p.requires_grad = False
        self._target = None

    def conv_precision(self):
        """Context for code that differentiates through this loss (``loss.backward()``): the backward convolutions of the three
        networks must run at the same precision as their forward ones (e4s_b200.optimization.invert enters it)."""
        return conv_precision(self.exact)

    # -------------------------------------------------------------------------------------------------- features
    def _views(self, img: torch.Tensor):
        """The inputs the three networks start from: the LPIPS pyramid and, where a level already is the 512 / 256 pooling the
        parsing / identity loss wants, that level (no second pooling pass)."""
        pyr = pyramid(img, self.lpips_sizes)
        by_size = {int(t.shape[2]): t for t in pyr if t.shape[2] == t.shape[3]}
        return pyr, by_size.get(256, img), by_size.get(512, img)

    def _features(self, img: torch.Tensor) -> Dict[str, object]:
        pyr, for_id, for_parsing = self._views(img)
        out: Dict[str, object] = {}
        with conv_precision(self.exact):
            if self.lpips_loss is not None:
                out["lpips"] = [self.lpips_loss.features(t) for t in pyr]
            if self.id_loss is not None:
                out["id"] = self.id_loss.extract_feats(for_id)
            if self.face_parsing_loss is not None:
                out["parsing"] = self.face_parsing_loss.extract_feats(for_parsing)
        return out

    def set_target(self, img: torch.Tensor) -> None:
        """Cache everything that depends on the target image alone."""
        with torch.no_grad():
            feats = self._features(img)
            self._target = {"img": img.detach(), "lpips": feats.get("lpips"),
                            "id": None if "id" not in feats else [f.detach() for f in feats["id"]],
                            "parsing": None if "parsing" not in feats else [f.detach() for f in feats["parsing"]]}

    # ------------------------------------------------------------------------------------------------------ loss
    def forward(self, img_recon: torch.Tensor, img: Optional[torch.Tensor] = None, return_terms: bool = False):
        """Loss of a reconstruction against `img` (or against the cached target when img is None)."""
        if img is not None:
            self.set_target(img)
        tgt = self._target
        if tgt is None:
            raise RuntimeError("InversionLoss: call set_target(img) first or pass img")
        feats = self._features(img_recon)
        n = img_recon.shape[0]
        terms = {}
        loss = img_recon.new_zeros(())
        if self.id_loss is not None:
            terms["loss_id"] = IDLoss.loss_from_feats(feats["id"], tgt["id"])
            loss = loss + terms["loss_id"] * self.id_lambda
        if self.l2_lambda > 0:
            terms["loss_l2"] = F.mse_loss(img_recon, tgt["img"])
            loss = loss + terms["loss_l2"] * self.l2_lambda
        if self.lpips_loss is not None:
            lp = img_recon.new_zeros(())
            for fx, fy in zip(feats["lpips"], tgt["lpips"]):
                lp = lp + self.lpips_loss.distance(fx, fy, n)
            terms["loss_lpips"] = lp
            loss = loss + lp * self.lpips_lambda
        if self.face_parsing_loss is not None:
            terms["loss_face_parsing"] = FaceParsingLoss.loss_from_feats(feats["parsing"], tgt["parsing"])
            loss = loss + terms["loss_face_parsing"] * self.face_parsing_lambda
        return (loss, terms) if return_terms else loss
