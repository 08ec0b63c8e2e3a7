"""TEST INFRASTRUCTURE - CPU restatement (torch, fp32) of GPEN's FullGenerator forward: the
blind-face-restoration network that is stage 2 of every swap (scripts/face_swap.py:208).  It runs on the same three ops
as the E4S generator (modulated convolution, upfirdn2d, fused bias + leaky ReLU); what differs is that there is no region
mask, the encoder's feature maps are CONCATENATED to the activations as "noise", and a strided conv stack encodes the input.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.  Pinned against the unmodified
reference model (which runs on the CPU as shipped) by oracle/make_golden_gpen.py -> tests/golden/gpen_vectors.npz.

Reference citations are path:line under /root/reference/src/pretrained/gpen/face_model/gpen_model.py.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import e4s_oracle as O

Tensor = torch.Tensor


def channels(narrow: float = 1.0, channel_multiplier: int = 2) -> Dict[int, int]:
    """:402-413 / :635-646."""
    cm = channel_multiplier
    return {4: int(512 * narrow), 8: int(512 * narrow), 16: int(512 * narrow), 32: int(512 * narrow),
            64: int(256 * cm * narrow), 128: int(128 * cm * narrow), 256: int(64 * cm * narrow),
            512: int(32 * cm * narrow), 1024: int(16 * cm * narrow), 2048: int(8 * cm * narrow)}


def param_shapes(size: int, style_dim: int = 512, n_mlp: int = 8, narrow: float = 1.0, channel_multiplier: int = 2):
    """state_dict keys and shapes of FullGenerator(size, style_dim, n_mlp, isconcat=True), :621-667 and :381-466."""
    ch = channels(narrow, channel_multiplier)
    log_size = int(math.log2(size))
    s = {}
    for i in range(n_mlp):
        s[f"generator.style.{i + 1}.weight"] = (style_dim, style_dim)
        s[f"generator.style.{i + 1}.bias"] = (style_dim,)
    s["generator.input.input"] = (1, ch[4], 4, 4)

    def styled(prefix, cin, cout, up):
        s[prefix + ".conv.weight"] = (1, cout, cin, 3, 3)
        if up:
            s[prefix + ".conv.blur.kernel"] = (4, 4)
        s[prefix + ".conv.modulation.weight"] = (cin, style_dim)
        s[prefix + ".conv.modulation.bias"] = (cin,)
        s[prefix + ".noise.weight"] = (1,)
        s[prefix + ".activate.bias"] = (2 * cout,)

    def to_rgb(prefix, cin, up):
        s[prefix + ".bias"] = (1, 3, 1, 1)
        if up:
            s[prefix + ".upsample.kernel"] = (4, 4)
        s[prefix + ".conv.weight"] = (1, 3, cin, 1, 1)
        s[prefix + ".conv.modulation.weight"] = (cin, style_dim)
        s[prefix + ".conv.modulation.bias"] = (cin,)

    styled("generator.conv1", ch[4], ch[4], False)
    to_rgb("generator.to_rgb1", 2 * ch[4], False)
    cin = ch[4]
    for j, i in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** i]
        styled(f"generator.convs.{2 * j}", 2 * cin, cout, True)
        styled(f"generator.convs.{2 * j + 1}", 2 * cout, cout, False)
        to_rgb(f"generator.to_rgbs.{j}", 2 * cout, True)
        cin = cout
    s["ecd0.0.0.weight"] = (ch[size], 3, 1, 1)
    s["ecd0.0.1.bias"] = (ch[size],)
    cin = ch[size]
    for n, i in enumerate(range(log_size, 2, -1)):
        cout = ch[2 ** (i - 1)]
        s[f"ecd{n + 1}.0.0.kernel"] = (4, 4)
        s[f"ecd{n + 1}.0.1.weight"] = (cout, cin, 3, 3)
        s[f"ecd{n + 1}.0.2.bias"] = (cout,)
        cin = cout
    s["final_linear.0.weight"] = (style_dim, ch[4] * 16)
    s["final_linear.0.bias"] = (style_dim,)
    return s


def synthetic_state(size: int, salt: int = 0, **kw) -> Dict[str, Tensor]:
    """Seeded stand-in for the GPEN-BFR checkpoint (cannot be downloaded): e4s_oracle.synthetic_state's conventions,
    with the encoder's Blur buffers set to what Blur.__init__ registers (make_kernel([1,3,3,1]), :72-85)."""
    st = O.synthetic_state(param_shapes(size, **kw), salt)
    for k in st:
        if k.startswith("ecd") and k.endswith(".kernel"):
            st[k] = O.make_fir((1, 3, 3, 1), 1.0)
    return st


def encode(p: Dict[str, Tensor], x: Tensor, size: int) -> List[Tensor]:
    """The ecd0 ... ecd{L-2} stack of FullGenerator.forward (:677-682): returns every stage's output (the "noise" maps),
    finest first.  ConvLayer (:558-605): [Blur pad (2,2)] -> EqualConv2d (stride 2, no padding | 1x1) -> FusedLeakyReLU."""
    feats = []
    w0 = p["ecd0.0.0.weight"]
    h = F.conv2d(x, w0 * (1.0 / math.sqrt(w0.shape[1] * w0.shape[2] ** 2)))                 # EqualConv2d :107-125
    h = O.fused_leaky_relu(h, p["ecd0.0.1.bias"])
    feats.append(h)
    for n in range(1, int(math.log2(size)) - 1):
        w = p[f"ecd{n}.0.1.weight"]
        h = O.upfirdn2d(h, p[f"ecd{n}.0.0.kernel"], pad=(2, 2))                             # p = (4-2)+(3-1) = 4 -> (2, 2)  :573-578
        h = F.conv2d(h, w * (1.0 / math.sqrt(w.shape[1] * 9)), stride=2, padding=0)
        h = O.fused_leaky_relu(h, p[f"ecd{n}.0.2.bias"])
        feats.append(h)
    return feats


def styled_conv(p: Dict[str, Tensor], prefix: str, x: Tensor, style: Tensor, noise: Tensor, upsample: bool) -> Tensor:
    """StyledConv.forward with isconcat=True (:350-357): conv -> cat(out, weight * noise) -> FusedLeakyReLU(2C)."""
    out = O.modulated_conv2d(x, style, p[prefix + ".conv.weight"], p[prefix + ".conv.modulation.weight"],
                             p[prefix + ".conv.modulation.bias"], demodulate=True, upsample=upsample)
    out = torch.cat((out, p[prefix + ".noise.weight"] * noise), dim=1)                      # NoiseInjection :296-300
    return O.fused_leaky_relu(out, p[prefix + ".activate.bias"])


def to_rgb(p: Dict[str, Tensor], prefix: str, x: Tensor, style: Tensor, skip) -> Tensor:
    """ToRGB.forward (:370-379)."""
    out = O.modulated_conv2d(x, style, p[prefix + ".conv.weight"], p[prefix + ".conv.modulation.weight"],
                             p[prefix + ".conv.modulation.bias"], demodulate=False)
    out = out + p[prefix + ".bias"]
    if skip is not None:
        out = out + O.upfirdn2d(skip, p[prefix + ".upsample.kernel"], up=2, pad=(2, 1))     # Upsample :35-52
    return out


def full_generator_forward(p: Dict[str, Tensor], x: Tensor, size: int, n_mlp: int = 8, lr_mlp: float = 0.01) -> Tensor:
    """FullGenerator.forward (:669-688) -> Generator.forward (:497-555) with one style and the encoder's noise list."""
    feats = encode(p, x, size)
    z = feats[-1].reshape(x.shape[0], -1)
    w = O.equal_linear(z, p["final_linear.0.weight"], p["final_linear.0.bias"], activation=True)           # :667
    # noise = every encoder map twice, coarsest first, first entry dropped (:684-685)
    noise = [f for f in reversed(feats) for _ in (0, 1)][1:]
    w = w * torch.rsqrt(w.pow(2).mean(dim=1, keepdim=True) + 1e-8)                                         # PixelNorm :22-26
    for i in range(n_mlp):
        w = O.equal_linear(w, p[f"generator.style.{i + 1}.weight"], p[f"generator.style.{i + 1}.bias"], lr_mul=lr_mlp,
                           activation=True)
    out = p["generator.input.input"].repeat(x.shape[0], 1, 1, 1)
    out = styled_conv(p, "generator.conv1", out, w, noise[0], False)
    skip = to_rgb(p, "generator.to_rgb1", out, w, None)
    for j in range(int(math.log2(size)) - 2):
        out = styled_conv(p, f"generator.convs.{2 * j}", out, w, noise[1 + 2 * j], True)
        out = styled_conv(p, f"generator.convs.{2 * j + 1}", out, w, noise[2 + 2 * j], False)
        skip = to_rgb(p, f"generator.to_rgbs.{j}", out, w, skip)
    return skip
