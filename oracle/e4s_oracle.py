"""CPU oracle for the E4S synthesis / inversion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``e4s_b200/`` imports this file; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` as the checker, and by
``bench.py`` for the ``cpu_baseline`` / ``--impl reference`` legs.

It is a restatement, in plain torch-on-CPU (fp32 by default, fp64 on request),
of the algorithm the reference executes, one function per reference symbol, each
citing the reference ``file:line`` it follows (paths relative to
``/root/reference``).  It deliberately keeps the reference's *structure* - per
sample modulated weights, one full convolution per region, mask-multiply and
sum - because it doubles as the CPU baseline that is timed next to the GPU path.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against the reference itself, imported in the build
container by ``oracle/make_golden.py`` (which also writes ``tests/golden/``).
``tests/test_oracle_golden.py`` re-checks the oracle against those committed
fixtures without needing ``/root/reference``.

All tensors are NCHW like the reference's.  Parameters are passed as a flat
``state`` dict that uses the reference's checkpoint key names, so
a real E4S checkpoint can be fed to the oracle unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SQRT2 = math.sqrt(2.0)

# Channel table of the synthesis network, src/models/stylegan2/model.py:481-491
# (channel_multiplier = 2, the only value the reference uses).
CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


# --------------------------------------------------------------------------- FIR
def make_fir(taps: Sequence[float], gain: float = 1.0, dtype=torch.float32) -> Tensor:
    """2-D FIR from 1-D taps: outer product normalised to unit sum, times `gain`.

    Follows make_kernel, src/models/stylegan2/model.py:23-31; Blur / Upsample
    multiply by factor**2 (model.py:39, 84-85) - pass that as `gain`.
    """
    t = torch.tensor(list(taps), dtype=torch.float64)
    k2 = torch.outer(t, t)
    k2 = k2 / k2.sum()
    return (k2 * gain).to(dtype)


def upfirdn2d(x: Tensor, fir: Tensor, up: int = 1, down: int = 1, pad=(0, 0)) -> Tensor:
    """Zero-stuff by `up`, pad (negative = crop), TRUE convolution with `fir`, keep every `down`-th.

    Semantics of upfirdn2d(), src/models/stylegan2/op/upfirdn2d.py:142-147 and its
    CPU spelling upfirdn2d_native :150-184 (out size :100-101, kernel flip :174).
    x is [N, C, H, W]; pad=(p0, p1) is applied to both axes.
    """
    n, c, h, w = x.shape
    kh, kw = fir.shape
    p0, p1 = int(pad[0]), int(pad[1])
    z = x.new_zeros(n, c, h * up, w * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    hh, ww = z.shape[2], z.shape[3]
    z = z[:, :, max(-p0, 0): hh - max(-p1, 0), max(-p0, 0): ww - max(-p1, 0)]
    flipped = torch.flip(fir.to(x.dtype), [0, 1]).reshape(1, 1, kh, kw)
    hh, ww = z.shape[2], z.shape[3]
    y = F.conv2d(z.reshape(n * c, 1, hh, ww), flipped)
    y = y.reshape(n, c, hh - kh + 1, ww - kw + 1)
    return y[:, :, ::down, ::down]


def upfirdn2d_out_size(size: int, up: int, down: int, p0: int, p1: int, k: int) -> int:
    """src/models/stylegan2/op/upfirdn2d.py:100-101."""
    return (size * up + p0 + p1 - k) // down + 1


# ---------------------------------------------------------------- bias + activation
def fused_leaky_relu(x: Tensor, bias: Optional[Tensor], negative_slope: float = 0.2,
                     scale: float = SQRT2) -> Tensor:
    """scale * leaky_relu(x + bias[c]).  fused_act.py:84-85 -> fused_bias_act_kernel.cu:18-49
    (act=3, grad=0); CPU spelling in src/pretrained/gpen/face_model/op/fused_act.py:96."""
    if bias is not None:
        shape = [1, -1] + [1] * (x.ndim - 2)
        x = x + bias.reshape(shape)
    return F.leaky_relu(x, negative_slope) * scale


def fused_leaky_relu_backward(grad_out: Tensor, out: Tensor, negative_slope: float = 0.2,
                              scale: float = SQRT2):
    """grad wrt input and bias, fused_act.py:27-36 (act=3, grad=1, ref = forward output)."""
    gx = torch.where(out > 0, grad_out, grad_out * negative_slope) * scale
    dims = [0] + list(range(2, gx.ndim))
    return gx, gx.sum(dims)


# --------------------------------------------------------------------- linear layers
def equal_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], lr_mul: float = 1.0,
                 activation: bool = False) -> Tensor:
    """EqualLinear.forward, src/models/stylegan2/model.py:154-164 (scale at :151)."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    y = F.linear(x, weight * scale)
    if activation:
        return fused_leaky_relu(y, bias * lr_mul)
    if bias is not None:
        y = y + bias * lr_mul
    return y


# ---------------------------------------------------------------- modulated conv
def modulated_conv2d(x: Tensor, style: Tensor, weight: Tensor, mod_weight: Tensor, mod_bias: Tensor,
                     demodulate: bool = True, upsample: bool = False,
                     blur_taps: Sequence[float] = (1, 3, 3, 1)) -> Tensor:
    """ModulatedConv2d.forward, fused branch, src/models/stylegan2/model.py:276-320.

    weight is the stored parameter [1, Cout, Cin, k, k]; style is the W-space latent [B, 512].
    Per-sample weights are materialised exactly as the reference does (:277-281) and
    the grouped convolution (:295-297 / :314-316) is run one sample at a time.
    """
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)                      # :276  [B, Cin]
    wmod = (1.0 / math.sqrt(cin * k * k)) * weight * s.reshape(b, 1, cin, 1, 1)   # :277
    if demodulate:
        d = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)             # :280
        wmod = wmod * d.reshape(b, cout, 1, 1, 1)                      # :281
    outs = []
    for i in range(b):
        xi = x[i:i + 1]
        if upsample:
            # conv_transpose2d stride 2, no padding -> [2H+1, 2W+1], then Blur pad (1,1)  :287-300
            yi = F.conv_transpose2d(xi, wmod[i].transpose(0, 1), stride=2, padding=0)
            fir = make_fir(blur_taps, gain=4.0, dtype=x.dtype)         # :206-213, :84-85
            p = (len(blur_taps) - 2) - (k - 1)
            yi = upfirdn2d(yi, fir, pad=((p + 1) // 2 + 1, p // 2 + 1))
        else:
            yi = F.conv2d(xi, wmod[i], padding=k // 2)                 # :312-318
        outs.append(yi)
    return torch.cat(outs, 0)


def nearest_resize(mask: Tensor, size: int) -> Tensor:
    """F.interpolate(mask, size, mode='nearest'), model.py:391 / :430 / psp_encoders.py:265.
    Index rule of ATen's legacy 'nearest': src = min(floor(dst * in/out), in-1) in float32."""
    return F.interpolate(mask, size=(size, size), mode="nearest")


def styled_conv(x: Tensor, style: Tensor, mask: Optional[Tensor], noise: Tensor, p: Dict[str, Tensor],
                prefix: str, upsample: bool, mask_op: bool) -> Tensor:
    """StyledConv.forward, src/models/stylegan2/model.py:382-406.

    style is [B, ncls, 512] when mask_op else [B, 512].  `noise` must be given
    ([B or 1, 1, H, W]); fresh-noise sampling (:333) is the caller's business.
    """
    wk = dict(weight=p[prefix + "conv.weight"], mod_weight=p[prefix + "conv.modulation.weight"],
              mod_bias=p[prefix + "conv.modulation.bias"])
    if not mask_op:
        out = modulated_conv2d(x, style, upsample=upsample, **wk)
    else:
        hout = x.shape[2] * (2 if upsample else 1)
        seg = nearest_resize(mask, hout)                               # :391
        out = None
        for c in range(style.shape[1]):                                # :395-398
            oc = modulated_conv2d(x, style[:, c], upsample=upsample, **wk) * seg[:, c:c + 1]
            out = oc if out is None else out + oc
    out = out + p[prefix + "noise.weight"] * noise                     # :335
    return fused_leaky_relu(out, p[prefix + "activate.bias"])          # :404


def to_rgb(x: Tensor, style: Tensor, mask: Optional[Tensor], skip: Optional[Tensor], p: Dict[str, Tensor],
           prefix: str, mask_op: bool) -> Tensor:
    """ToRGB.forward, src/models/stylegan2/model.py:422-448 (1x1 modconv, no demod :417)."""
    wk = dict(weight=p[prefix + "conv.weight"], mod_weight=p[prefix + "conv.modulation.weight"],
              mod_bias=p[prefix + "conv.modulation.bias"], demodulate=False)
    if not mask_op:
        out = modulated_conv2d(x, style, **wk)
    else:
        seg = nearest_resize(mask, x.shape[2])                         # :430
        out = None
        for c in range(style.shape[1]):                                # :434-437
            oc = modulated_conv2d(x, style[:, c], **wk) * seg[:, c:c + 1]
            out = oc if out is None else out + oc
    out = out + p[prefix + "bias"]                                     # :441
    if skip is not None:
        fir = make_fir((1, 3, 3, 1), gain=4.0, dtype=x.dtype)          # Upsample, :34-53
        out = out + upfirdn2d(skip, fir, up=2, pad=(2, 1))             # :444-446
    return out


def generator_layer_plan(size: int, remaining_layer_idx: int):
    """Which StyledConv / ToRGB modules are built with mask_op, model.py:529,537,545."""
    log_size = int(math.log2(size))
    K = remaining_layer_idx
    conv_mask, rgb_mask = [], []
    for i in range(3, log_size + 1):
        conv_mask.append(not (i > 2 + K // 2))
        rgb_mask.append(not (K != 17 and i >= 2 + K // 2))
    return log_size, conv_mask, rgb_mask


def generator_forward(p: Dict[str, Tensor], codes: Tensor, mask: Tensor, noise: List[Tensor], size: int,
                      remaining_layer_idx: int = 13, split_layer_idx: int = 5, prefix: str = ""):
    """Generator.forward with input_is_latent=True and a 4-D latent, model.py:576-667.

    codes: [B, ncls, n_latent, 512]; mask: [B, ncls, Hm, Wm] float one-hot; noise: list of
    num_layers tensors.  Returns (image, intermediate_feats) like the reference's
    (image, None, intermediate_feats).
    """
    K = remaining_layer_idx
    log_size, conv_mask, rgb_mask = generator_layer_plan(size, K)
    b = codes.shape[0]
    out = p[prefix + "input.input"].repeat(b, 1, 1, 1)                                     # :630
    out = styled_conv(out, codes[:, :, 0], mask, noise[0], p, prefix + "conv1.", False, True)   # :631
    skip = to_rgb(out, codes[:, :, 1], mask, None, p, prefix + "to_rgb1.", True)           # :632
    feats = None
    i = 1
    for r in range(log_size - 2):
        c1, c2, tr = f"{prefix}convs.{2 * r}.", f"{prefix}convs.{2 * r + 1}.", f"{prefix}to_rgbs.{r}."
        n1, n2 = noise[1 + 2 * r], noise[2 + 2 * r]
        if i < K:                                                                          # :639-653
            st = codes[:, :, i] if conv_mask[r] else codes[:, 0, i]
            out = styled_conv(out, st, mask, n1, p, c1, True, conv_mask[r])
            if i + 2 == split_layer_idx:
                feats = out
            st = codes[:, :, i + 1] if conv_mask[r] else codes[:, 0, i + 1]
            out = styled_conv(out, st, mask, n2, p, c2, False, conv_mask[r])
            if K == 17 or i + 2 != K:
                skip = to_rgb(out, codes[:, :, i + 2], mask, skip, p, tr, rgb_mask[r])
            else:
                skip = to_rgb(out, codes[:, 0, i + 2], mask, skip, p, tr, rgb_mask[r])
        else:                                                                              # :655-657
            out = styled_conv(out, codes[:, 0, i], mask, n1, p, c1, True, conv_mask[r])
            out = styled_conv(out, codes[:, 0, i + 1], mask, n2, p, c2, False, conv_mask[r])
            skip = to_rgb(out, codes[:, 0, i + 2], mask, skip, p, tr, rgb_mask[r])
        i += 2
    return skip, feats


# ------------------------------------------------------------------- Net3 pieces
def cal_style_codes(p: Dict[str, Tensor], style_vectors: Tensor, latent_avg: Tensor,
                    remaining_layer_idx: int = 13) -> Tensor:
    """Net3.cal_style_codes (start_from_latent_avg, not learn_in_w), src/models/networks.py:135-158;
    LocalMLP :15-39 (EqualLinear -> LeakyReLU(0.01) -> EqualLinear)."""
    b, ncls, _ = style_vectors.shape
    K = remaining_layer_idx
    nw = K if K != 17 else 18
    per_cls = []
    for j in range(ncls):
        h = equal_linear(style_vectors[:, j], p[f"MLPs.{j}.mlp.0.weight"], p[f"MLPs.{j}.mlp.0.bias"])
        h = F.leaky_relu(h, 0.01)
        h = equal_linear(h, p[f"MLPs.{j}.mlp.2.weight"], p[f"MLPs.{j}.mlp.2.bias"])
        per_cls.append(h.reshape(b, nw, 512))
    codes = torch.stack(per_cls, 1)
    if K != 17:
        codes = codes + latent_avg[:K].reshape(1, 1, K, 512)
        rest = latent_avg[K:].reshape(1, 1, -1, 512).expand(b, ncls, -1, -1)
        return torch.cat([codes, rest], 2)
    return codes + latent_avg.reshape(1, 1, 18, 512)


def label_to_onehot(label: Tensor, num_cls: int) -> Tensor:
    """labelMap2OneHot, src/utils/torch_utils.py:166-172.  label: int64 [B,1,H,W]."""
    b, _, h, w = label.shape
    return torch.zeros(b, num_cls, h, w).scatter_(1, label, 1.0)


def region_mean(feats: Tensor, mask: Tensor) -> Tensor:
    """FSEncoder_PSP.get_per_comp_styleCode, src/models/encoders/psp_encoders.py:264-283:
    nearest-resize the mask to the feature size, then per (sample, class) the mean feature over
    the pixels whose (bool) mask is set; zero where the region is empty."""
    seg = nearest_resize(mask, feats.shape[2]).bool()
    b, c = feats.shape[:2]
    out = feats.new_zeros(b, seg.shape[1], c)
    for i in range(b):
        for j in range(seg.shape[1]):
            area = int(seg[i, j].sum())
            if area > 0:
                out[i, j] = feats[i][:, seg[i, j]].mean(1)
    return out


# ---------------------------------------------------------------- RGI encoder
def _instance_norm(x: Tensor) -> Tensor:
    """nn.InstanceNorm2d defaults (affine=False, eps=1e-5, biased variance)."""
    return F.instance_norm(x, eps=1e-5)


def encoder_unit_specs():
    """(in_channel, depth, stride) of the 24 bottleneck units, psp_encoders.py:242-247 + helpers.py:25."""
    specs = []
    for cin, depth, n in ((64, 128, 3), (128, 256, 4), (256, 512, 14), (512, 512, 3)):
        specs.append((cin, depth, 2))
        specs.extend([(depth, depth, 1)] * (n - 1))
    return specs


def encoder_unit(x: Tensor, p: Dict[str, Tensor], prefix: str, cin: int, depth: int, stride: int) -> Tensor:
    """bottleneck_IR_SE_Ours.forward, src/models/encoders/helpers.py:122-144 (SEModule :56-72)."""
    if cin == depth:
        short = x[:, :, ::stride, ::stride]                       # MaxPool2d(1, stride)
    else:
        short = _instance_norm(F.conv2d(x, p[prefix + "shortcut_layer.0.weight"], stride=stride))
    r = _instance_norm(x)
    r = F.conv2d(r, p[prefix + "res_layer.1.weight"], padding=1)
    r = F.prelu(r, p[prefix + "res_layer.2.weight"])
    r = F.conv2d(r, p[prefix + "res_layer.3.weight"], stride=stride, padding=1)
    r = _instance_norm(r)
    g = r.mean((2, 3), keepdim=True)
    g = F.relu(F.conv2d(g, p[prefix + "res_layer.5.fc1.weight"]))
    g = torch.sigmoid(F.conv2d(g, p[prefix + "res_layer.5.fc2.weight"]))
    return r * g + short


def encoder_forward(p: Dict[str, Tensor], img256: Tensor, mask: Tensor, prefix: str = "encoder."):
    """FSEncoder_PSP.forward, src/models/encoders/psp_encoders.py:285-309.  img256 is the already
    bilinearly resized [B,3,256,256] input (networks.py:131).  Returns ([B,ncls,1280], zeros)."""
    x = F.conv2d(img256, p[prefix + "input_layer.0.weight"], padding=1)
    x = F.prelu(_instance_norm(x), p[prefix + "input_layer.2.weight"])
    taps = {}
    for i, (cin, depth, stride) in enumerate(encoder_unit_specs()):
        x = encoder_unit(x, p, f"{prefix}body.{i}.", cin, depth, stride)
        if i in (6, 20, 23):
            taps[i] = x
    codes = torch.cat([region_mean(taps[6], mask), region_mean(taps[20], mask), region_mean(taps[23], mask)], 2)
    return codes, torch.zeros_like(x)


def get_style_vectors(p: Dict[str, Tensor], img: Tensor, mask: Tensor):
    """Net3.get_style_vectors, src/models/networks.py:121-133."""
    return encoder_forward(p, F.interpolate(img, (256, 256), mode="bilinear"), mask)


# ------------------------------------------------------- synthetic parameters
def _key_seed(key: str) -> int:
    h = 2166136261
    for ch in key.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def synthetic_state(shapes: Dict[str, Sequence[int]], salt: int = 0) -> Dict[str, Tensor]:
    """Deterministic stand-in for a checkpoint (none ships with the reference).

    Every tensor is drawn from its own generator seeded by an FNV hash of its key, so the
    reference model (in the build container) and this repo's modules (on the GPU box) can be
    loaded with bit-identical parameters regardless of construction order.  Conventions follow
    the reference's initialisers (randn weights, modulation bias 1) except that parameters the
    reference zero-initialises (noise.weight, activate.bias, ToRGB bias) get small non-zero
    values so those code paths are exercised.
    """
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = torch.Generator().manual_seed(_key_seed(key) ^ salt)
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if key.endswith("modulation.bias"):
            t = 1.0 + 0.1 * t
        elif key.endswith("noise.weight") or key.endswith("activate.bias") or key.endswith(".bias"):
            t = 0.1 * t
        elif key.endswith("blur.kernel") or key.endswith("upsample.kernel"):
            t = make_fir((1, 3, 3, 1), gain=4.0)
        elif "encoder." in key and key.endswith(".weight") and len(shape) == 4:
            t = t * (1.0 / math.sqrt(shape[1] * shape[2] * shape[3]))   # keep the conv stack O(1)
        elif "encoder." in key and key.endswith(".weight") and len(shape) == 1:
            t = 0.25 + 0.05 * t                                          # PReLU slopes
        out[key] = t
    return out


def generator_param_shapes(size: int, style_dim: int = 512, prefix: str = "") -> Dict[str, tuple]:
    """Names/shapes of the synthesis-network parameters and buffers (SURVEY.md section 5 checkpoint
    contract; model.py:493-552).  The 8-layer mapping network `style.*` is not on the path."""
    log_size = int(math.log2(size))
    sh: Dict[str, tuple] = {prefix + "input.input": (1, CHANNELS[4], 4, 4)}

    def conv(name, cin, cout, k, up):
        sh[name + "conv.weight"] = (1, cout, cin, k, k)
        sh[name + "conv.modulation.weight"] = (cin, style_dim)
        sh[name + "conv.modulation.bias"] = (cin,)
        if up:
            sh[name + "conv.blur.kernel"] = (4, 4)

    def styled(name, cin, cout, up):
        conv(name, cin, cout, 3, up)
        sh[name + "noise.weight"] = (1,)
        sh[name + "activate.bias"] = (cout,)

    def rgb(name, cin, up):
        conv(name, cin, 3, 1, False)
        sh[name + "bias"] = (1, 3, 1, 1)
        if up:
            sh[name + "upsample.kernel"] = (4, 4)

    styled(prefix + "conv1.", CHANNELS[4], CHANNELS[4], False)
    rgb(prefix + "to_rgb1.", CHANNELS[4], False)
    cin = CHANNELS[4]
    for r, i in enumerate(range(3, log_size + 1)):
        cout = CHANNELS[2 ** i]
        styled(f"{prefix}convs.{2 * r}.", cin, cout, True)
        styled(f"{prefix}convs.{2 * r + 1}.", cout, cout, False)
        rgb(f"{prefix}to_rgbs.{r}.", cout, True)
        cin = cout
    for layer in range((log_size - 2) * 2 + 1):
        res = 2 ** ((layer + 5) // 2)
        sh[f"{prefix}noises.noise_{layer}"] = (1, 1, res, res)
    return sh


def mlp_param_shapes(num_cls: int, remaining_layer_idx: int = 13) -> Dict[str, tuple]:
    nw = remaining_layer_idx if remaining_layer_idx != 17 else 18
    sh = {}
    for j in range(num_cls):
        sh[f"MLPs.{j}.mlp.0.weight"] = (512, 1280)
        sh[f"MLPs.{j}.mlp.0.bias"] = (512,)
        sh[f"MLPs.{j}.mlp.2.weight"] = (512 * nw, 512)
        sh[f"MLPs.{j}.mlp.2.bias"] = (512 * nw,)
    return sh


def encoder_param_shapes(prefix: str = "encoder.") -> Dict[str, tuple]:
    sh = {prefix + "input_layer.0.weight": (64, 3, 3, 3), prefix + "input_layer.2.weight": (64,)}
    for i, (cin, depth, _s) in enumerate(encoder_unit_specs()):
        u = f"{prefix}body.{i}."
        if cin != depth:
            sh[u + "shortcut_layer.0.weight"] = (depth, cin, 1, 1)
        sh[u + "res_layer.1.weight"] = (depth, cin, 3, 3)
        sh[u + "res_layer.2.weight"] = (depth,)
        sh[u + "res_layer.3.weight"] = (depth, depth, 3, 3)
        sh[u + "res_layer.5.fc1.weight"] = (depth // 16, depth, 1, 1)
        sh[u + "res_layer.5.fc2.weight"] = (depth, depth // 16, 1, 1)
    return sh


def synthetic_inputs(batch: int, num_cls: int, size: int, mask_size: int, seed: int = 1, kind: str = "blobs"):
    """Seeded codes / one-hot mask / noise list for a synthesis call.

    kind='blobs' gives piecewise-constant label maps with a handful of regions (face-mask like);
    kind='iid' gives independent uniform labels per pixel (every tile sees every class)."""
    g = torch.Generator().manual_seed(seed)
    log_size = int(math.log2(size))
    n_latent = 2 * log_size - 2
    codes = torch.randn(batch, num_cls, n_latent, 512, generator=g)
    if kind == "iid":
        label = torch.randint(0, num_cls, (batch, 1, mask_size, mask_size), generator=g)
    else:
        coarse = max(4, mask_size // 16)
        lab = torch.randint(0, num_cls, (batch, 1, coarse, coarse), generator=g).float()
        label = F.interpolate(lab, size=(mask_size, mask_size), mode="nearest").long()
    mask = label_to_onehot(label, num_cls)
    noise = [torch.randn(batch, 1, 4, 4, generator=g)]
    for i in range(3, log_size + 1):
        noise.append(torch.randn(batch, 1, 2 ** i, 2 ** i, generator=g))
        noise.append(torch.randn(batch, 1, 2 ** i, 2 ** i, generator=g))
    return codes, mask, label, noise


# ------------------------------------------------------------- replayable golden-case inputs
MODCONV_CASES = {"plain": (24, 16, 3, True, False, 9), "up": (16, 24, 3, True, True, 6), "rgb": (24, 3, 1, False, False, 8)}
STYLEDCONV_CASES = {"plain": (16, 24, False, 8), "up": (24, 16, True, 8)}


def case_generator(tag: str) -> torch.Generator:
    return torch.Generator().manual_seed(_key_seed("golden-case/" + tag))


def modconv_case(tag: str):
    """(x [2,cin,hw,hw], latent [2,512]) of golden case modconv/<tag>; shared by make_golden.py and the tests."""
    cin, _cout, _k, _demod, _up, hw = MODCONV_CASES[tag]
    g = case_generator("modconv/" + tag)
    return torch.randn(2, cin, hw, hw, generator=g), torch.randn(2, 512, generator=g)


def styledconv_case(tag: str):
    cin, _cout, up, hw = STYLEDCONV_CASES[tag]
    g = case_generator("styledconv/" + tag)
    hout = 2 * hw if up else hw
    codes, mask, _label, _ = synthetic_inputs(2, 5, 16, 32, seed=3)
    return torch.randn(2, cin, hw, hw, generator=g), torch.randn(2, 1, hout, hout, generator=g), codes, mask


def torgb_case():
    g = case_generator("torgb")
    codes, mask, _label, _ = synthetic_inputs(2, 5, 16, 32, seed=3)
    return torch.randn(2, 24, 16, 16, generator=g), torch.randn(2, 3, 8, 8, generator=g), codes, mask


def net3_case():
    """(style vectors [2,12,1280], latent_avg [18,512], image [1,3,320,320], mask [1,12,256,256])."""
    g = case_generator("net3")
    sv = torch.randn(2, 12, 1280, generator=g)
    latent_avg = 0.5 * torch.randn(18, 512, generator=g)
    img = torch.randn(1, 3, 320, 320, generator=g)
    _, mask, _, _ = synthetic_inputs(1, 12, 64, 256, seed=21)
    return sv, latent_avg, img, mask


def region_mean_case():
    g = case_generator("region_mean")
    feats = torch.randn(2, 70, 16, 16, generator=g)
    lab = torch.randint(0, 3, (2, 1, 32, 32), generator=g)     # classes 3, 4 stay empty
    return feats, label_to_onehot(lab, 5)
