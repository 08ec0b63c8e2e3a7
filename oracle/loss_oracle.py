"""CPU restatement (functional torch, fp32) of the loss side of the inversion loop - TEST INFRASTRUCTURE, not product code:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

Follows, with file:line of /root/reference:
  calc_loss            scripts/optimization.py:88-122 (0.1 ID + 1.0 L2 + 0.8 sum_3 LPIPS + 0.1 parsing by default, optim_options.py:44-48)
  lpips                src/criteria/lpips/lpips.py:29-35, networks.py:42-83 (AlexNet features 1..12), utils.py:6-8
  id_loss              src/criteria/id_loss.py:25-57, src/models/encoders/model_irse.py:9-69 (IR-SE50, multi-scale taps after
                       units 2, 6, 20, 23), helpers.py:15-18,56-72,97-119
  parsing_loss         src/criteria/face_parsing/face_parsing_loss.py:46-78, unet.py:71-92, model_utils.py:177-203

PINNED: oracle/make_golden_losses.py imports the real reference classes (checkpoint downloads replaced by seeded state
dicts; the shipped parsing checkpoint used as is), asserts oracle == reference and stores the reference's outputs in
tests/golden/loss_vectors.npz.  State dicts use the reference modules' own keys.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _key_seed(key: str) -> int:
    h = 2166136261
    for ch in key.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def synthetic_loss_state(shapes: Dict[str, Sequence[int]], salt: int = 0, prelu_keys: Sequence[str] = ()) -> Dict[str, Tensor]:
    """Seeded stand-in for the loss networks' checkpoints (AlexNet + LPIPS linear layers and IR-SE50 cannot be downloaded).
    Convolution / linear weights ~ N(0, 2 / fan_in) in the ReLU networks (AlexNet, UNet) and N(0, 1 / fan_in) in the residual
    IR-SE50, whose last BatchNorm of every residual branch gets weight 0.25 (1 + 0.1 n): a plain He-initialised 24-unit residual
    stack doubles its activation variance per unit (|x| ~ 2500 at the output) and its input gradient then amplifies fp32
    rounding differences between devices to O(1) - a stand-in must be as well conditioned as a trained network.  BatchNorm
    otherwise weight 1 + 0.1 n, bias 0.1 n, running_mean 0.1 n, running_var 1 + 0.1 |n|; PReLU slopes 0.25 + 0.05 n; LPIPS
    linear weights |n| / C (non-negative like the trained ones).  Each tensor has its own generator seeded by a hash of its key."""
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = torch.Generator().manual_seed(_key_seed(key) ^ salt)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        if key.endswith("net.mean"):
            out[key] = torch.tensor([-.030, -.088, -.188]).reshape(shape)
            continue
        if key.endswith("net.std"):
            out[key] = torch.tensor([.458, .448, .450]).reshape(shape)
            continue
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if key.endswith("running_var"):
            t = 1.0 + 0.1 * t.abs()
        elif key.endswith("running_mean"):
            t = 0.1 * t
        elif key.startswith("lin.") or ".lin." in key:
            t = t.abs() / shape[1]
        elif len(shape) >= 2:                                        # conv / linear weight
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * math.sqrt((1.0 if key.startswith("facenet.") else 2.0) / fan_in)
        elif key.endswith(".bias"):
            t = 0.1 * t
        elif key in prelu_keys:                                      # PReLU slopes share the '.weight' suffix with BatchNorm scales
            t = 0.25 + 0.05 * t
        elif key.endswith("res_layer.4.weight"):                     # last BatchNorm of an IR-SE residual branch
            t = 0.25 * (1.0 + 0.1 * t)
        elif key.endswith(".weight") and len(shape) == 1:            # BatchNorm scale
            t = 1.0 + 0.1 * t
        out[key] = t
    return out


# ------------------------------------------------------------------------------------------------ LPIPS (AlexNet)
def normalize_activation(x: Tensor, eps: float = 1e-10) -> Tensor:          # utils.py:6-8
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True) + 1e-16) + eps)


def alexnet_features(sd: Dict[str, Tensor], x: Tensor, prefix: str = "net.") -> List[Tensor]:
    """networks.py:42-83: z-score, torchvision alexnet.features; outputs after ReLU 1..5, unit-normalised over channels."""
    x = (x - sd[prefix + "mean"]) / sd[prefix + "std"]
    L = prefix + "layers."
    out = []
    x = F.relu(F.conv2d(x, sd[L + "0.weight"], sd[L + "0.bias"], stride=4, padding=2))
    out.append(normalize_activation(x))
    x = F.max_pool2d(x, 3, 2)
    x = F.relu(F.conv2d(x, sd[L + "3.weight"], sd[L + "3.bias"], padding=2))
    out.append(normalize_activation(x))
    x = F.max_pool2d(x, 3, 2)
    x = F.relu(F.conv2d(x, sd[L + "6.weight"], sd[L + "6.bias"], padding=1))
    out.append(normalize_activation(x))
    x = F.relu(F.conv2d(x, sd[L + "8.weight"], sd[L + "8.bias"], padding=1))
    out.append(normalize_activation(x))
    x = F.relu(F.conv2d(x, sd[L + "10.weight"], sd[L + "10.bias"], padding=1))
    out.append(normalize_activation(x))
    return out


def lpips(sd: Dict[str, Tensor], x: Tensor, y: Tensor) -> Tensor:
    """lpips.py:29-35."""
    fx, fy = alexnet_features(sd, x), alexnet_features(sd, y)
    res = [F.conv2d((a - b) ** 2, sd[f"lin.{i}.1.weight"]).mean((2, 3), True) for i, (a, b) in enumerate(zip(fx, fy))]
    return torch.sum(torch.cat(res, 0)) / x.shape[0]


# ------------------------------------------------------------------------------------------- identity (IR-SE50)
def _bn(sd, key, x, eps=1e-5):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"], False, 0.0, eps)


IRSE50_UNITS = [(64, 64, 2)] + [(64, 64, 1)] * 2 + [(64, 128, 2)] + [(128, 128, 1)] * 3 + [(128, 256, 2)] + [(256, 256, 1)] * 13 + \
               [(256, 512, 2)] + [(512, 512, 1)] * 2                       # helpers.py:25-36 (in_channel, depth, stride)


def l2_norm(x: Tensor, axis: int = 1) -> Tensor:                            # helpers.py:15-18
    return x / torch.norm(x, 2, axis, True)


def irse50_features(sd: Dict[str, Tensor], x: Tensor, multi_scale: bool = True, prefix: str = "facenet.") -> List[Tensor]:
    """model_irse.py:44-69 with bottleneck_IR_SE units (helpers.py:97-119) and SEModule (helpers.py:56-72), eval mode."""
    p = prefix
    x = F.conv2d(x, sd[p + "input_layer.0.weight"], None, 1, 1)
    x = F.prelu(_bn(sd, p + "input_layer.1", x), sd[p + "input_layer.2.weight"])
    taps = []
    for i, (cin, depth, stride) in enumerate(IRSE50_UNITS):
        u = f"{p}body.{i}."
        if cin == depth:
            sc = F.max_pool2d(x, 1, stride)
        else:
            sc = _bn(sd, u + "shortcut_layer.1", F.conv2d(x, sd[u + "shortcut_layer.0.weight"], None, stride))
        r = _bn(sd, u + "res_layer.0", x)
        r = F.prelu(F.conv2d(r, sd[u + "res_layer.1.weight"], None, 1, 1), sd[u + "res_layer.2.weight"])
        r = _bn(sd, u + "res_layer.4", F.conv2d(r, sd[u + "res_layer.3.weight"], None, stride, 1))
        gate = F.adaptive_avg_pool2d(r, 1)
        gate = torch.sigmoid(F.conv2d(F.relu(F.conv2d(gate, sd[u + "res_layer.5.fc1.weight"])), sd[u + "res_layer.5.fc2.weight"]))
        x = r * gate + sc
        if multi_scale and i in (2, 6, 20, 23):
            taps.append(l2_norm(x.reshape(x.size(0), -1)))
    x = _bn(sd, p + "output_layer.0", x)                                   # Dropout is the identity in eval mode
    x = F.linear(x.reshape(x.size(0), -1), sd[p + "output_layer.3.weight"], sd[p + "output_layer.3.bias"])
    x = F.batch_norm(x, sd[p + "output_layer.4.running_mean"], sd[p + "output_layer.4.running_var"], sd[p + "output_layer.4.weight"],
                     sd[p + "output_layer.4.bias"], False, 0.0, 1e-5)
    return taps + [l2_norm(x)]


def id_extract_feats(sd, x, multi_scale=True):
    """id_loss.py:25-30."""
    x = F.adaptive_avg_pool2d(x, (256, 256)) if x.shape[2] != 256 else x
    x = x[:, :, 35:223, 32:220]
    x = F.adaptive_avg_pool2d(x, (112, 112))
    return irse50_features(sd, x, multi_scale)


def _cos_loss(y_hat_feats_ms, y_feats_ms):
    """id_loss.py:38-55 == face_parsing_loss.py:59-76: per scale, mean over samples of 1 - <y_hat_i, y_i>."""
    loss_all = 0
    for a, b in zip(y_hat_feats_ms, y_feats_ms):
        loss = 0
        for i in range(a.shape[0]):
            loss = loss + (1 - a[i].dot(b[i]))
        loss_all = loss_all + loss / a.shape[0]
    return loss_all


def id_loss(sd, y_hat, y, multi_scale=True):
    return _cos_loss(id_extract_feats(sd, y_hat, multi_scale), [f.detach() for f in id_extract_feats(sd, y, multi_scale)])


# ---------------------------------------------------------------------------------------------- parsing (UNet)
def _unet_conv2(sd, key, x):
    """model_utils.py:177-203 with batch norm."""
    for c in ("conv1", "conv2"):
        x = F.relu(_bn(sd, f"{key}.{c}.1", F.conv2d(x, sd[f"{key}.{c}.0.weight"], sd[f"{key}.{c}.0.bias"], 1, 1)))
    return x


def unet_extract_feats(sd: Dict[str, Tensor], x: Tensor, prefix: str = "G.") -> List[Tensor]:
    """unet.py:71-92."""
    c1 = _unet_conv2(sd, prefix + "conv1", x)
    c2 = _unet_conv2(sd, prefix + "conv2", F.max_pool2d(c1, 2))
    c3 = _unet_conv2(sd, prefix + "conv3", F.max_pool2d(c2, 2))
    c4 = _unet_conv2(sd, prefix + "conv4", F.max_pool2d(c3, 2))
    ce = _unet_conv2(sd, prefix + "center", F.max_pool2d(c4, 2))
    bs = x.size(0)
    return [l2_norm(t.reshape(bs, -1)) for t in (c1, c2, c3, c4, ce)]


def parsing_extract_feats(sd, x):
    """face_parsing_loss.py:46-49."""
    x = F.adaptive_avg_pool2d(x, (512, 512)) if x.shape[2] != 512 else x
    return unet_extract_feats(sd, x)


def parsing_loss(sd, y_hat, y):
    return _cos_loss(parsing_extract_feats(sd, y_hat), [f.detach() for f in parsing_extract_feats(sd, y)])


# ----------------------------------------------------------------------------------------------------- calc_loss
def calc_loss(states: Dict[str, Dict[str, Tensor]], img: Tensor, img_recon: Tensor, id_lambda=0.1, l2_lambda=1.0, lpips_lambda=0.8,
              face_parsing_lambda=0.1, sizes=(1024, 512, 256)):
    """scripts/optimization.py:88-122.  states: {"lpips": ..., "id": ..., "parsing": ...} state dicts.  Returns (loss, terms)."""
    terms = {}
    loss = 0.0
    if id_lambda > 0:
        terms["loss_id"] = id_loss(states["id"], img_recon, img)
        loss = loss + terms["loss_id"] * id_lambda
    if l2_lambda > 0:
        terms["loss_l2"] = F.mse_loss(img_recon, img)
        loss = loss + terms["loss_l2"] * l2_lambda
    if lpips_lambda > 0:
        lp = 0
        for s in sizes:
            lp = lp + lpips(states["lpips"], F.adaptive_avg_pool2d(img_recon, (s, s)), F.adaptive_avg_pool2d(img, (s, s)))
        terms["loss_lpips"] = lp
        loss = loss + lp * lpips_lambda
    if face_parsing_lambda > 0:
        terms["loss_face_parsing"] = parsing_loss(states["parsing"], img_recon, img)
        loss = loss + terms["loss_face_parsing"] * face_parsing_lambda
    return loss, terms


def golden_inputs():
    """The seeded image triple of tests/golden/loss_vectors.npz: a target in [-1, 1], a reconstruction close to it (what an
    inversion step sees) and an unrelated one, each [2, 3, 256, 256]."""
    g = torch.Generator().manual_seed(5)
    B, S = 2, 256
    img = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    recon = (img + 0.15 * torch.randn(B, 3, S, S, generator=g)).clamp(-1, 1)
    far = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    return img, recon, far


def loss_states(salt: int = 0):
    """Seeded state dicts for the three loss modules, keyed like e4s_b200.criteria / the reference (shapes from the architecture)."""
    shapes_lpips = {"net.mean": (1, 3, 1, 1), "net.std": (1, 3, 1, 1)}
    for idx, (co, ci, k) in zip((0, 3, 6, 8, 10), ((64, 3, 11), (192, 64, 5), (384, 192, 3), (256, 384, 3), (256, 256, 3))):
        shapes_lpips[f"net.layers.{idx}.weight"] = (co, ci, k, k)
        shapes_lpips[f"net.layers.{idx}.bias"] = (co,)
    for i, nc in enumerate((64, 192, 384, 256, 256)):
        shapes_lpips[f"lin.{i}.1.weight"] = (1, nc, 1, 1)

    def bn(d, key, c):
        d[key + ".weight"], d[key + ".bias"], d[key + ".running_mean"], d[key + ".running_var"] = (c,), (c,), (c,), (c,)
        d[key + ".num_batches_tracked"] = ()

    sid, prelu = {}, []
    p = "facenet."
    sid[p + "input_layer.0.weight"] = (64, 3, 3, 3)
    bn(sid, p + "input_layer.1", 64)
    sid[p + "input_layer.2.weight"] = (64,)
    prelu.append(p + "input_layer.2.weight")
    for i, (cin, depth, stride) in enumerate(IRSE50_UNITS):
        u = f"{p}body.{i}."
        if cin != depth:
            sid[u + "shortcut_layer.0.weight"] = (depth, cin, 1, 1)
            bn(sid, u + "shortcut_layer.1", depth)
        bn(sid, u + "res_layer.0", cin)
        sid[u + "res_layer.1.weight"] = (depth, cin, 3, 3)
        sid[u + "res_layer.2.weight"] = (depth,)
        prelu.append(u + "res_layer.2.weight")
        sid[u + "res_layer.3.weight"] = (depth, depth, 3, 3)
        bn(sid, u + "res_layer.4", depth)
        sid[u + "res_layer.5.fc1.weight"] = (depth // 16, depth, 1, 1)
        sid[u + "res_layer.5.fc2.weight"] = (depth, depth // 16, 1, 1)
    bn(sid, p + "output_layer.0", 512)
    sid[p + "output_layer.3.weight"], sid[p + "output_layer.3.bias"] = (512, 512 * 7 * 7), (512,)
    bn(sid, p + "output_layer.4", 512)

    spar = {}
    f = [16, 32, 64, 128, 256]

    def uc(key, ci, co):
        for c, a in (("conv1", ci), ("conv2", co)):
            spar[f"{key}.{c}.0.weight"], spar[f"{key}.{c}.0.bias"] = (co, a, 3, 3), (co,)
            bn(spar, f"{key}.{c}.1", co)

    uc("G.conv1", 3, f[0]), uc("G.conv2", f[0], f[1]), uc("G.conv3", f[1], f[2]), uc("G.conv4", f[2], f[3]), uc("G.center", f[3], f[4])
    for name, ci, co in (("G.up_concat4", f[4], f[3]), ("G.up_concat3", f[3], f[2]), ("G.up_concat2", f[2], f[1]), ("G.up_concat1", f[1], f[0])):
        uc(name + ".conv", ci, co)
        spar[name + ".up.weight"], spar[name + ".up.bias"] = (ci, co, 2, 2), (co,)
    spar["G.final.weight"], spar["G.final.bias"] = (19, f[0], 1, 1), (19,)
    return {"lpips": synthetic_loss_state(shapes_lpips, salt), "id": synthetic_loss_state(sid, salt + 1, set(prelu)),
            "parsing": synthetic_loss_state(spar, salt + 2)}
