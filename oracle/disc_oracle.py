"""CPU restatement (functional torch) of the StyleGAN2 discriminator forward - TEST INFRASTRUCTURE.

Follows /root/reference/src/models/stylegan2/model.py: ConvLayer :670-716 (Blur pad ((p+1)//2, p//2), p = 2 + (k - 1), before
a stride-2 EqualConv2d; FusedLeakyReLU when bias, else ScaledLeakyReLU), ResBlock :719-737, Discriminator.forward :778-799
(minibatch standard deviation over groups of 4).  State dict keys are the reference module's.  PINNED by
oracle/make_golden_disc.py against the imported reference (tests/golden/disc_vectors.npz)."""
import math

import torch
import torch.nn.functional as F

from . import e4s_oracle as O

CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


def conv_layer(sd, key, x, kernel_size, downsample=False, bias=True, activate=True):
    """One ConvLayer; `key` is the nn.Sequential's prefix (its children are numbered as in the reference)."""
    idx = 0
    stride, padding = 1, kernel_size // 2
    if downsample:
        p = 2 + (kernel_size - 1)
        x = O.upfirdn2d(x, O.make_fir((1, 3, 3, 1)), pad=((p + 1) // 2, p // 2))
        idx, stride, padding = 1, 2, 0
    w = sd[f"{key}.{idx}.weight"]
    scale = 1.0 / math.sqrt(w.shape[1] * kernel_size ** 2)
    cb = sd.get(f"{key}.{idx}.bias") if (bias and not activate) else None
    x = F.conv2d(x, w * scale, cb, stride=stride, padding=padding)
    if activate:
        if bias:
            x = O.fused_leaky_relu(x, sd[f"{key}.{idx + 1}.bias"])
        else:
            x = F.leaky_relu(x, 0.2) * math.sqrt(2)
    return x


def discriminator_forward(sd, img, size):
    x = conv_layer(sd, "convs.0", img, 1)
    log_size = int(math.log2(size))
    for n, _ in enumerate(range(log_size, 2, -1), 1):
        k = f"convs.{n}"
        out = conv_layer(sd, k + ".conv1", x, 3)
        out = conv_layer(sd, k + ".conv2", out, 3, downsample=True)
        skip = conv_layer(sd, k + ".skip", x, 1, downsample=True, activate=False, bias=False)
        x = (out + skip) / math.sqrt(2)
    b, c, h, w = x.shape
    group = min(b, 4)
    sdv = x.view(group, -1, 1, c, h, w)
    sdv = torch.sqrt(sdv.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdims=True).squeeze(2)
    x = torch.cat([x, sdv.repeat(group, 1, h, w)], 1)
    x = conv_layer(sd, "final_conv", x, 3)
    x = x.view(b, -1)
    x = O.equal_linear(x, sd["final_linear.0.weight"], sd["final_linear.0.bias"], activation=True)
    return O.equal_linear(x, sd["final_linear.1.weight"], sd["final_linear.1.bias"])


def synthetic_inputs(batch, size, seed=3):
    return torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(seed)).clamp(-2, 2)
