"""Surface-compatible stand-in for src/models/stylegan2/op/conv2d_gradfix.py.

The reference's custom double-backward conv is only active on torch 1.7/1.8 (could_use_op, :78-92); on
every later torch - including the 1.12 its environment pins - it forwards to torch.nn.functional
(:34-42, :66-75).  This module keeps that public surface (``conv2d``, ``conv_transpose2d``,
``no_weight_gradients``, ``enabled``) for callers such as src/criteria/adv_loss.py:4.  The modulated
convolutions of the synthesis network do not come through here: they run on the e4s_b200 kernels.
"""
import contextlib

from torch.nn import functional as F

enabled = True
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    previous = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = previous


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(input, weight, bias, stride, padding, dilation, groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return F.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)
