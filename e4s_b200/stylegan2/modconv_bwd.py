"""Backward of the fused StyledConv / ToRGB ops (input-, style- and noise-gradients; weights are frozen)."""
from __future__ import annotations


def save_for_styled_backward(ctx, *a):
    raise NotImplementedError("e4s_b200: gradients through StyledConv are not available in this build")


def styled_backward(ctx, gy):
    raise NotImplementedError


def save_for_torgb_backward(ctx, *a):
    raise NotImplementedError("e4s_b200: gradients through ToRGB are not available in this build")


def torgb_backward(ctx, g):
    raise NotImplementedError
