"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the mask stage of the E4S face-swapping pipeline
(SURVEY.md section 8f.3): shape swapping of two 12-class parsing maps, the foreground mask, and the flat
box dilation / erosion that build the blending masks.  All integer / index work: the bar is bit-exact.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; nothing under
e4s_b200/ does.  Pinned against the imported reference by oracle/make_golden_masks.py
(tests/golden/mask_pipeline_vectors.npz).

Reference citations are path:line under /root/reference.
"""
from __future__ import annotations

import numpy as np

PLACEHOLDER = 99          # src/utils/swap_face_mask.py:42 ("a place-holder magic number")


def swap_head_mask(source: np.ndarray, target: np.ndarray, hair_first: bool = True):
    """swap_head_mask_revisit_considerGlass, src/utils/swap_face_mask.py:33-83.

    source (the driven face D) and target (T): integer label maps with the 12 classes of
    faceParser_label_list_detailed (:27-29).  Returns (swapped label map, hole map in {0, 255}).
    Restated as a per-pixel decision list - the reference's sequence of masked assignments touches every
    pixel independently, later assignments overriding earlier ones.
    """
    assert source.shape == target.shape
    s = source.astype(np.int64).ravel()
    t = target.astype(np.int64).ravel()
    res = np.zeros_like(t)
    # background, neck, ear, ear rings of the target (:42-45); hair first (:47-48)
    res[t == 0] = PLACEHOLDER
    res[t == 8] = 8
    res[t == 7] = 7
    res[t == 11] = 11
    if hair_first:
        res[t == 4] = 4
    # inner face of the source wherever the target is not background (:51-56).  The guard reads `res`, which the
    # loop itself never sets to the placeholder, so it is the state after the target pass.
    for c in (1, 2, 3, 5, 6, 9):
        res[(s == c) & (res != PLACEHOLDER)] = c
    if not hair_first:
        res[t == 4] = 4                         # :66-67
    res[t == 10] = 10                           # eye glasses of the target (:70)
    hole = np.where(res == 0, 255, 0)           # :74-78 (both branches give 255 * (res == 0))
    res[res == 0] = 6                           # missing pixels become skin (:76)
    res[res == PLACEHOLDER] = 0                 # :81
    return res.reshape(target.shape).astype(target.dtype), hole.reshape(target.shape).astype(target.dtype)


def foreground_mask(swapped: np.ndarray, hole: np.ndarray) -> np.ndarray:
    """scripts/face_swap.py:280-284: background = classes {0, 11, 4}; holes are foreground.  Returns uint8 0/1."""
    bg = (swapped == 0) | (swapped == 11) | (swapped == 4)
    return ((~bg) | (hole == 255)).astype(np.uint8)


def _box_reduce(mask: np.ndarray, radius: int, op, neutral) -> np.ndarray:
    """Flat (2r+1)^2 box structuring element, origin at the centre, 'geodesic' border: positions outside the image are
    ignored (padded with -max_val for dilation, +max_val for erosion: src/utils/morphology.py:83-86, 170-173).  A box is
    separable, so the window reduction runs over rows, then columns, as 2r+1 shifted views of a neutrally padded copy."""
    out = mask
    for axis in (-2, -1):
        n = out.shape[axis]
        pad = [(0, 0)] * out.ndim
        pad[axis] = (radius, radius)
        padded = np.pad(out, pad, constant_values=neutral)
        acc = None
        for k in range(2 * radius + 1):
            view = np.take(padded, range(k, k + n), axis=axis)
            acc = view if acc is None else op(acc, view)
        out = acc
    return out.astype(mask.dtype)


def box_dilation(mask: np.ndarray, radius: int) -> np.ndarray:
    """dilation(mask, ones(2r+1, 2r+1), engine='convolution'), src/utils/morphology.py:23-106, on a 0/1 mask:
    max over the window of (value + 0)."""
    return _box_reduce(mask, radius, np.maximum, mask.min() if mask.size else 0)


def box_erosion(mask: np.ndarray, radius: int) -> np.ndarray:
    """erosion(...), src/utils/morphology.py:109-197: min over the window of (value - 0)."""
    return _box_reduce(mask, radius, np.minimum, mask.max() if mask.size else 0)


def create_masks(mask: np.ndarray, outer_dilation: int = 0, operation: str = "dilation"):
    """scripts/face_swap.py:30-48.  mask: 0/1 array [..., H, W].  Returns (content, border, full)."""
    r = outer_dilation
    if operation == "dilation":
        full = box_dilation(mask, r)
        border = full.astype(np.int64) - mask
    elif operation == "erosion":
        full = box_erosion(mask, r)
        border = mask.astype(np.int64) - full
    elif operation == "expansion":
        full = box_dilation(mask, r)
        border = full.astype(np.int64) - box_erosion(mask, r)
    else:
        raise UnboundLocalError("local variable 'border_mask' referenced before assignment")   # what the reference does
    border = np.clip(border, 0, 1).astype(mask.dtype)
    return mask, border, full


def swap_comp_style_vector(sv1: np.ndarray, sv2: np.ndarray, comp_indices, below_face_interpolation: bool = False):
    """scripts/face_swap.py:117-146: [1, ncls, C] texture vectors of target (sv1) and source (sv2)."""
    out = sv1.copy()
    for c in comp_indices:
        out[:, c, :] = sv2[:, c, :]
    if sv2[:, 7, :].sum() == 0:                 # no ear region in the source (:132-133)
        out[:, 7, :] = (sv1[:, 7, :] + sv2[:, 7, :]) / 2
    if sv2[:, 9, :].sum() == 0:                 # no teeth region in the source (:136-137)
        out[:, 9, :] = sv1[:, 9, :]
    if below_face_interpolation:
        out[:, 8, :] = (sv1[:, 8, :] + sv2[:, 8, :]) / 2
    return out
