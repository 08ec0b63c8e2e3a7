"""Steps 3-5 (and the mask half of step 6) of the reference's face-swapping pipeline, batched and device-resident.

Mirrors `faceSwapping_pipeline`, scripts/face_swap.py:149-330, between the points where it holds the driven face D,
the target face T and their 12-class parsing maps (:217-236) and where it hands the swapped face and the blending
masks to the PIL/cv2 paste code (:291-330):

    (3) texture vectors of D and T with the RGI encoder                         :238-239  Net3.get_style_vectors
    (4) shape swap of the two parsing maps, texture swap of the vectors         :253, :262-263
    (5) swapped mask + swapped vectors -> generator                             :268-276  Net3.cal_style_codes, gen_img
    (6a) foreground of the swapped mask and its dilated / border masks          :279-289  create_masks

The reference runs this for ONE pair, moves the maps through numpy / PIL between the steps (four host round trips) and
synchronises twice inside swap_comp_style_vector.  Here a batch of pairs stays on the device from the label maps to
the image; steps 1-2 (dlib alignment, face-vid2vid re-enactment, GPEN, the face parser) and the paste-back of step 6
need third-party networks and CPU image libraries and are outside the hot path (DESIGN.md
section 8).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import masks as M

# regions whose texture comes from the driven face: everything but background, hair, ear rings, eye glasses
# (scripts/face_swap.py:262)
TARGET_KEPT_REGIONS = (0, 4, 11, 10)


@dataclass
class SwapResult:
    image: torch.Tensor            # [B, 3, S, S] swapped faces (generator range, like Net3.gen_img)
    swapped_label: torch.Tensor    # [B, H, W] uint8 recomposed parsing maps
    hole_map: torch.Tensor         # [B, H, W] uint8, 255 where no region claimed the pixel
    content_mask: torch.Tensor     # [B, 1, H, W] float 0/1: foreground of the swapped mask
    border_mask: torch.Tensor      # [B, 1, H, W]
    full_mask: torch.Tensor        # [B, 1, H, W] dilated (or expanded) foreground
    style_vectors: torch.Tensor    # [B, ncls, 1280] swapped texture vectors


@torch.no_grad()
def swap_faces(net, driven: torch.Tensor, target: torch.Tensor, driven_label: torch.Tensor, target_label: torch.Tensor,
               outer_dilation: int = 5, lap_bld: bool = False, hair_first: bool = True,
               belowFace_interpolation: bool = False, noise=None) -> SwapResult:
    """driven / target: [B, 3, H, W] normalised faces (CUDA); *_label: [B, H, W] or [B, 1, H, W] integer parsing maps with
    the 12 classes of faceParser_label_list_detailed.  `net` is an e4s_b200.networks.Net3; `noise`: optional list of per-layer
    noise maps (Generator.forward's `noise=`), fresh Gaussian noise like the reference otherwise."""
    ncls = net.opts.num_seg_cls
    if driven_label.ndim == 4:
        driven_label = driven_label[:, 0]
    if target_label.ndim == 4:
        target_label = target_label[:, 0]
    d_lab = driven_label.to(torch.uint8).contiguous()
    t_lab = target_label.to(torch.uint8).contiguous()
    # (3) texture vectors of D and T: one encoder pass over both halves of the batch
    b = driven.shape[0]
    onehot = M.labelMap2OneHot(torch.cat([d_lab, t_lab]), ncls)
    vectors, _ = net.get_style_vectors(torch.cat([driven, target]), onehot)
    d_vec, t_vec = vectors[:b], vectors[b:]
    # (4) shape swap and texture swap
    swapped, hole, fg = M.swap_head_mask_with_foreground(d_lab, t_lab, hair_first=hair_first)
    comp = sorted(set(range(ncls)) - set(TARGET_KEPT_REGIONS))
    vec = M.swap_comp_style_vector(t_vec, d_vec, comp, belowFace_interpolation=belowFace_interpolation)
    # (5) generator
    codes = net.cal_style_codes(vec)
    image, _, _ = net.gen_img(None, codes, M.labelMap2OneHot(swapped, ncls), noise=noise)
    # (6a) blending masks
    content, border, full = M.create_masks(fg[:, None].float(), outer_dilation=outer_dilation,
                                           operation="expansion" if lap_bld else "dilation")
    return SwapResult(image, swapped, hole, content, border, full, vec)
