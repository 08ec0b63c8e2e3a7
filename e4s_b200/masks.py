"""Region-mask utilities on the GPU, bit-exact (SURVEY.md section 8a last row, section 8f.3).

`labelMap2OneHot` mirrors src/utils/torch_utils.py:166-172.  The 19 -> 12 class conversion of
CelebAMask-HQ labels mirrors __celebAHQ_masks_to_faceParser_mask_detailed, src/datasets/dataset.py:153-209,
which is a pure per-pixel table lookup; the table below restates it.
"""
import torch

from . import kernels as K

# index = CelebAMask-HQ label (0 background, 1 skin, 2 nose, 3 eye_g, 4 l_eye, 5 r_eye, 6 l_brow, 7 r_brow,
# 8 l_ear, 9 r_ear, 10 mouth, 11 u_lip, 12 l_lip, 13 hair, 14 hat, 15 ear_r, 16 neck_l, 17 neck, 18 cloth)
# value = 12-class label (0 background, 1 lip, 2 eyebrows, 3 eyes, 4 hair, 5 nose, 6 skin, 7 ears,
# 8 belowface, 9 mouth, 10 eye_glass, 11 ear_rings)
CELEBA19_TO_12 = [0, 6, 5, 10, 3, 3, 2, 2, 7, 7, 9, 1, 1, 4, 0, 11, 0, 8, 0] + [0] * (256 - 19)


def labelMap2OneHot(label, num_cls):
    """[B, 1, H, W] integer label map -> [B, num_cls, H, W] float one-hot."""
    return K.label_to_onehot(label, num_cls)


def celeba19_to_12(label_u8: torch.Tensor) -> torch.Tensor:
    lut = torch.tensor(CELEBA19_TO_12, dtype=torch.uint8, device=label_u8.device)
    return K.label_remap(label_u8.to(torch.uint8), lut)
