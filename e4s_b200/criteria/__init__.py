"""Loss networks of the inversion loop: mirrors of src/criteria/{lpips,id_loss,face_parsing}
plus `InversionLoss`, the reference's `Optimizer.calc_loss` (scripts/optimization.py:88-122) with the target-image features
cached (the reference recomputes them every step: id_loss.py:33, lpips.py:30, face_parsing_loss.py:55)."""
from .lpips import LPIPS
from .id_loss import IDLoss
from .face_parsing import FaceParsingLoss, unet
from .inversion_loss import InversionLoss

__all__ = ["LPIPS", "IDLoss", "FaceParsingLoss", "unet", "InversionLoss"]
