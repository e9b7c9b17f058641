"""Mirror of geometry/sampling.py: bilinear sampling with wrap-around in x and y."""
import torch

from .. import torch_ops  # noqa: F401  (registers torch.ops.msi.*)


def resample(image, pixels):
    """sampling.py:135-197.  image [N,H,W,C], pixels [N,h,w,2] (x = [...,0], y = [...,1]) ->
    [N,h,w,C]; weights from the un-wrapped corners, indices floor-mod wrapped."""
    return torch.ops.msi.resample(image.contiguous().float(), pixels.contiguous().float())


def bilinear_wrapper2(imgs, coords):
    """sampling.py:59-67."""
    return resample(imgs, coords)
