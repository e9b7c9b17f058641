"""Mirror of geometry/sampling.py: bilinear sampling with wrap-around in x and y."""
from .. import ops


def resample(image, pixels):
    """sampling.py:135-197.  image [N,H,W,C], pixels [N,h,w,2] (x = [...,0], y = [...,1]) ->
    [N,h,w,C]; weights from the un-wrapped corners, indices floor-mod wrapped."""
    return ops.resample(image, pixels)


def bilinear_wrapper2(imgs, coords):
    """sampling.py:59-67."""
    return resample(imgs, coords)
