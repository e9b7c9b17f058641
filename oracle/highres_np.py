"""Oracle (test infrastructure): the high-res plane-streamed re-render of the reference driver.

Follows /root/reference test.py:284-394: for every PSV plane i, build ONE high-res PSV plane for both
eyes (:312-317), bilinearly upsample that plane's low-res blend weight and alpha with
``align_corners=True`` (:319-325), blend (:327-334), reproject the single layer to the target position
(``msi_render_equirect_view_single``, :338), and over-composite on the host in NumPy (:374-382).

[TF-1.14 ResizeBilinear, align_corners=True, legacy scaler]:  scale = (in - 1) / float(out - 1);
in = out_index * scale;  lower = floor(in);  upper = min(ceil(in), in_size - 1);  lerp = in - lower;
top = tl + (tr - tl) * x_lerp;  bottom = bl + (br - bl) * x_lerp;  out = top + (bottom - top) * y_lerp,
all in float32.

PINNED bit for bit (tests/test_reference_golden.py::test_high_res_rerender_bit_exact) to the reference run of
oracle/refrun/run_reference.py::run_highres: the graph part of test.py:300-340 goes through the reference's own
MSI.format_network_input (one tensor plane) and msi_render_equirect_view_single; what is restated there, as here, is
the per-plane feed loop and host-side composite of test.py:354-383 (script code inside main()) and the align-corners
bilinear resize [TF-1.14 tf.image.resize].
"""
from __future__ import annotations

import numpy as np

from . import geometry_np as g
from . import msi_np

F32 = np.float32


def resize_bilinear_align_corners(img, out_h, out_w, dt=F32):
    """img [B, h, w, C] -> [B, out_h, out_w, C]."""
    img = np.asarray(img, dtype=dt)
    B, h, w, C = img.shape

    def weights(in_size, out_size):
        scale = dt((in_size - 1) / dt(out_size - 1)) if out_size > 1 else dt(0)
        pos = (np.arange(out_size).astype(dt) * scale).astype(dt)
        lo = np.floor(pos)
        hi = np.minimum(np.ceil(pos), in_size - 1)
        return np.maximum(lo, 0).astype(np.int64), hi.astype(np.int64), (pos - lo).astype(dt)

    y0, y1, ly = weights(h, out_h)
    x0, x1, lx = weights(w, out_w)
    lx = lx[None, None, :, None]
    ly = ly[None, :, None, None]
    tl = img[:, y0][:, :, x0]
    tr = img[:, y0][:, :, x1]
    bl = img[:, y1][:, :, x0]
    br = img[:, y1][:, :, x1]
    top = tl + (tr - tl) * lx
    bottom = bl + (br - bl) * lx
    return (top + (bottom - top) * ly).astype(dt)


def high_res_rerender(hres_ref_image, hres_src_image, blend_weights, alphas, ref_pose, src_pose, intrinsics,
                      tgt_pos, psv_planes, dt=F32):
    """test.py:296-382 for one frame.  hres images [1, Hh, Wh, 3] in [0, 1]; blend_weights / alphas
    [1, h, w, P] (the low-res net outputs saved by the low-res pass); tgt_pos [1, 3].
    Returns (hres_output [Hh, Wh, 3] in [-1, 1], hres_depth [Hh, Wh, 3] in [0, 1))."""
    P = len(psv_planes)
    ref = msi_np.preprocess_image(hres_ref_image, dt)
    src = msi_np.preprocess_image(hres_src_image, dt)
    _, Hh, Wh, _ = ref.shape
    eye = np.eye(4, dtype=dt)[None]
    hres_output = None
    hres_depth = None
    for i in range(P):
        plane = [psv_planes[i]]
        net_input = msi_np.format_network_input(ref, src, ref_pose, src_pose, plane, intrinsics, dt=dt)  # [1,Hh,Wh,6]
        uw = resize_bilinear_align_corners(blend_weights[..., i:i + 1], Hh, Wh, dt)
        ua = resize_bilinear_align_corners(alphas[..., i:i + 1], Hh, Wh, dt)
        ufg = net_input[..., 0:3]
        ubg = net_input[..., 3:6]
        rgb = uw * ufg + (dt(1) - uw) * ubg
        rgba = np.concatenate([rgb, ua], axis=3).reshape(1, Hh, Wh, 1, 4)
        proj = msi_np.msi_render_equirect_view_single(rgba, eye, tgt_pos, plane, dt=dt)[0]  # [1, Hh, Wh, 4]
        cur = proj.astype(dt)
        rgb_p = cur[..., :3]
        alpha = cur[..., 3:]
        alpha3 = np.tile(alpha, (1, 1, 1, 3))
        if i == 0:
            hres_output = rgb_p
            hres_depth = np.zeros_like(alpha3)
        else:
            hres_output = hres_output * (dt(1.0) - alpha) + rgb_p * alpha
            hres_depth = dt(i / P) * alpha3 + hres_depth * (dt(1.0) - alpha3)
    return hres_output[0], hres_depth[0]


def deprocess_high_res(hres_output, hres_depth):
    """test.py:384-386 + utils.write_image (:76-81): ((x+1)/2)*255 resp. d*255, clipped, cast to uint8."""
    out = np.clip(((hres_output + 1.0) / 2.0) * 255.0, 0, 255).astype("uint8")
    dep = np.clip(hres_depth * 255.0, 0, 255).astype("uint8")
    return out, dep
