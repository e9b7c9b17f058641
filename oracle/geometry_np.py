"""Oracle (test infrastructure): NumPy restatement of the reference geometry.

Follows /root/reference geometry/spherical.py, geometry/sampling.py and
geometry/projector.py op-for-op.  Every array op is evaluated in ``dt``
(float32 by default, float64 for the error-budget twin) with one rounding per
op and no FMA contraction -- which is what a TF-1.14 graph of separate
elementwise kernels does [TF-1.14].  Python-float constants such as
``np.pi / width`` are evaluated in float64 and rounded once to ``dt`` where
they meet an array, as ``tf.convert_to_tensor`` does.

PINNED bit for bit to the reference's own code run over a restated TF op layer (tests/golden/reference_run.npz) -- see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# spherical.py
# --------------------------------------------------------------------------- #
def linspace_tf(start, stop, num, dt=F32):
    """[TF-1.14 LinSpace CPU kernel]: ``start + step * i`` evaluated in T with
    ``step = (stop - start) / (num - 1)``; start/stop are first rounded to T.
    (Later TF versions pin the last element to ``stop``; 1.14 does not.)"""
    start = dt(start)
    stop = dt(stop)
    if num == 1:
        return np.array([start], dtype=dt)
    step = dt((stop - start) / dt(num - 1))
    i = np.arange(num).astype(dt)
    return (start + step * i).astype(dt)


def lat_long_axes(shape, dt=F32):
    """1-D axes of spherical.lat_long_grid (spherical.py:42-44): pixel-centre
    longitudes s[W] and latitudes t[H]."""
    H, W = shape
    s = linspace_tf(-np.pi + np.pi / W, np.pi - np.pi / W, W, dt)
    t = linspace_tf(-np.pi / 2.0 + np.pi / (2 * H), np.pi / 2.0 - np.pi / (2 * H), H, dt)
    return s, t


def lat_long_grid(shape, dt=F32):
    """spherical.py:42-44 -- tf.meshgrid (xy indexing) of the two linspaces.
    Returns S, T of shape [H, W]."""
    s, t = lat_long_axes(shape, dt)
    S, T = np.meshgrid(s, t)  # xy indexing: S[h, w] = s[w], T[h, w] = t[h]
    return S.astype(dt), T.astype(dt)


def theta_phi_to_pixels(theta, phi, width, height, dt=F32):
    """spherical.py:54-68."""
    u = theta + dt(np.pi)
    u = u - dt(np.pi / width)
    u = u / dt(2 * np.pi - (2 * np.pi / width))
    u = u * dt(width - 1)
    v = (phi + dt(0.5 * np.pi) - dt(0.5 * np.pi / height)) / dt(np.pi - np.pi / height)
    v = v * dt(height - 1)
    return np.stack([u, v], axis=-1)


def backproject_spherical(S, T, depth, intrinsics=None, dt=F32):
    """spherical.py:116-129.  S, T: [H, W]; depth: [P] -> x, y, z: [P, H, W].
    Note the bracketing ``depth * (cos(S) * cosT)``."""
    depth = np.asarray(depth, dtype=dt).reshape(-1, 1, 1)
    S = S[None].astype(dt)
    T = T[None].astype(dt)
    cosT = np.cos(T)
    x = depth * (np.cos(S) * cosT)
    y = depth * np.sin(T)
    z = depth * (np.sin(S) * cosT)
    return x, y, z


def project_ods(points, order, pose, intrinsics, width, height, dt=F32, return_aux=False):
    """spherical.py:170-233, tuple branch (:176-177, no y-negation).

    points: (x, y, z) each [P, H, W]; order: +1 (ref/left) or -1 (src/right);
    intrinsics[0][0][0] is the ODS baseline radius (:181).
    Returns uv [P, H, W, 2]; pixels with disc < 0 are set to (1, 1) (:226-229).
    """
    x, y, z = points
    one = dt(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = dt(np.asarray(intrinsics)[0][0][0])
        f = r * r - (np.square(x) + np.square(z))
        z_larger_x = np.greater(np.abs(z), np.abs(x))
        px = np.where(z_larger_x, x, z)
        pz = np.where(z_larger_x, z, x)

        # Solve quadratic (:187-192)
        pz_square = np.square(pz)
        a = one + np.square(px) / pz_square
        b = dt(-2) * f * px / pz_square
        c = f + np.square(f) / pz_square
        disc = np.square(b) - dt(4) * a * c

        # Direction vector from point (:195-202)
        s = dt(-order) * np.sign(pz) * np.sqrt(disc)
        s = np.where(z_larger_x, s, -s)

        dx = (-b + s) / (dt(2) * a)
        dz = (f - px * dx) / pz

        dx_final = np.where(z_larger_x, -dx, -dz)
        dz_final = np.where(z_larger_x, -dz, -dx)
        dx = dx_final
        dz = dz_final
        dy = y

        # Angles from direction vector (:208-219)
        theta = -np.arctan2(dz, dx)
        phi = np.arctan2(dy, np.sqrt(np.square(dx) + np.square(dz)))
        nan_mask = np.isnan(phi)
        phi = np.where(nan_mask, np.ones_like(phi), phi)

        pos_phi = np.ones_like(dx) * dt(np.pi / 2)
        neg_phi = np.ones_like(dx) * dt(np.pi / 2) * dt(-1.0)
        pos_phi_mask = np.less_equal(phi, dt(np.pi / 2))
        neg_phi_mask = np.greater_equal(phi, dt(-np.pi / 2))
        phi = np.where(pos_phi_mask, phi, pos_phi)
        phi = np.where(neg_phi_mask, phi, neg_phi)

        # Pixel coords (:222-223)
        u = ((theta + dt(np.pi) - dt(np.pi / width)) / dt(2 * np.pi - 2 * np.pi / width)) * dt(width - 1)
        v = ((phi + dt(0.5 * np.pi) - dt(0.5 * np.pi / height)) / dt(np.pi - np.pi / height)) * dt(height - 1)

        # Keep valid parts (:226-229)
        valid_mask = np.greater_equal(disc, dt(0))
        ones = np.ones_like(u)
        u = np.where(valid_mask, u, ones)
        v = np.where(valid_mask, v, ones)
    uv = np.stack([u, v], axis=-1).astype(dt)
    if return_aux:
        return uv, {"disc": disc, "valid": valid_mask}
    return uv


def project_spherical(points, order, pose, intrinsics, width, height, dt=F32):
    """spherical.py:235-246."""
    x, y, z = points
    theta = -np.arctan2(z, x)
    phi = np.arctan2(y, np.sqrt(np.square(x) + np.square(z)))
    return theta_phi_to_pixels(theta, phi, width, height, dt)


def _matvec_rows(mat, vecs, dt):
    """Row-by-row ``mat @ stack(vecs)`` with the k-sum evaluated left to right,
    one rounding per op: ((m0*v0 + m1*v1) + m2*v2) [+ m3*v3].
    [TF-1.14 MatMul/BatchMatMul on CPU: Eigen, pip wheel built without FMA;
    the accumulation order over k = 0..K-1 is taken as sequential.]"""
    out = []
    for i in range(mat.shape[0]):
        acc = dt(mat[i, 0]) * vecs[0]
        for k in range(1, len(vecs)):
            acc = acc + dt(mat[i, k]) * vecs[k]
        out.append(acc)
    return out


def intersect_sphere(pos, center, radius, num_planes, num_batch, width, height, dt=F32):
    """spherical.py:268-326 (+ project_spherical :235-246).

    pos: [4, 4] target pose [R|t]; center: [3] (or [3, 1]) target offset, read as
    (cx, cy, cz) = (center[2], center[1], center[0]) (:286-288); radius: [L].
    Returns uv [L, H, W, 2]."""
    pos = np.asarray(pos, dtype=dt)
    center = np.asarray(center, dtype=dt).reshape(-1)
    radius = np.asarray(radius, dtype=dt).reshape(num_planes, 1, 1)
    S, T = lat_long_grid((height, width), dt)
    S = np.broadcast_to(S[None], (num_planes, height, width))
    T = np.broadcast_to(T[None], (num_planes, height, width))

    cosT = np.cos(T)
    rx = np.cos(S) * cosT
    ry = np.sin(T)
    rz = np.sin(S) * cosT

    cx, cy, cz = center[2], center[1], center[0]

    rx, ry, rz = _matvec_rows(pos[:3, :3], [rx, ry, rz], dt)
    pt = _matvec_rows(pos, [cx, cy, cz, dt(1)], dt)
    cx, cy, cz = pt[0], pt[1], pt[2]

    with np.errstate(invalid="ignore", divide="ignore"):
        a = rx * rx + ry * ry + rz * rz
        b = dt(2) * (rx * cx + ry * cy + rz * cz)
        c = cx * cx + cy * cy + cz * cz - radius * radius
        disc = np.square(b) - dt(4) * a * c
        t = (-b + np.sqrt(disc)) / (dt(2) * a)
        x = cx + t * rx
        y = cy + t * ry
        z = cz + t * rz
        return project_spherical((x, y, z), 1, None, None, width, height, dt).astype(dt)


def intersect_ods(pose, center, order, intrinsics, radius, num_planes, num_batch, width, height, dt=F32):
    """spherical.py:328-365 (+ transform_ray :70-94, get_sphere_intersections :96-111,
    project_spherical :235-246).  pose [4,4]; order +1 (left) / -1 (right); intrinsics[0][0][0] =
    baseline; radius [L].  ``center`` is accepted and ignored, as in the reference.  uv [L,H,W,2]."""
    pose = np.asarray(pose, dtype=dt)
    radius = np.asarray(radius, dtype=dt).reshape(num_planes, 1, 1)
    baseline = dt(np.asarray(intrinsics)[0][0][0])
    S, T = lat_long_grid((height, width), dt)
    S = np.broadcast_to(S[None], (num_planes, height, width))
    T = np.broadcast_to(T[None], (num_planes, height, width))
    cosT = np.cos(T)
    rx = np.cos(S) * cosT
    ry = np.sin(T)
    rz = -np.sin(S) * cosT
    cx = -np.sin(S) * baseline * dt(order)
    cy = np.zeros_like(S)
    cz = -np.cos(S) * baseline * dt(order)
    rx, ry, rz = _matvec_rows(pose[:3, :3], [rx, ry, rz], dt)
    pt = _matvec_rows(pose, [cx, cy, cz, np.ones_like(cx)], dt)
    cx, cy, cz = pt[0], pt[1], pt[2]
    with np.errstate(invalid="ignore", divide="ignore"):
        a = rx * rx + ry * ry + rz * rz
        b = dt(2) * (rx * cx + ry * cy + rz * cz)
        c = cx * cx + cy * cy + cz * cz - radius * radius
        disc = np.square(b) - dt(4) * a * c
        t = (-b + np.sqrt(disc)) / (dt(2) * a)
        x = cx + t * rx
        y = cy + t * ry
        z = cz + t * rz
        return project_spherical((x, y, z), order, None, intrinsics, width, height, dt).astype(dt)


def uv_grid(shape, dt=F32):
    """spherical.py:46-48: pixel-centre coordinates in (-1, 1).  [TF-1.14 LinSpace]"""
    h, w = shape
    return np.meshgrid(linspace_tf(-1.0 + 1.0 / w, 1.0 - 1.0 / w, w, dt), linspace_tf(-1.0 + 1.0 / h, 1.0 - 1.0 / h, h, dt))


def viewing_window_pose(viewing_window, dt=F32):
    """projector.py:80-85: [R | 0] with R = tfg rotation_matrix_3d.from_euler([0, vw*pi/2, 0]).
    [tensorflow_graphics 1.0.0, absent here: R = Rz Ry Rx built from float32 sines / cosines of the
    float32 angles; for (0, a, 0) that is [[cos a, 0, sin a], [0, 1, 0], [-sin a, 0, cos a]].]"""
    a = dt(viewing_window * np.pi / 2.0)
    sy, cy = np.sin(a, dtype=dt), np.cos(a, dtype=dt)
    return np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]], dtype=dt)


def intersect_perspective(pos, center, radius, num_planes, num_batch, width, height, tgt_width, tgt_height,
                          intrinsics=None, dt=F32):
    """spherical.py:367-401 (+ transform_ray :70-94, get_sphere_intersections :96-111, project_spherical
    :235-246).  pos [4,4]; center [3] (or [3,1]); radius [L].  uv [L, tgt_height, tgt_width, 2] into the
    width x height ERP layers."""
    pos = np.asarray(pos, dtype=dt)
    center = np.asarray(center, dtype=dt).reshape(3)
    radius = np.asarray(radius, dtype=dt).reshape(num_planes, 1, 1)
    S, T = uv_grid((tgt_height, tgt_width), dt)
    S = np.broadcast_to(S[None], (num_planes, tgt_height, tgt_width))
    T = np.broadcast_to(T[None], (num_planes, tgt_height, tgt_width))
    rx = S * dt(0.1)
    ry = T * dt(0.05)
    rz = -np.ones_like(S) * dt(0.05)
    cx = np.full_like(S, center[0])
    cy = np.full_like(S, center[1])
    cz = np.full_like(S, -center[2])
    rx, ry, rz = _matvec_rows(pos[:3, :3], [rx, ry, rz], dt)
    pt = _matvec_rows(pos, [cx, cy, cz, np.ones_like(cx)], dt)
    cx, cy, cz = pt[0], pt[1], pt[2]
    with np.errstate(invalid="ignore", divide="ignore"):
        a = rx * rx + ry * ry + rz * rz
        b = dt(2) * (rx * cx + ry * cy + rz * cz)
        c = cx * cx + cy * cy + cz * cz - radius * radius
        disc = np.square(b) - dt(4) * a * c
        t = (-b + np.sqrt(disc)) / (dt(2) * a)
        x = cx + t * rx
        y = cy + t * ry
        z = cz + t * rz
        return project_spherical((x, y, z), 1, None, None, width, height, dt).astype(dt)


# --------------------------------------------------------------------------- #
# sampling.py
# --------------------------------------------------------------------------- #
def resample(image, pixels, dt=F32, return_aux=False):
    """sampling.py:135-197 -- bilinear gather with wrap-around in x AND y.

    image: [N, H, W, C]; pixels: [N, h, w, 2] with x = pixels[..., 0],
    y = pixels[..., 1].  Weights come from the UN-wrapped corners (:157-160),
    indices are floor-mod wrapped (:162-165) [TF-1.14 tf.mod = floor-mod]."""
    image = np.asarray(image, dtype=dt)
    N, ph, pw, _ = pixels.shape
    _, height, width, C = image.shape
    x = pixels[..., 0].reshape(-1).astype(dt)
    y = pixels[..., 1].reshape(-1).astype(dt)

    x0 = np.floor(x).astype(np.int32)
    x1 = x0 + 1
    y0 = np.floor(y).astype(np.int32)
    y1 = y0 + 1

    diff_x0 = x - x0.astype(dt)
    diff_y0 = y - y0.astype(dt)
    diff_x1 = x1.astype(dt) - x
    diff_y1 = y1.astype(dt) - y

    x0 = np.mod(x0 + width, width)
    y0 = np.mod(y0 + height, height)
    x1 = np.mod(x1 + width, width)
    y1 = np.mod(y1 + height, height)

    b = np.repeat(np.arange(N), ph * pw)
    pa = image[b, y0, x0]
    pb = image[b, y0, x1]
    pc = image[b, y1, x0]
    pd = image[b, y1, x1]

    area_a = (diff_y1 * diff_x1)[:, None]
    area_b = (diff_y1 * diff_x0)[:, None]
    area_c = (diff_y0 * diff_x1)[:, None]
    area_d = (diff_y0 * diff_x0)[:, None]

    # tf.add_n of four inputs: ((a + b) + c) + d  [TF-1.14 AddN]
    res = ((area_a * pa + area_b * pb) + area_c * pc) + area_d * pd
    res = res.reshape(N, ph, pw, C).astype(dt)
    if return_aux:
        aux = {
            "x0": x0.reshape(N, ph, pw), "y0": y0.reshape(N, ph, pw),
            "x1": x1.reshape(N, ph, pw), "y1": y1.reshape(N, ph, pw),
        }
        return res, aux
    return res


bilinear_wrapper2 = resample  # sampling.py:59-67


# --------------------------------------------------------------------------- #
# projector.py
# --------------------------------------------------------------------------- #
def apply_pose(points, pose, dt=F32):
    """projector.py:275-291.  points: (x, y, z) each [P, H, W]; pose: [P, 4, 4]
    (all P copies identical on our path) -> homogeneous pose . [x y z 1]."""
    x, y, z = points
    pose = np.asarray(pose, dtype=dt)
    outs = [[], [], []]
    for p in range(x.shape[0]):
        rows = _matvec_rows(pose[p], [x[p], y[p], z[p], dt(1)], dt)
        for k in range(3):
            outs[k].append(np.broadcast_to(rows[k], x[p].shape))
    return tuple(np.stack(o).astype(dt) for o in outs)


def sweep_one(image, order, depths, pose, intrinsics, dt=F32, return_aux=False):
    """projector.py:129-170 with (st_fun, backproj_fun, proj_fun) =
    (lat_long_grid, backproject_spherical, project_ods), i.e. ods_sphere_sweep
    (:209-211).  image: [B, H, W, C]; pose: [B, 4, 4]; intrinsics: [B, 3, 3].
    Returns [B, H, W, C * P] with channel = p * C + c."""
    image = np.asarray(image, dtype=dt)
    B, H, W, C = image.shape
    depths = np.asarray(depths, dtype=dt)
    P = depths.shape[0]
    S, T = lat_long_grid((H, W), dt)
    out = []
    auxs = []
    for i in range(B):
        intrinsic = np.asarray(intrinsics, dtype=dt)[i:i + 1]
        pose_tiled = np.broadcast_to(np.asarray(pose, dtype=dt)[i:i + 1], (P, 4, 4))
        points = backproject_spherical(S, T, depths, intrinsic, dt)
        points = apply_pose(points, pose_tiled, dt)
        uv, aux = project_ods(points, order, pose_tiled, intrinsic, W, H, dt, return_aux=True)
        image_tiled = np.broadcast_to(image[i:i + 1], (P, H, W, C))
        res, raux = resample(image_tiled, uv, dt, return_aux=True)
        res = np.transpose(res, (1, 2, 0, 3))  # [H, W, P, C]
        out.append(res)
        aux.update(raux)
        aux["uv"] = uv
        auxs.append(aux)
    out = np.stack(out).reshape(B, H, W, C * P).astype(dt)
    if return_aux:
        return out, auxs
    return out


ods_sphere_sweep = sweep_one  # projector.py:209-211


def projective_forward_sphere(src_images, intrinsics, tgt_pose_rt, tgt_pos, depths, dt=F32,
                              return_aux=False):
    """projector.py:34-62.  src_images: [L, B, H, W, C]; tgt_pose_rt: [B, 4, 4];
    tgt_pos: [B, 3]; depths: [L, B] -> [L, B, H, W, C]."""
    src_images = np.asarray(src_images, dtype=dt)
    L, B, H, W, C = src_images.shape
    depths = np.asarray(depths, dtype=dt)
    coords = []
    for i in range(B):
        coords.append(intersect_sphere(tgt_pose_rt[i], tgt_pos[i], depths[:, i], L, B, W, H, dt))
    coords = np.stack(coords, axis=0).transpose(1, 0, 2, 3, 4)  # [L, B, H, W, 2]
    proj = []
    auxs = []
    for l in range(L):
        r, a = resample(src_images[l], coords[l], dt, return_aux=True)
        proj.append(r)
        auxs.append(a)
    proj = np.stack(proj, axis=0)
    if return_aux:
        return proj, coords, auxs
    return proj


def over_composite(rgbas, dt=F32):
    """projector.py:246-265.  rgbas: list (back to front) of [B, H, W, 4]."""
    output = None
    for i in range(len(rgbas)):
        rgb = rgbas[i][..., 0:3]
        alpha = rgbas[i][..., 3:]
        if i == 0:
            output = rgb
        else:
            rgb_by_alpha = rgb * alpha
            output = rgb_by_alpha + output * (dt(1.0) - alpha)
    return output.astype(dt)


def over_composite_depth(rgbas, dt=F32):
    """projector.py:225-244.  ``i / len(rgbas)`` is true division
    (``from __future__ import division``, projector.py:23)."""
    n = len(rgbas)
    output = None
    for i in range(n):
        alpha_image = np.tile(rgbas[i][..., 3:], (1, 1, 1, 3))
        if i == 0:
            output = np.zeros_like(alpha_image)
        else:
            output = dt(i / n) * alpha_image + output * (dt(1.0) - alpha_image)
    return output.astype(dt)
