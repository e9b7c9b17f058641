// Device-side geometry of the MSI path, strict IEEE float32.
//
// This translation unit is compiled with -fmad=false and the default
// -prec-div=true -prec-sqrt=true -ftz=false so that every +,-,*,/ and sqrt is one
// correctly rounded float32 operation, evaluated in the same order as the
// reference's graph of separate TF ops.  With bit-identical cos/sin tables the
// quadratic discriminant of project_ods -- and therefore the `disc < 0` mask
// that snaps 1 % of the PSV samples to pixel (1,1) -- is then bit-identical to
// a float32 CPU evaluation.  Only atan2f differs from a CPU libm by a few ulp.
#pragma once
#include "common.cuh"

namespace msi {

// Python-float constants of theta_phi_to_pixels / project_ods, evaluated in double on the
// host exactly as the reference's Python expressions and rounded once to float32.
struct ErpConsts {
    float pi;         // np.pi
    float pi_w;       // np.pi / width
    float den_u;      // 2*np.pi - 2*np.pi/width
    float wm1;        // width - 1
    float half_pi;    // 0.5*np.pi  (== np.pi/2)
    float half_pi_h;  // 0.5*np.pi/height
    float den_v;      // np.pi - np.pi/height
    float hm1;        // height - 1
    float ku;         // (width - 1) / den_u   (fast render chain: one multiply instead of divide + multiply)
    float kv;         // (height - 1) / den_v
};

inline ErpConsts make_erp_consts(int H, int W) {
    const double pi = 3.141592653589793;
    ErpConsts c;
    c.pi = (float)pi;
    c.pi_w = (float)(pi / (double)W);
    c.den_u = (float)(2.0 * pi - 2.0 * pi / (double)W);
    c.wm1 = (float)(W - 1);
    c.half_pi = (float)(0.5 * pi);
    c.half_pi_h = (float)(0.5 * pi / (double)H);
    c.den_v = (float)(pi - pi / (double)H);
    c.hm1 = (float)(H - 1);
    c.ku = (float)((double)(W - 1) / (2.0 * pi - 2.0 * pi / (double)W));
    c.kv = (float)((double)(H - 1) / (pi - pi / (double)H));
    return c;
}

// spherical.project_ods (spherical.py:170-233, tuple branch) for one 3-D point.
// order = +1 / -1, r = ODS baseline.  Returns false (and u = v = 1) when disc < 0.
// The part of project_ods that does not depend on `order` (the eye): the quadratic of :187-192 and
// sqrt(disc).  Both ODS eyes of one 3-D point share it, so the sweep evaluates it once per
// (pixel, plane) when the two eyes have the same pose.
struct OdsQuad {
    float f, px, pz, a, b, sq, sgn, y;
    bool zl, valid;
};

__device__ __forceinline__ OdsQuad ods_quadratic(float x, float y, float z, float r) {
    OdsQuad q;
    q.f = r * r - (x * x + z * z);
    q.zl = fabsf(z) > fabsf(x);
    q.px = q.zl ? x : z;
    q.pz = q.zl ? z : x;
    const float pz2 = q.pz * q.pz;
    q.a = 1.0f + (q.px * q.px) / pz2;
    q.b = ((-2.0f * q.f) * q.px) / pz2;
    const float c = q.f + (q.f * q.f) / pz2;
    const float disc = q.b * q.b - (4.0f * q.a) * c;
    q.sgn = (q.pz > 0.0f) ? 1.0f : ((q.pz < 0.0f) ? -1.0f : 0.0f);
    q.sq = sqrtf(disc);
    q.valid = disc >= 0.0f;
    q.y = y;
    return q;
}

// The eye-dependent rest of project_ods (:195-229).  Returns false (and u = v = 1) when disc < 0.
__device__ __forceinline__ bool ods_finish(const OdsQuad& q, float order, const ErpConsts& k, float& u, float& v) {
    float s = ((-order) * q.sgn) * q.sq;
    s = q.zl ? s : -s;
    const float dx0 = (-q.b + s) / (2.0f * q.a);
    const float dz0 = (q.f - q.px * dx0) / q.pz;
    const float dx = q.zl ? -dx0 : -dz0;
    const float dz = q.zl ? -dz0 : -dx0;
    const float theta = -atan2f(dz, dx);
    float phi = atan2f(q.y, sqrtf(dx * dx + dz * dz));
    if (phi != phi) phi = 1.0f;
    phi = (phi <= k.half_pi) ? phi : k.half_pi;
    phi = (phi >= -k.half_pi) ? phi : -k.half_pi;
    u = ((theta + k.pi - k.pi_w) / k.den_u) * k.wm1;
    v = ((phi + k.half_pi - k.half_pi_h) / k.den_v) * k.hm1;
    if (!q.valid) {
        u = 1.0f;
        v = 1.0f;
    }
    return q.valid;
}

__device__ __forceinline__ bool project_ods_point(float x, float y, float z, float order, float r,
                                                  const ErpConsts& k, float& u, float& v) {
    const OdsQuad q = ods_quadratic(x, y, z, r);
    return ods_finish(q, order, k, u, v);
}

// backproject_spherical (spherical.py:116-129) + apply_pose (projector.py:275-291) of one grid point
__device__ __forceinline__ void sweep_point(float cs, float sn, float ct, float st, float depth,
                                            const float* __restrict__ pose, float& x, float& y, float& z) {
    const float x0 = depth * (cs * ct);
    const float y0 = depth * st;
    const float z0 = depth * (sn * ct);
    x = ((pose[0] * x0 + pose[1] * y0) + pose[2] * z0) + pose[3];
    y = ((pose[4] * x0 + pose[5] * y0) + pose[6] * z0) + pose[7];
    z = ((pose[8] * x0 + pose[9] * y0) + pose[10] * z0) + pose[11];
}

// backproject_spherical (spherical.py:116-129) + apply_pose (projector.py:275-291) +
// project_ods for one (pixel, plane, eye).  pose = 16 floats row-major.
__device__ __forceinline__ bool sweep_uv(float cs, float sn, float ct, float st, float depth,
                                         const float* __restrict__ pose, float order, float r,
                                         const ErpConsts& k, float& u, float& v) {
    const float x0 = depth * (cs * ct);
    const float y0 = depth * st;
    const float z0 = depth * (sn * ct);
    const float x = ((pose[0] * x0 + pose[1] * y0) + pose[2] * z0) + pose[3];
    const float y = ((pose[4] * x0 + pose[5] * y0) + pose[6] * z0) + pose[7];
    const float z = ((pose[8] * x0 + pose[9] * y0) + pose[10] * z0) + pose[11];
    return project_ods_point(x, y, z, order, r, k, u, v);
}

// intersect_sphere split in two: the layer-independent ray of one target pixel (:279-311, and the
// a, b coefficients of :313-314) ...
struct SphereRay {
    float rx, ry, rz, cx, cy, cz, a, b;
};

__device__ __forceinline__ SphereRay sphere_ray(float cs, float sn, float ct, float st, const float* __restrict__ pos,
                                                const float* __restrict__ center) {
    SphereRay q;
    const float rx0 = cs * ct;
    const float ry0 = st;
    const float rz0 = sn * ct;
    q.rx = (pos[0] * rx0 + pos[1] * ry0) + pos[2] * rz0;
    q.ry = (pos[4] * rx0 + pos[5] * ry0) + pos[6] * rz0;
    q.rz = (pos[8] * rx0 + pos[9] * ry0) + pos[10] * rz0;
    const float c0 = center[2], c1 = center[1], c2 = center[0];  // (cx,cy,cz) = (center[2],[1],[0]) :286-288
    q.cx = ((pos[0] * c0 + pos[1] * c1) + pos[2] * c2) + pos[3];
    q.cy = ((pos[4] * c0 + pos[5] * c1) + pos[6] * c2) + pos[7];
    q.cz = ((pos[8] * c0 + pos[9] * c1) + pos[10] * c2) + pos[11];
    q.a = (q.rx * q.rx + q.ry * q.ry) + q.rz * q.rz;
    q.b = 2.0f * ((q.rx * q.cx + q.ry * q.cy) + q.rz * q.cz);
    return q;
}

// The ray of spherical.intersect_ods (spherical.py:328-365) + transform_ray (:70-94): an ODS eye
// (order = +1 left / -1 right, viewing-circle radius = baseline) looks along
// (cosS cosT, sinT, -sinS cosT) from (-sinS b order, 0, -cosS b order); both are moved by `pose`.
// The target offset is NOT used by the reference here.
__device__ __forceinline__ SphereRay sphere_ray_ods(float cs, float sn, float ct, float st,
                                                    const float* __restrict__ pos, float order, float baseline) {
    SphereRay q;
    const float rx0 = cs * ct;
    const float ry0 = st;
    const float rz0 = (-sn) * ct;
    const float cx0 = ((-sn) * baseline) * order;
    const float cy0 = 0.0f;
    const float cz0 = ((-cs) * baseline) * order;
    q.rx = (pos[0] * rx0 + pos[1] * ry0) + pos[2] * rz0;
    q.ry = (pos[4] * rx0 + pos[5] * ry0) + pos[6] * rz0;
    q.rz = (pos[8] * rx0 + pos[9] * ry0) + pos[10] * rz0;
    q.cx = ((pos[0] * cx0 + pos[1] * cy0) + pos[2] * cz0) + pos[3];
    q.cy = ((pos[4] * cx0 + pos[5] * cy0) + pos[6] * cz0) + pos[7];
    q.cz = ((pos[8] * cx0 + pos[9] * cy0) + pos[10] * cz0) + pos[11];
    q.a = (q.rx * q.rx + q.ry * q.ry) + q.rz * q.rz;
    q.b = 2.0f * ((q.rx * q.cx + q.ry * q.cy) + q.rz * q.cz);
    return q;
}

// The ray of spherical.intersect_perspective (spherical.py:367-401) + transform_ray (:70-94): pixel
// (S, T) of uv_grid looks along (0.1 S, 0.05 T, -0.05) (hard-coded intrinsics, :385-387) from
// (center[0], center[1], -center[2]) (:390-392); both are moved by `pos`.
__device__ __forceinline__ SphereRay sphere_ray_perspective(float S, float T, const float* __restrict__ pos,
                                                            const float* __restrict__ center) {
    SphereRay q;
    const float rx0 = S * 0.1f;
    const float ry0 = T * 0.05f;
    const float rz0 = -1.0f * 0.05f;
    const float cx0 = center[0], cy0 = center[1], cz0 = -center[2];
    q.rx = (pos[0] * rx0 + pos[1] * ry0) + pos[2] * rz0;
    q.ry = (pos[4] * rx0 + pos[5] * ry0) + pos[6] * rz0;
    q.rz = (pos[8] * rx0 + pos[9] * ry0) + pos[10] * rz0;
    q.cx = ((pos[0] * cx0 + pos[1] * cy0) + pos[2] * cz0) + pos[3];
    q.cy = ((pos[4] * cx0 + pos[5] * cy0) + pos[6] * cz0) + pos[7];
    q.cz = ((pos[8] * cx0 + pos[9] * cy0) + pos[10] * cz0) + pos[11];
    q.a = (q.rx * q.rx + q.ry * q.ry) + q.rz * q.rz;
    q.b = 2.0f * ((q.rx * q.cx + q.ry * q.cy) + q.rz * q.cz);
    return q;
}

// ... and the per-layer hit + projection (:315-326, :235-246, :54-68).  Same operations in the same
// order as sphere_uv below, so both forms give identical bits.
__device__ __forceinline__ void sphere_hit_uv(const SphereRay& q, float radius, const ErpConsts& k, float& u, float& v) {
    const float c = ((q.cx * q.cx + q.cy * q.cy) + q.cz * q.cz) - radius * radius;
    const float disc = q.b * q.b - (4.0f * q.a) * c;
    const float t = (-q.b + sqrtf(disc)) / (2.0f * q.a);
    const float x = q.cx + t * q.rx;
    const float y = q.cy + t * q.ry;
    const float z = q.cz + t * q.rz;
    const float theta = -atan2f(z, x);
    const float phi = atan2f(y, sqrtf(x * x + z * z));
    u = theta + k.pi;
    u = u - k.pi_w;
    u = u / k.den_u;
    u = u * k.wm1;
    v = (phi + k.half_pi - k.half_pi_h) / k.den_v;
    v = v * k.hm1;
}

// spherical.intersect_sphere (spherical.py:268-326) + project_spherical (:235-246) +
// theta_phi_to_pixels (:54-68) for one (pixel, layer).
__device__ __forceinline__ void sphere_uv(float cs, float sn, float ct, float st,
                                          const float* __restrict__ pos, const float* __restrict__ center,
                                          float radius, const ErpConsts& k, float& u, float& v) {
    const float rx0 = cs * ct;
    const float ry0 = st;
    const float rz0 = sn * ct;
    const float rx = (pos[0] * rx0 + pos[1] * ry0) + pos[2] * rz0;
    const float ry = (pos[4] * rx0 + pos[5] * ry0) + pos[6] * rz0;
    const float rz = (pos[8] * rx0 + pos[9] * ry0) + pos[10] * rz0;
    // (cx, cy, cz) = (center[2], center[1], center[0])   spherical.py:286-288
    const float c0 = center[2], c1 = center[1], c2 = center[0];
    const float cx = ((pos[0] * c0 + pos[1] * c1) + pos[2] * c2) + pos[3];
    const float cy = ((pos[4] * c0 + pos[5] * c1) + pos[6] * c2) + pos[7];
    const float cz = ((pos[8] * c0 + pos[9] * c1) + pos[10] * c2) + pos[11];
    const float a = (rx * rx + ry * ry) + rz * rz;
    const float b = 2.0f * ((rx * cx + ry * cy) + rz * cz);
    const float c = ((cx * cx + cy * cy) + cz * cz) - radius * radius;
    const float disc = b * b - (4.0f * a) * c;
    const float t = (-b + sqrtf(disc)) / (2.0f * a);
    const float x = cx + t * rx;
    const float y = cy + t * ry;
    const float z = cz + t * rz;
    const float theta = -atan2f(z, x);
    const float phi = atan2f(y, sqrtf(x * x + z * z));
    u = theta + k.pi;
    u = u - k.pi_w;
    u = u / k.den_u;
    u = u * k.wm1;
    v = (phi + k.half_pi - k.half_pi_h) / k.den_v;
    v = v * k.hm1;
}

// sampling.resample corner set-up (sampling.py:152-165): floor, un-wrapped weights,
// floor-mod wrapped indices.
struct Bilinear {
    int x0, x1, y0, y1;
    float wa, wb, wc, wd;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

__device__ __forceinline__ int floor_mod(int a, int n) {
    int m = a % n;
    return m < 0 ? m + n : m;
}

// floor_mod(a + n, n) (sampling.py:162-165).  On the sweep / render paths a lies in [-n, 2n), where
// two conditional subtractions replace the integer division; anything else takes the general route.
__device__ __forceinline__ int wrap_index(int a, int n) {
    const int t = a + n;
    if ((unsigned)t < (unsigned)(3 * n)) {
        int m = t;
        m = (m >= n) ? m - n : m;
        m = (m >= n) ? m - n : m;
        return m;
    }
    return floor_mod(t, n);
}

__device__ __forceinline__ Bilinear bilinear_setup(float x, float y, int W, int H) {
    Bilinear s;
    const int x0 = (int)floorf(x);
    const int y0 = (int)floorf(y);
    const int x1 = x0 + 1;
    const int y1 = y0 + 1;
    const float dx0 = x - (float)x0;
    const float dy0 = y - (float)y0;
    const float dx1 = (float)x1 - x;
    const float dy1 = (float)y1 - y;
    s.x0 = wrap_index(x0, W);
    s.y0 = wrap_index(y0, H);
    s.x1 = wrap_index(x1, W);
    s.y1 = wrap_index(y1, H);
    s.wa = dy1 * dx1;
    s.wb = dy1 * dx0;
    s.wc = dy0 * dx1;
    s.wd = dy0 * dx0;
    return s;
}

// tf.add_n of the four weighted corners: ((a + b) + c) + d
__device__ __forceinline__ float blend4(const Bilinear& s, float pa, float pb, float pc, float pd) {
    return ((s.wa * pa + s.wb * pb) + s.wc * pc) + s.wd * pd;
}

// sampling.resample corner set-up for coordinates known to lie in [-n, 2n) (every coordinate project_ods
// returns: u in [-0.5, W - 0.5], v in [-0.5, H - 0.5], or the (1, 1) of an invalid sample): the same
// floor / weights / floor-mod as bilinear_setup without its general-range branch.  The final unsigned
// clamp only matters for a corrupt table (it keeps every tap inside the image).
struct Taps {
    unsigned a, b, c, d;   // pixel indices of (y0,x0) (y0,x1) (y1,x0) (y1,x1), frame base `ib` included
    float wa, wb, wc, wd;
};
__device__ __forceinline__ Taps taps_in_range(float x, float y, int W, int H, unsigned ib) {
    const int x0 = __float2int_rd(x), y0 = __float2int_rd(y);
    const float fx0 = (float)x0, fy0 = (float)y0;
    const float dx0 = x - fx0, dy0 = y - fy0;
    const float dx1 = (fx0 + 1.0f) - x, dy1 = (fy0 + 1.0f) - y;  // (float)(x0 + 1) == (float)x0 + 1 for |x0| < 2^24
    int xa = x0 + (x0 < 0 ? W : 0), ya = y0 + (y0 < 0 ? H : 0);
    xa -= (xa >= W ? W : 0);
    ya -= (ya >= H ? H : 0);
    int xb = xa + 1, yb = ya + 1;
    xb -= (xb >= W ? W : 0);
    yb -= (yb >= H ? H : 0);
    const unsigned ux0 = min((unsigned)xa, (unsigned)(W - 1)), ux1 = min((unsigned)xb, (unsigned)(W - 1));
    const unsigned r0 = min((unsigned)ya, (unsigned)(H - 1)) * (unsigned)W + ib;
    const unsigned r1 = min((unsigned)yb, (unsigned)(H - 1)) * (unsigned)W + ib;
    Taps t;
    t.a = r0 + ux0;
    t.b = r0 + ux1;
    t.c = r1 + ux0;
    t.d = r1 + ux1;
    t.wa = dy1 * dx1;
    t.wb = dy1 * dx0;
    t.wc = dy0 * dx1;
    t.wd = dy0 * dx0;
    return t;
}


// ------------------------------------------------------------------------------------------
// Fast form of the render-side chain (intersect_sphere + project_spherical + theta_phi_to_pixels,
// spherical.py:268-326, 235-246, 54-68) used by the fused render kernel.  The render side has no
// validity mask, so nothing here needs the reference's exact bits: approximate reciprocal / square
// root, FMA, a polynomial atan2 (max abs error 2.5e-7 rad) and the pixel scaling folded into one
// multiply.  Coordinates agree with the strict chain (sphere_hit_uv) to ~1e-4 px; the contract
// tested is 1e-3 px, floor() flips only at knife-edge coordinates.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_sqrtf(float x) {  // MUFU.SQRT, ~1 ulp
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcpf(float x) {  // MUFU.RCP, ~1 ulp (no scaling for denormal / huge x)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float q = mn * fast_rcpf(mx);
    q = (mx == 0.0f) ? 0.0f : q;  // atan2(0, 0) = 0
    const float s = q * q;
    // atan(q) / q on [0, 1] as a degree-7 polynomial in q^2 (minimax fit: 1.3e-7 abs in float32)
    float p = -0.004054469987750053f;
    p = __fmaf_rn(p, s, 0.02186259813606739f);
    p = __fmaf_rn(p, s, -0.05591179430484772f);
    p = __fmaf_rn(p, s, 0.09642156958580017f);
    p = __fmaf_rn(p, s, -0.13908612728118896f);
    p = __fmaf_rn(p, s, 0.19946561753749847f);
    p = __fmaf_rn(p, s, -0.33329859375953674f);
    p = __fmaf_rn(p, s, 0.9999993443489075f);
    float r = p * q;
    r = (ay > ax) ? (1.5707963267948966f - r) : r;
    r = (x < 0.0f) ? (3.141592653589793f - r) : r;
    return copysignf(r, y);
}

// layer-independent part of a ray for the fast chain
struct FastRay {
    float rx, ry, rz, cx, cy, cz;
    float b, b2, a4, inv2a, cc;  // b, b^2, 4a, 1/(2a), |c|^2
    float pad;
};
__device__ __forceinline__ FastRay make_fast_ray(const SphereRay& q) {
    FastRay f;
    f.rx = q.rx;
    f.ry = q.ry;
    f.rz = q.rz;
    f.cx = q.cx;
    f.cy = q.cy;
    f.cz = q.cz;
    f.b = q.b;
    f.b2 = q.b * q.b;
    f.a4 = 4.0f * q.a;
    f.inv2a = 1.0f / (2.0f * q.a);
    f.cc = (q.cx * q.cx + q.cy * q.cy) + q.cz * q.cz;
    f.pad = 0.0f;
    return f;
}
__device__ __forceinline__ void sphere_hit_uv_fast(const FastRay& q, float radius2, const ErpConsts& k, float& u, float& v) {
    const float c = q.cc - radius2;
    const float disc = __fmaf_rn(-q.a4, c, q.b2);
    const float t = (fast_sqrtf(disc) - q.b) * q.inv2a;
    const float x = __fmaf_rn(t, q.rx, q.cx);
    const float y = __fmaf_rn(t, q.ry, q.cy);
    const float z = __fmaf_rn(t, q.rz, q.cz);
    const float theta = -fast_atan2f(z, x);
    const float h = fast_sqrtf(__fmaf_rn(x, x, z * z));
    const float phi = fast_atan2f(y, h);
    u = ((theta + k.pi) - k.pi_w) * k.ku;
    v = ((phi + k.half_pi) - k.half_pi_h) * k.kv;
}

}  // namespace msi
