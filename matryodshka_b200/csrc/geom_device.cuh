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
    return c;
}

// spherical.project_ods (spherical.py:170-233, tuple branch) for one 3-D point.
// order = +1 / -1, r = ODS baseline.  Returns false (and u = v = 1) when disc < 0.
__device__ __forceinline__ bool project_ods_point(float x, float y, float z, float order, float r,
                                                  const ErpConsts& k, float& u, float& v) {
    const float f = r * r - (x * x + z * z);
    const bool zl = fabsf(z) > fabsf(x);
    const float px = zl ? x : z;
    const float pz = zl ? z : x;
    const float pz2 = pz * pz;
    const float a = 1.0f + (px * px) / pz2;
    const float b = ((-2.0f * f) * px) / pz2;
    const float c = f + (f * f) / pz2;
    const float disc = b * b - (4.0f * a) * c;
    const float sgn = (pz > 0.0f) ? 1.0f : ((pz < 0.0f) ? -1.0f : 0.0f);
    float s = ((-order) * sgn) * sqrtf(disc);
    s = zl ? s : -s;
    const float dx0 = (-b + s) / (2.0f * a);
    const float dz0 = (f - px * dx0) / pz;
    const float dx = zl ? -dx0 : -dz0;
    const float dz = zl ? -dz0 : -dx0;
    const float theta = -atan2f(dz, dx);
    float phi = atan2f(y, sqrtf(dx * dx + dz * dz));
    if (phi != phi) phi = 1.0f;
    phi = (phi <= k.half_pi) ? phi : k.half_pi;
    phi = (phi >= -k.half_pi) ? phi : -k.half_pi;
    u = ((theta + k.pi - k.pi_w) / k.den_u) * k.wm1;
    v = ((phi + k.half_pi - k.half_pi_h) / k.den_v) * k.hm1;
    const bool valid = disc >= 0.0f;
    if (!valid) {
        u = 1.0f;
        v = 1.0f;
    }
    return valid;
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
    s.x0 = floor_mod(x0 + W, W);
    s.y0 = floor_mod(y0 + H, H);
    s.x1 = floor_mod(x1 + W, W);
    s.y1 = floor_mod(y1 + H, H);
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

}  // namespace msi
