// Stage 1 (plane-sweep volume), stage 2 1/2 (RGBA assembly) and stage 3 (reprojection +
// over-composite) of the MSI inference path.  HBM-bound gather / blend kernels.
// Compiled with -fmad=false (see geom_device.cuh).
#include "geom_device.cuh"

namespace msi {

// ------------------------------------------------------------------------------------------
// K1  psv_build: one thread per (pixel, eye, plane); coordinates live in registers, the four
// bilinear taps come from the (L1/L2-resident, 2.5 MB) source image, and the 3 output floats
// of 256 consecutive threads are staged in shared memory so that the block writes 3 KB of
// contiguous PSV with 128-bit stores (float32) / 64-bit-per-8-channels stores (fp16 hi/lo).
// ------------------------------------------------------------------------------------------
struct PsvParams {
    const void* img[2];
    const float* poses;      // [B,2,16]
    const float* baselines;  // [B]
    const float* depths;     // [P]
    const float *cos_s, *sin_s, *cos_t, *sin_t;
    int B, H, W, P;
    int preprocess;
    float* out_f32;
    __half* out_hi;
    __half* out_lo;
    int c_stride;
    ErpConsts k;
    // cached sweep coordinates (msi_sweep_table_build): [table_frames][H*W][P] float4 (u_ref, v_ref, u_src, v_src),
    // table_frames = 1 (one rig shared by every frame of the batch) or B; null = evaluate the chain
    const float4* table;
    int table_frames;
};

template <typename T>
__device__ __forceinline__ float load_img(const T* img, size_t off, int preprocess);
template <>
__device__ __forceinline__ float load_img<float>(const float* img, size_t off, int preprocess) {
    float v = __ldg(img + off);
    return preprocess ? (v * 2.0f - 1.0f) : v;
}
template <>
__device__ __forceinline__ float load_img<uint8_t>(const uint8_t* img, size_t off, int preprocess) {
    // tf.image.convert_image_dtype(uint8 -> float32): cast * (1.0 / 255)
    float v = (float)__ldg(img + off) * (float)(1.0 / 255);
    return preprocess ? (v * 2.0f - 1.0f) : v;
}

template <typename T, bool TABLE>
__global__ void __launch_bounds__(256) psv_build_kernel(PsvParams p) {
    __shared__ __align__(16) float stage[768];
    const long long total = (long long)p.B * p.H * p.W * 2 * p.P;
    const long long base = (long long)blockIdx.x * 256;
    const long long idx = base + threadIdx.x;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    long long pix = 0;
    int e = 0, pl = 0;
    if (idx < total) {
        pl = (int)(idx % p.P);
        e = (int)((idx / p.P) & 1);
        pix = idx / (2 * p.P);
        const int j = (int)(pix % p.W);
        const int i = (int)((pix / p.W) % p.H);
        const int b = (int)(pix / ((long long)p.W * p.H));
        float u, v;
        if (TABLE) {
            const size_t row = (size_t)(p.table_frames == 1 ? 0 : b) * p.H * p.W + (size_t)i * p.W + j;
            const float4 uv = __ldcs(p.table + row * p.P + pl);
            u = e ? uv.z : uv.x;
            v = e ? uv.w : uv.y;
        } else {
            sweep_uv(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                     __ldg(p.depths + pl), p.poses + (b * 2 + e) * 16, e == 0 ? 1.0f : -1.0f,
                     __ldg(p.baselines + b), p.k, u, v);
        }
        const Bilinear s = bilinear_setup(u, v, p.W, p.H);
        const T* img = reinterpret_cast<const T*>(p.img[e]);
        const size_t ib = (size_t)b * p.H * p.W;
        const size_t oa = (ib + (size_t)s.y0 * p.W + s.x0) * 3;
        const size_t ob = (ib + (size_t)s.y0 * p.W + s.x1) * 3;
        const size_t oc = (ib + (size_t)s.y1 * p.W + s.x0) * 3;
        const size_t od = (ib + (size_t)s.y1 * p.W + s.x1) * 3;
        r0 = blend4(s, load_img<T>(img, oa + 0, p.preprocess), load_img<T>(img, ob + 0, p.preprocess),
                    load_img<T>(img, oc + 0, p.preprocess), load_img<T>(img, od + 0, p.preprocess));
        r1 = blend4(s, load_img<T>(img, oa + 1, p.preprocess), load_img<T>(img, ob + 1, p.preprocess),
                    load_img<T>(img, oc + 1, p.preprocess), load_img<T>(img, od + 1, p.preprocess));
        r2 = blend4(s, load_img<T>(img, oa + 2, p.preprocess), load_img<T>(img, ob + 2, p.preprocess),
                    load_img<T>(img, oc + 2, p.preprocess), load_img<T>(img, od + 2, p.preprocess));
    }
    const bool dense = (base + 256 <= total) && (p.c_stride == 6 * p.P);
    if (dense) {
        // thread idx owns floats [3*idx, 3*idx+3) of the dense [B,H,W,6P] tensor
        stage[3 * threadIdx.x + 0] = r0;
        stage[3 * threadIdx.x + 1] = r1;
        stage[3 * threadIdx.x + 2] = r2;
        __syncthreads();
        const size_t fbase = (size_t)base * 3;  // multiple of 768
        if (p.out_f32 != nullptr && threadIdx.x < 192) {
            const float4 q = reinterpret_cast<const float4*>(stage)[threadIdx.x];
            reinterpret_cast<float4*>(p.out_f32 + fbase)[threadIdx.x] = q;
        }
        if (p.out_hi != nullptr && threadIdx.x < 96) {
            __align__(16) __half hi[8];
            __align__(16) __half lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) split_half(stage[8 * threadIdx.x + q] * MSI_ACT_SCALE, hi[q], lo[q]);
            reinterpret_cast<uint4*>(p.out_hi + fbase)[threadIdx.x] = *reinterpret_cast<const uint4*>(hi);
            if (p.out_lo != nullptr)
                reinterpret_cast<uint4*>(p.out_lo + fbase)[threadIdx.x] = *reinterpret_cast<const uint4*>(lo);
        }
    } else if (idx < total) {
        const int ch = e * 3 * p.P + pl * 3;
        if (p.out_f32 != nullptr) {
            float* o = p.out_f32 + (size_t)pix * 6 * p.P + ch;
            o[0] = r0;
            o[1] = r1;
            o[2] = r2;
        }
        if (p.out_hi != nullptr) {
            const size_t o = (size_t)pix * p.c_stride + ch;
            const float r[3] = {r0, r1, r2};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                __half hi, lo;
                split_half(r[q] * MSI_ACT_SCALE, hi, lo);
                p.out_hi[o + q] = hi;
                if (p.out_lo != nullptr) p.out_lo[o + q] = lo;
            }
        }
    }
}

// K1, two-eye form (used when the caller provides scratch and 256 % P == 0).
//   pre-pass: both images -> float4 RGBX, preprocessed once per pixel (not once per tap);
//   main:     one thread per (pixel, plane) produces BOTH eyes: when the two eyes have the same pose
//             (ODS data: both identity) the quadratic of project_ods is evaluated once and only the
//             sign of its root differs (order = +1 / -1), bit for bit as in two separate evaluations;
//             taps are four 128-bit loads per eye; the block stages its 256/P pixels x 6P channels
//             (6 KB, contiguous in the PSV) in shared memory and stores 128-bit.
template <typename T>
__global__ void __launch_bounds__(256)
prep_images_kernel(const T* __restrict__ ref, const T* __restrict__ src, long long npix, int preprocess,
                   float4* __restrict__ out, float scale = 1.0f) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= 2 * npix) return;
    const int e = idx >= npix;
    const long long pix = idx - (e ? npix : 0);
    const T* img = e ? src : ref;
    float4 o;
    // scale is 1 or MSI_ACT_SCALE (a power of two: exact, and it commutes with the bilinear blend bit for bit)
    o.x = load_img<T>(img, (size_t)pix * 3 + 0, preprocess) * scale;
    o.y = load_img<T>(img, (size_t)pix * 3 + 1, preprocess) * scale;
    o.z = load_img<T>(img, (size_t)pix * 3 + 2, preprocess) * scale;
    o.w = 0.f;
    out[idx] = o;
}

__device__ __forceinline__ void sample_rgbx(const float4* __restrict__ img, size_t ib, int W, int H, float u, float v,
                                            float& r0, float& r1, float& r2) {
    const Bilinear s = bilinear_setup(u, v, W, H);
    const float4 pa = __ldg(img + ib + (size_t)s.y0 * W + s.x0);
    const float4 pb = __ldg(img + ib + (size_t)s.y0 * W + s.x1);
    const float4 pc = __ldg(img + ib + (size_t)s.y1 * W + s.x0);
    const float4 pd = __ldg(img + ib + (size_t)s.y1 * W + s.x1);
    r0 = blend4(s, pa.x, pb.x, pc.x, pd.x);
    r1 = blend4(s, pa.y, pb.y, pc.y, pd.y);
    r2 = blend4(s, pa.z, pb.z, pc.z, pd.z);
}

__global__ void __launch_bounds__(256) psv_build_pair_kernel(PsvParams p, const float4* __restrict__ rgbx) {
    __shared__ __align__(16) float stage[1536];
    __shared__ int s_b[64], s_i[64], s_j[64], s_same[64];
    const int P = p.P;
    const unsigned ppb = 256u / (unsigned)P;  // pixels per block (<= 64: P >= 4)
    const unsigned npix = (unsigned)p.B * (unsigned)p.H * (unsigned)p.W;  // < 2^31 (checked on the host)
    const unsigned pix0 = blockIdx.x * ppb;
    if (threadIdx.x < ppb) {
        // decode the block's pixels once (32-bit), and compare the two eye poses of their frame once
        const unsigned px = pix0 + threadIdx.x;
        const unsigned row = px / (unsigned)p.W;
        const unsigned bb = row / (unsigned)p.H;
        s_j[threadIdx.x] = (int)(px - row * (unsigned)p.W);
        s_i[threadIdx.x] = (int)(row - bb * (unsigned)p.H);
        s_b[threadIdx.x] = (int)bb;
        int same = 1;
        if (px < npix) {
            const float* pose0 = p.poses + (size_t)bb * 32;
            for (int q = 0; q < 12; ++q) same &= (__ldg(pose0 + q) == __ldg(pose0 + 16 + q)) ? 1 : 0;
        }
        s_same[threadIdx.x] = same;
    }
    __syncthreads();
    const unsigned lp = threadIdx.x / (unsigned)P;
    const int pl = (int)(threadIdx.x - lp * (unsigned)P);
    const unsigned pix = pix0 + lp;
    if (pix < npix) {
        const int j = s_j[lp], i = s_i[lp], b = s_b[lp];
        const bool same = s_same[lp] != 0;
        const float cs = __ldg(p.cos_s + j), sn = __ldg(p.sin_s + j), ct = __ldg(p.cos_t + i), st = __ldg(p.sin_t + i);
        const float depth = __ldg(p.depths + pl);
        const float r = __ldg(p.baselines + b);
        const float* pose0 = p.poses + (size_t)b * 32;
        const float* pose1 = pose0 + 16;
        float x, y, z;
        sweep_point(cs, sn, ct, st, depth, pose0, x, y, z);
        const OdsQuad q0 = ods_quadratic(x, y, z, r);
        OdsQuad q1 = q0;
        if (!same) {
            sweep_point(cs, sn, ct, st, depth, pose1, x, y, z);
            q1 = ods_quadratic(x, y, z, r);
        }
        float u0, v0, u1, v1;
        ods_finish(q0, 1.0f, p.k, u0, v0);
        ods_finish(q1, -1.0f, p.k, u1, v1);
        const size_t ib = (size_t)b * p.H * p.W;
        float* s0 = stage + lp * 6 * P + 3 * pl;
        sample_rgbx(rgbx, ib, p.W, p.H, u0, v0, s0[0], s0[1], s0[2]);
        float* s1 = s0 + 3 * P;
        sample_rgbx(rgbx + npix, ib, p.W, p.H, u1, v1, s1[0], s1[1], s1[2]);
    }
    __syncthreads();
    const unsigned nvalid = (npix - pix0 < ppb) ? (npix - pix0) : ppb;  // pixels of this block inside the tensor
    const int nfl = (int)nvalid * 6 * P;                                  // multiple of 6P
    const size_t fbase = (size_t)pix0 * 6 * P;                            // multiple of 1536 floats
    if (p.out_f32 != nullptr)
        for (int q = threadIdx.x; q < nfl / 4; q += 256)
            reinterpret_cast<float4*>(p.out_f32 + fbase)[q] = reinterpret_cast<const float4*>(stage)[q];
    if (p.out_hi != nullptr)
        for (int q = threadIdx.x; q < nfl / 8; q += 256) {
            __align__(16) __half hi[8];
            __align__(16) __half lo[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) split_half(stage[8 * q + t] * MSI_ACT_SCALE, hi[t], lo[t]);
            reinterpret_cast<uint4*>(p.out_hi + fbase)[q] = *reinterpret_cast<const uint4*>(hi);
            if (p.out_lo != nullptr) reinterpret_cast<uint4*>(p.out_lo + fbase)[q] = *reinterpret_cast<const uint4*>(lo);
        }
}

// ------------------------------------------------------------------------------------------
// K1 with cached coordinates.  The sweep coordinates depend only on (eye poses, baseline, depths,
// H, W) -- not on the images -- and a rig is static over a sequence (every frame of the reference's
// data has identity eye poses and one baseline, data_loader.py:146-174), so they are evaluated ONCE
// by sweep_table_kernel with the exact chain above (same device functions, same bits) and the
// per-frame kernel is a pure gather: one 128-bit table load, 8 taps, 6 blends per (pixel, plane).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sweep_table_kernel(PsvParams p, float4* __restrict__ table) {
    // thread per (frame, pixel, plane), plane fastest = the table's layout
    const long long total = (long long)p.B * p.H * p.W * p.P;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int pl = (int)(idx % p.P);
    const long long pix = idx / p.P;
    const int j = (int)(pix % p.W);
    const int i = (int)((pix / p.W) % p.H);
    const int b = (int)(pix / ((long long)p.W * p.H));
    const float cs = __ldg(p.cos_s + j), sn = __ldg(p.sin_s + j), ct = __ldg(p.cos_t + i), st = __ldg(p.sin_t + i);
    const float depth = __ldg(p.depths + pl);
    const float r = __ldg(p.baselines + b);
    float4 o;
    sweep_uv(cs, sn, ct, st, depth, p.poses + (size_t)b * 32, 1.0f, r, p.k, o.x, o.y);
    sweep_uv(cs, sn, ct, st, depth, p.poses + (size_t)b * 32 + 16, -1.0f, r, p.k, o.z, o.w);
    table[idx] = o;
}

// packed form of split_half for two values: the same two roundings per value, with the packing
// converts (full rate) instead of four scalar ones
__device__ __forceinline__ void split_half2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(hi);
    lo = __floats2half2_rn(a - hf.x, b - hf.y);
}

// One bilinear sample of an RGBX image (taps a, b, c, d already loaded): tf.add_n order, no FMA (blend4).
#define MSI_BLEND3(dst, t, a, b, c, d)                                              \
    do {                                                                            \
        (dst)[0] = (((t).wa * (a).x + (t).wb * (b).x) + (t).wc * (c).x) + (t).wd * (d).x; \
        (dst)[1] = (((t).wa * (a).y + (t).wb * (b).y) + (t).wc * (c).y) + (t).wd * (d).y; \
        (dst)[2] = (((t).wa * (a).z + (t).wb * (b).z) + (t).wc * (c).z) + (t).wd * (d).z; \
    } while (0)

// The gather kernel.  grid (ceil(groups / kGatherGroups), B); a group = ppb = 256 / P consecutive pixels
// (P a power of two), thread = (pixel of the group, plane), both eyes.  A block walks kGatherGroups
// consecutive groups and loads the NEXT group's table entry before it works on the current one, so the
// table's DRAM latency hides behind the taps and blends (measured: the one-group form stalled 9 of 12
// warps on the table -> taps dependency).  Per group and thread: one 128-bit table load, eight 128-bit
// taps, six blends.  When the 2 x 2 footprints of a whole warp lie inside the image (all but the seam
// columns and the pole rows) the four taps are base, +1, +W, +W+1 of one address; otherwise the
// floor-mod wrap of sampling.resample (taps_in_range).  rgbx holds the images pre-scaled by
// MSI_ACT_SCALE (exact: a power of two), so the staged values are the conv operand before the split.
constexpr int kGatherGroups = 4;
// min blocks per SM = the register budget.  Measured (scripts/exp_ab.sh): 8 blocks (32 registers, taps loaded in pairs)
// 75 us, 6 blocks 76, 4 blocks (all eight taps in flight) 80, unconstrained 115: occupancy beats per-thread ILP here
#ifndef MSI_GATHER_MINBLOCKS
#define MSI_GATHER_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(256, MSI_GATHER_MINBLOCKS) psv_gather_pair_kernel(PsvParams p, const float4* __restrict__ rgbx, int log2p) {
    __shared__ __align__(16) float stage[2][1536];
    const int P = p.P, W = p.W, H = p.H;
    const unsigned ppb = 256u >> log2p;
    const unsigned HW = (unsigned)H * (unsigned)W;
    const unsigned lp = threadIdx.x >> log2p;
    const unsigned pl = threadIdx.x & (unsigned)(P - 1);
    const unsigned ib = blockIdx.y * HW;  // < 2^31 (checked on the host)
    const unsigned tb = (p.table_frames == 1 ? 0u : ib);
    const float4* img0 = rgbx + ib;
    const float4* img1 = img0 + (size_t)p.B * HW;
    const unsigned g0 = blockIdx.x * kGatherGroups;
    const float4 dummy = make_float4(1.f, 1.f, 1.f, 1.f);  // lanes past the frame sample pixel (1, 1); nothing is stored for them
    float4 uv_next = dummy;
    {
        const unsigned pixf = g0 * ppb + lp;
        if (pixf < HW) uv_next = __ldcs(p.table + (((size_t)(tb + pixf) << log2p) + pl));
    }
#pragma unroll 1
    for (int it = 0; it < kGatherGroups; ++it) {
        const unsigned pixf0 = (g0 + it) * ppb;
        if (pixf0 >= HW) break;
        const float4 uv = uv_next;
        uv_next = dummy;
        if (it + 1 < kGatherGroups) {
            const unsigned pixn = pixf0 + ppb + lp;
            if (pixn < HW) uv_next = __ldcs(p.table + (((size_t)(tb + pixn) << log2p) + pl));
        }
        float* s0 = stage[it & 1] + (lp * 6u << log2p) + 3u * pl;
        float* s1 = s0 + 3 * P;
        const int x0 = __float2int_rd(uv.x), y0 = __float2int_rd(uv.y), x1 = __float2int_rd(uv.z), y1 = __float2int_rd(uv.w);
        const bool inside = ((unsigned)x0 < (unsigned)(W - 1)) & ((unsigned)y0 < (unsigned)(H - 1)) &
                            ((unsigned)x1 < (unsigned)(W - 1)) & ((unsigned)y1 < (unsigned)(H - 1));
        if (__all_sync(0xffffffffu, inside)) {
            const float4* q0 = img0 + (unsigned)(y0 * W + x0);
            const float4* q1 = img1 + (unsigned)(y1 * W + x1);
            float4 a0 = __ldg(q0), b0 = __ldg(q0 + 1), c0 = __ldg(q0 + W), d0 = __ldg(q0 + W + 1);
            float4 a1 = __ldg(q1), b1 = __ldg(q1 + 1), c1 = __ldg(q1 + W), d1 = __ldg(q1 + W + 1);
            // all eight taps are issued before the first one is consumed (the compiler otherwise pairs them up to save
            // registers, and the kernel is bound by load latency, not by occupancy)
            asm volatile("" : "+f"(a0.x), "+f"(b0.x), "+f"(c0.x), "+f"(d0.x), "+f"(a1.x), "+f"(b1.x), "+f"(c1.x), "+f"(d1.x));
            Taps t0, t1;
            {
                const float fx = (float)x0, fy = (float)y0;
                const float dx0 = uv.x - fx, dy0 = uv.y - fy, dx1 = (fx + 1.0f) - uv.x, dy1 = (fy + 1.0f) - uv.y;
                t0.wa = dy1 * dx1;
                t0.wb = dy1 * dx0;
                t0.wc = dy0 * dx1;
                t0.wd = dy0 * dx0;
            }
            {
                const float fx = (float)x1, fy = (float)y1;
                const float dx0 = uv.z - fx, dy0 = uv.w - fy, dx1 = (fx + 1.0f) - uv.z, dy1 = (fy + 1.0f) - uv.w;
                t1.wa = dy1 * dx1;
                t1.wb = dy1 * dx0;
                t1.wc = dy0 * dx1;
                t1.wd = dy0 * dx0;
            }
            MSI_BLEND3(s0, t0, a0, b0, c0, d0);
            MSI_BLEND3(s1, t1, a1, b1, c1, d1);
        } else {
            const Taps t0 = taps_in_range(uv.x, uv.y, W, H, 0u);
            const Taps t1 = taps_in_range(uv.z, uv.w, W, H, 0u);
            const float4 a0 = __ldg(img0 + t0.a), b0 = __ldg(img0 + t0.b), c0 = __ldg(img0 + t0.c), d0 = __ldg(img0 + t0.d);
            const float4 a1 = __ldg(img1 + t1.a), b1 = __ldg(img1 + t1.b), c1 = __ldg(img1 + t1.c), d1 = __ldg(img1 + t1.d);
            MSI_BLEND3(s0, t0, a0, b0, c0, d0);
            MSI_BLEND3(s1, t1, a1, b1, c1, d1);
        }
        __syncthreads();
        // (two stage buffers: the next iteration's writers cannot overtake this iteration's readers by more
        // than one barrier)
        const float* st = stage[it & 1];
        const unsigned nvalid = (HW - pixf0 < ppb) ? (HW - pixf0) : ppb;
        const int nfl = (int)nvalid * 6 * P;
        const size_t fbase = ((size_t)ib + pixf0) * 6 * P;
        if (p.out_f32 != nullptr)
            for (int q = threadIdx.x; q < nfl / 4; q += 256) {
                float4 v = reinterpret_cast<const float4*>(st)[q];
                v.x *= 1.0f / MSI_ACT_SCALE;
                v.y *= 1.0f / MSI_ACT_SCALE;
                v.z *= 1.0f / MSI_ACT_SCALE;
                v.w *= 1.0f / MSI_ACT_SCALE;
                reinterpret_cast<float4*>(p.out_f32 + fbase)[q] = v;
            }
        if (p.out_hi != nullptr) {
            const int q = threadIdx.x;  // 1536 floats = 192 groups of 8: one group per thread
            if (q < nfl / 8) {
                const float4 a = reinterpret_cast<const float4*>(st)[2 * q];
                const float4 c = reinterpret_cast<const float4*>(st)[2 * q + 1];
                __half2 h0, h1, h2, h3, l0, l1, l2, l3;
                split_half2(a.x, a.y, h0, l0);
                split_half2(a.z, a.w, h1, l1);
                split_half2(c.x, c.y, h2, l2);
                split_half2(c.z, c.w, h3, l3);
                uint4 hv, lv;
                hv.x = *reinterpret_cast<const unsigned*>(&h0);
                hv.y = *reinterpret_cast<const unsigned*>(&h1);
                hv.z = *reinterpret_cast<const unsigned*>(&h2);
                hv.w = *reinterpret_cast<const unsigned*>(&h3);
                lv.x = *reinterpret_cast<const unsigned*>(&l0);
                lv.y = *reinterpret_cast<const unsigned*>(&l1);
                lv.z = *reinterpret_cast<const unsigned*>(&l2);
                lv.w = *reinterpret_cast<const unsigned*>(&l3);
                reinterpret_cast<uint4*>(p.out_hi + fbase)[q] = hv;
                if (p.out_lo != nullptr) reinterpret_cast<uint4*>(p.out_lo + fbase)[q] = lv;
            }
        }
    }
}

__global__ void __launch_bounds__(256) sweep_coords_kernel(PsvParams p, float* uv, uint8_t* valid) {
    // output order [B,2,P,H,W]: thread per element, W fastest
    const long long total = (long long)p.B * 2 * p.P * p.H * p.W;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % p.W);
    const int i = (int)((idx / p.W) % p.H);
    const int pl = (int)((idx / ((long long)p.W * p.H)) % p.P);
    const int e = (int)((idx / ((long long)p.W * p.H * p.P)) & 1);
    const int b = (int)(idx / ((long long)p.W * p.H * p.P * 2));
    float u, v;
    const bool ok = sweep_uv(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                             __ldg(p.depths + pl), p.poses + (b * 2 + e) * 16, e == 0 ? 1.0f : -1.0f,
                             __ldg(p.baselines + b), p.k, u, v);
    uv[2 * idx + 0] = u;
    uv[2 * idx + 1] = v;
    if (valid != nullptr) valid[idx] = ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// K4  rgba_assemble: thread per (pixel, layer); lanes = layers, so pred / PSV reads and the
// float4 RGBA store of a warp are contiguous.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rgba_assemble_kernel(const float* __restrict__ pred, const float* __restrict__ psv_f32,
                     const __half* __restrict__ psv_hi, const __half* __restrict__ psv_lo, int c_stride,
                     long long npix, int L, int pred_stride, float4* __restrict__ rgba, float* __restrict__ bw_out,
                     float* __restrict__ al_out) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= npix * L) return;
    const long long pix = idx / L;
    const int l = (int)(idx % L);
    const float w = (__ldg(pred + pix * pred_stride + l) + 1.0f) / 2.0f;
    const float al = (__ldg(pred + pix * pred_stride + L + l) + 1.0f) / 2.0f;
    float fg[3], bg[3];
    if (psv_f32 != nullptr) {
        const float* f = psv_f32 + pix * 6 * L + 3 * l;
        const float* g = psv_f32 + pix * 6 * L + 3 * (L + l);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            fg[c] = __ldg(f + c);
            bg[c] = __ldg(g + c);
        }
    } else {
        const size_t f = (size_t)pix * c_stride + 3 * l;
        const size_t g = (size_t)pix * c_stride + 3 * (L + l);
        const float inv = 1.0f / MSI_ACT_SCALE;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            fg[c] = (__half2float(psv_hi[f + c]) + __half2float(psv_lo[f + c])) * inv;
            bg[c] = (__half2float(psv_hi[g + c]) + __half2float(psv_lo[g + c])) * inv;
        }
    }
    const float omw = 1.0f - w;
    float4 o;
    o.x = w * fg[0] + omw * bg[0];
    o.y = w * fg[1] + omw * bg[1];
    o.z = w * fg[2] + omw * bg[2];
    o.w = al;
    rgba[idx] = o;
    if (bw_out != nullptr) bw_out[idx] = w;
    if (al_out != nullptr) al_out[idx] = al;
}

// ------------------------------------------------------------------------------------------
// K5  render_composite: a block owns 32 consecutive output pixels.
//   phase 1 (all 8 warps): lanes = layers.  Each thread computes the ray/sphere hit of its
//     (pixel, layer), gathers the 4 bilinear RGBA taps as float4 (a warp's taps of one source
//     texel are one contiguous 512-byte run of [.., L, 4]) and parks the sample in shared memory;
//   phase 2 (4 warps): lane = pixel, warp = channel (r, g, b, depth).  Each thread runs the
//     back-to-front recurrence of over_composite / over_composite_depth in the reference's
//     exact order, so colour AND depth come out of ONE reprojection (the reference does two).
// ------------------------------------------------------------------------------------------
struct RenderParams {
    const float4* rgba;      // [B,H,W,L] float4
    const float* pose_rt;    // [B,16]
    const float* tgt_pos;    // [B,3]
    const float* depths;     // [L]
    const float *cos_s, *sin_s, *cos_t, *sin_t;
    int B, H, W, L;
    float* out_rgb;
    float* out_depth;
    uint8_t* out_rgb_u8;
    uint8_t* out_depth_u8;
    ErpConsts k;
    int ods_mode;            // 0: target ERP view (intersect_sphere); 1: ODS eye view (intersect_ods);
                             // 2: perspective view (intersect_perspective; cos_s / cos_t hold the uv_grid axes)
    int oH, oW;              // output image size (= H, W except for the perspective view)
    // Fused output all-gather: besides its local outputs the kernel stores the uint8
    // view into the gathered buffer [world * B, H, W, 3] of EVERY rank -- through the NVSwitch
    // multicast address of the symmetric buffer when there is one (one multimem.st per word), else
    // with one peer store per rank -- so no collective kernel runs after it.
    uint8_t* const* peer_u8;  // device array of n_peers base pointers (null: no gather)
    int n_peers;
    uint8_t* mc_u8;           // multicast address of the same buffer (null: per-peer stores)
    long long gather_off;     // byte offset of this rank's first frame in the gathered buffer
    float ods_order;         // +1 / -1
    const float* baselines;  // [B] (ODS mode)
};

__device__ __forceinline__ float4 sample_layer(const RenderParams& p, int b, int i, int j, int l) {
    float u, v;
    sphere_uv(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
              p.pose_rt + b * 16, p.tgt_pos + b * 3, __ldg(p.depths + l), p.k, u, v);
    const Bilinear s = bilinear_setup(u, v, p.W, p.H);
    const size_t ib = (size_t)b * p.H * p.W;
    const float4 pa = __ldg(p.rgba + (ib + (size_t)s.y0 * p.W + s.x0) * p.L + l);
    const float4 pb = __ldg(p.rgba + (ib + (size_t)s.y0 * p.W + s.x1) * p.L + l);
    const float4 pc = __ldg(p.rgba + (ib + (size_t)s.y1 * p.W + s.x0) * p.L + l);
    const float4 pd = __ldg(p.rgba + (ib + (size_t)s.y1 * p.W + s.x1) * p.L + l);
    float4 o;
    o.x = blend4(s, pa.x, pb.x, pc.x, pd.x);
    o.y = blend4(s, pa.y, pb.y, pc.y, pd.y);
    o.z = blend4(s, pa.z, pb.z, pc.z, pd.z);
    o.w = blend4(s, pa.w, pb.w, pc.w, pd.w);
    return o;
}

__device__ __forceinline__ uint8_t to_u8(float x) {
    // tf.image.convert_image_dtype(float -> uint8, saturate=False): truncating cast of x * 255.5
    return (uint8_t)(__float2int_rz(x * 255.5f) & 0xff);
}

__global__ void __launch_bounds__(256) render_composite_kernel(RenderParams p) {
    extern __shared__ float sm[];  // [4][32][L+1] samples, then [L] depth fractions
    const int L = p.L;
    const int ld = L + 1;
    float* frac = sm + 4 * 32 * ld;
    __shared__ SphereRay s_ray[32];
    __shared__ int s_b[32];
    __shared__ __align__(16) uint8_t s_u8[96];
    const long long npix = (long long)p.B * p.oH * p.oW;
    const long long pix0 = (long long)blockIdx.x * 32;

    for (int l = threadIdx.x; l < L; l += 256) frac[l] = (float)((double)l / (double)L);
    if (threadIdx.x >= 224) {
        // the layer-independent ray of each of the block's 32 pixels, once
        const int q = threadIdx.x - 224;
        const long long pix = pix0 + q;
        int b = 0;
        SphereRay ray = {};
        if (pix < npix) {
            const int j = (int)(pix % p.oW);
            const int i = (int)((pix / p.oW) % p.oH);
            b = (int)(pix / ((long long)p.oW * p.oH));
            if (p.ods_mode == 2)
                ray = sphere_ray_perspective(__ldg(p.cos_s + j), __ldg(p.cos_t + i), p.pose_rt + b * 16, p.tgt_pos + b * 3);
            else if (p.ods_mode)
                ray = sphere_ray_ods(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                                     p.pose_rt + b * 16, p.ods_order, __ldg(p.baselines + b));
            else
                ray = sphere_ray(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                                 p.pose_rt + b * 16, p.tgt_pos + b * 3);
        }
        s_ray[q] = ray;
        s_b[q] = b;
    }
    __syncthreads();

    {
        // walk (pixel, layer) incrementally: no division in the sample loop
        int q = threadIdx.x / L;
        int l = threadIdx.x - q * L;
        const int dq = 256 / L, dl = 256 - dq * L;
        for (int s = threadIdx.x; s < 32 * L; s += 256) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pix0 + q < npix) {
                float u, v;
                sphere_hit_uv(s_ray[q], __ldg(p.depths + l), p.k, u, v);
                const Bilinear sb = bilinear_setup(u, v, p.W, p.H);
                const float4* base = p.rgba + (size_t)s_b[q] * p.H * p.W * L + l;
                const float4 pa = __ldg(base + ((size_t)sb.y0 * p.W + sb.x0) * L);
                const float4 pb = __ldg(base + ((size_t)sb.y0 * p.W + sb.x1) * L);
                const float4 pc = __ldg(base + ((size_t)sb.y1 * p.W + sb.x0) * L);
                const float4 pd = __ldg(base + ((size_t)sb.y1 * p.W + sb.x1) * L);
                o.x = blend4(sb, pa.x, pb.x, pc.x, pd.x);
                o.y = blend4(sb, pa.y, pb.y, pc.y, pd.y);
                o.z = blend4(sb, pa.z, pb.z, pc.z, pd.z);
                o.w = blend4(sb, pa.w, pb.w, pc.w, pd.w);
            }
            sm[(0 * 32 + q) * ld + l] = o.x;
            sm[(1 * 32 + q) * ld + l] = o.y;
            sm[(2 * 32 + q) * ld + l] = o.z;
            sm[(3 * 32 + q) * ld + l] = o.w;
            q += dq;
            l += dl;
            if (l >= L) {
                l -= L;
                ++q;
            }
        }
    }
    __syncthreads();

    if (threadIdx.x < 128) {
        const int ch = threadIdx.x >> 5;
        const int q = threadIdx.x & 31;
        const long long pix = pix0 + q;
        const float* col = sm + (ch * 32 + q) * ld;
        const float* alp = sm + (3 * 32 + q) * ld;
        float out;
        if (ch < 3) {
            out = col[0];  // alpha of the farthest layer is ignored (projector.py:257-259)
            for (int l = 1; l < L; ++l) {
                const float a = alp[l];
                out = col[l] * a + out * (1.0f - a);
            }
        } else {
            out = 0.0f;
            for (int l = 1; l < L; ++l) {
                const float a = alp[l];
                out = frac[l] * a + out * (1.0f - a);
            }
        }
        if (pix < npix) {
            if (ch < 3) {
                if (p.out_rgb != nullptr) p.out_rgb[pix * 3 + ch] = out;
                const uint8_t q8 = to_u8((out + 1.0f) / 2.0f);
                if (p.out_rgb_u8 != nullptr) p.out_rgb_u8[pix * 3 + ch] = q8;
                if (p.peer_u8 != nullptr) s_u8[q * 3 + ch] = q8;
            } else {
                if (p.out_depth != nullptr) {
                    p.out_depth[pix * 3 + 0] = out;
                    p.out_depth[pix * 3 + 1] = out;
                    p.out_depth[pix * 3 + 2] = out;
                }
                if (p.out_depth_u8 != nullptr) {
                    const uint8_t d = to_u8(out);
                    p.out_depth_u8[pix * 3 + 0] = d;
                    p.out_depth_u8[pix * 3 + 1] = d;
                    p.out_depth_u8[pix * 3 + 2] = d;
                }
            }
        }
    }
    if (p.peer_u8 != nullptr) {
        // the block's 32 pixels x 3 bytes = 24 words, contiguous in the gathered buffer
        __syncthreads();
        const long long dst = p.gather_off + pix0 * 3;  // multiple of 4: pix0 is a multiple of 32
        const int nbytes = (int)min((long long)96, (npix - pix0) * 3);
        if (nbytes == 96) {
            if (threadIdx.x < 24) {
                const unsigned int w = reinterpret_cast<const unsigned int*>(s_u8)[threadIdx.x];
                if (p.mc_u8 != nullptr) {
                    asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(p.mc_u8 + dst + 4 * threadIdx.x), "r"(w) : "memory");
                } else {
                    for (int k = 0; k < p.n_peers; ++k)
                        *reinterpret_cast<unsigned int*>(p.peer_u8[k] + dst + 4 * threadIdx.x) = w;
                }
            }
        } else if ((int)threadIdx.x < nbytes) {  // ragged last block: byte stores to every peer
            for (int k = 0; k < p.n_peers; ++k) p.peer_u8[k][dst + threadIdx.x] = s_u8[threadIdx.x];
        }
    }
}

// K5, second form (default): a block owns NPIX = 64 consecutive output pixels.
//   set-up: 64 threads evaluate the layer-independent ray of their pixel once (FastRay in shared memory);
//   phase 1 (all 8 warps): lanes = layers; per (pixel, layer) the fast chain of geom_device.cuh -- no IEEE
//     divide / square root, polynomial atan2 -- then four 128-bit taps (a warp's taps of one texel are one
//     512-byte run of [.., L, 4]) and the no-FMA blend of sampling.resample; samples parked in shared memory;
//   phase 2 (all 256 threads): thread = (channel r/g/b/depth, pixel) runs over_composite /
//     over_composite_depth in the reference's order (projector.py:225-265);
//   epilogue: the block's 768 B of float rgb, 768 B of depth and 2 x 192 B of uint8 leave as 128-bit stores.
template <int NPIX>
__global__ void __launch_bounds__(256) render_composite_v2_kernel(RenderParams p) {
    extern __shared__ float sm[];  // [4][NPIX][L+1] samples, then [L] depth fractions, then [L] radius^2
    const int L = p.L;
    const int ld = L + 1;
    float* frac = sm + 4 * NPIX * ld;
    float* rad2 = frac + L;
    __shared__ __align__(16) FastRay s_ray[NPIX];
    __shared__ unsigned s_ib[NPIX];
    __shared__ __align__(16) float s_out[4][NPIX];
    __shared__ __align__(16) uint8_t s_u8[2][NPIX * 3];
    const long long npix = (long long)p.B * p.oH * p.oW;
    const long long pix0 = (long long)blockIdx.x * NPIX;

    for (int l = threadIdx.x; l < L; l += 256) {
        frac[l] = (float)((double)l / (double)L);
        const float r = __ldg(p.depths + l);
        rad2[l] = r * r;
    }
    if (threadIdx.x >= 256 - NPIX) {
        const int q = threadIdx.x - (256 - NPIX);
        const long long pix = pix0 + q;
        int b = 0;
        SphereRay ray = {};
        if (pix < npix) {
            const int j = (int)(pix % p.oW);
            const int i = (int)((pix / p.oW) % p.oH);
            b = (int)(pix / ((long long)p.oW * p.oH));
            if (p.ods_mode == 2)
                ray = sphere_ray_perspective(__ldg(p.cos_s + j), __ldg(p.cos_t + i), p.pose_rt + b * 16, p.tgt_pos + b * 3);
            else if (p.ods_mode)
                ray = sphere_ray_ods(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                                     p.pose_rt + b * 16, p.ods_order, __ldg(p.baselines + b));
            else
                ray = sphere_ray(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                                 p.pose_rt + b * 16, p.tgt_pos + b * 3);
        }
        s_ray[q] = make_fast_ray(ray);
        s_ib[q] = (unsigned)b * (unsigned)(p.H * p.W);
    }
    __syncthreads();

    {
        // every thread runs the same number of iterations (warp votes inside); lanes past the work sample pixel (1, 1)
        int q = threadIdx.x / L;
        int l = threadIdx.x - q * L;
        const int dq = 256 / L, dl = 256 - dq * L;
        const int W = p.W, H = p.H;
        const int iters = (NPIX * L + 255) / 256;
        int s = threadIdx.x;
#pragma unroll 1
        for (int it = 0; it < iters; ++it, s += 256) {
            const bool live = (s < NPIX * L) && (pix0 + q < npix);
            float u = 1.0f, v = 1.0f;
            unsigned ib = 0;
            if (live) {
                sphere_hit_uv_fast(s_ray[q], rad2[l], p.k, u, v);
                ib = s_ib[q];
            }
            const float4* base = p.rgba + (live ? l : 0);
            const int x0 = __float2int_rd(u), y0 = __float2int_rd(v);
            const bool inside = ((unsigned)x0 < (unsigned)(W - 1)) & ((unsigned)y0 < (unsigned)(H - 1));
            float4 pa, pb, pc, pd;
            Taps t;
            if (__all_sync(0xffffffffu, inside)) {
                // 2 x 2 footprint inside the image for the whole warp: one address, offsets 1, W, W + 1 texels
                const float4* q0 = base + (size_t)(ib + (unsigned)(y0 * W + x0)) * L;
                pa = __ldg(q0);
                pb = __ldg(q0 + L);
                pc = __ldg(q0 + (size_t)W * L);
                pd = __ldg(q0 + (size_t)W * L + L);
                const float fx = (float)x0, fy = (float)y0;
                const float dx0 = u - fx, dy0 = v - fy, dx1 = (fx + 1.0f) - u, dy1 = (fy + 1.0f) - v;
                t.wa = dy1 * dx1;
                t.wb = dy1 * dx0;
                t.wc = dy0 * dx1;
                t.wd = dy0 * dx0;
            } else {
                t = taps_in_range(u, v, W, H, ib);
                pa = __ldg(base + (size_t)t.a * L);
                pb = __ldg(base + (size_t)t.b * L);
                pc = __ldg(base + (size_t)t.c * L);
                pd = __ldg(base + (size_t)t.d * L);
            }
            if (s < NPIX * L) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    o.x = ((t.wa * pa.x + t.wb * pb.x) + t.wc * pc.x) + t.wd * pd.x;
                    o.y = ((t.wa * pa.y + t.wb * pb.y) + t.wc * pc.y) + t.wd * pd.y;
                    o.z = ((t.wa * pa.z + t.wb * pb.z) + t.wc * pc.z) + t.wd * pd.z;
                    o.w = ((t.wa * pa.w + t.wb * pb.w) + t.wc * pc.w) + t.wd * pd.w;
                }
                sm[(0 * NPIX + q) * ld + l] = o.x;
                sm[(1 * NPIX + q) * ld + l] = o.y;
                sm[(2 * NPIX + q) * ld + l] = o.z;
                sm[(3 * NPIX + q) * ld + l] = o.w;
            }
            q += dq;
            l += dl;
            if (l >= L) {
                l -= L;
                ++q;
            }
        }
    }
    __syncthreads();

    if (threadIdx.x < 4 * NPIX) {
        const int ch = threadIdx.x / NPIX;
        const int q = threadIdx.x - ch * NPIX;
        const float* col = sm + (ch * NPIX + q) * ld;
        const float* alp = sm + (3 * NPIX + q) * ld;
        float out;
        if (ch < 3) {
            out = col[0];  // alpha of the farthest layer is ignored (projector.py:257-259)
            for (int l = 1; l < L; ++l) {
                const float a = alp[l];
                out = col[l] * a + out * (1.0f - a);
            }
            s_u8[0][q * 3 + ch] = to_u8((out + 1.0f) / 2.0f);
        } else {
            out = 0.0f;
            for (int l = 1; l < L; ++l) {
                const float a = alp[l];
                out = frac[l] * a + out * (1.0f - a);
            }
            const uint8_t d = to_u8(out);
            s_u8[1][q * 3 + 0] = d;
            s_u8[1][q * 3 + 1] = d;
            s_u8[1][q * 3 + 2] = d;
        }
        s_out[ch][q] = out;
    }
    __syncthreads();

    const int nvalid = (int)min((long long)NPIX, npix - pix0);
    if (nvalid == NPIX) {
        // full block: every output run starts 16-byte aligned (pix0 is a multiple of 64)
        const int t = threadIdx.x;
        if (t < NPIX * 3 / 4) {  // float rgb: float4 #t = elements 4t .. 4t+3 of [NPIX][3]
            if (p.out_rgb != nullptr) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = 4 * t + k;
                    v[k] = s_out[e % 3][e / 3];
                }
                reinterpret_cast<float4*>(p.out_rgb + pix0 * 3)[t] = make_float4(v[0], v[1], v[2], v[3]);
            }
        } else if (t >= 64 && t < 64 + NPIX * 3 / 4) {
            if (p.out_depth != nullptr) {
                const int tt = t - 64;
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = s_out[3][(4 * tt + k) / 3];
                reinterpret_cast<float4*>(p.out_depth + pix0 * 3)[tt] = make_float4(v[0], v[1], v[2], v[3]);
            }
        } else if (t >= 128 && t < 128 + NPIX * 3 / 16) {
            if (p.out_rgb_u8 != nullptr)
                reinterpret_cast<uint4*>(p.out_rgb_u8 + pix0 * 3)[t - 128] = reinterpret_cast<const uint4*>(s_u8[0])[t - 128];
        } else if (t >= 160 && t < 160 + NPIX * 3 / 16) {
            if (p.out_depth_u8 != nullptr)
                reinterpret_cast<uint4*>(p.out_depth_u8 + pix0 * 3)[t - 160] = reinterpret_cast<const uint4*>(s_u8[1])[t - 160];
        }
    } else if ((int)threadIdx.x < nvalid * 3) {  // ragged last block: element-wise
        const int e = threadIdx.x, q = e / 3, ch = e - 3 * q;
        if (p.out_rgb != nullptr) p.out_rgb[pix0 * 3 + e] = s_out[ch][q];
        if (p.out_depth != nullptr) p.out_depth[pix0 * 3 + e] = s_out[3][q];
        if (p.out_rgb_u8 != nullptr) p.out_rgb_u8[pix0 * 3 + e] = s_u8[0][e];
        if (p.out_depth_u8 != nullptr) p.out_depth_u8[pix0 * 3 + e] = s_u8[1][e];
    }
    if (p.peer_u8 != nullptr) {
        // the block's NPIX pixels x 3 bytes, contiguous in the gathered buffer: multimem / peer word stores
        const long long dst = p.gather_off + pix0 * 3;  // multiple of 4
        if (nvalid == NPIX) {
            if (threadIdx.x >= 192 && threadIdx.x < 192 + NPIX * 3 / 4) {
                const int w4 = threadIdx.x - 192;
                const unsigned int w = reinterpret_cast<const unsigned int*>(s_u8[0])[w4];
                if (p.mc_u8 != nullptr) {
                    asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(p.mc_u8 + dst + 4 * w4), "r"(w) : "memory");
                } else {
                    for (int k = 0; k < p.n_peers; ++k) *reinterpret_cast<unsigned int*>(p.peer_u8[k] + dst + 4 * w4) = w;
                }
            }
        } else if ((int)threadIdx.x < nvalid * 3) {
            for (int k = 0; k < p.n_peers; ++k) p.peer_u8[k][dst + threadIdx.x] = s_u8[0][threadIdx.x];
        }
    }
}

// the coordinates the fused render kernel samples at (fast chain), for parity accounting
__global__ void __launch_bounds__(256) sphere_coords_fast_kernel(RenderParams p, float* uv) {
    // [B,L,H,W]
    const long long total = (long long)p.B * p.L * p.H * p.W;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % p.W);
    const int i = (int)((idx / p.W) % p.H);
    const int l = (int)((idx / ((long long)p.W * p.H)) % p.L);
    const int b = (int)(idx / ((long long)p.W * p.H * p.L));
    const SphereRay ray = sphere_ray(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
                                     p.pose_rt + b * 16, p.tgt_pos + b * 3);
    const float r = __ldg(p.depths + l);
    float u, v;
    sphere_hit_uv_fast(make_fast_ray(ray), r * r, p.k, u, v);
    uv[2 * idx + 0] = u;
    uv[2 * idx + 1] = v;
}

__global__ void __launch_bounds__(256) sphere_coords_kernel(RenderParams p, float* uv) {
    // [B,L,H,W]
    const long long total = (long long)p.B * p.L * p.H * p.W;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % p.W);
    const int i = (int)((idx / p.W) % p.H);
    const int l = (int)((idx / ((long long)p.W * p.H)) % p.L);
    const int b = (int)(idx / ((long long)p.W * p.H * p.L));
    float u, v;
    sphere_uv(__ldg(p.cos_s + j), __ldg(p.sin_s + j), __ldg(p.cos_t + i), __ldg(p.sin_t + i),
              p.pose_rt + b * 16, p.tgt_pos + b * 3, __ldg(p.depths + l), p.k, u, v);
    uv[2 * idx + 0] = u;
    uv[2 * idx + 1] = v;
}

__global__ void __launch_bounds__(256) project_layers_kernel(RenderParams p, float4* out) {
    // thread per (b, i, j, l), l fastest (coalesced gathers); out [L,B,H,W] float4
    const long long total = (long long)p.B * p.H * p.W * p.L;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int l = (int)(idx % p.L);
    const long long pix = idx / p.L;
    const int j = (int)(pix % p.W);
    const int i = (int)((pix / p.W) % p.H);
    const int b = (int)(pix / ((long long)p.W * p.H));
    out[(size_t)l * p.B * p.H * p.W + pix] = sample_layer(p, b, i, j, l);
}

// ------------------------------------------------------------------------------------------
// stand-alone sampling.resample and projector.over_composite[_depth]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ image, const float* __restrict__ coords, int N, int H, int W, int C,
                int h, int w, float* __restrict__ out) {
    const long long total = (long long)N * h * w;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int n = (int)(idx / ((long long)h * w));
    const Bilinear s = bilinear_setup(__ldg(coords + 2 * idx), __ldg(coords + 2 * idx + 1), W, H);
    const float* base = image + (size_t)n * H * W * C;
    const float* pa = base + ((size_t)s.y0 * W + s.x0) * C;
    const float* pb = base + ((size_t)s.y0 * W + s.x1) * C;
    const float* pc = base + ((size_t)s.y1 * W + s.x0) * C;
    const float* pd = base + ((size_t)s.y1 * W + s.x1) * C;
    float* o = out + (size_t)idx * C;
    for (int c = 0; c < C; ++c) o[c] = blend4(s, __ldg(pa + c), __ldg(pb + c), __ldg(pc + c), __ldg(pd + c));
}

__global__ void __launch_bounds__(256)
over_composite_kernel(const float4* __restrict__ layers, int L, long long npix, int depth_mode,
                      float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= npix) return;
    float4 first = __ldg(layers + idx);
    float r = depth_mode ? 0.f : first.x, g = depth_mode ? 0.f : first.y, b = depth_mode ? 0.f : first.z;
    for (int l = 1; l < L; ++l) {
        const float4 c = __ldg(layers + (size_t)l * npix + idx);
        const float a = c.w;
        const float oma = 1.0f - a;
        if (depth_mode) {
            const float f = (float)((double)l / (double)L);
            r = f * a + r * oma;
            g = r;
            b = r;
        } else {
            r = c.x * a + r * oma;
            g = c.y * a + g * oma;
            b = c.z * a + b * oma;
        }
    }
    out[idx * 3 + 0] = r;
    out[idx * 3 + 1] = g;
    out[idx * 3 + 2] = b;
}

}  // namespace msi

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace msi;

static int fill_psv_params(PsvParams& p, const void* ref, const void* src, int preprocess, const float* poses,
                           const float* baselines, const float* depths, const float* cos_s, const float* sin_s,
                           const float* cos_t, const float* sin_t, int B, int H, int W, int P) {
    MSI_CHECK_ARG(B > 0 && H > 1 && W > 1 && P > 0, "psv: bad shape B=%d H=%d W=%d P=%d", B, H, W, P);
    MSI_CHECK_ARG(poses && baselines && depths && cos_s && sin_s && cos_t && sin_t, "psv: null table/pose pointer");
    p.img[0] = ref;
    p.img[1] = src;
    p.poses = poses;
    p.baselines = baselines;
    p.depths = depths;
    p.cos_s = cos_s;
    p.sin_s = sin_s;
    p.cos_t = cos_t;
    p.sin_t = sin_t;
    p.B = B;
    p.H = H;
    p.W = W;
    p.P = P;
    p.preprocess = preprocess;
    p.out_f32 = nullptr;
    p.out_hi = nullptr;
    p.out_lo = nullptr;
    p.c_stride = 6 * P;
    p.k = make_erp_consts(H, W);
    p.table = nullptr;
    p.table_frames = 0;
    return MSI_OK;
}

extern "C" size_t msi_psv_scratch_bytes(int B, int H, int W) {
    return (size_t)2 * (size_t)B * (size_t)H * (size_t)W * sizeof(float4);
}

extern "C" int msi_psv_build(const void* ref, const void* src, int img_dtype, int preprocess, const float* poses,
                             const float* baselines, const float* depths, const float* cos_s, const float* sin_s,
                             const float* cos_t, const float* sin_t, int B, int H, int W, int P, float* out_f32,
                             void* out_hi, void* out_lo, int c_stride, void* scratch, size_t scratch_bytes,
                             void* stream) {
    PsvParams p;
    int rc = fill_psv_params(p, ref, src, preprocess, poses, baselines, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, P);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(ref && src, "psv: null image pointer");
    MSI_CHECK_ARG(out_f32 || out_hi, "psv: no output requested");
    MSI_CHECK_ARG(img_dtype == MSI_IMG_F32 || img_dtype == MSI_IMG_U8, "psv: bad img_dtype %d", img_dtype);
    if (out_hi) MSI_CHECK_ARG(c_stride >= 6 * P && c_stride % 8 == 0, "psv: c_stride %d must be >= 6P and a multiple of 8", c_stride);
    p.out_f32 = out_f32;
    p.out_hi = reinterpret_cast<__half*>(out_hi);
    p.out_lo = reinterpret_cast<__half*>(out_lo);
    p.c_stride = out_hi ? c_stride : 6 * P;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_hi && c_stride != 6 * P) {
        const size_t bytes = (size_t)B * H * W * c_stride * sizeof(__half);
        MSI_CUDA(cudaMemsetAsync(out_hi, 0, bytes, st));
        if (out_lo) MSI_CUDA(cudaMemsetAsync(out_lo, 0, bytes, st));
    }
    const long long npix = (long long)B * H * W;
    const bool pair = scratch != nullptr && scratch_bytes >= msi_psv_scratch_bytes(B, H, W) && (256 % P == 0) &&
                      npix < (1LL << 31) &&
                      (P % 4 == 0) && p.c_stride == 6 * P && ((uintptr_t)scratch % 16 == 0);
    if (pair) {
        float4* rgbx = reinterpret_cast<float4*>(scratch);
        if (img_dtype == MSI_IMG_F32)
            prep_images_kernel<float><<<ceil_div(2 * npix, 256), 256, 0, st>>>(
                reinterpret_cast<const float*>(ref), reinterpret_cast<const float*>(src), npix, preprocess, rgbx);
        else
            prep_images_kernel<uint8_t><<<ceil_div(2 * npix, 256), 256, 0, st>>>(
                reinterpret_cast<const uint8_t*>(ref), reinterpret_cast<const uint8_t*>(src), npix, preprocess, rgbx);
        MSI_LAUNCH_CHECK();
        psv_build_pair_kernel<<<ceil_div(npix, 256 / P), 256, 0, st>>>(p, rgbx);
        MSI_LAUNCH_CHECK();
        return MSI_OK;
    }
    const long long total = npix * 2 * P;
    const int grid = ceil_div(total, 256);
    if (img_dtype == MSI_IMG_F32)
        psv_build_kernel<float, false><<<grid, 256, 0, st>>>(p);
    else
        psv_build_kernel<uint8_t, false><<<grid, 256, 0, st>>>(p);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" size_t msi_sweep_table_bytes(int frames, int H, int W, int P) {
    if (frames <= 0 || H <= 0 || W <= 0 || P <= 0) return 0;
    return (size_t)frames * (size_t)H * (size_t)W * (size_t)P * sizeof(float4);
}

extern "C" int msi_sweep_table_build(const float* poses, const float* baselines, const float* depths, const float* cos_s,
                                     const float* sin_s, const float* cos_t, const float* sin_t, int frames, int H, int W,
                                     int P, void* table, void* stream) {
    PsvParams p;
    int rc = fill_psv_params(p, nullptr, nullptr, 0, poses, baselines, depths, cos_s, sin_s, cos_t, sin_t, frames, H, W, P);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(table != nullptr && ((uintptr_t)table % 16 == 0), "sweep_table_build: table must be a 16-byte aligned device pointer");
    const long long total = (long long)frames * H * W * P;
    sweep_table_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p, reinterpret_cast<float4*>(table));
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_psv_gather(const void* ref, const void* src, int img_dtype, int preprocess, const void* table,
                              int table_frames, int B, int H, int W, int P, float* out_f32, void* out_hi, void* out_lo,
                              int c_stride, void* scratch, size_t scratch_bytes, void* stream) {
    MSI_CHECK_ARG(B > 0 && H > 1 && W > 1 && P > 0, "psv_gather: bad shape B=%d H=%d W=%d P=%d", B, H, W, P);
    MSI_CHECK_ARG(ref && src, "psv_gather: null image pointer");
    MSI_CHECK_ARG(table != nullptr && ((uintptr_t)table % 16 == 0), "psv_gather: table must be a 16-byte aligned device pointer");
    MSI_CHECK_ARG(table_frames == 1 || table_frames == B, "psv_gather: table_frames=%d must be 1 or B=%d", table_frames, B);
    MSI_CHECK_ARG(out_f32 || out_hi, "psv_gather: no output requested");
    MSI_CHECK_ARG(img_dtype == MSI_IMG_F32 || img_dtype == MSI_IMG_U8, "psv_gather: bad img_dtype %d", img_dtype);
    if (out_hi) MSI_CHECK_ARG(c_stride >= 6 * P && c_stride % 8 == 0, "psv_gather: c_stride %d must be >= 6P and a multiple of 8", c_stride);
    PsvParams p = {};
    p.img[0] = ref;
    p.img[1] = src;
    p.B = B;
    p.H = H;
    p.W = W;
    p.P = P;
    p.preprocess = preprocess;
    p.out_f32 = out_f32;
    p.out_hi = reinterpret_cast<__half*>(out_hi);
    p.out_lo = reinterpret_cast<__half*>(out_lo);
    p.c_stride = out_hi ? c_stride : 6 * P;
    p.table = reinterpret_cast<const float4*>(table);
    p.table_frames = table_frames;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_hi && c_stride != 6 * P) {
        const size_t bytes = (size_t)B * H * W * c_stride * sizeof(__half);
        MSI_CUDA(cudaMemsetAsync(out_hi, 0, bytes, st));
        if (out_lo) MSI_CUDA(cudaMemsetAsync(out_lo, 0, bytes, st));
    }
    const long long npix = (long long)B * H * W;
    const bool pair = scratch != nullptr && scratch_bytes >= msi_psv_scratch_bytes(B, H, W) && (256 % P == 0) &&
                      npix < (1LL << 31) && B <= 65535 && (P % 4 == 0) && p.c_stride == 6 * P && ((uintptr_t)scratch % 16 == 0);
    if (pair) {
        float4* rgbx = reinterpret_cast<float4*>(scratch);
        if (img_dtype == MSI_IMG_F32)
            prep_images_kernel<float><<<ceil_div(2 * npix, 256), 256, 0, st>>>(
                reinterpret_cast<const float*>(ref), reinterpret_cast<const float*>(src), npix, preprocess, rgbx, MSI_ACT_SCALE);
        else
            prep_images_kernel<uint8_t><<<ceil_div(2 * npix, 256), 256, 0, st>>>(
                reinterpret_cast<const uint8_t*>(ref), reinterpret_cast<const uint8_t*>(src), npix, preprocess, rgbx, MSI_ACT_SCALE);
        MSI_LAUNCH_CHECK();
        const dim3 grid((unsigned)ceil_div(ceil_div((long long)H * W, 256 / P), kGatherGroups), (unsigned)B);
        int log2p = 0;
        while ((1 << log2p) < P) ++log2p;  // 256 % P == 0: P is a power of two
        psv_gather_pair_kernel<<<grid, 256, 0, st>>>(p, rgbx, log2p);
        MSI_LAUNCH_CHECK();
        return MSI_OK;
    }
    const long long total = npix * 2 * P;
    const int grid = ceil_div(total, 256);
    if (img_dtype == MSI_IMG_F32)
        psv_build_kernel<float, true><<<grid, 256, 0, st>>>(p);
    else
        psv_build_kernel<uint8_t, true><<<grid, 256, 0, st>>>(p);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_sweep_coords(const float* poses, const float* baselines, const float* depths, const float* cos_s,
                                const float* sin_s, const float* cos_t, const float* sin_t, int B, int H, int W, int P,
                                float* uv, uint8_t* valid, void* stream) {
    PsvParams p;
    int rc = fill_psv_params(p, nullptr, nullptr, 0, poses, baselines, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, P);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(uv != nullptr, "sweep_coords: null uv");
    const long long total = (long long)B * 2 * P * H * W;
    sweep_coords_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, uv, valid);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_rgba_assemble(const float* pred, const float* psv_f32, const void* psv_hi, const void* psv_lo,
                                 int c_stride, int B, int H, int W, int L, float* rgba, float* blend_weights,
                                 float* alphas, void* stream) {
    return msi_rgba_assemble_strided(pred, 2 * L, 2 * L, psv_f32, psv_hi, psv_lo, c_stride, B, H, W, L, MSI_COLOR_BLEND_PSV,
                                     rgba, blend_weights, alphas, nullptr, stream);
}

// The other `which_color_pred` schemes of infer_msi (matryodshka/msi.py:166-273).  pred has n_pred
// channels per pixel: blend_bg [w(L) | alpha(L) | bg rgb(3)], blend_bg_psv [w(L) | alpha(L) | bg_w(L) |
// bg rgb(3)], alpha_only [alpha(L)].  Same op order as the reference's elementwise graph (no FMA).
__global__ void __launch_bounds__(256)
rgba_assemble_ex_kernel(const float* __restrict__ pred, int pred_stride, const float* __restrict__ psv_f32,
                        const __half* __restrict__ psv_hi, const __half* __restrict__ psv_lo, int c_stride,
                        long long npix, int L, int mode, float4* __restrict__ rgba, float* __restrict__ bw_out,
                        float* __restrict__ al_out, float* __restrict__ bgw_out) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= npix * L) return;
    const long long pix = idx / L;
    const int l = (int)(idx % L);
    const float* pp = pred + pix * pred_stride;
    float fg[3], bg[3];
    if (psv_f32 != nullptr) {
        const float* f = psv_f32 + pix * 6 * L + 3 * l;
        const float* g = psv_f32 + pix * 6 * L + 3 * (L + l);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            fg[c] = __ldg(f + c);
            bg[c] = __ldg(g + c);
        }
    } else {
        const size_t f = (size_t)pix * c_stride + 3 * l;
        const size_t g = (size_t)pix * c_stride + 3 * (L + l);
        const float inv = 1.0f / MSI_ACT_SCALE;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            fg[c] = (__half2float(psv_hi[f + c]) + __half2float(psv_lo[f + c])) * inv;
            bg[c] = (__half2float(psv_hi[g + c]) + __half2float(psv_lo[g + c])) * inv;
        }
    }
    float4 o;
    float w = 0.f, al, bgw = 0.f;
    if (mode == MSI_COLOR_ALPHA_ONLY) {  // msi.py:251-267
        al = (__ldg(pp + l) + 1.0f) / 2.0f;
        o.x = fg[0];
        o.y = fg[1];
        o.z = fg[2];
    } else {
        w = (__ldg(pp + l) + 1.0f) / 2.0f;
        al = (__ldg(pp + L + l) + 1.0f) / 2.0f;
        const float omw = 1.0f - w;
        if (mode == MSI_COLOR_BLEND_BG) {  // msi.py:166-190: blend the reference PSV with ONE predicted background
            const float* b3 = pp + 2 * L;
            o.x = w * fg[0] + omw * __ldg(b3 + 0);
            o.y = w * fg[1] + omw * __ldg(b3 + 1);
            o.z = w * fg[2] + omw * __ldg(b3 + 2);
        } else {  // MSI_COLOR_BLEND_BG_PSV, msi.py:209-247: blend both PSVs, then with the predicted background
            bgw = (__ldg(pp + 2 * L + l) + 1.0f) / 2.0f;
            const float ombg = 1.0f - bgw;
            const float* b3 = pp + 3 * L;
            o.x = bgw * (w * fg[0] + omw * bg[0]) + ombg * __ldg(b3 + 0);
            o.y = bgw * (w * fg[1] + omw * bg[1]) + ombg * __ldg(b3 + 1);
            o.z = bgw * (w * fg[2] + omw * bg[2]) + ombg * __ldg(b3 + 2);
        }
    }
    o.w = al;
    rgba[idx] = o;
    if (bw_out != nullptr) bw_out[idx] = w;
    if (al_out != nullptr) al_out[idx] = al;
    if (bgw_out != nullptr) bgw_out[idx] = bgw;
}

extern "C" int msi_rgba_assemble_ex(const float* pred, int n_pred, const float* psv_f32, const void* psv_hi,
                                    const void* psv_lo, int c_stride, int B, int H, int W, int L, int mode, float* rgba,
                                    float* blend_weights, float* alphas, float* bg_blend_weights, void* stream) {
    return msi_rgba_assemble_strided(pred, n_pred, n_pred, psv_f32, psv_hi, psv_lo, c_stride, B, H, W, L, mode, rgba,
                                     blend_weights, alphas, bg_blend_weights, stream);
}

// All colour schemes, with the prediction's pixel stride separate from its channel count: the tensor-core
// net pads its head to a multiple of 64 output channels, and the pipeline reads that buffer in place.
extern "C" int msi_rgba_assemble_strided(const float* pred, int n_pred, int pred_stride, const float* psv_f32,
                                         const void* psv_hi, const void* psv_lo, int c_stride, int B, int H, int W, int L,
                                         int mode, float* rgba, float* blend_weights, float* alphas,
                                         float* bg_blend_weights, void* stream) {
    MSI_CHECK_ARG(B > 0 && H > 0 && W > 0 && L > 0, "rgba_assemble: bad shape");
    MSI_CHECK_ARG(pred && rgba, "rgba_assemble: null pred/rgba");
    MSI_CHECK_ARG(psv_f32 || (psv_hi && psv_lo), "rgba_assemble: need psv_f32 or the hi/lo pair");
    const int need = mode == MSI_COLOR_BLEND_PSV ? 2 * L : mode == MSI_COLOR_BLEND_BG ? 2 * L + 3
                     : mode == MSI_COLOR_BLEND_BG_PSV ? 3 * L + 3 : mode == MSI_COLOR_ALPHA_ONLY ? L : -1;
    MSI_CHECK_ARG(need > 0, "rgba_assemble: unknown mode %d", mode);
    MSI_CHECK_ARG(n_pred == need, "rgba_assemble: mode %d needs %d prediction channels, got %d (L=%d)", mode, need, n_pred, L);
    MSI_CHECK_ARG(pred_stride >= n_pred, "rgba_assemble: pred_stride %d < n_pred %d", pred_stride, n_pred);
    if (!psv_f32) MSI_CHECK_ARG(c_stride >= 6 * L, "rgba_assemble: c_stride %d < 6L", c_stride);
    const long long npix = (long long)B * H * W;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const __half* hi = reinterpret_cast<const __half*>(psv_hi);
    const __half* lo = reinterpret_cast<const __half*>(psv_lo);
    if (mode == MSI_COLOR_BLEND_PSV)
        rgba_assemble_kernel<<<ceil_div(npix * L, 256), 256, 0, st>>>(pred, psv_f32, hi, lo, c_stride, npix, L, pred_stride,
                                                                      reinterpret_cast<float4*>(rgba), blend_weights, alphas);
    else
        rgba_assemble_ex_kernel<<<ceil_div(npix * L, 256), 256, 0, st>>>(pred, pred_stride, psv_f32, hi, lo, c_stride, npix, L,
                                                                         mode, reinterpret_cast<float4*>(rgba), blend_weights,
                                                                         alphas, bg_blend_weights);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

static int fill_render_params(RenderParams& p, const float* rgba, const float* pose_rt, const float* tgt_pos,
                              const float* depths, const float* cos_s, const float* sin_s, const float* cos_t,
                              const float* sin_t, int B, int H, int W, int L) {
    MSI_CHECK_ARG(B > 0 && H > 1 && W > 1 && L > 0, "render: bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
    MSI_CHECK_ARG(pose_rt && tgt_pos && depths && cos_s && sin_s && cos_t && sin_t, "render: null pointer");
    p.rgba = reinterpret_cast<const float4*>(rgba);
    p.pose_rt = pose_rt;
    p.tgt_pos = tgt_pos;
    p.depths = depths;
    p.cos_s = cos_s;
    p.sin_s = sin_s;
    p.cos_t = cos_t;
    p.sin_t = sin_t;
    p.B = B;
    p.H = H;
    p.W = W;
    p.L = L;
    p.out_rgb = nullptr;
    p.out_depth = nullptr;
    p.out_rgb_u8 = nullptr;
    p.out_depth_u8 = nullptr;
    p.k = make_erp_consts(H, W);
    p.ods_mode = 0;
    p.oH = H;
    p.oW = W;
    p.ods_order = 1.0f;
    p.baselines = nullptr;
    p.peer_u8 = nullptr;
    p.n_peers = 0;
    p.mc_u8 = nullptr;
    p.gather_off = 0;
    return MSI_OK;
}

// Launch of the fused render kernel: the 64-pixel fast-chain form by default; MSI_RENDER_V1=1 selects the
// first form (32 pixels per block, strict IEEE chain of sphere_hit_uv) for A/B parity and timing runs.
static int launch_render(const RenderParams& p, long long npix, int L, cudaStream_t st) {
    static int v1 = -1;
    if (v1 < 0) {
        const char* env = getenv("MSI_RENDER_V1");
        v1 = (env && atoi(env) != 0) ? 1 : 0;
    }
    if (v1) {
        const size_t smem = (size_t)(4 * 32 * (L + 1) + L) * sizeof(float);
        MSI_CHECK_ARG(smem <= 200 * 1024, "render: L=%d needs %zu B of shared memory", L, smem);
        static std::atomic<size_t> opted{48 * 1024};
        if (smem > opted.load()) {
            MSI_CUDA(cudaFuncSetAttribute(render_composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            opted.store(smem);
        }
        render_composite_kernel<<<ceil_div(npix, 32), 256, smem, st>>>(p);
        MSI_LAUNCH_CHECK();
        return MSI_OK;
    }
    const size_t smem64 = (size_t)(4 * 64 * (L + 1) + 2 * L) * sizeof(float);
    const size_t smem32 = (size_t)(4 * 32 * (L + 1) + 2 * L) * sizeof(float);
    MSI_CHECK_ARG(smem32 <= 200 * 1024, "render: L=%d needs %zu B of shared memory", L, smem32);
    if (smem64 <= 100 * 1024) {
        static std::atomic<size_t> opted{48 * 1024};
        if (smem64 > opted.load()) {
            MSI_CUDA(cudaFuncSetAttribute(render_composite_v2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
            opted.store(smem64);
        }
        render_composite_v2_kernel<64><<<ceil_div(npix, 64), 256, smem64, st>>>(p);
    } else {
        static std::atomic<size_t> opted{48 * 1024};
        if (smem32 > opted.load()) {
            MSI_CUDA(cudaFuncSetAttribute(render_composite_v2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            opted.store(smem32);
        }
        render_composite_v2_kernel<32><<<ceil_div(npix, 32), 256, smem32, st>>>(p);
    }
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_render_composite(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                                    const float* depths, const float* cos_s, const float* sin_s, const float* cos_t,
                                    const float* sin_t, int B, int H, int W, int L, float* out_rgb, float* out_depth,
                                    uint8_t* out_rgb_u8, uint8_t* out_depth_u8, void* stream) {
    return msi_render_composite_gather(rgba, tgt_pose_rt, tgt_pos, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L, out_rgb,
                                       out_depth, out_rgb_u8, out_depth_u8, nullptr, 0, nullptr, 0, stream);
}

extern "C" int msi_render_composite_gather(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                                           const float* depths, const float* cos_s, const float* sin_s,
                                           const float* cos_t, const float* sin_t, int B, int H, int W, int L,
                                           float* out_rgb, float* out_depth, uint8_t* out_rgb_u8, uint8_t* out_depth_u8,
                                           uint8_t* const* peer_rgb_u8, int n_peers, uint8_t* multicast_rgb_u8,
                                           long long first_frame, void* stream) {
    RenderParams p;
    int rc = fill_render_params(p, rgba, tgt_pose_rt, tgt_pos, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(rgba != nullptr, "render: null rgba");
    MSI_CHECK_ARG(out_rgb || out_depth || out_rgb_u8 || out_depth_u8 || peer_rgb_u8, "render: no output requested");
    if (peer_rgb_u8 != nullptr) {
        MSI_CHECK_ARG(n_peers >= 1 && first_frame >= 0, "render_gather: n_peers=%d first_frame=%lld", n_peers, first_frame);
        p.peer_u8 = peer_rgb_u8;
        p.n_peers = n_peers;
        p.mc_u8 = multicast_rgb_u8;
        p.gather_off = first_frame * (long long)H * W * 3;
        MSI_CHECK_ARG(p.gather_off % 4 == 0, "render_gather: H*W*3 must be a multiple of 4");
    }
    p.out_rgb = out_rgb;
    p.out_depth = out_depth;
    p.out_rgb_u8 = out_rgb_u8;
    p.out_depth_u8 = out_depth_u8;
    return launch_render(p, (long long)B * H * W, L, reinterpret_cast<cudaStream_t>(stream));
}

// MSI.msi_render_ods_view (msi.py:502-525) -> projector.projective_forward_ods (projector.py:101-127) ->
// spherical.intersect_ods (spherical.py:328-365) + over_composite: the MSI seen from one ODS eye.
extern "C" int msi_render_ods(const float* rgba, const float* pose_rt, float order, const float* baselines,
                              const float* depths, const float* cos_s, const float* sin_s, const float* cos_t,
                              const float* sin_t, int B, int H, int W, int L, float* out_rgb, uint8_t* out_rgb_u8,
                              void* stream) {
    RenderParams p;
    int rc = fill_render_params(p, rgba, pose_rt, pose_rt, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(rgba && baselines && (out_rgb || out_rgb_u8), "render_ods: null pointer");
    MSI_CHECK_ARG(order == 1.0f || order == -1.0f, "render_ods: order must be +1 or -1");
    p.out_rgb = out_rgb;
    p.out_rgb_u8 = out_rgb_u8;
    p.ods_mode = 1;
    p.ods_order = order;
    p.baselines = baselines;
    return launch_render(p, (long long)B * H * W, L, reinterpret_cast<cudaStream_t>(stream));
}

// MSI.msi_render_perspective_view (msi.py:475-500) -> projector.projective_forward_sphere_to_perspective
// (projector.py:64-99) -> spherical.intersect_perspective (spherical.py:367-401) + over_composite.
// pose_rt [B,16] is the viewing-window rotation the reference builds itself (projector.py:80-85);
// s_axis [oW] / t_axis [oH] are the uv_grid axes (spherical.py:46-48).
extern "C" int msi_render_perspective(const float* rgba, const float* pose_rt, const float* tgt_pos, const float* depths,
                                      const float* s_axis, const float* t_axis, int B, int H, int W, int L, int oH,
                                      int oW, float* out_rgb, uint8_t* out_rgb_u8, void* stream) {
    RenderParams p;
    int rc = fill_render_params(p, rgba, pose_rt, tgt_pos, depths, s_axis, s_axis, t_axis, t_axis, B, H, W, L);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(rgba && (out_rgb || out_rgb_u8), "render_perspective: null pointer");
    MSI_CHECK_ARG(oH > 0 && oW > 0, "render_perspective: bad output size %dx%d", oH, oW);
    p.out_rgb = out_rgb;
    p.out_rgb_u8 = out_rgb_u8;
    p.ods_mode = 2;
    p.oH = oH;
    p.oW = oW;
    return launch_render(p, (long long)B * oH * oW, L, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int msi_intersect_sphere_coords(const float* tgt_pose_rt, const float* tgt_pos, const float* depths,
                                           const float* cos_s, const float* sin_s, const float* cos_t,
                                           const float* sin_t, int B, int H, int W, int L, float* uv, void* stream) {
    RenderParams p;
    int rc = fill_render_params(p, nullptr, tgt_pose_rt, tgt_pos, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(uv != nullptr, "intersect_sphere_coords: null uv");
    const long long total = (long long)B * L * H * W;
    sphere_coords_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, uv);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

// fast = 1: the coordinates of the fast chain the fused render kernel samples at (parity accounting);
// fast = 0: msi_intersect_sphere_coords.
extern "C" int msi_intersect_sphere_coords_ex(const float* tgt_pose_rt, const float* tgt_pos, const float* depths,
                                              const float* cos_s, const float* sin_s, const float* cos_t,
                                              const float* sin_t, int B, int H, int W, int L, int fast, float* uv,
                                              void* stream) {
    if (!fast) return msi_intersect_sphere_coords(tgt_pose_rt, tgt_pos, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L, uv, stream);
    RenderParams p;
    int rc = fill_render_params(p, nullptr, tgt_pose_rt, tgt_pos, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(uv != nullptr, "intersect_sphere_coords_ex: null uv");
    const long long total = (long long)B * L * H * W;
    sphere_coords_fast_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, uv);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_project_layers(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                                  const float* depths, const float* cos_s, const float* sin_s, const float* cos_t,
                                  const float* sin_t, int B, int H, int W, int L, float* out, void* stream) {
    RenderParams p;
    int rc = fill_render_params(p, rgba, tgt_pose_rt, tgt_pos, depths, cos_s, sin_s, cos_t, sin_t, B, H, W, L);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(rgba && out, "project_layers: null pointer");
    const long long total = (long long)B * H * W * L;
    project_layers_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p, reinterpret_cast<float4*>(out));
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

// ------------------------------------------------------------------------------------------
// Point-wise forms of the geometry/spherical.py functions (the stage-level API of the mirror):
// one kernel, the operation selected by `op`.
// ------------------------------------------------------------------------------------------
namespace msi {
struct PointOpParams {
    int op;
    long long n;          // points per plane
    int planes;           // backproject: number of depths
    const float *a, *b, *c;  // inputs (meaning depends on op)
    const float* pose;    // apply_pose: [planes][16] or one [16]
    int pose_per_plane;
    float order, r;
    float *o0, *o1, *o2;  // outputs
    uint8_t* valid;
    ErpConsts k;
};

__global__ void __launch_bounds__(256) point_op_kernel(PointOpParams p) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = p.n * (p.op <= MSI_OP_APPLY_POSE ? p.planes : 1);
    if (idx >= total) return;
    switch (p.op) {
        case MSI_OP_BACKPROJECT_SPHERICAL: {  // a = S, b = T [n]; c = depth [planes] -> x, y, z [planes][n]
            const long long i = idx % p.n;
            const int pl = (int)(idx / p.n);
            const float S = __ldg(p.a + i), T = __ldg(p.b + i), d = __ldg(p.c + pl);
            const float ct = cosf(T);
            p.o0[idx] = d * (cosf(S) * ct);
            p.o1[idx] = d * sinf(T);
            p.o2[idx] = d * (sinf(S) * ct);
            break;
        }
        case MSI_OP_APPLY_POSE: {  // a, b, c = x, y, z [planes][n]
            const int pl = (int)(idx / p.n);
            const float* m = p.pose + (p.pose_per_plane ? pl * 16 : 0);
            const float x = __ldg(p.a + idx), y = __ldg(p.b + idx), z = __ldg(p.c + idx);
            p.o0[idx] = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
            p.o1[idx] = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
            p.o2[idx] = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
            break;
        }
        case MSI_OP_PROJECT_ODS: {  // a, b, c = x, y, z [n] -> uv [n][2], valid [n]
            float u, v;
            const bool ok = project_ods_point(__ldg(p.a + idx), __ldg(p.b + idx), __ldg(p.c + idx), p.order, p.r, p.k, u, v);
            p.o0[2 * idx] = u;
            p.o0[2 * idx + 1] = v;
            if (p.valid != nullptr) p.valid[idx] = ok ? 1 : 0;
            break;
        }
        case MSI_OP_PROJECT_SPHERICAL: {  // a, b, c = x, y, z [n] -> uv [n][2]
            const float x = __ldg(p.a + idx), y = __ldg(p.b + idx), z = __ldg(p.c + idx);
            const float theta = -atan2f(z, x);
            const float phi = atan2f(y, sqrtf(x * x + z * z));
            float u = theta + p.k.pi;
            u = u - p.k.pi_w;
            u = u / p.k.den_u;
            p.o0[2 * idx] = u * p.k.wm1;
            p.o0[2 * idx + 1] = ((phi + p.k.half_pi - p.k.half_pi_h) / p.k.den_v) * p.k.hm1;
            break;
        }
        default: {  // MSI_OP_THETA_PHI_TO_PIXELS: a = theta, b = phi [n] -> uv [n][2]
            float u = __ldg(p.a + idx) + p.k.pi;
            u = u - p.k.pi_w;
            u = u / p.k.den_u;
            p.o0[2 * idx] = u * p.k.wm1;
            p.o0[2 * idx + 1] = ((__ldg(p.b + idx) + p.k.half_pi - p.k.half_pi_h) / p.k.den_v) * p.k.hm1;
            break;
        }
    }
}
}  // namespace msi

extern "C" int msi_point_op(int op, const float* a, const float* b, const float* c, long long n, int planes,
                            const float* pose, int pose_per_plane, float order, float baseline, int H, int W,
                            float* o0, float* o1, float* o2, uint8_t* valid, void* stream) {
    MSI_CHECK_ARG(op >= MSI_OP_BACKPROJECT_SPHERICAL && op <= MSI_OP_THETA_PHI_TO_PIXELS, "point_op: bad op %d", op);
    MSI_CHECK_ARG(a && b && o0 && n > 0, "point_op: null pointer or n <= 0");
    if (op <= MSI_OP_APPLY_POSE) MSI_CHECK_ARG(c && o1 && o2 && planes > 0, "point_op: op %d needs c, o1, o2, planes", op);
    if (op == MSI_OP_APPLY_POSE) MSI_CHECK_ARG(pose != nullptr, "point_op: apply_pose needs a pose");
    if (op == MSI_OP_PROJECT_ODS || op == MSI_OP_PROJECT_SPHERICAL) MSI_CHECK_ARG(c != nullptr, "point_op: needs z");
    if (op >= MSI_OP_PROJECT_ODS) MSI_CHECK_ARG(H > 1 && W > 1, "point_op: needs the ERP size");
    PointOpParams p;
    p.op = op;
    p.n = n;
    p.planes = planes > 0 ? planes : 1;
    p.a = a;
    p.b = b;
    p.c = c;
    p.pose = pose;
    p.pose_per_plane = pose_per_plane;
    p.order = order;
    p.r = baseline;
    p.o0 = o0;
    p.o1 = o1;
    p.o2 = o2;
    p.valid = valid;
    p.k = make_erp_consts(H > 1 ? H : 2, W > 1 ? W : 2);
    const long long total = n * (op <= MSI_OP_APPLY_POSE ? p.planes : 1);
    point_op_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_resample(const float* image, const float* coords, int N, int H, int W, int C, int h, int w,
                            float* out, void* stream) {
    MSI_CHECK_ARG(image && coords && out, "resample: null pointer");
    MSI_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && h > 0 && w > 0, "resample: bad shape");
    const long long total = (long long)N * h * w;
    resample_kernel<<<ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(image, coords, N, H, W,
                                                                                             C, h, w, out);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_over_composite(const float* layers, int L, int B, int H, int W, int depth_mode, float* out,
                                  void* stream) {
    MSI_CHECK_ARG(layers && out, "over_composite: null pointer");
    MSI_CHECK_ARG(L > 0 && B > 0 && H > 0 && W > 0, "over_composite: bad shape");
    const long long npix = (long long)B * H * W;
    over_composite_kernel<<<ceil_div(npix, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(layers), L, npix, depth_mode, out);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

// ------------------------------------------------------------------------------------------
// High-res plane-streamed re-render (reference driver test.py:284-394): per PSV plane,
//   A  highres_plane_kernel: one high-res PSV plane for both eyes (format_network_input with a
//      single depth, :312-317) + the plane's low-res blend weight / alpha upsampled bilinearly with
//      align_corners (:319-325) + the blend (:327-334) -> one RGBA layer [Hh, Wh] float4;
//   B  highres_composite_kernel: reproject that layer to the target position
//      (msi_render_equirect_view_single, :338) and fold it into the running colour / depth
//      composite (:374-382) in place -- the reference does this step in NumPy on the host.
// ------------------------------------------------------------------------------------------
namespace msi {

struct UpW {
    int lo, hi;
    float lerp;
};
// [TF-1.14 ResizeBilinear, align_corners, legacy scaler]
__device__ __forceinline__ UpW up_weights(int o, float scale, int in_size) {
    UpW w;
    const float in = (float)o * scale;
    const float fl = floorf(in);
    w.lo = max((int)fl, 0);
    w.hi = min((int)ceilf(in), in_size - 1);
    w.lerp = in - fl;
    return w;
}

template <typename T>
__global__ void __launch_bounds__(256)
highres_plane_kernel(PsvParams p, const float* __restrict__ bw, const float* __restrict__ al, int lh, int lw, int L,
                     int plane, float sy, float sx, float4* __restrict__ rgba) {
    // p: B = 1, P = 1 (depths points at the plane's depth), images = the high-res pair
    const long long npix = (long long)p.H * p.W;
    const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
    if (pix >= npix) return;
    const int j = (int)(pix % p.W);
    const int i = (int)(pix / p.W);
    const float cs = __ldg(p.cos_s + j), sn = __ldg(p.sin_s + j), ct = __ldg(p.cos_t + i), st = __ldg(p.sin_t + i);
    const float depth = __ldg(p.depths);
    const float r = __ldg(p.baselines);
    float rgb[2][3];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        float u, v;
        sweep_uv(cs, sn, ct, st, depth, p.poses + e * 16, e == 0 ? 1.0f : -1.0f, r, p.k, u, v);
        const Bilinear s = bilinear_setup(u, v, p.W, p.H);
        const T* img = reinterpret_cast<const T*>(p.img[e]);
        const size_t oa = ((size_t)s.y0 * p.W + s.x0) * 3, ob = ((size_t)s.y0 * p.W + s.x1) * 3;
        const size_t oc = ((size_t)s.y1 * p.W + s.x0) * 3, od = ((size_t)s.y1 * p.W + s.x1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            rgb[e][c] = blend4(s, load_img<T>(img, oa + c, p.preprocess), load_img<T>(img, ob + c, p.preprocess),
                               load_img<T>(img, oc + c, p.preprocess), load_img<T>(img, od + c, p.preprocess));
    }
    const UpW wy = up_weights(i, sy, lh), wx = up_weights(j, sx, lw);
    auto up = [&](const float* m) {
        const float tl = __ldg(m + ((size_t)wy.lo * lw + wx.lo) * L + plane);
        const float tr = __ldg(m + ((size_t)wy.lo * lw + wx.hi) * L + plane);
        const float bl = __ldg(m + ((size_t)wy.hi * lw + wx.lo) * L + plane);
        const float br = __ldg(m + ((size_t)wy.hi * lw + wx.hi) * L + plane);
        const float top = tl + (tr - tl) * wx.lerp;
        const float bottom = bl + (br - bl) * wx.lerp;
        return top + (bottom - top) * wy.lerp;
    };
    const float uw = up(bw), ua = up(al);
    const float omw = 1.0f - uw;
    float4 o;
    o.x = uw * rgb[0][0] + omw * rgb[1][0];
    o.y = uw * rgb[0][1] + omw * rgb[1][1];
    o.z = uw * rgb[0][2] + omw * rgb[1][2];
    o.w = ua;
    rgba[pix] = o;
}

__global__ void __launch_bounds__(256)
highres_composite_kernel(RenderParams p, int plane, int num_planes, float* __restrict__ acc_rgb,
                         float* __restrict__ acc_depth) {
    // p: B = 1, L = 1, rgba = the plane's layer, depths points at the plane's depth
    const long long npix = (long long)p.H * p.W;
    const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
    if (pix >= npix) return;
    const int j = (int)(pix % p.W);
    const int i = (int)(pix / p.W);
    const float4 s = sample_layer(p, 0, i, j, 0);
    float* o = acc_rgb + pix * 3;
    float* d = acc_depth + pix * 3;
    if (plane == 0) {
        o[0] = s.x;
        o[1] = s.y;
        o[2] = s.z;
        d[0] = d[1] = d[2] = 0.0f;
    } else {
        const float oma = 1.0f - s.w;
        o[0] = o[0] * oma + s.x * s.w;
        o[1] = o[1] * oma + s.y * s.w;
        o[2] = o[2] * oma + s.z * s.w;
        const float f = (float)((double)plane / (double)num_planes);
        const float dd = f * s.w + d[0] * oma;
        d[0] = d[1] = d[2] = dd;
    }
}

}  // namespace msi

extern "C" int msi_highres_plane(const void* hres_ref, const void* hres_src, int img_dtype, int preprocess,
                                 const float* poses, const float* baseline, const float* depth, const float* cos_s,
                                 const float* sin_s, const float* cos_t, const float* sin_t, int Hh, int Wh,
                                 const float* blend_weights, const float* alphas, int lh, int lw, int L, int plane,
                                 float* rgba, void* stream) {
    PsvParams p;
    int rc = fill_psv_params(p, hres_ref, hres_src, preprocess, poses, baseline, depth, cos_s, sin_s, cos_t, sin_t, 1, Hh,
                             Wh, 1);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(hres_ref && hres_src && blend_weights && alphas && rgba, "highres_plane: null pointer");
    MSI_CHECK_ARG(lh > 1 && lw > 1 && L > 0 && plane >= 0 && plane < L, "highres_plane: bad low-res shape / plane");
    MSI_CHECK_ARG(img_dtype == MSI_IMG_F32 || img_dtype == MSI_IMG_U8, "highres_plane: bad img_dtype %d", img_dtype);
    const float sy = (float)(lh - 1) / (float)(Hh - 1);
    const float sx = (float)(lw - 1) / (float)(Wh - 1);
    const long long npix = (long long)Hh * Wh;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (img_dtype == MSI_IMG_F32)
        highres_plane_kernel<float><<<ceil_div(npix, 256), 256, 0, st>>>(p, blend_weights, alphas, lh, lw, L, plane, sy, sx,
                                                                         reinterpret_cast<float4*>(rgba));
    else
        highres_plane_kernel<uint8_t><<<ceil_div(npix, 256), 256, 0, st>>>(p, blend_weights, alphas, lh, lw, L, plane, sy,
                                                                           sx, reinterpret_cast<float4*>(rgba));
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

extern "C" int msi_highres_composite(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                                     const float* depth, const float* cos_s, const float* sin_s, const float* cos_t,
                                     const float* sin_t, int Hh, int Wh, int plane, int num_planes, float* acc_rgb,
                                     float* acc_depth, void* stream) {
    RenderParams p;
    int rc = fill_render_params(p, rgba, tgt_pose_rt, tgt_pos, depth, cos_s, sin_s, cos_t, sin_t, 1, Hh, Wh, 1);
    if (rc != MSI_OK) return rc;
    MSI_CHECK_ARG(rgba && acc_rgb && acc_depth, "highres_composite: null pointer");
    MSI_CHECK_ARG(plane >= 0 && plane < num_planes, "highres_composite: plane %d outside [0, %d)", plane, num_planes);
    const long long npix = (long long)Hh * Wh;
    highres_composite_kernel<<<ceil_div(npix, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p, plane, num_planes, acc_rgb, acc_depth);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}
