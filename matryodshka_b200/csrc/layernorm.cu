// LayerNorm over (H, W, C) per frame + ReLU (slim.layer_norm, nets.py:484-485 arg_scope),
// input splitting into the fp16 hi/lo operand format, and the coord-channel bias table.
//
// slim.layer_norm [TF-1.14]: mean/variance over axes 1..3, then tf.nn.batch_normalization with
// eps = 1e-12:  inv = rsqrt(var + eps) * gamma;  y = x * inv + (beta - mean * inv).
// 13.1 M elements per frame after conv1_1, so the reduction is grid-wide: per-block (sum, sumsq)
// partials in float64, a one-block finalize, and a fused normalise + ReLU + fp16 hi/lo split pass.
#include <cuda_fp8.h>

#include "net_internal.cuh"

namespace msi {

static constexpr int kLnChunk = 256 * 4 * 8;  // elements per partial block: 256 threads x 8 float4

int ln_partials_count(long long n_per_sample) { return ceil_div(n_per_sample, kLnChunk); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
ln_partial_kernel(const float* __restrict__ raw, long long n_per_sample, double2* __restrict__ partials,
                  int n_partials) {
    const int b = blockIdx.y;
    const float* x = raw + (size_t)b * n_per_sample;
    const long long start = (long long)blockIdx.x * kLnChunk;
    const long long end = min(start + (long long)kLnChunk, n_per_sample);
    float s = 0.f, q = 0.f;
    // n_per_sample is a multiple of 8 (C % 8 == 0) so float4 access is aligned
    for (long long i = start + threadIdx.x * 4; i < end; i += 256 * 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
        s += (v.x + v.y) + (v.z + v.w);
        q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    __shared__ double sh[2][8];
    double ds = warp_sum((double)s), dq = warp_sum((double)q);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = ds;
        sh[1][threadIdx.x >> 5] = dq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
        for (int w = 0; w < 8; ++w) {
            a += sh[0][w];
            c += sh[1][w];
        }
        partials[(size_t)b * n_partials + blockIdx.x] = make_double2(a, c);
    }
}

__global__ void __launch_bounds__(256)
ln_finalize_kernel(const double2* __restrict__ partials, int n_partials, long long n_per_sample,
                   float2* __restrict__ stats) {
    const int b = blockIdx.x;
    double s = 0, q = 0;
    for (int i = threadIdx.x; i < n_partials; i += 256) {
        const double2 p = partials[(size_t)b * n_partials + i];
        s += p.x;
        q += p.y;
    }
    __shared__ double sh[2][8];
    s = warp_sum(s);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = s;
        sh[1][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
        for (int w = 0; w < 8; ++w) {
            a += sh[0][w];
            c += sh[1][w];
        }
        const double mean = a / (double)n_per_sample;
        double var = c / (double)n_per_sample - mean * mean;
        if (var < 0) var = 0;
        stats[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-12)));
    }
}

// thread = UNROLL x 8 consecutive channels; the raw tensor is read once (streaming load).  UNROLL = 4 for the large
// tensors (more bytes in flight per thread), 2 for the small ones (more blocks: they are latency-bound)
// Offset (in elements) of element `lin` of a dense [rows, W, C] tensor inside the wrap-padded
// [rows, W + 2 x_pad, C] layout, and the offsets of its wrap copies (or -1): column x < x_pad is
// repeated right of the image, column x >= W - x_pad left of it (nets.py:288-295 wrap_pad).
__device__ __forceinline__ void wrap_offsets(long long lin, int W, int C, int x_pad, long long& main_off, long long& copy_off) {
    const long long pix = lin / C;
    const int c = (int)(lin - pix * C);
    const long long row = pix / W;
    const int x = (int)(pix - row * W);
    const long long Wp = W + 2 * x_pad;
    main_off = (row * Wp + x + x_pad) * C + c;
    copy_off = -1;
    if (x < x_pad) copy_off = (row * Wp + x + W + x_pad) * C + c;
    if (x >= W - x_pad) copy_off = (row * Wp + x - W + x_pad) * C + c;
}

// e4m3 pair of two floats (round to nearest, saturating): low byte = a
__device__ __forceinline__ unsigned short to_e4m3x2(float a, float b) {
    return (unsigned short)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}

// byte offset, inside the q8 tensor, of the hi8 run of 8 channels starting at element offset `e` of the fp16-shaped
// tensor [.., C] (e % 8 == 0, C % 64 == 0): chunk k = c / 64 holds [hi8 x 64 | lo8 x 64]; the lo8 run is 64 bytes on
__device__ __forceinline__ size_t q8_offset(size_t e, int C) {
    const size_t pix = e / C;
    const int c = (int)(e - pix * C);
    return pix * (size_t)C * 2 + (size_t)(c >> 6) * 128 + (c & 63);
}

template <bool WRAP, int UNROLL>
__global__ void __launch_bounds__(256)
ln_apply_kernel(const float* raw, long long n_per_sample, int C, const float2* stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ out_hi,
                __half* __restrict__ out_lo, uint8_t* __restrict__ out_q8, int W, int x_pad) {
    // `raw` and `stats` are written by the PREVIOUS kernel, which may still be running when this one starts
    // (programmatic dependent launch): they are deliberately not `const __restrict__` -- that would license the compiler
    // to treat them as read-only for the kernel's lifetime and hoist their loads above griddepcontrol.wait (it did, with
    // `stats`: LayerNorm then ran on the zeroed statistics of a conv kernel that had not finished) -- and the statistics
    // are read with a volatile asm load, which cannot move across the wait.
    // programmatic dependent launch: let the next conv kernel set itself up, then wait for the conv
    // kernel that produced `raw` and `stats` (no-ops when launched without the attribute)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = blockIdx.y;
    float2 st;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(st.x), "=f"(st.y) : "l"(stats + b) : "memory");
    // A thread's 8 channels are the same for all of its groups when a block's 2048 elements are whole pixels
    // (C divides 2048: every ngf that is a power of two): scale and shift are then formed once per thread
    const bool fixed_c = (2048 % C) == 0;
    float inv[8], sh[8];
    auto load_affine = [&](int c0) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            inv[q] = st.y * gm[q];
            sh[q] = bt[q] - st.x * inv[q];
        }
    };
    const int c_fixed = (int)((threadIdx.x * 8u) % (unsigned)C);
    if (fixed_c) load_affine(c_fixed);
    // UNROLL groups of 8 elements per thread, all loads issued before the first use
    long long idx[UNROLL];
    float4 v0[UNROLL], v1[UNROLL];
#pragma unroll
    for (int g = 0; g < UNROLL; ++g) {
        idx[g] = (((long long)blockIdx.x * UNROLL + g) * 256 + threadIdx.x) * 8;
        if (idx[g] < n_per_sample) {
            const size_t off = (size_t)b * n_per_sample + idx[g];
            v0[g] = __ldcs(reinterpret_cast<const float4*>(raw + off));
            v1[g] = __ldcs(reinterpret_cast<const float4*>(raw + off + 4));
        }
    }
#pragma unroll
    for (int g = 0; g < UNROLL; ++g) {
        if (idx[g] >= n_per_sample) continue;
        const size_t off = (size_t)b * n_per_sample + idx[g];
        const int c0 = fixed_c ? c_fixed : (int)(idx[g] % C);
        if (!fixed_c) load_affine(c0);
        // (the q8 tensor: chunk c0 / 64 of the pixel holds [hi8 x 64 | lo8 x 64]; no division: the pixel base is off - c0)
        const size_t q8_in_pix = (size_t)(c0 >> 6) * 128 + (size_t)(c0 & 63);
        const float x[8] = {v0[g].x, v0[g].y, v0[g].z, v0[g].w, v1[g].x, v1[g].y, v1[g].z, v1[g].w};
        __align__(16) __half2 hi[4];
        __align__(16) __half2 lo[4];
        float hf[8], lf[8];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            // (same roundings as the scalar form: y -> fp16 hi, residual y - hi -> fp16 lo; packed converts)
            const float y0 = fmaxf(x[q] * inv[q] + sh[q], 0.f) * MSI_ACT_SCALE;
            const float y1 = fmaxf(x[q + 1] * inv[q + 1] + sh[q + 1], 0.f) * MSI_ACT_SCALE;
            hi[q >> 1] = __floats2half2_rn(y0, y1);
            const float2 h2 = __half22float2(hi[q >> 1]);
            hf[q] = h2.x;
            hf[q + 1] = h2.y;
            lf[q] = y0 - h2.x;
            lf[q + 1] = y1 - h2.y;
            lo[q >> 1] = __floats2half2_rn(lf[q], lf[q + 1]);
        }
        uint2 h8 = make_uint2(0u, 0u), l8 = make_uint2(0u, 0u);
        if (out_q8 != nullptr) {
            const float sh_hi = 1.0f / (float)(1 << kFp8HiShift), sh_lo = (float)(1 << kFp8LoShift);
            h8.x = (unsigned)to_e4m3x2(hf[0] * sh_hi, hf[1] * sh_hi) | ((unsigned)to_e4m3x2(hf[2] * sh_hi, hf[3] * sh_hi) << 16);
            h8.y = (unsigned)to_e4m3x2(hf[4] * sh_hi, hf[5] * sh_hi) | ((unsigned)to_e4m3x2(hf[6] * sh_hi, hf[7] * sh_hi) << 16);
            l8.x = (unsigned)to_e4m3x2(lf[0] * sh_lo, lf[1] * sh_lo) | ((unsigned)to_e4m3x2(lf[2] * sh_lo, lf[3] * sh_lo) << 16);
            l8.y = (unsigned)to_e4m3x2(lf[4] * sh_lo, lf[5] * sh_lo) | ((unsigned)to_e4m3x2(lf[6] * sh_lo, lf[7] * sh_lo) << 16);
        }
        if (!WRAP) {
            *reinterpret_cast<uint4*>(out_hi + off) = *reinterpret_cast<const uint4*>(hi);
            if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + off) = *reinterpret_cast<const uint4*>(lo);
            if (out_q8 != nullptr) {
                uint8_t* q = out_q8 + (off - (size_t)c0) * 2 + q8_in_pix;
                *reinterpret_cast<uint2*>(q) = h8;
                *reinterpret_cast<uint2*>(q + 64) = l8;
            }
        } else {
            long long m, cp;
            wrap_offsets((long long)off, W, C, x_pad, m, cp);  // `off` runs over [B * rows, W, C]
            *reinterpret_cast<uint4*>(out_hi + m) = *reinterpret_cast<const uint4*>(hi);
            if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + m) = *reinterpret_cast<const uint4*>(lo);
            if (out_q8 != nullptr) {
                uint8_t* q = out_q8 + ((size_t)m - (size_t)c0) * 2 + q8_in_pix;
                *reinterpret_cast<uint2*>(q) = h8;
                *reinterpret_cast<uint2*>(q + 64) = l8;
            }
            if (cp >= 0) {
                *reinterpret_cast<uint4*>(out_hi + cp) = *reinterpret_cast<const uint4*>(hi);
                if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + cp) = *reinterpret_cast<const uint4*>(lo);
                if (out_q8 != nullptr) {
                    uint8_t* q = out_q8 + ((size_t)cp - (size_t)c0) * 2 + q8_in_pix;
                    *reinterpret_cast<uint2*>(q) = h8;
                    *reinterpret_cast<uint2*>(q + 64) = l8;
                }
            }
        }
    }
}

int ln_forward(const float* raw, int B, long long n_per_sample, int C, const float* gamma, const float* beta,
               double2* partials, int n_partials, float2* stats, __half* out_hi, __half* out_lo, uint8_t* out_q8,
               bool stats_ready, bool pdl, int W, int x_pad, cudaStream_t st) {
    MSI_CHECK_ARG(C % 8 == 0, "layer_norm: C=%d must be a multiple of 8", C);
    MSI_CHECK_ARG(out_q8 == nullptr || C % 64 == 0, "layer_norm: the e4m3 copy needs C=%d to be a multiple of 64", C);
    if (!stats_ready) {
        // stand-alone statistics (SIMT back end); the tcgen05 conv kernel produces `stats` itself
        const int np = ln_partials_count(n_per_sample);
        MSI_CHECK_ARG(np <= n_partials, "layer_norm: %d partial slots < %d", n_partials, np);
        ln_partial_kernel<<<dim3(np, B), 256, 0, st>>>(raw, n_per_sample, partials, np);
        MSI_LAUNCH_CHECK();
        ln_finalize_kernel<<<B, 256, 0, st>>>(partials, np, n_per_sample, stats);
        MSI_LAUNCH_CHECK();
    }
    // 4 groups per thread where there are enough blocks to fill the GPU several times over, else 2
    const int unroll = (n_per_sample * B >= (4ll << 20)) ? 4 : 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(n_per_sample, 256 * 8 * unroll), B);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && stats_ready && pdl_enabled()) ? 1 : 0;
    auto kern = (x_pad > 0) ? (unroll == 4 ? ln_apply_kernel<true, 4> : ln_apply_kernel<true, 2>)
                            : (unroll == 4 ? ln_apply_kernel<false, 4> : ln_apply_kernel<false, 2>);
    MSI_CUDA(cudaLaunchKernelEx(&cfg, kern, raw, n_per_sample, C, (const float2*)stats, gamma, beta, out_hi, out_lo, out_q8, W,
                                x_pad));
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

// float32 NHWC [npix, C] -> fp16 hi/lo [npix, c_stride] scaled by MSI_ACT_SCALE (pad channels = 0)
__global__ void __launch_bounds__(256)
split_input_kernel(const float* __restrict__ in, long long npix, int C, int c_stride, __half* __restrict__ hi,
                   __half* __restrict__ lo, int W, int x_pad) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= npix * c_stride) return;
    const long long pix = idx / c_stride;
    const int c = (int)(idx % c_stride);
    __half h = __float2half_rn(0.f), l = h;
    if (c < C) split_half(__ldg(in + pix * C + c) * MSI_ACT_SCALE, h, l);
    if (x_pad == 0) {
        hi[idx] = h;
        lo[idx] = l;
    } else {
        long long m, cp;
        wrap_offsets(idx, W, c_stride, x_pad, m, cp);
        hi[m] = h;
        lo[m] = l;
        if (cp >= 0) {
            hi[cp] = h;
            lo[cp] = l;
        }
    }
}

int split_input(const float* in, long long npix, int C, int c_stride, __half* hi, __half* lo, int W, int x_pad,
                cudaStream_t st) {
    split_input_kernel<<<ceil_div(npix * c_stride, 256), 256, 0, st>>>(in, npix, C, c_stride, hi, lo, W, x_pad);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

__global__ void __launch_bounds__(256)
wrap_copy_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, long long total, int W, int c_stride,
                 int x_pad, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;  // 8 halves = 16 bytes
    if (idx >= total) return;
    long long m, cp;
    wrap_offsets(idx, W, c_stride, x_pad, m, cp);
    const uint4 h = *reinterpret_cast<const uint4*>(in_hi + idx), l = *reinterpret_cast<const uint4*>(in_lo + idx);
    *reinterpret_cast<uint4*>(out_hi + m) = h;
    *reinterpret_cast<uint4*>(out_lo + m) = l;
    if (cp >= 0) {
        *reinterpret_cast<uint4*>(out_hi + cp) = h;
        *reinterpret_cast<uint4*>(out_lo + cp) = l;
    }
}

int wrap_copy(const __half* in_hi, const __half* in_lo, long long rows, int W, int c_stride, int x_pad, __half* out_hi,
              __half* out_lo, cudaStream_t st) {
    MSI_CHECK_ARG(c_stride % 8 == 0, "wrap_copy: c_stride=%d must be a multiple of 8", c_stride);
    const long long total = rows * W * c_stride;
    wrap_copy_kernel<<<ceil_div(total / 8, 256), 256, 0, st>>>(in_hi, in_lo, total, W, c_stride, x_pad, out_hi, out_lo);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

__global__ void __launch_bounds__(256)
merge_activation_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, const uint8_t* __restrict__ q8,
                        long long npix, int C, int c_stride, float* __restrict__ out, int W, int x_pad) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= npix * C) return;
    const long long pix = idx / C;
    const int c = (int)(idx % C);
    const long long row = pix / W;
    const size_t opix = (size_t)(row * (W + 2 * x_pad) + (pix - row * W) + x_pad);
    const size_t o = opix * c_stride + c;
    float l;
    if (lo != nullptr) {
        l = __half2float(lo[o]);
    } else {  // the e4m3 residual
        const __nv_fp8_storage_t b = q8[opix * (size_t)c_stride * 2 + (size_t)(c >> 6) * 128 + 64 + (c & 63)];
        l = __half2float(__half(__nv_cvt_fp8_to_halfraw(b, __NV_E4M3))) * (1.0f / (float)(1 << kFp8LoShift));
    }
    out[idx] = (__half2float(hi[o]) + l) * (1.0f / MSI_ACT_SCALE);
}

int merge_activation(const __half* hi, const __half* lo, const uint8_t* q8, long long npix, int C, int c_stride, float* out,
                     int W, int x_pad, cudaStream_t st) {
    MSI_CHECK_ARG(lo != nullptr || q8 != nullptr, "merge_activation: neither residual format is stored");
    merge_activation_kernel<<<ceil_div(npix * C, 256), 256, 0, st>>>(hi, lo, q8, npix, C, c_stride, out, W, x_pad);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

// Coord-channel fold (nets.py:260-270).  The appended channel |sin(lat_row)| depends only on the
// input row, and SAME zero padding zeroes it outside the image, so its contribution to an output
// (ho, wo, n) is a function of ho and of WHICH kw taps are inside the image at wo:
//   cbias[ho][mask][n] = sum_{kh in-bounds} coord[h(ho,kh)] * sum_{kw in mask} Wc[kh][kw][n]
// with mask = 3 bits (kw in-bounds).  The tensor-core GEMM then runs on K = 9*Cin (a multiple of
// 64) and the epilogue adds cbias[ho][mask(wo)][n].
__global__ void __launch_bounds__(256)
coord_bias_kernel(const float* __restrict__ w /*[k,k,cin+1,cout]*/, const float* __restrict__ coord_rows, int k,
                  int cin, int cout, int Hin, int Hout, int stride, int rate, int pad_t, float* __restrict__ cbias) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= Hout * 8 * cout) return;
    const int n = idx % cout;
    const int mask = (idx / cout) & 7;
    const int ho = idx / (cout * 8);
    float acc = 0.f;
    for (int kh = 0; kh < k; ++kh) {
        const int h = ho * stride + kh * rate - pad_t;
        if (h < 0 || h >= Hin) continue;
        float ws = 0.f;
        for (int kw = 0; kw < k; ++kw)
            if (mask & (1 << kw)) ws += w[((size_t)(kh * k + kw) * (cin + 1) + cin) * cout + n];
        acc += coord_rows[h] * ws;
    }
    cbias[idx] = acc;
}

int coord_bias_build(const LayerPlan& L, const float* coord_rows_dev, cudaStream_t st) {
    const int total = L.Hout * 8 * L.cout;
    coord_bias_kernel<<<ceil_div(total, 256), 256, 0, st>>>(L.w_f32, coord_rows_dev, L.k, L.cin_total, L.cout, L.Hin,
                                                           L.Hout, L.stride, L.rate, L.pad_t, L.cbias);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

}  // namespace msi
