// Tensor-core back end of the conv net (MSI_CONV_TCGEN05): implicit-GEMM convolution on the 5th-gen
// tensor cores of sm_100a.  One kernel serves the 3x3 convs (stride 1 / 2, dilation 1 / 2), the 4x4
// stride-2 transposed convs (as four 2x2 output-parity sub-convolutions) and the 1x1 head.
//
//   D[128 pixels, N couts] (fp32, TMEM) += A[128 pixels, 64 ch] (smem) * W[N couts, 64 ch]^T (smem)
//
// * A operand: the activation tensor is NHWC fp16, so one (tap, 64-channel chunk) of the im2col
//   matrix for a BH x BW patch of output pixels is a 4-D TMA box {64 ch, BW, BH, 1} whose start is
//   shifted by the tap offset; out-of-bounds rows/columns are zero-filled by TMA, which IS the SAME
//   padding (including TF's asymmetric 0-before/1-after for stride 2), and `elementStrides` = 2
//   walks every other pixel for the stride-2 layers.  The box lands in shared memory as 128 rows of
//   128 bytes with the 128-byte swizzle = the canonical K-major UMMA layout.  No im2col buffer.
// * W operand: weights pre-packed K-major [class][Cout][K], K = tap * Cin + c, TMA box {64, N}.
// * fp16x3 precision (default): activations and weights are stored as fp16 hi + lo pairs
//   (x = hi + lo to ~22 bits) and every product is hi*hi + lo*hi + hi*lo into the same fp32
//   accumulator -- three MMAs per 16-wide K step -- because a single fp16 (or tf32) pass leaves
//   7e-3 max-abs on the net output, above the path's 1e-3 bar.  MSI_PREC_FP16 issues one.
// * skip connections: the deconvs read their two concatenated sources through two tensor maps.
// * the coord channel of coord_conv2d is folded out of the GEMM into a bias table (layernorm.cu).
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocator), warps 2-5 epilogue
// (tcgen05.ld -> scale/bias/tanh -> float32 NHWC stores).  smem ring of kStages (full/empty mbarriers).
#include <cuda.h>

#include "net_internal.cuh"

namespace msi {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // fp16 elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kThreads = 192;
constexpr int kMaxSmem = 227 * 1024 - 2048;  // dynamic limit: 227 KB minus this kernel's static shared memory

struct TcParams {
    int n_tile;        // UMMA N
    int split;         // 1: fp16x3, 0: fp16
    int stages;
    int stage_bytes;
    int BH, BW;        // output-pixel patch of one M tile, BH * BW = 128
    int tiles_x, tiles_y;
    int Mh, Mw;        // output positions per class
    int in_stride;     // conv stride (TMA traversal stride)
    int out_stride;    // 1, or 2 for deconv classes
    int ncls;
    int nsrc;
    int chunks[2];     // 64-channel chunks per source
    int cs_total;      // packed channels per tap (sum of source channel strides)
    TapList taps[4];
    int Hout, Wout, cout;
    int kind;
    float unscale;     // 1 / (MSI_ACT_SCALE * MSI_WEIGHT_SCALE)
    const float* cbias;  // [Hout][8][cout] or null
    int cb_k, cb_stride, cb_rate, cb_pad_l, cb_Win;  // to derive the kw in-bounds mask of a column
    const float* bias;   // head
    float* out;
};

struct TcPlan {
    TcParams p;
    CUtensorMap a_map[2][2];  // [source][hi/lo]
    CUtensorMap w_map[2];     // hi/lo
    dim3 grid;
    int smem_bytes;
};

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost TMA transaction must fault the kernel, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("msi conv_tc: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128-byte swizzle smem matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major), SBO>>4 [32,46) = 1024 B between
// 8-row groups, version=1 [46,48), layout_type=SWIZZLE_128B(2) [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D=f32 (1<<4), A=B=f16 (0), K-major A and B, N>>3 at 17, M>>4 at 24.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// ---- the kernel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_tcgen05_kernel(const __grid_constant__ CUtensorMap a0_hi, const __grid_constant__ CUtensorMap a0_lo,
                          const __grid_constant__ CUtensorMap a1_hi, const __grid_constant__ CUtensorMap a1_lo,
                          const __grid_constant__ CUtensorMap w_hi, const __grid_constant__ CUtensorMap w_lo,
                          const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tile_x = blockIdx.x % p.tiles_x;
    const int tile_y = blockIdx.x / p.tiles_x;
    const int n0 = blockIdx.y * p.n_tile;
    const int cls = blockIdx.z % p.ncls;
    const int b = blockIdx.z / p.ncls;
    const int ox0 = tile_x * p.BW;
    const int oy0 = tile_y * p.BH;

    const int w_tile_bytes = p.n_tile * kBlockK * 2;
    const int chunks_total = p.chunks[0] + p.chunks[1];
    const TapList& taps = p.taps[cls];
    const int n_iters = taps.n * chunks_total;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&a0_hi);
        prefetch_tmap(&w_hi);
        if (p.split) {
            prefetch_tmap(&a0_lo);
            prefetch_tmap(&w_lo);
        }
        if (p.nsrc == 2) prefetch_tmap(&a1_hi);
    }
    if (warp == 1) {
        // allocate n_tile TMEM columns (power of two >= 32); the same warp frees them
        const uint32_t dst = smem_u32(&tmem_base_smem);
        if (p.n_tile == 64)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(dst) : "memory");
        else if (p.n_tile == 128)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst) : "memory");
        else
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)((kATileBytes + w_tile_bytes) * (p.split ? 2 : 1));
            int it = 0;
            for (int t = 0; t < taps.n; ++t) {
                const int cx = ox0 * p.in_stride + taps.dx[t];
                const int cy = oy0 * p.in_stride + taps.dy[t];
                for (int ch = 0; ch < chunks_total; ++ch, ++it) {
                    const int stage = it % p.stages;
                    const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sa_hi = smem + (size_t)stage * p.stage_bytes;
                    uint8_t* sw_hi = sa_hi + kATileBytes;
                    uint8_t* sa_lo = sw_hi + w_tile_bytes;
                    uint8_t* sw_lo = sa_lo + kATileBytes;
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    const bool second = ch >= p.chunks[0];
                    const int c0 = (second ? ch - p.chunks[0] : ch) * kBlockK;
                    const int kk = t * p.cs_total + ch * kBlockK;  // packed K offset (sources are laid back to back)
                    tma_load_4d(sa_hi, second ? &a1_hi : &a0_hi, &full_bar[stage], c0, cx, cy, b);
                    tma_load_3d(sw_hi, &w_hi, &full_bar[stage], kk, n0, cls);
                    if (p.split) {
                        tma_load_4d(sa_lo, second ? &a1_lo : &a0_lo, &full_bar[stage], c0, cx, cy, b);
                        tma_load_3d(sw_lo, &w_lo, &full_bar[stage], kk, n0, cls);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.n_tile);
            for (int it = 0; it < n_iters; ++it) {
                const int stage = it % p.stages;
                const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&full_bar[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa_hi = smem_u32(smem + (size_t)stage * p.stage_bytes);
                const uint32_t sw_hi = sa_hi + kATileBytes;
                const uint32_t sa_lo = sw_hi + w_tile_bytes;
                const uint32_t sw_lo = sa_lo + kATileBytes;
                const uint64_t da_hi = make_desc(sa_hi), dw_hi = make_desc(sw_hi);
                const uint64_t da_lo = make_desc(sa_lo), dw_lo = make_desc(sw_lo);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per K step inside the swizzle row
                    umma_f16(tmem_base, da_hi + adv, dw_hi + adv, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    if (p.split) {
                        umma_f16(tmem_base, da_lo + adv, dw_hi + adv, idesc, 1u);
                        umma_f16(tmem_base, da_hi + adv, dw_lo + adv, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            }
            umma_commit(&tmem_full_bar);  // accumulator complete
        }
    } else {
        // =============================== epilogue (warps 2..5) ===============================
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;       // M index inside the tile
        const int ly = row / p.BW;
        const int lx = row - ly * p.BW;
        const int my = oy0 + ly, mx = ox0 + lx;    // output position inside the class grid
        const bool valid = (my < p.Mh) && (mx < p.Mw);
        int oy = my, ox = mx;
        if (p.out_stride == 2) {
            oy = my * 2 + (cls >> 1);
            ox = mx * 2 + (cls & 1);
        }
        float* orow = p.out + (((size_t)b * p.Hout + oy) * p.Wout + ox) * p.cout + n0;
        const float* cb = nullptr;
        if (p.cbias != nullptr && valid) {
            int mask = 0;
            for (int kw = 0; kw < p.cb_k; ++kw) {
                const int ix = ox * p.cb_stride + kw * p.cb_rate - p.cb_pad_l;
                if (ix >= 0 && ix < p.cb_Win) mask |= 1 << kw;
            }
            cb = p.cbias + ((size_t)oy * 8 + mask) * p.cout + n0;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int c = 0; c < p.n_tile; c += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v;
                    v.x = __uint_as_float(r[j + 0]) * p.unscale;
                    v.y = __uint_as_float(r[j + 1]) * p.unscale;
                    v.z = __uint_as_float(r[j + 2]) * p.unscale;
                    v.w = __uint_as_float(r[j + 3]) * p.unscale;
                    if (cb != nullptr) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(cb + c + j));
                        v.x += q.x;
                        v.y += q.y;
                        v.z += q.z;
                        v.w += q.w;
                    }
                    if (p.kind == kHead) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
                        v.x = tanhf(v.x + q.x);
                        v.y = tanhf(v.y + q.y);
                        v.z = tanhf(v.z + q.z);
                        v.w = tanhf(v.w + q.w);
                    }
                    *reinterpret_cast<float4*>(orow + c + j) = v;
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.n_tile == 64)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
        else if (p.n_tile == 128)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

// ---- weight packing ----------------------------------------------------------------------------
// w_f32 (TF layout) -> fp16 hi/lo [cls][cout][K], K = tap * cs_total + packed channel, x MSI_WEIGHT_SCALE.
struct PackParams {
    const float* w;
    __half* hi;
    __half* lo;
    int kind, ncls, cout, K, cs_total, cin_total;
    int nsrc, cin[2], cstride[2];
    TapList taps[4];
};

__global__ void __launch_bounds__(256) pack_weights_kernel(PackParams q) {
    const long long total = (long long)q.ncls * q.cout * q.K;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx % q.K);
    const int n = (int)((idx / q.K) % q.cout);
    const int cls = (int)(idx / ((long long)q.K * q.cout));
    const int t = k / q.cs_total;
    int pc = k - t * q.cs_total;  // packed channel
    int c = -1;                   // channel in the concatenated TF tensor
    if (pc < q.cstride[0]) {
        if (pc < q.cin[0]) c = pc;
    } else if (q.nsrc == 2) {
        pc -= q.cstride[0];
        if (pc < q.cin[1]) c = q.cin[0] + pc;
    }
    float v = 0.f;
    if (c >= 0) {
        const int wt = q.taps[cls].wtap[t];
        if (q.kind == kDeconv)
            v = q.w[((size_t)wt * q.cout + n) * q.cin_total + c];
        else if (q.kind == kConv)
            v = q.w[((size_t)wt * (q.cin_total + 1) + c) * q.cout + n];
        else
            v = q.w[(size_t)c * q.cout + n];
    }
    __half h, l;
    split_half(v * MSI_WEIGHT_SCALE, h, l);
    q.hi[idx] = h;
    q.lo[idx] = l;
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int encode_act_map(CUtensorMap* m, const __half* base, int C, int W, int H, int B, int box_w, int box_h, int estride) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return MSI_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d box=%dx%d stride=%d) failed: %d", C, W, H, B,
                  box_w, box_h, estride, (int)r);
        return MSI_ERR_CUDA;
    }
    return MSI_OK;
}

int encode_w_map(CUtensorMap* m, const __half* base, int K, int cout, int ncls, int n_tile) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return MSI_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)cout, (cuuint64_t)ncls};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)n_tile, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights K=%d cout=%d ncls=%d) failed: %d", K, cout, ncls, (int)r);
        return MSI_ERR_CUDA;
    }
    return MSI_OK;
}

// BH x BW = 128 patch that wastes the fewest rows on the Mh x Mw grid (ties: wider).
void pick_tile(int Mh, int Mw, int& BH, int& BW) {
    long long best = -1;
    for (int bw = 128; bw >= 8; bw >>= 1) {
        const int bh = kBlockM / bw;
        const long long cover = (long long)((Mw + bw - 1) / bw) * bw * ((Mh + bh - 1) / bh) * bh;
        if (best < 0 || cover < best) {
            best = cover;
            BW = bw;
            BH = bh;
        }
    }
}

}  // namespace

int conv_tc_plan_create(LayerPlan& L, const ActBuf* srcs, int max_batch, int precision) {
    TcPlan* plan = new TcPlan();
    memset(plan, 0, sizeof(TcPlan));
    TcParams& p = plan->p;
    p.split = (precision == MSI_PREC_FP16X3) ? 1 : 0;
    p.n_tile = (L.cout % 128 == 0) ? 128 : L.cout;
    if (!(p.n_tile == 64 || p.n_tile == 128 || p.n_tile == 256)) {
        delete plan;
        set_error("conv_tc: layer %s has cout=%d; the tcgen05 back end needs 64, or a multiple of 128", L.scope, L.cout);
        return MSI_ERR_UNSUPPORTED;
    }
    p.kind = L.kind;
    p.nsrc = L.nsrc;
    p.cs_total = 0;
    for (int s = 0; s < 2; ++s) {
        p.chunks[s] = 0;
        if (s < L.nsrc) {
            if (srcs[s].c_stride % kBlockK != 0) {
                delete plan;
                set_error("conv_tc: layer %s source %d channel stride %d is not a multiple of 64", L.scope, s, srcs[s].c_stride);
                return MSI_ERR_UNSUPPORTED;
            }
            p.chunks[s] = srcs[s].c_stride / kBlockK;
            p.cs_total += srcs[s].c_stride;
        }
    }
    if (L.kind == kDeconv) {
        p.Mh = L.Hin;
        p.Mw = L.Win;
        p.in_stride = 1;
        p.out_stride = 2;
        p.ncls = 4;
        for (int c = 0; c < 4; ++c) p.taps[c] = deconv_taps(c >> 1, c & 1);
    } else {
        p.Mh = L.Hout;
        p.Mw = L.Wout;
        p.in_stride = L.stride;
        p.out_stride = 1;
        p.ncls = 1;
        p.taps[0] = conv_taps(L);
    }
    pick_tile(p.Mh, p.Mw, p.BH, p.BW);
    p.tiles_x = (p.Mw + p.BW - 1) / p.BW;
    p.tiles_y = (p.Mh + p.BH - 1) / p.BH;
    p.Hout = L.Hout;
    p.Wout = L.Wout;
    p.cout = L.cout;
    p.unscale = 1.0f / (MSI_ACT_SCALE * MSI_WEIGHT_SCALE);
    p.cbias = (L.kind == kConv) ? L.cbias : nullptr;
    p.cb_k = L.k;
    p.cb_stride = L.stride;
    p.cb_rate = L.rate;
    p.cb_pad_l = L.pad_l;
    p.cb_Win = L.Win;
    p.bias = (L.kind == kHead) ? L.bias : nullptr;
    p.stage_bytes = (kATileBytes + p.n_tile * kBlockK * 2) * (p.split ? 2 : 1);
    p.stages = (kMaxSmem - 1024) / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) {
        delete plan;
        set_error("conv_tc: layer %s does not fit 2 pipeline stages", L.scope);
        return MSI_ERR_UNSUPPORTED;
    }
    plan->smem_bytes = p.stages * p.stage_bytes + 1024;
    plan->grid = dim3(p.tiles_x * p.tiles_y, L.cout / p.n_tile, max_batch * p.ncls);

    int rc = MSI_OK;
    for (int s = 0; s < L.nsrc && rc == MSI_OK; ++s) {
        rc = encode_act_map(&plan->a_map[s][0], srcs[s].hi, srcs[s].c_stride, srcs[s].W, srcs[s].H, max_batch, p.BW,
                            p.BH, p.in_stride);
        if (rc == MSI_OK)
            rc = encode_act_map(&plan->a_map[s][1], srcs[s].lo, srcs[s].c_stride, srcs[s].W, srcs[s].H, max_batch,
                                p.BW, p.BH, p.in_stride);
    }
    if (rc == MSI_OK && L.nsrc == 1) {
        plan->a_map[1][0] = plan->a_map[0][0];
        plan->a_map[1][1] = plan->a_map[0][1];
    }
    if (rc == MSI_OK) rc = encode_w_map(&plan->w_map[0], L.w_hi, L.K, L.cout, L.ncls, p.n_tile);
    if (rc == MSI_OK) rc = encode_w_map(&plan->w_map[1], L.w_lo, L.K, L.cout, L.ncls, p.n_tile);
    if (rc != MSI_OK) {
        delete plan;
        return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kMaxSmem);
        if (e != cudaSuccess) {
            delete plan;
            set_error("cudaFuncSetAttribute(conv_igemm_tcgen05_kernel, %d) failed: %s", kMaxSmem, cudaGetErrorString(e));
            return MSI_ERR_CUDA;
        }
        attr_set = true;
    }
    L.tc_plan = plan;
    return MSI_OK;
}

void conv_tc_plan_destroy(LayerPlan& L) {
    if (L.tc_plan) {
        delete reinterpret_cast<TcPlan*>(L.tc_plan);
        L.tc_plan = nullptr;
    }
}

int conv_tc_pack_weights(LayerPlan& L, const ActBuf* srcs, cudaStream_t st) {
    PackParams q;
    q.w = L.w_f32;
    q.hi = L.w_hi;
    q.lo = L.w_lo;
    q.kind = L.kind;
    q.ncls = L.ncls;
    q.cout = L.cout;
    q.K = L.K;
    q.cin_total = L.cin_total;
    q.nsrc = L.nsrc;
    q.cs_total = 0;
    for (int s = 0; s < 2; ++s) {
        q.cin[s] = (s < L.nsrc) ? L.cin[s] : 0;
        q.cstride[s] = (s < L.nsrc) ? srcs[s].c_stride : 0;
        q.cs_total += q.cstride[s];
    }
    if (L.kind == kDeconv)
        for (int c = 0; c < 4; ++c) q.taps[c] = deconv_taps(c >> 1, c & 1);
    else
        q.taps[0] = conv_taps(L);
    const long long total = (long long)L.ncls * L.cout * L.K;
    pack_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(q);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

int conv_tc_forward(const LayerPlan& L, int B, float* out, cudaStream_t st) {
    TcPlan* plan = reinterpret_cast<TcPlan*>(L.tc_plan);
    if (!plan) {
        set_error("conv_tc_forward: layer %s has no plan", L.scope);
        return MSI_ERR_STATE;
    }
    TcParams p = plan->p;
    p.out = out;
    dim3 grid = plan->grid;
    grid.z = B * p.ncls;
    conv_igemm_tcgen05_kernel<<<grid, kThreads, plan->smem_bytes, st>>>(plan->a_map[0][0], plan->a_map[0][1],
                                                                        plan->a_map[1][0], plan->a_map[1][1],
                                                                        plan->w_map[0], plan->w_map[1], p);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

}  // namespace msi
