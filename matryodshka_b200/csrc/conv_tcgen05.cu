// Tensor-core back end of the conv net (MSI_CONV_TCGEN05): implicit-GEMM convolution on the 5th-gen
// tensor cores of sm_100a.  One persistent, warp-specialised kernel serves the 3x3 convs (stride
// 1 / 2, dilation 1 / 2), the 4x4 stride-2 transposed convs (as four 2x2 output-parity
// sub-convolutions) and the 1x1 head.
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
//   (x = hi + lo to ~22 bits) and every product is hi*hi + hi*lo + lo*hi in fp32, because a single
//   fp16 (or tf32) pass leaves 7e-3 max-abs on the net output, above the path's 1e-3 bar.  W_hi and
//   W_lo sit back to back in shared memory, so  A_hi x [W_hi | W_lo]  is ONE MMA of width 2N (A_hi
//   is read from shared memory once, not twice) into accumulator columns [0, 2N), and  A_lo x W_hi
//   accumulates into columns [0, N); the epilogue adds the two halves.  MSI_PREC_FP16 issues one MMA.
// * skip connections: the deconvs read their two concatenated sources through two tensor maps.
// * the coord channel of coord_conv2d is folded out of the GEMM into a bias table (layernorm.cu).
// * LayerNorm statistics (sum, sum of squares per frame) are accumulated in the epilogue, one partial
//   per (frame, CTA, warp), and the last CTA to finish reduces them to (mean, rstd): the global
//   reduction costs no extra pass over the activation and no extra launch.
//
// Persistent CTA = 6 warps, one CTA per SM, static round-robin over output tiles:
//   warp 0   TMA producer (elected lane) -> smem ring of kStages (full / empty mbarriers)
//   warp 1   MMA issuer (elected lane) + TMEM allocator; accumulators double-buffered in TMEM so the
//            epilogue of tile i overlaps the main loop of tile i+1 (tmem_full / tmem_empty mbarriers)
//   warps 2-5 epilogue: tcgen05.ld -> un-scale + coord bias (+ bias, tanh for the head) -> float32
//            NHWC stores + LayerNorm partial sums
#include <cuda.h>
#include <cuda_fp8.h>

#include <algorithm>

#include "net_internal.cuh"

namespace msi {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // fp16 elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kEpiWarps = 8;                        // two warps per TMEM lane quarter, half the columns each
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxDynSmem = 227 * 1024 - 4096;      // 227 KB minus this kernel's static shared memory
// Shared memory the persistent conv CTAs leave FREE on their SM (MSI_CONV_SMEM_RESERVE, bytes; default 0), so that a
// block of another frame's sweep kernel (12.3 KB) can be resident beside them (runtime.MSIFrameLanes).
inline int smem_reserve() {
    static int r = -1;
    if (r < 0) {
        const char* env = getenv("MSI_CONV_SMEM_RESERVE");
        r = env ? atoi(env) : 0;
        if (r < 0 || r > 64 * 1024) r = 0;
    }
    return r;
}
// __launch_bounds__ is given 512 threads although the kernels launch 320 / 352: that caps ptxas at
// 128 registers per thread, which leaves a third of the register file to the kernels of the OTHER
// frames in flight (runtime.MSIFrameLanes) so that they can co-reside with the persistent conv CTA
// (measured with 3 lanes: 778 vs 759 frames/s; one frame at a time is unchanged).
#ifndef MSI_REG_CAP_THREADS
#define MSI_REG_CAP_THREADS 512
#endif
constexpr int kRegCapThreads = MSI_REG_CAP_THREADS;

struct TcParams {
    int stages;
    int BH, BW;          // output-pixel patch of one M tile, BH * BW = 128
    int tiles_x, tiles_y, n_tiles;
    int m_tiles;           // B * tiles_y * tiles_x pixel tiles per column (set per forward)
    int units_per_col;     // ceil(m_tiles / CL)
    int total_units;       // ncls * n_tiles * units_per_col
    int Mh, Mw;          // output positions per class
    int in_stride;       // conv stride (TMA traversal stride)
    int out_stride;      // 1, or 2 for deconv classes
    int ncls;
    int nsrc;
    int chunks[2];       // 64-channel chunks per source
    int cs_total;        // packed channels per tap (sum of source channel strides)
    TapList taps[4];
    int Hout, Wout, cout;
    int kind;
    float unscale;       // 1 / (MSI_ACT_SCALE * MSI_WEIGHT_SCALE)
    const float* cbias;  // [Hout][8][cout] or null
    int cb_k, cb_stride, cb_rate, cb_pad_l, cb_Win;  // to derive the kw in-bounds mask of a column
    const float* bias;   // head
    float* out;
    // LayerNorm statistics
    int do_stats;
    int B;
    int n_partials;      // slots per frame (>= gridDim.x * kEpiWarps)
    double2* partials;   // [B][n_partials], zeroed before the launch
    unsigned int* counter;  // zeroed before the launch
    float2* stats;       // [B] (mean, rstd)
    double n_per_sample;
    // halo kernel (conv_halo_tcgen05_kernel): the A operand of all taps of one 64-channel chunk is ONE
    // halo tile in shared memory; tap (dy, dx) = the same tile read through a row-shifted descriptor
    int orient;          // 0: x is the fast (8-wide) tile dimension, BW = 8, BH = 16; 1: y fast, BH = 8, BW = 16
    int PF, PS;          // halo extent along the fast / slow dimension (pixels)
    int a_rows;          // PF * PS smem rows of 128 bytes per precision half
    int a_slot_bytes;    // [hi halo | lo halo], rounded up to 1024
    int a_stages;
    int T;               // taps per W ring slot
    int w_stages;
    int halo_x0[4], halo_y0[4];  // per class (or parity plane): halo origin in INPUT pixels relative to (ox0, oy0) * in_stride
    // Stride-2 convs: the input splits into 4 parity planes (y & 1, x & 1); input pixel 2 o + d = plane (d mod 2), plane
    // pixel o + floor(d / 2), so every tap is again a row-shifted view of a dense halo tile -- of ITS parity plane.  A
    // 64-channel chunk is therefore `vpar` = 4 halo tiles (TMA element stride 2), each serving the taps of one plane
    // (4 + 2 + 2 + 1 for a 3x3 kernel); the taps are ordered plane by plane (taps[0], and so the packed weights).
    // vpar = 1: one tile per chunk (stride 1, deconv classes).  `sel` below = parity plane (vpar = 4) or class.
    int vpar;
    int v_n[4];                  // taps served by the halo tile of `sel`
    int v_off[4][9];             // smem row offset of every such tap inside the tile
    int w_ntaps;                 // taps per (class, chunk) in the packed weights
    int stage_out;       // 1: the epilogue transposes its tiles through shared memory (coalesced stores)
    int x_off;           // added to every TMA x coordinate: the wrap padding of the source rows (MSI_NET_WRAP)
    // MSI_NET_WRAP deconvs: slim.layer_norm sees the FULL output of the VALID transposed conv over the
    // wrap-padded input, (2H+10) x (2W+10), and only then is it cropped [5:-5] (nets.py:431-436).  The
    // tiles therefore cover the class grid extended by 3 on every side (grid_off = -3): everything is
    // computed and counted in the statistics (positions the reference does not have come out as exact
    // zeros: their inputs are TMA out-of-bounds fill), only the positions inside the crop are stored.
    int grid_off;
    int stat_all;
    long long* trace;    // debugging: CTA 0 records clock64() of its pipeline events here (null = off)
    // debugging (MSI_BEACON=1): every CTA posts the phase each of its roles has reached into host-mapped memory, slot
    // (beacon_seq, blockIdx.x), so that a stalled launch can be read from the host while the GPU is stuck
    long long* beacon;
    int beacon_seq;
    int pair_relinquish;  // MSI_PAIR_RELINQUISH=1 (A/B switch): the pair relinquishes its TMEM permit after the prologue's cluster barrier
    // Head fused with the RGBA assembly (MSI.infer_msi `blend_psv`, msi.py:130-147): the epilogue turns a pixel's
    // L blend weights and L alphas into its L RGBA layers, reading the two PSV eyes of that pixel from the net's
    // own input operand (fp16 hi + lo).  rgba == null: the plain head (tanh -> pred).
    float4* rgba;            // [B,Hout,Wout,L] float4
    const __half* psv_hi;    // [B,H,Wp,psv_cstride], pixel x at column x + psv_xpad
    const __half* psv_lo;
    int psv_cstride, psv_Wp, psv_xpad;
};

struct TcPlan {
    TcParams p;
    int pair;                 // 1: the halo kernel runs as CTA pairs (tcgen05 cta_group::2, W rows split between the CTAs)
    int fp8x;                 // 1: MSI_PREC_FP16_FP8X layer (cross terms in e4m3; the "lo" maps point at the e4m3 copies)
    int halo;                 // 1: conv_halo_tcgen05_kernel (a_map[s][0] = 5-D hi+lo map, w_map[0] = 4-D [kb][hi|lo][cout][64])
    int n_tile, split;
    int cl;                   // cluster size along M: CTAs of a cluster multicast the W tile to each other
    CUtensorMap a_map[2][2];  // [source][hi/lo]
    CUtensorMap w_map[2];     // hi/lo
    int grid;
    int smem_bytes;
};

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute
// may start while its predecessor is still running; pdl_wait() blocks until the predecessor grid has
// completed and its writes are visible, pdl_trigger() lets the successor grid start being scheduled.
// Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking test (try_wait suspends the thread for up to a hardware time limit when the phase
// has not completed -- measured ~8000 clk on B200 -- so it cannot be used to peek at a barrier).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost TMA transaction must fault the kernel, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("msi conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128-byte swizzle smem matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major), SBO>>4 [32,46) = 1024 B between
// 8-row groups, version=1 [46,48), layout_type=SWIZZLE_128B(2) [61,64).
constexpr uint64_t kDescBase = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) { return kDescBase | (uint64_t)((saddr >> 4) & 0x3FFF); }

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
// kind::f8f6f4 with e4m3 operands (the instruction descriptor's format fields are 0 for e4m3 as they are for f16): K = 32
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
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
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
// tanh(x) = 1 - 2 / (1 + e^(2x)); ex2.approx + fast division: abs error < 2e-7 on the whole range
// (e^(2x) -> inf gives 1, -> 0 gives -1), ~6 instructions instead of ~30 for tanhf.
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Work decomposition.  A "column" is one weight tile (output-parity class, N tile); its m_tiles =
// B * tiles_y * tiles_x pixel tiles all multiply the same W.  A work unit is CL consecutive pixel
// tiles of one column, handled by the CL CTAs of a cluster in lockstep so that they can share the W
// tile (each CTA fetches 1/CL of it and multicasts it to the others).  When m_tiles is not a multiple
// of CL the last unit's surplus CTAs recompute the last real tile with their stores masked off.
struct TileCoord {
    int b, cls, n0, ox0, oy0;
    bool dummy;
};
__device__ __forceinline__ TileCoord decode_unit(const TcParams& p, int unit, int rank, int cl, int n_tile) {
    TileCoord t;
    const int col = unit / p.units_per_col;
    int m = (unit - col * p.units_per_col) * cl + rank;
    t.dummy = m >= p.m_tiles;
    if (t.dummy) m = p.m_tiles - 1;
    const int per_frame = p.tiles_x * p.tiles_y;
    t.b = m / per_frame;
    const int r = m - t.b * per_frame;
    const int ty = r / p.tiles_x;
    const int tx = r - ty * p.tiles_x;
    t.cls = col / p.n_tiles;
    t.n0 = (col - t.cls * p.n_tiles) * n_tile;
    t.ox0 = tx * p.BW + p.grid_off;
    t.oy0 = ty * p.BH + p.grid_off;
    return t;
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a W slice into the same smem offset of every CTA in cta_mask; each destination CTA's
// mbarrier (same offset) receives the complete_tx for the bytes written into its own shared memory.
__device__ __forceinline__ void tma_load_3d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  int c2, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
        : "memory");
}
// tcgen05.commit arriving on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(cta_mask)
        : "memory");
}

// ---- CTA-pair (cta_group::2) forms -----------------------------------------------------------------
// In a cluster of two CTAs the shared-window address of the peer differs in bit 24; clearing it
// addresses the leader (even rank) CTA's copy of a barrier (CUTLASS: Sm100MmaPeerBitMask).
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;
// TMA loads issued by BOTH CTAs of the pair; the transaction bytes count on the LEADER's barrier
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kLeaderMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kLeaderMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// one MMA over both CTAs: M = 256 (128 rows from each CTA's A tile), B rows split between the two
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on this offset in BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the leader CTA's barrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kLeaderMask) : "memory");
}
// CTA pair: warp 1 of BOTH CTAs executes the allocation (same warp id, same shared-memory offset for the result: one
// allocation at the same columns of both CTAs' TMEM).  The allocation permit of a cta_group::2 allocation belongs to the
// PAIR: relinquishing it right after the own alloc, as the single-CTA kernels do, is a race -- when the leader's
// relinquish overtakes the peer's alloc (seen once in ~8 runs of bench.py, never in isolation), the peer's alloc blocks
// forever and the pair spins in its prologue, the leader in barrier.cluster.wait (profiles/r2_hang_beacon.log).  The
// pair kernel therefore never relinquishes: a persistent conv CTA owns its SM (no other CTA that allocates TMEM fits
// beside it), so there is nobody to hand the permit to before the CTA exits.
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
                 : "memory");
}
// MSI_PAIR_RELINQUISH=1 only (the form that ran 24 clean bench runs; kept as an A/B switch for scripts/stress_bench.sh):
// relinquish AFTER the cluster barrier that follows the allocation, when both CTAs' allocs have returned
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_pair(int n) {  // M = 256 across the CTA pair
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// ---- debugging beacon (MSI_BEACON) ----
constexpr int kBeaconSeqs = 1024, kBeaconCtas = 160, kBeaconSlots = 8;
// slot 0: thread 0 (1 entered, 2 set-up + cluster sync done, 3 past griddepcontrol.wait, 9 at the final sync, 0 exited);
// 1: MMA warp (1 in its loop, 2 done); 2: epilogue warp 0 (1 in its loop, 2 loop done, 3 statistics done);
// 3: A producer (1 in its loop, 2 done); 4: W producer (1 in its loop, 2 done); 5: units finished by the MMA warp
#define MSI_BEACON(p, k, v)                                                                                              \
    do {                                                                                                                  \
        if ((p).beacon != nullptr)                                                                                        \
            ((volatile long long*)(p).beacon)[((size_t)(p).beacon_seq * kBeaconCtas + blockIdx.x) * kBeaconSlots + (k)] = (v); \
    } while (0)

// ---- debugging trace (MSI_TC_TRACE) ----
constexpr int kTraceRegion = 1024, kTraceRegions = 10;
__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// region 9: [0] = launch counter; launch i records event k (nanoseconds, %globaltimer) at [8 + 16 * i + k]
__device__ __forceinline__ void trace_g(long long* tr, int launch, int k) {
    if (tr != nullptr && launch >= 0 && launch < 60) tr[9 * kTraceRegion + 8 + 16 * launch + k] = globaltimer_ns();
}
__device__ __forceinline__ void trace_ev(long long* tr, int region, int idx) {
    if (tr != nullptr && blockIdx.x == 0 && idx < kTraceRegion) tr[region * kTraceRegion + idx] = clock64();
}

// Epilogue role, shared by both kernels: tcgen05.ld -> un-scale + coord bias (+ bias, tanh for the
// head) -> float32 NHWC stores + LayerNorm partial sums; the last CTA finalises (mean, rstd).
// ew = index of this warp among the kEpiWarps epilogue warps, quarter = its TMEM lane quarter
// (hardware: warp id % 4), (lx, ly) = position of this thread's accumulator row inside the M tile.
// STAGE: a lane owns one accumulator ROW, so a direct float4 store touches 32 different 128-byte
// lines per warp instruction (measured: ~7400 clk of epilogue per 128 x 128 tile, and the burst slows
// the TMA loads of the next tile).  With STAGE each warp writes its 32 rows x 32 columns into a
// private 4 KB shared-memory tile (16-byte chunks XOR-swizzled by the row: conflict-free) and reads it
// back so that 8 lanes cover one row: a warp store then touches 4 full lines.
template <int N_TILE, int SPLIT, int CL, bool STAGE = false, bool PAIR = false>
__device__ __forceinline__ void epilogue_role(const TcParams& p, const int ew, const int quarter, const int lane,
                                              const int lx, const int ly, const int cluster_id, const int n_clusters,
                                              const int cta_rank, const uint32_t tmem_base, const uint32_t tfull0,
                                              const uint32_t tempty0, int* s_is_last, double (*s_red)[kEpiWarps],
                                              const int tr_launch = -1, const uint32_t stage_base = 0) {
    constexpr int kAccCols = SPLIT ? 2 * N_TILE : N_TILE;
    const uint32_t stage_w = stage_base + (uint32_t)ew * 4096u + (uint32_t)lane * 128u;  // this lane's row (write side)
    const int t_sub = lane >> 3, t_chunk = lane & 7;  // read side: row 4k + t_sub, 16-byte chunk t_chunk
    const int c_begin = (ew >> 2) * (N_TILE / 2), c_end = c_begin + N_TILE / 2;  // this warp's columns
    float s_sum = 0.f, s_sq = 0.f;
    int cur_b = -1;
    int local = 0;
    if (ew == 0 && lane == 0) MSI_BEACON(p, 2, 1);
    for (int unit = cluster_id; unit < p.total_units; unit += n_clusters, ++local) {
        const TileCoord tc = decode_unit(p, unit, cta_rank, CL, N_TILE);
        if (ew == 0 && lane == 0) MSI_BEACON(p, 6, (long long)local);
        if (p.do_stats && tc.b != cur_b) {
            if (cur_b >= 0) {
                // this (frame, CTA, warp) slot belongs to this warp alone: plain read-modify-write
                const double ds = warp_sum_d((double)s_sum), dq = warp_sum_d((double)s_sq);
                if (lane == 0) {
                    double2* slot = &p.partials[(size_t)cur_b * p.n_partials + blockIdx.x * kEpiWarps + ew];
                    const double2 old = *slot;
                    *slot = make_double2(old.x + ds, old.y + dq);
                }
                s_sum = 0.f;
                s_sq = 0.f;
            }
            cur_b = tc.b;
        }
        const int my = tc.oy0 + ly, mx = tc.ox0 + lx;  // output position inside the class grid
        const bool valid = (my >= 0) && (mx >= 0) && (my < p.Mh) && (mx < p.Mw) && !tc.dummy;  // stored
        const bool counted = valid || (p.stat_all != 0 && !tc.dummy);                          // computed + in the statistics
        int oy = my, ox = mx;
        if (p.out_stride == 2) {
            oy = my * 2 + (tc.cls >> 1);
            ox = mx * 2 + (tc.cls & 1);
        }
        float* orow = p.out + (((size_t)tc.b * p.Hout + oy) * p.Wout + ox) * p.cout + tc.n0;
        float* tptr[STAGE ? 8 : 1];  // STAGE: where this lane stores chunk t_chunk of tile rows 4k + t_sub (null = masked)
        if (STAGE) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int trow = quarter * 32 + 4 * k + t_sub;
                const int tlx = (p.orient == 0) ? (trow & 7) : (trow >> 3);
                const int tly = (p.orient == 0) ? (trow >> 3) : (trow & 7);
                const int my2 = tc.oy0 + tly, mx2 = tc.ox0 + tlx;
                int oy2 = my2, ox2 = mx2;
                if (p.out_stride == 2) {
                    oy2 = my2 * 2 + (tc.cls >> 1);
                    ox2 = mx2 * 2 + (tc.cls & 1);
                }
                const bool ok = (my2 >= 0) && (mx2 >= 0) && (my2 < p.Mh) && (mx2 < p.Mw) && !tc.dummy;
                tptr[k] = ok ? p.out + (((size_t)tc.b * p.Hout + oy2) * p.Wout + ox2) * p.cout + tc.n0 + t_chunk * 4 : nullptr;
            }
        }
        const float* cb = nullptr;
        if (p.cbias != nullptr && valid) {
            int mask = 0;
            for (int kw = 0; kw < p.cb_k; ++kw) {
                const int ix = ox * p.cb_stride + kw * p.cb_rate - p.cb_pad_l;
                if (ix >= 0 && ix < p.cb_Win) mask |= 1 << kw;
            }
            cb = p.cbias + ((size_t)oy * 8 + mask) * p.cout + tc.n0;
        }
        const int acc = local & 1;
        const uint32_t use = (uint32_t)(local >> 1);
        mbar_wait(tfull0 + 8u * acc, use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.trace != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0 && local < 512) p.trace[6 * 1024 + 2 * local] = clock64();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccCols);
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 32) {
            uint32_t r[32];
            uint32_t r2[SPLIT ? 32 : 1];
            if (!(PAIR && SPLIT)) {  // (a pair in MSI_PREC_FP16_FP8X has ONE accumulator column per cout)
                tmem_ld32(taddr + (uint32_t)c, r);
                if (SPLIT) tmem_ld32(taddr + (uint32_t)(N_TILE + c), r2);
            } else {
                // CTA-pair layout of the 2N accumulator columns (see conv_halo_tcgen05_kernel<.., PAIR>):
                // cout c < N/2: hi.hi + lo.hi in column c, hi.lo in column 3N/2 + c;
                // cout c >= N/2: hi.hi in column N/2 + c, hi.lo + lo.hi in column c
                const bool low = c < N_TILE / 2;
                tmem_ld32(taddr + (uint32_t)(low ? c : N_TILE / 2 + c), r);
                tmem_ld32(taddr + (uint32_t)(low ? 3 * N_TILE / 2 + c : c), r2);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c + 32 >= c_end) {
                // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (PAIR)
                        mbar_arrive_leader(tempty0 + 8u * acc);  // the pair's one MMA thread lives in the leader CTA
                    else
                        mbar_arrive(tempty0 + 8u * acc);
                }
            }
            if (counted) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v;
                    // (explicit roundings: the fused RGBA epilogue of the head repeats this arithmetic bit for bit)
                    if (SPLIT) {
                        v.x = __fmul_rn(__uint_as_float(r[j + 0]) + __uint_as_float(r2[j + 0]), p.unscale);
                        v.y = __fmul_rn(__uint_as_float(r[j + 1]) + __uint_as_float(r2[j + 1]), p.unscale);
                        v.z = __fmul_rn(__uint_as_float(r[j + 2]) + __uint_as_float(r2[j + 2]), p.unscale);
                        v.w = __fmul_rn(__uint_as_float(r[j + 3]) + __uint_as_float(r2[j + 3]), p.unscale);
                    } else {
                        v.x = __fmul_rn(__uint_as_float(r[j + 0]), p.unscale);
                        v.y = __fmul_rn(__uint_as_float(r[j + 1]), p.unscale);
                        v.z = __fmul_rn(__uint_as_float(r[j + 2]), p.unscale);
                        v.w = __fmul_rn(__uint_as_float(r[j + 3]), p.unscale);
                    }
                    if (cb != nullptr) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(cb + c + j));
                        v.x += q.x;
                        v.y += q.y;
                        v.z += q.z;
                        v.w += q.w;
                    }
                    if (p.kind == kHead) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(p.bias + tc.n0 + c + j));
                        v.x = fast_tanh(__fadd_rn(v.x, q.x));
                        v.y = fast_tanh(__fadd_rn(v.y, q.y));
                        v.z = fast_tanh(__fadd_rn(v.z, q.z));
                        v.w = fast_tanh(__fadd_rn(v.w, q.w));
                    }
                    s_sum += (v.x + v.y) + (v.z + v.w);
                    s_sq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                    if (STAGE)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_w + (uint32_t)((((j >> 2) ^ (lane & 7))) << 4)),
                                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                                     : "memory");
                    else if (valid)
                        *reinterpret_cast<float4*>(orow + c + j) = v;
                }
            }
            if (STAGE) {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int lrow = 4 * k + t_sub;
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"(stage_base + (uint32_t)ew * 4096u + (uint32_t)lrow * 128u + (uint32_t)((t_chunk ^ (lrow & 7)) << 4))
                                 : "memory");
                    if (tptr[k] != nullptr) *reinterpret_cast<float4*>(tptr[k] + c) = v;
                }
                __syncwarp();
            }
        }
        if (p.trace != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0 && local < 512) p.trace[6 * 1024 + 2 * local + 1] = clock64();
    }
    if (ew == 0 && lane == 0) trace_g(p.trace, tr_launch, 5);
    if (ew == 0 && lane == 0) MSI_BEACON(p, 2, 2);
    if (p.do_stats) {
        if (cur_b >= 0) {
            const double ds = warp_sum_d((double)s_sum), dq = warp_sum_d((double)s_sq);
            if (lane == 0) {
                double2* slot = &p.partials[(size_t)cur_b * p.n_partials + blockIdx.x * kEpiWarps + ew];
                const double2 old = *slot;
                *slot = make_double2(old.x + ds, old.y + dq);
            }
        }
        // last CTA to finish turns the partials into (mean, rstd) per frame.  Barrier, then ONE
        // thread fences at GPU scope and bumps the counter: the barrier orders the other threads'
        // partial-sum writes before it and the fence is cumulative (the grid-sync pattern).
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        const int et = ew * 32 + lane;  // index inside the epilogue group
        if (et == 0) {
            __threadfence();
            *s_is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        if (*s_is_last) {
            __threadfence();
            const int used = gridDim.x * kEpiWarps;
            for (int b = 0; b < p.B; ++b) {
                double s = 0, q = 0;
                for (int i = et; i < used; i += 32 * kEpiWarps) {
                    const double2 v = __ldcg(&p.partials[(size_t)b * p.n_partials + i]);
                    s += v.x;
                    q += v.y;
                }
                s = warp_sum_d(s);
                q = warp_sum_d(q);
                if (lane == 0) {
                    s_red[0][ew] = s;
                    s_red[1][ew] = q;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (et == 0) {
                    double a = 0, c = 0;
                    for (int w = 0; w < kEpiWarps; ++w) {
                        a += s_red[0][w];
                        c += s_red[1][w];
                    }
                    const double mean = a / p.n_per_sample;
                    double var = c / p.n_per_sample - mean * mean;
                    if (var < 0) var = 0;
                    p.stats[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-12)));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            }
            // trace: the finalising CTA is the last one to finish its epilogue = the end of the grid's work
            if (ew == 0 && lane == 0 && p.trace != nullptr) {
                const long long n = p.trace[9 * kTraceRegion] - 1;  // index of the current launch (CTA 0 bumped it at entry)
                trace_g(p.trace, (int)n, 8);
            }
        }
        if (ew == 0 && lane == 0) trace_g(p.trace, tr_launch, 6);
    }
    if (ew == 0 && lane == 0) MSI_BEACON(p, 2, 3);
}

// Epilogue of the head fused with the RGBA assembly (replaces the `pred` round trip through HBM and the separate
// rgba_assemble launch).  N_TILE = 2L: accumulator columns [0, L) are a pixel's blend weights, [L, 2L) its alphas.
//   1. the two warps of a TMEM lane quarter (ew and ew + 4) read the weight / alpha halves of their 32 pixels,
//      apply bias + tanh and (x + 1) / 2 (msi.py:131-133) and park them in a shared-memory tile, 16-byte chunks
//      XOR-swizzled by the row so that both the row-wise writes and the layer-wise reads are conflict-free;
//   2. named barrier of the pair; then each warp assembles 16 of the 32 pixels with lanes = layer pairs: the PSV
//      taps of a pixel (fg = reference eye, bg = source eye; 3L halves each, hi and lo) are contiguous 4-byte loads
//      across the lanes, rgb = w * fg + (1 - w) * bg without FMA (the arithmetic of rgba_assemble_kernel, bit for
//      bit), and the L float4 of a pixel leave as one contiguous run.
template <int N_TILE>
__device__ __forceinline__ void epilogue_head_rgba(const TcParams& p, const int ew, const int quarter, const int lane,
                                                   const int cluster_id, const int n_clusters, const uint32_t tmem_base,
                                                   const uint32_t tfull0, const uint32_t tempty0, const uint32_t stage_base) {
    constexpr int L = N_TILE / 2;
    constexpr int kAccCols = 2 * N_TILE;
    constexpr uint32_t kRowBytes = N_TILE * 4;
    constexpr int kLanesPerPixel = L / 4;             // a lane owns four layers (4k .. 4k + 3)
    constexpr int kPixPerIter = 32 / kLanesPerPixel;  // L = 32: four pixels per warp step; L = 64: two
    const uint32_t qbase = stage_base + (uint32_t)quarter * 32u * kRowBytes;
    const int half = ew >> 2;  // 0: this warp reads the blend weights out of TMEM, 1: the alphas
    const int sub = lane / kLanesPerPixel, k = lane - sub * kLanesPerPixel;
    const float inv = 1.0f / MSI_ACT_SCALE;
    int local = 0;
    for (int unit = cluster_id; unit < p.total_units; unit += n_clusters, ++local) {
        const TileCoord tc = decode_unit(p, unit, 0, 1, N_TILE);
        const int acc = local & 1;
        const uint32_t use = (uint32_t)(local >> 1);
        mbar_wait(tfull0 + 8u * acc, use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccCols);
#pragma unroll 1
        for (int c0 = 0; c0 < L; c0 += 32) {
            const int c = half * L + c0;
            uint32_t r[32], r2[32];
            tmem_ld32(taddr + (uint32_t)c, r);
            tmem_ld32(taddr + (uint32_t)(N_TILE + c), r2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c0 + 32 >= L) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(p.bias + c + j));
                float4 v;
                v.x = __fadd_rn(fast_tanh(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 0]) + __uint_as_float(r2[j + 0]), p.unscale), q.x)), 1.0f) * 0.5f;
                v.y = __fadd_rn(fast_tanh(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 1]) + __uint_as_float(r2[j + 1]), p.unscale), q.y)), 1.0f) * 0.5f;
                v.z = __fadd_rn(fast_tanh(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 2]) + __uint_as_float(r2[j + 2]), p.unscale), q.z)), 1.0f) * 0.5f;
                v.w = __fadd_rn(fast_tanh(__fadd_rn(__fmul_rn(__uint_as_float(r[j + 3]) + __uint_as_float(r2[j + 3]), p.unscale), q.w)), 1.0f) * 0.5f;
                const int ch = (c + j) >> 2;  // 16-byte chunk of the row
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(qbase + (uint32_t)lane * kRowBytes + (uint32_t)((ch >> 3) << 7) +
                                                                             (uint32_t)(((ch & 7) ^ (lane & 7)) << 4)),
                             "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                             : "memory");
            }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        // Assembly.  A lane owns FOUR layers (4k .. 4k + 3) of one pixel: its PSV taps are 12 consecutive halves = three
        // 8-byte loads per array (fg / bg x hi / lo), its weights and alphas one 16-byte chunk each, its output four
        // float4 = 64 contiguous bytes.  Two steps (kPixPerIter pixels each) are processed together with all 24 loads
        // issued before the first use: the epilogue has only 8 warps per SM, so the bytes in flight per warp decide
        // whether this phase runs at HBM speed (measured: one pixel pair per step with 4-byte loads took 106 us per
        // frame, more than the unfused head + the separate assembly kernel).
#pragma unroll 1
        for (int i = 0; i < 16; i += 2 * kPixPerIter) {
            uint2 wfh[2][3], wfl[2][3], wbh[2][3], wbl[2][3];
            float4 w4[2], a4[2];
            bool valid[2];
            float4* dst[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int rr = half * 16 + i + u * kPixPerIter + sub;  // row of the quarter = pixel
                const int m = quarter * 32 + rr;                       // pixel of the M tile
                const int ly = m / p.BW, lx = m - ly * p.BW;
                const int oy = tc.oy0 + ly, ox = tc.ox0 + lx;
                valid[u] = (oy < p.Mh) && (ox < p.Mw) && !tc.dummy;
                const int chw = k, cha = (L >> 2) + k;  // 16-byte chunks of the row holding w[4k..4k+3] / alpha[4k..4k+3]
                const uint32_t rowaddr = qbase + (uint32_t)rr * kRowBytes;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(w4[u].x), "=f"(w4[u].y), "=f"(w4[u].z), "=f"(w4[u].w)
                             : "r"(rowaddr + (uint32_t)((chw >> 3) << 7) + (uint32_t)(((chw & 7) ^ (rr & 7)) << 4))
                             : "memory");
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(a4[u].x), "=f"(a4[u].y), "=f"(a4[u].z), "=f"(a4[u].w)
                             : "r"(rowaddr + (uint32_t)((cha >> 3) << 7) + (uint32_t)(((cha & 7) ^ (rr & 7)) << 4))
                             : "memory");
                const int oyc = valid[u] ? oy : 0, oxc = valid[u] ? ox : 0;  // masked lanes read pixel (0, 0) and store nothing
                const size_t poff = (((size_t)tc.b * p.Hout + oyc) * p.psv_Wp + oxc + p.psv_xpad) * p.psv_cstride + 12 * k;
                const uint2* fh = reinterpret_cast<const uint2*>(p.psv_hi + poff);
                const uint2* fl = reinterpret_cast<const uint2*>(p.psv_lo + poff);
                const uint2* bh = reinterpret_cast<const uint2*>(p.psv_hi + poff + 3 * L);
                const uint2* bl = reinterpret_cast<const uint2*>(p.psv_lo + poff + 3 * L);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    wfh[u][t] = __ldg(fh + t);
                    wfl[u][t] = __ldg(fl + t);
                    wbh[u][t] = __ldg(bh + t);
                    wbl[u][t] = __ldg(bl + t);
                }
                dst[u] = p.rgba + (((size_t)tc.b * p.Hout + oyc) * p.Wout + oxc) * L + 4 * k;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float fg[12], bg[12];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const unsigned hw[2] = {wfh[u][t].x, wfh[u][t].y}, lw[2] = {wfl[u][t].x, wfl[u][t].y};
                    const unsigned gw[2] = {wbh[u][t].x, wbh[u][t].y}, mw[2] = {wbl[u][t].x, wbl[u][t].y};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&gw[e]));
                        const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&mw[e]));
                        fg[4 * t + 2 * e] = __fmul_rn(__fadd_rn(a.x, b.x), inv);
                        fg[4 * t + 2 * e + 1] = __fmul_rn(__fadd_rn(a.y, b.y), inv);
                        bg[4 * t + 2 * e] = __fmul_rn(__fadd_rn(c.x, d.x), inv);
                        bg[4 * t + 2 * e + 1] = __fmul_rn(__fadd_rn(c.y, d.y), inv);
                    }
                }
                const float wv[4] = {w4[u].x, w4[u].y, w4[u].z, w4[u].w};
                const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
                if (valid[u]) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float om = __fsub_rn(1.0f, wv[j]);
                        float4 o;
                        o.x = __fadd_rn(__fmul_rn(wv[j], fg[3 * j + 0]), __fmul_rn(om, bg[3 * j + 0]));
                        o.y = __fadd_rn(__fmul_rn(wv[j], fg[3 * j + 1]), __fmul_rn(om, bg[3 * j + 1]));
                        o.z = __fadd_rn(__fmul_rn(wv[j], fg[3 * j + 2]), __fmul_rn(om, bg[3 * j + 2]));
                        o.w = av[j];
                        dst[u][j] = o;
                    }
                }
            }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");  // the tile may be overwritten
    }
}

// ---- the kernel ------------------------------------------------------------------------------
template <int N_TILE, int SPLIT, int CL>
__global__ void __launch_bounds__(kRegCapThreads, 1)
conv_igemm_tcgen05_kernel(const __grid_constant__ CUtensorMap a0_hi, const __grid_constant__ CUtensorMap a0_lo,
                          const __grid_constant__ CUtensorMap a1_hi, const __grid_constant__ CUtensorMap a1_lo,
                          const __grid_constant__ CUtensorMap w_hi, const __grid_constant__ CUtensorMap w_lo,
                          const __grid_constant__ TcParams p) {
    // CL = cluster size along M (1 or 2).  The W tensor maps have box {64, N_TILE / CL}.
    constexpr int kWTileBytes = N_TILE * kBlockK * 2;
    constexpr int kWSliceRows = N_TILE / CL;
    constexpr int kWSliceBytes = kWSliceRows * kBlockK * 2;
    constexpr uint16_t kCtaMask = (uint16_t)((1u << CL) - 1u);
    const int cta_rank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;
    const int cluster_id = blockIdx.x / CL;
    const int n_clusters = gridDim.x / CL;
    // SPLIT: 0 = one fp16 MMA; 1 = fp16x3 (A_hi x [W_hi | W_lo] and A_lo x W_hi); 2 = MSI_PREC_FP16_FP8X (A_hi x W_hi in
    // fp16, [A_hi8 | A_lo8] x [W_lo8 | W_hi8] in e4m3 into the SAME accumulator columns: the "lo" tiles hold the e4m3 copies)
    constexpr int kStageBytes = (kATileBytes + kWTileBytes) * (SPLIT ? 2 : 1);
    constexpr int kAccCols = (SPLIT == 1) ? 2 * N_TILE : N_TILE;  // TMEM columns of one accumulator
    constexpr int kTmemCols = 2 * kAccCols;                       // double-buffered
    constexpr uint32_t kTxBytes = (uint32_t)kStageBytes;
    // stage layout: [A_hi][A_lo][W_hi][W_lo] (SPLIT) or [A][W]
    constexpr int kOffALo = kATileBytes;
    constexpr int kOffWHi = SPLIT ? 2 * kATileBytes : kATileBytes;
    constexpr int kOffWLo = kOffWHi + kWTileBytes;

    extern __shared__ uint8_t smem_raw[];
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 1);
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int s_is_last;
    __shared__ double s_red[2][kEpiWarps];
    __shared__ int s_dx[4][9], s_dy[4][9];
    __shared__ int s_ntaps[4];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int chunks_total = p.chunks[0] + p.chunks[1];

    if (threadIdx.x < 36) {
        const int c = threadIdx.x / 9, t = threadIdx.x % 9;
        s_dx[c][t] = p.taps[c].dx[t];
        s_dy[c][t] = p.taps[c].dy[t];
        if (t == 0) s_ntaps[c] = p.taps[c].n;
    }
    // 32-bit shared addresses of the barriers, computed ONCE: converting a __shared__ pointer inside the
    // issue loops costs an S2UR + address rebuild per use, and those loops are latency-critical
    // (measured: the barrier ping-pong alone was ~500 cycles per K iteration before this).
    const uint32_t full0 = smem_u32(&full_bar[0]);
    const uint32_t empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]);
    const uint32_t tempty0 = smem_u32(&tmem_empty_bar[0]);
    if (threadIdx.x == 64) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, CL);  // every CTA of the cluster must have consumed the slot
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8u * a, 1);
            mbar_init(tempty0 + 8u * a, kEpiWarps);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&a0_hi);
        prefetch_tmap(&w_hi);
        if (SPLIT) {
            prefetch_tmap(&a0_lo);
            prefetch_tmap(&w_lo);
        }
        if (p.nsrc == 2) {
            prefetch_tmap(&a1_hi);
            if (SPLIT) prefetch_tmap(&a1_lo);
        }
    }
    if (warp == 1) tmem_alloc<kTmemCols>(&tmem_base_smem);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast targets them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 2);
    pdl_trigger();
    pdl_wait();  // everything above overlapped the previous kernel's tail; its output is read below
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 3);

    if (warp == 0) {
        // =============================== TMA producer ===============================
        const bool leader = elect_one();
        // loop state lives in registers; nothing in the inner loop re-reads kernel parameters
        const int chunks0 = p.chunks[0], cs_total = p.cs_total, in_stride = p.in_stride;
        const int total_units = p.total_units;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t sa = smem_base, full_a = full0, empty_a = empty0;
        for (int unit = cluster_id; unit < total_units; unit += n_clusters) {
            const TileCoord tc = decode_unit(p, unit, cta_rank, CL, N_TILE);
            const int ntaps = s_ntaps[tc.cls];
            const int bx = tc.ox0 * in_stride, by = tc.oy0 * in_stride;
            // W slice this CTA fetches: all N rows (CL == 1) or rows [rank * N/CL, (rank+1) * N/CL)
            const uint32_t so = (CL == 1) ? 0u : (uint32_t)(cta_rank * kWSliceBytes);
            const int n_row = tc.n0 + ((CL == 1) ? 0 : cta_rank * kWSliceRows);
            for (int t = 0; t < ntaps; ++t) {
                const int cx = bx + s_dx[tc.cls][t] + p.x_off;
                const int cy = by + s_dy[tc.cls][t];
                int kk = t * cs_total;
                for (int ch = 0; ch < chunks_total; ++ch, kk += kBlockK) {
                    mbar_wait(empty_a, phase ^ 1u);
                    if (leader) {
                        mbar_expect_tx(full_a, kTxBytes);
                        const bool second = ch >= chunks0;
                        const int c0 = (second ? ch - chunks0 : ch) * kBlockK;
                        tma_load_4d(sa, second ? &a1_hi : &a0_hi, full_a, c0, cx, cy, tc.b);
                        if (SPLIT) tma_load_4d(sa + kOffALo, second ? &a1_lo : &a0_lo, full_a, c0, cx, cy, tc.b);
                        if (CL == 1) {
                            tma_load_3d(sa + kOffWHi, &w_hi, full_a, kk, n_row, tc.cls);
                            if (SPLIT) tma_load_3d(sa + kOffWLo, &w_lo, full_a, kk, n_row, tc.cls);
                        } else {
                            tma_load_3d_mcast(sa + kOffWHi + so, &w_hi, full_a, kk, n_row, tc.cls, kCtaMask);
                            if (SPLIT) tma_load_3d_mcast(sa + kOffWLo + so, &w_lo, full_a, kk, n_row, tc.cls, kCtaMask);
                        }
                    }
                    __syncwarp();
                    sa += kStageBytes;
                    full_a += 8;
                    empty_a += 8;
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1u;
                        sa = smem_base;
                        full_a = full0;
                        empty_a = empty0;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        const bool leader = elect_one();
        constexpr uint32_t idesc_wide = make_idesc(SPLIT == 1 ? 2 * N_TILE : N_TILE);
        constexpr uint32_t idesc_n = make_idesc(N_TILE);
        constexpr uint64_t kDescWLo = (uint64_t)(kOffWLo >> 4);
        const int total_units = p.total_units, units_per_col = p.units_per_col, n_tiles = p.n_tiles;
        int stage = 0;
        uint32_t phase = 0;
        // descriptor of the current stage's A_hi tile, advanced incrementally (the address field is
        // (smem address >> 4); every stage base is 1024-byte aligned and below 256 KB)
        const uint64_t desc0 = make_desc(smem_base);
        constexpr uint64_t kDescStage = (uint64_t)(kStageBytes >> 4);
        constexpr uint64_t kDescWHi = (uint64_t)(kOffWHi >> 4);
        constexpr uint64_t kDescALo = (uint64_t)(kOffALo >> 4);
        uint64_t da = desc0;
        uint32_t full_a = full0, empty_a = empty0;
        int local = 0;
        for (int unit = cluster_id; unit < total_units; unit += n_clusters, ++local) {
            const int cls = (unit / units_per_col) / n_tiles;
            const int n_iters = s_ntaps[cls] * chunks_total;
            const int acc = local & 1;
            const uint32_t use = (uint32_t)(local >> 1);
            mbar_wait(tempty0 + 8u * acc, (use & 1u) ^ 1u);  // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
            const uint32_t tfull_a = tfull0 + 8u * acc;
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(full_a, phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (leader) {
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);  // 32 bytes per K step inside the swizzle row (>> 4)
                        // A_hi x [W_hi | W_lo] -> columns [0, 2N)   (fp16 mode: A x W -> [0, N))
                        umma_f16(d_tmem, da + adv, da + kDescWHi + adv, idesc_wide, (it > 0 || k > 0) ? 1u : 0u);
                        // A_lo x W_hi -> accumulates into columns [0, N)
                        if (SPLIT == 1) umma_f16(d_tmem, da + kDescALo + adv, da + kDescWHi + adv, idesc_n, 1u);
                    }
                    if (SPLIT == 2) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {  // 128 e4m3 per row = 4 K steps of 32: [hi8 | lo8] x [w_lo8 | w_hi8]
                            const uint64_t adv = (uint64_t)(k * 2);
                            umma_f8(d_tmem, da + kDescALo + adv, da + kDescWLo + adv, idesc_n, 1u);
                        }
                    }
                    // free the smem slot when these MMAs retire -- in every CTA of the cluster, because
                    // the peers' next multicast W slices land in this CTA's slot too
                    if (CL == 1)
                        umma_commit(empty_a);
                    else
                        umma_commit_mcast(empty_a, kCtaMask);
                    if (it == n_iters - 1) umma_commit(tfull_a);  // accumulator complete
                }
                __syncwarp();
                da += kDescStage;
                full_a += 8;
                empty_a += 8;
                if (++stage == stages) {
                    stage = 0;
                    phase ^= 1u;
                    da = desc0;
                    full_a = full0;
                    empty_a = empty0;
                }
            }
        }
    } else {
        // =============================== epilogue (warps 2..9) ===============================
        const int row = (warp & 3) * 32 + lane;  // M index inside the tile
        const int ly = row / p.BW;
        if (SPLIT == 1 && CL == 1 && p.rgba != nullptr)
            epilogue_head_rgba<N_TILE>(p, warp - 2, warp & 3, lane, cluster_id, n_clusters, tmem_base, tfull0, tempty0,
                                       smem_base + (uint32_t)(stages * kStageBytes));
        else
            epilogue_role<N_TILE, SPLIT == 1, CL>(p, warp - 2, warp & 3, lane, row - ly * p.BW, ly, cluster_id, n_clusters, cta_rank,
                                                  tmem_base, tfull0, tempty0, &s_is_last, s_red);
    }

    if (threadIdx.x == 0) MSI_BEACON(p, 0, 9);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still write its smem / barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc<kTmemCols>(tmem_base);
    }
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 0);
}

// ---- the halo kernel ---------------------------------------------------------------------------
// Same GEMM, different operand delivery.  Measured on B200 (scripts/probe_umma.cu, profiles/): one SM
// ingests at most ~55-65 B/clk through TMA, and the kernel above re-fetches the activation tile for
// every tap (9x for a 3x3 conv): 64 KB per 768 MMA-clocks = 85 B/clk, so it is bound by operand
// delivery, not by the tensor pipe.  Here the producer loads, per 64-channel chunk, ONE halo tile
// {64 ch, PF, PS} x {hi, lo} (a single 5-D TMA box) and the MMA warp reads all taps out of it: a
// K-major SWIZZLE_128B descriptor may start at any 128-byte row and use any stride between 8-row
// groups (probe A), so tap (dy, dx) is the halo tile's descriptor advanced by (dy * PF + dx) rows
// with SBO = PF rows.  That needs one 8-pixel group per tile row: the M tile is 16 x 8 pixels (or
// 8 x 16 through a tensor map with H and W swapped).  The weights of T taps arrive as one 4-D box
// [tap][hi|lo][N][64] from a K-block-major packing.  Ingest drops to 25-45 B/clk.
//   warp 0  A producer (halo ring, a_stages slots)      warp 2  W producer (w_stages slots of T taps)
//   warp 1  MMA issuer + TMEM allocator                 warps 3-10  epilogue (shared with the kernel above)
constexpr int kHaloThreads = 96 + 32 * kEpiWarps;
// FP8X (MSI_PREC_FP16_FP8X): the "lo" halves of the halo tile and of the weight slot hold the e4m3 copies
// ([hi8 | lo8] per activation row, [w_lo8 | w_hi8] per weight row); per tap and chunk the MMA warp issues four fp16 MMAs
// (K = 16: hi x hi) and four e4m3 MMAs (K = 32: both cross terms) of width N into ONE set of N accumulator columns --
// 8 N-wide MMAs where fp16x3 issues 4 of width 2N and 4 of width N.
template <int N_TILE, int T, bool PAIR, bool FP8X = false>
__global__ void __launch_bounds__(kRegCapThreads, 1)
conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap a0, const __grid_constant__ CUtensorMap a1,
                         const __grid_constant__ CUtensorMap wmap, const __grid_constant__ TcParams p) {
    constexpr int kAccCols = FP8X ? N_TILE : 2 * N_TILE;
    constexpr int kTmemCols = 2 * kAccCols;
    // PAIR: two CTAs of a cluster share every weight tile (tcgen05 cta_group::2, M = 256 = both CTAs'
    // pixel tiles): each holds HALF of the rows, arranged [W_hi rows r*N/2.. | W_lo rows (1-r)*N/2..] for
    // rank r, so that the wide MMA (N = 2N_TILE: first half of the columns from the leader's rows,
    // second half from the peer's) and the narrow one (A_lo x the first N/2 rows of each CTA = W_hi)
    // put the three partial products of a cout into two accumulator columns (see epilogue_role).
    // Halves the weight bytes per SM: a deeper W ring in the same shared memory.
    constexpr int kWTapBytes = (PAIR ? 1 : 2) * N_TILE * kBlockK * 2;  // this CTA's W rows of one tap
    const int cta_rank = PAIR ? (int)(blockIdx.x & 1u) : 0;
    const int cluster_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const bool is_leader = cta_rank == 0;

    extern __shared__ uint8_t smem_raw[];
    __shared__ int s_launch;  // trace only
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 1);
    if (p.trace != nullptr && threadIdx.x == 0) {
        s_launch = (blockIdx.x == 0) ? (int)atomicAdd((unsigned long long*)&p.trace[9 * kTraceRegion], 1ull) : -1;
        trace_g(p.trace, s_launch, 0);
    }
    __shared__ __align__(8) uint64_t a_full_bar[4];
    __shared__ __align__(8) uint64_t a_empty_bar[4];
    __shared__ __align__(8) uint64_t w_full_bar[8];
    __shared__ __align__(8) uint64_t w_empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int s_is_last;
    __shared__ double s_red[2][kEpiWarps];
    __shared__ int s_off[4][9];  // smem row offset of every tap inside the halo tile
    __shared__ int s_ntaps[4];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int chunks_total = p.chunks[0] + p.chunks[1];
    const int a_stages = p.a_stages, w_stages = p.w_stages;
    const uint32_t a_slot_bytes = (uint32_t)p.a_slot_bytes;
    const uint32_t w_slot_bytes = (uint32_t)(T * kWTapBytes);
    const uint32_t w_ring = smem_base + (uint32_t)a_stages * a_slot_bytes;

    if (threadIdx.x < 36) {
        const int c = threadIdx.x / 9, t = threadIdx.x % 9;
        s_off[c][t] = p.v_off[c][t];
        if (t == 0) s_ntaps[c] = p.v_n[c];
    }
    const int vpar = p.vpar;
    const uint32_t afull0 = smem_u32(&a_full_bar[0]);
    const uint32_t aempty0 = smem_u32(&a_empty_bar[0]);
    const uint32_t wfull0 = smem_u32(&w_full_bar[0]);
    const uint32_t wempty0 = smem_u32(&w_empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]);
    const uint32_t tempty0 = smem_u32(&tmem_empty_bar[0]);
    if (threadIdx.x == 96) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(afull0 + 8u * s, 1);
            mbar_init(aempty0 + 8u * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8u * s, 1);
            mbar_init(tempty0 + 8u * s, PAIR ? 2 * kEpiWarps : kEpiWarps);  // PAIR: both CTAs' epilogues arrive on the leader
        }
        for (int s = 0; s < 8; ++s) {
            mbar_init(wfull0 + 8u * s, 1);
            mbar_init(wempty0 + 8u * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&a0);
        if (p.nsrc == 2) prefetch_tmap(&a1);
    }
    if (warp == 2 && lane == 0) prefetch_tmap(&wmap);
    if (PAIR) {
        // both CTAs of the pair are running and in step before either issues the pair's TMEM allocation (a cta_group::2
        // allocation is executed by one warp of EACH CTA; the Blackwell guide: "sync both CTAs before TMEM alloc")
        __syncthreads();
        cluster_sync_all();
    }
    if (warp == 1) {
        if (lane == 0) MSI_BEACON(p, 1, 10);
        if (PAIR)
            tmem_alloc_pair<kTmemCols>(&tmem_base_smem);
        else
            tmem_alloc<kTmemCols>(&tmem_base_smem);
        if (lane == 0) MSI_BEACON(p, 1, 11);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 12);
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything targets them
    if (PAIR && p.pair_relinquish != 0 && warp == 1) tmem_relinquish_pair();  // (A/B switch; default: never, see tmem_alloc_pair)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    const int n_ctas = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;  // stride of the unit loops (clusters)
    if (threadIdx.x == 0) trace_ev(p.trace, 7, 0);
    const int tr_launch = (p.trace != nullptr) ? s_launch : -1;
    if (threadIdx.x == 0) trace_g(p.trace, tr_launch, 1);
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 2);
    // Programmatic dependent launch: the set-up above and the first weight loads below overlap the
    // tail of the previous kernel (the weights do not depend on it); every other role waits here.
    pdl_trigger();
    if (warp != 2) pdl_wait();
    if (threadIdx.x == 0) trace_g(p.trace, tr_launch, 2);
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 3);

    if (warp == 0) {
        // =============================== A (halo) producer ===============================
        if (lane == 0) {
            const int chunks0 = p.chunks[0];
            const uint32_t a_tx = (uint32_t)(2 * p.a_rows * 128);
            int tr_i = 0;
            int stage = 0;
            uint32_t phase = 0;
            MSI_BEACON(p, 3, 1);
            for (int unit = cluster_id; unit < p.total_units; unit += n_ctas) {
                const TileCoord tc = decode_unit(p, unit, cta_rank, PAIR ? 2 : 1, N_TILE);
                const int bx = tc.ox0 * p.in_stride + p.x_off, by = tc.oy0 * p.in_stride;
                for (int ch = 0; ch < chunks_total; ++ch) {
                    const bool second = ch >= chunks0;
                    for (int v = 0; v < vpar; ++v) {  // (stride 2: the halo tiles of the four parity planes)
                        const int sel = (vpar > 1) ? v : tc.cls;
                        const int hx = bx + p.halo_x0[sel], hy = by + p.halo_y0[sel];
                        const int cf = (p.orient == 0) ? hx : hy, cs = (p.orient == 0) ? hy : hx;
                        mbar_wait(aempty0 + 8u * stage, phase ^ 1u);
                        trace_ev(p.trace, 4, tr_i++);
                        if (!PAIR) {
                            mbar_expect_tx(afull0 + 8u * stage, a_tx);
                            tma_load_5d(smem_base + (uint32_t)stage * a_slot_bytes, second ? &a1 : &a0, afull0 + 8u * stage,
                                        (second ? ch - chunks0 : ch) * kBlockK, cf, cs, 0, tc.b);
                        } else {
                            // both CTAs' halos count on the leader's barrier, which the leader arms for both
                            if (is_leader) mbar_expect_tx(afull0 + 8u * stage, 2u * a_tx);
                            tma_load_5d_pair(smem_base + (uint32_t)stage * a_slot_bytes, second ? &a1 : &a0, afull0 + 8u * stage,
                                             (second ? ch - chunks0 : ch) * kBlockK, cf, cs, 0, tc.b);
                        }
                        if (++stage == a_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
            MSI_BEACON(p, 3, 2);
        }
    } else if (warp == 2) {
        // =============================== W producer ===============================
        if (lane == 0) {
            int stage = 0;
            int tr_i = 0;
            uint32_t phase = 0;
            MSI_BEACON(p, 4, 1);
            for (int unit = cluster_id; unit < p.total_units; unit += n_ctas) {
                const TileCoord tc = decode_unit(p, unit, cta_rank, PAIR ? 2 : 1, N_TILE);
                const int ntaps = p.w_ntaps;
                int kb = tc.cls * chunks_total * ntaps;
                const int n_slots = chunks_total * (ntaps / T);
                for (int sl = 0; sl < n_slots; ++sl, kb += T) {
                    mbar_wait(wempty0 + 8u * stage, phase ^ 1u);
                    trace_ev(p.trace, 2, tr_i);
                    if (!PAIR) {
                        mbar_expect_tx(wfull0 + 8u * stage, w_slot_bytes);
                        tma_load_4d(w_ring + (uint32_t)stage * w_slot_bytes, &wmap, wfull0 + 8u * stage, 0, tc.n0, 0, kb);
                    } else {  // this rank's half of the rows ([kb][rank][cout][64] packing)
                        if (is_leader) mbar_expect_tx(wfull0 + 8u * stage, 2u * w_slot_bytes);
                        tma_load_4d_pair(w_ring + (uint32_t)stage * w_slot_bytes, &wmap, wfull0 + 8u * stage, 0, tc.n0,
                                         cta_rank, kb);
                    }
                    trace_ev(p.trace, 3, tr_i++);
                    if (++stage == w_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
            MSI_BEACON(p, 4, 2);
        }
    } else if (warp == 1 && (!PAIR || is_leader)) {
        // =============================== MMA issuer (PAIR: the leader CTA issues for both) ===============================
        // The issue thread is the critical resource (measured: ~80 clk per MMA issue + ~300 clk per
        // barrier round trip), so the full barrier of the NEXT slot is tested before this slot's MMAs
        // are issued (its latency hides behind them) and the tap offsets are read before the wait.
        const bool leader = elect_one();
        constexpr uint32_t idesc_wide = PAIR ? make_idesc_pair(2 * N_TILE) : make_idesc(2 * N_TILE);
        constexpr uint32_t idesc_n = PAIR ? make_idesc_pair(N_TILE) : make_idesc(N_TILE);
        // FP8X: the e4m3 weight rows follow this CTA's fp16 rows inside a tap (N rows, or N / 2 for a pair)
        constexpr uint64_t kW8Adv = (uint64_t)(((PAIR ? N_TILE / 2 : N_TILE) * kBlockK * 2) >> 4);
        constexpr uint64_t kDescFlags = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t a_desc_hi = kDescFlags | ((uint64_t)(p.PF * 8) << 32);  // SBO = PF rows of 128 bytes (>> 4)
        const uint64_t lo_adv = (uint64_t)(p.a_rows * 8);                       // hi halo -> lo halo
        int a_stage = 0, w_stage = 0;
        uint32_t a_phase = 0, w_phase = 0;
        bool a_ready = false, w_ready = false, t_ready = false;
        int local = 0;
        int tr_slot = 0, tr_chunk = 0;
        trace_ev(leader ? p.trace : nullptr, 7, 1);
        if (leader) MSI_BEACON(p, 1, 1);
        const int units_per_cls = p.units_per_col * p.n_tiles;
        for (int unit = cluster_id; unit < p.total_units; unit += n_ctas, ++local) {
            const int cls = (p.ncls > 1) ? unit / units_per_cls : 0;  // (a division only for the deconv classes)
            if (leader) MSI_BEACON(p, 5, (long long)local);
            const int acc = local & 1;
            const uint32_t use = (uint32_t)(local >> 1);
            // the epilogue has drained this accumulator (tested ahead, during the previous unit's last slot)
            if (!t_ready) mbar_wait(tempty0 + 8u * acc, (use & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            trace_ev(leader ? p.trace : nullptr, 8, local);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
            const uint32_t next_tempty = tempty0 + 8u * (uint32_t)(acc ^ 1);
            const uint32_t next_tparity = ((uint32_t)((local + 1) >> 1) & 1u) ^ 1u;
            uint32_t accumulate = 0u;
            const int vchunks = chunks_total * vpar;  // halo tiles per unit (stride 2: four parity planes per chunk)
            for (int vc = 0, v = 0; vc < vchunks; ++vc, v = (v + 1 == vpar) ? 0 : v + 1) {
                const int sel = (vpar > 1) ? v : cls;
                const int slots_per_chunk = s_ntaps[sel] / T;
                const bool last_tile = vc == vchunks - 1;
                if (!a_ready) mbar_wait(afull0 + 8u * a_stage, a_phase);
                if (tr_chunk == 0 && leader) trace_g(p.trace, tr_launch, 3);
                trace_ev(leader ? p.trace : nullptr, 5, tr_chunk++);
                const uint64_t da_tile = a_desc_hi | (uint64_t)(((smem_base + (uint32_t)a_stage * a_slot_bytes) >> 4) & 0x3FFF);
                if (++a_stage == a_stages) {
                    a_stage = 0;
                    a_phase ^= 1u;
                }
                a_ready = false;
                const uint32_t a_done_bar = aempty0 + 8u * (a_stage == 0 ? a_stages - 1 : a_stage - 1);
                for (int sl = 0; sl < slots_per_chunk; ++sl) {
                    int off[T];
#pragma unroll
                    for (int t = 0; t < T; ++t) off[t] = s_off[sel][sl * T + t] * 8;
                    const uint32_t cur = (uint32_t)w_stage;
                    if (!w_ready) mbar_wait(wfull0 + 8u * cur, w_phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (++w_stage == w_stages) {
                        w_stage = 0;
                        w_phase ^= 1u;
                    }
                    // Non-blocking tests of the barriers the NEXT iteration needs, issued before this
                    // slot's MMAs so that their latency hides behind the issue: the next W slot always,
                    // the next chunk's halo and the next unit's accumulator during a chunk's last slot
                    // (a blocking wait at a chunk / unit boundary drains the MMA queue: measured +400 /
                    // +1200 clk per boundary).
                    w_ready = mbar_test_wait(wfull0 + 8u * w_stage, w_phase);
                    if (sl == slots_per_chunk - 1) {
                        a_ready = mbar_test_wait(afull0 + 8u * a_stage, a_phase);
                        t_ready = last_tile ? mbar_test_wait(next_tempty, next_tparity) : false;
                    }
                    if (leader) {
                        trace_ev(p.trace, 0, tr_slot);
                        uint64_t db = make_desc(w_ring + cur * w_slot_bytes);
#pragma unroll
                        for (int t = 0; t < T; ++t, db += (uint64_t)(kWTapBytes >> 4)) {
                            const uint64_t da = da_tile + (uint64_t)off[t];
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);
                                if (FP8X) {
                                    if (!PAIR)
                                        umma_f16(d_tmem, da + adv, db + adv, idesc_n, accumulate);       // A_hi x W_hi
                                    else
                                        umma_f16_pair(d_tmem, da + adv, db + adv, idesc_n, accumulate);
                                    accumulate = 1u;
                                } else if (!PAIR) {
                                    umma_f16(d_tmem, da + adv, db + adv, idesc_wide, accumulate);  // A_hi x [W_hi | W_lo]
                                    accumulate = 1u;
                                    umma_f16(d_tmem, da + lo_adv + adv, db + adv, idesc_n, 1u);    // A_lo x W_hi
                                } else {
                                    umma_f16_pair(d_tmem, da + adv, db + adv, idesc_wide, accumulate);
                                    accumulate = 1u;
                                    umma_f16_pair(d_tmem, da + lo_adv + adv, db + adv, idesc_n, 1u);
                                }
                            }
                            if (FP8X) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k) {  // [A_hi8 | A_lo8] x [W_lo8 | W_hi8], K = 32 per step
                                    const uint64_t adv = (uint64_t)(k * 2);
                                    if (!PAIR)
                                        umma_f8(d_tmem, da + lo_adv + adv, db + kW8Adv + adv, idesc_n, 1u);
                                    else
                                        umma_f8_pair(d_tmem, da + lo_adv + adv, db + kW8Adv + adv, idesc_n, 1u);
                                }
                            }
                        }
                        if (!PAIR) {
                            umma_commit(wempty0 + 8u * cur);
                            if (sl == slots_per_chunk - 1) {
                                umma_commit(a_done_bar);
                                if (last_tile) umma_commit(tfull0 + 8u * acc);
                            }
                        } else {  // the slots and the accumulators of BOTH CTAs
                            umma_commit_pair(wempty0 + 8u * cur);
                            if (sl == slots_per_chunk - 1) {
                                umma_commit_pair(a_done_bar);
                                if (last_tile) umma_commit_pair(tfull0 + 8u * acc);
                            }
                        }
                        trace_ev(p.trace, 1, tr_slot);
                    }
                    ++tr_slot;
                    __syncwarp();
                }
            }
        }
        if (leader) trace_g(p.trace, tr_launch, 4);
        if (leader) MSI_BEACON(p, 1, 2);
    } else if (warp >= 3) {
        // =============================== epilogue (warps 3..10) ===============================
        const int row = (warp & 3) * 32 + lane;  // M index inside the tile: 8-pixel group = row / 8
        const int lx = (p.orient == 0) ? (row & 7) : (row >> 3);
        const int ly = (p.orient == 0) ? (row >> 3) : (row & 7);
        if (p.stage_out)
            epilogue_role<N_TILE, FP8X ? 0 : 1, PAIR ? 2 : 1, true, PAIR>(p, warp - 3, warp & 3, lane, lx, ly, cluster_id, n_ctas,
                                                                          cta_rank, tmem_base, tfull0, tempty0, &s_is_last, s_red,
                                                                          tr_launch, w_ring + (uint32_t)w_stages * w_slot_bytes);
        else
            epilogue_role<N_TILE, FP8X ? 0 : 1, PAIR ? 2 : 1, false, PAIR>(p, warp - 3, warp & 3, lane, lx, ly, cluster_id, n_ctas,
                                                                           cta_rank, tmem_base, tfull0, tempty0, &s_is_last, s_red,
                                                                           tr_launch);
    }

    if (threadIdx.x == 0) MSI_BEACON(p, 0, 9);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) trace_ev(p.trace, 7, 2);
    if (threadIdx.x == 0 && *(volatile int*)&s_is_last >= 0) trace_g(p.trace, tr_launch, 7);
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 10);
    if (PAIR) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still touch its memory / barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR)
            tmem_dealloc_pair<kTmemCols>(tmem_base);
        else
            tmem_dealloc<kTmemCols>(tmem_base);
    }
    if (threadIdx.x == 0) MSI_BEACON(p, 0, 0);
}

// e4m3 weight copies of MSI_PREC_FP16_FP8X (scales: net_internal.cuh)
__device__ __forceinline__ uint8_t w_lo8(float w_scaled, __half h) {
    return (uint8_t)__nv_cvt_float_to_fp8((w_scaled - __half2float(h)) * (float)(1 << kFp8HiShift), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint8_t w_hi8(__half h) {
    return (uint8_t)__nv_cvt_float_to_fp8(__half2float(h) * (1.0f / (float)(1 << kFp8LoShift)), __NV_SATFINITE, __NV_E4M3);
}

// ---- weight packing ----------------------------------------------------------------------------
// w_f32 (TF layout) -> fp16 hi/lo [cls][cout][K], K = tap * cs_total + packed channel, x MSI_WEIGHT_SCALE.
struct PackParams {
    const float* w;
    __half* hi;
    __half* lo;
    int fp8x;  // MSI_PREC_FP16_FP8X layer: the "lo" rows hold [w_lo8 x 64 | w_hi8 x 64] (e4m3) instead of fp16 residuals
    int kind, ncls, cout, K, cs_total, cin_total, coord;
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
            v = q.w[((size_t)wt * (q.cin_total + q.coord) + c) * q.cout + n];
        else
            v = q.w[(size_t)c * q.cout + n];
    }
    __half h, l;
    split_half(v * MSI_WEIGHT_SCALE, h, l);
    q.hi[idx] = h;
    if (!q.fp8x) {
        q.lo[idx] = l;
    } else {
        // the row of 64-channel K block kc as 128 bytes: w_lo8 (pairs with A_hi8), then w_hi8 (pairs with A_lo8)
        uint8_t* row = reinterpret_cast<uint8_t*>(q.lo) + ((idx - (k & 63)) * 2);
        row[k & 63] = w_lo8(v * MSI_WEIGHT_SCALE, h);
        row[64 + (k & 63)] = w_hi8(h);
    }
}

// K-block-major packing of the halo kernel: [kb][hi|lo][cout][64] with kb = (cls * chunks + chunk) *
// ntaps + tap, so that the [W_hi | W_lo] tiles of T consecutive taps of one chunk are ONE 4-D TMA box.
// pair_tile > 0 (CTA-pair kernel, pair_tile = N_TILE): the second index is the CTA rank instead, and the
// rows of every N tile are arranged [W_hi of couts r * N/2 .. | W_lo of couts (1 - r) * N/2 ..] for rank r.
__global__ void __launch_bounds__(256) pack_weights_halo_kernel(PackParams q, int chunks_total, int ntaps, int pair_tile) {
    const long long total = (long long)q.ncls * chunks_total * ntaps * q.cout * kBlockK;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int k64 = (int)(idx % kBlockK);
    const int n = (int)((idx / kBlockK) % q.cout);
    const int kb = (int)(idx / ((long long)kBlockK * q.cout));
    const int t = kb % ntaps;
    const int chunk = (kb / ntaps) % chunks_total;
    const int cls = kb / (ntaps * chunks_total);
    int pc = chunk * kBlockK + k64;  // packed channel
    int c = -1;                      // channel in the concatenated TF tensor
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
        else
            v = q.w[((size_t)wt * (q.cin_total + q.coord) + c) * q.cout + n];
    }
    __half h, l;
    split_half(v * MSI_WEIGHT_SCALE, h, l);
    if (q.fp8x) {
        // fp16 row and e4m3 row ([w_lo8 x 64 | w_hi8 x 64]) of this cout.  Single CTA: planes [hi rows | e4m3 rows] of the
        // K block.  Pair: rank r holds couts [r N/2, (r+1) N/2) of every N tile, fp16 rows first, their e4m3 rows after.
        size_t row_h, row_8;
        if (pair_tile == 0) {
            row_h = ((size_t)kb * 2 + 0) * q.cout + n;
            row_8 = ((size_t)kb * 2 + 1) * q.cout + n;
        } else {
            const int half_rows = pair_tile / 2;
            const int tile0 = (n / pair_tile) * pair_tile, in_tile = n % pair_tile;
            const int r = in_tile / half_rows, i = in_tile % half_rows;
            row_h = ((size_t)kb * 2 + r) * q.cout + tile0 + i;
            row_8 = ((size_t)kb * 2 + r) * q.cout + tile0 + half_rows + i;
        }
        q.hi[row_h * kBlockK + k64] = h;
        uint8_t* row = reinterpret_cast<uint8_t*>(q.hi + row_8 * kBlockK);
        row[k64] = w_lo8(v * MSI_WEIGHT_SCALE, h);
        row[64 + k64] = w_hi8(h);
    } else if (pair_tile == 0) {
        q.hi[(((size_t)kb * 2 + 0) * q.cout + n) * kBlockK + k64] = h;
        q.hi[(((size_t)kb * 2 + 1) * q.cout + n) * kBlockK + k64] = l;
    } else {
        const int half_rows = pair_tile / 2;
        const int tile0 = (n / pair_tile) * pair_tile, in_tile = n % pair_tile;
        const int r = in_tile / half_rows, i = in_tile % half_rows;  // W_hi of this cout lives in rank r, W_lo in rank 1 - r
        q.hi[(((size_t)kb * 2 + r) * q.cout + tile0 + i) * kBlockK + k64] = h;
        q.hi[(((size_t)kb * 2 + (1 - r)) * q.cout + tile0 + half_rows + i) * kBlockK + k64] = l;
    }
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

// 5-D activation map of the halo kernel: {C, fast, slow, hi|lo, B}, box {64, PF, PS, 2, 1}.  The hi
// and lo tensors are two carve-outs of one workspace, so "lo" is "hi" at a constant byte offset.
int encode_act_map5(CUtensorMap* m, const __half* hi, const __half* lo, int C, int W, int H, int B, int orient, int PF,
                    int PS, int estride) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return MSI_ERR_CUDA;
    }
    const long long hl = (const char*)lo - (const char*)hi;
    if (hl <= 0 || hl % 16 != 0) {
        set_error("conv_tc: hi/lo activation buffers are not at a positive 16-byte-aligned offset");
        return MSI_ERR_UNSUPPORTED;
    }
    const cuuint64_t row = (cuuint64_t)C * 2, img_row = (cuuint64_t)W * C * 2;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)(orient == 0 ? W : H), (cuuint64_t)(orient == 0 ? H : W), 2,
                          (cuuint64_t)B};
    cuuint64_t strides[4] = {orient == 0 ? row : img_row, orient == 0 ? img_row : row, (cuuint64_t)hl,
                             (cuuint64_t)H * W * C * 2};
    // estride = 2 (stride-2 convs): every second pixel of the traversed 2 PF x 2 PS window = one parity plane's halo tile
    cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)(PF * estride), (cuuint32_t)(PS * estride), 2, 1};
    cuuint32_t es[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)hi, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(halo activation C=%d W=%d H=%d B=%d orient=%d box=%dx%d) failed: %d", C, W, H, B,
                  orient, PF, PS, (int)r);
        return MSI_ERR_CUDA;
    }
    return MSI_OK;
}

// 4-D weight map of the halo kernel over [kb][hi|lo][cout][64], box {64, n_tile, 2, T}
int encode_w_map4(CUtensorMap* m, const __half* base, int cout, int nkb, int n_tile, int T, int planes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return MSI_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)kBlockK, (cuuint64_t)cout, 2, (cuuint64_t)nkb};
    cuuint64_t strides[3] = {(cuuint64_t)kBlockK * 2, (cuuint64_t)cout * kBlockK * 2, (cuuint64_t)2 * cout * kBlockK * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)n_tile, (cuuint32_t)planes, (cuuint32_t)T};  // planes: 2 = [hi | lo], 1 = one rank (pair)
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(halo weights cout=%d nkb=%d n_tile=%d T=%d) failed: %d", cout, nkb, n_tile, T, (int)r);
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

int num_sms() {
    static int n = 0;
    if (n > 0) return n;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0) {
        cudaGetLastError();
        n = 0;
        return 148;
    }
    return n;
}

template <int N_TILE, int SPLIT, int CL>
int launch_tc(const TcPlan* plan, const TcParams& p, bool pdl, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = conv_igemm_tcgen05_kernel<N_TILE, SPLIT, CL>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_igemm_tcgen05_kernel<%d,%d,%d>, %d) failed: %s", N_TILE, SPLIT, CL,
                      kMaxDynSmem, cudaGetErrorString(e));
            return MSI_ERR_CUDA;
        }
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, plan->a_map[0][0], plan->a_map[0][1], plan->a_map[1][0],
                                       plan->a_map[1][1], plan->w_map[0], plan->w_map[1], p);
    if (e != cudaSuccess) {
        set_error("cudaLaunchKernelEx(conv_igemm_tcgen05_kernel<%d,%d,%d>, grid %d) failed: %s", N_TILE, SPLIT, CL,
                  plan->grid, cudaGetErrorString(e));
        return MSI_ERR_CUDA;
    }
    return MSI_OK;
}

template <int N_TILE>
int launch_tc_nt(const TcPlan* plan, const TcParams& p, bool pdl, cudaStream_t st) {
    if (plan->cl == 2)
        return plan->split ? launch_tc<N_TILE, 1, 2>(plan, p, pdl, st) : launch_tc<N_TILE, 0, 2>(plan, p, pdl, st);
    return plan->split ? launch_tc<N_TILE, 1, 1>(plan, p, pdl, st) : launch_tc<N_TILE, 0, 1>(plan, p, pdl, st);
}

// MSI_BEACON=1: host-mapped progress beacons of every conv launch (see MSI_BEACON in the kernels)
long long* g_beacon_host = nullptr;
long long* g_beacon_dev = nullptr;
int g_beacon_next = 0;
struct BeaconInfo {
    char scope[16];
    int grid, halo, pair, fp8x, n_tile;
};
BeaconInfo g_beacon_info[kBeaconSeqs];
bool beacon_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* env = getenv("MSI_BEACON");
        on = (env && atoi(env) == 1) ? 1 : 0;
    }
    return on == 1;
}
long long* beacon_buffer() {
    if (!beacon_enabled()) return nullptr;
    if (!g_beacon_host) {
        const size_t bytes = sizeof(long long) * kBeaconSeqs * kBeaconCtas * kBeaconSlots;
        if (cudaHostAlloc(&g_beacon_host, bytes, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&g_beacon_dev, g_beacon_host, 0) != cudaSuccess) {
            cudaGetLastError();
            g_beacon_host = g_beacon_dev = nullptr;
            return nullptr;
        }
        memset(g_beacon_host, 0, bytes);
    }
    return g_beacon_dev;
}

long long* g_trace_dev = nullptr;
long long* trace_buffer() {
    if (!g_trace_dev) {
        if (cudaMalloc(&g_trace_dev, sizeof(long long) * kTraceRegion * kTraceRegions) != cudaSuccess) {
            cudaGetLastError();
            g_trace_dev = nullptr;
        } else {
            cudaMemset(g_trace_dev, 0, sizeof(long long) * kTraceRegion * kTraceRegions);
        }
    }
    return g_trace_dev;
}

// Launch attribute for programmatic dependent launch (MSI_PDL=0 turns it off): the kernel may begin
// before the previous kernel in the stream has finished; it calls griddepcontrol.wait before
// touching anything that kernel wrote.  `first` = no kernel precedes it in the forward (memset).
template <int N_TILE, int T, bool PAIR, bool FP8X = false>
int launch_halo(const TcPlan* plan, const TcParams& p, bool pdl, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = conv_halo_tcgen05_kernel<N_TILE, T, PAIR, FP8X>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_halo_tcgen05_kernel<%d,%d>, %d) failed: %s", N_TILE, T, kMaxDynSmem,
                      cudaGetErrorString(e));
            return MSI_ERR_CUDA;
        }
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan->grid);
    cfg.blockDim = dim3(kHaloThreads);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (PAIR) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, plan->a_map[0][0], plan->a_map[1][0], plan->w_map[0], p);
    if (e != cudaSuccess) {
        set_error("cudaLaunchKernelEx(conv_halo_tcgen05_kernel<%d,%d>, grid %d) failed: %s", N_TILE, T, plan->grid,
                  cudaGetErrorString(e));
        return MSI_ERR_CUDA;
    }
    return MSI_OK;
}

// MSI_CONV_HALO_S2=1: the stride-2 convs (conv1_2, conv2_2, conv3_3) run on the halo kernel with four parity-plane halo
// tiles per chunk (TcParams::vpar).  Built, parity-tested (tests/test_gpu_net.py::test_stride2_halo_matches_per_tap_kernel)
// and measured on B200: 27.1 / 28.0 / 28.5 us against 25.9 / 27.3 / 25.1 us on the per-tap kernel -- no gain, so the
// default stays off.  The pipeline trace (profiles/r2_trace_conv{1_2,3_3}_s2.log) shows why: the strided halo loads
// complete ~4-5 K clocks after issue and the MMA loop runs at ~760 clocks per tap against 592 at the pipe's rate (per-SM
// TMA ingest, ~50 B/clk, is the limit in both forms: these layers have one or two N tiles per pixel tile, so there is
// little operand reuse to win), and 14 of the 26 us of a launch are fixed costs outside the MMA loop (first loads 4 us,
// last epilogue 3.7 us, store drain + statistics 2.5 us, CTA skew + finalisation 3.5 us).
bool stride2_halo_enabled() {
    const char* env = getenv("MSI_CONV_HALO_S2");
    return env && atoi(env) == 1;
}

// Halo-kernel plan: tile orientation, halo extents, ring sizes, 5-D / 4-D tensor maps.  Returns
// MSI_ERR_UNSUPPORTED when the layer does not fit (the caller then uses the per-tap kernel).
int plan_halo(TcPlan* plan, const LayerPlan& L, const ActBuf* srcs, int max_batch) {
    TcParams& p = plan->p;
    const int ext = -2 * p.grid_off;  // wrap deconv: tiles cover the class grid extended by 3 on every side
    const int c0 = ((p.Mw + ext + 7) / 8) * ((p.Mh + ext + 15) / 16), c1 = ((p.Mw + ext + 15) / 16) * ((p.Mh + ext + 7) / 8);
    p.orient = (c1 < c0) ? 1 : 0;
    p.BW = p.orient == 0 ? 8 : 16;
    p.BH = p.orient == 0 ? 16 : 8;
    int ext_x = -1, ext_y = -1;
    const bool s2 = (L.kind == kConv && p.in_stride == 2);
    p.vpar = s2 ? 4 : 1;
    p.w_ntaps = p.taps[0].n;
    int rel_x[4][9], rel_y[4][9];  // tap position inside the halo tile of its `sel` (class or parity plane), in tile pixels
    if (s2) {
        // input pixel 2 o + d = parity plane (d mod 2), plane pixel o + floor(d / 2)
        if (plan->n_tile != 128 || p.ncls != 1 || p.taps[0].n > 9) return MSI_ERR_UNSUPPORTED;
        const TapList src = p.taps[0];
        int par[9], qx[9], qy[9];
        int qminx = 1 << 20, qmaxx = -(1 << 20), qminy = 1 << 20, qmaxy = -(1 << 20);
        for (int t = 0; t < src.n; ++t) {
            const int px = ((src.dx[t] % 2) + 2) % 2, py = ((src.dy[t] % 2) + 2) % 2;
            qx[t] = (src.dx[t] - px) / 2;
            qy[t] = (src.dy[t] - py) / 2;
            par[t] = py * 2 + px;
            qminx = std::min(qminx, qx[t]);
            qmaxx = std::max(qmaxx, qx[t]);
            qminy = std::min(qminy, qy[t]);
            qmaxy = std::max(qmaxy, qy[t]);
        }
        TapList dst = src;
        int n = 0;
        for (int v = 0; v < 4; ++v) {  // taps plane by plane: the order of the packed weights and of the W ring
            p.v_n[v] = 0;
            for (int t = 0; t < src.n; ++t) {
                if (par[t] != v) continue;
                dst.dx[n] = src.dx[t];
                dst.dy[n] = src.dy[t];
                dst.wtap[n] = src.wtap[t];
                rel_x[v][p.v_n[v]] = qx[t] - qminx;
                rel_y[v][p.v_n[v]] = qy[t] - qminy;
                ++p.v_n[v];
                ++n;
            }
            if (p.v_n[v] == 0) return MSI_ERR_UNSUPPORTED;
            p.halo_x0[v] = 2 * qminx + (v & 1);
            p.halo_y0[v] = 2 * qminy + (v >> 1);
        }
        p.taps[0] = dst;
        ext_x = p.BW + qmaxx - qminx;
        ext_y = p.BH + qmaxy - qminy;
    } else {
        for (int c = 0; c < p.ncls; ++c) {
            int minx = 1 << 20, maxx = -(1 << 20), miny = 1 << 20, maxy = -(1 << 20);
            for (int t = 0; t < p.taps[c].n; ++t) {
                minx = std::min(minx, p.taps[c].dx[t]);
                maxx = std::max(maxx, p.taps[c].dx[t]);
                miny = std::min(miny, p.taps[c].dy[t]);
                maxy = std::max(maxy, p.taps[c].dy[t]);
            }
            p.halo_x0[c] = minx;
            p.halo_y0[c] = miny;
            const int ex = p.BW + maxx - minx, ey = p.BH + maxy - miny;
            if (c > 0 && (ex != ext_x || ey != ext_y || p.taps[c].n != p.taps[0].n)) return MSI_ERR_UNSUPPORTED;
            ext_x = ex;
            ext_y = ey;
            p.v_n[c] = p.taps[c].n;
            for (int t = 0; t < p.taps[c].n; ++t) {
                rel_x[c][t] = p.taps[c].dx[t] - minx;
                rel_y[c][t] = p.taps[c].dy[t] - miny;
            }
        }
        for (int c = p.ncls; c < 4; ++c) p.v_n[c] = 0;
    }
    p.PF = p.orient == 0 ? ext_x : ext_y;
    p.PS = p.orient == 0 ? ext_y : ext_x;
    for (int c = 0; c < 4; ++c)
        for (int t = 0; t < 9; ++t)
            p.v_off[c][t] = (t < p.v_n[c]) ? ((p.orient == 0) ? rel_y[c][t] * p.PF + rel_x[c][t] : rel_x[c][t] * p.PF + rel_y[c][t]) : 0;
    p.a_rows = p.PF * p.PS;
    p.a_slot_bytes = (2 * p.a_rows * 128 + 1023) / 1024 * 1024;
    // stride 2: four smaller tiles per chunk, the shortest serving ONE tap -- a third slot keeps the loads two tiles ahead
    p.a_stages = s2 ? 3 : 2;
    const int ntaps = p.taps[0].n;
    p.T = (plan->n_tile == 64) ? ((ntaps % 3 == 0) ? 3 : 2) : 1;
    {
        const char* env = getenv("MSI_HALO_T");  // experiment: taps per W slot for the Cout = 64 layers
        if (env && plan->n_tile == 64 && atoi(env) == 1) p.T = 1;  // measured: 667 clk per tap (issue-thread bound) vs 572 for T = 3
    }
    if (ntaps % p.T != 0 || p.PS > 256 || p.PF > 256 || p.PF * p.in_stride > 256 || p.PS * p.in_stride > 256) return MSI_ERR_UNSUPPORTED;
    if (s2 && p.T != 1) return MSI_ERR_UNSUPPORTED;  // (the planes serve 4 + 2 + 2 + 1 taps)
    if (plan->pair && plan->n_tile == 64 && p.T == 1) return MSI_ERR_UNSUPPORTED;  // (no such instantiation)
    // A pair synchronises its two CTAs at every unit boundary (accumulator hand-over across the cluster): with only
    // a few W slots per unit that costs more than the halved weight traffic wins (measured: conv8_2, 3 slots per
    // unit, 56 us as pairs vs 51 us; every layer with >= 8 slots per unit is 10-18 % faster as pairs).
    if (plan->pair && (p.chunks[0] + p.chunks[1]) * (ntaps / p.T) < 6) return MSI_ERR_UNSUPPORTED;
    const int w_slot = p.T * (plan->pair ? 1 : 2) * plan->n_tile * kBlockK * 2;  // pair: this CTA's half of the rows
    const int budget = kMaxDynSmem - smem_reserve() - 1024 - p.a_stages * p.a_slot_bytes;
    const int stage_bytes = kEpiWarps * 4096;  // one 32 x 32 float tile per epilogue warp
    {
        const char* env = getenv("MSI_CONV_STAGE");
        const int with = (budget - stage_bytes) / w_slot;
        p.stage_out = (!(env && atoi(env) == 0) && with >= (p.T == 1 ? 3 : 2)) ? 1 : 0;
        p.w_stages = p.stage_out ? with : budget / w_slot;
    }
    if (p.w_stages > 8) p.w_stages = 8;
    if (p.w_stages < 2) return MSI_ERR_UNSUPPORTED;
    plan->smem_bytes = 1024 + p.a_stages * p.a_slot_bytes + p.w_stages * w_slot + (p.stage_out ? stage_bytes : 0);
    p.tiles_x = (p.Mw + ext + p.BW - 1) / p.BW;
    p.tiles_y = (p.Mh + ext + p.BH - 1) / p.BH;
    // Pairs only where ONE frame has at least two pixel tiles: the choice must not depend on the batch size, because the
    // pair kernel sums a cout's three partial products in a different order than the single-CTA kernel and a frame's
    // result has to be the same bits whatever batch it is part of (tests: batch of 3 == three single frames).
    if (plan->pair && p.tiles_x * p.tiles_y < 2) return MSI_ERR_UNSUPPORTED;
    if (L.w_lo != L.w_hi + (size_t)L.ncls * L.cout * L.K) return MSI_ERR_UNSUPPORTED;  // one [.. hi|lo ..] buffer
    int rc = MSI_OK;
    for (int s = 0; s < L.nsrc && rc == MSI_OK; ++s)
        rc = encode_act_map5(&plan->a_map[s][0], srcs[s].hi, plan->fp8x ? reinterpret_cast<const __half*>(srcs[s].q8) : srcs[s].lo,
                             srcs[s].c_stride, srcs[s].Wp, srcs[s].H, max_batch, p.orient, p.PF, p.PS, p.in_stride);
    if (rc == MSI_OK && L.nsrc == 1) plan->a_map[1][0] = plan->a_map[0][0];
    const int nkb = L.ncls * (p.chunks[0] + p.chunks[1]) * ntaps;
    if (rc == MSI_OK) rc = encode_w_map4(&plan->w_map[0], L.w_hi, L.cout, nkb, plan->n_tile, p.T, plan->pair ? 1 : 2);
    return rc;
}

}  // namespace

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* env = getenv("MSI_PDL");
        on = (env && atoi(env) == 0) ? 0 : 1;
    }
    return on == 1;
}

int conv_tc_plan_create(LayerPlan& L, const ActBuf* srcs, int max_batch, int precision) {
    TcPlan* plan = new TcPlan();
    memset(plan, 0, sizeof(TcPlan));
    TcParams& p = plan->p;
    plan->split = (precision == MSI_PREC_FP16X3 || precision == MSI_PREC_FP16_FP8X) ? 1 : 0;
    plan->n_tile = (L.cout % 128 == 0) ? 128 : 64;
    plan->fp8x = (L.fp8x && plan->n_tile == 128) ? 1 : 0;
    if (L.fp8x && !plan->fp8x) {
        delete plan;
        set_error("conv_tc: layer %s is marked fp8x but its N tile is %d", L.scope, plan->n_tile);
        return MSI_ERR_STATE;
    }
    if (L.cout % plan->n_tile != 0) {
        delete plan;
        set_error("conv_tc: layer %s has cout=%d; the tcgen05 back end needs a multiple of 64", L.scope, L.cout);
        return MSI_ERR_UNSUPPORTED;
    }
    for (int s = 0; s < L.nsrc; ++s) {
        const void* second = plan->fp8x ? (const void*)srcs[s].q8 : (const void*)srcs[s].lo;
        if (srcs[s].hi == nullptr || (plan->split && second == nullptr)) {
            delete plan;
            set_error("conv_tc: layer %s source %d lacks the %s operand", L.scope, s, plan->fp8x ? "e4m3" : "fp16 lo");
            return MSI_ERR_STATE;
        }
    }
    p.kind = L.kind;
    p.nsrc = L.nsrc;
    p.x_off = srcs[0].x_pad;
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
    const bool wrap_deconv = (L.kind == kDeconv) && srcs[0].x_pad > 0;
    p.grid_off = wrap_deconv ? -3 : 0;
    p.stat_all = wrap_deconv ? 1 : 0;
    const int ext = wrap_deconv ? 6 : 0;  // tiles cover [-3, M + 3)
    pick_tile(p.Mh + ext, p.Mw + ext, p.BH, p.BW);
    p.tiles_x = (p.Mw + ext + p.BW - 1) / p.BW;
    p.tiles_y = (p.Mh + ext + p.BH - 1) / p.BH;
    p.n_tiles = L.cout / plan->n_tile;
    // Optional: pair CTAs along M so that each fetches half of the W tile and multicasts it to its
    // peer (MSI_CONV_CLUSTER=2).  Measured on B200 (profiles/): no gain -- 1.100 ms vs 1.087 ms per
    // frame -- because the kernel is bound by bytes delivered INTO each SM, which multicast does not
    // reduce; so the default stays 1.
    {
        const char* env = getenv("MSI_CONV_CLUSTER");
        const int want = env ? atoi(env) : 1;
        plan->cl = (want >= 2 && p.tiles_x * p.tiles_y >= 2) ? 2 : 1;
    }
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
    const int stage_bytes = (kATileBytes + plan->n_tile * kBlockK * 2) * (plan->split ? 2 : 1);
    // the head keeps room for the [128 pixels][n_tile] float tile of its fused RGBA epilogue
    const int rgba_tile_bytes = (L.kind == kHead) ? kBlockM * plan->n_tile * 4 : 0;
    p.stages = (kMaxDynSmem - smem_reserve() - 1024 - rgba_tile_bytes) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) {
        delete plan;
        set_error("conv_tc: layer %s does not fit 2 pipeline stages", L.scope);
        return MSI_ERR_UNSUPPORTED;
    }
    plan->smem_bytes = p.stages * stage_bytes + 1024 + rgba_tile_bytes;
    p.do_stats = (L.kind != kHead) ? 1 : 0;
    p.n_partials = L.n_partials;
    p.partials = L.partials;
    p.counter = L.counter;
    p.stats = L.stats;
    p.n_per_sample = (double)L.Hout * L.Wout * L.cout;
    if (wrap_deconv) p.n_per_sample = (double)(L.Hout + 10) * (L.Wout + 10) * L.cout;  // LayerNorm before the crop

    {
        const char* env = getenv("MSI_CONV_HALO");
        const bool want = !(env && atoi(env) == 0) && plan->split && plan->cl == 1 &&
                          ((L.kind == kConv && (L.stride == 1 || (L.stride == 2 && stride2_halo_enabled()))) || L.kind == kDeconv);
        if (want) {
            TcPlan saved = *plan;
            // CTA pairs (tcgen05 cta_group::2): the two CTAs of a cluster take two consecutive pixel tiles of a
            // column (cl = 2) and each holds half of the weight rows.  Default on (MSI_CONV_PAIR=0 turns it off);
            // plan_halo declines it for layers with too little work per unit, which then run one CTA per unit.
            const char* pe = getenv("MSI_CONV_PAIR");
            plan->pair = !(pe && atoi(pe) == 0) ? 1 : 0;
            if (plan->pair) plan->cl = 2;
            int hrc = plan_halo(plan, L, srcs, max_batch);
            if (hrc != MSI_OK && plan->pair) {
                *plan = saved;
                hrc = plan_halo(plan, L, srcs, max_batch);
            }
            if (hrc == MSI_OK) {
                plan->halo = 1;
                const char* tr = getenv("MSI_TC_TRACE");  // debugging: scope name of the layer to trace
                plan->p.trace = (tr && strcmp(tr, L.scope) == 0) ? trace_buffer() : nullptr;
                L.tc_plan = plan;
                return MSI_OK;
            }
            *plan = saved;  // does not fit: per-tap kernel
        }
    }
    int rc = MSI_OK;
    for (int s = 0; s < L.nsrc && rc == MSI_OK; ++s) {
        rc = encode_act_map(&plan->a_map[s][0], srcs[s].hi, srcs[s].c_stride, srcs[s].Wp, srcs[s].H, max_batch, p.BW,
                            p.BH, p.in_stride);
        if (rc == MSI_OK)
            rc = encode_act_map(&plan->a_map[s][1], plan->fp8x ? reinterpret_cast<const __half*>(srcs[s].q8) : srcs[s].lo,
                                srcs[s].c_stride, srcs[s].Wp, srcs[s].H, max_batch, p.BW, p.BH, p.in_stride);
    }
    if (rc == MSI_OK && L.nsrc == 1) {
        plan->a_map[1][0] = plan->a_map[0][0];
        plan->a_map[1][1] = plan->a_map[0][1];
    }
    if (rc == MSI_OK) rc = encode_w_map(&plan->w_map[0], L.w_hi, L.K, L.cout, L.ncls, plan->n_tile / plan->cl);
    if (rc == MSI_OK) rc = encode_w_map(&plan->w_map[1], L.w_lo, L.K, L.cout, L.ncls, plan->n_tile / plan->cl);
    if (rc != MSI_OK) {
        delete plan;
        return rc;
    }
    L.tc_plan = plan;
    return MSI_OK;
}

}  // namespace msi

// Debugging hook (not part of include/msi_b200.h): prints, for every conv launch slot whose CTAs have not all exited, the
// phase each role of each such CTA has reached (MSI_BEACON=1; readable while the GPU is stuck).  Returns the number of
// unfinished CTAs.
extern "C" int msi_debug_beacon_dump(void) {
    using namespace msi;
    if (!g_beacon_host) {
        fprintf(stderr, "[beacon] not enabled (MSI_BEACON=1)\n");
        return -1;
    }
    int open_ctas = 0;
    for (int seq = 0; seq < kBeaconSeqs; ++seq) {
        const BeaconInfo& bi = g_beacon_info[seq];
        int n_open = 0;
        for (int c = 0; c < kBeaconCtas; ++c)
            if (((volatile long long*)g_beacon_host)[((size_t)seq * kBeaconCtas + c) * kBeaconSlots] != 0) ++n_open;
        if (n_open == 0) continue;
        open_ctas += n_open;
        fprintf(stderr, "[beacon] seq %d layer %s grid %d halo %d pair %d fp8x %d n_tile %d: %d CTAs not exited\n", seq, bi.scope,
                bi.grid, bi.halo, bi.pair, bi.fp8x, bi.n_tile, n_open);
        int shown = 0;
        for (int c = 0; c < kBeaconCtas && shown < 24; ++c) {
            volatile long long* b = (volatile long long*)g_beacon_host + ((size_t)seq * kBeaconCtas + c) * kBeaconSlots;
            if (b[0] == 0) continue;
            fprintf(stderr, "[beacon]    cta %3d: main %lld mma %lld (unit %lld) epi %lld (unit %lld) A-prod %lld W-prod %lld\n", c,
                    b[0], b[1], b[5], b[2], b[6], b[3], b[4]);
            ++shown;
        }
    }
    fprintf(stderr, "[beacon] %d CTAs not exited in total\n", open_ctas);
    return open_ctas;
}

// Debugging hook (not part of include/msi_b200.h): copies the pipeline-event clocks that CTA 0 of the
// layer named by MSI_TC_TRACE recorded during its last launch.  Returns the number of entries.
extern "C" int msi_debug_conv_trace(long long* host_out, int max_entries) {
    const int n = msi::kTraceRegion * msi::kTraceRegions;
    if (!msi::g_trace_dev || !host_out || max_entries < n) return 0;
    if (cudaMemcpy(host_out, msi::g_trace_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
}

namespace msi {

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
    q.coord = L.coord ? 1 : 0;
    q.fp8x = L.fp8x ? 1 : 0;
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
    const TcPlan* plan = reinterpret_cast<const TcPlan*>(L.tc_plan);
    if (plan && plan->halo)  // the halo plan's own tap order (stride 2: plane by plane)
        for (int c = 0; c < plan->p.ncls; ++c) q.taps[c] = plan->p.taps[c];
    if (plan && plan->halo)
        pack_weights_halo_kernel<<<ceil_div(total, 256), 256, 0, st>>>(q, plan->p.chunks[0] + plan->p.chunks[1],
                                                                        plan->p.taps[0].n, plan->pair ? plan->n_tile : 0);
    else
        pack_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(q);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

// The caller has zeroed L.partials / L.counter on `st` before this launch (net.cu does one memset
// for all layers per forward).  On return L.stats holds (mean, rstd) per frame.
bool conv_tc_can_fuse_rgba(const LayerPlan& L) {
    const TcPlan* plan = reinterpret_cast<const TcPlan*>(L.tc_plan);
    return plan && L.kind == kHead && !plan->halo && plan->split && plan->cl == 1 && plan->n_tile == L.cout &&
           (L.cout == 64 || L.cout == 128);
}

int conv_tc_forward(const LayerPlan& L, int B, float* out, bool after_kernel, cudaStream_t st, const HeadFuse* fuse) {
    TcPlan* plan = reinterpret_cast<TcPlan*>(L.tc_plan);
    if (!plan) {
        set_error("conv_tc_forward: layer %s has no plan", L.scope);
        return MSI_ERR_STATE;
    }
    TcParams p = plan->p;
    p.out = out;
    p.B = B;
    p.rgba = nullptr;
    if (fuse != nullptr) {
        if (!conv_tc_can_fuse_rgba(L)) {
            set_error("conv_tc_forward: layer %s cannot take the fused RGBA epilogue", L.scope);
            return MSI_ERR_UNSUPPORTED;
        }
        p.rgba = reinterpret_cast<float4*>(fuse->rgba);
        p.psv_hi = fuse->psv_hi;
        p.psv_lo = fuse->psv_lo;
        p.psv_cstride = fuse->c_stride;
        p.psv_Wp = fuse->Wp;
        p.psv_xpad = fuse->x_pad;
    }
    p.beacon = beacon_buffer();
    p.beacon_seq = 0;
    {
        static int relq = -1;
        if (relq < 0) {
            const char* env = getenv("MSI_PAIR_RELINQUISH");
            relq = (env && atoi(env) == 1) ? 1 : 0;
        }
        p.pair_relinquish = relq;
    }
    const int cl = plan->cl;
    p.m_tiles = B * p.tiles_x * p.tiles_y;
    p.units_per_col = (p.m_tiles + cl - 1) / cl;
    p.total_units = p.ncls * p.n_tiles * p.units_per_col;
    int clusters = num_sms();
    if (clusters > kMaxPersistentCtas) clusters = kMaxPersistentCtas;
    {
        // experiment (MSI_CONV_MAX_CTAS): a conv launch takes at most this many SMs, so that two frames' conv kernels run side
        // by side and one's start-up / last epilogue overlaps the other's MMA loop (runtime.MSIFrameLanes)
        static int cap = -1;
        if (cap < 0) {
            const char* env = getenv("MSI_CONV_MAX_CTAS");
            cap = env ? atoi(env) : 0;
            if (cap < 2) cap = 0;
        }
        if (cap > 0 && clusters > cap) clusters = cap;
    }
    clusters /= cl;
    if (clusters > p.total_units) clusters = p.total_units;
    const int grid = clusters * cl;
    plan->grid = grid;
    if (p.do_stats && grid * kEpiWarps > p.n_partials) {
        set_error("conv_tc_forward: %d partial slots < %d", p.n_partials, grid * kEpiWarps);
        return MSI_ERR_STATE;
    }
    if (p.beacon != nullptr) {
        p.beacon_seq = g_beacon_next;
        g_beacon_next = (g_beacon_next + 1) % kBeaconSeqs;
        BeaconInfo& bi = g_beacon_info[p.beacon_seq];
        strncpy(bi.scope, L.scope, sizeof(bi.scope) - 1);
        bi.scope[sizeof(bi.scope) - 1] = 0;
        bi.grid = grid;
        bi.halo = plan->halo;
        bi.pair = plan->pair;
        bi.fp8x = plan->fp8x;
        bi.n_tile = plan->n_tile;
    }
    int rc;
    const bool pdl = after_kernel && pdl_enabled();
    if (plan->fp8x && plan->halo) {
        rc = plan->pair ? launch_halo<128, 1, true, true>(plan, p, pdl, st) : launch_halo<128, 1, false, true>(plan, p, pdl, st);
    } else if (plan->fp8x) {
        if (plan->cl != 1) {
            set_error("conv_tc_forward: the per-tap fp8x kernel has no cluster form");
            return MSI_ERR_UNSUPPORTED;
        }
        rc = launch_tc<128, 2, 1>(plan, p, pdl, st);
    } else if (plan->halo && plan->pair) {
        if (plan->n_tile == 128)
            rc = launch_halo<128, 1, true>(plan, p, pdl, st);
        else if (p.T == 3)
            rc = launch_halo<64, 3, true>(plan, p, pdl, st);
        else
            rc = launch_halo<64, 2, true>(plan, p, pdl, st);
    } else if (plan->halo) {
        if (plan->n_tile == 128)
            rc = launch_halo<128, 1, false>(plan, p, pdl, st);
        else if (p.T == 3)
            rc = launch_halo<64, 3, false>(plan, p, pdl, st);
        else if (p.T == 1)
            rc = launch_halo<64, 1, false>(plan, p, pdl, st);
        else
            rc = launch_halo<64, 2, false>(plan, p, pdl, st);
    } else {
        rc = (plan->n_tile == 64) ? launch_tc_nt<64>(plan, p, pdl, st) : launch_tc_nt<128>(plan, p, pdl, st);
    }
    if (rc != MSI_OK) return rc;
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

}  // namespace msi
