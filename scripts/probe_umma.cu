// Hardware probe for the conv kernel design (not product code; run with gpurun):
//   A. does a tcgen05 K-major SWIZZLE_128B shared-memory descriptor accept a start address that is
//      shifted by whole 128-byte rows (not 1024-byte aligned) and a stride-byte-offset that is not a
//      multiple of 1024?  (needed to read the 9 taps of a 3x3 conv from ONE halo tile in smem)
//   B. how many bytes per clock can one SM ingest through TMA from L2?
//   C. issue rate of SS-mode MMAs of N = 64 / 128 / 256 (with and without concurrent TMA ingest)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/probe_umma scripts/probe_umma.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("probe: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
constexpr uint64_t kDescFixed = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    return kDescFixed | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)(base_off & 7u) << 49) |
           (uint64_t)((saddr >> 4) & 0x3FFF);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// ---------------------------------------------------------------------------------------------
// A. descriptor probe
struct DescCase {
    int shift_rows;  // start address = A tile + shift_rows * 128 bytes
    int sbo_rows;    // stride between 8-row groups, in 128-byte rows
    int bo_mode;     // 0: base_offset = 0; 1: base_offset = (start >> 7) & 7
};
constexpr int kARows = 256, kBN = 64, kMaxCases = 16;

__global__ void __launch_bounds__(128, 1)
desc_probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, const DescCase* cases,
                  int ncases, float* out /* [ncases][128][64] */) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_smem;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + kARows * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar_load), 1);
        mbar_init(smem_u32(&bar_mma), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_smem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&bar_load), (kARows + kBN) * 128);
        tma_load_2d(sA, &ta, smem_u32(&bar_load), 0, 0);
        tma_load_2d(sB, &tb, smem_u32(&bar_load), 0, 0);
    }
    mbar_wait(smem_u32(&bar_load), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cs = 0; cs < ncases; ++cs) {
        if (threadIdx.x == 0) {
            const DescCase c = cases[cs];
            const uint32_t start = sA + (uint32_t)c.shift_rows * 128u;
            const uint32_t bo = c.bo_mode ? ((start >> 7) & 7u) : 0u;
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_desc(start + 32u * k, (uint32_t)c.sbo_rows * 128u, bo);
                const uint64_t db = make_desc(sB + 32u * k, 1024u, 0u);
                umma_f16(tmem, da, db, make_idesc(128, kBN), k > 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&bar_mma));
        }
        mbar_wait(smem_u32(&bar_mma), (uint32_t)(cs & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < kBN; c0 += 32) {
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[((size_t)cs * 128 + row) * kBN + c0 + j] = __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------
// B/C. ingest and MMA rate probe.  Warps 0..n_prod-1 (lane 0) = TMA producers, each with its own ring
// of `stages` slots; a slot = ops_per_slot boxes of box_rows x 128 bytes on ONE mbarrier; a producer
// waits for its own full barrier before re-arming a slot (pure delivery, nothing consumes).  Warp 4
// lane 0 = MMA issuer: batches of `batch` SS MMAs (M=128, N=mma_n, K=16), commit per batch, at most
// `inflight` batches outstanding.
struct RateParams {
    int n_prod;        // producer warps (0..4)
    int n_slots;       // slot fills per producer
    int ops_per_slot;  // TMA ops per slot (per expect_tx)
    int box_rows;      // rows of 128 bytes per op (<= 256)
    int stages;
    int region_rows;      // rows of the global tensor each CTA cycles through
    int cta_stride_rows;  // offset between CTAs' regions (0 = everyone reads the same lines)
    int n_batches;     // MMA batches per CTA (0 = none)
    int batch;
    int inflight;      // 1 or 2
    int mma_n;
    int a_shift_rows;  // A descriptor start shifted by this many 128-byte rows
    int a_sbo_rows;    // A descriptor SBO in rows (8 = canonical)
    int alt_n;         // 0 = all MMAs have N = mma_n; else shapes alternate between mma_n and alt_n ...
    int alt_group;     // ... in runs of alt_group MMAs of the same shape
    long long* lat_out;  // block 0 writes: [0] = clocks of (1 MMA + commit + wait), [1] = clocks of one TMA op round trip
};

__global__ void __launch_bounds__(160, 1) rate_probe_kernel(const __grid_constant__ CUtensorMap tm, const RateParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4][8];
    __shared__ __align__(8) uint64_t mma_bar[2];
    __shared__ uint32_t tmem_base_smem;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // zero the MMA operand region (first 56 KB) so that the MMAs run on finite numbers
    for (int i = threadIdx.x; i < 56 * 1024 / 16; i += 160)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + 16u * i), "r"(0u) : "memory");
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; ++w)
            for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&full_bar[w][s]), 1);
        mbar_init(smem_u32(&mma_bar[0]), 1);
        mbar_init(smem_u32(&mma_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_smem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    if (warp < p.n_prod && lane == 0 && p.n_slots > 0) {
        // rings live above the 56 KB operand region
        const uint32_t box_bytes = (uint32_t)p.box_rows * 128u;
        const uint32_t slot_bytes = box_bytes * (uint32_t)p.ops_per_slot;
        const uint32_t ring = base + 56 * 1024 + (uint32_t)warp * slot_bytes * (uint32_t)p.stages;
        int row = 0;
        const int row0 = ((int)blockIdx.x * 4 + warp) * p.cta_stride_rows;
        int stage = 0;
        uint32_t phase = 0;
        if (blockIdx.x == 0 && warp == 0 && p.lat_out) {
            const long long t0 = clock64();
            mbar_expect_tx(smem_u32(&full_bar[0][7]), box_bytes);
            tma_load_2d(ring, &tm, smem_u32(&full_bar[0][7]), 0, row0);
            mbar_wait(smem_u32(&full_bar[0][7]), 0);
            p.lat_out[1] = clock64() - t0;
        }
        for (int i = 0; i < p.n_slots; ++i) {
            const uint32_t fb = smem_u32(&full_bar[warp][stage]);
            if (i >= p.stages) mbar_wait(fb, phase ^ 1u);  // previous fill of this slot has landed
            mbar_expect_tx(fb, slot_bytes);
            for (int o = 0; o < p.ops_per_slot; ++o) {
                tma_load_2d(ring + (uint32_t)stage * slot_bytes + (uint32_t)o * box_bytes, &tm, fb, 0, row0 + row);
                row += p.box_rows;
                if (row + p.box_rows > p.region_rows) row = 0;
            }
            if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        for (int s = 0; s < p.stages && s < p.n_slots; ++s) {
            const int uses = (p.n_slots - s + p.stages - 1) / p.stages;
            mbar_wait(smem_u32(&full_bar[warp][s]), (uint32_t)((uses - 1) & 1));
        }
    } else if (warp == 4 && lane == 0 && p.n_batches > 0) {
        const uint32_t idesc = make_idesc(128, p.mma_n);
        const uint64_t da0 = make_desc(base + (uint32_t)p.a_shift_rows * 128u, (uint32_t)p.a_sbo_rows * 128u, 0u);
        const uint64_t db0 = make_desc(base + 24 * 1024, 1024u, 0u);  // B: up to 256 rows = 32 KB
        if (blockIdx.x == 0 && p.lat_out) {
            const long long t0 = clock64();
            umma_f16(tmem, da0, db0, idesc, 0u);
            umma_commit(smem_u32(&mma_bar[1]));
            mbar_wait(smem_u32(&mma_bar[1]), 0);
            p.lat_out[0] = clock64() - t0;
        } else {
            umma_f16(tmem, da0, db0, idesc, 0u);
            umma_commit(smem_u32(&mma_bar[1]));
            mbar_wait(smem_u32(&mma_bar[1]), 0);
        }
        // from here: batch b commits to mma_bar[b & 1]; mma_bar[1] has completed one phase already
        uint32_t uses[2] = {0u, 1u};
        for (int b = 0; b < p.n_batches; ++b) {
            if (p.alt_n == 0) {
                for (int i = 0; i < p.batch; ++i) {
                    const uint64_t adv = (uint64_t)((i & 3) * 2);
                    umma_f16(tmem, da0 + adv, db0 + adv, idesc, 1u);
                }
            } else {
                const uint32_t idesc2 = make_idesc(128, p.alt_n);
                for (int i = 0; i < p.batch; ++i) {
                    const uint64_t adv = (uint64_t)((i & 3) * 2);
                    const bool second = ((i / p.alt_group) & 1) != 0;
                    umma_f16(tmem, da0 + adv + (second ? 512u : 0u), db0 + adv, second ? idesc2 : idesc, 1u);
                }
            }
            umma_commit(smem_u32(&mma_bar[b & 1]));
            uses[b & 1]++;
            const int wb = (p.inflight >= 2) ? b - 1 : b;  // batch to wait for
            if (wb >= 0) mbar_wait(smem_u32(&mma_bar[wb & 1]), (uses[wb & 1] - 1u) & 1u);
        }
        if (p.inflight >= 2) {
            const int wb = p.n_batches - 1;
            mbar_wait(smem_u32(&mma_bar[wb & 1]), (uses[wb & 1] - 1u) & 1u);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
        printf("no cuTensorMapEncodeTiled\n");
        exit(2);
    }
    return (EncodeTiledFn)p;
}
static CUtensorMap map2d(const void* base, uint64_t rows, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(2);
    }
    return m;
}

static float a_val(int r, int c) {
    if (c == 0) return (float)(r % 16 - 8);  // channels 0, 1 encode the row index: every row is distinct
    if (c == 1) return (float)(r / 16 - 8);
    return (float)(((r * 7 + c * 3) % 13) - 6);
}
static float b_val(int n, int c) { return (float)(((n * 5 + c * 11) % 7) - 3); }

static void run_desc_probe() {
    std::vector<__half> hA((size_t)kARows * 64), hB((size_t)kBN * 64);
    for (int r = 0; r < kARows; ++r)
        for (int c = 0; c < 64; ++c) hA[(size_t)r * 64 + c] = __float2half(a_val(r, c));
    for (int n = 0; n < kBN; ++n)
        for (int c = 0; c < 64; ++c) hB[(size_t)n * 64 + c] = __float2half(b_val(n, c));
    __half *dA, *dB;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap ta = map2d(dA, kARows, kARows), tb = map2d(dB, kBN, kBN);
    std::vector<DescCase> cases = {
        {0, 8, 0}, {1, 8, 0}, {1, 8, 1}, {3, 8, 0}, {3, 8, 1}, {8, 8, 0},  {0, 10, 0}, {0, 10, 1},
        {11, 10, 0}, {11, 10, 1}, {22, 10, 0}, {22, 10, 1}, {26, 12, 0}, {5, 9, 0}, {0, 16, 0}, {2, 16, 0},
    };
    const int nc = (int)cases.size();
    DescCase* dC;
    float* dOut;
    CK(cudaMalloc(&dC, nc * sizeof(DescCase)));
    CK(cudaMemcpy(dC, cases.data(), nc * sizeof(DescCase), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dOut, (size_t)nc * 128 * kBN * 4));
    CK(cudaMemset(dOut, 0, (size_t)nc * 128 * kBN * 4));
    const int smem = (kARows + kBN) * 128 + 1024;
    CK(cudaFuncSetAttribute(desc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    desc_probe_kernel<<<1, 128, smem>>>(ta, tb, dC, nc, dOut);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)nc * 128 * kBN);
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    // reference rows: ref[r][n] = sum_c A[r][c] * B[n][c]
    std::vector<float> ref((size_t)kARows * kBN);
    for (int r = 0; r < kARows; ++r)
        for (int n = 0; n < kBN; ++n) {
            float s = 0;
            for (int c = 0; c < 64; ++c) s += a_val(r, c) * b_val(n, c);
            ref[(size_t)r * kBN + n] = s;
        }
    printf("== A. descriptor probe (M=128 N=64 K=64, A tile = 256 smem rows, swizzle 128B) ==\n");
    for (int cs = 0; cs < nc; ++cs) {
        int ok = 0, nomatch = 0, other = 0;
        int first_bad = -1, first_bad_src = -2;
        for (int m = 0; m < 128; ++m) {
            const int want = cases[cs].shift_rows + (m / 8) * cases[cs].sbo_rows + (m % 8);
            int found = -1;
            for (int r = 0; r < kARows && found < 0; ++r)
                if (memcmp(&out[((size_t)cs * 128 + m) * kBN], &ref[(size_t)r * kBN], kBN * 4) == 0) found = r;
            if (want < kARows && found == want)
                ++ok;
            else {
                if (found < 0) ++nomatch; else ++other;
                if (first_bad < 0) {
                    first_bad = m;
                    first_bad_src = found;
                }
            }
        }
        printf("case %2d shift=%2d sbo_rows=%2d bo_mode=%d : rows ok %3d/128, matched another row %3d, no match %3d",
               cs, cases[cs].shift_rows, cases[cs].sbo_rows, cases[cs].bo_mode, ok, other, nomatch);
        if (first_bad >= 0) printf("  (first bad m=%d reads smem row %d)", first_bad, first_bad_src);
        printf("\n");
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(dOut);
}

static float time_rate(const CUtensorMap& tm, const RateParams& p, int grid, int smem, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    rate_probe_kernel<<<grid, 160, smem>>>(tm, p);  // warm
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) rate_probe_kernel<<<grid, 160, smem>>>(tm, p);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / reps;
}

static void run_rate_probe() {
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const uint64_t rows = 1u << 21;  // 256 MB tensor
    __half* dT;
    CK(cudaMalloc(&dT, rows * 128));
    CK(cudaMemset(dT, 0, rows * 128));
    long long* dLat;
    CK(cudaMalloc(&dLat, 16));
    CK(cudaMemset(dLat, 0, 16));
    const int smem = 224 * 1024;
    CK(cudaFuncSetAttribute(rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    printf("== B. TMA ingest per SM (grid %d, clock %.3f GHz nominal); own 256 KB region per producer (L2 hits) ==\n", sms, ghz);
    const int boxes[] = {256, 128, 64, 32, 16};
    for (int bi = 0; bi < 5; ++bi)
        for (int n_prod = 1; n_prod <= 4; n_prod *= 2)
            for (int ops = 1; ops <= 4; ops *= 2) {
                RateParams p = {};
                p.n_prod = n_prod;
                p.ops_per_slot = ops;
                p.box_rows = boxes[bi];
                p.stages = 3;
                const long long slot = (long long)p.box_rows * 128 * ops;
                if (slot * p.stages * n_prod > (224 - 56 - 2) * 1024) continue;
                p.n_slots = (int)((16ll << 20) / slot / n_prod);  // 16 MB per CTA
                p.region_rows = 2048;
                p.cta_stride_rows = 2048;
                p.lat_out = dLat;
                CUtensorMap tm = map2d(dT, rows, p.box_rows);
                const float ms = time_rate(tm, p, sms, smem, 3);
                const double bytes = (double)p.n_slots * slot * n_prod;
                const double clk = ms * 1e-3 * ghz * 1e9;
                long long lat[2];
                CK(cudaMemcpy(lat, dLat, 16, cudaMemcpyDeviceToHost));
                printf("box %3d rows (%4.1f KB) producers %d ops/barrier %d: %.3f ms  %.1f B/clk/SM  %.0f clk/op/SM  total %.2f TB/s  (1-op round trip %lld clk)\n",
                       p.box_rows, p.box_rows / 8.0, n_prod, ops, ms, bytes / clk, clk / ((double)p.n_slots * ops * n_prod),
                       bytes * sms / ms * 1e-9, lat[1]);
            }
    printf("== C. SS MMA rate (M=128, K=16) ==\n");
    const int ns[] = {64, 128, 256};
    for (int ni = 0; ni < 3; ++ni)
        for (int batch = 4; batch <= 64; batch *= 4)
            for (int inflight = 1; inflight <= 2; ++inflight)
                for (int shifted = 0; shifted < 2; ++shifted) {
                    if (shifted && !(batch == 16 && inflight == 2)) continue;
                    RateParams p = {};
                    p.n_batches = 32768 / batch;
                    p.batch = batch;
                    p.inflight = inflight;
                    p.mma_n = ns[ni];
                    p.a_shift_rows = shifted ? 11 : 0;
                    p.a_sbo_rows = shifted ? 10 : 8;
                    p.lat_out = dLat;
                    CUtensorMap tm = map2d(dT, rows, 256);
                    const float ms = time_rate(tm, p, sms, smem, 3);
                    const double clk = ms * 1e-3 * ghz * 1e9;
                    long long lat[2];
                    CK(cudaMemcpy(lat, dLat, 16, cudaMemcpyDeviceToHost));
                    printf("N=%3d batch %2d inflight %d %s: %.3f ms, %.1f clk/MMA (ideal %d), %.0f TFLOP/s chip  (1 MMA+commit+wait %lld clk)\n",
                           p.mma_n, batch, inflight, shifted ? "A shifted 11 rows, SBO 10 rows" : "canonical A", ms,
                           clk / 32768.0, p.mma_n / 2, 2.0 * 128 * p.mma_n * 16 * 32768.0 * sms / (ms * 1e-3) * 1e-12, lat[0]);
                }
    printf("== C2. alternating shapes N=256 / N=128 (ideal average 96 clk/MMA) ==\n");
    for (int batch = 8; batch <= 32; batch *= 2)
        for (int group = 1; group <= batch / 2; group *= 2) {
            RateParams p = {};
            p.n_batches = 32768 / batch;
            p.batch = batch;
            p.inflight = 2;
            p.mma_n = 256;
            p.alt_n = 128;
            p.alt_group = group;
            p.a_sbo_rows = 8;
            p.lat_out = dLat;
            CUtensorMap tm = map2d(dT, rows, 256);
            const float ms = time_rate(tm, p, sms, smem, 3);
            const double clk = ms * 1e-3 * ghz * 1e9;
            printf("batch %2d, runs of %2d same-shape MMAs: %.3f ms, %.1f clk/MMA\n", batch, group, ms, clk / 32768.0);
        }
    printf("== D. MMA (N=256, batch 16, 2 in flight) with concurrent TMA ingest ==\n");
    for (int cfg = 0; cfg < 4; ++cfg) {
        RateParams p = {};
        p.n_batches = 32768 / 16;
        p.batch = 16;
        p.inflight = 2;
        p.mma_n = 256;
        p.a_sbo_rows = 8;
        p.n_prod = (cfg & 1) ? 2 : 1;
        p.ops_per_slot = 1;
        p.box_rows = (cfg & 2) ? 256 : 128;
        p.stages = 2;
        p.region_rows = 2048;
        p.cta_stride_rows = 2048;
        const double mma_clk = 32768.0 * 128;
        p.n_slots = (int)(mma_clk * 64.0 / (p.box_rows * 128) / p.n_prod);  // offer 64 B/clk
        p.lat_out = dLat;
        CUtensorMap tm = map2d(dT, rows, p.box_rows);
        const float ms = time_rate(tm, p, sms, smem, 3);
        const double clk = ms * 1e-3 * ghz * 1e9;
        printf("producers %d box %d rows: %.3f ms, %.1f clk/MMA (ideal 128), ingest offered/achieved >= %.1f B/clk/SM\n", p.n_prod,
               p.box_rows, ms, clk / 32768.0, (double)p.n_slots * p.n_prod * p.box_rows * 128 / clk);
    }
    cudaFree(dT);
    cudaFree(dLat);
}

int main(int argc, char** argv) {
    CK(cudaSetDevice(0));
    run_desc_probe();
    if (argc > 1 && !strcmp(argv[1], "--desc-only")) return 0;
    run_rate_probe();
    return 0;
}
