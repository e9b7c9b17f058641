// fp32 CUDA-core implicit-GEMM convolution (MSI_CONV_SIMT): the bring-up / cross-check back end
// of the net.  Same operands as the tensor-core path (fp16 hi/lo activations, recombined to ~22
// significand bits; float32 weights in the TF layout), float32 accumulation, and the coord channel
// of coord_conv2d (nets.py:260-270) evaluated directly instead of folded -- so it checks the
// tcgen05 kernel's weight packing, coord fold and tap geometry independently.
//
// Tile: 64 output positions x 64 output channels per 256-thread block, 4x4 per thread, K-step 16.
#include "net_internal.cuh"

namespace msi {

struct SimtParams {
    const __half* hi[2];
    const __half* lo[2];
    int cin[2], cstride[2], nsrc;
    int cin_total;
    int has_coord;            // conv layers: virtual channel cin_total = |sin(lat_row)|
    const float* coord_rows;  // [Hin]
    const float* w;
    const float* bias;        // head only
    int kind;
    int Hin, Win, Hout, Wout, cout;
    int in_stride;            // conv stride (deconv classes: 1)
    int Mh, Mw;               // grid of output positions per class (conv: Hout x Wout; deconv: Hin x Win)
    TapList taps[4];          // per class
    int ncls;
    float* out;
};

__device__ __forceinline__ float simt_load_a(const SimtParams& p, int b, int iy, int ix, int c) {
    if (iy < 0 || iy >= p.Hin || ix < 0 || ix >= p.Win) return 0.f;
    if (c < p.cin_total) {
        int s = 0;
        if (p.nsrc == 2 && c >= p.cin[0]) {
            s = 1;
            c -= p.cin[0];
        }
        const size_t o = (((size_t)b * p.Hin + iy) * p.Win + ix) * p.cstride[s] + c;
        return (__half2float(p.hi[s][o]) + __half2float(p.lo[s][o])) * (1.0f / MSI_ACT_SCALE);
    }
    if (p.has_coord && c == p.cin_total) return __ldg(p.coord_rows + iy);
    return 0.f;
}

__device__ __forceinline__ float simt_load_w(const SimtParams& p, int wtap, int c, int n) {
    const int kc = p.cin_total + (p.has_coord ? 1 : 0);
    if (c >= kc || n >= p.cout) return 0.f;
    if (p.kind == kDeconv) return __ldg(p.w + ((size_t)wtap * p.cout + n) * p.cin_total + c);
    return __ldg(p.w + ((size_t)wtap * kc + c) * p.cout + n);
}

__global__ void __launch_bounds__(256) conv_simt_kernel(SimtParams p) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int cls = blockIdx.z % p.ncls;
    const int b = blockIdx.z / p.ncls;
    const int m0 = blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const int M = p.Mh * p.Mw;
    const TapList& taps = p.taps[cls];
    const int kc = p.cin_total + (p.has_coord ? 1 : 0);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // A-load assignment: pixel = tid / 4, channels (tid % 4) * 4 .. +3
    const int am = m0 + (tid >> 2);
    const int aoy = (am < M) ? am / p.Mw : 0;
    const int aox = (am < M) ? am % p.Mw : 0;
    // W-load assignment: kk = tid / 16, n = (tid % 16) * 4 .. +3
    const int wk = tid >> 4;
    const int wn = (tid & 15) * 4;

    for (int t = 0; t < taps.n; ++t) {
        const int iy = aoy * p.in_stride + taps.dy[t];
        const int ix = aox * p.in_stride + taps.dx[t];
        for (int c0 = 0; c0 < kc; c0 += 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + (tid & 3) * 4 + q;
                As[(tid & 3) * 4 + q][tid >> 2] = (am < M) ? simt_load_a(p, b, iy, ix, c) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) Ws[wk][wn + q] = simt_load_w(p, taps.wtap[t], c0 + wk, n0 + wn + q);
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    const int py = (p.kind == kDeconv) ? (cls >> 1) : 0;
    const int px = (p.kind == kDeconv) ? (cls & 1) : 0;
    const int os = (p.kind == kDeconv) ? 2 : 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int oy = (m / p.Mw) * os + py;
        const int ox = (m % p.Mw) * os + px;
        float* o = p.out + (((size_t)b * p.Hout + oy) * p.Wout + ox) * p.cout;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.cout) continue;
            float v = acc[i][j];
            if (p.kind == kHead) v = tanhf(v + __ldg(p.bias + n));
            o[n] = v;
        }
    }
}

int conv_simt_forward(const LayerPlan& L, const ActBuf* srcs, int B, float* out, cudaStream_t st) {
    if (!L.coord || srcs[0].x_pad != 0) {
        set_error("conv_simt: the wrap-pad net variant is not built for the SIMT back end");
        return MSI_ERR_UNSUPPORTED;
    }
    SimtParams p;
    p.nsrc = L.nsrc;
    for (int s = 0; s < 2; ++s) {
        p.hi[s] = (s < L.nsrc) ? srcs[s].hi : nullptr;
        p.lo[s] = (s < L.nsrc) ? srcs[s].lo : nullptr;
        p.cin[s] = (s < L.nsrc) ? L.cin[s] : 0;
        p.cstride[s] = (s < L.nsrc) ? srcs[s].c_stride : 0;
    }
    p.cin_total = L.cin_total;
    p.has_coord = (L.kind == kConv) ? 1 : 0;
    p.coord_rows = nullptr;
    p.w = L.w_f32;
    p.bias = L.bias;
    p.kind = L.kind;
    p.Hin = L.Hin;
    p.Win = L.Win;
    p.Hout = L.Hout;
    p.Wout = L.Wout;
    p.cout = L.cout;
    p.out = out;
    if (L.kind == kDeconv) {
        p.in_stride = 1;
        p.Mh = L.Hin;
        p.Mw = L.Win;
        p.ncls = 4;
        for (int c = 0; c < 4; ++c) p.taps[c] = deconv_taps(c >> 1, c & 1);
    } else {
        p.in_stride = L.stride;
        p.Mh = L.Hout;
        p.Mw = L.Wout;
        p.ncls = 1;
        p.taps[0] = conv_taps(L);
    }
    // coord rows live right behind the cbias table (see net.cu)
    if (p.has_coord) p.coord_rows = L.cbias + (size_t)L.Hout * 8 * L.cout;
    dim3 grid(ceil_div((long long)p.Mh * p.Mw, 64), ceil_div(L.cout, 64), B * p.ncls);
    conv_simt_kernel<<<grid, 256, 0, st>>>(p);
    MSI_LAUNCH_CHECK();
    return MSI_OK;
}

}  // namespace msi
