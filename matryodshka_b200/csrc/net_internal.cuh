// Internal description of the conv net (nets.py:471-515) shared by the host orchestration
// (net.cu) and the conv back ends (conv_simt.cu, conv_tcgen05.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace msi {

enum LayerKind { kConv = 0, kDeconv = 1, kHead = 2 };

// MSI_PREC_FP16_FP8X.  x * w ~= x_hi * w_hi (fp16 MMA) + x_hi8 * w_lo8 + x_lo8 * w_hi8 (ONE e4m3 MMA of twice the K):
// the cross terms are ~2^-11 of the main term, so 4 significand bits carry them to ~2^-15 of the product.  The scales
// pair up so that every product has the scale MSI_ACT_SCALE * MSI_WEIGHT_SCALE of the main term:
//   x_hi8 = e4m3(x_hi * 2^-kFp8HiShift)   w_lo8 = e4m3(w_lo * 2^+kFp8HiShift)
//   x_lo8 = e4m3(x_lo * 2^+kFp8LoShift)   w_hi8 = e4m3(w_hi * 2^-kFp8LoShift)
// (x_hi = fp16(16 x) up to ~450 for |x| < 28 -> x_hi8 < 112; |x_lo| <= 2^-11 x_hi -> x_lo8 < 56; measured on the net:
// max-abs 2.3e-4 on the prediction at 160 x 320 against 1.5e-5 for fp16x3 and 6.5e-3 for one fp16 pass,
// scripts/exp_fp8_cross.py.)
constexpr int kFp8HiShift = 2;
constexpr int kFp8LoShift = 9;

// persistent conv kernel: at most this many CTAs (B200 has 148 SMs); 4 statistic slots per CTA
constexpr int kMaxPersistentCtas = 160;

// One activation tensor in the workspace: post-LayerNorm+ReLU values scaled by MSI_ACT_SCALE and
// split into fp16 hi + lo, NHWC with channel stride c_stride (>= C, multiple of 64 for tcgen05).
struct ActBuf {
    __half* hi = nullptr;
    __half* lo = nullptr;     // fp16 residual (consumers on the fp16x3 product); may be null when only q8 is consumed
    // MSI_PREC_FP16_FP8X: e4m3 copies for the cross terms, shaped like the fp16 tensor (2 bytes per channel): per pixel
    // and 64-channel chunk, 64 bytes hi8 = e4m3(hi * 2^-kFp8HiShift) then 64 bytes lo8 = e4m3(lo * 2^kFp8LoShift)
    uint8_t* q8 = nullptr;
    int H = 0, W = 0, C = 0, c_stride = 0;
    // MSI_NET_WRAP: rows are stored x_pad pixels wider on both sides (Wp = W + 2 * x_pad) and the pad
    // columns hold the circular wrap of the row, so a TMA box that leaves the image along x reads the
    // other side of the panorama while rows above / below the image are still zero-filled
    int x_pad = 0, Wp = 0;
};

struct LayerPlan {
    char scope[16];
    int kind;
    int k, stride, rate;
    int nsrc;
    int src[2];  // activation indices; -1 = network input
    int cin[2];
    int cin_total;
    int cout;
    int Hin, Win, Hout, Wout;
    int pad_t, pad_l;  // SAME padding before (conv)
    bool coord = true; // the 3x3 convs take an extra |sin(latitude)| input channel (msi_coord_train_net)
    bool fp8x = false; // this layer forms its cross terms in e4m3 (MSI_PREC_FP16_FP8X and an N tile of 128)
    // arena (parameters)
    float* w_f32 = nullptr;   // TF layout as loaded
    float* gamma = nullptr;
    float* beta = nullptr;
    float* bias = nullptr;    // head only
    float* cbias = nullptr;   // coord-channel bias table [Hout][8][cout] (conv only)
    __half* w_hi = nullptr;   // packed K-major [ncls][cout][K], scaled by MSI_WEIGHT_SCALE
    __half* w_lo = nullptr;
    int K = 0;                // packed reduction length = ntaps * sum(c_stride of sources)
    int ncls = 1;             // 4 output-parity classes for deconv
    bool loaded = false;
    // workspace
    float* raw = nullptr;     // [B,Hout,Wout,cout] pre-LayerNorm conv output (head: pred, caller's buffer)
    double2* partials = nullptr;  // [B][n_partials] (sum, sumsq)
    float2* stats = nullptr;      // [B] (mean, rstd)
    unsigned int* counter = nullptr;  // CTA completion counter of the fused statistics (tcgen05 back end)
    int n_partials = 0;
    int out_act = -1;         // index of the activation this layer produces
    void* tc_plan = nullptr;  // back-end private (TMA descriptors, tile shape)
};

// Geometry of one gather tap of the implicit GEMM.
struct TapList {
    int n;
    int dy[9], dx[9];  // input offset added to (o * in_stride)
    int wtap[9];       // index of the (kh,kw) weight slice
};

// conv: input = o*stride + k*rate - pad.  deconv class (py,px): input = o + d, output = 2*o + parity.
inline TapList conv_taps(const LayerPlan& L) {
    TapList t;
    t.n = L.k * L.k;
    for (int kh = 0; kh < L.k; ++kh)
        for (int kw = 0; kw < L.k; ++kw) {
            const int i = kh * L.k + kw;
            t.dy[i] = kh * L.rate - L.pad_t;
            t.dx[i] = kw * L.rate - L.pad_l;
            t.wtap[i] = i;
        }
    return t;
}

// 4x4 stride-2 SAME transposed conv, oy = 2*iy - 1 + kh:
//   oy even (py=0): kh=1 -> iy=o, kh=3 -> iy=o-1;   oy odd (py=1): kh=0 -> iy=o+1, kh=2 -> iy=o
inline void deconv_axis(int parity, int t, int& k, int& d) {
    if (parity == 0) {
        k = (t == 0) ? 1 : 3;
        d = (t == 0) ? 0 : -1;
    } else {
        k = (t == 0) ? 0 : 2;
        d = (t == 0) ? 1 : 0;
    }
}

inline TapList deconv_taps(int py, int px) {
    TapList t;
    t.n = 4;
    for (int th = 0; th < 2; ++th)
        for (int tw = 0; tw < 2; ++tw) {
            int kh, dy, kw, dx;
            deconv_axis(py, th, kh, dy);
            deconv_axis(px, tw, kw, dx);
            const int i = th * 2 + tw;
            t.dy[i] = dy;
            t.dx[i] = dx;
            t.wtap[i] = kh * 4 + kw;
        }
    return t;
}

// back ends -------------------------------------------------------------------------------
int conv_simt_forward(const LayerPlan& L, const ActBuf* srcs, int B, float* out, cudaStream_t st);

int conv_tc_plan_create(LayerPlan& L, const ActBuf* srcs, int max_batch, int precision);
void conv_tc_plan_destroy(LayerPlan& L);
// after_kernel: the previous operation in the stream is one of this library's kernels that calls
// griddepcontrol.wait itself, so this launch may use programmatic dependent launch
// fuse != null (head only, conv_tc_can_fuse_rgba): the epilogue assembles the RGBA layers (msi.py:130-147) from the
// prediction and the PSV operand instead of storing the prediction; `out` is then unused
struct HeadFuse {
    float* rgba;          // [B,H,W,L,4]
    const __half* psv_hi; // the net's input operand [B,H,Wp,c_stride], pixel x at column x + x_pad
    const __half* psv_lo;
    int c_stride, Wp, x_pad;
};
bool conv_tc_can_fuse_rgba(const LayerPlan& L);
int conv_tc_forward(const LayerPlan& L, int B, float* out, bool after_kernel, cudaStream_t st, const HeadFuse* fuse = nullptr);
bool pdl_enabled();
int conv_tc_pack_weights(LayerPlan& L, const ActBuf* srcs, cudaStream_t st);

// LayerNorm ---------------------------------------------------------------------------------
int ln_partials_count(long long n_per_sample);
// out_lo / out_q8: either may be null (the formats the consumers of this activation read)
int ln_forward(const float* raw, int B, long long n_per_sample, int C, const float* gamma, const float* beta,
               double2* partials, int n_partials, float2* stats, __half* out_hi, __half* out_lo, uint8_t* out_q8,
               bool stats_ready, bool pdl, int W, int x_pad, cudaStream_t st);
// W / x_pad: row width and wrap padding of the fp16 tensor (x_pad = 0: dense rows)
int split_input(const float* in, long long npix, int C, int c_stride, __half* hi, __half* lo, int W, int x_pad,
                cudaStream_t st);
// lo == null: the residual is taken from the e4m3 copy q8 (~4 bits of it)
int merge_activation(const __half* hi, const __half* lo, const uint8_t* q8, long long npix, int C, int c_stride, float* out,
                     int W, int x_pad, cudaStream_t st);
// dense fp16 hi/lo [rows, W, c_stride] -> wrap-padded [rows, W + 2 x_pad, c_stride]
int wrap_copy(const __half* in_hi, const __half* in_lo, long long rows, int W, int c_stride, int x_pad, __half* out_hi,
              __half* out_lo, cudaStream_t st);
int coord_bias_build(const LayerPlan& L, const float* coord_rows_dev, cudaStream_t st);

}  // namespace msi
