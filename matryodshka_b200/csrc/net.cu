// Host orchestration of the conv net (nets.msi_coord_train_net, nets.py:471-515) behind the
// msi_net_* C ABI: layer table, workspace / arena carving, parameter packing, and the launch
// sequence  conv -> LayerNorm(stats, finalize, normalise+ReLU+split)  x17  -> 1x1 head + tanh.
// All work is enqueued on the caller's stream without host synchronisation, so one forward can be
// captured into a CUDA graph.
#include <math.h>
#include <string.h>

#include <string>

#include "net_internal.cuh"

using namespace msi;

struct msi_net {
    int H, W, c_in, c_out, ngf, max_batch, conv_impl, precision;
    int variant = MSI_NET_COORD;  // MSI_NET_WRAP: nets.msi_train_net (wrap_pad, no coord channel)
    int in_c_stride;
    std::vector<LayerPlan> layers;
    std::vector<ActBuf> acts;  // acts[0] = network input; acts[i + 1] = output of layer i
    std::vector<std::vector<float>> coord_rows_host;
    size_t ws_bytes = 0, arena_bytes = 0;
    // carve-out offsets (filled at create, turned into pointers at bind)
    struct Off {
        size_t w_f32, gamma, beta, bias, cbias, w_hi, w_lo;  // arena
        size_t raw, partials, counter, stats, act_hi, act_lo, act_q8;  // workspace
        bool need_lo, need_q8;  // which residual formats the consumers of this layer's activation read
    };
    std::vector<Off> off;
    size_t in_hi_off = 0, in_lo_off = 0;
    size_t stat_begin = 0, stat_end = 0;  // [partials | counter] of every layer: one memset per forward
    char* ws = nullptr;
    char* arena = nullptr;
    bool bound = false;
    const void* bound_in_hi = nullptr;
};

namespace {

struct ArchRow {
    const char* scope;
    int kind, mult, k, stride, rate, nsrc;
    const char* src[2];
};

// nets.py:486-515
const ArchRow kArch[] = {
    {"conv1_1", kConv, 1, 3, 1, 1, 1, {"input", nullptr}},
    {"conv1_2", kConv, 2, 3, 2, 1, 1, {"conv1_1", nullptr}},
    {"conv2_1", kConv, 2, 3, 1, 1, 1, {"conv1_2", nullptr}},
    {"conv2_2", kConv, 4, 3, 2, 1, 1, {"conv2_1", nullptr}},
    {"conv3_1", kConv, 4, 3, 1, 1, 1, {"conv2_2", nullptr}},
    {"conv3_2", kConv, 4, 3, 1, 1, 1, {"conv3_1", nullptr}},
    {"conv3_3", kConv, 8, 3, 2, 1, 1, {"conv3_2", nullptr}},
    {"conv4_1", kConv, 8, 3, 1, 2, 1, {"conv3_3", nullptr}},
    {"conv4_2", kConv, 8, 3, 1, 2, 1, {"conv4_1", nullptr}},
    {"conv4_3", kConv, 8, 3, 1, 2, 1, {"conv4_2", nullptr}},
    {"conv6_1", kDeconv, 4, 4, 2, 1, 2, {"conv4_3", "conv3_3"}},
    {"conv6_2", kConv, 4, 3, 1, 1, 1, {"conv6_1", nullptr}},
    {"conv6_3", kConv, 4, 3, 1, 1, 1, {"conv6_2", nullptr}},
    {"conv7_1", kDeconv, 2, 4, 2, 1, 2, {"conv6_3", "conv2_2"}},
    {"conv7_2", kConv, 2, 3, 1, 1, 1, {"conv7_1", nullptr}},
    {"conv8_1", kDeconv, 1, 4, 2, 1, 2, {"conv7_2", "conv1_2"}},
    {"conv8_2", kConv, 1, 3, 1, 1, 1, {"conv8_1", nullptr}},
    {"color_pred", kHead, 0, 1, 1, 1, 1, {"conv8_2", nullptr}},
};
const int kNumLayers = sizeof(kArch) / sizeof(kArch[0]);

size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

// [TF-1.14] SAME padding before
int same_pad_before(int n, int k, int s, int rate) {
    const int k_eff = (k - 1) * rate + 1;
    const int out = (n + s - 1) / s;
    int total = (out - 1) * s + k_eff - n;
    if (total < 0) total = 0;
    return total / 2;
}

int find_act(const msi_net* net, const char* scope) {
    if (strcmp(scope, "input") == 0) return 0;
    for (int i = 0; i < (int)net->layers.size(); ++i)
        if (strcmp(net->layers[i].scope, scope) == 0) return i + 1;
    return -1;
}

}  // namespace

extern "C" int msi_net_create(msi_net** out, int H, int W, int c_in, int c_out, int ngf, int max_batch,
                              int conv_impl, int precision) {
    return msi_net_create_ex(out, H, W, c_in, c_out, ngf, max_batch, conv_impl, precision, MSI_NET_COORD);
}

extern "C" int msi_net_create_ex(msi_net** out, int H, int W, int c_in, int c_out, int ngf, int max_batch,
                                 int conv_impl, int precision, int variant) {
    MSI_CHECK_ARG(out != nullptr, "net_create: null out");
    MSI_CHECK_ARG(variant == MSI_NET_COORD || variant == MSI_NET_WRAP, "net_create: variant=%d", variant);
    if (variant == MSI_NET_WRAP && conv_impl != MSI_CONV_TCGEN05) {
        set_error("net_create: MSI_NET_WRAP (msi_train_net) is built for the tcgen05 back end only");
        return MSI_ERR_UNSUPPORTED;
    }
    MSI_CHECK_ARG(H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "net_create: H=%d W=%d must be multiples of 8", H, W);
    MSI_CHECK_ARG(c_in > 0 && c_out > 0 && c_out % 4 == 0, "net_create: bad channels c_in=%d c_out=%d", c_in, c_out);
    MSI_CHECK_ARG(ngf >= 8 && ngf % 8 == 0, "net_create: ngf=%d must be a multiple of 8", ngf);
    MSI_CHECK_ARG(max_batch >= 1, "net_create: max_batch=%d", max_batch);
    MSI_CHECK_ARG(conv_impl == MSI_CONV_TCGEN05 || conv_impl == MSI_CONV_SIMT, "net_create: conv_impl=%d", conv_impl);
    MSI_CHECK_ARG(precision == MSI_PREC_FP16X3 || precision == MSI_PREC_FP16 || precision == MSI_PREC_FP16_FP8X,
                  "net_create: precision=%d", precision);
    if (conv_impl == MSI_CONV_TCGEN05) {
        if (ngf % 64 != 0 || c_out % 16 != 0 || c_out > 256) {
            set_error("net_create: the tcgen05 back end needs ngf %% 64 == 0 and c_out %% 16 == 0, c_out <= 256 (got ngf=%d c_out=%d)", ngf, c_out);
            return MSI_ERR_UNSUPPORTED;
        }
    }
    msi_net* net = new msi_net();
    net->H = H;
    net->W = W;
    net->c_in = c_in;
    net->c_out = c_out;
    net->ngf = ngf;
    net->max_batch = max_batch;
    net->conv_impl = conv_impl;
    net->precision = precision;
    net->variant = variant;
    net->in_c_stride = (int)align_up((size_t)c_in, 64);
    // wrap_pad needs 1 column for the 3x3 / deconv taps and 2 for the dilated convs: 2 everywhere
    const int x_pad = (variant == MSI_NET_WRAP) ? 2 : 0;

    net->acts.resize(kNumLayers + 1);
    net->acts[0].H = H;
    net->acts[0].W = W;
    net->acts[0].C = c_in;
    net->acts[0].c_stride = net->in_c_stride;
    net->acts[0].x_pad = x_pad;
    net->acts[0].Wp = W + 2 * x_pad;
    net->layers.resize(kNumLayers);
    net->off.resize(kNumLayers);
    net->coord_rows_host.resize(kNumLayers);

    // MSI_PREC_FP16_FP8X: the layers with an N tile of 128 form their cross terms in e4m3 and read the e4m3 copy of
    // their inputs; the others (Cout = 64, the head) read the fp16 residual.  An activation stores what its consumers read.
    std::vector<bool> layer_fp8x(kNumLayers, false), need_lo(kNumLayers + 1, false), need_q8(kNumLayers + 1, false);
    {
        auto act_of = [&](const char* scope) {
            if (strcmp(scope, "input") == 0) return 0;
            for (int j = 0; j < kNumLayers; ++j)
                if (strcmp(kArch[j].scope, scope) == 0) return j + 1;
            return -1;
        };
        for (int i = 0; i < kNumLayers; ++i) {
            const int cout = (kArch[i].kind == kHead) ? c_out : ngf * kArch[i].mult;
            layer_fp8x[i] = conv_impl == MSI_CONV_TCGEN05 && precision == MSI_PREC_FP16_FP8X && kArch[i].kind != kHead &&
                            cout % 128 == 0;
            for (int s = 0; s < kArch[i].nsrc; ++s) {
                const int ai = act_of(kArch[i].src[s]);
                (layer_fp8x[i] ? need_q8 : need_lo)[ai] = true;
            }
        }
    }
    size_t a = 0, w = 0;
    const size_t B = (size_t)max_batch;
    net->in_hi_off = w;
    w += align_up(B * H * (W + 2 * x_pad) * net->in_c_stride * sizeof(__half));
    net->in_lo_off = w;
    w += align_up(B * H * (W + 2 * x_pad) * net->in_c_stride * sizeof(__half));

    for (int i = 0; i < kNumLayers; ++i) {
        const ArchRow& r = kArch[i];
        LayerPlan& L = net->layers[i];
        memset(L.scope, 0, sizeof(L.scope));
        strncpy(L.scope, r.scope, sizeof(L.scope) - 1);
        L.kind = r.kind;
        L.k = r.k;
        L.stride = r.stride;
        L.rate = r.rate;
        L.nsrc = r.nsrc;
        L.cin_total = 0;
        for (int s = 0; s < r.nsrc; ++s) {
            L.src[s] = find_act(net, r.src[s]);
            L.cin[s] = net->acts[L.src[s]].C;
            L.cin_total += L.cin[s];
        }
        if (r.nsrc == 1) {
            L.src[1] = -1;
            L.cin[1] = 0;
        }
        const ActBuf& in0 = net->acts[L.src[0]];
        L.Hin = in0.H;
        L.Win = in0.W;
        L.cout = (r.kind == kHead) ? c_out : ngf * r.mult;
        if (r.kind == kDeconv) {
            L.Hout = L.Hin * 2;
            L.Wout = L.Win * 2;
            L.pad_t = L.pad_l = 0;
            L.ncls = 4;
        } else {
            L.Hout = (L.Hin + r.stride - 1) / r.stride;
            L.Wout = (L.Win + r.stride - 1) / r.stride;
            L.pad_t = same_pad_before(L.Hin, r.k, r.stride, r.rate);
            L.pad_l = same_pad_before(L.Win, r.k, r.stride, r.rate);
            if (variant == MSI_NET_WRAP && r.kind == kConv) {
                // wrap_pad(x, rate, rate) + VALID (nets.py:404-423): `rate` rows / columns before, whatever the stride
                L.pad_t = L.pad_l = r.rate;
                L.Hout = (L.Hin + 2 * r.rate - ((r.k - 1) * r.rate + 1)) / r.stride + 1;
                L.Wout = (L.Win + 2 * r.rate - ((r.k - 1) * r.rate + 1)) / r.stride + 1;
            }
            L.ncls = 1;
        }
        L.coord = (variant == MSI_NET_COORD);
        L.fp8x = layer_fp8x[i];
        L.out_act = i + 1;
        ActBuf& o = net->acts[i + 1];
        o.H = L.Hout;
        o.W = L.Wout;
        o.C = L.cout;
        o.c_stride = L.cout;
        o.x_pad = x_pad;
        o.Wp = L.Wout + 2 * x_pad;

        // packed reduction length: taps x (sum of source channel strides)
        int cs_total = 0;
        for (int s = 0; s < L.nsrc; ++s) cs_total += net->acts[L.src[s]].c_stride;
        const int ntaps = (r.kind == kDeconv) ? 4 : r.k * r.k;
        L.K = ntaps * cs_total;

        msi_net::Off& f = net->off[i];
        size_t wcount;
        if (r.kind == kConv)
            wcount = (size_t)r.k * r.k * (L.cin_total + (L.coord ? 1 : 0)) * L.cout;
        else if (r.kind == kDeconv)
            wcount = (size_t)r.k * r.k * L.cout * L.cin_total;
        else
            wcount = (size_t)L.cin_total * L.cout;
        f.w_f32 = a;
        a += align_up(wcount * sizeof(float));
        f.gamma = a;
        a += align_up(L.cout * sizeof(float));
        f.beta = a;
        a += align_up(L.cout * sizeof(float));
        f.bias = a;
        a += align_up(L.cout * sizeof(float));
        f.cbias = a;
        if (r.kind == kConv) a += align_up(((size_t)L.Hout * 8 * L.cout + L.Hin) * sizeof(float));
        f.w_hi = a;
        a += align_up((size_t)L.ncls * L.cout * L.K * sizeof(__half));
        f.w_lo = a;
        a += align_up((size_t)L.ncls * L.cout * L.K * sizeof(__half));

        const size_t n_per = (size_t)L.Hout * L.Wout * L.cout;
        L.n_partials = ln_partials_count((long long)n_per);
        if (L.n_partials < kMaxPersistentCtas * 8) L.n_partials = kMaxPersistentCtas * 8;
        f.raw = w;
        if (r.kind != kHead) w += align_up(B * n_per * sizeof(float));
        f.stats = w;
        w += align_up(B * sizeof(float2));
        const size_t n_act = (size_t)L.Hout * o.Wp * L.cout;  // wrap-padded rows
        f.act_hi = w;
        if (r.kind != kHead) w += align_up(B * n_act * sizeof(__half));
        f.need_lo = need_lo[i + 1] || !need_q8[i + 1];   // (an activation nobody reads in e4m3 keeps the fp16 pair)
        f.need_q8 = need_q8[i + 1];
        f.act_lo = w;
        if (r.kind != kHead && f.need_lo) w += align_up(B * n_act * sizeof(__half));
        f.act_q8 = w;
        if (r.kind != kHead && f.need_q8) w += align_up(B * n_act * sizeof(__half));

        if (r.kind == kConv) {
            // nets.py:262-263: |sin(linspace(-pi/2, pi/2, H))| in float64, cast to float32
            std::vector<float>& rows = net->coord_rows_host[i];
            rows.resize(L.Hin);
            const double pi = 3.141592653589793;
            const double start = -pi / 2.0, stop = pi / 2.0;
            const double step = (L.Hin > 1) ? (stop - start) / (double)(L.Hin - 1) : 0.0;
            for (int h = 0; h < L.Hin; ++h) {
                const double lat = (h == L.Hin - 1 && L.Hin > 1) ? stop : start + step * (double)h;
                rows[h] = (float)fabs(sin(lat));
            }
        }
    }
    net->stat_begin = w;
    for (int i = 0; i < kNumLayers; ++i) {
        msi_net::Off& f = net->off[i];
        f.partials = w;
        w += align_up(B * net->layers[i].n_partials * sizeof(double2));
        f.counter = w;
        w += align_up(sizeof(unsigned int));
    }
    net->stat_end = w;
    net->ws_bytes = w;
    net->arena_bytes = a;
    *out = net;
    return MSI_OK;
}

extern "C" void msi_net_destroy(msi_net* net) {
    if (!net) return;
    for (auto& L : net->layers) conv_tc_plan_destroy(L);
    delete net;
}

extern "C" size_t msi_net_workspace_bytes(const msi_net* net) { return net ? net->ws_bytes : 0; }
extern "C" size_t msi_net_arena_bytes(const msi_net* net) { return net ? net->arena_bytes : 0; }
extern "C" int msi_net_input_c_stride(const msi_net* net) { return net ? net->in_c_stride : 0; }

extern "C" int msi_net_num_launches_per_forward(const msi_net* net) {
    if (!net) return 0;
    // tcgen05: conv (statistics fused) + normalise per layer, + head;  SIMT: conv + 3 LayerNorm kernels
    return (kNumLayers - 1) * (net->conv_impl == MSI_CONV_TCGEN05 ? 2 : 4) + 1;
}

extern "C" int msi_net_bind(msi_net* net, void* workspace, size_t workspace_bytes, void* arena, size_t arena_bytes) {
    MSI_CHECK_ARG(net && workspace && arena, "net_bind: null pointer");
    MSI_CHECK_ARG(workspace_bytes >= net->ws_bytes, "net_bind: workspace %zu < %zu", workspace_bytes, net->ws_bytes);
    MSI_CHECK_ARG(arena_bytes >= net->arena_bytes, "net_bind: arena %zu < %zu", arena_bytes, net->arena_bytes);
    MSI_CHECK_ARG(((uintptr_t)workspace % 1024) == 0 && ((uintptr_t)arena % 1024) == 0,
                  "net_bind: workspace and arena must be 1024-byte aligned");
    net->ws = (char*)workspace;
    net->arena = (char*)arena;
    net->acts[0].hi = (__half*)(net->ws + net->in_hi_off);
    net->acts[0].lo = (__half*)(net->ws + net->in_lo_off);
    for (int i = 0; i < kNumLayers; ++i) {
        LayerPlan& L = net->layers[i];
        const msi_net::Off& f = net->off[i];
        L.w_f32 = (float*)(net->arena + f.w_f32);
        L.gamma = (float*)(net->arena + f.gamma);
        L.beta = (float*)(net->arena + f.beta);
        L.bias = (float*)(net->arena + f.bias);
        L.cbias = (L.kind == kConv && L.coord) ? (float*)(net->arena + f.cbias) : nullptr;
        L.w_hi = (__half*)(net->arena + f.w_hi);
        L.w_lo = (__half*)(net->arena + f.w_lo);
        L.raw = (L.kind != kHead) ? (float*)(net->ws + f.raw) : nullptr;
        L.partials = (double2*)(net->ws + f.partials);
        L.counter = (unsigned int*)(net->ws + f.counter);
        L.stats = (float2*)(net->ws + f.stats);
        if (L.kind != kHead) {
            net->acts[i + 1].hi = (__half*)(net->ws + f.act_hi);
            net->acts[i + 1].lo = f.need_lo ? (__half*)(net->ws + f.act_lo) : nullptr;
            net->acts[i + 1].q8 = f.need_q8 ? (uint8_t*)(net->ws + f.act_q8) : nullptr;
        }
        L.loaded = false;
    }
    net->bound_in_hi = nullptr;
    if (net->conv_impl == MSI_CONV_TCGEN05) {
        for (int i = 0; i < kNumLayers; ++i) {
            LayerPlan& L = net->layers[i];
            ActBuf srcs[2];
            for (int s = 0; s < L.nsrc; ++s) srcs[s] = net->acts[L.src[s]];
            conv_tc_plan_destroy(L);
            int rc = conv_tc_plan_create(L, srcs, net->max_batch, net->precision);
            if (rc != MSI_OK) return rc;
        }
    }
    net->bound = true;
    return MSI_OK;
}

extern "C" int msi_net_load_layer(msi_net* net, const char* scope, const float* weights, const float* gamma,
                                  const float* beta, const float* bias, void* stream) {
    MSI_CHECK_ARG(net && scope && weights, "net_load_layer: null pointer");
    if (!net->bound) {
        set_error("net_load_layer: call msi_net_bind first");
        return MSI_ERR_STATE;
    }
    const int ai = find_act(net, scope);
    MSI_CHECK_ARG(ai >= 1, "net_load_layer: unknown scope '%s'", scope);
    const int i = ai - 1;
    LayerPlan& L = net->layers[i];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t wcount;
    if (L.kind == kConv)
        wcount = (size_t)L.k * L.k * (L.cin_total + (L.coord ? 1 : 0)) * L.cout;
    else if (L.kind == kDeconv)
        wcount = (size_t)L.k * L.k * L.cout * L.cin_total;
    else
        wcount = (size_t)L.cin_total * L.cout;
    MSI_CUDA(cudaMemcpyAsync(L.w_f32, weights, wcount * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (L.kind == kHead) {
        MSI_CHECK_ARG(bias != nullptr, "net_load_layer: '%s' needs a bias", scope);
        MSI_CUDA(cudaMemcpyAsync(L.bias, bias, L.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        MSI_CHECK_ARG(gamma && beta, "net_load_layer: '%s' needs LayerNorm gamma and beta", scope);
        MSI_CUDA(cudaMemcpyAsync(L.gamma, gamma, L.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
        MSI_CUDA(cudaMemcpyAsync(L.beta, beta, L.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (L.kind == kConv && L.coord) {
        float* rows_dev = L.cbias + (size_t)L.Hout * 8 * L.cout;
        MSI_CUDA(cudaMemcpyAsync(rows_dev, net->coord_rows_host[i].data(), L.Hin * sizeof(float),
                                 cudaMemcpyHostToDevice, st));
        int rc = coord_bias_build(L, rows_dev, st);
        if (rc != MSI_OK) return rc;
    }
    if (net->conv_impl == MSI_CONV_TCGEN05) {
        ActBuf srcs[2];
        for (int s = 0; s < L.nsrc; ++s) srcs[s] = net->acts[L.src[s]];
        int rc = conv_tc_pack_weights(L, srcs, st);
        if (rc != MSI_OK) return rc;
    }
    L.loaded = true;
    return MSI_OK;
}

struct ProfCtx;
static int net_forward_impl(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B,
                            float* pred, void* stream, ProfCtx* ev, float* rgba = nullptr);

extern "C" int msi_net_forward(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B,
                               float* pred, void* stream) {
    return net_forward_impl(net, in_f32, in_hi, in_lo, B, pred, stream, nullptr);
}

// The head's epilogue can assemble the RGBA layers itself when the net is the `blend_psv` net of the tensor-core
// back end: c_in = 6P PSV channels, c_out = 2L with L = P, fp16x3 operands, L = 32 or 64 (one N tile).
extern "C" int msi_net_can_fuse_rgba(const msi_net* net) {
    if (!net || !net->bound || net->conv_impl != MSI_CONV_TCGEN05 ||
        (net->precision != MSI_PREC_FP16X3 && net->precision != MSI_PREC_FP16_FP8X))
        return 0;
    if (net->c_in != 3 * net->c_out) return 0;
    return conv_tc_can_fuse_rgba(net->layers[kNumLayers - 1]) ? 1 : 0;
}

extern "C" int msi_net_forward_rgba(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B,
                                    float* rgba, void* stream) {
    MSI_CHECK_ARG(net && rgba, "net_forward_rgba: null pointer");
    if (!msi_net_can_fuse_rgba(net)) {
        set_error("net_forward_rgba: this net cannot fuse the RGBA assembly (needs tcgen05, fp16x3, c_in = 3 c_out, c_out 64 or 128)");
        return MSI_ERR_UNSUPPORTED;
    }
    return net_forward_impl(net, in_f32, in_hi, in_lo, B, nullptr, stream, nullptr, rgba);
}

// Measurement forward: every layer runs once (the real forward), then its conv kernel and its
// LayerNorm kernel are each re-launched kProfReps times back to back between two CUDA events on
// `stream`.  Timing the warm repeats keeps CPU launch latency (which exceeds the duration of the
// small layers' kernels when they are launched one by one) out of the kernel durations.  The
// repeats rewrite identical outputs; the tcgen05 kernel's completion counter is past the grid size
// on a repeat, so it does not re-finalise the LayerNorm statistics.  Synchronises the stream and
// returns per-layer milliseconds per launch (host arrays of msi_net_num_layers() floats; ln_ms of the
// head is 0).  Not capturable into a graph.
static const int kProfReps = 4;
struct ProfCtx {
    std::vector<cudaEvent_t> ev;   // per layer: conv begin / end, LayerNorm begin / end of every repeat
    void* flush = nullptr;         // cold-L2 mode: this buffer (larger than the L2) is overwritten before every timed launch
    size_t flush_bytes = 0;
    cudaEvent_t& at(int layer, int which, int rep) { return ev[((size_t)layer * 4 + which) * kProfReps + rep]; }
};
static int net_forward_prof(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B, float* pred,
                            void* stream, ProfCtx* prof, float* rgba);

extern "C" int msi_net_forward_profiled_flush(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo,
                                              int B, float* pred, void* stream, float* conv_ms_host, float* ln_ms_host,
                                              void* flush, size_t flush_bytes, float* rgba) {
    MSI_CHECK_ARG(conv_ms_host && ln_ms_host, "net_forward_profiled: null output");
    if (rgba != nullptr && !msi_net_can_fuse_rgba(net)) {
        set_error("net_forward_profiled: this net cannot fuse the RGBA assembly");
        return MSI_ERR_UNSUPPORTED;
    }
    ProfCtx prof;
    prof.ev.resize((size_t)kNumLayers * 4 * kProfReps);
    prof.flush = flush;
    prof.flush_bytes = flush ? flush_bytes : 0;
    for (auto& e : prof.ev) MSI_CUDA(cudaEventCreate(&e));
    int rc = net_forward_prof(net, in_f32, in_hi, in_lo, B, pred, stream, &prof, rgba);
    if (rc == MSI_OK) {
        cudaError_t e = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) {
            set_error("net_forward_profiled: %s", cudaGetErrorString(e));
            rc = MSI_ERR_CUDA;
        }
    }
    if (rc == MSI_OK) {
        for (int i = 0; i < kNumLayers; ++i) {
            // warm mode: one event pair around the kProfReps back-to-back launches (rep 0 slots);
            // cold mode: one pair per launch, the flush in between is not timed
            const int pairs = prof.flush ? kProfReps : 1;
            float conv = 0.f, ln = 0.f;
            for (int r = 0; r < pairs; ++r) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, prof.at(i, 0, r), prof.at(i, 1, r));
                conv += ms;
                if (net->layers[i].kind != kHead) {
                    cudaEventElapsedTime(&ms, prof.at(i, 2, r), prof.at(i, 3, r));
                    ln += ms;
                }
            }
            conv_ms_host[i] = conv / kProfReps;
            ln_ms_host[i] = ln / kProfReps;
        }
    }
    for (auto& e : prof.ev) cudaEventDestroy(e);
    return rc;
}

extern "C" int msi_net_forward_profiled(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo,
                                        int B, float* pred, void* stream, float* conv_ms_host, float* ln_ms_host) {
    return msi_net_forward_profiled_flush(net, in_f32, in_hi, in_lo, B, pred, stream, conv_ms_host, ln_ms_host, nullptr, 0,
                                          nullptr);
}

extern "C" int msi_net_num_layers(const msi_net* net) { return net ? kNumLayers : 0; }
extern "C" const char* msi_net_layer_scope(const msi_net* net, int i) {
    return (net && i >= 0 && i < kNumLayers) ? net->layers[i].scope : nullptr;
}
// Algorithmic FLOPs (2 x MACs, coord channels counted) of layer i for one frame.
extern "C" double msi_net_layer_flops(const msi_net* net, int i) {
    if (!net || i < 0 || i >= kNumLayers) return 0.0;
    const LayerPlan& L = net->layers[i];
    if (L.kind == kConv) return 2.0 * L.Hout * L.Wout * L.cout * (double)(L.cin_total + (L.coord ? 1 : 0)) * L.k * L.k;
    if (L.kind == kDeconv) return 2.0 * L.Hin * L.Win * (double)L.cin_total * L.cout * L.k * L.k;
    return 2.0 * L.Hout * L.Wout * (double)L.cin_total * L.cout;
}

static int net_forward_prof(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B, float* pred,
                            void* stream, ProfCtx* prof, float* rgba) {
    return net_forward_impl(net, in_f32, in_hi, in_lo, B, pred, stream, prof, rgba);
}

static int net_forward_impl(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B,
                            float* pred, void* stream, ProfCtx* ev, float* rgba) {
    MSI_CHECK_ARG(net && (pred || rgba), "net_forward: null pointer");
    MSI_CHECK_ARG(B >= 1 && B <= net->max_batch, "net_forward: B=%d outside [1, %d]", B, net->max_batch);
    MSI_CHECK_ARG(in_f32 || (in_hi && in_lo), "net_forward: need in_f32 or the hi/lo pair");
    if (!net->bound) {
        set_error("net_forward: call msi_net_bind first");
        return MSI_ERR_STATE;
    }
    for (auto& L : net->layers)
        if (!L.loaded) {
            set_error("net_forward: layer '%s' has no parameters loaded", L.scope);
            return MSI_ERR_STATE;
        }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    const size_t in_elems = (size_t)B * net->H * net->W * net->in_c_stride;
    if (in_f32) {
        rc = split_input(in_f32, (long long)B * net->H * net->W, net->c_in, net->in_c_stride, net->acts[0].hi,
                         net->acts[0].lo, net->W, net->acts[0].x_pad, st);
        if (rc != MSI_OK) return rc;
    } else if (net->acts[0].x_pad > 0) {
        // the caller's operand is dense [B,H,W,c]; the workspace copy carries the wrap columns
        MSI_CHECK_ARG(in_hi != net->acts[0].hi, "net_forward: the wrap-padded input buffer cannot be the operand itself");
        rc = wrap_copy(reinterpret_cast<const __half*>(in_hi), reinterpret_cast<const __half*>(in_lo),
                       (long long)B * net->H, net->W, net->in_c_stride, net->acts[0].x_pad, net->acts[0].hi,
                       net->acts[0].lo, st);
        if (rc != MSI_OK) return rc;
    } else if (in_hi != net->acts[0].hi) {
        // keep the TMA descriptors pointing at the workspace copy of the input
        MSI_CUDA(cudaMemcpyAsync(net->acts[0].hi, in_hi, in_elems * sizeof(__half), cudaMemcpyDeviceToDevice, st));
        MSI_CUDA(cudaMemcpyAsync(net->acts[0].lo, in_lo, in_elems * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    }
    const bool tc = net->conv_impl == MSI_CONV_TCGEN05;
    if (tc)  // zero the LayerNorm partial sums and CTA counters of every layer
        MSI_CUDA(cudaMemsetAsync(net->ws + net->stat_begin, 0, net->stat_end - net->stat_begin, st));
    for (int i = 0; i < kNumLayers; ++i) {
        LayerPlan& L = net->layers[i];
        ActBuf srcs[2];
        for (int s = 0; s < L.nsrc; ++s) srcs[s] = net->acts[L.src[s]];
        float* out = (L.kind == kHead) ? pred : L.raw;
        const int conv_runs = ev ? 1 + kProfReps : 1;
        const bool cold = ev && ev->flush != nullptr;
        for (int r = 0; r < conv_runs; ++r) {
            if (cold && r >= 1) MSI_CUDA(cudaMemsetAsync(ev->flush, r, ev->flush_bytes, st));
            if (ev && (r == 1 || (cold && r >= 1))) MSI_CUDA(cudaEventRecord(ev->at(i, 0, cold ? r - 1 : 0), st));
            if (net->conv_impl == MSI_CONV_SIMT) {
                rc = conv_simt_forward(L, srcs, B, out, st);
            } else if (L.kind == kHead && rgba != nullptr) {
                // fused RGBA assembly: the PSV eyes of a pixel come from the net's own input operand
                HeadFuse hf;
                hf.rgba = rgba;
                hf.psv_hi = net->acts[0].hi;
                hf.psv_lo = net->acts[0].lo;
                hf.c_stride = net->acts[0].c_stride;
                hf.Wp = net->acts[0].Wp;
                hf.x_pad = net->acts[0].x_pad;
                rc = conv_tc_forward(L, B, nullptr, /*after_kernel=*/true, st, &hf);
            } else {  // the first conv follows a memset / copy, every later one follows our own LayerNorm kernel
                rc = conv_tc_forward(L, B, out, /*after_kernel=*/(i > 0 || r > 0) && !(cold && r >= 1), st);
            }
            if (rc != MSI_OK) return rc;
            if (cold && r >= 1) MSI_CUDA(cudaEventRecord(ev->at(i, 1, r - 1), st));
        }
        if (ev && !cold) MSI_CUDA(cudaEventRecord(ev->at(i, 1, 0), st));
        if (L.kind != kHead) {
            const long long n_per = (long long)L.Hout * L.Wout * L.cout;
            for (int r = 0; r < conv_runs; ++r) {
                if (cold && r >= 1) MSI_CUDA(cudaMemsetAsync(ev->flush, r, ev->flush_bytes, st));
                if (ev && (r == 1 || (cold && r >= 1))) MSI_CUDA(cudaEventRecord(ev->at(i, 2, cold ? r - 1 : 0), st));
                rc = ln_forward(L.raw, B, n_per, L.cout, L.gamma, L.beta, L.partials, L.n_partials, L.stats,
                                net->acts[i + 1].hi, net->acts[i + 1].lo, net->acts[i + 1].q8, /*stats_ready=*/tc,
                                /*pdl=*/tc && !(cold && r >= 1), L.Wout, net->acts[i + 1].x_pad, st);
                if (rc != MSI_OK) return rc;
                if (cold && r >= 1) MSI_CUDA(cudaEventRecord(ev->at(i, 3, r - 1), st));
            }
            if (ev && !cold) MSI_CUDA(cudaEventRecord(ev->at(i, 3, 0), st));
        }
    }
    return MSI_OK;
}

// Pointers of the workspace copy of the network input, so that msi_psv_build can write the PSV
// operand in place (no copy in msi_net_forward).
extern "C" int msi_net_input_buffers(msi_net* net, void** hi, void** lo) {
    MSI_CHECK_ARG(net && hi && lo, "net_input_buffers: null pointer");
    if (!net->bound) {
        set_error("net_input_buffers: call msi_net_bind first");
        return MSI_ERR_STATE;
    }
    if (net->acts[0].x_pad > 0) {
        set_error("net_input_buffers: the MSI_NET_WRAP input is stored wrap-padded; pass a dense operand to msi_net_forward");
        return MSI_ERR_UNSUPPORTED;
    }
    *hi = net->acts[0].hi;
    *lo = net->acts[0].lo;
    return MSI_OK;
}

extern "C" int msi_net_read_activation(msi_net* net, const char* scope, int B, float* out, void* stream) {
    MSI_CHECK_ARG(net && scope && out, "net_read_activation: null pointer");
    const int ai = find_act(net, scope);
    MSI_CHECK_ARG(ai >= 0 && ai < kNumLayers, "net_read_activation: unknown or un-normalised scope '%s'", scope);
    const ActBuf& a = net->acts[ai];
    return merge_activation(a.hi, a.lo, a.q8, (long long)B * a.H * a.W, a.C, a.c_stride, out, a.W, a.x_pad,
                            reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int msi_net_read_raw(msi_net* net, const char* scope, int B, float* out, void* stream) {
    MSI_CHECK_ARG(net && scope && out, "net_read_raw: null pointer");
    const int ai = find_act(net, scope);
    MSI_CHECK_ARG(ai >= 1 && ai < kNumLayers, "net_read_raw: unknown or un-normalised scope '%s'", scope);
    const LayerPlan& L = net->layers[ai - 1];
    MSI_CUDA(cudaMemcpyAsync(out, L.raw, (size_t)B * L.Hout * L.Wout * L.cout * sizeof(float),
                             cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return MSI_OK;
}
