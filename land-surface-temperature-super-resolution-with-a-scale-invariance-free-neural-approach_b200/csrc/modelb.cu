// ModelB_2.forward (model.py:608-645) and its backward as one host-side launch plan over
// the kernels of this library.  Every convolution stores its RAW (pre-BatchNorm) output;
// BatchNorm + ReLU are applied by whoever consumes it (the next convolution's load stage,
// the pooling / residual / up-sample kernels), so no activation tensor is materialised
// twice and the backward pass recomputes activations from the same raw tensors.
//
// The caller owns the workspace; this file only carves it up.
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace {

struct ConvDesc { int cin, cout, level; };

struct Net {
    ConvDesc conv[SIFNN_MODELB_NCONV];
    int64_t w_off[SIFNN_MODELB_NCONV];
    int64_t gamma_off[SIFNN_MODELB_NBN], beta_off[SIFNN_MODELB_NBN], bn_off[SIFNN_MODELB_NBN];
    int64_t bias_off, n_params, bn_total;
    int d[4];
    bool ok;
};

Net build_net(const sifnn_modelb_cfg* cfg) {
    Net n{};
    n.ok = false;
    if (!cfg) return n;
    const int in = cfg->in_channels, d0 = cfg->down[0], d1 = cfg->down[1], d2 = cfg->down[2], d3 = cfg->down[3];
    if (in <= 0 || d0 <= 0 || d1 != 2 * d0 || d2 != 2 * d1 || d3 != 2 * d2) return n;  // cat([up, skip]) must match UpBlock's in_channels
    const int half = d3 / 2;
    const ConvDesc t[SIFNN_MODELB_NCONV] = {
        {in, d0, 0}, {d0, d0, 0},                       // inbloc
        {d0, d0, 1}, {d0, d0, 1}, {d0, d1, 1},          // db1
        {d1, d1, 2}, {d1, d1, 2}, {d1, d2, 2},          // db2
        {d2, d2, 3}, {d2, d2, 3}, {d2, half, 3},        // db3
        {d3, d3 / 2, 2}, {d3 / 2, d2 / 2, 2},           // ub1
        {d2, d2 / 2, 1}, {d2 / 2, d1 / 2, 1},           // ub2
        {d1, d1 / 2, 0}, {d1 / 2, d0, 0},               // ub3
        {d0, 1, 0},                                     // outlay
    };
    int64_t off = 0, boff = 0;
    for (int i = 0; i < SIFNN_MODELB_NCONV; ++i) {
        n.conv[i] = t[i];
        n.w_off[i] = off;
        off += (int64_t)t[i].cout * t[i].cin * 9;
        if (i < SIFNN_MODELB_NBN) {
            n.gamma_off[i] = off; off += t[i].cout;
            n.beta_off[i] = off; off += t[i].cout;
            n.bn_off[i] = boff; boff += t[i].cout;
        } else {
            n.bias_off = off; off += t[i].cout;
        }
    }
    n.n_params = off;
    n.bn_total = boff;
    n.d[0] = d0; n.d[1] = d1; n.d[2] = d2; n.d[3] = d3;
    n.ok = true;
    return n;
}

// tcgen05 path on by default where the shape is eligible; SIFNN_DISABLE_TC=1 (or sifnn_set_tensor_cores(0)) forces the
// fp32 SIMT kernels everywhere ("strict fp32" mode: no tensor-core accumulation rounding).
int g_tc_mode = -1;
bool tc_enabled() {
    if (g_tc_mode < 0) {
        const char* e = getenv("SIFNN_DISABLE_TC");
        g_tc_mode = (e && e[0] && e[0] != '0') ? 0 : 1;
    }
    return g_tc_mode == 1;
}

// Which convolution kernel serves a layer (K = channels of the tensor that is read, O = channels written):
//   FS  fold + shift (conv3x3_fs.cu): widths that are multiples of 128;  FF  full fold (conv3x3_ff.cu): 32- and 64-pixel-wide levels;
//   TC  round-1 tcgen05 kernels (any other eligible shape, more than 64 input channels);  SIMT  fp32 direct convolution.
enum ConvKind { KIND_SIMT = 0, KIND_TC = 1, KIND_FF = 2, KIND_FS = 3 };
ConvKind conv_kind(int K, int O, int H, int W, bool tc_ok, bool forward = false) {
    if (!tc_enabled()) return KIND_SIMT;
    if (forward ? sifnn::conv3x3_fs_fwd_preferred(K, O, H, W) : sifnn::conv3x3_fs_supported(K, O, H, W)) return KIND_FS;
    if (sifnn::conv3x3_ff_supported(K, O, H, W)) return KIND_FF;
    return tc_ok ? KIND_TC : KIND_SIMT;
}

struct Carver {
    char* base;
    size_t off;
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Workspace {
    float* raw[SIFNN_MODELB_NBN];
    float* P[3];
    float* R[3];
    float* U[3];
    float *scale, *shift, *mean, *invstd;  // bn_total each
    double* stats;                          // 2 * bn_total (layer i at 2*bn_off[i]); the 32 tickets of the fused finalize follow directly (one memset)
    unsigned int* bn_tickets;               // [SIFNN_MODELB_NBN] (rounded up to 32)
    double* bsums;                          // 2 * bn_total
    // training only
    float* g[SIFNN_MODELB_NBN];
    float* gR[3];
    float* gU[3];
    void* wgrad_ws;
    void* wgrad_part[SIFNN_MODELB_NCONV];   // per-layer partial sums of the tensor-core weight gradients (reduced in one launch per backward phase)
    void* wprep_f[SIFNN_MODELB_NCONV];  // hi/lo-split weights per layer, forward layout (tensor-core path; all prepared by ONE launch per pass)
    void* wprep_d[SIFNN_MODELB_NCONV];  // same, data-gradient layout (training only)
    float* wedge_d[SIFNN_MODELB_NCONV]; // fp32 edge taps of the fold + shift data gradient (training only)
    size_t bytes;
};

Workspace carve(const Net& n, void* base, int B, int H, int W, int train) {
    Workspace w{};
    Carver c{static_cast<char*>(base), 0};
    const size_t hw[4] = {(size_t)H * W, (size_t)H * W / 4, (size_t)H * W / 16, (size_t)H * W / 64};
    for (int i = 0; i < SIFNN_MODELB_NBN; ++i) w.raw[i] = c.take<float>((size_t)B * n.conv[i].cout * hw[n.conv[i].level]);
    for (int k = 0; k < 3; ++k) {
        const size_t sz = (size_t)B * n.d[k] * hw[k + 1];
        w.P[k] = c.take<float>(sz);
        w.R[k] = c.take<float>(sz);
    }
    // U1 (d3 @ L2), U2 (d2 @ L1), U3 (d1 @ L0)
    for (int k = 0; k < 3; ++k) w.U[k] = c.take<float>((size_t)B * n.d[3 - k] * hw[2 - k]);
    w.scale = c.take<float>(n.bn_total);
    w.shift = c.take<float>(n.bn_total);
    w.mean = c.take<float>(n.bn_total);
    w.invstd = c.take<float>(n.bn_total);
    w.stats = c.take<double>(2 * n.bn_total);
    w.bn_tickets = c.take<unsigned int>(32);
    w.bsums = c.take<double>(2 * n.bn_total);
    for (int i = 0; i < SIFNN_MODELB_NCONV; ++i)
        w.wprep_f[i] = c.take<char>(sifnn_conv3x3_tc_wprep_bytes((n.conv[i].cin + 7) / 8 * 8, n.conv[i].cout));
    if (train) {
        for (int i = 0; i < SIFNN_MODELB_NCONV; ++i) {
            w.wprep_d[i] = c.take<char>(sifnn_conv3x3_tc_wprep_bytes((n.conv[i].cout + 7) / 8 * 8, n.conv[i].cin));
            w.wedge_d[i] = c.take<float>(sifnn::conv3x3_fs_wedge_bytes(n.conv[i].cout, n.conv[i].cin) / sizeof(float));
        }
        for (int i = 0; i < SIFNN_MODELB_NBN; ++i) w.g[i] = c.take<float>((size_t)B * n.conv[i].cout * hw[n.conv[i].level]);
        for (int k = 0; k < 3; ++k) w.gR[k] = c.take<float>((size_t)B * n.d[k] * hw[k + 1]);
        for (int k = 0; k < 3; ++k) w.gU[k] = c.take<float>((size_t)B * n.d[3 - k] * hw[2 - k]);
        size_t mx = 0;
        for (int i = 0; i < SIFNN_MODELB_NCONV; ++i) {
            const int s = n.conv[i].level;
            const size_t b = sifnn_conv3x3_wgrad_workspace(B, n.conv[i].cin, n.conv[i].cout, H >> s, W >> s);
            const size_t b2 = sifnn_conv3x3_wgrad_tc_workspace(B, n.conv[i].cin, n.conv[i].cout, H >> s, W >> s);
            const size_t b3 = sifnn_conv3x3_wgrad_km_workspace(B, n.conv[i].cin, n.conv[i].cout, H >> s, W >> s);
            if (b > mx) mx = b;
            w.wgrad_part[i] = c.take<char>(b2 > b3 ? b2 : b3);
        }
        w.wgrad_ws = c.take<char>(mx);
    }
    c.off = (c.off + 255) & ~(size_t)255;
    w.bytes = c.off;
    return w;
}

struct EvalTable {
    int64_t gamma_off[SIFNN_MODELB_NBN], beta_off[SIFNN_MODELB_NBN], bn_off[SIFNN_MODELB_NBN + 1];
};

__global__ void bn_eval_affine_all_kernel(const float* __restrict__ params, const float* __restrict__ rm, const float* __restrict__ rv,
                                          float* scale, float* shift, const EvalTable t, int total) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    int l = 0;
#pragma unroll 1
    while (l + 1 < SIFNN_MODELB_NBN && j >= t.bn_off[l + 1]) ++l;
    const int c = j - (int)t.bn_off[l];
    const float invstd = 1.0f / sqrtf(rv[j] + 1e-5f);
    const float sc = params[t.gamma_off[l] + c] * invstd;
    scale[j] = sc;
    shift[j] = fmaf(-rm[j], sc, params[t.beta_off[l] + c]);
}

bool shape_ok(int B, int H, int W) { return B > 0 && B <= 65535 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0; }

}  // namespace

extern "C" void sifnn_set_tensor_cores(int on) { g_tc_mode = on ? 1 : 0; }
extern "C" int sifnn_get_tensor_cores(void) { return tc_enabled() ? 1 : 0; }

extern "C" int64_t sifnn_modelb_param_layout(const sifnn_modelb_cfg* cfg, int64_t* w_off, int64_t* gamma_off, int64_t* beta_off,
                                             int64_t* bias_off, int64_t* bn_off, int64_t* bn_total) {
    const Net n = build_net(cfg);
    if (!n.ok) { sifnn::set_error("modelb: unsupported configuration (need down[k+1] == 2*down[k])"); return -1; }
    if (w_off) memcpy(w_off, n.w_off, sizeof(n.w_off));
    if (gamma_off) memcpy(gamma_off, n.gamma_off, sizeof(n.gamma_off));
    if (beta_off) memcpy(beta_off, n.beta_off, sizeof(n.beta_off));
    if (bias_off) *bias_off = n.bias_off;
    if (bn_off) memcpy(bn_off, n.bn_off, sizeof(n.bn_off));
    if (bn_total) *bn_total = n.bn_total;
    return n.n_params;
}

extern "C" int64_t sifnn_modelb_decoder_offset(const sifnn_modelb_cfg* cfg) {
    const Net n = build_net(cfg);
    return n.ok ? n.w_off[11] : -1;
}

extern "C" size_t sifnn_modelb_workspace_bytes(const sifnn_modelb_cfg* cfg, int B, int H, int W, int train) {
    const Net n = build_net(cfg);
    if (!n.ok || !shape_ok(B, H, W)) return 0;
    return carve(n, nullptr, B, H, W, train).bytes;
}

extern "C" int sifnn_modelb_forward(const sifnn_modelb_cfg* cfg, const float* params, float* running_mean, float* running_var,
                                    const float* x, float* y, void* workspace, int B, int H, int W, int train, sifnn_stream_t stream) {
    const Net n = build_net(cfg);
    SIFNN_REQUIRE(n.ok, "modelb_forward: unsupported configuration (need down[k+1] == 2*down[k])");
    SIFNN_REQUIRE(params && running_mean && running_var && x && y && workspace, "modelb_forward: null pointer");
    SIFNN_REQUIRE(shape_ok(B, H, W), "modelb_forward: need H, W multiples of 8 and 1 <= B <= 65535 (got B=%d H=%d W=%d)", B, H, W);
    const Workspace w = carve(n, workspace, B, H, W, train);
    cudaStream_t st = sifnn::as_stream(stream);
    const int hs[4] = {H, H / 2, H / 4, H / 8}, ws[4] = {W, W / 2, W / 4, W / 8};
    if (train) {
        SIFNN_CUDA(cudaMemsetAsync(w.stats, 0, (size_t)(reinterpret_cast<char*>(w.bn_tickets + 32) - reinterpret_cast<char*>(w.stats)), st));   // sums + tickets
    } else {
        EvalTable t{};
        for (int i = 0; i < SIFNN_MODELB_NBN; ++i) { t.gamma_off[i] = n.gamma_off[i]; t.beta_off[i] = n.beta_off[i]; t.bn_off[i] = n.bn_off[i]; }
        t.bn_off[SIFNN_MODELB_NBN] = n.bn_total;
        bn_eval_affine_all_kernel<<<((int)n.bn_total + 127) / 128, 128, 0, st>>>(params, running_mean, running_var, w.scale, w.shift, t, (int)n.bn_total);
        SIFNN_TRY(sifnn::check_launch("bn_eval_affine_all_kernel"));
    }

    // tensor-core layers: split the weights of all of them up front, one launch per kernel family
    auto fwd_kind = [&](int i) {
        const ConvDesc& c = n.conv[i];
        const bool tc_ok = c.cout <= 64 && sifnn_conv3x3_tc_supported(c.cin, c.cout, hs[c.level], ws[c.level]) != 0;
        return conv_kind(c.cin, c.cout, hs[c.level], ws[c.level], tc_ok, true);
    };
    {
        sifnn::TcPrepJob jobs[SIFNN_MODELB_NCONV];
        const float* pw[2][SIFNN_MODELB_NCONV];
        void* pp[2][SIFNN_MODELB_NCONV];
        int pK[2][SIFNN_MODELB_NCONV], pO[2][SIFNN_MODELB_NCONV], pso[2][SIFNN_MODELB_NCONV], psk[2][SIFNN_MODELB_NCONV], pfl[2][SIFNN_MODELB_NCONV];
        int nj = 0, np[2] = {0, 0};
        for (int i = 0; i < SIFNN_MODELB_NCONV; ++i) {
            const ConvKind k = fwd_kind(i);
            if (k == KIND_TC) jobs[nj++] = sifnn::tc_prep_job_fwd(params + n.w_off[i], w.wprep_f[i], n.conv[i].cin, n.conv[i].cout, ws[n.conv[i].level]);
            if (k == KIND_FF || k == KIND_FS) {
                const int f = (k == KIND_FS) ? 1 : 0, j = np[f]++;
                pw[f][j] = params + n.w_off[i]; pp[f][j] = w.wprep_f[i]; pK[f][j] = n.conv[i].cin; pO[f][j] = n.conv[i].cout;
                pso[f][j] = n.conv[i].cin * 9; psk[f][j] = 9; pfl[f][j] = 0;
            }
        }
        SIFNN_TRY(sifnn::tc_prep_many(jobs, nj, st));
        if (np[0]) SIFNN_TRY(sifnn::ff_prep(pw[0], pp[0], pK[0], pO[0], pso[0], psk[0], pfl[0], np[0], st));
        if (np[1]) SIFNN_TRY(sifnn::fs_prep(pw[1], pp[1], nullptr, pK[1], pO[1], pso[1], psk[1], pfl[1], np[1], st));
    }
    // conv i reading `in` (plain if aff < 0, else BatchNorm+ReLU of layer `aff` applied on load)
    auto conv = [&](int i, const float* in, int aff, float* out) -> int {
        const ConvDesc& c = n.conv[i];
        const int l = c.level;
        const bool bn = i < SIFNN_MODELB_NBN;
        const float* isc = aff >= 0 ? w.scale + n.bn_off[aff] : nullptr;
        const float* ish = aff >= 0 ? w.shift + n.bn_off[aff] : nullptr;
        double* st_ptr = (bn && train) ? w.stats + 2 * n.bn_off[i] : nullptr;
        // BatchNorm finalize: inside the convolution (last CTA) for the fs / ff kernels, a separate launch otherwise
        sifnn::BnTail tail{};
        bool fused_tail = false;
        if (bn && train) {
            const int64_t o = n.bn_off[i];
            tail = sifnn::BnTail{w.bn_tickets + i, params + n.gamma_off[i], params + n.beta_off[i], running_mean + o, running_var + o, w.scale + o, w.shift + o,
                                 w.mean + o, w.invstd + o, (double)B * hs[l] * ws[l]};
        }
        switch (fwd_kind(i)) {
            case KIND_FS:
                SIFNN_TRY(sifnn::conv3x3_fwd_fs_prepped(in, nullptr, 0, isc, ish, w.wprep_f[i], out, st_ptr, 0, B, c.cin, c.cout, hs[l], ws[l], st, st_ptr ? &tail : nullptr));
                fused_tail = st_ptr != nullptr;
                break;
            case KIND_FF:
                SIFNN_TRY(sifnn::conv3x3_fwd_ff_prepped(in, nullptr, 0, isc, ish, w.wprep_f[i], out, st_ptr, 0, B, c.cin, c.cout, hs[l], ws[l], st, st_ptr ? &tail : nullptr));
                fused_tail = st_ptr != nullptr;
                break;
            case KIND_TC:
                SIFNN_TRY(sifnn::conv3x3_fwd_tc_prepped(in, isc, ish, w.wprep_f[i], bn ? nullptr : params + n.bias_off, out, st_ptr, B, c.cin, c.cout, hs[l],
                                                        ws[l], st));
                break;
            default:
                SIFNN_TRY(sifnn_conv3x3_fwd(in, isc, ish, params + n.w_off[i], bn ? nullptr : params + n.bias_off, out, st_ptr, B, c.cin, c.cout, hs[l],
                                            ws[l], stream));
        }
        if (bn && train && !fused_tail) {
            const int64_t o = n.bn_off[i];
            SIFNN_TRY(sifnn_bn_train_finalize(w.stats + 2 * o, params + n.gamma_off[i], params + n.beta_off[i], running_mean + o, running_var + o,
                                              w.scale + o, w.shift + o, w.mean + o, w.invstd + o, c.cout, (double)B * hs[l] * ws[l], stream));
        }
        return 0;
    };
    auto sc = [&](int i) { return w.scale + n.bn_off[i]; };
    auto sh = [&](int i) { return w.shift + n.bn_off[i]; };

    SIFNN_TRY(conv(0, x, -1, w.raw[0]));
    SIFNN_TRY(conv(1, w.raw[0], 0, w.raw[1]));
    for (int k = 0; k < 3; ++k) {  // db1..db3
        const int src = 3 * k + 1, c0 = 3 * k + 2, c1 = 3 * k + 3, last = 3 * k + 4;
        const int C = n.d[k];
        SIFNN_TRY(sifnn_act_avgpool2_fwd(w.raw[src], sc(src), sh(src), w.P[k], B, C, hs[k], ws[k], stream));
        SIFNN_TRY(conv(c0, w.P[k], -1, w.raw[c0]));
        SIFNN_TRY(conv(c1, w.raw[c0], c0, w.raw[c1]));
        SIFNN_TRY(sifnn_act_residual_fwd(w.P[k], w.raw[c1], sc(c1), sh(c1), w.R[k], B, C, hs[k + 1] * ws[k + 1], stream));
        SIFNN_TRY(conv(last, w.R[k], -1, w.raw[last]));
    }
    for (int k = 0; k < 3; ++k) {  // ub1..ub3
        const int low = (k == 0) ? 10 : 10 + 2 * k, skip = 7 - 3 * k, c0 = 11 + 2 * k, c1 = 12 + 2 * k;
        const int ll = 3 - k;  // level of the low-resolution input
        SIFNN_TRY(sifnn_act_upcat_fwd(w.raw[low], sc(low), sh(low), w.raw[skip], sc(skip), sh(skip), w.U[k], B, n.conv[low].cout,
                                      n.conv[skip].cout, hs[ll], ws[ll], stream));
        SIFNN_TRY(conv(c0, w.U[k], -1, w.raw[c0]));
        SIFNN_TRY(conv(c1, w.raw[c0], c0, w.raw[c1]));
    }
    SIFNN_TRY(conv(17, w.raw[16], 16, y));
    return 0;
}

extern "C" int sifnn_modelb_backward(const sifnn_modelb_cfg* cfg, const float* params, const float* x, const float* dy, float* grads,
                                     void* workspace, int B, int H, int W, int phase, sifnn_stream_t stream) {
    const Net n = build_net(cfg);
    SIFNN_REQUIRE(n.ok, "modelb_backward: unsupported configuration");
    SIFNN_REQUIRE(params && x && dy && grads && workspace, "modelb_backward: null pointer");
    SIFNN_REQUIRE(shape_ok(B, H, W), "modelb_backward: bad shape B=%d H=%d W=%d", B, H, W);
    SIFNN_REQUIRE(phase >= 0 && phase <= 2, "modelb_backward: phase must be 0, 1 or 2");
    const Workspace w = carve(n, workspace, B, H, W, 1);
    cudaStream_t st = sifnn::as_stream(stream);
    const int hs[4] = {H, H / 2, H / 4, H / 8}, ws[4] = {W, W / 2, W / 4, W / 8};

    auto sc = [&](int i) { return w.scale + n.bn_off[i]; };
    auto sh = [&](int i) { return w.shift + n.bn_off[i]; };
    // gradient of conv i's weights; input = raw[aff] through BN+ReLU, or a plain tensor
    // Tensor-core weight gradients leave per-CTA partials in the layer's own buffer; one launch per backward phase sums them all (the 16 small
    // reduce launches were ~5 us each on the critical path of the step).
    sifnn::ReduceJob rjobs[SIFNN_MODELB_NCONV];
    int nrjobs = 0;
    auto flush_reduces = [&]() -> int {
        const int rc = sifnn::wgrad_reduce_many(rjobs, nrjobs, st);
        nrjobs = 0;
        return rc;
    };
    auto wgrad = [&](int i, const float* in, int aff, const float* g) -> int {
        const ConvDesc& c = n.conv[i];
        const float* isc = aff >= 0 ? sc(aff) : nullptr;
        const float* ish = aff >= 0 ? sh(aff) : nullptr;
        // 16-bit K-major kernel (wgrad_km.cu) wherever it takes the shape; the round-1 TF32 kernel otherwise.  (With the deeper raw ring km ties or
        // beats wgrad_tc on the wide layers too: 64->32 @128 118 vs 125 us, 128->64 @64 127 vs 126, 64->32 @64 42 vs 42; profiles/r2za_profile_wgrad.log.)
        const bool km_wins = true;
        int slots = 0;
        if (tc_enabled() && i != 17 && km_wins && sifnn_conv3x3_wgrad_km_supported(c.cin, c.cout, hs[c.level], ws[c.level])) {
            SIFNN_TRY(sifnn::wgrad_km_partials(in, isc, ish, g, w.wgrad_part[i], B, c.cin, c.cout, hs[c.level], ws[c.level], st, &slots));
        } else if (tc_enabled() && i != 17 && sifnn_conv3x3_wgrad_tc_supported(c.cin, c.cout, hs[c.level], ws[c.level])) {
            SIFNN_TRY(sifnn::wgrad_tc_partials(in, isc, ish, g, w.wgrad_part[i], B, c.cin, c.cout, hs[c.level], ws[c.level], st, &slots));
        } else {
            return sifnn_conv3x3_wgrad(in, isc, ish, g, grads + n.w_off[i], i == 17 ? grads + n.bias_off : nullptr, w.wgrad_ws, B, c.cin, c.cout, hs[c.level],
                                       ws[c.level], stream);
        }
        rjobs[nrjobs++] = sifnn::ReduceJob{static_cast<const float*>(w.wgrad_part[i]), grads + n.w_off[i], c.cout * c.cin * 9, slots};
        return 0;
    };
    auto dgrad_kind = [&](int i) {
        const ConvDesc& c = n.conv[i];
        if (i == 0) return KIND_SIMT;
        const bool tc_ok = sifnn_conv3x3_tc_supported(c.cout, c.cin, hs[c.level], ws[c.level]) != 0;
        return conv_kind(c.cout, c.cin, hs[c.level], ws[c.level], tc_ok);
    };
    {   // data-gradient weight splits of the layers this call touches, one launch per kernel family (decoder = conv 11..17, encoder = conv 1..10)
        sifnn::TcPrepJob jobs[SIFNN_MODELB_NCONV];
        const float* pw[2][SIFNN_MODELB_NCONV];
        void* pp[2][SIFNN_MODELB_NCONV];
        float* pe[SIFNN_MODELB_NCONV];
        int pK[2][SIFNN_MODELB_NCONV], pO[2][SIFNN_MODELB_NCONV], pso[2][SIFNN_MODELB_NCONV], psk[2][SIFNN_MODELB_NCONV], pfl[2][SIFNN_MODELB_NCONV];
        int nj = 0, np[2] = {0, 0};
        const int lo = (phase == 2) ? 1 : ((phase == 1) ? 11 : 1), hi = (phase == 2) ? 10 : 17;
        for (int i = lo; i <= hi; ++i) {
            const ConvKind k = dgrad_kind(i);
            if (k == KIND_TC) jobs[nj++] = sifnn::tc_prep_job_dgrad(params + n.w_off[i], w.wprep_d[i], n.conv[i].cin, n.conv[i].cout, ws[n.conv[i].level]);
            if (k == KIND_FF || k == KIND_FS) {
                const int f = (k == KIND_FS) ? 1 : 0, j = np[f]++;
                pw[f][j] = params + n.w_off[i]; pp[f][j] = w.wprep_d[i]; pK[f][j] = n.conv[i].cout; pO[f][j] = n.conv[i].cin;
                pso[f][j] = 9; psk[f][j] = n.conv[i].cin * 9; pfl[f][j] = 1;
                if (f == 1) pe[j] = w.wedge_d[i];
            }
        }
        SIFNN_TRY(sifnn::tc_prep_many(jobs, nj, st));
        if (np[0]) SIFNN_TRY(sifnn::ff_prep(pw[0], pp[0], pK[0], pO[0], pso[0], psk[0], pfl[0], np[0], st));
        if (np[1]) SIFNN_TRY(sifnn::fs_prep(pw[1], pp[1], pe, pK[1], pO[1], pso[1], psk[1], pfl[1], np[1], st));
    }
    auto dgrad = [&](int i, const float* g, float* dx, int accumulate) -> int {
        const ConvDesc& c = n.conv[i];
        switch (dgrad_kind(i)) {
            case KIND_FS:   // complete data gradient, padding adjoint included
                return sifnn::conv3x3_dgrad_fs_prepped(g, w.wprep_d[i], w.wedge_d[i], dx, accumulate, B, c.cin, c.cout, hs[c.level], ws[c.level], st);
            case KIND_FF:
                return sifnn::conv3x3_dgrad_ff_prepped(g, w.wprep_d[i], dx, accumulate, B, c.cin, c.cout, hs[c.level], ws[c.level], st);
            case KIND_TC:
                // the round-1 kernel folds the top/bottom row terms of the padding adjoint in as extra tap MMAs; the column terms (+ corners) follow
                SIFNN_TRY(sifnn::conv3x3_dgrad_tc_main_prepped(g, w.wprep_d[i], dx, accumulate, B, c.cin, c.cout, hs[c.level], ws[c.level], st));
                return sifnn::conv3x3_dgrad_border_cols(g, params + n.w_off[i], dx, B, c.cin, c.cout, hs[c.level], ws[c.level], st);
            default:
                return sifnn_conv3x3_dgrad(g, params + n.w_off[i], dx, accumulate, B, c.cin, c.cout, hs[c.level], ws[c.level], stream);
        }
    };
    // BatchNorm+ReLU backward of layer i: dY (gradient w.r.t. the activated output) -> dx (w.r.t. raw[i])
    auto bnbwd = [&](int i, const float* dY, float* dx) -> int {
        const ConvDesc& c = n.conv[i];
        const int64_t o = n.bn_off[i];
        const int HW = hs[c.level] * ws[c.level];
        SIFNN_TRY(sifnn_bn_relu_bwd_reduce(dY, w.raw[i], w.scale + o, w.shift + o, w.mean + o, w.invstd + o, w.bsums + 2 * o, B, c.cout, HW, stream));
        return sifnn_bn_relu_bwd_apply(dY, w.raw[i], w.scale + o, w.shift + o, w.mean + o, w.invstd + o, params + n.gamma_off[i], w.bsums + 2 * o, dx,
                                       grads + n.gamma_off[i], grads + n.beta_off[i], B, c.cout, HW, stream);
    };

    if (phase == 0 || phase == 1) {
        SIFNN_CUDA(cudaMemsetAsync(w.bsums, 0, sizeof(double) * 2 * n.bn_total, st));
        SIFNN_TRY(wgrad(17, w.raw[16], 16, dy));
        SIFNN_TRY(dgrad(17, dy, w.g[16], 0));
        for (int k = 2; k >= 0; --k) {  // ub3, ub2, ub1
            const int low = (k == 0) ? 10 : 10 + 2 * k, skip = 7 - 3 * k, c0 = 11 + 2 * k, c1 = 12 + 2 * k;
            const int ll = 3 - k;
            SIFNN_TRY(bnbwd(c1, w.g[c1], w.g[c1]));
            SIFNN_TRY(wgrad(c1, w.raw[c0], c0, w.g[c1]));
            SIFNN_TRY(dgrad(c1, w.g[c1], w.g[c0], 0));
            SIFNN_TRY(bnbwd(c0, w.g[c0], w.g[c0]));
            SIFNN_TRY(wgrad(c0, w.U[k], -1, w.g[c0]));
            SIFNN_TRY(dgrad(c0, w.g[c0], w.gU[k], 0));
            SIFNN_TRY(sifnn_upcat_bwd(w.gU[k], w.g[low], w.g[skip], B, n.conv[low].cout, n.conv[skip].cout, hs[ll], ws[ll], stream));
        }
        SIFNN_TRY(flush_reduces());   // decoder gradients complete (the first all-reduce bucket of a data-parallel step)
    }
    if (phase == 0 || phase == 2) {
        for (int k = 2; k >= 0; --k) {  // db3, db2, db1
            const int src = 3 * k + 1, c0 = 3 * k + 2, c1 = 3 * k + 3, last = 3 * k + 4;
            const int C = n.d[k];
            SIFNN_TRY(bnbwd(last, w.g[last], w.g[last]));
            SIFNN_TRY(wgrad(last, w.R[k], -1, w.g[last]));
            SIFNN_TRY(dgrad(last, w.g[last], w.gR[k], 0));
            SIFNN_TRY(bnbwd(c1, w.gR[k], w.g[c1]));           // residual branch; gR[k] stays = dL/dP so far
            SIFNN_TRY(wgrad(c1, w.raw[c0], c0, w.g[c1]));
            SIFNN_TRY(dgrad(c1, w.g[c1], w.g[c0], 0));
            SIFNN_TRY(bnbwd(c0, w.g[c0], w.g[c0]));
            SIFNN_TRY(wgrad(c0, w.P[k], -1, w.g[c0]));
            SIFNN_TRY(dgrad(c0, w.g[c0], w.gR[k], 1));          // += : gR[k] is now dL/dP
            SIFNN_TRY(sifnn_avgpool2_bwd(w.gR[k], w.g[src], 1, B, C, hs[k], ws[k], stream));  // += onto the skip gradient
        }
        SIFNN_TRY(bnbwd(1, w.g[1], w.g[1]));
        SIFNN_TRY(wgrad(1, w.raw[0], 0, w.g[1]));
        SIFNN_TRY(dgrad(1, w.g[1], w.g[0], 0));
        SIFNN_TRY(bnbwd(0, w.g[0], w.g[0]));
        SIFNN_TRY(wgrad(0, x, -1, w.g[0]));
        SIFNN_TRY(flush_reduces());
    }
    return 0;
}
