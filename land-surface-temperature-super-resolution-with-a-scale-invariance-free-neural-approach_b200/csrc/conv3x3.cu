// 3x3 convolution, fp32 SIMT direct form: forward (replicate padding, optional fused
// BatchNorm+ReLU prologue, BatchNorm statistics epilogue) and data gradient (the same
// kernel with zero padding and transposed/flipped weights, plus a border pass that adds
// the adjoint of the replicate padding).
//
// Replaces nn.Conv2d(k=3,padding=1,padding_mode='replicate') at model.py:135,138,507,605
// (cuDNN/mkldnn implicit GEMM + a materialised replication_pad2d in the reference) and
// its autograd.
//
// Tiling: one CTA = 256 threads = 8 warps computes a (ROWS x 32) pixel tile for CO_T
// output channels of one image.  Lanes run along W (coalesced, bank-conflict free);
// each thread owns 8 rows x CPT output channels in registers (64 accumulators for
// CPT = 8) and, per input channel, reads a 10x3 input window plus 9*CPT broadcast
// weights from shared memory: 576 FFMA per 48 LDS.
#include "common.cuh"

namespace {

constexpr int TW = 32;
constexpr int CI_CHUNK = 8;
constexpr int PY = 8;
// shared-memory input rows: [3] = left halo, [4..35] = interior (16B aligned), [36] = right halo
constexpr int IN_STRIDE = 40;
constexpr int IN_X0 = 3;
constexpr int MAX_K = 512;  // largest channel count the fused BatchNorm prologue supports

enum { PAD_REPLICATE = 0, PAD_ZERO = 1 };

struct ConvArgs {
    const float* in;
    const float* in_scale;
    const float* in_shift;
    const float* w;
    const float* bias;
    float* out;
    double* stats;
    int B, K, O, H, W;  // K input channels, O output channels of THIS op
    int w_so, w_sk, w_flip;
    int accumulate;
    int tiles_x;
    int vec_ok;  // W % 4 == 0 and 16-byte aligned base: interior columns may be copied 16 B at a time
};

// Stage one chunk of input channels (with halo) and its weights into shared memory with
// cp.async.  Vector path (16 B) for the 32 interior columns when the tile lies fully inside
// the image in x; the two halo columns and the fallback path are 4 B copies.  Zero padding
// (PAD_ZERO) is produced by the zero-fill form of cp.async (src-size 0).
template <int IN_ROWS, int CO_T, int PAD>
__device__ __forceinline__ void conv_issue_fill(float* in_st, float* w_st, const ConvArgs& a, const float* in_b, int c0, int nci,
                                                int x0, int y0, int o0, bool vec_ok, int tid) {
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    if (vec_ok) {
        for (int idx = tid; idx < nci * IN_ROWS * 10; idx += 256) {
            const int ci = idx / (IN_ROWS * 10);
            const int rem = idx - ci * (IN_ROWS * 10);
            const int r = rem / 10, j = rem - r * 10;
            int gy = y0 + r - 1;
            bool ok = true;
            if (PAD == PAD_REPLICATE) gy = min(max(gy, 0), H - 1);
            else ok = gy >= 0 && gy < H;
            const float* src = in_b + (size_t)(c0 + ci) * plane + (size_t)(ok ? gy : 0) * W;
            float* dst = in_st + ci * IN_PLANE + r * IN_STRIDE;
            if (j < 8) {
                sifnn::cp_async16(dst + IN_X0 + 1 + 4 * j, src + x0 + 4 * j, ok ? 16 : 0);
            } else if (j == 8) {
                int gx = x0 - 1;
                if (PAD == PAD_REPLICATE) gx = max(gx, 0);
                else ok = ok && gx >= 0;
                sifnn::cp_async4(dst + IN_X0, src + (ok ? gx : 0), ok ? 4 : 0);
            } else {
                int gx = x0 + TW;
                if (PAD == PAD_REPLICATE) gx = min(gx, W - 1);
                else ok = ok && gx < W;
                sifnn::cp_async4(dst + IN_X0 + 1 + TW, src + (ok ? gx : 0), ok ? 4 : 0);
            }
        }
    } else {
        constexpr int IN_COLS = TW + 2;
        for (int idx = tid; idx < nci * IN_ROWS * IN_COLS; idx += 256) {
            const int ci = idx / (IN_ROWS * IN_COLS);
            const int rem = idx - ci * (IN_ROWS * IN_COLS);
            const int r = rem / IN_COLS, c = rem - r * IN_COLS;
            int gy = y0 + r - 1, gx = x0 + c - 1;
            bool ok = true;
            if (PAD == PAD_REPLICATE) {
                gy = min(max(gy, 0), H - 1);
                gx = min(max(gx, 0), W - 1);
            } else {
                ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
            }
            sifnn::cp_async4(in_st + ci * IN_PLANE + r * IN_STRIDE + IN_X0 + c,
                             in_b + (size_t)(c0 + ci) * plane + (ok ? (size_t)gy * W + gx : 0), ok ? 4 : 0);
        }
    }
    // weights: w_st[ci][tap][o]
    for (int idx = tid; idx < nci * 9 * CO_T; idx += 256) {
        const int o = idx % CO_T;
        const int t = (idx / CO_T) % 9;
        const int ci = idx / (9 * CO_T);
        const bool ok = o0 + o < a.O;
        sifnn::cp_async4(w_st + idx, a.w + (ok ? (size_t)(o0 + o) * a.w_so + (size_t)(c0 + ci) * a.w_sk + (a.w_flip ? 8 - t : t) : 0), ok ? 4 : 0);
    }
}

// BatchNorm + ReLU of the producing layer applied in place to the elements THIS thread copied
// (visible to it after cp.async.wait_group, before the CTA barrier).
template <int IN_ROWS>
__device__ __forceinline__ void conv_affine_pass(float* in_st, const float* sc_s, const float* sh_s, int c0, int nci, bool vec_ok, int tid) {
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    if (vec_ok) {
        for (int idx = tid; idx < nci * IN_ROWS * 10; idx += 256) {
            const int ci = idx / (IN_ROWS * 10);
            const int rem = idx - ci * (IN_ROWS * 10);
            const int r = rem / 10, j = rem - r * 10;
            const float sc = sc_s[c0 + ci], sh = sh_s[c0 + ci];
            float* dst = in_st + ci * IN_PLANE + r * IN_STRIDE;
            if (j < 8) {
                float4* q = reinterpret_cast<float4*>(dst + IN_X0 + 1 + 4 * j);
                float4 v = *q;
                v.x = sifnn::act_affine_relu(v.x, sc, sh); v.y = sifnn::act_affine_relu(v.y, sc, sh);
                v.z = sifnn::act_affine_relu(v.z, sc, sh); v.w = sifnn::act_affine_relu(v.w, sc, sh);
                *q = v;
            } else {
                float* q = dst + (j == 8 ? IN_X0 : IN_X0 + 1 + TW);
                *q = sifnn::act_affine_relu(*q, sc, sh);
            }
        }
    } else {
        constexpr int IN_COLS = TW + 2;
        for (int idx = tid; idx < nci * IN_ROWS * IN_COLS; idx += 256) {
            const int ci = idx / (IN_ROWS * IN_COLS);
            const int rem = idx - ci * (IN_ROWS * IN_COLS);
            const int r = rem / IN_COLS, c = rem - r * IN_COLS;
            float* q = in_st + ci * IN_PLANE + r * IN_STRIDE + IN_X0 + c;
            *q = sifnn::act_affine_relu(*q, sc_s[c0 + ci], sh_s[c0 + ci]);
        }
    }
}

template <int CPT, int WARPS_CO, int PAD, bool AFFINE>
__global__ void __launch_bounds__(256, 2) conv3x3_kernel(const ConvArgs a) {
    constexpr int WARPS_ROW = 8 / WARPS_CO;
    constexpr int ROWS = PY * WARPS_ROW;
    constexpr int CO_T = CPT * WARPS_CO;
    constexpr int IN_ROWS = ROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int STAGE = CI_CHUNK * IN_PLANE + CI_CHUNK * 9 * CO_T;  // floats per pipeline stage

    extern __shared__ __align__(16) float smem[];
    __shared__ float sc_s[AFFINE ? MAX_K : 1], sh_s[AFFINE ? MAX_K : 1];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wc = warp % WARPS_CO;
    const int wr = warp / WARPS_CO;
    const int tx = blockIdx.x % a.tiles_x;
    const int ty = blockIdx.x / a.tiles_x;
    const int x0 = tx * TW;
    const int y0 = ty * ROWS;
    const int o0 = blockIdx.y * CO_T;
    const int b = blockIdx.z;
    const int H = a.H, W = a.W, K = a.K;
    const size_t plane = (size_t)H * W;
    const float* in_b = a.in + (size_t)b * K * plane;
    const bool vec_ok = a.vec_ok && (x0 + TW <= W);

    if (AFFINE) {
        for (int i = tid; i < K; i += 256) { sc_s[i] = __ldg(a.in_scale + i); sh_s[i] = __ldg(a.in_shift + i); }
    }

    float acc[PY][CPT];
#pragma unroll
    for (int i = 0; i < PY; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

    const int nchunks = (K + CI_CHUNK - 1) / CI_CHUNK;
    conv_issue_fill<IN_ROWS, CO_T, PAD>(smem, smem + CI_CHUNK * IN_PLANE, a, in_b, 0, min(CI_CHUNK, K), x0, y0, o0, vec_ok, tid);
    sifnn::cp_async_commit();

    for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * CI_CHUNK;
        const int nci = min(CI_CHUNK, K - c0);
        float* in_s = smem + (ch & 1) * STAGE;
        float* w_s = in_s + CI_CHUNK * IN_PLANE;
        if (ch + 1 < nchunks) {  // prefetch the next chunk into the other stage while this one is consumed
            float* nin = smem + ((ch + 1) & 1) * STAGE;
            conv_issue_fill<IN_ROWS, CO_T, PAD>(nin, nin + CI_CHUNK * IN_PLANE, a, in_b, c0 + CI_CHUNK, min(CI_CHUNK, K - c0 - CI_CHUNK), x0, y0, o0, vec_ok, tid);
            sifnn::cp_async_commit();
            sifnn::cp_async_wait<1>();
        } else {
            sifnn::cp_async_wait<0>();
        }
        if (AFFINE) {
            if (ch == 0) __syncthreads();  // sc_s / sh_s
            conv_affine_pass<IN_ROWS>(in_s, sc_s, sh_s, c0, nci, vec_ok, tid);
        }
        __syncthreads();

        const float* ip = in_s + (wr * PY) * IN_STRIDE + IN_X0 + lane;
        const float* wp = w_s + wc * CPT;
#pragma unroll 1
        for (int ci = 0; ci < nci; ++ci) {
            float v[PY + 2][3];
#pragma unroll
            for (int r = 0; r < PY + 2; ++r)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) v[r][kx] = ip[ci * IN_PLANE + r * IN_STRIDE + kx];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float wv[CPT];
                    const float* wq = wp + (ci * 9 + ky * 3 + kx) * CO_T;
                    if (CPT % 4 == 0) {
#pragma unroll
                        for (int j = 0; j < CPT; j += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(wq + j);
                            wv[j] = t4.x; wv[j + 1] = t4.y; wv[j + 2] = t4.z; wv[j + 3] = t4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CPT; ++j) wv[j] = wq[j];
                    }
#pragma unroll
                    for (int py = 0; py < PY; ++py) {
                        const float x = v[py + ky][kx];
#pragma unroll
                        for (int j = 0; j < CPT; ++j) acc[py][j] = fmaf(x, wv[j], acc[py][j]);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this stage before it is refilled two iterations later
    }

    // ---- epilogue ---------------------------------------------------------------------------
    const int x = x0 + lane;
    const bool xok = x < W;
    float s1[CPT], s2[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int o = o0 + wc * CPT + j;
        if (o < a.O) {
            const float bv = a.bias ? __ldg(a.bias + o) : 0.f;
            float* op = a.out + ((size_t)b * a.O + o) * plane;
#pragma unroll
            for (int py = 0; py < PY; ++py) {
                const int y = y0 + wr * PY + py;
                if (xok && y < H) {
                    float r = acc[py][j] + bv;
                    const size_t off = (size_t)y * W + x;
                    if (a.accumulate) r += op[off];
                    op[off] = r;
                    s1[j] += r;
                    s2[j] = fmaf(r, r, s2[j]);
                }
            }
        }
    }
    if (a.stats) {
        __syncthreads();  // in_s is dead now; reuse it
        float* red = smem; // [WARPS_ROW][CO_T][2]
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const float t1 = sifnn::warp_sum(s1[j]);
            const float t2 = sifnn::warp_sum(s2[j]);
            if (lane == 0) {
                red[(wr * CO_T + wc * CPT + j) * 2 + 0] = t1;
                red[(wr * CO_T + wc * CPT + j) * 2 + 1] = t2;
            }
        }
        __syncthreads();
        if (tid < CO_T && o0 + tid < a.O) {
            double d1 = 0.0, d2 = 0.0;
#pragma unroll
            for (int r = 0; r < WARPS_ROW; ++r) {
                d1 += (double)red[(r * CO_T + tid) * 2 + 0];
                d2 += (double)red[(r * CO_T + tid) * 2 + 1];
            }
            atomicAdd(a.stats + o0 + tid, d1);
            atomicAdd(a.stats + a.O + o0 + tid, d2);
        }
    }
}

// Adjoint of the replicate padding for the data gradient: the zero-padded transposed
// convolution misses the taps that the forward pass read through a clamped index.  Those
// are four 1-D problems (top / bottom row, left / right column):
//   top    : dx[k][0][q]   += sum_o sum_kx w[o][k][0][kx] * dy[o][0][q-kx+1]   (+ corner cross terms)
//   bottom : dx[k][H-1][q] += sum_o sum_kx w[o][k][2][kx] * dy[o][H-1][q-kx+1]
//   left   : dx[k][p][0]   += sum_o sum_ky w[o][k][ky][0] * dy[o][p-ky+1][0]
//   right  : dx[k][p][W-1] += sum_o sum_ky w[o][k][ky][2] * dy[o][p-ky+1][W-1]
// One CTA = 32 consecutive border pixels of one side of one image (lanes), 8 warps split
// the input channels; the three dy values a lane needs per output channel are loaded once
// and reused for every k, the weights come from shared memory as broadcast float4.
// Launched twice (rows, then columns) so the corner pixels are updated without a race.
constexpr int BORDER_OC = 16;  // output-channel chunk staged in shared memory

__global__ void __launch_bounds__(256) dgrad_border_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                           int Cin, int Cout, int H, int W, int cols_pass) {
    extern __shared__ __align__(16) float ws[];  // [BORDER_OC][4 taps (3 main + corner)][Cin_pad]
    const int Cp = (Cin + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int L = cols_pass ? H : W;              // side length
    const int segs = (L + 31) / 32;
    const int side = blockIdx.x / segs;           // 0: top/left, 1: bottom/right
    const int i = (blockIdx.x % segs) * 32 + lane;  // position along the side
    const bool active = i < L;
    // main taps t0 + j*tstep (j = 0..2) read dy at (r0 + j*dr, c0 + j*dc)
    int t0, tstep, r0, c0, dr, dc, p, q;
    if (!cols_pass) {
        p = side ? H - 1 : 0; q = i;
        t0 = side ? 6 : 0; tstep = 1; r0 = p; dr = 0; c0 = q + 1; dc = -1;
    } else {
        p = i; q = side ? W - 1 : 0;
        t0 = side ? 2 : 0; tstep = 3; r0 = p + 1; dr = -1; c0 = q; dc = 0;
    }
    // corner cross term (rows pass only): tap (ky_e, kx_e) on dy[r_e][c_e]
    int tc = -1;
    if (!cols_pass && active) {
        if (q == 0) tc = (side ? 6 : 0);
        else if (q == W - 1) tc = (side ? 8 : 2);
    }
    const size_t plane = (size_t)H * W;
    const float* dyb = dy + (size_t)b * Cout * plane;
    bool ok[3];
    size_t off[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int r = r0 + j * dr, c = c0 + j * dc;
        ok[j] = active && r >= 0 && r < H && c >= 0 && c < W;
        off[j] = ok[j] ? (size_t)r * W + c : 0;
    }
    const size_t offc = (size_t)p * W + q;  // corner term reads dy at the pixel itself

    for (int kbase = 0; kbase < Cin; kbase += 64) {   // each warp: 8 input channels at a time (loop trip count is CTA-uniform)
        const int kb = kbase + warp * 8;
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
        for (int ob = 0; ob < Cout; ob += BORDER_OC) {
            const int no = min(BORDER_OC, Cout - ob);
            __syncthreads();
            // stage w[ob..ob+no)[all k][the 3 main taps + both possible corner taps]; layout [o][slot][k]
            for (int idx = threadIdx.x; idx < no * 5 * Cp; idx += 256) {
                const int k = idx % Cp;
                const int slot = (idx / Cp) % 5;
                const int o = idx / (5 * Cp);
                int t;
                if (slot < 3) t = t0 + slot * tstep;
                else t = (slot == 3) ? (side ? 6 : 0) : (side ? 8 : 2);
                ws[idx] = (k < Cin) ? __ldg(w + ((size_t)(ob + o) * Cin + k) * 9 + t) : 0.f;
            }
            __syncthreads();
            if (kb < Cin) {
                for (int o = 0; o < no; ++o) {
                    const float* dyo = dyb + (size_t)(ob + o) * plane;
                    float d[4];
#pragma unroll
                    for (int j = 0; j < 3; ++j) d[j] = ok[j] ? __ldg(dyo + off[j]) : 0.f;
                    d[3] = (tc >= 0) ? __ldg(dyo + offc) : 0.f;
                    const float* wo = ws + (size_t)o * 5 * Cp + kb;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float4 wa = *reinterpret_cast<const float4*>(wo + j * Cp);
                        const float4 wb = *reinterpret_cast<const float4*>(wo + j * Cp + 4);
                        acc[0] = fmaf(wa.x, d[j], acc[0]); acc[1] = fmaf(wa.y, d[j], acc[1]);
                        acc[2] = fmaf(wa.z, d[j], acc[2]); acc[3] = fmaf(wa.w, d[j], acc[3]);
                        acc[4] = fmaf(wb.x, d[j], acc[4]); acc[5] = fmaf(wb.y, d[j], acc[5]);
                        acc[6] = fmaf(wb.z, d[j], acc[6]); acc[7] = fmaf(wb.w, d[j], acc[7]);
                    }
                    if (tc >= 0) {  // lane-divergent but only the two corner lanes of a rows pass take it
                        const float* wc = wo + ((tc == 0 || tc == 6) ? 3 : 4) * Cp;
#pragma unroll
                        for (int u = 0; u < 8; ++u) acc[u] = fmaf(wc[u], d[3], acc[u]);
                    }
                }
            }
        }
        if (active && kb < Cin) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kb + u < Cin) dx[((size_t)b * Cin + kb + u) * plane + (size_t)p * W + q] += acc[u];
        }
    }
}

template <int CPT, int WARPS_CO, int PAD, bool AFFINE>
int launch_conv(const ConvArgs& a0, cudaStream_t st) {
    constexpr int WARPS_ROW = 8 / WARPS_CO;
    constexpr int ROWS = PY * WARPS_ROW;
    constexpr int CO_T = CPT * WARPS_CO;
    constexpr int IN_PLANE = (ROWS + 2) * IN_STRIDE;
    constexpr size_t smem = 2 * (size_t)(CI_CHUNK * IN_PLANE + CI_CHUNK * 9 * CO_T) * sizeof(float);  // two pipeline stages
    static bool attr_done = false;
    auto kern = conv3x3_kernel<CPT, WARPS_CO, PAD, AFFINE>;
    if (!attr_done) {
        SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    ConvArgs a = a0;
    a.tiles_x = (a.W + TW - 1) / TW;
    a.vec_ok = (a.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
    const int tiles_y = (a.H + ROWS - 1) / ROWS;
    dim3 grid(a.tiles_x * tiles_y, (a.O + CO_T - 1) / CO_T, a.B);
    kern<<<grid, 256, smem, st>>>(a);
    return sifnn::check_launch("conv3x3_kernel");
}

template <int PAD, bool AFFINE>
int dispatch_conv(const ConvArgs& a, cudaStream_t st) {
    if (a.O <= 4) return launch_conv<1, 1, PAD, AFFINE>(a, st);
    if (a.O % 32 == 0) return launch_conv<8, 4, PAD, AFFINE>(a, st);
    return launch_conv<8, 2, PAD, AFFINE>(a, st);
}

}  // namespace

extern "C" int sifnn_conv3x3_fwd(const float* in, const float* in_scale, const float* in_shift, const float* w,
                                 const float* bias, float* out, double* stats, int B, int Cin, int Cout, int H, int W,
                                 sifnn_stream_t stream) {
    SIFNN_REQUIRE(in && w && out, "conv3x3_fwd: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_fwd: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(in_scale == nullptr || Cin <= MAX_K, "conv3x3_fwd: fused BatchNorm prologue supports at most %d input channels", MAX_K);
    SIFNN_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && B <= 65535, "conv3x3_fwd: bad shape B=%d Cin=%d Cout=%d H=%d W=%d", B, Cin, Cout, H, W);
    ConvArgs a{};
    a.in = in; a.in_scale = in_scale; a.in_shift = in_shift; a.w = w; a.bias = bias; a.out = out; a.stats = stats;
    a.B = B; a.K = Cin; a.O = Cout; a.H = H; a.W = W;
    a.w_so = Cin * 9; a.w_sk = 9; a.w_flip = 0; a.accumulate = 0;
    cudaStream_t st = sifnn::as_stream(stream);
    return in_scale ? dispatch_conv<PAD_REPLICATE, true>(a, st) : dispatch_conv<PAD_REPLICATE, false>(a, st);
}

extern "C" int sifnn_conv3x3_dgrad(const float* dy, const float* w, float* dx, int accumulate, int B, int Cin, int Cout,
                                   int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dy && w && dx, "conv3x3_dgrad: null pointer");
    SIFNN_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H >= 2 && W >= 2 && B <= 65535, "conv3x3_dgrad: bad shape B=%d Cin=%d Cout=%d H=%d W=%d", B, Cin, Cout, H, W);
    ConvArgs a{};
    a.in = dy; a.w = w; a.out = dx;
    a.B = B; a.K = Cout; a.O = Cin; a.H = H; a.W = W;
    a.w_so = 9; a.w_sk = Cin * 9; a.w_flip = 1; a.accumulate = accumulate ? 1 : 0;
    cudaStream_t st = sifnn::as_stream(stream);
    SIFNN_TRY((dispatch_conv<PAD_ZERO, false>(a, st)));
    const int Cp = (Cin + 3) & ~3;
    const size_t bsmem = (size_t)BORDER_OC * 5 * (Cp + 8) * sizeof(float);
    for (int cols_pass = 0; cols_pass < 2; ++cols_pass) {
        const int L = cols_pass ? H : W;
        dim3 grid(2 * ((L + 31) / 32), B);
        dgrad_border_kernel<<<grid, 256, bsmem, st>>>(dy, w, dx, Cin, Cout, H, W, cols_pass);
        if (cols_pass == 0) SIFNN_TRY(sifnn::check_launch("dgrad_border_kernel"));
    }
    return sifnn::check_launch("dgrad_border_kernel");
}
