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
};

template <int CPT, int WARPS_CO, int PAD, bool AFFINE>
__global__ void __launch_bounds__(256, 2) conv3x3_kernel(const ConvArgs a) {
    constexpr int WARPS_ROW = 8 / WARPS_CO;
    constexpr int ROWS = PY * WARPS_ROW;
    constexpr int CO_T = CPT * WARPS_CO;
    constexpr int IN_ROWS = ROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int IN_COLS = TW + 2;

    extern __shared__ __align__(16) float smem[];
    float* in_s = smem;                        // CI_CHUNK * IN_PLANE
    float* w_s = in_s + CI_CHUNK * IN_PLANE;   // CI_CHUNK * 9 * CO_T

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

    float acc[PY][CPT];
#pragma unroll
    for (int i = 0; i < PY; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < K; c0 += CI_CHUNK) {
        const int nci = min(CI_CHUNK, K - c0);
        __syncthreads();
        // ---- stage the input chunk (with halo) ------------------------------------------
        for (int idx = tid; idx < nci * IN_ROWS * IN_COLS; idx += 256) {
            const int ci = idx / (IN_ROWS * IN_COLS);
            const int rem = idx - ci * (IN_ROWS * IN_COLS);
            const int r = rem / IN_COLS;
            const int c = rem - r * IN_COLS;
            int gy = y0 + r - 1, gx = x0 + c - 1;
            float v;
            if (PAD == PAD_REPLICATE) {
                gy = min(max(gy, 0), H - 1);
                gx = min(max(gx, 0), W - 1);
                v = __ldg(in_b + (size_t)(c0 + ci) * plane + (size_t)gy * W + gx);
                if (AFFINE) v = sifnn::act_affine_relu(v, __ldg(a.in_scale + c0 + ci), __ldg(a.in_shift + c0 + ci));
            } else {
                v = 0.f;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    v = __ldg(in_b + (size_t)(c0 + ci) * plane + (size_t)gy * W + gx);
                    if (AFFINE) v = sifnn::act_affine_relu(v, __ldg(a.in_scale + c0 + ci), __ldg(a.in_shift + c0 + ci));
                }
            }
            in_s[ci * IN_PLANE + r * IN_STRIDE + IN_X0 + c] = v;
        }
        // ---- stage the weights: w_s[ci][tap][o] ---------------------------------------------
        for (int idx = tid; idx < nci * 9 * CO_T; idx += 256) {
            const int o = idx % CO_T;
            const int t = (idx / CO_T) % 9;
            const int ci = idx / (9 * CO_T);
            float v = 0.f;
            if (o0 + o < a.O) v = __ldg(a.w + (size_t)(o0 + o) * a.w_so + (size_t)(c0 + ci) * a.w_sk + (a.w_flip ? 8 - t : t));
            w_s[idx] = v;
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
// convolution misses the taps that the forward pass read through a clamped index.
// One thread per (image, input channel, border pixel).
__global__ void dgrad_border_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                    int B, int Cin, int Cout, int H, int W) {
    const int nb = 2 * W + 2 * (H - 2);
    const long long total = (long long)B * Cin * nb;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(idx % nb);
        const int k = (int)((idx / nb) % Cin);
        const int b = (int)(idx / ((long long)nb * Cin));
        int p, q;
        if (e < W) { p = 0; q = e; }
        else if (e < 2 * W) { p = H - 1; q = e - W; }
        else if (e < 2 * W + (H - 2)) { p = e - 2 * W + 1; q = 0; }
        else { p = e - 2 * W - (H - 2) + 1; q = W - 1; }
        // extra (clamped) taps: row side (ky_e reads dy row r_e), column side (kx_e reads dy col c_e)
        const int ky_e = (p == 0) ? 0 : ((p == H - 1) ? 2 : -1);
        const int r_e = (p == 0) ? 0 : H - 1;
        const int kx_e = (q == 0) ? 0 : ((q == W - 1) ? 2 : -1);
        const int c_e = (q == 0) ? 0 : W - 1;
        float sum = 0.f;
        for (int o = 0; o < Cout; ++o) {
            const float* wk = w + ((size_t)o * Cin + k) * 9;
            const float* dyo = dy + ((size_t)b * Cout + o) * H * W;
            if (ky_e >= 0) {
                for (int kx = 0; kx < 3; ++kx) {
                    const int c = q - kx + 1;
                    if (c >= 0 && c < W) sum = fmaf(__ldg(wk + ky_e * 3 + kx), __ldg(dyo + (size_t)r_e * W + c), sum);
                }
                if (kx_e >= 0) sum = fmaf(__ldg(wk + ky_e * 3 + kx_e), __ldg(dyo + (size_t)r_e * W + c_e), sum);
            }
            if (kx_e >= 0) {
                for (int ky = 0; ky < 3; ++ky) {
                    const int r = p - ky + 1;
                    if (r >= 0 && r < H) sum = fmaf(__ldg(wk + ky * 3 + kx_e), __ldg(dyo + (size_t)r * W + c_e), sum);
                }
            }
        }
        dx[((size_t)b * Cin + k) * H * W + (size_t)p * W + q] += sum;
    }
}

template <int CPT, int WARPS_CO, int PAD, bool AFFINE>
int launch_conv(const ConvArgs& a0, cudaStream_t st) {
    constexpr int WARPS_ROW = 8 / WARPS_CO;
    constexpr int ROWS = PY * WARPS_ROW;
    constexpr int CO_T = CPT * WARPS_CO;
    constexpr int IN_PLANE = (ROWS + 2) * IN_STRIDE;
    constexpr size_t smem = (size_t)(CI_CHUNK * IN_PLANE + CI_CHUNK * 9 * CO_T) * sizeof(float);
    static bool attr_done = false;
    auto kern = conv3x3_kernel<CPT, WARPS_CO, PAD, AFFINE>;
    if (!attr_done) {
        SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    ConvArgs a = a0;
    a.tiles_x = (a.W + TW - 1) / TW;
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
    const long long total = (long long)B * Cin * (2 * W + 2 * (H - 2));
    const int threads = 128;
    long long nblk = (total + threads - 1) / threads;
    if (nblk > 148 * 16) nblk = 148 * 16;
    const int blocks = (int)nblk;
    dgrad_border_kernel<<<blocks, threads, 0, st>>>(dy, w, dx, B, Cin, Cout, H, W);
    return sifnn::check_launch("dgrad_border_kernel");
}
