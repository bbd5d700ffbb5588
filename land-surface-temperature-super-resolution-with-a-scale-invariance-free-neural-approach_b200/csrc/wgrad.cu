// Weight gradient of the 3x3 replicate-padded convolution (autograd of model.py:135;
// loss.backward() at train_model_B_gradFTM.py:119), fp32 SIMT.
//
//   dW[o][k][ky][kx] = sum_{b,y,x} dy[b][o][y][x] * act(in)[b][k][clamp(y+ky-1)][clamp(x+kx-1)]
//
// A reduction over B*H*W pixels into O*K*9 outputs.  One CTA owns an (O_CHUNK x K_CHUNK)
// block of the output and strides over pixel tiles (16 rows x 32 columns); each warp owns
// an (OT x KT x 9) register tile, lanes run along W and walk down the 16 rows with a
// sliding 3x3 window: 72 FFMA per 10 conflict-free LDS.  Lanes are reduced by shuffles
// once per CTA; per-CTA partials are summed in a fixed order by a second kernel, so the
// result is deterministic (no atomics).
#include "common.cuh"

namespace {

constexpr int WROWS = 16;
constexpr int IN_STRIDE = 40;
constexpr int IN_X0 = 3;
constexpr int IN_COLS = 34;

struct WgradArgs {
    const float* in;
    const float* in_scale;
    const float* in_shift;
    const float* dy;
    float* partial;       // [S][O][K][9]
    float* bias_partial;  // [S][O]
    int B, K, O, H, W;
    int tiles_x, tiles_per_img, total_tiles;
    int vec_ok;
};

// cp.async staging of one pixel tile: K_CHUNK input planes with halo (replicate-clamped) and
// O_CHUNK dy planes (zero-filled outside the image / past the last channel).
template <int O_CHUNK, int K_CHUNK, int NT>
__device__ __forceinline__ void wgrad_issue_fill(float* in_s, float* dy_s, const WgradArgs& a, int b, int x0, int y0, int k0, int o0,
                                                 bool vec_ok, int tid) {
    constexpr int IN_ROWS = WROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int DY_PLANE = WROWS * 32;
    const int H = a.H, W = a.W, K = a.K, O = a.O;
    const size_t plane = (size_t)H * W;
    if (vec_ok) {
        for (int idx = tid; idx < K_CHUNK * IN_ROWS * 10; idx += NT) {
            const int kk = idx / (IN_ROWS * 10);
            const int rem = idx - kk * (IN_ROWS * 10);
            const int r = rem / 10, j = rem - r * 10;
            const bool ok = k0 + kk < K;
            const int gy = min(max(y0 + r - 1, 0), H - 1);
            const float* src = a.in + ((size_t)b * K + (ok ? k0 + kk : 0)) * plane + (size_t)gy * W;
            float* dst = in_s + kk * IN_PLANE + r * IN_STRIDE;
            if (j < 8) sifnn::cp_async16(dst + IN_X0 + 1 + 4 * j, src + x0 + 4 * j, ok ? 16 : 0);
            else if (j == 8) sifnn::cp_async4(dst + IN_X0, src + max(x0 - 1, 0), ok ? 4 : 0);
            else sifnn::cp_async4(dst + IN_X0 + 33, src + min(x0 + 32, W - 1), ok ? 4 : 0);
        }
        for (int idx = tid; idx < O_CHUNK * WROWS * 8; idx += NT) {
            const int oo = idx / (WROWS * 8);
            const int rem = idx - oo * (WROWS * 8);
            const int r = rem >> 3, j = rem & 7;
            const bool ok = (o0 + oo < O) && (y0 + r < H);
            const float* src = a.dy + ((size_t)b * O + (ok ? o0 + oo : 0)) * plane + (size_t)(ok ? y0 + r : 0) * W + x0 + 4 * j;
            sifnn::cp_async16(dy_s + oo * DY_PLANE + r * 32 + 4 * j, src, ok ? 16 : 0);
        }
    } else {
        for (int idx = tid; idx < K_CHUNK * IN_ROWS * IN_COLS; idx += NT) {
            const int kk = idx / (IN_ROWS * IN_COLS);
            const int rem = idx - kk * (IN_ROWS * IN_COLS);
            const int r = rem / IN_COLS, c = rem - r * IN_COLS;
            const bool ok = k0 + kk < K;
            const int gy = min(max(y0 + r - 1, 0), H - 1);
            const int gx = min(max(x0 + c - 1, 0), W - 1);
            sifnn::cp_async4(in_s + kk * IN_PLANE + r * IN_STRIDE + IN_X0 + c,
                             a.in + ((size_t)b * K + (ok ? k0 + kk : 0)) * plane + (size_t)gy * W + gx, ok ? 4 : 0);
        }
        for (int idx = tid; idx < O_CHUNK * DY_PLANE; idx += NT) {
            const int oo = idx / DY_PLANE;
            const int rem = idx - oo * DY_PLANE;
            const int r = rem >> 5, c = rem & 31;
            const bool ok = (o0 + oo < O) && (y0 + r < H) && (x0 + c < W);
            sifnn::cp_async4(dy_s + idx, a.dy + (ok ? ((size_t)b * O + o0 + oo) * plane + (size_t)(y0 + r) * W + x0 + c : 0), ok ? 4 : 0);
        }
    }
}

template <int K_CHUNK, int NT>
__device__ __forceinline__ void wgrad_affine_pass(float* in_s, const float* sc_s, const float* sh_s, int K, int k0, bool vec_ok, int tid) {
    constexpr int IN_ROWS = WROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    if (vec_ok) {
        for (int idx = tid; idx < K_CHUNK * IN_ROWS * 10; idx += NT) {
            const int kk = idx / (IN_ROWS * 10);
            const int rem = idx - kk * (IN_ROWS * 10);
            const int r = rem / 10, j = rem - r * 10;
            if (k0 + kk >= K) continue;  // zero-filled planes stay zero (they meet zero... any dy, result unused)
            const float sc = sc_s[kk], sh = sh_s[kk];
            float* dst = in_s + kk * IN_PLANE + r * IN_STRIDE;
            if (j < 8) {
                float4* q = reinterpret_cast<float4*>(dst + IN_X0 + 1 + 4 * j);
                float4 v = *q;
                v.x = sifnn::act_affine_relu(v.x, sc, sh); v.y = sifnn::act_affine_relu(v.y, sc, sh);
                v.z = sifnn::act_affine_relu(v.z, sc, sh); v.w = sifnn::act_affine_relu(v.w, sc, sh);
                *q = v;
            } else {
                float* q = dst + (j == 8 ? IN_X0 : IN_X0 + 33);
                *q = sifnn::act_affine_relu(*q, sc, sh);
            }
        }
    } else {
        for (int idx = tid; idx < K_CHUNK * IN_ROWS * IN_COLS; idx += NT) {
            const int kk = idx / (IN_ROWS * IN_COLS);
            const int rem = idx - kk * (IN_ROWS * IN_COLS);
            const int r = rem / IN_COLS, c = rem - r * IN_COLS;
            if (k0 + kk >= K) continue;
            float* q = in_s + kk * IN_PLANE + r * IN_STRIDE + IN_X0 + c;
            *q = sifnn::act_affine_relu(*q, sc_s[kk], sh_s[kk]);
        }
    }
}

template <int OT, int KT, int O_CHUNK, int K_CHUNK, bool AFFINE, bool BIAS>
__global__ void __launch_bounds__((O_CHUNK / OT) * (K_CHUNK / KT) * 32, 1) wgrad_kernel(const WgradArgs a) {
    constexpr int N_OT = O_CHUNK / OT;
    constexpr int N_KT = K_CHUNK / KT;
    constexpr int NT = N_OT * N_KT * 32;
    constexpr int IN_ROWS = WROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int DY_PLANE = WROWS * 32;
    constexpr int STAGE = K_CHUNK * IN_PLANE + O_CHUNK * DY_PLANE;

    extern __shared__ __align__(16) float smem[];
    __shared__ float sc_s[K_CHUNK], sh_s[K_CHUNK];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ot = warp % N_OT, kt = warp / N_OT;
    const int s = blockIdx.x, S = gridDim.x;
    const int k0 = blockIdx.y * K_CHUNK, o0 = blockIdx.z * O_CHUNK;
    const int K = a.K, O = a.O;

    if (AFFINE && tid < K_CHUNK) {
        sc_s[tid] = (k0 + tid < K) ? __ldg(a.in_scale + k0 + tid) : 0.f;
        sh_s[tid] = (k0 + tid < K) ? __ldg(a.in_shift + k0 + tid) : 0.f;
    }

    float acc[OT][KT][9];
    float bsum[OT];
#pragma unroll
    for (int j = 0; j < OT; ++j) {
        bsum[j] = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[j][k][t] = 0.f;
    }

    auto tile_origin = [&](int tile, int& b, int& x0, int& y0) {
        b = tile / a.tiles_per_img;
        const int t = tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        x0 = tx * 32;
        y0 = ty * WROWS;
    };

    int it = 0;
    if (s < a.total_tiles) {
        int b, x0, y0;
        tile_origin(s, b, x0, y0);
        wgrad_issue_fill<O_CHUNK, K_CHUNK, NT>(smem, smem + K_CHUNK * IN_PLANE, a, b, x0, y0, k0, o0, a.vec_ok && x0 + 32 <= a.W, tid);
    }
    sifnn::cp_async_commit();
    for (int tile = s; tile < a.total_tiles; tile += S, ++it) {
        float* in_s = smem + (it & 1) * STAGE;
        float* dy_s = in_s + K_CHUNK * IN_PLANE;
        int b, x0, y0;
        tile_origin(tile, b, x0, y0);
        const bool vec_ok = a.vec_ok && x0 + 32 <= a.W;
        if (tile + S < a.total_tiles) {
            int nb, nx0, ny0;
            tile_origin(tile + S, nb, nx0, ny0);
            float* nin = smem + ((it + 1) & 1) * STAGE;
            wgrad_issue_fill<O_CHUNK, K_CHUNK, NT>(nin, nin + K_CHUNK * IN_PLANE, a, nb, nx0, ny0, k0, o0, a.vec_ok && nx0 + 32 <= a.W, tid);
            sifnn::cp_async_commit();
            sifnn::cp_async_wait<1>();
        } else {
            sifnn::cp_async_wait<0>();
        }
        if (AFFINE) {
            if (it == 0) __syncthreads();  // sc_s / sh_s
            wgrad_affine_pass<K_CHUNK, NT>(in_s, sc_s, sh_s, K, k0, vec_ok, tid);
        }
        __syncthreads();

        const float* ip = in_s + (kt * KT) * IN_PLANE + IN_X0 + lane;
        const float* dp = dy_s + (ot * OT) * DY_PLANE + lane;
        float win[KT][3][3];
#pragma unroll
        for (int k = 0; k < KT; ++k)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                win[k][0][kx] = ip[k * IN_PLANE + 0 * IN_STRIDE + kx];
                win[k][1][kx] = ip[k * IN_PLANE + 1 * IN_STRIDE + kx];
            }
#pragma unroll
        for (int r = 0; r < WROWS; ++r) {
#pragma unroll
            for (int k = 0; k < KT; ++k)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) win[k][2][kx] = ip[k * IN_PLANE + (r + 2) * IN_STRIDE + kx];
            float d[OT];
#pragma unroll
            for (int j = 0; j < OT; ++j) d[j] = dp[j * DY_PLANE + r * 32];
#pragma unroll
            for (int j = 0; j < OT; ++j) {
                if (BIAS) bsum[j] += d[j];
#pragma unroll
                for (int k = 0; k < KT; ++k)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) acc[j][k][ky * 3 + kx] = fmaf(d[j], win[k][ky][kx], acc[j][k][ky * 3 + kx]);
            }
#pragma unroll
            for (int k = 0; k < KT; ++k)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    win[k][0][kx] = win[k][1][kx];
                    win[k][1][kx] = win[k][2][kx];
                }
        }
        __syncthreads();  // stage may be refilled by the prefetch of the iteration after next
    }

#pragma unroll
    for (int j = 0; j < OT; ++j) {
        const int o = o0 + ot * OT + j;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int kk = k0 + kt * KT + k;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float v = sifnn::warp_sum(acc[j][k][t]);
                if (lane == 0 && o < O && kk < K) a.partial[(((size_t)s * O + o) * K + kk) * 9 + t] = v;
            }
        }
        if (BIAS) {
            const float v = sifnn::warp_sum(bsum[j]);
            if (lane == 0 && kt == 0 && blockIdx.y == 0 && o < O) a.bias_partial[(size_t)s * O + o] = v;
        }
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += partial[(size_t)s * n + i];
    out[i] = acc;
}

struct Plan {
    int variant;  // 0 general, 1 small-K, 2 single-output (outlay)
    int o_chunks, k_chunks, S, tiles_x, tiles_y, total_tiles;
};

Plan make_plan(int B, int K, int O, int H, int W) {
    Plan p{};
    int oc, kc;
    if (O == 1) { p.variant = 2; oc = 1; kc = 16; }
    else if (K <= 2) { p.variant = 1; oc = 16; kc = 2; }
    else { p.variant = 0; oc = 16; kc = 8; }
    p.o_chunks = (O + oc - 1) / oc;
    p.k_chunks = (K + kc - 1) / kc;
    p.tiles_x = (W + 31) / 32;
    p.tiles_y = (H + WROWS - 1) / WROWS;
    p.total_tiles = B * p.tiles_x * p.tiles_y;
    const int pairs = p.o_chunks * p.k_chunks;
    int S = sifnn::num_sms() / pairs;
    if (S < 1) S = 1;
    if (S > p.total_tiles) S = p.total_tiles;
    p.S = S;
    return p;
}

template <int OT, int KT, int O_CHUNK, int K_CHUNK, bool BIAS>
int launch_wgrad(const WgradArgs& a, const Plan& p, bool affine, cudaStream_t st) {
    constexpr int NT = (O_CHUNK / OT) * (K_CHUNK / KT) * 32;
    constexpr size_t smem = 2 * (size_t)(K_CHUNK * (WROWS + 2) * IN_STRIDE + O_CHUNK * WROWS * 32) * sizeof(float);  // two stages
    auto k_aff = wgrad_kernel<OT, KT, O_CHUNK, K_CHUNK, true, BIAS>;
    auto k_pln = wgrad_kernel<OT, KT, O_CHUNK, K_CHUNK, false, BIAS>;
    static bool attr_done = false;
    if (!attr_done) {
        SIFNN_CUDA(cudaFuncSetAttribute(k_aff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SIFNN_CUDA(cudaFuncSetAttribute(k_pln, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid(p.S, p.k_chunks, p.o_chunks);
    if (affine) k_aff<<<grid, NT, smem, st>>>(a);
    else k_pln<<<grid, NT, smem, st>>>(a);
    return sifnn::check_launch("wgrad_kernel");
}

}  // namespace

extern "C" size_t sifnn_conv3x3_wgrad_workspace(int B, int Cin, int Cout, int H, int W) {
    if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return 0;
    const Plan p = make_plan(B, Cin, Cout, H, W);
    return ((size_t)p.S * Cout * Cin * 9 + (size_t)p.S * Cout) * sizeof(float);
}

extern "C" int sifnn_conv3x3_wgrad(const float* in, const float* in_scale, const float* in_shift, const float* dy,
                                   float* dw, float* dbias, void* workspace, int B, int Cin, int Cout, int H, int W,
                                   sifnn_stream_t stream) {
    SIFNN_REQUIRE(in && dy && dw && workspace, "conv3x3_wgrad: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_wgrad: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "conv3x3_wgrad: bad shape");
    SIFNN_REQUIRE(dbias == nullptr || Cout == 1, "conv3x3_wgrad: dbias is only produced for Cout == 1 (outlay)");
    const Plan p = make_plan(B, Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    WgradArgs a{};
    a.in = in; a.in_scale = in_scale; a.in_shift = in_shift; a.dy = dy;
    a.partial = static_cast<float*>(workspace);
    a.bias_partial = a.partial + (size_t)p.S * Cout * Cin * 9;
    a.B = B; a.K = Cin; a.O = Cout; a.H = H; a.W = W;
    a.tiles_x = p.tiles_x; a.tiles_per_img = p.tiles_x * p.tiles_y; a.total_tiles = p.total_tiles;
    a.vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    const bool affine = in_scale != nullptr;
    if (p.variant == 2) SIFNN_TRY((launch_wgrad<1, 1, 1, 16, true>(a, p, affine, st)));
    else if (p.variant == 1) SIFNN_TRY((launch_wgrad<2, 1, 16, 2, false>(a, p, affine, st)));
    else SIFNN_TRY((launch_wgrad<4, 2, 16, 8, false>(a, p, affine, st)));
    const int n = Cout * Cin * 9;
    wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(a.partial, dw, n, p.S);
    SIFNN_TRY(sifnn::check_launch("wgrad_reduce_kernel"));
    if (dbias) {
        wgrad_reduce_kernel<<<1, 32, 0, st>>>(a.bias_partial, dbias, Cout, p.S);
        SIFNN_TRY(sifnn::check_launch("wgrad_reduce_kernel(bias)"));
    }
    return 0;
}
