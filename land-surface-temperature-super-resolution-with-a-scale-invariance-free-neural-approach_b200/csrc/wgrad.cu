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
constexpr int WGRAD_STAGES = 2;
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
// Fast path: every thread owns at most ONE fixed slot -- a 16 B column group of an input row
// (threads 0..143), a halo element (160..195) or a 16 B group of a dy row (256..383) -- and
// walks it over the chunk's channel planes, so the per-tile index arithmetic is a handful of
// instructions.  The generic per-element path handles tiles that stick out of the image in x.
__device__ __forceinline__ int wgrad_role(int tid) { return tid < 144 ? 1 : ((tid >= 160 && tid < 196) ? 2 : ((tid >= 256 && tid < 384) ? 3 : 0)); }

template <int O_CHUNK, int K_CHUNK>
__device__ __forceinline__ void wgrad_fill_vec(float* in_s, float* dy_s, const WgradArgs& a, int b, int x0, int y0, int k0, int o0, int tid) {
    constexpr int IN_ROWS = WROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int DY_PLANE = WROWS * 32;
    const int H = a.H, W = a.W, K = a.K, O = a.O;
    const size_t plane = (size_t)H * W;
    const int role = wgrad_role(tid);
    if (role == 1 || role == 2) {
        int r, dcol, gx;
        if (role == 1) { r = tid >> 3; const int j = tid & 7; dcol = IN_X0 + 1 + 4 * j; gx = x0 + 4 * j; }
        else { const int ps = tid - 160; r = ps >> 1; const int right = ps & 1; dcol = right ? IN_X0 + 33 : IN_X0; gx = right ? min(x0 + 32, W - 1) : max(x0 - 1, 0); }
        const int gy = min(max(y0 + r - 1, 0), H - 1);
        const int nk = min(K_CHUNK, K - k0);
        const float* src = a.in + ((size_t)b * K + k0) * plane + (size_t)gy * W + gx;
        float* dst = in_s + r * IN_STRIDE + dcol;
#pragma unroll
        for (int kk = 0; kk < K_CHUNK; ++kk) {
            const bool ok = kk < nk;
            if (role == 1) sifnn::cp_async16(dst, ok ? src : a.in, ok ? 16 : 0);
            else sifnn::cp_async4(dst, ok ? src : a.in, ok ? 4 : 0);
            dst += IN_PLANE;
            src += plane;
        }
    } else if (role == 3) {
        const int t = tid - 256;
        const int r = t >> 3, j = t & 7;
        const bool rok = y0 + r < H;
        const int no = min(O_CHUNK, O - o0);
        const float* src = a.dy + ((size_t)b * O + o0) * plane + (size_t)(rok ? y0 + r : 0) * W + x0 + 4 * j;
        float* dst = dy_s + r * 32 + 4 * j;
#pragma unroll
        for (int oo = 0; oo < O_CHUNK; ++oo) {
            const bool ok = rok && oo < no;
            sifnn::cp_async16(dst, ok ? src : a.dy, ok ? 16 : 0);
            dst += DY_PLANE;
            src += plane;
        }
    }
}

template <int K_CHUNK>
__device__ __forceinline__ void wgrad_affine_vec(float* in_s, const float* sc_s, const float* sh_s, int K, int k0, int tid) {
    constexpr int IN_ROWS = WROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    const int role = wgrad_role(tid);
    const int nk = min(K_CHUNK, K - k0);
    if (role == 1) {
        float* dst = in_s + (tid >> 3) * IN_STRIDE + IN_X0 + 1 + 4 * (tid & 7);
#pragma unroll
        for (int kk = 0; kk < K_CHUNK; ++kk) {
            if (kk < nk) {
                float4 v = *reinterpret_cast<float4*>(dst);
                const float c = sc_s[kk], h = sh_s[kk];
                v.x = sifnn::act_affine_relu(v.x, c, h); v.y = sifnn::act_affine_relu(v.y, c, h);
                v.z = sifnn::act_affine_relu(v.z, c, h); v.w = sifnn::act_affine_relu(v.w, c, h);
                *reinterpret_cast<float4*>(dst) = v;
            }
            dst += IN_PLANE;
        }
    } else if (role == 2) {
        const int ps = tid - 160;
        float* dst = in_s + (ps >> 1) * IN_STRIDE + ((ps & 1) ? IN_X0 + 33 : IN_X0);
#pragma unroll
        for (int kk = 0; kk < K_CHUNK; ++kk) {
            if (kk < nk) *dst = sifnn::act_affine_relu(*dst, sc_s[kk], sh_s[kk]);
            dst += IN_PLANE;
        }
    }
}

template <int O_CHUNK, int K_CHUNK, int NT>
__device__ __forceinline__ void wgrad_issue_fill(float* in_s, float* dy_s, const WgradArgs& a, int b, int x0, int y0, int k0, int o0,
                                                 bool vec_ok, int tid) {
    constexpr int IN_ROWS = WROWS + 2;
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int DY_PLANE = WROWS * 32;
    const int H = a.H, W = a.W, K = a.K, O = a.O;
    const size_t plane = (size_t)H * W;
    if (vec_ok) {
        wgrad_fill_vec<O_CHUNK, K_CHUNK>(in_s, dy_s, a, b, x0, y0, k0, o0, tid);
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
        wgrad_affine_vec<K_CHUNK>(in_s, sc_s, sh_s, K, k0, tid);
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
    sifnn::f32x2_t acc2[OT][9];  // KT == 2: (k0, k1) pairs, used instead of acc
    float bsum[OT];
#pragma unroll
    for (int j = 0; j < OT; ++j) {
        bsum[j] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc2[j][t] = 0ull;
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

    // NST-deep cp.async ring: the tiles it+1 .. it+NST-1 are in flight while tile `it` is consumed (one tile of
    // compute is ~2.4 us at full FMA rate, about one DRAM round trip under load, so a single tile of look-ahead
    // does not cover it).  One commit group per tile, empty groups past the end keep the wait counts uniform.
    constexpr int NST = WGRAD_STAGES;
    auto issue = [&](int tile, int slot) {
        if (tile < a.total_tiles) {
            int b, x0, y0;
            tile_origin(tile, b, x0, y0);
            float* nin = smem + slot * STAGE;
            wgrad_issue_fill<O_CHUNK, K_CHUNK, NT>(nin, nin + K_CHUNK * IN_PLANE, a, b, x0, y0, k0, o0, a.vec_ok && x0 + 32 <= a.W, tid);
        }
        sifnn::cp_async_commit();
    };
    int it = 0;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) issue(s + p * S, p);
    for (int tile = s; tile < a.total_tiles; tile += S, ++it) {
        const int slot = it % NST;
        float* in_s = smem + slot * STAGE;
        float* dy_s = in_s + K_CHUNK * IN_PLANE;
        int b, x0, y0;
        tile_origin(tile, b, x0, y0);
        const bool vec_ok = a.vec_ok && x0 + 32 <= a.W;
        issue(tile + (NST - 1) * S, (it + NST - 1) % NST);   // refills the slot consumed in the previous iteration
        sifnn::cp_async_wait<NST - 1>();
        if (AFFINE) {
            if (it == 0) __syncthreads();  // sc_s / sh_s
            wgrad_affine_pass<K_CHUNK, NT>(in_s, sc_s, sh_s, K, k0, vec_ok, tid);
        }
        __syncthreads();

        const float* ip = in_s + (kt * KT) * IN_PLANE + IN_X0 + lane;
        const float* dp = dy_s + (ot * OT) * DY_PLANE + lane;
        if (KT == 2) {
            // packed path (FFMA2): each accumulator pair holds the two input channels of this warp's k-tile;
            // dy is the scalar-broadcast operand.
            sifnn::f32x2_t win2[3][3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                win2[0][kx] = sifnn::pack2(ip[0 * IN_STRIDE + kx], ip[IN_PLANE + 0 * IN_STRIDE + kx]);
                win2[1][kx] = sifnn::pack2(ip[1 * IN_STRIDE + kx], ip[IN_PLANE + 1 * IN_STRIDE + kx]);
            }
#pragma unroll
            for (int r = 0; r < WROWS; ++r) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) win2[2][kx] = sifnn::pack2(ip[(r + 2) * IN_STRIDE + kx], ip[IN_PLANE + (r + 2) * IN_STRIDE + kx]);
                float d[OT];
#pragma unroll
                for (int j = 0; j < OT; ++j) d[j] = dp[j * DY_PLANE + r * 32];
                if (BIAS) {
#pragma unroll
                    for (int j = 0; j < OT; ++j) bsum[j] += d[j];
                }
                // window-stationary order: the 64-bit window pair stays in the operand-reuse cache across the OT
                // dy scalars (see tools/ffma2_probe2.cu)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int j = 0; j < OT; ++j)
                            acc2[j][ky * 3 + kx] = sifnn::fma2(sifnn::pack2(d[j], d[j]), win2[ky][kx], acc2[j][ky * 3 + kx]);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    win2[0][kx] = win2[1][kx];
                    win2[1][kx] = win2[2][kx];
                }
            }
        } else {
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
                float av = acc[j][k][t];
                if (KT == 2) {
                    float lo, hi;
                    sifnn::unpack2(acc2[j][t], lo, hi);
                    av = k ? hi : lo;
                }
                const float v = sifnn::warp_sum(av);
                if (lane == 0 && o < O && kk < K) a.partial[(((size_t)s * O + o) * K + kk) * 9 + t] = v;
            }
        }
        if (BIAS) {
            const float v = sifnn::warp_sum(bsum[j]);
            if (lane == 0 && kt == 0 && blockIdx.y == 0 && o < O) a.bias_partial[(size_t)s * O + o] = v;
        }
    }
}

// out[i] = sum_s partial[s][i] in a fixed order (deterministic): 32 slice-strided sums per output, combined in order.  Outputs
// n .. n + nb - 1 are the bias gradient (own partial array), so one launch finishes both.
__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int S,
                                                           const float* __restrict__ bias_partial, float* __restrict__ bias_out, int nb) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    const float* src = i < n ? partial + i : bias_partial + (i - n);
    const int stride = i < n ? n : nb;
    float acc = 0.f;
    if (i < n + nb)
        for (int s = ty; s < S; s += 32) acc += __ldg(src + (size_t)s * stride);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && i < n + nb) {
        float t = red[0][tx];
#pragma unroll
        for (int j = 1; j < 32; ++j) t += red[j][tx];
        if (i < n) out[i] = t;
        else bias_out[i - n] = t;
    }
}

// ---- one output channel (the network's `outlay`): dW[k][ky][kx] = sum dy[y][x] * act(in)[k][clamp(y+ky-1)][clamp(x+kx-1)] ----
// The register-tiled kernel above needs many (o, k) pairs per warp to amortise its shared-memory reads; with one output channel
// it is load-bound.  Here a thread owns 4 columns x 8 rows of dy (kept in registers for all input channels) and, per channel,
// slides over the 10 input rows straight from global memory like conv3x3_to1_kernel: 288 FMAs into 9 sums, which are reduced
// over the warp by shuffles and added to the warp's row of shared accumulators by lane 0.  CTAs stride over the pixel blocks; a
// CTA's 4 warp rows are summed in order into its partial, the partials by wgrad_reduce_kernel: deterministic, no atomics.
constexpr int W1_MAXK = 64;
constexpr int W1_ROWS = 8;

template <bool AFFINE>
__global__ void __launch_bounds__(128, 4) wgrad_to1_kernel(const WgradArgs a) {
    __shared__ float red[4][W1_MAXK * 9 + 1];
    __shared__ float sc_s[W1_MAXK], sh_s[W1_MAXK];
    const int K = a.K, H = a.H, W = a.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4 * (W1_MAXK * 9 + 1); i += 128) (&red[0][0])[i] = 0.f;
    if (AFFINE)
        for (int i = tid; i < K; i += 128) { sc_s[i] = __ldg(a.in_scale + i); sh_s[i] = __ldg(a.in_shift + i); }
    __syncthreads();
    const int w4 = W >> 2;
    const int bands = (H + W1_ROWS - 1) / W1_ROWS;
    const int total = a.B * bands * w4;
    const size_t plane = (size_t)H * W;
    float bsum = 0.f;
    for (int base = blockIdx.x * 128; base < total; base += gridDim.x * 128) {
        const int u = base + tid;
        const bool live = u < total;
        const int uu = live ? u : 0;
        const int x0 = (uu % w4) << 2;
        const int band = (uu / w4) % bands;
        const int b = uu / (w4 * bands);
        const int y0 = band * W1_ROWS;
        const float* dyb = a.dy + (size_t)b * plane;
        float d[W1_ROWS][4];
#pragma unroll
        for (int r = 0; r < W1_ROWS; ++r) {
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && y0 + r < H) m = __ldg(reinterpret_cast<const float4*>(dyb + (size_t)(y0 + r) * W + x0));
            d[r][0] = m.x; d[r][1] = m.y; d[r][2] = m.z; d[r][3] = m.w;
            bsum += (m.x + m.y) + (m.z + m.w);
        }
        const int xl = max(x0 - 1, 0), xr = min(x0 + 4, W - 1);
        int rowoff[W1_ROWS + 2];
#pragma unroll
        for (int r = 0; r < W1_ROWS + 2; ++r) rowoff[r] = min(max(y0 - 1 + r, 0), H - 1) * W;
        const float* in_b = a.in + (size_t)b * K * plane;
#pragma unroll 1
        for (int ci = 0; ci < K; ++ci) {
            const float* ip = in_b + (size_t)ci * plane;
            const float sc = AFFINE ? sc_s[ci] : 1.f, sh = AFFINE ? sh_s[ci] : 0.f;
            float v[W1_ROWS + 2][6];
#pragma unroll
            for (int r = 0; r < W1_ROWS + 2; ++r) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(ip + rowoff[r] + x0));
                v[r][0] = __ldg(ip + rowoff[r] + xl);
                v[r][1] = m.x; v[r][2] = m.y; v[r][3] = m.z; v[r][4] = m.w;
                v[r][5] = __ldg(ip + rowoff[r] + xr);
            }
            if (AFFINE) {
#pragma unroll
                for (int r = 0; r < W1_ROWS + 2; ++r)
#pragma unroll
                    for (int i = 0; i < 6; ++i) v[r][i] = sifnn::act_affine_relu(v[r][i], sc, sh);
            }
            float sum[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) sum[t] = 0.f;
#pragma unroll
            for (int r = 0; r < W1_ROWS; ++r)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int i = 0; i < 4; ++i) sum[ky * 3 + kx] = fmaf(d[r][i], v[r + ky][i + kx], sum[ky * 3 + kx]);
#pragma unroll
            for (int t = 0; t < 9; ++t) sum[t] = sifnn::warp_sum(sum[t]);
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < 9; ++t) red[warp][ci * 9 + t] += sum[t];
            }
        }
    }
    bsum = sifnn::warp_sum(bsum);
    if (lane == 0) red[warp][W1_MAXK * 9] = bsum;
    __syncthreads();
    const int n = K * 9;
    for (int i = tid; i < n; i += 128)
        a.partial[(size_t)blockIdx.x * n + i] = ((red[0][i] + red[1][i]) + red[2][i]) + red[3][i];
    if (tid == 0)
        a.bias_partial[blockIdx.x] = ((red[0][W1_MAXK * 9] + red[1][W1_MAXK * 9]) + red[2][W1_MAXK * 9]) + red[3][W1_MAXK * 9];
}

static bool to1_wgrad_eligible(int K, int O, int W, const float* in, const float* dy) {
    return O == 1 && K <= W1_MAXK && W % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
}

static int to1_wgrad_ctas(int B, int H, int W) {
    const int units = B * ((H + W1_ROWS - 1) / W1_ROWS) * (W / 4);
    return max(1, min(4 * sifnn::num_sms(), (units + 127) / 128));
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
    constexpr size_t smem = WGRAD_STAGES * (size_t)(K_CHUNK * (WROWS + 2) * IN_STRIDE + O_CHUNK * WROWS * 32) * sizeof(float);
    auto k_aff = wgrad_kernel<OT, KT, O_CHUNK, K_CHUNK, true, BIAS>;
    auto k_pln = wgrad_kernel<OT, KT, O_CHUNK, K_CHUNK, false, BIAS>;
    static sifnn::PerDeviceOnce attr_once;   // the attribute is per device: one flag per device, not one per process
    if (attr_once.first_time()) {
        SIFNN_CUDA(cudaFuncSetAttribute(k_aff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SIFNN_CUDA(cudaFuncSetAttribute(k_pln, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    size_t S = (size_t)p.S;
    if (Cout == 1 && Cin <= W1_MAXK && W % 4 == 0) S = max(S, (size_t)to1_wgrad_ctas(B, H, W));   // wgrad_to1_kernel: one partial per CTA
    return (S * Cout * Cin * 9 + S * Cout) * sizeof(float);
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
    int S = p.S;
    if (to1_wgrad_eligible(Cin, Cout, W, in, dy)) {
        S = to1_wgrad_ctas(B, H, W);
        a.bias_partial = a.partial + (size_t)S * Cin * 9;
        if (affine) wgrad_to1_kernel<true><<<S, 128, 0, st>>>(a);
        else wgrad_to1_kernel<false><<<S, 128, 0, st>>>(a);
        SIFNN_TRY(sifnn::check_launch("wgrad_to1_kernel"));
    } else if (p.variant == 2) SIFNN_TRY((launch_wgrad<1, 1, 1, 16, true>(a, p, affine, st)));
    else if (p.variant == 1) SIFNN_TRY((launch_wgrad<2, 1, 16, 2, false>(a, p, affine, st)));
    else SIFNN_TRY((launch_wgrad<4, 2, 16, 8, false>(a, p, affine, st)));
    const int n = Cout * Cin * 9;
    const int nb = dbias ? Cout : 0;
    wgrad_reduce_kernel<<<(n + nb + 31) / 32, 1024, 0, st>>>(a.partial, dw, n, S, a.bias_partial, dbias, nb);
    SIFNN_TRY(sifnn::check_launch("wgrad_reduce_kernel"));
    return 0;
}
