// Fused SIF-NN-SR losses: forward scalars AND dLoss/dSR in one pass over the images.
//
//   ds     = Huber( bicubic/4( G_0.1 * reflect4(SR) ) , LST )                 (both variants)
//            train_model_B_gradFTM.py:99-106, utils.py:1671-1706  (K12)
//   SR1 pl = Huber( Sobel4(SR) , gamma * Sobel4(NDVI) ), zero 'same' padding  (K11)
//            train_model_B_predef_filters.py:38-42,120-130
//   SR2 pl = Huber( SR - G_0.25*SR , gamma * (NDVI - G_0.25*NDVI) ), reflect4 (K13)
//            train_model_B_gradFTM.py:108-114, utils.py:1833-1860
//   loss   = alpha*ds + (1-alpha)*pl                                           (K14)
//
// Every operator is linear in SR, so pl is evaluated on x = SR - gamma*NDVI, and the
// un-normalise / re-normalise pair around the down-scaling (weights sum to one) is
// folded away: ds is evaluated directly on the normalised SR.  The 9x9 PSFs are rank-1,
// and blur + bicubic/4 (taps [-3,19,19,-3]/32, stride 4) collapses to one separable
// 12-tap stride-4 filter h = d (*) g  (SURVEY section 2.1).  The adjoints of the
// reflect-padded operators are applied in gather form from small host-built per-axis
// tables, so no atomics are needed for the gradient.
//
// One CTA = one 64x64 tile of one image; halo values are recomputed from L2-resident
// data, so DRAM traffic stays at the algorithmic read SR + NDVI + LST, write dSR.
#include "common.cuh"

namespace {

constexpr int T = 64;           // output tile
constexpr int HALO = 8;
constexpr int TS = T + 2 * HALO;  // 80: staged SR / x tile
constexpr int NI = T / 4 + 2;     // 18 low-res rows/cols touched by a tile
constexpr int LNT = 1024;         // threads per CTA: the 128 KB tile buffers allow one CTA per SM, so the CTA itself has to fill the SM --
                                  // with 256 threads the kernel ran at 5 % of the HBM roofline (84 us for 25.7 MB), every phase waiting on 8 warps

__device__ __forceinline__ int reflect_idx(int p, int n) { return p < 0 ? -p : (p >= n ? 2 * (n - 1) - p : p); }
__device__ __forceinline__ float huber(float d) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float huber_grad(float d) { return fminf(fmaxf(d, -1.f), 1.f); }

struct LossArgs {
    const float* sr;
    const float* ndvi;
    const float* lst;
    const float* tab_ds;   // [H][3]
    const float* h12;      // [12]
    const float* tab_lp;   // [H][9]
    const float* g9;       // [9]
    double* losses;
    float* dsr;
    float alpha, gamma;
    int B, H, W;
};

// the four predefined 3x3 gradient filters of train_model_B_predef_filters.py:38-42 are compile-time tables inside the kernel (SB)

template <int KIND>
__global__ void __launch_bounds__(LNT, 1) loss_kernel(const LossArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* S = sm;                    // [TS][TS]   SR (0 outside the image)
    float* X = S + TS * TS;           // [TS][TS]   SR - gamma*NDVI (0 outside the image)
    float* T1 = X + TS * TS;          // [NI][TS]   row-filtered SR at low-res rows
    float* PD = T1 + NI * TS;         // [NI][NI]   alpha/Nds * huber'(d)  (padded to 328)
    float* PE = PD + 328;             // SR1: [4][66][66]; SR2: [72][72] then TMP [80][72], TMP2 [72][64]
    __shared__ float s_h[12], s_g[9];
    __shared__ float red[2][LNT / 32];

    const int tid = threadIdx.x;
    const int H = a.H, W = a.W, h = H / 4, w = W / 4;
    const int tiles_x = W / T;
    const int b = blockIdx.y;
    const int r0 = (blockIdx.x / tiles_x) * T, c0 = (blockIdx.x % tiles_x) * T;
    const int I0 = r0 / 4 - 1, J0 = c0 / 4 - 1;
    const float* sr = a.sr + (size_t)b * H * W;
    const float* nd = a.ndvi + (size_t)b * H * W;
    const float alpha = a.alpha, gamma = a.gamma;
    const float c_ds = alpha / ((float)a.B * (float)h * (float)w);
    const float n_p = (KIND == 1 ? 4.f : 1.f) * (float)a.B * (float)H * (float)W;
    const float c_p = (1.f - alpha) / n_p;

    if (tid < 12) s_h[tid] = a.h12[tid];
    if (KIND == 2 && tid < 9) s_g[tid] = a.g9[tid];

    // ---- P0: stage SR and x ------------------------------------------------------------------
    for (int idx = tid; idx < TS * TS; idx += LNT) {
        const int rr = idx / TS, cc = idx - rr * TS;
        const int r = r0 - HALO + rr, c = c0 - HALO + cc;
        float s = 0.f, x = 0.f;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            s = __ldg(sr + (size_t)r * W + c);
            x = fmaf(-gamma, __ldg(nd + (size_t)r * W + c), s);
        }
        S[idx] = s;
        X[idx] = x;
    }
    __syncthreads();

    float acc_ds = 0.f, acc_p = 0.f;

    // ---- P1: ds row pass  T1[i][cc] = sum_t h[t] * SR[refl(4I-4+t)][c] -------------------------------
    for (int idx = tid; idx < NI * TS; idx += LNT) {
        const int i = idx / TS, cc = idx - i * TS;
        const int I = I0 + i, c = c0 - HALO + cc;
        float v = 0.f;
        if (I >= 0 && I < h && c >= 0 && c < W) {
#pragma unroll
            for (int t = 0; t < 12; ++t) v = fmaf(s_h[t], S[(reflect_idx(4 * I - 4 + t, H) - (r0 - HALO)) * TS + cc], v);
        }
        T1[idx] = v;
    }
    __syncthreads();
    // ---- P2: ds column pass, Huber, huber' -------------------------------------------------------
    for (int idx = tid; idx < NI * NI; idx += LNT) {
        const int i = idx / NI, j = idx - i * NI;
        const int I = I0 + i, J = J0 + j;
        float psi = 0.f;
        if (I >= 0 && I < h && J >= 0 && J < w) {
            float v = 0.f;
#pragma unroll
            for (int t = 0; t < 12; ++t) v = fmaf(s_h[t], T1[i * TS + reflect_idx(4 * J - 4 + t, W) - (c0 - HALO)], v);
            const float d = v - __ldg(a.lst + ((size_t)b * h + I) * w + J);
            psi = c_ds * huber_grad(d);
            if (i >= 1 && i <= T / 4 && j >= 1 && j <= T / 4) acc_ds += huber(d);
        }
        PD[idx] = psi;
    }

    // ---- P3: perceptual residual e and huber'(e) ----------------------------------------------------
    if (KIND == 1) {
        constexpr int E = T + 2;  // 66
        // one item = one pixel of the 66x66 halo region, all four filters: the nine x values are read once and the filter coefficients are
        // compile-time constants (half of them zero), instead of one item per (filter, pixel) with dynamically indexed coefficients
        for (int idx = tid; idx < E * E; idx += LNT) {
            const int rr = idx / E, cc = idx - rr * E;
            const int r = r0 - 1 + rr, c = c0 - 1 + cc;
            float psi[4] = {0.f, 0.f, 0.f, 0.f};
            if (r >= 0 && r < H && c >= 0 && c < W) {
                const float* xp = X + (rr + HALO - 2) * TS + (cc + HALO - 2);  // x[r-1][c-1]
                float xv[9];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) xv[ky * 3 + kx] = xp[ky * TS + kx];
                const bool inner = rr >= 1 && rr <= T && cc >= 1 && cc <= T;
                constexpr float SB[4][9] = {{1, 2, 1, 0, 0, 0, -1, -2, -1}, {1, 0, -1, 2, 0, -2, 1, 0, -1}, {2, 1, 0, 1, 0, -1, 0, -1, -2}, {0, 1, 2, -1, 0, 1, -2, -1, 0}};
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    float e = 0.f;
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        if (SB[f][t] != 0.f) e = fmaf(SB[f][t], xv[t], e);
                    psi[f] = c_p * huber_grad(e);
                    if (inner) acc_p += huber(e);
                }
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) PE[f * E * E + idx] = psi[f];
        }
    } else {
        constexpr int E = T + 8;  // 72
        float* TMP = PE + E * E;  // [TS][E]
        // row-direction blur: TMP[rr][cc] = sum_n g[n] * x[r][refl(c+n)],  r = r0-8+rr, c = c0-4+cc
        for (int idx = tid; idx < TS * E; idx += LNT) {
            const int rr = idx / E, cc = idx - rr * E;
            const int r = r0 - HALO + rr, c = c0 - 4 + cc;
            float v = 0.f;
            if (r >= 0 && r < H && c >= 0 && c < W) {
#pragma unroll
                for (int n = 0; n < 9; ++n) v = fmaf(s_g[n], X[rr * TS + reflect_idx(c + n - 4, W) - (c0 - HALO)], v);
            }
            TMP[idx] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < E * E; idx += LNT) {
            const int rr = idx / E, cc = idx - rr * E;
            const int r = r0 - 4 + rr, c = c0 - 4 + cc;
            float psi = 0.f;
            if (r >= 0 && r < H && c >= 0 && c < W) {
                float v = 0.f;
#pragma unroll
                for (int m = 0; m < 9; ++m) v = fmaf(s_g[m], TMP[(reflect_idx(r + m - 4, H) - (r0 - HALO)) * E + cc], v);
                const float e = X[(rr + 4) * TS + (cc + 4)] - v;
                psi = c_p * huber_grad(e);
                if (rr >= 4 && rr < T + 4 && cc >= 4 && cc < T + 4) acc_p += huber(e);
            }
            PE[idx] = psi;
        }
    }
    __syncthreads();

    // ---- loss scalars ----------------------------------------------------------------------------------
    {
        const float t_ds = sifnn::warp_sum(acc_ds), t_p = sifnn::warp_sum(acc_p);
        if ((tid & 31) == 0) { red[0][tid >> 5] = t_ds; red[1][tid >> 5] = t_p; }
        __syncthreads();
        if (tid == 0) {
            double d = 0.0, p = 0.0;
            for (int i = 0; i < LNT / 32; ++i) { d += (double)red[0][i]; p += (double)red[1][i]; }
            d /= (double)a.B * h * w;
            p /= (double)n_p;
            atomicAdd(a.losses + 0, d);
            atomicAdd(a.losses + 1, p);
            atomicAdd(a.losses + 2, (double)alpha * d + (1.0 - (double)alpha) * p);
        }
    }
    if (a.dsr == nullptr) return;

    // ---- P4: gradient ------------------------------------------------------------------------------------
    float* out = a.dsr + (size_t)b * H * W;
    if (KIND == 2) {
        constexpr int E = T + 8;
        float* TMP2 = PE + E * E + TS * E;  // [E][T]
        // TMP2[rr][cc] = sum_j A[c][j] * PE[rr][cc + j],  c = c0 + cc
        for (int idx = tid; idx < E * T; idx += LNT) {
            const int rr = idx / T, cc = idx - rr * T;
            const float* A = a.tab_lp + (size_t)(c0 + cc) * 9;
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) v = fmaf(__ldg(A + j), PE[rr * E + cc + j], v);
            TMP2[idx] = v;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < T * T; idx += LNT) {
        const int rr = idx / T, cc = idx - rr * T;
        const int r = r0 + rr, c = c0 + cc;
        // down-sampling term
        const float* tr = a.tab_ds + (size_t)r * 3;
        const float* tc = a.tab_ds + (size_t)c * 3;
        const int ib = r / 4 - 1 - I0, jb = c / 4 - 1 - J0;  // == rr/4, cc/4
        float g = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) row = fmaf(__ldg(tc + j), PD[(ib + i) * NI + jb + j], row);
            g = fmaf(__ldg(tr + i), row, g);
        }
        if (KIND == 1) {
            constexpr int E = T + 2;
            constexpr float SB[4][9] = {{1, 2, 1, 0, 0, 0, -1, -2, -1}, {1, 0, -1, 2, 0, -2, 1, 0, -1}, {2, 1, 0, 1, 0, -1, 0, -1, -2}, {0, 1, 2, -1, 0, 1, -2, -1, 0}};
#pragma unroll
            for (int f = 0; f < 4; ++f)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        if (SB[f][ky * 3 + kx] != 0.f) g = fmaf(SB[f][ky * 3 + kx], PE[f * E * E + (rr + 2 - ky) * E + (cc + 2 - kx)], g);
        } else {
            constexpr int E = T + 8;
            const float* TMP2 = PE + E * E + TS * E;
            const float* A = a.tab_lp + (size_t)r * 9;
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < 9; ++i) v = fmaf(__ldg(A + i), TMP2[(rr + i) * T + cc], v);
            g += PE[(rr + 4) * E + cc + 4] - v;
        }
        out[(size_t)r * W + c] = g;
    }
}

}  // namespace

extern "C" int sifnn_loss_fwd_bwd(int kind, const float* sr, const float* ndvi, const float* lst, const float* tab_ds,
                                  const float* h12, const float* tab_lp, const float* g9, float alpha, float gamma,
                                  double* losses, float* dsr, int B, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(kind == 1 || kind == 2, "loss: kind must be 1 (SR1) or 2 (SR2)");
    SIFNN_REQUIRE(sr && ndvi && lst && tab_ds && h12 && losses, "loss: null pointer");
    SIFNN_REQUIRE(kind == 1 || (tab_lp && g9), "loss: SR2 needs tab_lp and g9");
    SIFNN_REQUIRE(B > 0 && B <= 65535 && H == W && H >= 64 && H % 64 == 0, "loss: need square images with H %% 64 == 0 (got %dx%d)", H, W);
    LossArgs a{};
    a.sr = sr; a.ndvi = ndvi; a.lst = lst; a.tab_ds = tab_ds; a.h12 = h12; a.tab_lp = tab_lp; a.g9 = g9;
    a.losses = losses; a.dsr = dsr; a.alpha = alpha; a.gamma = gamma; a.B = B; a.H = H; a.W = W;
    constexpr int PE1 = 4 * 66 * 66;
    constexpr int PE2 = 72 * 72 + 80 * 72 + 72 * 64;
    constexpr size_t smem = (size_t)(2 * TS * TS + NI * TS + 328 + (PE1 > PE2 ? PE1 : PE2)) * sizeof(float);
    static sifnn::PerDeviceOnce attr_once;   // the attribute is per device: one flag per device, not one per process
    if (attr_once.first_time()) {
        SIFNN_CUDA(cudaFuncSetAttribute(loss_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SIFNN_CUDA(cudaFuncSetAttribute(loss_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid((H / T) * (W / T), B);
    if (kind == 1) loss_kernel<1><<<grid, LNT, smem, sifnn::as_stream(stream)>>>(a);
    else loss_kernel<2><<<grid, LNT, smem, sifnn::as_stream(stream)>>>(a);
    return sifnn::check_launch("loss_kernel");
}
