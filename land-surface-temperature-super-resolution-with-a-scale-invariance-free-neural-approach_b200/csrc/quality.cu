// PSNR / SSIM of a batch on the device: the reference computes them on the host with skimage after two
// device->host copies of (B,1,256,256) per step (utils.py:548-578, train_model_B_gradFTM.py:126-127).
//
//   psnr = mean_i 10 log10(R^2 / mse_i),   R = max(target) - min(target) over the WHOLE batch (utils.py:551)
//   ssim = mean_i mean_{interior} S,        skimage.metrics.structural_similarity defaults: 7x7 uniform window, sample covariance
//          (normalised by 49/48), K1 = 0.01, K2 = 0.03, data_range = R, mean over the map cropped by 3 pixels per side
//
// Three launches: batch min/max; one pass per 32x32 tile of every image (38x38 halo tile in shared memory, separable running
// sums of x, y, x^2, y^2, xy; the tile's squared error rides along); finalize.  HBM-bound: both tensors are read once (+ halo).
#include "common.cuh"

namespace {

__device__ __forceinline__ int float_to_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct QWork { int mn, mx; int pad[2]; double acc[1]; };   // acc: [B][2] = (sum of squared error, sum of S)

__global__ void quality_init_kernel(QWork* w, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { w->mn = float_to_ordered(3.0e38f); w->mx = float_to_ordered(-3.0e38f); }
    if (i < 2 * B) w->acc[i] = 0.0;
}

__global__ void __launch_bounds__(256) quality_minmax_kernel(const float* __restrict__ t, long long n, QWork* w) {
    float mn = 3.0e38f, mx = -3.0e38f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(t + i);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&w->mn, float_to_ordered(mn)); atomicMax(&w->mx, float_to_ordered(mx)); }
}

constexpr int QT = 32;          // SSIM-map tile
constexpr int QH = QT + 6;      // with the 3-pixel window halo

__global__ void __launch_bounds__(256) quality_tile_kernel(const float* __restrict__ pred, const float* __restrict__ targ, QWork* w, int H, int W) {
    __shared__ float xs[QH][QH + 1], ys[QH][QH + 1];
    __shared__ float hs[5][QH][QT + 1];   // horizontal 7-sums of x, y, xx, yy, xy for the 38 rows x 32 centre columns
    const int tid = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * QT, y0 = blockIdx.y * QT;     // tile of centre positions
    const size_t plane = (size_t)H * W;
    const float* p = pred + (size_t)b * plane;
    const float* t = targ + (size_t)b * plane;
    float se = 0.f;
    for (int i = tid; i < QH * QH; i += 256) {
        const int r = i / QH, c = i - r * QH;
        const int gy = y0 + r - 3, gx = x0 + c - 3;
        float a = 0.f, d = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            a = __ldg(t + (size_t)gy * W + gx);    // x = target (first argument of structural_similarity in the reference)
            d = __ldg(p + (size_t)gy * W + gx);
            if (r >= 3 && r < 3 + QT && c >= 3 && c < 3 + QT) { const float e = a - d; se = fmaf(e, e, se); }   // each pixel belongs to one tile
        }
        xs[r][c] = a; ys[r][c] = d;
    }
    __syncthreads();
    for (int i = tid; i < QH * QT; i += 256) {
        const int r = i / QT, c = i - r * QT;
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const float a = xs[r][c + k], d = ys[r][c + k];
            sx += a; sy += d; sxx = fmaf(a, a, sxx); syy = fmaf(d, d, syy); sxy = fmaf(a, d, sxy);
        }
        hs[0][r][c] = sx; hs[1][r][c] = sy; hs[2][r][c] = sxx; hs[3][r][c] = syy; hs[4][r][c] = sxy;
    }
    __syncthreads();
    const float R = ordered_to_float(w->mx) - ordered_to_float(w->mn);
    const double C1 = (0.01 * (double)R) * (0.01 * (double)R), C2 = (0.03 * (double)R) * (0.03 * (double)R);
    double ssum = 0.0;
    for (int i = tid; i < QT * QT; i += 256) {
        const int r = i / QT, c = i - r * QT;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= 3 && gy < H - 3 && gx >= 3 && gx < W - 3) {
            float s[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 7; ++k) acc += hs[q][r + k][c];
                s[q] = acc;
            }
            const double ux = s[0] / 49.0, uy = s[1] / 49.0, uxx = s[2] / 49.0, uyy = s[3] / 49.0, uxy = s[4] / 49.0;
            const double cn = 49.0 / 48.0;
            const double vx = cn * (uxx - ux * ux), vy = cn * (uyy - uy * uy), vxy = cn * (uxy - ux * uy);
            ssum += ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
        }
    }
    ssum = sifnn::warp_sum_d(ssum);
    const double sed = sifnn::warp_sum_d((double)se);
    if ((tid & 31) == 0) { atomicAdd(&w->acc[2 * b], sed); atomicAdd(&w->acc[2 * b + 1], ssum); }
}

__global__ void quality_finalize_kernel(const QWork* w, float* out, int B, int H, int W) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double R = (double)ordered_to_float(w->mx) - (double)ordered_to_float(w->mn);
    double ps = 0.0, ss = 0.0;
    for (int b = 0; b < B; ++b) {
        const double mse = w->acc[2 * b] / ((double)H * W);
        ps += 10.0 * log10(R * R / mse);
        ss += w->acc[2 * b + 1] / ((double)(H - 6) * (W - 6));
    }
    out[0] = (float)(ps / B);
    out[1] = (float)(ss / B);
}

}  // namespace

extern "C" size_t sifnn_quality_workspace_bytes(int B) { return B > 0 ? sizeof(QWork) + sizeof(double) * 2 * (size_t)B : 0; }

extern "C" int sifnn_quality_psnr_ssim(const float* pred, const float* target, float* out2, void* workspace, int B, int H, int W,
                                       sifnn_stream_t stream) {
    SIFNN_REQUIRE(pred && target && out2 && workspace, "quality_psnr_ssim: null pointer");
    SIFNN_REQUIRE(B > 0 && B <= 65535 && H >= 7 && W >= 7, "quality_psnr_ssim: need images of at least 7x7 (got B=%d H=%d W=%d)", B, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    QWork* w = static_cast<QWork*>(workspace);
    quality_init_kernel<<<(2 * B + 255) / 256, 256, 0, st>>>(w, B);
    SIFNN_TRY(sifnn::check_launch("quality_init_kernel"));
    const long long n = (long long)B * H * W;
    const int blocks = (int)((n + 256 * 8 - 1) / (256 * 8) < 4 * sifnn::num_sms() ? (n + 256 * 8 - 1) / (256 * 8) : 4 * sifnn::num_sms());
    quality_minmax_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(target, n, w);
    SIFNN_TRY(sifnn::check_launch("quality_minmax_kernel"));
    dim3 grid((W + QT - 1) / QT, (H + QT - 1) / QT, B);
    quality_tile_kernel<<<grid, 256, 0, st>>>(pred, target, w, H, W);
    SIFNN_TRY(sifnn::check_launch("quality_tile_kernel"));
    quality_finalize_kernel<<<1, 32, 0, st>>>(w, out2, B, H, W);
    return sifnn::check_launch("quality_finalize_kernel");
}
