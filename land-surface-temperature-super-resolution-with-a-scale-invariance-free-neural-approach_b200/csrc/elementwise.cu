// HBM-bound glue of the ModelB hot path: BatchNorm statistics / backward, AvgPool,
// residual add, bilinear(align_corners) up-sample + concat and their adjoints, and the
// bicubic x4 + concat input stage.  All kernels are streaming: coalesced (float4 where
// alignment allows), grid-stride, grid sized as a multiple of the SM count.
#include "common.cuh"
#include <cstdlib>

namespace {

inline int grid_for(long long work_items, int threads, int max_blocks_per_sm = 16) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)sifnn::num_sms() * max_blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ------------------------------------------------------------------------------------------
// BatchNorm (model.py:136,139,508): finalize training statistics / eval affine
// ------------------------------------------------------------------------------------------
__global__ void bn_train_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* running_mean, float* running_var,
                                         float* scale, float* shift, float* save_mean, float* save_invstd, int C, double n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = stats[c] / n;
    double var = stats[C + c] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + 1e-5));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = fmaf(-(float)mean, sc, beta[c]);
    save_mean[c] = (float)mean;
    save_invstd[c] = invstd;
    if (running_mean) {
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        running_mean[c] = (float)(0.9 * (double)running_mean[c] + 0.1 * mean);
        running_var[c] = (float)(0.9 * (double)running_var[c] + 0.1 * unbiased);
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float* scale, float* shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float invstd = 1.0f / sqrtf(rv[c] + 1e-5f);
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = fmaf(-rm[c], sc, beta[c]);
}

// ------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward
// ------------------------------------------------------------------------------------------
// grid = (chunks, C, B); each CTA reduces a slice of one (b, c) plane.
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_kernel(const float* __restrict__ dY, const float* __restrict__ raw,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 double* sums, int C, int HW, int reverse) {
    sifnn::pdl_wait_and_trigger();
    const int c = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    const int b = reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int bx = reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
    const size_t base = ((size_t)b * C + c) * HW;
    const float4* g4 = reinterpret_cast<const float4*>(dY + base);
    const float4* x4 = reinterpret_cast<const float4*>(raw + base);
    float s1 = 0.f, s2 = 0.f;
    const int n4 = HW >> 2;
    for (int i = bx * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 g = __ldg(g4 + i), x = __ldg(x4 + i);
        const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float d = fmaf(xv[j], sc, sh) > 0.f ? gv[j] : 0.f;
            s1 += d;
            s2 = fmaf(d, (xv[j] - mu) * is, s2);
        }
    }
    __shared__ float r1[8], r2[8];
    s1 = sifnn::warp_sum(s1);
    s2 = sifnn::warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double d1 = 0.0, d2 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { d1 += (double)r1[i]; d2 += (double)r2[i]; }
        atomicAdd(sums + c, d1);
        atomicAdd(sums + C + c, d2);
    }
}

__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float* __restrict__ dY, const float* __restrict__ raw,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const double* __restrict__ sums,
                                                                float* dx, float* dgamma, float* dbeta, int C, int HW, double inv_n, int reverse) {
    sifnn::pdl_wait_and_trigger();
    // reverse = 1: walk the tensors from the end.  The reduce pass that ran just before read dY and raw front to back, so their tails are what the
    // 126 MB L2 still holds; on the 256^2 layers (2 x 134 MB) a front-to-back second pass would miss everywhere.
    const int c = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    const int b = reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int bx = reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
    const float m1 = (float)(sums[c] * inv_n), m2 = (float)(sums[C + c] * inv_n);
    const float gi = gamma[c] * is;
    if (b == 0 && bx == 0 && threadIdx.x == 0) {
        if (dbeta) dbeta[c] = (float)sums[c];
        if (dgamma) dgamma[c] = (float)sums[C + c];
    }
    const size_t base = ((size_t)b * C + c) * HW;
    const float4* g4 = reinterpret_cast<const float4*>(dY + base);
    const float4* x4 = reinterpret_cast<const float4*>(raw + base);
    float4* o4 = reinterpret_cast<float4*>(dx + base);
    const int n4 = HW >> 2;
    for (int i = bx * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 g = g4[i], x = __ldg(x4 + i);
        const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w};
        float ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float d = fmaf(xv[j], sc, sh) > 0.f ? gv[j] : 0.f;
            const float xh = (xv[j] - mu) * is;
            ov[j] = gi * (d - m1 - xh * m2);
        }
        o4[i] = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
}

// ------------------------------------------------------------------------------------------
// AvgPool2d(2) of the activated tensor (model.py:504,529) and its adjoint
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_avgpool2_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float* __restrict__ out,
                                                           long long total, int C, int H, int W) {
    sifnn::pdl_wait_and_trigger();
    const int Ho = H >> 1, Wo = W >> 1, Wo2 = Wo >> 1;  // each thread: 2 output pixels = 4 input columns
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int xo2 = (int)(idx % Wo2);
        const int yo = (int)((idx / Wo2) % Ho);
        const long long bc = idx / ((long long)Wo2 * Ho);
        const int c = (int)(bc % C);
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float* p = raw + (size_t)bc * H * W + (size_t)(2 * yo) * W + 4 * xo2;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + W));
        float2 o;
        o.x = 0.25f * (((sifnn::act_affine_relu(a.x, sc, sh) + sifnn::act_affine_relu(a.y, sc, sh)) + sifnn::act_affine_relu(b.x, sc, sh)) + sifnn::act_affine_relu(b.y, sc, sh));
        o.y = 0.25f * (((sifnn::act_affine_relu(a.z, sc, sh) + sifnn::act_affine_relu(a.w, sc, sh)) + sifnn::act_affine_relu(b.z, sc, sh)) + sifnn::act_affine_relu(b.w, sc, sh));
        *reinterpret_cast<float2*>(out + (size_t)bc * Ho * Wo + (size_t)yo * Wo + 2 * xo2) = o;
    }
}

// din (B,C,H,W) (+)= 0.25 * dout (B,C,H/2,W/2) replicated 2x2
__global__ void __launch_bounds__(256) avgpool2_bwd_kernel(const float* __restrict__ dout, float* din, long long total, int H, int W, int accumulate) {
    sifnn::pdl_wait_and_trigger();
    const int Ho = H >> 1, Wo = W >> 1, Wo2 = Wo >> 1;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int xo2 = (int)(idx % Wo2);
        const int yo = (int)((idx / Wo2) % Ho);
        const long long bc = idx / ((long long)Wo2 * Ho);
        const float2 g = __ldg(reinterpret_cast<const float2*>(dout + (size_t)bc * Ho * Wo + (size_t)yo * Wo + 2 * xo2));
        const float gx = 0.25f * g.x, gy = 0.25f * g.y;
        float* p = din + (size_t)bc * H * W + (size_t)(2 * yo) * W + 4 * xo2;
        float4 a = make_float4(gx, gx, gy, gy), b = a;
        if (accumulate) {
            const float4 pa = *reinterpret_cast<float4*>(p), pb = *reinterpret_cast<float4*>(p + W);
            a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
            b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
        }
        *reinterpret_cast<float4*>(p) = a;
        *reinterpret_cast<float4*>(p + W) = b;
    }
}

// ------------------------------------------------------------------------------------------
// Residual add (model.py:311-312)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_residual_kernel(const float* __restrict__ x, const float* __restrict__ raw,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           float* __restrict__ out, long long total4, int C, int HW4) {
    sifnn::pdl_wait_and_trigger();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((idx / HW4) % C);
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + idx);
        const float4 r = __ldg(reinterpret_cast<const float4*>(raw) + idx);
        float4 o;
        o.x = a.x + sifnn::act_affine_relu(r.x, sc, sh);
        o.y = a.y + sifnn::act_affine_relu(r.y, sc, sh);
        o.z = a.z + sifnn::act_affine_relu(r.z, sc, sh);
        o.w = a.w + sifnn::act_affine_relu(r.w, sc, sh);
        reinterpret_cast<float4*>(out)[idx] = o;
    }
}

// ------------------------------------------------------------------------------------------
// Bilinear x2 (align_corners=True) + concat (model.py:207,236,247) and adjoint
// ------------------------------------------------------------------------------------------
struct UpCoord { int i0, i1; float w0, w1; };
__device__ __forceinline__ UpCoord up_coord(int dst, int n_in, float rscale) {
    // ATen area_pixel_compute_source_index(align_corners=true): src = scale * dst, scale = (in-1)/(out-1) in fp32
    const float s = rscale * (float)dst;
    UpCoord u;
    u.i0 = (int)s;
    u.i1 = u.i0 + ((u.i0 < n_in - 1) ? 1 : 0);
    u.w1 = s - (float)u.i0;
    u.w0 = 1.0f - u.w1;
    return u;
}

// grid = (ceil(Ho*Wo/4 / 256), C1+C2, B); each thread writes 4 consecutive output pixels (one float4)
__global__ void __launch_bounds__(256) act_upcat_kernel(const float* __restrict__ low, const float* __restrict__ lsc, const float* __restrict__ lsh,
                                                        const float* __restrict__ skip, const float* __restrict__ ssc, const float* __restrict__ ssh,
                                                        float* __restrict__ out, int C1, int C2, int H, int W, float ry, float rx) {
    const int Ho = 2 * H, Wo = 2 * W, Wq = Wo >> 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Ho * Wq) return;
    const int y = t / Wq, x = (t - y * Wq) * 4;
    const int c = blockIdx.y, b = blockIdx.z;
    float4 v;
    if (c < C1) {
        const float sc = __ldg(lsc + c), sh = __ldg(lsh + c);
        const UpCoord uy = up_coord(y, H, ry);
        const float* p0 = low + ((size_t)b * C1 + c) * H * W + (size_t)uy.i0 * W;
        const float* p1 = low + ((size_t)b * C1 + c) * H * W + (size_t)uy.i1 * W;
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const UpCoord ux = up_coord(x + j, W, rx);
            const float v00 = sifnn::act_affine_relu(__ldg(p0 + ux.i0), sc, sh);
            const float v01 = sifnn::act_affine_relu(__ldg(p0 + ux.i1), sc, sh);
            const float v10 = sifnn::act_affine_relu(__ldg(p1 + ux.i0), sc, sh);
            const float v11 = sifnn::act_affine_relu(__ldg(p1 + ux.i1), sc, sh);
            r[j] = uy.w0 * (ux.w0 * v00 + ux.w1 * v01) + uy.w1 * (ux.w0 * v10 + ux.w1 * v11);
        }
        v = make_float4(r[0], r[1], r[2], r[3]);
    } else {
        const int cs = c - C1;
        const float sc = __ldg(ssc + cs), sh = __ldg(ssh + cs);
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(skip + ((size_t)b * C2 + cs) * Ho * Wo + (size_t)y * Wo + x));
        v = make_float4(sifnn::act_affine_relu(s4.x, sc, sh), sifnn::act_affine_relu(s4.y, sc, sh),
                        sifnn::act_affine_relu(s4.z, sc, sh), sifnn::act_affine_relu(s4.w, sc, sh));
    }
    *reinterpret_cast<float4*>(out + ((size_t)b * (C1 + C2) + c) * Ho * Wo + (size_t)y * Wo + x) = v;
}

// Same operation for C1 == C2, C1 % 4 == 0 (every decoder level).  act_upcat_kernel is bound by instruction issue on its up-sampled
// half (ncu: 82 % issue active, ~375 instructions per thread, most of them source coordinates and 64-bit addresses) while its skip
// half is a pure copy.  Here a thread owns four output pixels of UPC = 4 up-sampled channels AND of 4 skip channels: coordinates and
// source offsets are computed once, the skip loads are issued first and the gathers of channel u + 1 before channel u is
// interpolated, so every warp carries both kinds of work and the copy's latency hides behind the interpolation.
// grid = (ceil(Ho*Wo/4 / 256), C1/4, B).
constexpr int UPC = 4;
__global__ void __launch_bounds__(256, 2) act_upcat4_kernel(const float* __restrict__ low, const float* __restrict__ lsc, const float* __restrict__ lsh,
                                                            const float* __restrict__ skip, const float* __restrict__ ssc, const float* __restrict__ ssh,
                                                            float* __restrict__ out, int C1, int H, int W, float ry, float rx) {
    sifnn::pdl_wait_and_trigger();
    const int Ho = 2 * H, Wo = 2 * W, Wq = Wo >> 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Ho * Wq) return;
    const int y = t / Wq, x = (t - y * Wq) * 4;
    const int b = blockIdx.z;
    const size_t oplane = (size_t)Ho * Wo;
    const int ooff = y * Wo + x;
    const int c0 = blockIdx.y * UPC;
    float4 s4[UPC];
    {
        const float* sp = skip + ((size_t)b * C1 + c0) * oplane + ooff;
#pragma unroll
        for (int u = 0; u < UPC; ++u) s4[u] = __ldg(reinterpret_cast<const float4*>(sp + u * oplane));
    }
    const UpCoord uy = up_coord(y, H, ry);
    int o0[4], o1[4];
    float wx0[4], wx1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const UpCoord ux = up_coord(x + j, W, rx);
        o0[j] = ux.i0; o1[j] = ux.i1;
        wx0[j] = ux.w0; wx1[j] = ux.w1;
    }
    const float* p0 = low + ((size_t)b * C1 + c0) * H * W + uy.i0 * W;
    const float* p1 = low + ((size_t)b * C1 + c0) * H * W + uy.i1 * W;
    float* q = out + ((size_t)b * 2 * C1 + c0) * oplane + ooff;
    float* qs = q + (size_t)C1 * oplane;
    float v[4][4], n[4][4];
    auto fetch = [&](float (&d)[4][4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            d[j][0] = __ldg(p0 + o0[j]); d[j][1] = __ldg(p0 + o1[j]);
            d[j][2] = __ldg(p1 + o0[j]); d[j][3] = __ldg(p1 + o1[j]);
        }
        p0 += (size_t)H * W; p1 += (size_t)H * W;
    };
    fetch(n);
#pragma unroll
    for (int u = 0; u < UPC; ++u) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) v[j][e] = n[j][e];
        if (u + 1 < UPC) fetch(n);
        const float sc = __ldg(lsc + c0 + u), sh = __ldg(lsh + c0 + u);
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v00 = sifnn::act_affine_relu(v[j][0], sc, sh), v01 = sifnn::act_affine_relu(v[j][1], sc, sh);
            const float v10 = sifnn::act_affine_relu(v[j][2], sc, sh), v11 = sifnn::act_affine_relu(v[j][3], sc, sh);
            r[j] = uy.w0 * (wx0[j] * v00 + wx1[j] * v01) + uy.w1 * (wx0[j] * v10 + wx1[j] * v11);
        }
        *reinterpret_cast<float4*>(q + u * oplane) = make_float4(r[0], r[1], r[2], r[3]);
        const float k_sc = __ldg(ssc + c0 + u), k_sh = __ldg(ssh + c0 + u);
        *reinterpret_cast<float4*>(qs + u * oplane) =
            make_float4(sifnn::act_affine_relu(s4[u].x, k_sc, k_sh), sifnn::act_affine_relu(s4[u].y, k_sc, k_sh),
                        sifnn::act_affine_relu(s4[u].z, k_sc, k_sh), sifnn::act_affine_relu(s4[u].w, k_sc, k_sh));
    }
}

// dlow[b][c][i][k] = sum_{y,x} wy(i,y) wx(k,x) dout[b][c][y][x]  (gather form of the adjoint)
__global__ void __launch_bounds__(256) upcat_bwd_low_kernel(const float* __restrict__ dout, float* __restrict__ dlow, long long total,
                                                            int C1, int C2, int H, int W, float ry, float rx) {
    const int Ho = 2 * H, Wo = 2 * W, Ct = C1 + C2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % W);
        const int i = (int)((idx / W) % H);
        const int c = (int)((idx / ((long long)W * H)) % C1);
        const long long b = idx / ((long long)W * H * C1);
        const float* g = dout + ((size_t)b * Ct + c) * Ho * Wo;
        float wyv[6], wxv[6];
        int ys[6], xs[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int y = 2 * i - 2 + t;
            ys[t] = y;
            float wgt = 0.f;
            if (y >= 0 && y < Ho) {
                const UpCoord u = up_coord(y, H, ry);
                if (u.i0 == i) wgt += u.w0;
                if (u.i1 == i) wgt += u.w1;
            }
            wyv[t] = wgt;
            const int x = 2 * k - 2 + t;
            xs[t] = x;
            wgt = 0.f;
            if (x >= 0 && x < Wo) {
                const UpCoord u = up_coord(x, W, rx);
                if (u.i0 == k) wgt += u.w0;
                if (u.i1 == k) wgt += u.w1;
            }
            wxv[t] = wgt;
        }
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            if (wyv[t] != 0.f) {
                float row = 0.f;
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    if (wxv[u] != 0.f) row = fmaf(wxv[u], __ldg(g + (size_t)ys[t] * Wo + xs[u]), row);
                acc = fmaf(wyv[t], row, acc);
            }
        }
        dlow[idx] = acc;
    }
}

// Same adjoint, tiled: one CTA = (plane, band of 8 low-resolution rows).  The 20 contributing rows of dout are staged in shared memory with
// coalesced 16-byte loads and stores; a thread then owns one low-resolution column k and LR of the band's rows: it reduces the
// 2 LR + 4 staged rows it needs along x into registers (6-tap gather, weights computed once per thread) and finishes along y from
// those registers -- one barrier, no intermediate buffer.  (The first version wrote the x pass back to shared memory, staged with
// scalar stores and divided by W at run time: ncu showed 78 % issue activity and 1.6 M store bank conflicts, 70 us at 16 x 128^2, B = 32.)
// W in {32, 64, 128} (the decoder's levels; LR = W / 32); other shapes take the generic kernel above.
constexpr int UPB_TL = 8, UPB_NR = 2 * UPB_TL + 4;
__device__ __forceinline__ void up_gather_weights(int i, int n_in, float r, float* wv) {
    // weight of output index 2i-2+t (t = 0..5) on low-resolution index i
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const int y = 2 * i - 2 + t;
        float wgt = 0.f;
        if (y >= 0 && y < 2 * n_in) {
            const UpCoord u = up_coord(y, n_in, r);
            if (u.i0 == i) wgt += u.w0;
            if (u.i1 == i) wgt += u.w1;
        }
        wv[t] = wgt;
    }
}
template <int LR>
__global__ void __launch_bounds__(256) upcat_bwd_low_tiled_kernel(const float* __restrict__ dout, float* __restrict__ dlow, float* __restrict__ dskip,
                                                                  int C1, int C2, int H, float ry, float rx) {
    sifnn::pdl_wait_and_trigger();
    constexpr int W = 32 * LR, Wo = 2 * W, Wq = Wo / 4, STRIDE = Wo + 8;   // staged row: column x at index x + 4, zeros at -2, -1, Wo, Wo + 1
    __shared__ __align__(16) float in_s[UPB_NR * STRIDE];
    __shared__ float wy_s[UPB_TL][6];      // the y weights depend on the row only: one thread per low-resolution row computes them once
    const int Ho = 2 * H;
    const int tid = threadIdx.x;
    const int plane = blockIdx.y;          // b * C1 + c
    const int b = plane / C1, c = plane - b * C1;
    const int i_lo = blockIdx.x * UPB_TL;
    const int y_lo = 2 * i_lo - 2;
    const float* g = dout + ((size_t)b * (C1 + C2) + c) * Ho * Wo;
    // C1 == C2 (dskip != nullptr): this CTA also copies the band's 16 output rows of skip channel c, LR 16-byte pieces per thread; the
    // loads are issued here and stored at the end, so the copy's latency hides behind the adjoint (it was a launch of its own)
    float4 sk[LR];
    if (dskip) {
        const float* gs = dout + ((size_t)b * (C1 + C2) + C1 + c) * Ho * Wo;
#pragma unroll
        for (int u = 0; u < LR; ++u) {
            const int idx = tid + u * 256, r = idx / Wq, x4 = idx % Wq;
            const int y = 2 * i_lo + r;
            sk[u] = y < Ho ? __ldg(reinterpret_cast<const float4*>(gs + (size_t)y * Wo) + x4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int idx = tid; idx < UPB_NR * Wq; idx += 256) {
        const int r = idx / Wq, x4 = idx % Wq;
        const int y = y_lo + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < Ho) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)y * Wo) + x4);
        *reinterpret_cast<float4*>(in_s + r * STRIDE + 4 + 4 * x4) = v;
    }
    if (tid < UPB_NR * 4) {
        const int r = tid >> 2, e = tid & 3;
        in_s[r * STRIDE + (e < 2 ? 2 + e : Wo + 2 + e)] = 0.f;
    }
    if (tid >= 128 && tid < 128 + UPB_TL) {
        float wv[6];
        up_gather_weights(i_lo + tid - 128, H, ry, wv);
#pragma unroll
        for (int t = 0; t < 6; ++t) wy_s[tid - 128][t] = wv[t];
    }
    __syncthreads();
    const int k = tid % W, il0 = (tid / W) * LR;      // this thread: column k, band rows il0 .. il0 + LR - 1
    float wx[6];
    up_gather_weights(k, W, rx, wx);
    float xr[2 * LR + 4];                              // staged rows 2 il0 .. 2 il0 + 2 LR + 3, reduced along x
#pragma unroll
    for (int r = 0; r < 2 * LR + 4; ++r) {
        const float2* p = reinterpret_cast<const float2*>(in_s + (2 * il0 + r) * STRIDE + 2 * k + 2);   // columns 2k-2 .. 2k+3
        const float2 a0 = p[0], a1 = p[1], a2 = p[2];
        xr[r] = fmaf(wx[0], a0.x, fmaf(wx[1], a0.y, fmaf(wx[2], a1.x, fmaf(wx[3], a1.y, fmaf(wx[4], a2.x, wx[5] * a2.y)))));
    }
#pragma unroll
    for (int l = 0; l < LR; ++l) {
        const int i = i_lo + il0 + l;
        if (i < H) {
            float acc = 0.f;
#pragma unroll
            for (int t = 0; t < 6; ++t) acc = fmaf(wy_s[il0 + l][t], xr[2 * l + t], acc);
            dlow[((size_t)plane * H + i) * W + k] = acc;
        }
    }
    if (dskip) {
        float* ds = dskip + ((size_t)b * C2 + c) * Ho * Wo;
#pragma unroll
        for (int u = 0; u < LR; ++u) {
            const int idx = tid + u * 256, r = idx / Wq, x4 = idx % Wq;
            const int y = 2 * i_lo + r;
            if (y < Ho) *(reinterpret_cast<float4*>(ds + (size_t)y * Wo) + x4) = sk[u];
        }
    }
}

__global__ void __launch_bounds__(256) upcat_bwd_skip_kernel(const float* __restrict__ dout, float* __restrict__ dskip, long long total4,
                                                             int C1, int C2, int HWo4) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
        const long long per_img = (long long)C2 * HWo4;
        const long long b = idx / per_img;
        const long long r = idx - b * per_img;
        reinterpret_cast<float4*>(dskip)[idx] = __ldg(reinterpret_cast<const float4*>(dout) + (b * (C1 + C2) + C1) * (long long)HWo4 + r);
    }
}

// ------------------------------------------------------------------------------------------
// Bicubic x4 (Keys a=-0.75, half-pixel centres, clamped borders == cv2.INTER_CUBIC,
// utils.py:163-180) fused with the channel concat (train_model_B_gradFTM.py:94)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
    const float A = -0.75f;
    const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
    c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
    c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
    c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// grid = (ceil(H*W/4 / 256), B): each thread produces 4 consecutive output pixels (one float4 of each channel).  For a x4
// up-sampling the four outputs of a source column j use source columns j-2..j+2 and the phases t = .625, .875, .125, .375.
__global__ void __launch_bounds__(256) bicubic4_cat_kernel(const float* __restrict__ lst, const float* __restrict__ ndvi, float* __restrict__ xout,
                                                           int h, int w) {
    const int H = 4 * h, W = 4 * w;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= H * w) return;
    const int y = t / w, j = t - y * w;
    const int b = blockIdx.y;
    const float sy = ((float)y + 0.5f) * 0.25f - 0.5f;
    const float fy = floorf(sy);
    float cy[4];
    cubic_coeffs(sy - fy, cy);
    const int iy = (int)fy;
    const float* p = lst + (size_t)b * h * w;
    // row-interpolated values at the 5 source columns j-2 .. j+2 (clamped)
    float col[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        const int xx = min(max(j - 2 + d, 0), w - 1);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = min(max(iy - 1 + a, 0), h - 1);
            acc = fmaf(cy[a], __ldg(p + (size_t)yy * w + xx), acc);
        }
        col[d] = acc;
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float sx = ((float)(4 * j + i) + 0.5f) * 0.25f - 0.5f;
        const float fx = floorf(sx);
        float cx[4];
        cubic_coeffs(sx - fx, cx);
        const int base = (int)fx - 1 - (j - 2);  // 0 for i = 0,1 and 1 for i = 2,3
        o[i] = cx[0] * col[base] + cx[1] * col[base + 1] + cx[2] * col[base + 2] + cx[3] * col[base + 3];
    }
    const size_t off = (size_t)b * 2 * H * W + (size_t)y * W + 4 * j;
    *reinterpret_cast<float4*>(xout + off) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(xout + off + (size_t)H * W) = __ldg(reinterpret_cast<const float4*>(ndvi + (size_t)b * H * W + (size_t)y * W + 4 * j));
}

}  // namespace

// ==========================================================================================
extern "C" int sifnn_bn_train_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                                       float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                                       int C, double n, sifnn_stream_t stream) {
    SIFNN_REQUIRE(stats && gamma && beta && scale && shift && save_mean && save_invstd && C > 0 && n > 0, "bn_train_finalize: bad arguments");
    SIFNN_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_train_finalize: running buffers must both be set or both NULL");
    bn_train_finalize_kernel<<<(C + 63) / 64, 64, 0, sifnn::as_stream(stream)>>>(stats, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, C, n);
    return sifnn::check_launch("bn_train_finalize_kernel");
}

extern "C" int sifnn_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                    float* scale, float* shift, int C, sifnn_stream_t stream) {
    SIFNN_REQUIRE(gamma && beta && running_mean && running_var && scale && shift && C > 0, "bn_eval_affine: bad arguments");
    bn_eval_affine_kernel<<<(C + 63) / 64, 64, 0, sifnn::as_stream(stream)>>>(gamma, beta, running_mean, running_var, scale, shift, C);
    return sifnn::check_launch("bn_eval_affine_kernel");
}

// SIFNN_BN_REVERSE: 0 both passes front to back; 1 (default) apply back to front (starts with what reduce left in L2); 2 reduce back to front
// (starts with what the producer of dY left in L2), apply front to back
static int bn_bwd_order() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_BN_REVERSE"); v = e ? atoi(e) : 1; if (v < 0 || v > 2) v = 1; }
    return v;
}

static int bn_bwd_chunks(int B, int C, int HW) {
    const int n4 = HW / 4;
    long long want = (long long)sifnn::num_sms() * 8 / ((long long)B * C) + 1;
    long long maxc = (n4 + 255) / 256;
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    return (int)want;
}

extern "C" int sifnn_bn_relu_bwd_reduce(const float* dY, const float* raw, const float* scale, const float* shift,
                                        const float* save_mean, const float* save_invstd, double* sums, int B, int C, int HW,
                                        sifnn_stream_t stream) {
    SIFNN_REQUIRE(dY && raw && scale && shift && save_mean && save_invstd && sums, "bn_relu_bwd_reduce: null pointer");
    SIFNN_REQUIRE(B > 0 && C > 0 && HW > 0 && HW % 4 == 0 && B <= 65535 && C <= 65535, "bn_relu_bwd_reduce: bad shape (HW must be a multiple of 4)");
    dim3 grid(bn_bwd_chunks(B, C, HW), C, B);
    SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, bn_relu_bwd_reduce_kernel, grid, dim3(256), (size_t)0, sifnn::as_stream(stream), dY, raw, scale, shift, save_mean, save_invstd, sums, C, HW,
                                 bn_bwd_order() == 2 ? 1 : 0));
    return sifnn::check_launch("bn_relu_bwd_reduce_kernel");
}

extern "C" int sifnn_bn_relu_bwd_apply(const float* dY, const float* raw, const float* scale, const float* shift,
                                       const float* save_mean, const float* save_invstd, const float* gamma, const double* sums,
                                       float* dx, float* dgamma, float* dbeta, int B, int C, int HW, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dY && raw && scale && shift && save_mean && save_invstd && gamma && sums && dx, "bn_relu_bwd_apply: null pointer");
    SIFNN_REQUIRE(B > 0 && C > 0 && HW > 0 && HW % 4 == 0 && B <= 65535 && C <= 65535, "bn_relu_bwd_apply: bad shape (HW must be a multiple of 4)");
    dim3 grid(bn_bwd_chunks(B, C, HW), C, B);
    const int reverse = bn_bwd_order() == 1 ? 1 : 0;
    SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, bn_relu_bwd_apply_kernel, grid, dim3(256), (size_t)0, sifnn::as_stream(stream), dY, raw, scale, shift, save_mean, save_invstd, gamma, sums, dx, dgamma, dbeta,
                                 C, HW, 1.0 / ((double)B * HW), reverse));
    return sifnn::check_launch("bn_relu_bwd_apply_kernel");
}

extern "C" int sifnn_act_avgpool2_fwd(const float* raw, const float* scale, const float* shift, float* out, int B, int C, int H, int W,
                                      sifnn_stream_t stream) {
    SIFNN_REQUIRE(raw && scale && shift && out, "act_avgpool2_fwd: null pointer");
    SIFNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 4 == 0, "act_avgpool2_fwd: need even H and W %% 4 == 0");
    const long long total = (long long)B * C * (H / 2) * (W / 4);
    SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, act_avgpool2_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, sifnn::as_stream(stream), raw, scale, shift, out, total, C, H, W));
    return sifnn::check_launch("act_avgpool2_kernel");
}

extern "C" int sifnn_avgpool2_bwd(const float* dout, float* din, int accumulate, int B, int C, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dout && din, "avgpool2_bwd: null pointer");
    SIFNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 4 == 0, "avgpool2_bwd: need even H and W %% 4 == 0");
    const long long total = (long long)B * C * (H / 2) * (W / 4);
    SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, avgpool2_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, sifnn::as_stream(stream), dout, din, total, H, W, accumulate ? 1 : 0));
    return sifnn::check_launch("avgpool2_bwd_kernel");
}

extern "C" int sifnn_act_residual_fwd(const float* x, const float* raw, const float* scale, const float* shift, float* out, int B, int C,
                                      int HW, sifnn_stream_t stream) {
    SIFNN_REQUIRE(x && raw && scale && shift && out, "act_residual_fwd: null pointer");
    SIFNN_REQUIRE(B > 0 && C > 0 && HW > 0 && HW % 4 == 0, "act_residual_fwd: HW must be a multiple of 4");
    const long long total4 = (long long)B * C * (HW / 4);
    SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, act_residual_kernel, dim3(grid_for(total4, 256)), dim3(256), (size_t)0, sifnn::as_stream(stream), x, raw, scale, shift, out, total4, C, HW / 4));
    return sifnn::check_launch("act_residual_kernel");
}

static inline float up_ratio(int n_in) { return n_in > 1 ? (float)(n_in - 1) / (float)(2 * n_in - 1) : 0.f; }

extern "C" int sifnn_act_upcat_fwd(const float* low, const float* low_scale, const float* low_shift, const float* skip,
                                   const float* skip_scale, const float* skip_shift, float* out, int B, int C1, int C2, int H, int W,
                                   sifnn_stream_t stream) {
    SIFNN_REQUIRE(low && low_scale && low_shift && skip && skip_scale && skip_shift && out, "act_upcat_fwd: null pointer");
    SIFNN_REQUIRE(B > 0 && C1 > 0 && C2 > 0 && H > 0 && W > 0, "act_upcat_fwd: bad shape");
    SIFNN_REQUIRE(W % 2 == 0 && B <= 65535 && C1 + C2 <= 65535, "act_upcat_fwd: need even W, B and channels <= 65535");
    if (C1 == C2 && C1 % UPC == 0) {
        dim3 grid4((2 * H * (2 * W / 4) + 255) / 256, C1 / UPC, B);
        SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, act_upcat4_kernel, grid4, dim3(256), (size_t)0, sifnn::as_stream(stream), low, low_scale, low_shift, skip, skip_scale, skip_shift, out, C1, H, W,
                                        up_ratio(H), up_ratio(W)));
        return sifnn::check_launch("act_upcat4_kernel");
    }
    dim3 grid((2 * H * (2 * W / 4) + 255) / 256, C1 + C2, B);
    act_upcat_kernel<<<grid, 256, 0, sifnn::as_stream(stream)>>>(low, low_scale, low_shift, skip, skip_scale, skip_shift, out, C1, C2, H, W,
                                                                   up_ratio(H), up_ratio(W));
    return sifnn::check_launch("act_upcat_kernel");
}

extern "C" int sifnn_upcat_bwd(const float* dout, float* dlow, float* dskip, int B, int C1, int C2, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dout && dlow && dskip, "upcat_bwd: null pointer");
    SIFNN_REQUIRE(B > 0 && C1 > 0 && C2 > 0 && H > 0 && W > 0 && (H * W) % 1 == 0, "upcat_bwd: bad shape");
    cudaStream_t st = sifnn::as_stream(stream);
    const long long total = (long long)B * C1 * H * W;
    if ((W == 32 || W == 64 || W == 128) && (long long)B * C1 <= 65535) {
        dim3 grid((H + UPB_TL - 1) / UPB_TL, B * C1);
        float* fused_skip = (C1 == C2) ? dskip : nullptr;
        if (W == 128) SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, upcat_bwd_low_tiled_kernel<4>, grid, dim3(256), (size_t)0, st, dout, dlow, fused_skip, C1, C2, H, up_ratio(H), up_ratio(W)));
        else if (W == 64) SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, upcat_bwd_low_tiled_kernel<2>, grid, dim3(256), (size_t)0, st, dout, dlow, fused_skip, C1, C2, H, up_ratio(H), up_ratio(W)));
        else SIFNN_CUDA(sifnn::launch_pdl_if(sifnn::pdl_mode() != 0, upcat_bwd_low_tiled_kernel<1>, grid, dim3(256), (size_t)0, st, dout, dlow, fused_skip, C1, C2, H, up_ratio(H), up_ratio(W)));
        if (fused_skip) return sifnn::check_launch("upcat_bwd_low_tiled_kernel");
        SIFNN_TRY(sifnn::check_launch("upcat_bwd_low_tiled_kernel"));
    } else {
        upcat_bwd_low_kernel<<<grid_for(total, 256), 256, 0, st>>>(dout, dlow, total, C1, C2, H, W, up_ratio(H), up_ratio(W));
        SIFNN_TRY(sifnn::check_launch("upcat_bwd_low_kernel"));
    }
    const int HWo4 = H * W;  // (2H*2W)/4
    const long long total4 = (long long)B * C2 * HWo4;
    upcat_bwd_skip_kernel<<<grid_for(total4, 256), 256, 0, st>>>(dout, dskip, total4, C1, C2, HWo4);
    return sifnn::check_launch("upcat_bwd_skip_kernel");
}

extern "C" int sifnn_bicubic4_cat(const float* lst, const float* ndvi, float* x, int B, int h, int w, sifnn_stream_t stream) {
    SIFNN_REQUIRE(lst && ndvi && x && B > 0 && h > 0 && w > 0, "bicubic4_cat: bad arguments");
    SIFNN_REQUIRE(B <= 65535, "bicubic4_cat: batch too large");
    dim3 grid((4 * h * w + 255) / 256, B);
    bicubic4_cat_kernel<<<grid, 256, 0, sifnn::as_stream(stream)>>>(lst, ndvi, x, h, w);
    return sifnn::check_launch("bicubic4_cat_kernel");
}

// ==========================================================================================
// Whole-tile driver kernels (reference predict.py:84-103): gather 64x64 LST windows / 256x256 NDVI windows out of
// the MODIS tile with the per-window arithmetic fused (NDVI clip to [-1,1], z-score), and scatter the de-normalised
// super-resolved patches back into the 4x tile.  Patch p of the list covers window (wy[p], wx[p]) (window units).
// ==========================================================================================
namespace {

__global__ void __launch_bounds__(256) tile_gather_kernel(const float* __restrict__ lst_tile, const float* __restrict__ ndvi_tile,
                                                          const int* __restrict__ wy, const int* __restrict__ wx, float* __restrict__ lst_out,
                                                          float* __restrict__ ndvi_out, int P, int Wt, float mean_lst, float inv_std_lst,
                                                          float mean_ndvi, float inv_std_ndvi) {
    // one CTA per (patch, 16-row band of the 256x256 NDVI window); the 64x64 LST window rides along in band 0..3
    const int p = blockIdx.y, band = blockIdx.x;
    const int y0 = wy[p] * 64, x0 = wx[p] * 64;
    const float* nsrc = ndvi_tile + (size_t)(4 * y0 + band * 16) * (4 * Wt) + 4 * x0;
    float* ndst = ndvi_out + (size_t)p * 65536 + band * 16 * 256;
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
        const int r = i >> 6, c4 = i & 63;
        float4 v = __ldg(reinterpret_cast<const float4*>(nsrc + (size_t)r * (4 * Wt)) + c4);
        v.x = (fminf(fmaxf(v.x, -1.f), 1.f) - mean_ndvi) * inv_std_ndvi;
        v.y = (fminf(fmaxf(v.y, -1.f), 1.f) - mean_ndvi) * inv_std_ndvi;
        v.z = (fminf(fmaxf(v.z, -1.f), 1.f) - mean_ndvi) * inv_std_ndvi;
        v.w = (fminf(fmaxf(v.w, -1.f), 1.f) - mean_ndvi) * inv_std_ndvi;
        reinterpret_cast<float4*>(ndst + r * 256)[c4] = v;
    }
    if (band < 4) {
        const float* lsrc = lst_tile + (size_t)(y0 + band * 16) * Wt + x0;
        float* ldst = lst_out + (size_t)p * 4096 + band * 16 * 64;
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            ldst[r * 64 + c] = (__ldg(lsrc + (size_t)r * Wt + c) - mean_lst) * inv_std_lst;
        }
    }
}

__global__ void __launch_bounds__(256) tile_scatter_kernel(const float* __restrict__ sr, const int* __restrict__ wy, const int* __restrict__ wx,
                                                           float* __restrict__ out_tile, int P, int Wt, float mean_lst, float std_lst) {
    const int p = blockIdx.y, band = blockIdx.x;
    const int y0 = wy[p] * 256, x0 = wx[p] * 256;
    const float* src = sr + (size_t)p * 65536 + band * 16 * 256;
    float* dst = out_tile + (size_t)(y0 + band * 16) * (4 * Wt) + x0;
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
        const int r = i >> 6, c4 = i & 63;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + r * 256) + c4);
        v.x = fmaf(v.x, std_lst, mean_lst); v.y = fmaf(v.y, std_lst, mean_lst);
        v.z = fmaf(v.z, std_lst, mean_lst); v.w = fmaf(v.w, std_lst, mean_lst);
        reinterpret_cast<float4*>(dst + (size_t)r * (4 * Wt))[c4] = v;
    }
}

}  // namespace

extern "C" int sifnn_tile_gather(const float* lst_tile, const float* ndvi_tile, const int* wy, const int* wx, float* lst_out, float* ndvi_out,
                                 int P, int Ht, int Wt, float mean_lst, float std_lst, float mean_ndvi, float std_ndvi, sifnn_stream_t stream) {
    SIFNN_REQUIRE(lst_tile && ndvi_tile && wy && wx && lst_out && ndvi_out && P > 0 && P <= 65535, "tile_gather: bad arguments");
    SIFNN_REQUIRE(Ht >= 64 && Wt >= 64 && Wt % 4 == 0 && std_lst != 0.f && std_ndvi != 0.f, "tile_gather: bad tile shape / statistics");
    tile_gather_kernel<<<dim3(16, P), 256, 0, sifnn::as_stream(stream)>>>(lst_tile, ndvi_tile, wy, wx, lst_out, ndvi_out, P, Wt, mean_lst, 1.f / std_lst,
                                                                         mean_ndvi, 1.f / std_ndvi);
    return sifnn::check_launch("tile_gather_kernel");
}

extern "C" int sifnn_tile_scatter(const float* sr, const int* wy, const int* wx, float* out_tile, int P, int Ht, int Wt, float mean_lst,
                                  float std_lst, sifnn_stream_t stream) {
    SIFNN_REQUIRE(sr && wy && wx && out_tile && P > 0 && P <= 65535 && Ht >= 64 && Wt >= 64, "tile_scatter: bad arguments");
    tile_scatter_kernel<<<dim3(16, P), 256, 0, sifnn::as_stream(stream)>>>(sr, wy, wx, out_tile, P, Wt, mean_lst, std_lst);
    return sifnn::check_launch("tile_scatter_kernel");
}
