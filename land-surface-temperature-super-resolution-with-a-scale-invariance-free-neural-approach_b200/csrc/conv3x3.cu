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

// ---- staging (cp.async) -------------------------------------------------------------------
// Every thread owns a FIXED set of (tile row, column group) slots for the whole kernel, so the
// index arithmetic (row clamp / zero-fill decision, offsets) is done once per CTA; per chunk a
// thread just walks its slots over the chunk's channel planes (pointer += plane).  Vector
// slots copy 16 B of the 32 interior columns, scalar slots copy one halo element.  Zero padding
// (PAD_ZERO) is the zero-fill form of cp.async (src-size 0).  The generic per-element path is
// kept for tiles that are not fully inside the image in x, or unaligned tensors.
template <int IN_ROWS>
struct FillSlots {
    static constexpr int NV = IN_ROWS * 8;              // vector slots per channel plane
    static constexpr int NS = IN_ROWS * 2;              // scalar (halo) slots per channel plane
    static constexpr int VSL = (NV + 255) / 256;        // vector slots per thread
    static constexpr int SOFF = (((NV % 256) + 31) / 32 * 32 + NS <= 256) ? ((NV % 256) + 31) / 32 * 32 : 0;
    int vsrc[VSL], vdst[VSL], vbytes[VSL];              // vdst < 0: slot unused
    int ssrc, sdst, sbytes;
};

template <int IN_ROWS, int PAD>
__device__ __forceinline__ void conv_make_slots(FillSlots<IN_ROWS>& fs, int H, int W, int x0, int y0, int tid) {
    using FS = FillSlots<IN_ROWS>;
#pragma unroll
    for (int s = 0; s < FS::VSL; ++s) {
        const int pv = tid + 256 * s;
        fs.vdst[s] = -1; fs.vsrc[s] = 0; fs.vbytes[s] = 0;
        if (pv < FS::NV) {
            const int r = pv >> 3, j = pv & 7;
            int gy = y0 + r - 1;
            bool ok = true;
            if (PAD == PAD_REPLICATE) gy = min(max(gy, 0), H - 1);
            else ok = gy >= 0 && gy < H;
            fs.vdst[s] = r * IN_STRIDE + IN_X0 + 1 + 4 * j;
            fs.vsrc[s] = ok ? gy * W + x0 + 4 * j : 0;
            fs.vbytes[s] = ok ? 16 : 0;
        }
    }
    fs.sdst = -1; fs.ssrc = 0; fs.sbytes = 0;
    const int ps = tid - FS::SOFF;
    if (ps >= 0 && ps < FS::NS) {
        const int r = ps >> 1, right = ps & 1;
        int gy = y0 + r - 1, gx = right ? x0 + TW : x0 - 1;
        bool ok = true;
        if (PAD == PAD_REPLICATE) {
            gy = min(max(gy, 0), H - 1);
            gx = min(max(gx, 0), W - 1);
        } else {
            ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        }
        fs.sdst = r * IN_STRIDE + (right ? IN_X0 + 1 + TW : IN_X0);
        fs.ssrc = ok ? gy * W + gx : 0;
        fs.sbytes = ok ? 4 : 0;
    }
}

template <int IN_ROWS>
__device__ __forceinline__ void conv_fill_vec(float* in_st, const float* src_chunk, size_t plane, int nci, const FillSlots<IN_ROWS>& fs) {
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    using FS = FillSlots<IN_ROWS>;
#pragma unroll
    for (int s = 0; s < FS::VSL; ++s) {
        if (fs.vdst[s] >= 0) {
            const float* src = src_chunk + fs.vsrc[s];
            float* dst = in_st + fs.vdst[s];
            for (int ci = 0; ci < nci; ++ci) {
                sifnn::cp_async16(dst, src, fs.vbytes[s]);
                dst += IN_PLANE;
                src += plane;
            }
        }
    }
    if (fs.sdst >= 0) {
        const float* src = src_chunk + fs.ssrc;
        float* dst = in_st + fs.sdst;
        for (int ci = 0; ci < nci; ++ci) {
            sifnn::cp_async4(dst, src, fs.sbytes);
            dst += IN_PLANE;
            src += plane;
        }
    }
}

// BatchNorm + ReLU of the producing layer applied in place to the elements THIS thread copied
// (visible to it after cp.async.wait_group, before the CTA barrier).
template <int IN_ROWS>
__device__ __forceinline__ void conv_affine_vec(float* in_st, const float* sc, const float* sh, int nci, const FillSlots<IN_ROWS>& fs) {
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    using FS = FillSlots<IN_ROWS>;
#pragma unroll
    for (int s = 0; s < FS::VSL; ++s) {
        if (fs.vdst[s] >= 0) {
            float* dst = in_st + fs.vdst[s];
            for (int ci = 0; ci < nci; ++ci) {
                float4 v = *reinterpret_cast<float4*>(dst);
                const float c = sc[ci], h = sh[ci];
                v.x = sifnn::act_affine_relu(v.x, c, h); v.y = sifnn::act_affine_relu(v.y, c, h);
                v.z = sifnn::act_affine_relu(v.z, c, h); v.w = sifnn::act_affine_relu(v.w, c, h);
                *reinterpret_cast<float4*>(dst) = v;
                dst += IN_PLANE;
            }
        }
    }
    if (fs.sdst >= 0) {
        float* dst = in_st + fs.sdst;
        for (int ci = 0; ci < nci; ++ci) {
            *dst = sifnn::act_affine_relu(*dst, sc[ci], sh[ci]);
            dst += IN_PLANE;
        }
    }
}

// generic (slow) per-element staging: partial tiles in x, W % 4 != 0, unaligned tensors
template <int IN_ROWS, int PAD>
__device__ __forceinline__ void conv_fill_generic(float* in_st, const ConvArgs& a, const float* in_b, int c0, int nci, int x0, int y0, int tid) {
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int IN_COLS = TW + 2;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
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

template <int IN_ROWS>
__device__ __forceinline__ void conv_affine_generic(float* in_st, const float* sc, const float* sh, int nci, int tid) {
    constexpr int IN_PLANE = IN_ROWS * IN_STRIDE;
    constexpr int IN_COLS = TW + 2;
    for (int idx = tid; idx < nci * IN_ROWS * IN_COLS; idx += 256) {
        const int ci = idx / (IN_ROWS * IN_COLS);
        const int rem = idx - ci * (IN_ROWS * IN_COLS);
        const int r = rem / IN_COLS, c = rem - r * IN_COLS;
        float* q = in_st + ci * IN_PLANE + r * IN_STRIDE + IN_X0 + c;
        *q = sifnn::act_affine_relu(*q, sc[ci], sh[ci]);
    }
}

// weights: w_st[ci][tap][o]; each thread owns fixed (tap, o) slots and walks the chunk's channels
template <int CO_T>
struct WSlots {
    static constexpr int NW = 9 * CO_T;
    static constexpr int WSL = (NW + 255) / 256;
    long long src[WSL];  // offset of (o, c = 0, tap) in the weight tensor, or -1 if o >= O (zero-filled)
};

template <int CO_T>
__device__ __forceinline__ void conv_make_wslots(WSlots<CO_T>& ws, const ConvArgs& a, int o0, int tid) {
#pragma unroll
    for (int s = 0; s < WSlots<CO_T>::WSL; ++s) {
        const int wi = tid + 256 * s;
        ws.src[s] = -2;  // unused slot
        if (wi < WSlots<CO_T>::NW) {
            const int o = wi % CO_T, t = wi / CO_T;
            ws.src[s] = (o0 + o < a.O) ? (long long)(o0 + o) * a.w_so + (a.w_flip ? 8 - t : t) : -1;
        }
    }
}

template <int CO_T>
__device__ __forceinline__ void conv_fill_w(float* w_st, const ConvArgs& a, int c0, int nci, const WSlots<CO_T>& ws, int tid) {
#pragma unroll
    for (int s = 0; s < WSlots<CO_T>::WSL; ++s) {
        if (ws.src[s] != -2) {
            const bool ok = ws.src[s] >= 0;
            const float* src = a.w + (ok ? ws.src[s] + (long long)c0 * a.w_sk : 0);
            float* dst = w_st + tid + 256 * s;
            for (int ci = 0; ci < nci; ++ci) {
                sifnn::cp_async4(dst, src, ok ? 4 : 0);
                dst += 9 * CO_T;
                if (ok) src += a.w_sk;
            }
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

    // accumulators: packed pairs of output channels (FFMA2) when CPT is even, scalars otherwise
    constexpr bool PACKED = (CPT % 4 == 0);
    constexpr int NP = PACKED ? CPT / 2 : 1;
    sifnn::f32x2_t acc2[PY][NP];
    float acc[PY][PACKED ? 1 : CPT];
#pragma unroll
    for (int i = 0; i < PY; ++i) {
#pragma unroll
        for (int j = 0; j < NP; ++j) acc2[i][j] = 0ull;
#pragma unroll
        for (int j = 0; j < (PACKED ? 1 : CPT); ++j) acc[i][j] = 0.f;
    }

    FillSlots<IN_ROWS> fs;
    WSlots<CO_T> wsl;
    if (vec_ok) conv_make_slots<IN_ROWS, PAD>(fs, H, W, x0, y0, tid);
    conv_make_wslots<CO_T>(wsl, a, o0, tid);
    auto issue_fill = [&](float* in_st, int c0, int nci) {
        if (vec_ok) conv_fill_vec<IN_ROWS>(in_st, in_b + (size_t)c0 * plane, plane, nci, fs);
        else conv_fill_generic<IN_ROWS, PAD>(in_st, a, in_b, c0, nci, x0, y0, tid);
        conv_fill_w<CO_T>(in_st + CI_CHUNK * IN_PLANE, a, c0, nci, wsl, tid);
        sifnn::cp_async_commit();
    };

    const int nchunks = (K + CI_CHUNK - 1) / CI_CHUNK;
    issue_fill(smem, 0, min(CI_CHUNK, K));

    for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * CI_CHUNK;
        const int nci = min(CI_CHUNK, K - c0);
        float* in_s = smem + (ch & 1) * STAGE;
        float* w_s = in_s + CI_CHUNK * IN_PLANE;
        if (ch + 1 < nchunks) {  // prefetch the next chunk into the other stage while this one is consumed
            issue_fill(smem + ((ch + 1) & 1) * STAGE, c0 + CI_CHUNK, min(CI_CHUNK, K - c0 - CI_CHUNK));
            sifnn::cp_async_wait<1>();
        } else {
            sifnn::cp_async_wait<0>();
        }
        if (AFFINE) {
            if (ch == 0) __syncthreads();  // sc_s / sh_s
            if (vec_ok) conv_affine_vec<IN_ROWS>(in_s, sc_s + c0, sh_s + c0, nci, fs);
            else conv_affine_generic<IN_ROWS>(in_s, sc_s + c0, sh_s + c0, nci, tid);
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
                    const float* wq = wp + (ci * 9 + ky * 3 + kx) * CO_T;
                    if (PACKED) {
                        sifnn::f32x2_t w2[NP];
#pragma unroll
                        for (int j = 0; j < NP; j += 2) {
                            const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(wq + 2 * j);
                            w2[j] = t2.x; w2[j + 1] = t2.y;
                        }
                        // weight-stationary order: the 64-bit weight pair sits in the operand-reuse cache across the
                        // 8 rows, each FFMA2 then reads one scalar + one accumulator pair from the register file
                        // (98% of FMA peak in tools/ffma2_probe2.cu vs 87% for the x-stationary order)
#pragma unroll
                        for (int j = 0; j < NP; ++j) {
#pragma unroll
                            for (int py = 0; py < PY; ++py) {
                                const float x = v[py + ky][kx];
                                acc2[py][j] = sifnn::fma2(sifnn::pack2(x, x), w2[j], acc2[py][j]);
                            }
                        }
                    } else {
                        float wv[CPT];
#pragma unroll
                        for (int j = 0; j < CPT; ++j) wv[j] = wq[j];
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
                    float r;
                    if (PACKED) {
                        float lo, hi;
                        sifnn::unpack2(acc2[py][j >> 1], lo, hi);
                        r = ((j & 1) ? hi : lo) + bv;
                    } else {
                        r = acc[py][PACKED ? 0 : j] + bv;
                    }
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
constexpr int BORDER_KC = 32;  // input channels per CTA (4 warps x 8)

__global__ void __launch_bounds__(128) dgrad_border_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                           int Cin, int Cout, int H, int W, int mode) {
    // mode 0: rows pass incl. the corner cross terms; 1: columns pass; 2: columns pass incl. the corner cross terms (used when
    // the row terms were already folded into the main kernel as extra tap MMAs, see conv3x3_tc.cu)
    const int cols_pass = mode != 0;
    __shared__ __align__(16) float ws[BORDER_OC * 5 * BORDER_KC];  // [o][slot: 3 main taps + 2 corner taps][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int kc0 = blockIdx.z * BORDER_KC;
    const int kb = kc0 + warp * 8;
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
    // corner cross term, on dy at the pixel itself.  Rows pass: slot 3 = tap (ky_e, 0), slot 4 = tap (ky_e, 2);
    // columns pass (mode 2): slot 3 = tap (0, kx_e), slot 4 = tap (2, kx_e).
    int cslot = -1;
    if (mode == 0 && active) {
        if (q == 0) cslot = 3;
        else if (q == W - 1) cslot = 4;
    } else if (mode == 2 && active) {
        if (p == 0) cslot = 3;
        else if (p == H - 1) cslot = 4;
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
    const size_t offc = active ? (size_t)p * W + q : 0;

    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    for (int ob = 0; ob < Cout; ob += BORDER_OC) {
        const int no = min(BORDER_OC, Cout - ob);
        // all dy values of this chunk first: BORDER_OC x 4 independent loads in flight while the weights are staged
        float dv[BORDER_OC][4];
#pragma unroll
        for (int o = 0; o < BORDER_OC; ++o) {
            const float* dyo = dyb + (size_t)(ob + min(o, no - 1)) * plane;
#pragma unroll
            for (int j = 0; j < 3; ++j) dv[o][j] = (ok[j] && o < no) ? __ldg(dyo + off[j]) : 0.f;
            dv[o][3] = (cslot >= 0 && o < no) ? __ldg(dyo + offc) : 0.f;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < no * 5 * BORDER_KC; idx += 128) {
            const int k = idx % BORDER_KC;
            const int slot = (idx / BORDER_KC) % 5;
            const int o = idx / (5 * BORDER_KC);
            int t;
            if (slot < 3) t = t0 + slot * tstep;
            else if (!cols_pass) t = (slot == 3) ? (side ? 6 : 0) : (side ? 8 : 2);
            else t = (slot == 3) ? (side ? 2 : 0) : (side ? 8 : 6);
            ws[idx] = (kc0 + k < Cin) ? __ldg(w + ((size_t)(ob + o) * Cin + kc0 + k) * 9 + t) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < BORDER_OC; ++o) {
            if (o < no) {
                const float* wo = ws + o * 5 * BORDER_KC + warp * 8;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4 wa = *reinterpret_cast<const float4*>(wo + j * BORDER_KC);
                    const float4 wb = *reinterpret_cast<const float4*>(wo + j * BORDER_KC + 4);
                    const float d = dv[o][j];
                    acc[0] = fmaf(wa.x, d, acc[0]); acc[1] = fmaf(wa.y, d, acc[1]);
                    acc[2] = fmaf(wa.z, d, acc[2]); acc[3] = fmaf(wa.w, d, acc[3]);
                    acc[4] = fmaf(wb.x, d, acc[4]); acc[5] = fmaf(wb.y, d, acc[5]);
                    acc[6] = fmaf(wb.z, d, acc[6]); acc[7] = fmaf(wb.w, d, acc[7]);
                }
                if (cslot >= 0) {  // only the two corner lanes of a rows pass
                    const float* wc = wo + cslot * BORDER_KC;
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[u] = fmaf(wc[u], dv[o][3], acc[u]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (kb + u < Cin) dx[((size_t)b * Cin + kb + u) * plane + (size_t)p * W + q] += acc[u];
    }
}

// Columns pass of the padding adjoint (modes 1 and 2 of dgrad_border_kernel), restructured around what bounds it: every dy element of
// image column 0 / W-1 sits in its own 32-byte sector, so the pass is limited by the number of sector requests, not by bytes or FLOPs.
// One CTA = (side, 32-row segment, image, 32 input channels): the dy column segment (all Cout channels, +1 halo row each way) and the
// three weights of the side's kx for every (o, k) are staged in shared memory ONCE, lanes run along rows, warp w owns 4 input channels.
// Mode 2 folds the corner cross terms in the same way wgrad_tc.cu folds the padding along x: row 0 meets dy[0] a second time through
// the ky = 0 weight, row H-1 meets dy[H-1] a second time through the ky = 2 weight.
constexpr int BCOL_ROWS = 32;
__global__ void __launch_bounds__(256) dgrad_border_cols_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                                int Cin, int Cout, int H, int W, int corners) {
    extern __shared__ __align__(16) float bsm[];
    float* dys = bsm;                                   // [Cout][BCOL_ROWS + 2]: rows p0-1 .. p0+32
    float* wsm = bsm + ((Cout * (BCOL_ROWS + 2) + 3) & ~3);  // [Cout][3 ky][32 k], 16-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int segs = (H + BCOL_ROWS - 1) / BCOL_ROWS;
    const int side = blockIdx.x / segs;
    const int p0 = (blockIdx.x % segs) * BCOL_ROWS;
    const int b = blockIdx.y, kc0 = blockIdx.z * 32;
    const int q = side ? W - 1 : 0, kx = side ? 2 : 0;
    const size_t plane = (size_t)H * W;
    const float* dyb = dy + (size_t)b * Cout * plane + q;
    // the read half of the dx update is issued first: its DRAM latency hides behind the staging below
    const int p = p0 + lane;
    float dxv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int k = kc0 + 4 * warp + u;
        dxv[u] = (p < H && k < Cin) ? dx[((size_t)b * Cin + k) * plane + (size_t)p * W + q] : 0.f;
    }
    const int n_dy = Cout * (BCOL_ROWS + 2);
    for (int base = 0; base < n_dy; base += 4 * 256) {     // four independent sector requests in flight per thread
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * 256 + tid;
            const int o = idx / (BCOL_ROWS + 2), r = idx - o * (BCOL_ROWS + 2);
            const int y = p0 - 1 + r;
            v[u] = (idx < n_dy && y >= 0 && y < H) ? __ldg(dyb + (size_t)o * plane + (size_t)y * W) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * 256 + tid;
            if (idx < n_dy) dys[idx] = v[u];
        }
    }
    for (int idx = tid; idx < Cout * 96; idx += 256) {
        const int k = idx & 31, ky = (idx >> 5) % 3, o = idx / 96;
        wsm[idx] = (kc0 + k < Cin) ? __ldg(w + ((size_t)o * Cin + kc0 + k) * 9 + ky * 3 + kx) : 0.f;
    }
    __syncthreads();
    const bool top = corners && p == 0, bot = corners && p == H - 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int o = 0; o < Cout; ++o) {
        const float* dr = dys + o * (BCOL_ROWS + 2) + lane;   // dr[0] = dy[p-1], dr[1] = dy[p], dr[2] = dy[p+1]
        const float dc = dr[1];
        const float d0 = top ? dr[2] + dc : dr[2];            // ky = 0 reads row p+1
        const float d2 = bot ? dr[0] + dc : dr[0];            // ky = 2 reads row p-1
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + o * 96 + 4 * warp);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + o * 96 + 32 + 4 * warp);
        const float4 w2 = *reinterpret_cast<const float4*>(wsm + o * 96 + 64 + 4 * warp);
        acc[0] = fmaf(w0.x, d0, fmaf(w1.x, dc, fmaf(w2.x, d2, acc[0])));
        acc[1] = fmaf(w0.y, d0, fmaf(w1.y, dc, fmaf(w2.y, d2, acc[1])));
        acc[2] = fmaf(w0.z, d0, fmaf(w1.z, dc, fmaf(w2.z, d2, acc[2])));
        acc[3] = fmaf(w0.w, d0, fmaf(w1.w, dc, fmaf(w2.w, d2, acc[3])));
    }
    if (p < H) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = kc0 + 4 * warp + u;
            if (k < Cin) dx[((size_t)b * Cin + k) * plane + (size_t)p * W + q] = dxv[u] + acc[u];
        }
    }
}

// ---- one output channel (the network's `outlay`, 16 -> 1) -----------------------------------------------------------
// The register-tiled kernel above amortises every staged input value over CPT output channels; with one output channel
// there is nothing to amortise and it runs at a quarter of the HBM rate.  Here a thread owns a 4-column x TO1_ROWS-row
// block of the output and slides down the TO1_ROWS + 2 input rows of every channel: each row is loaded once (one 16-byte
// load + the two neighbours, straight from global memory -- the 3x reuse between row bands is L1 / L2 hits), gets the
// BatchNorm + ReLU prologue once, and feeds the three output rows it touches.  Replicate padding = clamped indices.
constexpr int TO1_MAXK = 64;
constexpr int TO1_THREADS = 128;

template <bool AFFINE, int TO1_ROWS>
__global__ void __launch_bounds__(TO1_THREADS, TO1_ROWS >= 4 ? 4 : 6) conv3x3_to1_kernel(const ConvArgs a) {
    __shared__ float w_s[TO1_MAXK * 9], sc_s[TO1_MAXK], sh_s[TO1_MAXK];
    const int K = a.K, H = a.H, W = a.W;
    for (int i = threadIdx.x; i < K * 9; i += TO1_THREADS) w_s[i] = __ldg(a.w + (size_t)(i / 9) * a.w_sk + (i % 9));
    if (AFFINE)
        for (int i = threadIdx.x; i < K; i += TO1_THREADS) { sc_s[i] = __ldg(a.in_scale + i); sh_s[i] = __ldg(a.in_shift + i); }
    __syncthreads();
    const int w4 = W >> 2;
    const int g = blockIdx.x * TO1_THREADS + threadIdx.x;
    const int band = g / w4;
    const int y0 = band * TO1_ROWS;
    if (y0 >= H) return;
    const int x0 = (g - band * w4) << 2;
    const int b = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* in_b = a.in + (size_t)b * K * plane;
    const int xl = max(x0 - 1, 0), xr = min(x0 + 4, W - 1);
    int rowoff[TO1_ROWS + 2];
#pragma unroll
    for (int r = 0; r < TO1_ROWS + 2; ++r) rowoff[r] = min(max(y0 - 1 + r, 0), H - 1) * W;
    float acc[TO1_ROWS][4];
#pragma unroll
    for (int r = 0; r < TO1_ROWS; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
#pragma unroll 1
    for (int ci = 0; ci < K; ++ci) {
        const float* ip = in_b + (size_t)ci * plane;
        float wv[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wv[t] = w_s[ci * 9 + t];
        const float sc = AFFINE ? sc_s[ci] : 1.f, sh = AFFINE ? sh_s[ci] : 0.f;
        float v[TO1_ROWS + 2][6];
#pragma unroll
        for (int r = 0; r < TO1_ROWS + 2; ++r) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(ip + rowoff[r] + x0));
            v[r][0] = __ldg(ip + rowoff[r] + xl);
            v[r][1] = m.x; v[r][2] = m.y; v[r][3] = m.z; v[r][4] = m.w;
            v[r][5] = __ldg(ip + rowoff[r] + xr);
        }
        if (AFFINE) {
#pragma unroll
            for (int r = 0; r < TO1_ROWS + 2; ++r)
#pragma unroll
                for (int i = 0; i < 6; ++i) v[r][i] = sifnn::act_affine_relu(v[r][i], sc, sh);
        }
#pragma unroll
        for (int r = 0; r < TO1_ROWS; ++r)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[r][i] = fmaf(v[r + ky][i + kx], wv[ky * 3 + kx], acc[r][i]);
    }
    const float bv = a.bias ? __ldg(a.bias) : 0.f;
    float* op = a.out + (size_t)b * plane;
#pragma unroll
    for (int r = 0; r < TO1_ROWS; ++r) {
        const int y = y0 + r;
        if (y < H)
            *reinterpret_cast<float4*>(op + (size_t)y * W + x0) = make_float4(acc[r][0] + bv, acc[r][1] + bv, acc[r][2] + bv, acc[r][3] + bv);
    }
}

static bool to1_eligible(const ConvArgs& a) {
    return a.O == 1 && a.K <= TO1_MAXK && a.W % 4 == 0 && !a.w_flip && !a.accumulate && !a.stats &&
           (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

template <bool AFFINE, int TO1_ROWS>
int launch_to1_rows(const ConvArgs& a, cudaStream_t st) {
    const int bands = (a.H + TO1_ROWS - 1) / TO1_ROWS;
    dim3 grid((bands * (a.W / 4) + TO1_THREADS - 1) / TO1_THREADS, a.B);
    conv3x3_to1_kernel<AFFINE, TO1_ROWS><<<grid, TO1_THREADS, 0, st>>>(a);
    return sifnn::check_launch("conv3x3_to1_kernel");
}

template <bool AFFINE>
int launch_to1(const ConvArgs& a, cudaStream_t st) {
    return launch_to1_rows<AFFINE, 4>(a, st);   // 2 / 3 rows per thread: 55 / 49 us against 49 us at B = 32, 256 x 256
}

template <int CPT, int WARPS_CO, int PAD, bool AFFINE>
int launch_conv(const ConvArgs& a0, cudaStream_t st) {
    constexpr int WARPS_ROW = 8 / WARPS_CO;
    constexpr int ROWS = PY * WARPS_ROW;
    constexpr int CO_T = CPT * WARPS_CO;
    constexpr int IN_PLANE = (ROWS + 2) * IN_STRIDE;
    constexpr size_t smem = 2 * (size_t)(CI_CHUNK * IN_PLANE + CI_CHUNK * 9 * CO_T) * sizeof(float);  // two pipeline stages
    static sifnn::PerDeviceOnce attr_once;   // the attribute is per device: one flag per device, not one per process
    auto kern = conv3x3_kernel<CPT, WARPS_CO, PAD, AFFINE>;
    if (attr_once.first_time()) {
        SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    if (PAD == PAD_REPLICATE && to1_eligible(a)) return launch_to1<AFFINE>(a, st);
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

// ---- data gradient of a one-output-channel layer (dy has ONE channel, dx has Cin) ----------------------------------------
// dx[k][p][q] = sum_{ky,kx} w[k][ky][kx] * S(p, ky, q, kx), where S gathers every dy[y][x] whose replicate-padded read position
// (clamp(y + ky - 1), clamp(x + kx - 1)) is (p, q): the regular term dy[p - ky + 1][q - kx + 1] plus, on the image border, the
// terms the padding folded onto the edge row / column.  S does not depend on k, so a thread builds the 36 values of its four
// pixels once per output row and then streams the Cin output channels: 36 FMAs + one 16-byte store per channel.  One pass,
// no separate border kernel; bound by the dx write.
constexpr int FROM1_ROWS = 4;

__global__ void __launch_bounds__(128, 4) dgrad_from1_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                             float* __restrict__ dx, int Cin, int H, int W) {
    __shared__ float w_s[TO1_MAXK * 9];
    for (int i = threadIdx.x; i < Cin * 9; i += 128) w_s[i] = __ldg(w + i);
    __syncthreads();
    const int w4 = W >> 2;
    const int g = blockIdx.x * 128 + threadIdx.x;
    const int band = g / w4;
    const int y0 = band * FROM1_ROWS;
    if (y0 >= H) return;
    const int x0 = (g - band * w4) << 2;
    const int b = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* dyb = dy + (size_t)b * plane;
    float D[FROM1_ROWS + 2][6];     // rows y0-1 .. y0+R, columns x0-1 .. x0+4; zero outside the image
#pragma unroll
    for (int r = 0; r < FROM1_ROWS + 2; ++r) {
        const int y = y0 - 1 + r;
        if (y >= 0 && y < H) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(dyb + (size_t)y * W + x0));
            D[r][0] = x0 > 0 ? __ldg(dyb + (size_t)y * W + x0 - 1) : 0.f;
            D[r][1] = m.x; D[r][2] = m.y; D[r][3] = m.z; D[r][4] = m.w;
            D[r][5] = x0 + 4 < W ? __ldg(dyb + (size_t)y * W + x0 + 4) : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < 6; ++j) D[r][j] = 0.f;
        }
    }
    float* dxb = dx + (size_t)b * Cin * plane;
#pragma unroll
    for (int i = 0; i < FROM1_ROWS; ++i) {
        const int p = y0 + i;
        if (p >= H) break;
        float T[3][3][4];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float rs[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float v = D[i - ky + 2][j];
                if (ky == 0 && p == 0) v += D[i + 1][j];          // row -1 of the padded input is row 0: dy[0] with ky = 0
                if (ky == 2 && p == H - 1) v += D[i + 1][j];      // row H of the padded input is row H-1: dy[H-1] with ky = 2
                rs[j] = v;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v = rs[c - kx + 2];
                    if (kx == 0 && x0 + c == 0) v += rs[c + 1];
                    if (kx == 2 && x0 + c == W - 1) v += rs[c + 1];
                    T[ky][kx][c] = v;
                }
        }
#pragma unroll 2
        for (int k = 0; k < Cin; ++k) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float wv = w_s[k * 9 + t];
                a0 = fmaf(T[t / 3][t % 3][0], wv, a0);
                a1 = fmaf(T[t / 3][t % 3][1], wv, a1);
                a2 = fmaf(T[t / 3][t % 3][2], wv, a2);
                a3 = fmaf(T[t / 3][t % 3][3], wv, a3);
            }
            *reinterpret_cast<float4*>(dxb + (size_t)k * plane + (size_t)p * W + x0) = make_float4(a0, a1, a2, a3);
        }
    }
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
    if (Cout == 1 && Cin <= TO1_MAXK && W % 4 == 0 && !accumulate && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
        const int bands = (H + FROM1_ROWS - 1) / FROM1_ROWS;
        dim3 grid((bands * (W / 4) + 127) / 128, B);
        dgrad_from1_kernel<<<grid, 128, 0, st>>>(dy, w, dx, Cin, H, W);
        return sifnn::check_launch("dgrad_from1_kernel");
    }
    SIFNN_TRY((dispatch_conv<PAD_ZERO, false>(a, st)));
    return sifnn_conv3x3_dgrad_border(dy, w, dx, B, Cin, Cout, H, W, stream);
}

static int launch_border(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H, int W, int mode, cudaStream_t st) {
    if (mode != 0) {
        const size_t smem = ((size_t)((Cout * (BCOL_ROWS + 2) + 3) & ~3) + (size_t)Cout * 96) * sizeof(float);
        static size_t smem_set = 48 * 1024;
        if (smem > smem_set) {
            SIFNN_CUDA(cudaFuncSetAttribute(dgrad_border_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set = smem;
        }
        dim3 grid(2 * ((H + BCOL_ROWS - 1) / BCOL_ROWS), B, (Cin + 31) / 32);
        dgrad_border_cols_kernel<<<grid, 256, smem, st>>>(dy, w, dx, Cin, Cout, H, W, mode == 2 ? 1 : 0);
        return sifnn::check_launch("dgrad_border_cols_kernel");
    }
    const int L = mode ? H : W;
    dim3 grid(2 * ((L + 31) / 32), B, (Cin + BORDER_KC - 1) / BORDER_KC);
    dgrad_border_kernel<<<grid, 128, 0, st>>>(dy, w, dx, Cin, Cout, H, W, mode);
    return sifnn::check_launch("dgrad_border_kernel");
}

extern "C" int sifnn_conv3x3_dgrad_border(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H, int W,
                                          sifnn_stream_t stream) {
    SIFNN_REQUIRE(dy && w && dx && B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && H >= 2 && W >= 2, "conv3x3_dgrad_border: bad arguments");
    cudaStream_t st = sifnn::as_stream(stream);
    SIFNN_TRY(launch_border(dy, w, dx, B, Cin, Cout, H, W, 0, st));
    return launch_border(dy, w, dx, B, Cin, Cout, H, W, 1, st);
}

namespace sifnn {
int conv3x3_dgrad_border_cols(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
    return launch_border(dy, w, dx, B, Cin, Cout, H, W, 2, st);
}
}  // namespace sifnn

extern "C" int sifnn_conv3x3_dgrad_tc(const float* dy, const float* w, float* dx, int accumulate, void* wprep, int B, int Cin, int Cout,
                                      int H, int W, sifnn_stream_t stream) {
    // the tensor-core kernel folds the top/bottom row terms in as extra tap MMAs; only the column terms (+ corners) are left
    SIFNN_TRY(sifnn_conv3x3_dgrad_tc_main(dy, w, dx, accumulate, wprep, B, Cin, Cout, H, W, stream));
    SIFNN_REQUIRE(H >= 2 && W >= 2, "conv3x3_dgrad_tc: bad shape");
    return launch_border(dy, w, dx, B, Cin, Cout, H, W, 2, sifnn::as_stream(stream));
}
