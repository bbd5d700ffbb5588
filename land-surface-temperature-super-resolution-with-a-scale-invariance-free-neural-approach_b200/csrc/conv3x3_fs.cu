// 3x3 convolution for images whose width is a multiple of 128: ky folded into the MMA's N dimension, kx summed by the tensor core through
// three SHIFTED VIEWS of one staged activation row ("fold + shift", round 2).
//
// Measurements behind the design (profiles/r2a_mma_probe2.log, r2j_ncu_full_conv3x3_ff_16x16x256_summary.csv):
//   * an M = 128 tcgen05.mma costs ~111 clocks for any N <= 144 (~134 at N = 192): the number of MMAs is what counts, not their width;
//   * the full-fold kernel (conv3x3_ff.cu: all nine taps in N, kx reduced by the epilogue with warp shuffles and a quadrant exchange)
//     needs only 3 MMAs per 16 input channels but ~6700 warp instructions per 128-pixel row on the CUDA cores: issue-bound
//     (55 % issue utilisation, tensor pipe 7 %).
// Here one input row of 128 (+2 halo) pixels is staged pixel-major, T[ch / 8][pixel][8 bf16]; in the K-major no-swizzle layout an MMA row
// is one 16-byte unit, so "pixel p + kx - 1" is just a start address:
//
//     D[p][(s, ky, o)] += A[p + kx - 1][c] * W[kx][(s, ky, o)][c]            3 kx x 2 MMAs (a_hi x [w_hi ; w_lo], N = 96; a_lo x w_hi, N = 48)
//     out[y] = P1 + D_r[ky = 2]  (complete),  P1 <- P0 + D_r[ky = 1],  P0 <- D_r[ky = 0]        rolling sums over input rows r in registers
//
// so the epilogue is TMEM load -> hi + cross -> two adds -> store: no shuffles, no inter-warp exchange, no barrier, no tile carries.
// Padding: replicate = the halo pixel is a copy of the edge pixel (columns) / the rolling-sum edge rule (rows).  Data gradient: zero halo
// (TMA zero fill) + the adjoint of the replicate padding: rows by the rolling-sum rule (own opposite tap), columns by a correction
// C[ky][o] = sum_c dy[edge pixel][c] * wf[c][o][ky][opposite kx] that the edge warps (or the transformer threads) compute in fp32 per row and the edge lane adds -- no
// border pass, no extra MMAs.
// Two output groups (32 channels) share one MMA pair (N = 192 / 96) when the layer has >= 32 output channels; more go to blockIdx.y.
//
// Warp roles (16 warps): 0..7 epilogue (TMEM lane quadrant = warp % 4, channel half = warp / 4), 8 TMA loader,
// 9 and 15 MMA issuers (even / odd steps: the per-step bookkeeping of one overlaps the MMAs of the other), 10..13 transformers
// (thread = pixel), 14 halo pixels.  Data gradient, column terms of the padding adjoint: + warps 16, 17 (one output group), or the
// transformer threads after their pixel (two groups; see fs_edge_warps).
#include "tc_common.cuh"

#include <cstdlib>
#include <cuda_fp16.h>

namespace {

using namespace sifnn_tc;

constexpr int FS_EPI_WARPS = 8, FS_LOAD_WARP = 8, FS_MMA_WARP = 9, FS_XF_WARP0 = 10, FS_XF_WARPS = 4, FS_HALO_WARP = 14, FS_MMA_WARP2 = 15;
// Data gradient: who computes the column terms of the padding adjoint.  One output group: two extra warps, 16 and 17 (18 warps = five per
// scheduler = a 96-register cap, which that epilogue fits).  Two output groups: the epilogue needs ~120 registers (at 96 it spilled: 2600
// instead of 1500 clocks per step), so the CTA stays at 16 warps and the transformer threads take one or two terms each after their pixel.
// Measured (profiles/r2z_fs_dgrad_edge_variants.log, us, layers 16->16@256 / 32->16@256 / 16->32@256 / 32->32@128 / 64->32@128):
//   two extra warps everywhere 88 / 220 / 135 /  97 / 162;   transformer threads everywhere 96 / 143 / 156 / 67 / 115;
//   halo warp 105 / 139 / 173 / 74 / 132;   one warp instead of the second MMA issuer (two groups) - / 152 / - / 115 / 214.
__host__ __device__ constexpr int fs_edge_warps(int pad, int ng) { return (pad == 1 && ng == 1) ? 2 : 0; }
constexpr int FS_EDGE_WARP0 = 16;
__host__ __device__ constexpr int fs_threads(int pad, int ng) { return (16 + fs_edge_warps(pad, ng)) * 32; }

// One chunk of the column terms: C[e][n] += sum_k dy[edge pixel e][k] * wf[e][k][n] for this thread's items (item = e * HALF + n, dealt round-robin
// to NT threads); after the last chunk the sums go to the edge row `eb` that the epilogue's edge lanes add.
template <int KC, int HALF, int NT, int PER>
__device__ __forceinline__ void fs_edge_chunk(float (&cacc)[PER], int et, int nitems, int ebase, const float* raw0, const float* wedge_s, int K, int c, int rpx,
                                              int Wt, bool last, float* eb) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if (et + i * NT < nitems) {
            const int item = ebase + et + i * NT;
            const float* dyp = raw0 + (item < HALF ? 4 : 3 + Wt);
            const float* we = wedge_s + (size_t)(item < HALF ? 0 : K - 1) * HALF + (size_t)c * KC * HALF + item;   // [e][k][n]
            float dv[KC], wv[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) { dv[k] = dyp[k * rpx]; wv[k] = we[k * HALF]; }
            float s0 = cacc[i], s1 = 0.f;
#pragma unroll
            for (int k = 0; k < KC; k += 2) { s0 = fmaf(dv[k], wv[k], s0); s1 = fmaf(dv[k + 1], wv[k + 1], s1); }
            float sacc = s0 + s1;
            if (last) { eb[item] = sacc; sacc = 0.f; }
            cacc[i] = sacc;
        }
    }
}
constexpr int FS_PX = 136;                 // staged pixels of a row piece: x0 - 4 .. x0 + 131 (16-byte aligned both ends); MMA row m of view kx = pixel index 3 + kx + m
constexpr int FS_A_TILE = 2 * FS_PX * 16;  // bytes of one (hi or lo) activation tile of a chunk: [2 q][136 pixels][16 B]
constexpr int FS_MAX_K = 64;               // input channels per launch
constexpr int FS_ES = 16;                  // ring of edge-correction rows (the halo warp runs at most operand ring + TMEM ring steps ahead of the epilogue)

#define FS_STAMP(ev, idx) do { if (DBG && a.trace && blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 256) a.trace[(ev) * 256 + (idx)] = clock64(); } while (0)

struct FsArgs {
    sifnn::BnTail tail;   // optional fused BatchNorm finalize (forward with statistics)
    int ablate;   // timing experiments only (SIFNN_FS_ABLATE): 1 skip the edge-column math, 2 skip the global stores, 4 skip the accumulate loads
    const float* in_scale;
    const float* in_shift;
    const unsigned char* wprep;   // [gridDim.y][chunk][kx][2 q][(s, ky, g, o) = 96 NG rows][16 B]
    const float* wedge;           // data gradient only: [gridDim.y][edge][K][(ky, g, o) = 48 NG] fp32
    float* out;
    double* stats;
    int B, K, O, H, W;
    int accumulate;
    int nrows;    // B * H
    int K1;       // channels that come from the first tensor map (== K without a second source)
    unsigned long long* trace;   // debug (sifnn_conv3x3_fs_trace): clock64 stamps of CTA (0,0), [event][step]
};

struct FsLayout {
    int edge, wedge, stat, w, a, raw, total;
    int AS, RS, raw_stage;
};
__host__ __device__ inline FsLayout fs_layout(int nchunks, int NG, int KC, bool pad1) {
    FsLayout L{};
    int off = 1024;               // [0, 1024): mbarriers + TMEM slot
    off += 2 * FS_MAX_K * 4;      // BatchNorm scale / shift of the input channels
    L.edge = off; if (pad1) off += FS_ES * 2 * 48 * NG * 4;
    L.wedge = off; if (pad1) off += 2 * nchunks * KC * 48 * NG * 4;   // fp32 weights of the column terms of the padding adjoint (data gradient)
    L.stat = off; if (!pad1) off += 4 * 2 * 16 * NG * 4;   // per-quadrant partial BatchNorm sums of the CTA
    off = (off + 1023) & ~1023;
    L.w = off; off += nchunks * 3 * 2 * 96 * NG * 16;
    L.AS = (NG == 1) ? (2 * nchunks < 8 ? (2 * nchunks < 4 ? 4 : 2 * nchunks) : 8) : 2 * nchunks;
    L.a = off; off += L.AS * 2 * FS_A_TILE;
    L.raw_stage = KC * FS_PX * 4;
    L.RS = 4;
    while (L.RS > 2 && off + L.RS * L.raw_stage > 224 * 1024) --L.RS;
    L.raw = off; off += L.RS * L.raw_stage;
    L.total = off;
    return L;
}

// A CTA's walk over its share of the B * H image rows: strips (one image, output rows [ya, yb)) x the T = W / 128 pieces of a row x input rows.
struct FsIter {
    int H, T, g1, b, ya, yb, t, r, rfirst, rlast;
    bool active;
    __device__ void begin(int g0) {
        if (g0 >= g1) { active = false; return; }
        b = g0 / H;
        ya = g0 - b * H;
        yb = min(g1 - b * H, H);
        t = 0;
        rfirst = max(ya - 1, 0);
        rlast = min(yb, H - 1);
        r = rfirst;
        active = true;
    }
    __device__ void init(int worker, int nworkers, int nrows, int H_, int T_) {
        H = H_; T = T_;
        const int g0 = (int)((long long)worker * nrows / nworkers);
        g1 = (int)((long long)(worker + 1) * nrows / nworkers);
        begin(g0);
    }
    __device__ void next() {
        if (!active) return;
        if (r < rlast) { ++r; return; }
        if (t + 1 < T) { ++t; r = rfirst; return; }
        begin(b * H + yb);
    }
    __device__ int count() const {
        FsIter c = *this;
        int n = 0;
        while (c.active) { n += c.T * (c.rlast - c.rfirst + 1); c.begin(c.b * c.H + c.yb); }
        return n;
    }
};

// position in a ring of n mbarrier-guarded slots without integer division: slot index, parity of the current use, and whether the slot was used before
struct FsRing {
    int idx, n;
    uint32_t phase;
    bool wrapped;
    __device__ explicit FsRing(int n_) : idx(0), n(n_), phase(0), wrapped(false) {}
    __device__ void next() { if (++idx == n) { idx = 0; phase ^= 1; wrapped = true; } }
};

__device__ __forceinline__ uint32_t fs_pack_bf16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

// FP16 pair: pack (round to nearest) and unpack
__device__ __forceinline__ uint32_t fs_pack_f16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void fs_unpack_f16x2(uint32_t h, float& lo_elem, float& hi_elem) {
    asm("{\n\t.reg .f16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(lo_elem), "=f"(hi_elem) : "r"(h));
}
constexpr float FS_LO_SCALE = 2048.f;   // FP16 split: the residual (and the weights' residual) is stored times 2^11, the cross block is rescaled by the epilogue

// one pixel of one chunk: raw fp32 (channel stride FS_PX) -> BatchNorm affine + ReLU -> hi / lo split -> the two 16-byte units of the pixel
template <int KIND, bool AFFINE>
__device__ __forceinline__ void fs_convert_pixel(const float* raw, int rpx, unsigned char* a_hi, int jdst, const float* sc, const float* sh) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if constexpr (KIND != 1) {
            uint32_t hp[4], lp[4];
            float scv[8], shv[8];   // the eight channels' BatchNorm scale / shift as four 16-byte broadcast loads instead of sixteen scalar ones
            if (AFFINE) {
                const float4 s0 = *reinterpret_cast<const float4*>(sc + 8 * q), s1 = *reinterpret_cast<const float4*>(sc + 8 * q + 4);
                const float4 h0 = *reinterpret_cast<const float4*>(sh + 8 * q), h1 = *reinterpret_cast<const float4*>(sh + 8 * q + 4);
                scv[0] = s0.x; scv[1] = s0.y; scv[2] = s0.z; scv[3] = s0.w; scv[4] = s1.x; scv[5] = s1.y; scv[6] = s1.z; scv[7] = s1.w;
                shv[0] = h0.x; shv[1] = h0.y; shv[2] = h0.z; shv[3] = h0.w; shv[4] = h1.x; shv[5] = h1.y; shv[6] = h1.z; shv[7] = h1.w;
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                float t0 = raw[(8 * q + e) * rpx], t1 = raw[(8 * q + e + 1) * rpx];
                if (AFFINE) {
                    t0 = sifnn::act_affine_relu(t0, scv[e], shv[e]);
                    t1 = sifnn::act_affine_relu(t1, scv[e + 1], shv[e + 1]);
                }
                if constexpr (KIND == 0) {
                    const uint32_t h = fs_pack_bf16x2(t0, t1);
                    const float r0 = t0 - __uint_as_float(h << 16), r1 = t1 - __uint_as_float(h & 0xffff0000u);
                    hp[e >> 1] = h;
                    lp[e >> 1] = fs_pack_bf16x2(r0, r1);
                } else {
                    const uint32_t h = fs_pack_f16x2(t0, t1);
                    float f0, f1;
                    fs_unpack_f16x2(h, f0, f1);
                    hp[e >> 1] = h;
                    lp[e >> 1] = fs_pack_f16x2((t0 - f0) * FS_LO_SCALE, (t1 - f1) * FS_LO_SCALE);
                }
            }
            *reinterpret_cast<uint4*>(a_hi + (size_t)(q * FS_PX + jdst) * 16) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4*>(a_hi + FS_A_TILE + (size_t)(q * FS_PX + jdst) * 16) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        } else {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = raw[(4 * q + e) * rpx];
                if (AFFINE) t = sifnn::act_affine_relu(t, sc[4 * q + e], sh[4 * q + e]);
                hi[e] = tf32_hi(t);
                lo[e] = t - hi[e];
            }
            *reinterpret_cast<float4*>(a_hi + (size_t)(q * FS_PX + jdst) * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a_hi + FS_A_TILE + (size_t)(q * FS_PX + jdst) * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

template <int KIND, int NG, int PAD, bool AFFINE, bool STATS, bool DBG, int MM>
__global__ void __launch_bounds__(fs_threads(PAD, NG), 1) conv3x3_fs_kernel(const FsArgs a, const __grid_constant__ CUtensorMap tmap,
                                                                   const __grid_constant__ CUtensorMap tmap2) {
    constexpr int KC = (KIND == 1) ? 8 : 16;
    constexpr int EDGE_WARPS = fs_edge_warps(PAD, NG), EDGE_WARP0 = FS_EDGE_WARP0;
    constexpr bool EDGE_IN_XF = (PAD == 1 && NG == 2);
    constexpr int NPROD = FS_XF_WARPS + 1 + EDGE_WARPS;   // warps that read a raw stage and arrive on the operand barrier
    constexpr int NSLOT = (NG == 1) ? 5 : 2;     // TMEM ring: 96 NG columns per row piece
    constexpr int SLOT = 96 * NG, HALF = 48 * NG;
    constexpr int W_TILE = 2 * SLOT * 16;        // bytes of the weight tile of one (chunk, kx)
    sifnn::pdl_wait_and_trigger();   // launched with launch_pdl: every global access below comes after the previous kernel of the stream
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nchunks = a.K / KC;
    const FsLayout L = fs_layout(nchunks, NG, KC, PAD == 1);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* raw_full = bars;            // [4]
    uint64_t* raw_empty = bars + 4;       // [4]
    uint64_t* a_full = bars + 8;          // [16]
    uint64_t* a_empty = bars + 24;        // [16]
    uint64_t* acc_full = bars + 40;       // [5]
    uint64_t* acc_empty = bars + 45;      // [5]
    uint64_t* w_full = bars + 50;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 52);
    float* sc_s = reinterpret_cast<float*>(smem + 1024);
    float* sh_s = sc_s + FS_MAX_K;
    float* edge_s = reinterpret_cast<float*>(smem + L.edge);   // [FS_ES][2 edges][48 NG]
    float* wedge_s = reinterpret_cast<float*>(smem + L.wedge); // [2 edges][K][48 NG]
    unsigned char* w_s = smem + L.w;
    unsigned char* a_s = smem + L.a;
    unsigned char* raw_s = smem + L.raw;
    const int AS = L.AS, RS = L.RS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W;
    // pixels of a row piece (MM = 64 serves the 64- and 32-pixel-wide levels: one image row per MMA, junk rows beyond).  A compile-time constant in the
    // M = 128 form, so the channel stride of the raw stage (rpx) folds into the immediate offsets of the transformers' shared-memory loads
    const int Wt = (MM == 128) ? 128 : (W < 128 ? W : 128);
    const int T = W / Wt;
    const int rpx = Wt + 8;                    // staged pixels per channel of a raw box: x0 - 4 .. x0 + Wt + 3
    const uint32_t raw_bytes = (uint32_t)(KC * rpx * 4);
    const int nworkers = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, NPROD); }
        for (int s = 0; s < AS; ++s) { mbar_init(a_full + s, NPROD); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < NSLOT; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, FS_EPI_WARPS); }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == FS_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (PAD == 1) {   // resident: with ~200 KB of shared memory there is next to no L1 left and every __ldg of these weights was an L2 round trip
        const float* wg = a.wedge + (size_t)blockIdx.y * 2 * a.K * HALF;
        for (int i = tid; i < 2 * a.K * HALF; i += fs_threads(PAD, NG)) wedge_s[i] = __ldg(wg + i);
    }
    if (AFFINE) {
        for (int i = tid; i < a.K; i += fs_threads(PAD, NG)) { sc_s[i] = a.in_scale ? __ldg(a.in_scale + i) : 1.f; sh_s[i] = a.in_shift ? __ldg(a.in_shift + i) : 0.f; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t w_bytes = (uint32_t)(nchunks * 3 * W_TILE);
    if (tid == 0) {   // the split weights of this CTA's output channels stay resident for the whole kernel
        mbar_arrive_expect_tx(w_full, w_bytes);
        bulk_g2s(w_s, a.wprep + (size_t)blockIdx.y * w_bytes, w_bytes, w_full);
    }
    FsIter it;
    it.init(blockIdx.x, nworkers, a.nrows, H, T);
    const int nsteps = it.count();

    if (warp == FS_LOAD_WARP) {
        // ======================= loader: one TMA box {136 pixels, 1 row, KC planes} per chunk; out-of-image pixels arrive as zeros =======================
        FsRing rr(RS);
        for (int step = 0; step < nsteps; ++step) {
            for (int c = 0; c < nchunks; ++c, rr.next()) {
                const int rs = rr.idx;
                if (lane == 0 && c == 0) FS_STAMP(11, step);
                if (rr.wrapped) mbar_wait(raw_empty + rs, rr.phase ^ 1);
                if (lane == 0) {
                    if (c == 0) FS_STAMP(0, step);
                    mbar_arrive_expect_tx(raw_full + rs, raw_bytes);
                    const int ch = c * KC;
                    unsigned char* dst = raw_s + (size_t)rs * L.raw_stage;
                    if (ch < a.K1) tma_load_3d(dst, &tmap, it.t * 128 - 4, it.r, it.b * a.K1 + ch, raw_full + rs);
                    else tma_load_3d(dst, &tmap2, it.t * 128 - 4, it.r, it.b * (a.K - a.K1) + (ch - a.K1), raw_full + rs);
                }
                __syncwarp();
            }
            it.next();
        }
    } else if (warp == FS_MMA_WARP || warp == FS_MMA_WARP2) {
        // ======================= MMA issuers (one thread each, alternating steps): 3 shifted views x (a_hi x [w_hi ; w_lo], a_lo x w_hi) per chunk =======================
        // The tensor pipe executes one of these MMAs per ~111 clocks and its queue is shallow, so every clock this thread spends between two
        // steps is a clock the pipe idles: one thread does everything (no warp-wide polls, no __syncwarp), its waits spin, and the descriptors
        // are advanced by constants instead of being rebuilt.
        constexpr uint32_t idesc1 = KIND == 0 ? make_idesc_bf16(MM, SLOT) : (KIND == 1 ? make_idesc(MM, SLOT) : make_idesc_f16(MM, SLOT));
        constexpr uint32_t idesc2 = KIND == 0 ? make_idesc_bf16(MM, HALF) : (KIND == 1 ? make_idesc(MM, HALF) : make_idesc_f16(MM, HALF));
        if (lane == 0) {
            mbar_wait_spin(w_full, 0);
            // start-address field = low 14 bits in 16-byte units; every offset below stays inside the 256 KB window, so plain 64-bit adds advance it
            const uint64_t da0 = make_desc(smem_u32(a_s) + 3 * 16, FS_PX * 16, 128);
            const uint64_t dw0 = make_desc(smem_u32(w_s), SLOT * 16, 128);
            FsRing ra(AS), rc(NSLOT);
            // Two issuers only when an operand slot comes back to the SAME issuer (slot period AS / nchunks even): a parity wait must never be two
            // phases behind, and the issuers do not wait for each other.  (TF32 with 64 input channels has period 1: issuer 0 does every step.)
            const bool two = ((AS / nchunks) & 1) == 0;
            const int mine = (warp == FS_MMA_WARP) ? 0 : 1;
            for (int step = 0; step < nsteps; ++step, rc.next()) {
                if (two ? ((step & 1) != mine) : (mine != 0)) {   // the other issuer's step: only advance the rings
                    for (int c = 0; c < nchunks; ++c) ra.next();
                    continue;
                }
                const int slot = rc.idx;
                FS_STAMP(12, step);
                if (rc.wrapped) mbar_wait_spin(acc_empty + slot, rc.phase ^ 1);
                FS_STAMP(4, step);
                const uint32_t d = tmem_base + slot * SLOT;
                for (int c = 0; c < nchunks; ++c, ra.next()) {
                    const int as = ra.idx;
                    mbar_wait_spin(a_full + as, ra.phase);
                    tc_fence_after();
                    if (c == 0) FS_STAMP(5, step);
                    const uint64_t da_hi = da0 + (uint64_t)(as * (2 * FS_A_TILE / 16));
                    const uint64_t da_lo = da_hi + (uint64_t)(FS_A_TILE / 16);
                    const uint64_t dw = dw0 + (uint64_t)(c * 3 * (W_TILE / 16));
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        if (KIND != 1) {
                            umma_bf16(d, da_hi + kx, dw + kx * (W_TILE / 16), idesc1, (c > 0 || kx > 0) ? 1u : 0u);   // hi x [hi ; lo]: also initialises the cross block
                            umma_bf16(d + HALF, da_lo + kx, dw + kx * (W_TILE / 16), idesc2, 1u);
                        } else {
                            umma_tf32(d, da_hi + kx, dw + kx * (W_TILE / 16), idesc1, (c > 0 || kx > 0) ? 1u : 0u);
                            umma_tf32(d + HALF, da_lo + kx, dw + kx * (W_TILE / 16), idesc2, 1u);
                        }
                    }
                    umma_commit(a_empty + as);   // chunk slot reusable once these MMAs have read it
                }
                umma_commit(acc_full + slot);
                FS_STAMP(6, step);
            }
        }
        __syncwarp();
    } else if (warp >= FS_XF_WARP0 && warp < FS_HALO_WARP) {
        // ======================= transformers: thread = pixel x0 + p (+ column terms of the padding adjoint, two-group data gradient) =======================
        const int p = tid - FS_XF_WARP0 * 32;
        float cacc[2] = {0.f, 0.f};
        FsRing rr(RS), ra(AS);
        for (int step = 0; step < nsteps; ++step) {
            int nitems = 0, ebase = 0;
            if (EDGE_IN_XF && !(a.ablate & 1)) {
                const bool img_l = (it.t == 0), img_r = (it.t == T - 1);
                if (img_l && img_r) nitems = 2 * HALF;
                else if (img_l || img_r) { nitems = HALF; ebase = img_l ? 0 : HALF; }
            }
            for (int c = 0; c < nchunks; ++c, rr.next(), ra.next()) {
                const int rs = rr.idx, as = ra.idx;
                if (p == 0 && c == 0) FS_STAMP(13, step);
                if (ra.wrapped) mbar_wait(a_empty + as, ra.phase ^ 1);
                if (p == 0 && c == 0) FS_STAMP(1, step);
                mbar_wait(raw_full + rs, rr.phase);
                if (p == 0 && c == 0) FS_STAMP(2, step);
                const float* raw0 = reinterpret_cast<const float*>(raw_s + (size_t)rs * L.raw_stage);
                if (p < Wt) fs_convert_pixel<KIND, AFFINE>(raw0 + 4 + p, rpx, a_s + (size_t)as * 2 * FS_A_TILE, 4 + p, sc_s + c * KC, sh_s + c * KC);
                if (EDGE_IN_XF)
                    fs_edge_chunk<KC, HALF, FS_XF_WARPS * 32, 2>(cacc, p, nitems, ebase, raw0, wedge_s, a.K, c, rpx, Wt, c == nchunks - 1,
                                                                 edge_s + (size_t)(step & (FS_ES - 1)) * 2 * HALF);
                fence_proxy_async();           // this thread's st.shared -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) { mbar_arrive(raw_empty + rs); mbar_arrive(a_full + as); }   // (the edge row is published before the MMAs of this step may start)
                if (p == 0 && c == nchunks - 1) FS_STAMP(3, step);
            }
            if (EDGE_IN_XF) it.next();
        }
    } else if (warp == FS_HALO_WARP) {
        // ======================= halo pixels: lanes 0 and 1 convert pixel x0 - 1 and x0 + 128 =======================
        FsRing rr(RS), ra(AS);
        for (int step = 0; step < nsteps; ++step) {
            const bool img_l = (it.t == 0), img_r = (it.t == T - 1);
            for (int c = 0; c < nchunks; ++c, rr.next(), ra.next()) {
                const int rs = rr.idx, as = ra.idx;
                if (ra.wrapped) mbar_wait(a_empty + as, ra.phase ^ 1);
                mbar_wait(raw_full + rs, rr.phase);
                const float* raw = reinterpret_cast<const float*>(raw_s + (size_t)rs * L.raw_stage);
                if (lane < 2) {
                    // replicate padding: the halo pixel outside the image is a copy of the edge pixel; zero padding: it arrived as zeros
                    const int jdst = lane == 0 ? 3 : 4 + Wt;
                    int jsrc = jdst;
                    if (PAD == 0) { if (lane == 0 && img_l) jsrc = 4; if (lane == 1 && img_r) jsrc = 3 + Wt; }
                    fs_convert_pixel<KIND, AFFINE>(raw + jsrc, rpx, a_s + (size_t)as * 2 * FS_A_TILE, jdst, sc_s + c * KC, sh_s + c * KC);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(raw_empty + rs); mbar_arrive(a_full + as); }
            }
            it.next();
        }
    } else if (PAD == 1 && warp >= EDGE_WARP0 && warp < EDGE_WARP0 + EDGE_WARPS) {
        // ======================= data gradient: column terms of the padding adjoint, fp32 on the CUDA cores =======================
        // C[e][n] = sum_k dy[edge pixel e][k] * wf[e][k][n]   (n = (ky, g, o); the tap that would have left the image comes back onto the edge pixel).
        // A row piece has one edge (W > 128: 48 NG items) or two (W = 128: 96 NG items), dealt round-robin to the threads; the weights are resident in
        // shared memory (with ~200 KB of it in use there is next to no L1 left: as __ldg they were L2 round trips).
        constexpr int NT = (EDGE_WARPS > 0 ? EDGE_WARPS : 1) * 32, PER = (2 * HALF + NT - 1) / NT;
        const int et = tid - EDGE_WARP0 * 32;
        float cacc[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) cacc[i] = 0.f;
        FsRing rr(RS), ra(AS);
        for (int step = 0; step < nsteps; ++step) {
            const bool img_l = (it.t == 0), img_r = (it.t == T - 1);
            int nitems = 0, ebase = 0;   // item = e * HALF + n
            if (!(a.ablate & 1)) {
                if (img_l && img_r) nitems = 2 * HALF;
                else if (img_l || img_r) { nitems = HALF; ebase = img_l ? 0 : HALF; }
            }
            for (int c = 0; c < nchunks; ++c, rr.next(), ra.next()) {
                const int rs = rr.idx, as = ra.idx;
                if (et == 0 && c == 0) FS_STAMP(9, step);
                if (ra.wrapped) mbar_wait(a_empty + as, ra.phase ^ 1);   // same gate as the other producers: never two arrivals in one phase
                mbar_wait(raw_full + rs, rr.phase);
                const float* raw0 = reinterpret_cast<const float*>(raw_s + (size_t)rs * L.raw_stage);
                fs_edge_chunk<KC, HALF, NT, PER>(cacc, et, nitems, ebase, raw0, wedge_s, a.K, c, rpx, Wt, c == nchunks - 1, edge_s + (size_t)(step & (FS_ES - 1)) * 2 * HALF);
                __syncwarp();
                if (lane == 0) { mbar_arrive(raw_empty + rs); mbar_arrive(a_full + as); }   // the edge row is published before the MMAs of this step may start
                if (et == 0 && c == nchunks - 1) FS_STAMP(15, step);
            }
            it.next();
        }
    } else {
        // ======================= epilogue: TMEM -> hi + cross -> rolling ky sums -> global (+ BatchNorm statistics) =======================
        const int quad = warp & 3, half = warp >> 2;
        // M = 128: accumulator row m lives in TMEM lane m.  M = 64: row m lives in lane (m % 16) + 32 * (m / 16), the lower 16 lanes of each quadrant.
        const int p = (MM == 128) ? quad * 32 + lane : quad * 16 + (lane & 15);
        const bool px_ok = (MM == 128 || lane < 16) && p < Wt;
        float P0[NG][8], P1[NG][8];
        float s1[STATS ? NG : 1][8], s2[STATS ? NG : 1][8];
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 8; ++j) { P0[g][j] = 0.f; P1[g][j] = 0.f; if constexpr (STATS) { s1[g][j] = 0.f; s2[g][j] = 0.f; } }
        const bool accum = a.accumulate != 0;
        const size_t plane = (size_t)H * W;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + half * 8;
        FsRing rc(NSLOT);
        for (int step = 0; step < nsteps; ++step, rc.next()) {
            const int b = it.b, r = it.r, ya = it.ya, yb = it.yb;
            const int x = it.t * 128 + p;
            const bool emit_prev = px_ok && (r - 1 >= ya);
            const bool emit_last = px_ok && (r == H - 1) && (yb == H);
            const int slot = rc.idx;
            if (tid == 0) FS_STAMP(14, step);
            mbar_wait(acc_full + slot, rc.phase);   // suspending wait: the epilogue runs behind the MMAs with five slots of slack, a spin would only take issue slots from the transformers
            tc_fence_after();
            if (tid == 0) FS_STAMP(7, step);
            const uint32_t tcol = tlane + slot * SLOT;
            float* const orow = a.out + ((size_t)b * a.O + blockIdx.y * NG * 16 + half * 8) * plane + x;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float Gk[3][8];
                {
                    float hi[3][8], cr[3][8];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        tmem_ld8(tcol + (ky * NG + g) * 16, hi[ky]);
                        tmem_ld8(tcol + HALF + (ky * NG + g) * 16, cr[ky]);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int j = 0; j < 8; ++j) Gk[ky][j] = (KIND == 2) ? fmaf(cr[ky][j], 1.f / FS_LO_SCALE, hi[ky][j]) : hi[ky][j] + cr[ky][j];
                }
                if (g == NG - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + slot);   // the slot may be overwritten (one arrival per warp)
                    if (tid == 0) FS_STAMP(8, step);
                }
                if (PAD == 1) {   // column terms of the padding adjoint, computed by the halo warp for this row
                    if (px_ok && (x == 0 || x == W - 1)) {
                        const float* eb = edge_s + (size_t)(step & (FS_ES - 1)) * 2 * HALF + (x == 0 ? 0 : HALF) + g * 16 + half * 8;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int j = 0; j < 8; ++j) Gk[ky][j] += eb[ky * NG * 16 + j];
                    }
                }
                // rolling sums over ky.  Top edge: the missing row above is replaced by this row's own ky = 0 (forward) / ky = 2 (data gradient) term.
                constexpr int ET = (PAD == 0) ? 0 : 2, EB = (PAD == 0) ? 2 : 0;
                if (r == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) P0[g][j] = Gk[ET][j];
                }
                float o_prev[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o_prev[j] = P1[g][j] + Gk[2][j];
                    P1[g][j] = P0[g][j] + Gk[1][j];
                    P0[g][j] = Gk[0][j];
                }
                auto emit = [&](int row, const float* v) {
                    float* op = orow + (size_t)(g * 16) * plane + (size_t)row * W;
                    float old[8];
                    if (a.ablate & 2) return;
                    if (accum && !(a.ablate & 4)) {   // all eight loads first: one memory round trip instead of eight dependent ones
#pragma unroll
                        for (int j = 0; j < 8; ++j) old[j] = __ldcg(op + (size_t)j * plane);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j, op += plane) {
                        float o = v[j];
                        if (accum) o += old[j];
                        *op = o;
                        if constexpr (STATS) { s1[g][j] += o; s2[g][j] = fmaf(o, o, s2[g][j]); }
                    }
                };
                if (emit_prev) emit(r - 1, o_prev);
                if (emit_last) {
                    float o_last[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o_last[j] = P1[g][j] + Gk[EB][j];
                    emit(H - 1, o_last);
                }
            }
            if (tid == 0) FS_STAMP(10, step);
            it.next();
        }
        if constexpr (STATS) if (a.stats) {
            // per-channel sums of this CTA: lanes by shuffle, the four quadrant warps through shared memory in a fixed order, then ONE fp64 atomic per
            // channel and CTA (148 per address instead of 592: the serialised atomics were ~10 us of a 16-channel layer)
            float* part = reinterpret_cast<float*>(smem + L.stat);   // [quad][stat][NG * 16]
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t1 = sifnn::warp_sum(s1[g][j]), t2 = sifnn::warp_sum(s2[g][j]);
                    if (lane == 0) {
                        part[(quad * 2 + 0) * 16 * NG + g * 16 + half * 8 + j] = t1;
                        part[(quad * 2 + 1) * 16 * NG + g * 16 + half * 8 + j] = t2;
                    }
                }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps
            if (tid < 2 * 16 * NG) {
                const int stat = tid / (16 * NG), ch = tid - stat * 16 * NG;
                const double v = (double)part[(0 * 2 + stat) * 16 * NG + ch] + (double)part[(1 * 2 + stat) * 16 * NG + ch] + (double)part[(2 * 2 + stat) * 16 * NG + ch] +
                                 (double)part[(3 * 2 + stat) * 16 * NG + ch];
                atomicAdd(a.stats + (size_t)stat * a.O + blockIdx.y * NG * 16 + ch, v);
                __threadfence();   // ordered before this CTA's ticket of the BatchNorm tail
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FS_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if constexpr (STATS) {
        if (a.stats && a.tail.counter) {
            __shared__ int tail_flag;
            sifnn::bn_tail_finalize(a.tail, a.stats, a.O, gridDim.x * gridDim.y, &tail_flag);
        }
    }
}

int g_fs_max_ctas = 0;
unsigned long long* g_fs_trace = nullptr;
// 64- and 32-pixel-wide levels with M = 64 MMAs (one image row each): OFF by default.  Measured (profiles/r2v_profile_fs_narrow.log): an M = 64
// MMA costs the tensor core as much as an M = 128 one, so the full-fold kernel, which packs two / four rows into every M = 128 MMA, is faster
// there (step 4.61 -> 5.02 ms with this path on).  SIFNN_FS_NARROW=1 or sifnn_conv3x3_fs_narrow(1) enables it (tests, A/B runs).
int g_fs_narrow = -1;
bool fs_narrow() {
    if (g_fs_narrow < 0) { const char* e = getenv("SIFNN_FS_NARROW"); g_fs_narrow = (e && e[0] == '1') ? 1 : 0; }
    return g_fs_narrow == 1;
}
bool fs_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_FS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <int KIND, int NG, int PAD, bool AFFINE, bool STATS, int MM>
int launch_fs(const FsArgs& a, const CUtensorMap& tm1, const CUtensorMap& tm2, int gx, int gy, cudaStream_t st) {
    constexpr int KC = (KIND == 1) ? 8 : 16;
    const FsLayout L = fs_layout(a.K / KC, NG, KC, PAD == 1);
    auto kern = conv3x3_fs_kernel<KIND, NG, PAD, AFFINE, STATS, false, MM>;
    if (a.trace) {
        if constexpr (!AFFINE && !STATS && MM == 128 && KIND != 1) kern = conv3x3_fs_kernel<KIND, NG, PAD, AFFINE, STATS, true, MM>;   // traced build: plain forward and data gradient
    }
    SIFNN_REQUIRE(L.total <= 227 * 1024, "conv3x3_fs: shared-memory budget exceeded (K=%d)", a.K);
    SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    SIFNN_CUDA(sifnn::launch_pdl(kern, dim3(gx, gy), dim3(fs_threads(PAD, NG)), (size_t)L.total, st, a, tm1, tm2));
    return sifnn::check_launch("conv3x3_fs_kernel");
}

template <int KIND, int PAD, bool AFFINE, bool STATS>
int dispatch_fs2(const FsArgs& a, const CUtensorMap& tm1, const CUtensorMap& tm2, int NG, int gx, int gy, cudaStream_t st) {
    if (a.W < 128) {   // 64- and 32-pixel-wide levels: M = 64 MMAs, one image row each
        if (NG == 1) return launch_fs<KIND, 1, PAD, AFFINE, STATS, 64>(a, tm1, tm2, gx, gy, st);
        return launch_fs<KIND, 2, PAD, AFFINE, STATS, 64>(a, tm1, tm2, gx, gy, st);
    }
    if (NG == 1) return launch_fs<KIND, 1, PAD, AFFINE, STATS, 128>(a, tm1, tm2, gx, gy, st);
    return launch_fs<KIND, 2, PAD, AFFINE, STATS, 128>(a, tm1, tm2, gx, gy, st);
}
template <int KIND>
int dispatch_fs1(int pad, bool affine, bool stats, const FsArgs& a, const CUtensorMap& tm1, const CUtensorMap& tm2, int NG, int gx, int gy, cudaStream_t st) {
    if (pad == 1) return dispatch_fs2<KIND, 1, false, false>(a, tm1, tm2, NG, gx, gy, st);
    if (affine) return stats ? dispatch_fs2<KIND, 0, true, true>(a, tm1, tm2, NG, gx, gy, st) : dispatch_fs2<KIND, 0, true, false>(a, tm1, tm2, NG, gx, gy, st);
    return stats ? dispatch_fs2<KIND, 0, false, true>(a, tm1, tm2, NG, gx, gy, st) : dispatch_fs2<KIND, 0, false, false>(a, tm1, tm2, NG, gx, gy, st);
}

// Forward shapes of the 64-pixel level that the M = 64 form wins against the full-fold kernel (tools/variants_fs.py, affine + statistics, B = 32):
// 32 -> 32 @64 33.0 vs 41.1 us, 32 -> 64 @64 51.8 vs 56.8 us.  (64 input channels, the 32-pixel level and every data gradient stay on conv3x3_ff.)
bool fs_fwd_narrow_wins(int K, int O, int H, int W) { return W == 64 && K == 32 && (O == 32 || O == 64) && H >= 1; }
bool fs_shape_ok(int K, int O, int H, int W) {
    return ((W % 128 == 0 && W >= 128 && W <= 4096) || ((W == 64 || W == 32) && fs_narrow())) && (K % 16 == 0) && K >= 16 && K <= FS_MAX_K && (O % 16 == 0) && O >= 16 && O <= 128 && (O == 16 || O % 32 == 0) && H >= 1;
}

int fs_groups(int K, int O, int kind) {   // output groups of 16 channels per CTA
    if (O < 32) return 1;
    if (kind == 1 && K > 32) return 1;          // TF32: 8 chunks of weights + operand ring do not fit with two groups
    return 2;
}

int run_fs(const sifnn::BnTail* tail, int pad, const float* in, const float* in2, int K1, const float* in_scale, const float* in_shift, const void* wprep, const float* wedge, float* out,
           double* stats, int accumulate, int B, int K, int O, int H, int W, cudaStream_t st) {
    SIFNN_REQUIRE(fs_shape_ok(K, O, H, W) || (pad == 0 && fs_fwd_narrow_wins(K, O, H, W)), "conv3x3_fs: unsupported shape K=%d O=%d H=%d W=%d", K, O, H, W);
    const int kind = sifnn::tc_split_kind(pad);
    SIFNN_REQUIRE(!(pad == 1 && kind == 2), "conv3x3_fs: the FP16 split is for the forward form only (gradients can be tiny)");
    const int KC = (kind == 1) ? 8 : 16;
    SIFNN_REQUIRE(!in2 || (K1 % KC == 0 && K1 > 0 && K1 < K), "conv3x3_fs: the split point of a two-source input must be a multiple of %d", KC);
    FsArgs a{};
    a.in_scale = in_scale; a.in_shift = in_shift; a.wprep = static_cast<const unsigned char*>(wprep); a.wedge = wedge; a.out = out; a.stats = stats;
    a.B = B; a.K = K; a.O = O; a.H = H; a.W = W; a.accumulate = accumulate ? 1 : 0;
    a.nrows = B * H;
    a.K1 = in2 ? K1 : K;
    a.trace = g_fs_trace;
    if (tail && stats) a.tail = *tail;
    { static int ablate = -1; if (ablate < 0) { const char* e = getenv("SIFNN_FS_ABLATE"); ablate = e ? atoi(e) : 0; } a.ablate = ablate; }   // read once per process
    const int NG = fs_groups(K, O, kind);
    const int gy = O / (16 * NG);
    int gx = sifnn::num_sms() / gy;
    if (g_fs_max_ctas > 0 && gx > g_fs_max_ctas) gx = g_fs_max_ctas;   // tests: long strips on small inputs
    if (gx > a.nrows) gx = a.nrows;
    if (gx < 1) gx = 1;
    CUtensorMap tm1, tm2;
    const int rpx = (W < 128 ? W : 128) + 8;
    SIFNN_REQUIRE(encode_planes_map(&tm1, in, W, H, (long long)B * a.K1, rpx, 1, KC), "conv3x3_fs: cuTensorMapEncodeTiled is unavailable or failed");
    if (in2) SIFNN_REQUIRE(encode_planes_map(&tm2, in2, W, H, (long long)B * (K - K1), rpx, 1, KC), "conv3x3_fs: cuTensorMapEncodeTiled failed (second source)");
    else tm2 = tm1;
    const bool affine = in_scale != nullptr;
    SIFNN_REQUIRE(pad == 0 || wedge, "conv3x3_fs: the data-gradient form needs the edge weights");
    if (kind == 0) return dispatch_fs1<0>(pad, affine, stats != nullptr, a, tm1, tm2, NG, gx, gy, st);
    if (kind == 1) return dispatch_fs1<1>(pad, affine, stats != nullptr, a, tm1, tm2, NG, gx, gy, st);
    return dispatch_fs1<2>(pad, affine, stats != nullptr, a, tm1, tm2, NG, gx, gy, st);
}

// Split weights in the exact shared-memory image of the kernel:
//   wprep [gy][chunk][kx][2 q][row = (s, ky, g, o): s * 48 NG + (ky * NG + g) * 16 + o][8 bf16 | 4 tf32]
//   wedge [gy][edge][k][(ky * NG + g) * 16 + o] fp32 (data gradient): tap (ky, kx = 2) for the left edge, (ky, kx = 0) for the right edge
struct FsPrepJob { const float* w; void* wprep; float* wedge; int K, O, NG, w_so, w_sk, flip, kind; };
constexpr int FS_PREP_MAX = 24;
struct FsPrepBatch { FsPrepJob j[FS_PREP_MAX]; };

__global__ void __launch_bounds__(256) fs_prep_kernel(const __grid_constant__ FsPrepBatch batch) {
    const FsPrepJob& J = batch.j[blockIdx.y];
    const int KC = (J.kind == 1) ? 8 : 16, E = KC / 2, NG = J.NG;
    const int nchunks = J.K / KC, gy = J.O / (16 * NG), rows = 96 * NG;
    const int total = gy * nchunks * 3 * 2 * rows * E;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int i = idx;
        const int e = i % E; i /= E;
        const int row = i % rows; i /= rows;
        const int q = i & 1; i >>= 1;
        const int kx = i % 3; i /= 3;
        const int c = i % nchunks;
        const int y = i / nchunks;
        const int lo = row >= 48 * NG;
        const int rr = lo ? row - 48 * NG : row;
        const int ky = rr / (16 * NG), g = (rr / 16) % NG, o = (y * NG + g) * 16 + (rr % 16), k = c * KC + q * E + e;
        const int tap = ky * 3 + kx;
        const float v = __ldg(J.w + (size_t)o * J.w_so + (size_t)k * J.w_sk + (J.flip ? 8 - tap : tap));
        if (J.kind == 0) {
            unsigned short h, l;
            bf16_split(v, h, l);
            static_cast<unsigned short*>(J.wprep)[idx] = lo ? l : h;
        } else if (J.kind == 2) {
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn((v - __half2float(h)) * FS_LO_SCALE);
            static_cast<__half*>(J.wprep)[idx] = lo ? l : h;
        } else {
            const float hi = tf32_hi(v);
            static_cast<float*>(J.wprep)[idx] = lo ? v - hi : hi;
        }
    }
    if (J.wedge) {
        const int te = gy * 2 * J.K * 48 * NG;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < te; idx += gridDim.x * blockDim.x) {
            int i = idx;
            const int n = i % (48 * NG); i /= 48 * NG;
            const int k = i % J.K; i /= J.K;
            const int e = i & 1;
            const int y = i >> 1;
            const int ky = n / (16 * NG), g = (n / 16) % NG, o = (y * NG + g) * 16 + (n % 16);
            const int tap = ky * 3 + (e == 0 ? 2 : 0);
            J.wedge[idx] = __ldg(J.w + (size_t)o * J.w_so + (size_t)k * J.w_sk + (J.flip ? 8 - tap : tap));
        }
    }
}

}  // namespace

namespace sifnn {

bool conv3x3_fs_supported(int K, int O, int H, int W) { return fs_enabled() && fs_shape_ok(K, O, H, W); }
bool conv3x3_fs_fwd_preferred(int K, int O, int H, int W) { return fs_enabled() && (fs_shape_ok(K, O, H, W) || fs_fwd_narrow_wins(K, O, H, W)); }
// bytes of the edge-weight table of a data-gradient launch (K = dy channels, O = dx channels)
size_t conv3x3_fs_wedge_bytes(int K, int O) { return (size_t)2 * K * 3 * O * sizeof(float); }

int fs_prep(const float* const* w, void* const* wprep, float* const* wedge, const int* K, const int* O, const int* w_so, const int* w_sk, const int* flip, int n,
            cudaStream_t st) {
    for (int i0 = 0; i0 < n; i0 += FS_PREP_MAX) {
        FsPrepBatch b{};
        const int m = n - i0 < FS_PREP_MAX ? n - i0 : FS_PREP_MAX;
        for (int i = 0; i < m; ++i) {
            const int kind = tc_split_kind(flip[i0 + i]);   // flip == 1: data-gradient layout
            b.j[i] = FsPrepJob{w[i0 + i], wprep[i0 + i], wedge ? wedge[i0 + i] : nullptr, K[i0 + i], O[i0 + i], fs_groups(K[i0 + i], O[i0 + i], kind), w_so[i0 + i], w_sk[i0 + i],
                               flip[i0 + i], kind};
        }
        fs_prep_kernel<<<dim3(32, m), 256, 0, st>>>(b);
        SIFNN_TRY(check_launch("fs_prep_kernel"));
    }
    return 0;
}

int conv3x3_fwd_fs_prepped(const float* in, const float* in2, int K1, const float* in_scale, const float* in_shift, const void* wprep, float* out, double* stats,
                           int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st, const BnTail* tail) {
    return run_fs(tail, 0, in, in2, K1, in_scale, in_shift, wprep, nullptr, out, stats, accumulate, B, Cin, Cout, H, W, st);
}
int conv3x3_dgrad_fs_prepped(const float* dy, const void* wprep, const float* wedge, float* dx, int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
    return run_fs(nullptr, 1, dy, nullptr, 0, nullptr, nullptr, wprep, wedge, dx, nullptr, accumulate, B, Cout, Cin, H, W, st);
}

}  // namespace sifnn

extern "C" void sifnn_conv3x3_fs_config(int kind, int max_ctas) { sifnn::tc_split_set(kind, kind == 2 ? 0 : kind); g_fs_max_ctas = max_ctas; }
extern "C" void sifnn_conv3x3_fs_trace(void* buf) { g_fs_trace = static_cast<unsigned long long*>(buf); }
extern "C" void sifnn_conv3x3_fs_narrow(int on) { g_fs_narrow = on ? 1 : 0; }
extern "C" int sifnn_conv3x3_fs_supported(int Cin, int Cout, int H, int W) { return sifnn::conv3x3_fs_supported(Cin, Cout, H, W) ? 1 : 0; }

extern "C" int sifnn_conv3x3_fwd_fs(const float* in, const float* in_scale, const float* in_shift, const float* w, float* out, double* stats, void* wprep,
                                    int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(in && w && out && wprep, "conv3x3_fwd_fs: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_fwd_fs: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE((fs_shape_ok(Cin, Cout, H, W) || fs_fwd_narrow_wins(Cin, Cout, H, W)) && B > 0 && B <= 65535, "conv3x3_fwd_fs: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    const int K = Cin, O = Cout, so = Cin * 9, sk = 9, flip = 0;
    SIFNN_TRY(sifnn::fs_prep(&w, &wprep, nullptr, &K, &O, &so, &sk, &flip, 1, st));
    return sifnn::conv3x3_fwd_fs_prepped(in, nullptr, 0, in_scale, in_shift, wprep, out, stats, 0, B, Cin, Cout, H, W, st);
}

// Complete data gradient (zero-padded transposed convolution + the adjoint of the replicate padding) in one launch.
// wprep: sifnn_conv3x3_tc_wprep_bytes(Cout, Cin) + 2 * Cout * 3 * Cin * 4 bytes (split weights, then the fp32 edge-tap table).
extern "C" int sifnn_conv3x3_dgrad_fs(const float* dy, const float* w, float* dx, int accumulate, void* wprep, int B, int Cin, int Cout, int H, int W,
                                      sifnn_stream_t stream) {
    SIFNN_REQUIRE(dy && w && dx && wprep, "conv3x3_dgrad_fs: null pointer");
    SIFNN_REQUIRE(fs_shape_ok(Cout, Cin, H, W) && B > 0 && B <= 65535, "conv3x3_dgrad_fs: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    const int K = Cout, O = Cin, so = 9, sk = Cin * 9, flip = 1;
    float* wedge = reinterpret_cast<float*>(static_cast<char*>(wprep) + sifnn_conv3x3_tc_wprep_bytes(Cout, Cin));
    SIFNN_TRY(sifnn::fs_prep(&w, &wprep, &wedge, &K, &O, &so, &sk, &flip, 1, st));
    return sifnn::conv3x3_dgrad_fs_prepped(dy, wprep, wedge, dx, accumulate, B, Cin, Cout, H, W, st);
}
