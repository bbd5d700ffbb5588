// 3x3 convolution with ALL NINE TAPS folded into the MMA's N dimension ("full fold", round 2).
//
// Why: tools/mma_probe2.cu (profiles/r2a_mma_probe2.log) shows that on this GPU an M = 128 tcgen05.mma costs ~111 clocks for ANY
// N <= 144, whatever the operand layout (no swizzle / 128-byte swizzle / A in TMEM), the kind (tf32 K = 8, bf16 K = 16) or the accumulator
// pattern.  The round-1 kernels issued N = 48 / 96 MMAs (one filter row per MMA) and were bound by exactly that floor.  Here one MMA
// produces, for 128 input pixels, the contribution of 16 (bf16) or 8 (tf32) input channels to all 9 taps of 16 output channels:
//
//     E[p][(ky, kx, o)] (+)= sum_c A[p][c] * Wt[(ky, kx, o)][c]                N = 144, a plain GEMM: no halo, no im2col, no geometry
//     out[y][x][o] = sum_{ky,kx} E[(cy(y + ky - 1), cx(x + kx - 1))][(ky, kx, o)]      formed by the epilogue (shift-and-add)
//
// * The geometry lives only in the epilogue.  A CTA walks down image rows; for input row r the epilogue reduces kx by +-1 lane shifts
//   (warp shuffles; quadrant edges through shared memory) and keeps two partial output rows per thread in registers:
//   out[r-1] = P1 + G_r[ky=2] (complete -> stored), P1 <- P0 + G_r[ky=1], P0 <- G_r[ky=0].
// * Padding is a rule on the edge terms, for free: forward (replicate) -- the missing neighbour term is replaced by the pixel's own
//   term of the same tap; data gradient (adjoint of replicate padding over a zero-padded transposed convolution) -- by the pixel's own
//   term of the OPPOSITE tap.  No border pass, no extra MMAs, for rows, columns and corners alike.
// * Because E[p] depends on pixel p alone, the 128 MMA rows can be ANY 128 pixels: one 128-pixel piece of a 256-wide row, a 128-wide
//   row, two 64-wide rows or four 32-wide rows taken from different row ranges ("segments") that walk in lockstep.
// * fp32 parity through a 3-term split folded into K: (a_lo, w_hi), (a_hi, w_lo), (a_hi, w_hi) accumulate into the same TMEM columns.
//   The accumulation chain of one column is 3 * C_in / KC MMAs (the 9 taps are summed by the epilogue in fp32), shorter than round 1's.
// * Work split: the B * H image rows are cut into gridDim.x * G equal contiguous ranges (one halo row at each cut), so every SM gets
//   the same number of rows; more than 32 output channels are split over blockIdx.y.
//
// Warp roles (22 warps): 0..15 epilogue (TMEM lane quadrant = warp % 4, channel quarter = warp / 4: the epilogue is a long chain of
// dependent instructions per warp, so it is spread over four warps per scheduler), 16 TMA loader, 17 MMA issuer, 18..21 transformers
// (raw fp32 rows -> BatchNorm affine + ReLU -> hi/lo split -> K-major operand tiles).
#include "tc_common.cuh"

#include <cstdlib>
#include <cuda_fp16.h>

namespace {

using namespace sifnn_tc;

constexpr int FF_EPI_WARPS = 16, FF_LOAD_WARP = 16, FF_MMA_WARP = 17, FF_XF_WARP0 = 18, FF_XF_WARPS = 4;
constexpr int FF_THREADS = (FF_XF_WARP0 + FF_XF_WARPS) * 32;   // 704
constexpr int FF_CS = 4;             // output channels per epilogue thread
constexpr int FF_N = 144;            // MMA N = TMEM columns of one accumulator slot: (ky, kx, o) = 9 x 16
constexpr int FF_NSLOT = 3;
constexpr int FF_MAX_STRIP = 64;     // output rows per strip when a row has two tiles (bounds the carry buffers)
constexpr int FF_MAX_K = 128;        // input channels per launch (more than 64: one output group per CTA and 16-bit splits only, see run_ff)
constexpr int FF_W_TILE = 2 * FF_N * 16;   // bytes of one (hi or lo) weight tile of a chunk: [2 q][144 rows][16 B]
constexpr int FF_A_TILE = 2 * 128 * 16;    // bytes of one (hi or lo) activation tile of a chunk: [2 q][128 pixels][16 B]

#define FF_STAMP(ev, idx) do { if (DBG && a.trace && blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 256) a.trace[(ev) * 256 + (idx)] = clock64(); } while (0)

struct FfArgs {
    sifnn::BnTail tail;   // optional fused BatchNorm finalize (forward with statistics)
    const float* in_scale;
    const float* in_shift;
    const unsigned char* wprep;   // [O / 16][chunk][hi, lo][2 q][144][16 B]
    float* out;
    double* stats;
    int B, K, O, H, W;
    int accumulate;
    int G;        // image-row segments per 128-pixel tile: 128 / W for W < 128, else 1
    int nrows;    // B * H
    int K1;       // channels that come from the first tensor map (== K without a second source)
    unsigned long long* trace;   // debug (sifnn_conv3x3_ff_trace): clock64 stamps of CTA (0,0), [event][step], 256 steps per event
    int epi_spin; // epilogue waits for an accumulator by polling (1) or suspended (0)
    int ablate;   // debug (sifnn_conv3x3_ff_debug): 1 no MMAs, 2 no epilogue math, 4 no TMEM loads, 8 no transform, 16 no global stores, 32 loads hit L2
};

// Shared-memory carve-up, identical on host and device
struct FfLayout {
    int xs, xbuf, stat, carry_l, carry_out, w, a, raw, total;
    int AS, RS, raw_stage;
};
__host__ __device__ inline FfLayout ff_layout(int nchunks, int NG, int KC, bool T2) {
    FfLayout L{};
    int off = 1024;               // [0, 1024): mbarriers + TMEM slot
    off += 2 * FF_MAX_K * 4;      // BatchNorm scale / shift of the input channels
    L.xs = off; off += 4 * NG * 2 * FF_CS * 4;                    // rolling state of the deferred tile-edge pixel (two-tile rows only)
    off = (off + 127) & ~127;
    L.xbuf = off; off += 4 * 2 * 4 * 2 * 3 * FF_CS * 4;       // quadrant-edge exchange [quarter][parity][quad][kind][3 ky][4]
    L.stat = off; off += 4 * 2 * 16 * NG * 4;                    // per-quadrant partial BatchNorm sums of the CTA
    L.carry_l = off; if (T2) off += (FF_MAX_STRIP + 2) * 3 * 16 * NG * 4;
    L.carry_out = off; if (T2) off += FF_MAX_STRIP * 16 * NG * 4;
    off = (off + 1023) & ~1023;
    L.w = off; off += nchunks * NG * 2 * FF_W_TILE;
    L.AS = (NG == 1) ? (2 * nchunks < 8 ? (2 * nchunks < 4 ? 4 : 2 * nchunks) : 8) : 2 * nchunks;
    L.a = off; off += L.AS * 2 * FF_A_TILE;
    L.raw_stage = KC * 128 * 4;
    L.RS = 4;
    while (L.RS > 2 && off + L.RS * L.raw_stage > 220 * 1024) --L.RS;
    L.raw = off; off += L.RS * L.raw_stage;
    L.total = off;
    return L;
}

// One segment's walk over its share of the image rows: strips (one image, output rows [ya, yb)) x tiles of a row x input rows.
struct FfIter {
    int H, T, g1, b, ya, yb, t, r, rfirst, rlast;
    bool active;
    __device__ void begin(int g0) {
        if (g0 >= g1) { active = false; return; }
        b = g0 / H;
        ya = g0 - b * H;
        yb = min(g1 - b * H, H);
        if (T > 1 && yb - ya > FF_MAX_STRIP) yb = ya + FF_MAX_STRIP;
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
    __device__ int count() const {   // steps from the current position (call right after init)
        FfIter c = *this;
        int n = 0;
        while (c.active) { n += c.T * (c.rlast - c.rfirst + 1); c.begin(c.b * c.H + c.yb); }
        return n;
    }
};

__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t h, float& lo_elem, float& hi_elem) {
    asm("{\n\t.reg .f16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(lo_elem), "=f"(hi_elem) : "r"(h));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

template <int KIND, int NG, int PAD, bool AFFINE, bool STATS, bool T2, bool DBG>
__global__ void __launch_bounds__(FF_THREADS, 1) conv3x3_ff_kernel(const FfArgs a, const __grid_constant__ CUtensorMap tmap,
                                                                   const __grid_constant__ CUtensorMap tmap2) {
    sifnn::pdl_wait_and_trigger();   // launched with launch_pdl: every global access below comes after the previous kernel of the stream
    constexpr int KC = (KIND == 1) ? 8 : 16;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nchunks = a.K / KC;
    const FfLayout L = ff_layout(nchunks, NG, KC, T2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* raw_full = bars;            // [4]
    uint64_t* raw_empty = bars + 4;       // [4]
    uint64_t* a_full = bars + 8;          // [16]
    uint64_t* a_empty = bars + 24;        // [16]
    uint64_t* acc_full = bars + 40;       // [3]
    uint64_t* acc_empty = bars + 43;      // [3]
    uint64_t* w_full = bars + 46;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);
    float* sc_s = reinterpret_cast<float*>(smem + 1024);
    float* sh_s = sc_s + FF_MAX_K;
    float* xs = reinterpret_cast<float*>(smem + L.xs);
    float* xbuf = reinterpret_cast<float*>(smem + L.xbuf);
    float* carry_l = reinterpret_cast<float*>(smem + L.carry_l);
    float* carry_out = reinterpret_cast<float*>(smem + L.carry_out);
    unsigned char* w_s = smem + L.w;
    unsigned char* a_s = smem + L.a;
    unsigned char* raw_s = smem + L.raw;
    const int AS = L.AS, RS = L.RS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, G = a.G;
    const int T = T2 ? 2 : 1;
    const int Wseg = W < 128 ? W : 128;
    const int nworkers = gridDim.x * G;

    if (tid == 0) {
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, FF_XF_WARPS); }
        for (int s = 0; s < AS; ++s) { mbar_init(a_full + s, FF_XF_WARPS); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < FF_NSLOT; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, FF_EPI_WARPS); }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == FF_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (AFFINE) {
        for (int i = tid; i < a.K; i += FF_THREADS) { sc_s[i] = a.in_scale ? __ldg(a.in_scale + i) : 1.f; sh_s[i] = a.in_shift ? __ldg(a.in_shift + i) : 0.f; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t w_bytes = (uint32_t)(nchunks * NG * 2 * FF_W_TILE);
    if (tid == 0) {   // the split weights of this CTA's output channels stay resident for the whole kernel
        mbar_arrive_expect_tx(w_full, w_bytes);
        bulk_g2s(w_s, a.wprep + (size_t)blockIdx.y * w_bytes, w_bytes, w_full);
    }

    // number of lockstep steps of this CTA = the longest walk among its segments
    int nsteps = 0;
    for (int g = 0; g < G; ++g) {
        FfIter it;
        it.init(blockIdx.x * G + g, nworkers, a.nrows, H, T);
        nsteps = max(nsteps, it.count());
    }

    if (warp == FF_LOAD_WARP) {
        // ======================= loader: G TMA boxes {Wseg pixels, 1 row, KC planes} per chunk =======================
        FfIter it;
        it.init(blockIdx.x * G + (lane < G ? lane : 0), nworkers, a.nrows, H, T);
        int gc = 0;
        for (int step = 0; step < nsteps; ++step) {
            const bool ld_real = it.active && !(DBG && (a.ablate & 32));
            const int bb = ld_real ? it.b : 0, rr = ld_real ? it.r : 0, x0 = ld_real ? it.t * 128 : 0;
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int rs = gc % RS;
                if (lane == 0 && c == 0) FF_STAMP(11, step);
                if (gc >= RS) mbar_wait(raw_empty + rs, ((gc / RS) - 1) & 1);
                if (lane == 0) { if (c == 0) FF_STAMP(0, step); mbar_arrive_expect_tx(raw_full + rs, (uint32_t)L.raw_stage); }
                __syncwarp();
                if (lane < G) {
                    const int ch = c * KC;
                    unsigned char* dst = raw_s + (size_t)rs * L.raw_stage + (size_t)lane * (KC * Wseg * 4);
                    if (ch < a.K1) tma_load_3d(dst, &tmap, x0, rr, bb * a.K1 + ch, raw_full + rs);
                    else tma_load_3d(dst, &tmap2, x0, rr, bb * (a.K - a.K1) + (ch - a.K1), raw_full + rs);
                }
                __syncwarp();
                if (lane == 0 && c == 0) FF_STAMP(12, step);
            }
            it.next();
        }
    } else if (warp == FF_MMA_WARP) {
        // ======================= MMA issuer (one thread): 3 MMAs of N = 144 per (chunk, output group) =======================
        constexpr uint32_t idesc = KIND == 0 ? make_idesc_bf16(128, FF_N) : (KIND == 1 ? make_idesc(128, FF_N) : make_idesc_f16(128, FF_N));
        mbar_wait(w_full, 0);
        int seq = 0;
        for (int step = 0; step < nsteps; ++step) {
#pragma unroll
            for (int g = 0; g < NG; ++g, ++seq) {
                const int slot = seq % FF_NSLOT;
                if (seq >= FF_NSLOT) mbar_wait(acc_empty + slot, ((seq / FF_NSLOT) - 1) & 1);
                tc_fence_after();
                if (lane == 0 && g == 0) FF_STAMP(4, step);
                const uint32_t d = tmem_base + slot * FF_N;
                for (int c = 0; c < nchunks; ++c) {
                    const int gc = step * nchunks + c, as = gc % AS;
                    if (g == 0) { mbar_wait(a_full + as, (gc / AS) & 1); tc_fence_after(); if (lane == 0 && c == 0) FF_STAMP(5, step); }
                    if (lane == 0) {
                        const uint32_t a_hi = smem_u32(a_s + (size_t)as * 2 * FF_A_TILE);
                        const uint32_t w_hi = smem_u32(w_s + (size_t)(g * nchunks + c) * 2 * FF_W_TILE);
                        const uint64_t da_hi = make_desc(a_hi, 128 * 16, 128), da_lo = make_desc(a_hi + FF_A_TILE, 128 * 16, 128);
                        const uint64_t dw_hi = make_desc(w_hi, FF_N * 16, 128), dw_lo = make_desc(w_hi + FF_W_TILE, FF_N * 16, 128);
                        if (DBG && (a.ablate & 1)) {
                        } else if (KIND != 1) {
                            umma_bf16(d, da_lo, dw_hi, idesc, c > 0 ? 1u : 0u);   // small terms first
                            umma_bf16(d, da_hi, dw_lo, idesc, 1u);
                            umma_bf16(d, da_hi, dw_hi, idesc, 1u);
                        } else {
                            umma_tf32(d, da_lo, dw_hi, idesc, c > 0 ? 1u : 0u);
                            umma_tf32(d, da_hi, dw_lo, idesc, 1u);
                            umma_tf32(d, da_hi, dw_hi, idesc, 1u);
                        }
                        if (g == NG - 1) umma_commit(a_empty + as);   // chunk slot reusable once these MMAs have read it
                    }
                    __syncwarp();
                }
                if (lane == 0) { umma_commit(acc_full + slot); if (g == NG - 1) FF_STAMP(6, step); }
                __syncwarp();
            }
        }
    } else if (warp >= FF_XF_WARP0) {
        // ======================= transformers: raw fp32 -> (BatchNorm, ReLU) -> hi / lo operand tiles, thread = pixel =======================
        const int p = tid - FF_XF_WARP0 * 32;          // MMA row
        const int seg = p / Wseg, xt = p - seg * Wseg;
        int gc = 0;
        for (int step = 0; step < nsteps; ++step) {
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int rs = gc % RS, as = gc % AS;
                if (gc >= AS) mbar_wait(a_empty + as, ((gc / AS) - 1) & 1);
                if (p == 0 && c == 0) FF_STAMP(1, step);
                mbar_wait(raw_full + rs, (gc / RS) & 1);
                if (p == 0 && c == 0) FF_STAMP(2, step);
                const float* raw = reinterpret_cast<const float*>(raw_s + (size_t)rs * L.raw_stage) + (size_t)seg * KC * Wseg + xt;
                unsigned char* a_hi = a_s + (size_t)as * 2 * FF_A_TILE;
                const int c0 = c * KC;
                if (!(DBG && (a.ablate & 8)))
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if constexpr (KIND != 1) {
                        uint32_t hp[4], lp[4];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            float t0 = raw[(8 * q + e) * Wseg], t1 = raw[(8 * q + e + 1) * Wseg];
                            if (AFFINE) {
                                t0 = sifnn::act_affine_relu(t0, sc_s[c0 + 8 * q + e], sh_s[c0 + 8 * q + e]);
                                t1 = sifnn::act_affine_relu(t1, sc_s[c0 + 8 * q + e + 1], sh_s[c0 + 8 * q + e + 1]);
                            }
                            if constexpr (KIND == 0) {
                                const uint32_t h = pack_bf16x2(t0, t1);
                                const float r0 = t0 - __uint_as_float(h << 16), r1 = t1 - __uint_as_float(h & 0xffff0000u);
                                hp[e >> 1] = h;
                                lp[e >> 1] = pack_bf16x2(r0, r1);
                            } else {   // FP16: 11 + 11 significant bits; the three products share one accumulator here, so the residual is NOT scaled
                                const uint32_t h = pack_f16x2(t0, t1);
                                float f0, f1;
                                unpack_f16x2(h, f0, f1);
                                hp[e >> 1] = h;
                                lp[e >> 1] = pack_f16x2(t0 - f0, t1 - f1);
                            }
                        }
                        *reinterpret_cast<uint4*>(a_hi + (size_t)(q * 128 + p) * 16) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        *reinterpret_cast<uint4*>(a_hi + FF_A_TILE + (size_t)(q * 128 + p) * 16) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                    } else {
                        float hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float t = raw[(4 * q + e) * Wseg];
                            if (AFFINE) t = sifnn::act_affine_relu(t, sc_s[c0 + 4 * q + e], sh_s[c0 + 4 * q + e]);
                            hi[e] = tf32_hi(t);
                            lo[e] = t - hi[e];
                        }
                        *reinterpret_cast<float4*>(a_hi + (size_t)(q * 128 + p) * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4*>(a_hi + FF_A_TILE + (size_t)(q * 128 + p) * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                fence_proxy_async();           // this thread's st.shared -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) { mbar_arrive(raw_empty + rs); mbar_arrive(a_full + as); }
                if (p == 0 && c == nchunks - 1) FF_STAMP(3, step);
            }
        }
    } else {
        // ======================= epilogue: TMEM -> kx shift-add -> rolling ky sums -> global (+ BatchNorm statistics) =======================
        constexpr int CS = FF_CS;
        const int quad = warp & 3, cq = warp >> 2;           // TMEM lane quadrant, channel quarter [4 cq, 4 cq + 4) of every 16-channel group
        const int p = quad * 32 + lane;
        const int seg = p / Wseg, xt = p - seg * Wseg;
        const bool seg_first = (xt == 0), seg_last = (xt == Wseg - 1);   // only ever true on lane 0 / lane 31
        FfIter it;
        it.init(blockIdx.x * G + seg, nworkers, a.nrows, H, T);
        float P0[NG][CS], P1[NG][CS];
        float s1[STATS ? NG : 1][CS], s2[STATS ? NG : 1][CS];
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < CS; ++j) { P0[g][j] = 0.f; P1[g][j] = 0.f; if constexpr (STATS) { s1[g][j] = 0.f; s2[g][j] = 0.f; } }
        const bool accum = a.accumulate != 0;
        const size_t plane = (size_t)H * W;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + cq * CS;
        int seq = 0;
        for (int step = 0; step < nsteps; ++step) {
            const bool act = it.active;
            const int b = it.b, r = it.r, t = it.t, ya = it.ya, yb = it.yb;
            const int x = t * 128 + xt;
            const bool img_l = (x == 0), img_r = (x == W - 1);
            // multipliers of the lane-shifted neighbour terms: the edge lanes of a warp get theirs after the exchange, except at the image
            // edge of the replicate form, where the shuffle's "own value" is exactly the padding rule
            const float mL = (lane == 0 && !(PAD == 0 && img_l)) ? 0.f : 1.f;
            const float mR = (lane == 31 && !(PAD == 0 && img_r)) ? 0.f : 1.f;
            const bool defer_r = T2 && seg_last && !img_r;     // right neighbour lives in the next tile pass
            const bool carry_in = T2 && seg_first && !img_l;   // left neighbour was published by the previous tile pass
            const bool emit_prev = act && (r - 1 >= ya);
            const bool emit_last = act && (r == H - 1) && (yb == H);
#pragma unroll
            for (int g = 0; g < NG; ++g, ++seq) {
                const int slot = seq % FF_NSLOT;
                if (a.epi_spin) mbar_wait_spin(acc_full + slot, (seq / FF_NSLOT) & 1);
                else mbar_wait(acc_full + slot, (seq / FF_NSLOT) & 1);
                tc_fence_after();
                if (tid == 0 && g == 0) FF_STAMP(7, step);
                const uint32_t tcol = tlane + slot * FF_N;
                float* xb = xbuf + (size_t)(((cq * 2 + (seq & 1)) * 4 + quad) * 2) * (3 * CS);   // [kind][ky][4]
                const int chan = g * 16 + cq * CS;                                                // channel offset inside this CTA's 16 NG channels
                float Gk[3][CS];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    float e0[CS], e1[CS], e2[CS];
                    if (DBG && (a.ablate & 4)) {
#pragma unroll
                        for (int j = 0; j < CS; ++j) { e0[j] = 0.f; e1[j] = 1.f; e2[j] = 2.f; }
                    } else {
                        tmem_ld4(tcol + (ky * 3 + 0) * 16, e0);
                        tmem_ld4(tcol + (ky * 3 + 1) * 16, e1);
                        tmem_ld4(tcol + (ky * 3 + 2) * 16, e2);
                        tmem_ld_wait();
                    }
                    if (DBG && (a.ablate & 2)) {
#pragma unroll
                        for (int j = 0; j < CS; ++j) Gk[ky][j] = e0[j] + e1[j] + e2[j];
                        continue;
                    }
                    if (lane == 31) {   // my kx = 0 terms are the right neighbour's left terms
                        *reinterpret_cast<float4*>(xb + ky * CS) = make_float4(e0[0], e0[1], e0[2], e0[3]);
                        if (T2 && defer_r) *reinterpret_cast<float4*>(carry_l + ((size_t)(r - it.rfirst) * 3 + ky) * (16 * NG) + chan) = make_float4(e0[0], e0[1], e0[2], e0[3]);
                    } else if (lane == 0) {   // my kx = 2 terms are the left neighbour's right terms
                        *reinterpret_cast<float4*>(xb + 3 * CS + ky * CS) = make_float4(e2[0], e2[1], e2[2], e2[3]);
                    }
#pragma unroll
                    for (int j = 0; j < CS; ++j) {
                        const float l = __shfl_up_sync(0xffffffffu, e0[j], 1);
                        const float rr = __shfl_down_sync(0xffffffffu, e2[j], 1);
                        Gk[ky][j] = fmaf(l, mL, fmaf(rr, mR, e1[j]));
                    }
                    if (PAD == 1 && (seg_first || seg_last)) {   // adjoint of the replicate padding: the term that would leave the image comes back
                        if (img_l) {
#pragma unroll
                            for (int j = 0; j < CS; ++j) Gk[ky][j] += e2[j];
                        } else if (img_r) {
#pragma unroll
                            for (int j = 0; j < CS; ++j) Gk[ky][j] += e0[j];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + slot);   // the slot may be overwritten (one arrival per warp)
                if (tid == 0 && g == 0) FF_STAMP(8, step);
                if (DBG && (a.ablate & 2)) {
                    if (Gk[0][0] + Gk[1][1] + Gk[2][2] == 1234.5f) a.out[0] = 1.f;
                    continue;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + cq) : "memory");   // the four quadrant warps of this channel quarter
                if (tid == 0 && g == 0) FF_STAMP(9, step);
                // edge lanes: add the neighbour terms that live in another warp (or came from the previous tile pass)
                if (lane == 0 || lane == 31) {
                    const float* nb = nullptr;
                    int stride = CS;
                    if (lane == 0) {
                        if (carry_in) { nb = carry_l + (size_t)(r - it.rfirst) * 3 * (16 * NG) + chan; stride = 16 * NG; }
                        else if (!seg_first) nb = xb - 2 * 3 * CS;          // quadrant to the left, kind 0
                    } else {
                        if (!seg_last) nb = xb + 2 * 3 * CS + 3 * CS;       // quadrant to the right, kind 1
                    }
                    if (nb) {
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const float4 v0 = *reinterpret_cast<const float4*>(nb + ky * stride);
                            Gk[ky][0] += v0.x; Gk[ky][1] += v0.y; Gk[ky][2] += v0.z; Gk[ky][3] += v0.w;
                        }
                    }
                }
                // rolling sums over ky.  Top edge: the missing row above is replaced by this row's own ky = 0 (forward) / ky = 2 (data gradient) term.
                constexpr int ET = (PAD == 0) ? 0 : 2, EB = (PAD == 0) ? 2 : 0;
                if (r == 0) {
#pragma unroll
                    for (int j = 0; j < CS; ++j) P0[g][j] = Gk[ET][j];
                }
                float o_prev[CS];
#pragma unroll
                for (int j = 0; j < CS; ++j) {
                    o_prev[j] = P1[g][j] + Gk[2][j];
                    P1[g][j] = P0[g][j] + Gk[1][j];
                    P0[g][j] = Gk[0][j];
                }
                const int obase = blockIdx.y * NG * 16 + chan;
                auto emit = [&](int row, const float* v) {
                    if (T2 && defer_r) {   // this pixel still misses its right neighbour: park it for the next tile pass
                        *reinterpret_cast<float4*>(carry_out + (size_t)(row - ya) * (16 * NG) + chan) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
                        float* op = a.out + ((size_t)b * a.O + obase) * plane + (size_t)row * W + x;
                        float old[CS];
                        if (accum) {   // all loads first: one memory round trip instead of CS dependent ones
#pragma unroll
                            for (int j = 0; j < CS; ++j) old[j] = __ldcg(op + (size_t)j * plane);
                        }
#pragma unroll
                        for (int j = 0; j < CS; ++j, op += plane) {
                            float o = v[j];
                            if (accum) o += old[j];
                            if (!(DBG && (a.ablate & 16))) *op = o;
                            if constexpr (STATS) { s1[g][j] += o; s2[g][j] = fmaf(o, o, s2[g][j]); }
                        }
                    }
                };
                if (emit_prev) emit(r - 1, o_prev);
                if (emit_last) {
                    float o_last[CS];
#pragma unroll
                    for (int j = 0; j < CS; ++j) o_last[j] = P1[g][j] + Gk[EB][j];
                    emit(H - 1, o_last);
                }
                if (T2 && carry_in && lane == 0) {
                    // finish pixel x - 1 (the last pixel of the previous tile pass): its parked value + the rolled kx = 2 terms of this pixel
                    float* x0s = xs + (size_t)((cq * NG + g) * 2) * CS;
                    float* x1s = x0s + CS;
                    const float* own = xb + 3 * CS;   // my own kx = 2 terms, published above
                    float* op0 = a.out + ((size_t)b * a.O + obase) * plane + (size_t)(x - 1);
#pragma unroll
                    for (int j = 0; j < CS; ++j) {
                        const float X0 = own[j], X1 = own[CS + j], X2 = own[2 * CS + j];
                        const float q0 = (r == 0) ? (PAD == 0 ? X0 : X2) : x0s[j];
                        const float ox_prev = x1s[j] + X2;
                        const float q1 = q0 + X1;
                        x1s[j] = q1;
                        x0s[j] = X0;
                        if (emit_prev) {
                            float* op = op0 + (size_t)j * plane + (size_t)(r - 1) * W;
                            float o = carry_out[(size_t)(r - 1 - ya) * (16 * NG) + chan + j] + ox_prev;
                            if (accum) o += __ldcg(op);
                            *op = o;
                            if constexpr (STATS) { s1[g][j] += o; s2[g][j] = fmaf(o, o, s2[g][j]); }
                        }
                        if (emit_last) {
                            float* op = op0 + (size_t)j * plane + (size_t)(H - 1) * W;
                            float o = carry_out[(size_t)(H - 1 - ya) * (16 * NG) + chan + j] + q1 + (PAD == 0 ? X2 : X0);
                            if (accum) o += __ldcg(op);
                            *op = o;
                            if constexpr (STATS) { s1[g][j] += o; s2[g][j] = fmaf(o, o, s2[g][j]); }
                        }
                    }
                }
            }
            if (tid == 0) FF_STAMP(10, step);
            it.next();
        }
        if constexpr (STATS) if (a.stats) {
            // per-channel sums of this CTA: lanes by shuffle, the four quadrant warps through shared memory in a fixed order, one fp64 atomic per channel and CTA
            float* part = reinterpret_cast<float*>(smem + L.stat);   // [quad][stat][NG * 16]
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int j = 0; j < CS; ++j) {
                    const float t1 = sifnn::warp_sum(s1[g][j]), t2 = sifnn::warp_sum(s2[g][j]);
                    if (lane == 0) {
                        part[(quad * 2 + 0) * 16 * NG + g * 16 + cq * CS + j] = t1;
                        part[(quad * 2 + 1) * 16 * NG + g * 16 + cq * CS + j] = t2;
                    }
                }
            asm volatile("bar.sync 5, 512;" ::: "memory");   // the sixteen epilogue warps
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
    if (warp == FF_MMA_WARP) {
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

// Operand precision of the 3-term split: BF16 (K = 16 per MMA; per-layer error ~5e-6) by default, TF32 (K = 8: twice the MMAs, ~4e-7) with
// SIFNN_FF_TF32=1 or sifnn_conv3x3_ff_config(1, ...).  SIFNN_FF=0 turns the full-fold kernel off in the network plan (round-1 kernels; A/B runs).
int g_ff_max_ctas = 0, g_ff_ablate = 0;
unsigned long long* g_ff_trace = nullptr;
bool ff_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_FF"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <int KIND, int NG, int PAD, bool AFFINE, bool STATS, bool T2>
int launch_ff(const FfArgs& a, const CUtensorMap& tm1, const CUtensorMap& tm2, int gx, int gy, cudaStream_t st) {
    constexpr int KC = (KIND == 1) ? 8 : 16;
    const FfLayout L = ff_layout(a.K / KC, NG, KC, T2);
    auto kern = conv3x3_ff_kernel<KIND, NG, PAD, AFFINE, STATS, T2, false>;
    if (a.ablate || a.trace) {
        // the traced / ablatable build exists for the plain bf16 / tf32 forward only (tools/trace_ff.py, tools/ablate_ff.py)
        if constexpr (PAD == 0 && !AFFINE && !STATS) kern = conv3x3_ff_kernel<KIND, NG, PAD, AFFINE, STATS, T2, true>;
    }
    SIFNN_REQUIRE(L.total <= 227 * 1024, "conv3x3_ff: shared-memory budget exceeded (K=%d)", a.K);
    SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    SIFNN_CUDA(sifnn::launch_pdl(kern, dim3(gx, gy), dim3(FF_THREADS), (size_t)L.total, st, a, tm1, tm2));
    return sifnn::check_launch("conv3x3_ff_kernel");
}

template <int KIND, int PAD, bool AFFINE, bool STATS>
int dispatch_ff2(const FfArgs& a, const CUtensorMap& tm1, const CUtensorMap& tm2, int NG, int gx, int gy, bool t2, cudaStream_t st) {
    if (NG == 1) return t2 ? launch_ff<KIND, 1, PAD, AFFINE, STATS, true>(a, tm1, tm2, gx, gy, st) : launch_ff<KIND, 1, PAD, AFFINE, STATS, false>(a, tm1, tm2, gx, gy, st);
    return t2 ? launch_ff<KIND, 2, PAD, AFFINE, STATS, true>(a, tm1, tm2, gx, gy, st) : launch_ff<KIND, 2, PAD, AFFINE, STATS, false>(a, tm1, tm2, gx, gy, st);
}
template <int KIND>
int dispatch_ff1(int pad, bool affine, bool stats, const FfArgs& a, const CUtensorMap& tm1, const CUtensorMap& tm2, int NG, int gx, int gy, bool t2, cudaStream_t st) {
    if (pad == 1) return dispatch_ff2<KIND, 1, false, false>(a, tm1, tm2, NG, gx, gy, t2, st);
    if (affine) return stats ? dispatch_ff2<KIND, 0, true, true>(a, tm1, tm2, NG, gx, gy, t2, st) : dispatch_ff2<KIND, 0, true, false>(a, tm1, tm2, NG, gx, gy, t2, st);
    return stats ? dispatch_ff2<KIND, 0, false, true>(a, tm1, tm2, NG, gx, gy, t2, st) : dispatch_ff2<KIND, 0, false, false>(a, tm1, tm2, NG, gx, gy, t2, st);
}

bool ff_shape_ok(int K, int O, int H, int W) {
    // 65..128 input channels: the resident weights (8 chunks) only fit beside the operand ring with one output group per CTA and 16-byte-per-row
    // 16-bit tiles; the TF32 split (twice the chunks) stays at 64
    const bool any_tf32 = sifnn::tc_split_kind(0) == 1 || sifnn::tc_split_kind(1) == 1;
    const int maxk = any_tf32 ? 64 : FF_MAX_K;
    return (W == 32 || W == 64 || W == 128 || W == 256) && (K % 16 == 0) && K >= 16 && K <= maxk && (O % 16 == 0) && O >= 16 && O <= 128 && H >= 1;
}

// in2 != nullptr: channels [K1, K) come from a second tensor (the two halves of a channel concat read in place)
int run_ff(const sifnn::BnTail* tail, int pad, const float* in, const float* in2, int K1, const float* in_scale, const float* in_shift, const void* wprep, float* out, double* stats,
           int accumulate, int B, int K, int O, int H, int W, cudaStream_t st) {
    SIFNN_REQUIRE(ff_shape_ok(K, O, H, W), "conv3x3_ff: unsupported shape K=%d O=%d H=%d W=%d", K, O, H, W);
    const int kind = sifnn::tc_split_kind(pad);
    SIFNN_REQUIRE(!(pad == 1 && kind == 2), "conv3x3_ff: the FP16 split is for the forward form only (gradients can be tiny)");
    const int KC = (kind == 1) ? 8 : 16;
    SIFNN_REQUIRE(!in2 || (K1 % KC == 0 && K1 > 0 && K1 < K), "conv3x3_ff: the split point of a two-source input must be a multiple of %d", KC);
    FfArgs a{};
    a.in_scale = in_scale; a.in_shift = in_shift; a.wprep = static_cast<const unsigned char*>(wprep); a.out = out; a.stats = stats;
    if (tail && stats) a.tail = *tail;
    { static int spin = -1; if (spin < 0) { const char* e = getenv("SIFNN_FF_EPI_SPIN"); spin = e ? atoi(e) : 0; } a.epi_spin = spin; }   // default suspended: on the 64- / 32-pixel levels the epilogue waits for the MMAs, polling only took issue slots (4.221 -> 4.215 ms)
    a.B = B; a.K = K; a.O = O; a.H = H; a.W = W; a.accumulate = accumulate ? 1 : 0;
    a.G = W < 128 ? 128 / W : 1;
    a.nrows = B * H;
    a.K1 = in2 ? K1 : K;
    a.ablate = g_ff_ablate;
    a.trace = g_ff_trace;
    // output groups per CTA: two where the weights fit; TF32 with 64 input channels keeps one (8 chunks of weights + operand ring)
    int NG = (O >= 32) ? 2 : 1;
    if (kind == 1 && K > 32) NG = 1;
    if (K > 64) NG = 1;
    const int gy = O / (16 * NG);
    int gx = sifnn::num_sms() / gy;
    if (g_ff_max_ctas > 0 && gx > g_ff_max_ctas) gx = g_ff_max_ctas;   // tests: long strips on small inputs
    const int maxw = (a.nrows + a.G - 1) / a.G;
    if (gx > maxw) gx = maxw;
    if (gx < 1) gx = 1;
    const int Wseg = W < 128 ? W : 128;
    CUtensorMap tm1, tm2;
    SIFNN_REQUIRE(encode_planes_map(&tm1, in, W, H, (long long)B * a.K1, Wseg, 1, KC), "conv3x3_ff: cuTensorMapEncodeTiled is unavailable or failed");
    if (in2) SIFNN_REQUIRE(encode_planes_map(&tm2, in2, W, H, (long long)B * (K - K1), Wseg, 1, KC), "conv3x3_ff: cuTensorMapEncodeTiled failed (second source)");
    else tm2 = tm1;
    const bool t2 = (W == 256);
    const bool affine = in_scale != nullptr;
    if (kind == 0) return dispatch_ff1<0>(pad, affine, stats != nullptr, a, tm1, tm2, NG, gx, gy, t2, st);
    if (kind == 1) return dispatch_ff1<1>(pad, affine, stats != nullptr, a, tm1, tm2, NG, gx, gy, t2, st);
    return dispatch_ff1<2>(pad, affine, stats != nullptr, a, tm1, tm2, NG, gx, gy, t2, st);
}

// split weights in the exact shared-memory image of the kernel: [O / 16][chunk][hi, lo][2 q][144 = (tap, o % 16)][8 bf16 | 4 tf32]
struct FfPrepJob { const float* w; void* wprep; int K, O, w_so, w_sk, flip, kind; };
constexpr int FF_PREP_MAX = 24;
struct FfPrepBatch { FfPrepJob j[FF_PREP_MAX]; };

__global__ void __launch_bounds__(256) ff_prep_kernel(const __grid_constant__ FfPrepBatch batch) {
    const FfPrepJob& J = batch.j[blockIdx.y];
    const int KC = (J.kind == 1) ? 8 : 16, E = KC / 2;
    const int nchunks = J.K / KC;
    const int total = (J.O / 16) * nchunks * 2 * 2 * FF_N * E;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int i = idx;
        const int e = i % E; i /= E;
        const int n = i % FF_N; i /= FF_N;
        const int q = i & 1; i >>= 1;
        const int lo = i & 1; i >>= 1;
        const int c = i % nchunks;
        const int grp = i / nchunks;
        const int tap = n / 16, o = grp * 16 + (n % 16), k = c * KC + q * E + e;
        const float v = __ldg(J.w + (size_t)o * J.w_so + (size_t)k * J.w_sk + (J.flip ? 8 - tap : tap));
        if (J.kind == 0) {
            unsigned short h, l;
            bf16_split(v, h, l);
            static_cast<unsigned short*>(J.wprep)[idx] = lo ? l : h;
        } else if (J.kind == 2) {
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            static_cast<__half*>(J.wprep)[idx] = lo ? l : h;
        } else {
            const float hi = tf32_hi(v);
            static_cast<float*>(J.wprep)[idx] = lo ? v - hi : hi;
        }
    }
}

}  // namespace

namespace sifnn {

bool conv3x3_ff_supported(int K, int O, int H, int W) { return ff_enabled() && ff_shape_ok(K, O, H, W); }

// K0 / Kn: the slice of input channels (forward) or of dy channels (data gradient) this launch covers (weights of the other channels are skipped)
int ff_prep(const float* const* w, void* const* wprep, const int* K, const int* O, const int* w_so, const int* w_sk, const int* flip, int n, cudaStream_t st) {
    for (int i0 = 0; i0 < n; i0 += FF_PREP_MAX) {
        FfPrepBatch b{};
        const int m = n - i0 < FF_PREP_MAX ? n - i0 : FF_PREP_MAX;
        for (int i = 0; i < m; ++i) b.j[i] = FfPrepJob{w[i0 + i], wprep[i0 + i], K[i0 + i], O[i0 + i], w_so[i0 + i], w_sk[i0 + i], flip[i0 + i], tc_split_kind(flip[i0 + i])};
        ff_prep_kernel<<<dim3(32, m), 256, 0, st>>>(b);
        SIFNN_TRY(check_launch("ff_prep_kernel"));
    }
    return 0;
}

int conv3x3_fwd_ff_prepped(const float* in, const float* in2, int K1, const float* in_scale, const float* in_shift, const void* wprep, float* out, double* stats,
                           int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st, const BnTail* tail) {
    return run_ff(tail, 0, in, in2, K1, in_scale, in_shift, wprep, out, stats, accumulate, B, Cin, Cout, H, W, st);
}
int conv3x3_dgrad_ff_prepped(const float* dy, const void* wprep, float* dx, int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
    return run_ff(nullptr, 1, dy, nullptr, 0, nullptr, nullptr, wprep, dx, nullptr, accumulate, B, Cout, Cin, H, W, st);
}

}  // namespace sifnn

extern "C" void sifnn_conv3x3_ff_config(int kind, int max_ctas) { sifnn::tc_split_set(kind, kind == 2 ? 0 : kind); g_ff_max_ctas = max_ctas; }
extern "C" void sifnn_conv3x3_ff_debug(int ablate) { g_ff_ablate = ablate; }
extern "C" void sifnn_conv3x3_ff_trace(void* buf) { g_ff_trace = static_cast<unsigned long long*>(buf); }
extern "C" int sifnn_conv3x3_ff_supported(int Cin, int Cout, int H, int W) { return sifnn::conv3x3_ff_supported(Cin, Cout, H, W) ? 1 : 0; }

extern "C" int sifnn_conv3x3_fwd_ff(const float* in, const float* in_scale, const float* in_shift, const float* w, float* out, double* stats, void* wprep,
                                    int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(in && w && out && wprep, "conv3x3_fwd_ff: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_fwd_ff: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(ff_shape_ok(Cin, Cout, H, W) && B > 0 && B <= 65535, "conv3x3_fwd_ff: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    const int K = Cin, O = Cout, so = Cin * 9, sk = 9, flip = 0;
    SIFNN_TRY(sifnn::ff_prep(&w, &wprep, &K, &O, &so, &sk, &flip, 1, st));
    return sifnn::conv3x3_fwd_ff_prepped(in, nullptr, 0, in_scale, in_shift, wprep, out, stats, 0, B, Cin, Cout, H, W, st);
}

// Complete data gradient (zero-padded transposed convolution + the adjoint of the replicate padding) in one launch.
extern "C" int sifnn_conv3x3_dgrad_ff(const float* dy, const float* w, float* dx, int accumulate, void* wprep, int B, int Cin, int Cout, int H, int W,
                                      sifnn_stream_t stream) {
    SIFNN_REQUIRE(dy && w && dx && wprep, "conv3x3_dgrad_ff: null pointer");
    SIFNN_REQUIRE(ff_shape_ok(Cout, Cin, H, W) && B > 0 && B <= 65535, "conv3x3_dgrad_ff: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    const int K = Cout, O = Cin, so = 9, sk = Cin * 9, flip = 1;
    SIFNN_TRY(sifnn::ff_prep(&w, &wprep, &K, &O, &so, &sk, &flip, 1, st));
    return sifnn::conv3x3_dgrad_ff_prepped(dy, wprep, dx, accumulate, B, Cin, Cout, H, W, st);
}
