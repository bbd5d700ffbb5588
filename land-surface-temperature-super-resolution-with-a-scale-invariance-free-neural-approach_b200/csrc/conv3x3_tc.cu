// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM),
// fp32-accurate through a 3-term TF32 split.  Same contract as conv3x3.cu (forward with
// replicate padding + fused BatchNorm/ReLU prologue + BatchNorm-statistics epilogue; data
// gradient with zero padding and flipped/transposed weights), used for the layers whose
// width is a multiple of 128 pixels and whose channel counts are multiples of 8 / {16,32,64}.
//
//   D[M = 128 pixels of one image row][N = C_out] += A_tap[M][K = 8 channels] * B_tap[N][K]   for 9 taps, K/8 chunks
//
// * A (activations) is staged ONCE per chunk as a pixel-major halo tile  T[q = ch/4][pixel][4 ch]
//   (16 bytes per pixel and channel quad).  In the K-major / no-swizzle canonical layout a row
//   of the MMA is exactly one 16-byte unit and rows are contiguous (SBO = 128 B), so the nine
//   taps are nine START ADDRESSES into the same tile -- no im2col, no shifted copies.
// * fp32 parity (rel 1e-4): a = a_hi + a_lo, w = w_hi + w_lo (TF32 each);
//   D = a_hi*w_hi + a_hi*w_lo + a_lo*w_hi.  The weight tile stacks [w_hi ; w_lo] along N, so one
//   MMA (N' = 2N) produces a_hi*w_hi and a_hi*w_lo in two TMEM column blocks and a second MMA
//   (N) adds a_lo*w_hi: two A fetches per tap instead of three.
// * Warp roles: warps 0..7 stage A (global -> BN/ReLU -> split -> st.shared), one elected thread
//   bulk-copies the pre-split weights of the chunk (cp.async.bulk, completes on the stage's
//   mbarrier); warp 8 issues tcgen05.mma from one thread and releases stages with
//   tcgen05.commit; warps 0..7 then read the accumulators back (tcgen05.ld) for the epilogue.
#include "tc_common.cuh"

#include <cstdlib>

#ifdef SIFNN_TC_TRACE
static unsigned long long* g_trace = nullptr;
extern "C" void sifnn_debug_set_trace(void* p) { g_trace = static_cast<unsigned long long*>(p); }
#define TC_STAMP(ev, idx) do { if (a.trace && blockIdx.x == 0 && (idx) < 64) a.trace[(ev) * 64 + (idx)] = clock64(); } while (0)
#else
#define TC_STAMP(ev, idx) do { } while (0)
#endif

namespace {

using namespace sifnn_tc;

constexpr int TC_KC = 8;           // channels per pipeline stage = one MMA K step (TF32: 32 bytes)

struct TcArgs {
    const float* in;
    const float* in_scale;
    const float* in_shift;
    const float* wprep;  // [K/8][9][2][2N][4] hi/lo-split weights (prep kernel below)
    const float* bias;
    float* out;
    double* stats;
    int B, K, O, H, W;
    int accumulate;
    int tiles_x, tiles_y, num_tiles;
    unsigned long long* trace;   // SIFNN_TC_TRACE builds only: clock64 stamps of CTA 0, [event][slot], 64 slots per event
    int ablate;   // SIFNN_TC_ABLATE builds only: bit 0 no MMAs, 1 no output stores, 2 no transform, 3 no accumulator drain work
};

// Warp roles of the persistent kernel (18 warps)
constexpr int TC_EPI_WARPS = 8;     // warps 0..7 : epilogue (TMEM lane quadrant = warp % 4, channel half = warp / 4)
constexpr int TC_LOAD_WARP = 8;     // warp 8     : bulk-copies raw activation rows (global -> shared, TMA unit)
constexpr int TC_MMA_WARP = 9;      // warp 9     : issues tcgen05.mma / tcgen05.commit from one thread
constexpr int TC_XF_WARP0 = 10;     // warps 10..17: transform raw rows -> (BatchNorm, ReLU) -> TF32 hi/lo pixel-major tiles
constexpr int TC_XF_THREADS = 256;
constexpr int TC_THREADS2 = (TC_XF_WARP0 + 8) * 32;

// N = output channels, R = output rows per tile, MM = pixels per MMA (the UMMA M: 128, or 64 for 64-pixel-wide images)
template <int N, int R, int MM>
struct TcSmem {
    static constexpr int TROWS = R + 2;
    static constexpr int PITCH = MM + 2;        // tile row: MM pixels + 2 halo columns
    static constexpr int RAW_ROW = MM + 8;      // staged raw row: gx = x0-4 .. x0+MM+3 (16-byte aligned both ends)
    static constexpr int A_TILE = 2 * TROWS * PITCH * 4;  // floats per (hi or lo) tile: [2 q][TROWS*PITCH px][4]
    static constexpr int B_TILE = 9 * 2 * 2 * N * 4;          // floats: [9 taps][2 q][2N rows][4]
    static constexpr int STAGE = 2 * A_TILE + B_TILE;
    static constexpr int RAW_STAGE = TC_KC * TROWS * RAW_ROW;  // floats
    static constexpr int RAW_STAGES = 3;
    static constexpr int CTRL_FLOATS = 512;  // barriers, TMEM slot, BatchNorm scale/shift (2 KB)
    static constexpr int BUDGET = 220 * 1024;
    static constexpr int REST = BUDGET - CTRL_FLOATS * 4 - RAW_STAGES * RAW_STAGE * 4;
    static constexpr int STAGES = (STAGE * 4 * 4 <= REST) ? 4 : ((STAGE * 4 * 3 <= REST) ? 3 : 2);
    static constexpr size_t BYTES = (size_t)(STAGES * STAGE + RAW_STAGES * RAW_STAGE + CTRL_FLOATS) * 4;
    static constexpr int ACC_COLS = R * 2 * N;  // TMEM columns of one accumulator set
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit the 512 TMEM columns");
    static_assert((STAGE * 4) % 128 == 0 && (RAW_STAGE * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(2 * N <= 256, "UMMA N limit");
};

// Persistent, warp-specialised implicit-GEMM convolution.  Four rings of mbarriers:
//   raw  : loader -> transformers   (raw fp32 rows, bulk-copied RAW_STAGES chunks ahead: hides DRAM/L2 latency)
//   ab   : transformers (+ weight bulk copy) -> MMA   (TF32 hi/lo A tile + [w_hi ; w_lo] B tile)
//   acc  : MMA -> epilogue   (two TMEM accumulator sets: the epilogue of tile i overlaps the MMAs of tile i+1)
template <int N, int R, int MM, int PAD, bool AFFINE>
__global__ void __launch_bounds__(TC_THREADS2, 1) conv3x3_tc_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tmap) {
    using SM = TcSmem<N, R, MM>;
    constexpr int TROWS = SM::TROWS;
    constexpr int TC_PITCH = SM::PITCH;
    constexpr int TC_RAW_ROW = SM::RAW_ROW;
    constexpr bool STATS = (PAD == 0);  // BatchNorm statistics exist only in the forward (replicate) form
    constexpr int S = SM::STAGES;
    constexpr int RS = SM::RAW_STAGES;
    extern __shared__ __align__(128) float smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* ab_full = bars;            // [4]
    uint64_t* ab_empty = bars + 4;       // [4]
    uint64_t* raw_full = bars + 8;       // [4]
    uint64_t* raw_empty = bars + 12;     // [4]
    uint64_t* acc_full = bars + 16;      // [2]
    uint64_t* acc_empty = bars + 18;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* sc_s = smem + 64;    // [<=128]
    float* sh_s = smem + 192;   // [<=128]
    float* stage0 = smem + SM::CTRL_FLOATS;
    float* raw0 = stage0 + (size_t)S * SM::STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, K = a.K;
    const size_t plane = (size_t)H * W;
    const int nchunks = K / TC_KC;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(ab_full + s, TC_XF_THREADS + 1); mbar_init(ab_empty + s, 1); }
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, TC_XF_THREADS); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, TC_EPI_WARPS * 32); }
        fence_mbar_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, SM::TMEM_COLS);
    if (AFFINE) {
        for (int i = tid; i < K; i += TC_THREADS2) { sc_s[i] = __ldg(a.in_scale + i); sh_s[i] = __ldg(a.in_shift + i); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int& b, int& y0, int& x0) {
        b = tile / tiles_per_img;
        const int t = tile - b * tiles_per_img;
        const int ty = t / a.tiles_x;
        y0 = ty * R;
        x0 = (t - ty * a.tiles_x) * MM;
    };

    if (warp == TC_LOAD_WARP) {
        // ======================= loader: one TMA box (8 ch x TROWS rows x 136 cols) per chunk, RAW_STAGES ahead =======================
        // Out-of-image elements are zero-filled by the TMA unit: exactly the zero padding of the data-gradient form;
        // for replicate padding the transformers re-index them onto the clamped row / column, which is inside the box.
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            for (int ch = 0; ch < nchunks; ++ch, ++g) {
                const int rs = g % RS;
                if (g >= RS) mbar_wait(raw_empty + rs, ((g / RS) - 1) & 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(raw_full + rs, SM::RAW_STAGE * 4);
                    tma_load_3d(raw0 + (size_t)rs * SM::RAW_STAGE, &tmap, x0 - 4, y0 - 1, b * K + ch * TC_KC, raw_full + rs);
                }
                __syncwarp();
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ======================= MMA issuer (one thread) =======================
        constexpr uint32_t idesc1 = make_idesc(MM, 2 * N);  // a_hi x [w_hi ; w_lo]
        constexpr uint32_t idesc2 = make_idesc(MM, N);      // a_lo x  w_hi
        int g = 0, it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int ab = it & 1;
            if (it >= 2) mbar_wait(acc_empty + ab, ((it >> 1) - 1) & 1);
            tc_fence_after();
            for (int ch = 0; ch < nchunks; ++ch, ++g) {
                const int s = g % S;
                mbar_wait(ab_full + s, (g / S) & 1);
                tc_fence_after();
                if (lane == 0) {
                    // The issuing thread is the critical path of the whole CTA (72 MMAs per chunk from ONE thread), so the
                    // descriptors are built once per stage and advanced by compile-time constants: the start-address field
                    // is the low 14 bits in 16-byte units, and every offset below stays inside the 256 KB window.
                    constexpr uint32_t LBO_A = TROWS * TC_PITCH * 16, LBO_B = 2 * N * 16, SBO = 128;
                    const uint32_t a_hi = smem_u32(stage0 + (size_t)s * SM::STAGE);
                    const uint64_t da_hi = make_desc(a_hi, LBO_A, SBO);
                    const uint64_t da_lo = da_hi + (uint64_t)(SM::A_TILE * 4 / 16);
                    const uint64_t db = da_hi + (uint64_t)(2 * SM::A_TILE * 4 / 16) - ((uint64_t)(LBO_A >> 4) << 16) + ((uint64_t)(LBO_B >> 4) << 16);
                    const uint32_t first = (ch != 0) ? 1u : 0u;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint32_t d = tmem_base + ab * SM::ACC_COLS + r * 2 * N;
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const int ky = t / 3, kx = t - 3 * ky;
                            const uint64_t oa = (uint64_t)((r + ky) * TC_PITCH + kx);  // 16-byte units
                            const uint64_t ob = (uint64_t)(t * 2 * 2 * N);
                            // Tensor-core fp32 accumulation truncates (measured: -6e-8 relative per accumulate step,
                            // tools/tc_accuracy_probe.py), so the chain into the BIG accumulator is kept as short as possible:
                            // columns [0,N) receive only a_hi*w_hi; both small cross terms go to columns [N,2N).
                            umma_tf32(d, da_hi + oa, db + ob, idesc1, t == 0 ? first : 1u);
                            umma_tf32(d + N, da_lo + oa, db + ob, idesc2, 1u);
                        }
                        if (PAD == 1) {
                            // Data-gradient form: adjoint of the forward's replicate padding along y.  The forward read image row 0
                            // (H-1) once more through the clamped index, so the first (last) image row receives its own dy row a
                            // second time through the ky = 0 (ky = 2) weights -- three extra tap MMAs on the tile row of the output
                            // row itself.  (The column terms and the four corners are added by dgrad_border_kernel, mode 2.)
                            const int y = y0 + r;
                            if (y == 0 || y == H - 1) {
                                const int tb = (y == 0) ? 6 : 0;  // flipped taps: (ty = 2, tx) carries w[ky = 0][2 - tx]; (ty = 0, tx) carries w[ky = 2][..]
#pragma unroll
                                for (int tx = 0; tx < 3; ++tx) {
                                    const uint64_t oa = (uint64_t)((r + 1) * TC_PITCH + tx);
                                    const uint64_t ob = (uint64_t)((tb + tx) * 2 * 2 * N);
                                    umma_tf32(d, da_hi + oa, db + ob, idesc1, 1u);
                                    umma_tf32(d + N, da_lo + oa, db + ob, idesc2, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(ab_empty + s);                        // stage reusable once these MMAs have read it
                    if (ch == nchunks - 1) umma_commit(acc_full + ab);  // accumulator set complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= TC_XF_WARP0) {
        // ======================= transformers: raw -> BN/ReLU -> TF32 hi/lo pixel-major tiles =======================
        const int xt = tid - TC_XF_WARP0 * 32;
        constexpr int ITEMS = 2 * TROWS * TC_PITCH;
        constexpr int NIT = (ITEMS + TC_XF_THREADS - 1) / TC_XF_THREADS;
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int jlo = (x0 == 0) ? 4 : 3, jhi = min(MM + 4, W - x0 + 3);  // staged columns that exist in the image (tiles may stick out)
            for (int ch = 0; ch < nchunks; ++ch, ++g) {
                const int s = g % S, rs = g % RS;
                float* a_hi = stage0 + (size_t)s * SM::STAGE;
                float* a_lo = a_hi + SM::A_TILE;
                const float* raw = raw0 + (size_t)rs * SM::RAW_STAGE;
                if (g >= S) mbar_wait(ab_empty + s, ((g / S) - 1) & 1);
                if (xt == 0) {  // weights of this chunk: one bulk copy straight into the stage
                    mbar_arrive_expect_tx(ab_full + s, SM::B_TILE * 4);
                    bulk_g2s(a_lo + SM::A_TILE, a.wprep + (size_t)ch * SM::B_TILE, SM::B_TILE * 4, ab_full + s);
                }
                mbar_wait(raw_full + rs, (g / RS) & 1);
                const int c0 = ch * TC_KC;
#pragma unroll
                for (int i = 0; i < NIT; ++i) {
                    const int item = xt + i * TC_XF_THREADS;
                    if (item < ITEMS) {
                        const int q = item / (TROWS * TC_PITCH);
                        const int rem = item - q * (TROWS * TC_PITCH);
                        const int rr = rem / TC_PITCH, px = rem - rr * TC_PITCH;
                        int j = px + 3, rj = rr;
                        if (PAD == 0) {  // replicate: clamp onto the image; the clamped element is inside the staged box
                            j = min(max(j, jlo), jhi);
                            rj = min(max(y0 + rr - 1, 0), H - 1) - (y0 - 1);
                        }
                        float hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float t = raw[((4 * q + e) * TROWS + rj) * TC_RAW_ROW + j];
                            if (AFFINE) t = sifnn::act_affine_relu(t, sc_s[c0 + 4 * q + e], sh_s[c0 + 4 * q + e]);
                            hi[e] = tf32_hi(t);
                            lo[e] = t - hi[e];
                        }
                        *reinterpret_cast<float4*>(a_hi + (size_t)item * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4*>(a_lo + (size_t)item * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                mbar_arrive(raw_empty + rs);   // raw slot consumed (generic-proxy reads done)
                fence_proxy_async();           // st.shared -> visible to the tensor core (async proxy)
                mbar_arrive(ab_full + s);
            }
        }
    } else {
        // ======================= epilogue: TMEM -> registers -> global (+ BatchNorm statistics) =======================
        // M = 128: accumulator row m lives in TMEM lane m.  M = 64: row m lives in lane (m % 16) + 32 * (m / 16), i.e. the
        // lower 16 lanes of each quadrant (cute "half subpartitions" atom), so only lanes 0..15 of a warp carry pixels.
        const int quad = warp & 3, half = warp >> 2;
        constexpr int NH = N / 2;
        constexpr int LANES = (MM == 128) ? 32 : 16;
        const bool lane_ok = lane < LANES;
        float s1[STATS ? NH : 1], s2[STATS ? NH : 1];
#pragma unroll
        for (int j = 0; j < (STATS ? NH : 1); ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        int it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int ab = it & 1;
            mbar_wait(acc_full + ab, (it >> 1) & 1);
            tc_fence_after();
            const int x = x0 + quad * LANES + lane;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const int y = y0 + r;
#pragma unroll
                for (int n0 = 0; n0 < NH; n0 += 8) {
                    const int n = half * NH + n0;
                    float d1[8], d2[8];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ab * SM::ACC_COLS + r * 2 * N + n;
                    tmem_ld8(taddr, d1);
                    tmem_ld8(taddr + N, d2);
                    tmem_ld_wait();
                    if (y < H && lane_ok && x < W) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float v = d1[j] + d2[j];
                            if (a.bias) v += __ldg(a.bias + n + j);
                            float* op = a.out + ((size_t)b * a.O + n + j) * plane + (size_t)y * W + x;
                            if (a.accumulate) v += *op;
                            *op = v;
                            if (STATS) {
                                s1[n0 + j] += v;
                                s2[n0 + j] = fmaf(v, v, s2[n0 + j]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty + ab);   // this accumulator set may be overwritten by tile it + 2
        }
        if (STATS && a.stats) {
            // per-channel sums over all pixels this CTA produced: lanes by shuffle, then one fp64 atomic per warp and channel
#pragma unroll
            for (int j = 0; j < (STATS ? NH : 1); ++j) {
                const float t1 = sifnn::warp_sum(s1[j]), t2 = sifnn::warp_sum(s2[j]);
                if (lane == 0) {
                    atomicAdd(a.stats + half * NH + j, (double)t1);
                    atomicAdd(a.stats + a.O + half * NH + j, (double)t2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SM::TMEM_COLS);
    }
}

// =====================================================================================================================
// kx folded into N  (128-pixel MMAs, N <= 32 output channels)
//
// The 9-tap form above fetches a 4 KB activation tile from shared memory for every (tap, hi/lo) MMA while the MMA itself is only
// 2N (N) columns wide: with 16 or 32 output channels the tensor core waits on its operand fetch.  Here the three kx taps of a
// filter row are stacked along N instead:
//
//   E[pixel p][(s, kx, o)] += A[p][8 ch] * W[ky][(s, kx, o)][8 ch]        one MMA pair per (row, ky): 6 instead of 18 per chunk and row
//   out[y][x][o] = E_kx0[x-1] + E_kx1[x] + E_kx2[x+1]                      (hi*hi block + cross block, summed by the epilogue)
//
// E_kx[p] depends on pixel p alone, so the tile has no halo columns, the kx shift becomes a +-1 LANE shift of the accumulator
// (warp shuffles, quadrant edges through a small shared-memory exchange, tile edges carried from one tile to the next -- a CTA walks
// the tiles of an image row in order), and padding is a rule on the edge term: replicate  E_kx0[-1] = E_kx0[0], E_kx2[W] = E_kx2[W-1];
// zero (data gradient): both vanish.
// TMEM: one row slot = 6N columns ([hi*hi : kx0 kx1 kx2][cross : kx0 kx1 kx2]), R slots used as a ring with per-row full/empty
// barriers (the epilogue drains row r of tile i while the MMAs of rows r+1.. and of tile i+1 run), instead of two full sets.
// =====================================================================================================================
// WCH > 0: the split weights of all WCH chunks stay resident in shared memory for the whole kernel (they are the same for every tile);
// WCH == 0: each stage carries the weights of its chunk (more chunks than fit).
// BF: 3-term BF16 split instead of TF32 (K = 16 channels per MMA and chunk: half the MMA instructions; opt-in)
template <int N, int R, int WCH, bool BF = false>
struct TcxSmem {
    static constexpr int TROWS = R + 2;
    static constexpr int KC = BF ? 16 : TC_KC;                  // input channels per chunk
    static constexpr int A_TILE = 2 * TROWS * 128 * 4;          // floats per (hi or lo) tile: [2 q][TROWS*128 px][4]
    static constexpr int B_TILE = 3 * 2 * 6 * N * 4;            // floats: [3 ky][2 q][6N rows = (s, kx, o)][4]
    static constexpr int STAGE = 2 * A_TILE + (WCH ? 0 : B_TILE);
    static constexpr int W_ALL = WCH * B_TILE;                  // floats
    static constexpr int RAW_STAGE = KC * TROWS * 128;          // floats
    static constexpr int CTRL_FLOATS = 512 + 2 * 2 * 2 * 4 * N + 2 * 2 * R * N;   // barriers / BatchNorm affine, edge exchange (2 buffers), tile carries (2 buffers)
    static constexpr int CTRL_PAD = (CTRL_FLOATS + 31) / 32 * 32;
    static constexpr int BUDGET = 222 * 1024;
    // two transformed stages are enough (transform and MMA are short); everything left goes to raw boxes in flight (up to 4), which is what
    // hides the load latency
    static constexpr int STAGES = 2;
    static constexpr int RAW_ROOM = (BUDGET - CTRL_PAD * 4 - W_ALL * 4 - STAGES * STAGE * 4) / (RAW_STAGE * 4);
    static constexpr int RAW_STAGES = RAW_ROOM >= 4 ? 4 : RAW_ROOM;
    static constexpr int REST = BUDGET - CTRL_PAD * 4 - RAW_STAGES * RAW_STAGE * 4 - W_ALL * 4;
    static_assert(RAW_STAGES >= 2, "at least two raw boxes in flight");
    static constexpr size_t BYTES = (size_t)(STAGES * STAGE + RAW_STAGES * RAW_STAGE + CTRL_PAD + W_ALL) * 4;
    static constexpr int ROW_COLS = 6 * N;
    // Row slots of the TMEM ring: one MORE than the rows of a tile where it fits.  With exactly R slots the MMA of (tile i+1, row r)
    // waits for the drain of (tile i, row r), which was committed only a few MMAs earlier: a barrier round trip (~1000 clocks) per row.
    // The spare slot lets the issue thread run one row ahead of the drain.
    static constexpr int NSLOT = ((R + 1) * ROW_COLS <= 512) ? R + 1 : R;
    static constexpr int TMEM_COLS = 512;
    static_assert(R * ROW_COLS <= 512 && NSLOT <= 5, "row slots must fit the 512 TMEM columns");
    static_assert(STAGE * 4 * 2 <= REST, "two stages must fit");
    static_assert((STAGE * 4) % 128 == 0 && (RAW_STAGE * 4) % 128 == 0 && (CTRL_PAD * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(6 * N <= 256 && R <= 4, "UMMA N limit / barrier slots");
};

// EW = number of epilogue warps: 8 (one group: TMEM lane quadrant = warp % 4, channel half = warp / 4, every row) or 16 (two such groups,
// group g drains the rows r with r % 2 == g: two rows in flight, because a row's drain is one long latency chain -- barrier wake-up,
// tcgen05.ld, named barrier, shuffles, stores).  With 16 epilogue warps the transformers get 4 warps instead of 8 (register budget).
template <int N, int R, int PAD, bool AFFINE, int EW, int WCH, bool BF>
__global__ void __launch_bounds__((EW + 2 + (EW == 16 ? 4 : 8)) * 32, 1) conv3x3_tcx_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tmap) {
    using SM = TcxSmem<N, R, WCH, BF>;
    constexpr int KC = SM::KC;
    constexpr int TROWS = SM::TROWS;
    constexpr int XW = (EW == 16) ? 4 : 8, XF_T = XW * 32, EG = EW / 8;   // transformer warps / threads, epilogue row groups
    constexpr int X_LOAD_WARP = EW, X_MMA_WARP = EW + 1, X_XF_WARP0 = EW + 2, X_THREADS = (EW + 2 + XW) * 32;
    static_assert(R % EG == 0, "rows split evenly over the epilogue groups");
    constexpr bool STATS = (PAD == 0);
    constexpr int S = SM::STAGES;
    constexpr int RS = SM::RAW_STAGES;
    extern __shared__ __align__(128) float smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* ab_full = bars;            // [4]
    uint64_t* ab_empty = bars + 4;       // [4]
    uint64_t* raw_full = bars + 8;       // [4]
    uint64_t* raw_empty = bars + 12;     // [4]
    uint64_t* acc_full = bars + 16;      // [NSLOT <= 5]
    uint64_t* acc_empty = bars + 22;     // [NSLOT <= 5]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
    uint64_t* w_full = bars + 30;        // resident weights have landed (WCH > 0)
    constexpr int NSLOT = SM::NSLOT;
    float* sc_s = smem + 64;    // [<=128]
    float* sh_s = smem + 192;   // [<=128]
    float* xch = smem + 512;                     // [2 groups][2 buffers][2: e0 of lane 31 | e2 of lane 0][4 quadrants][N]
    float* carry_e0 = xch + 2 * 2 * 2 * 4 * N;       // [2][R][N]: E_kx0 of the last pixel of the previous tile of this image row (buffer = tile parity)
    float* carry_out = carry_e0 + 2 * R * N;     // [2][R][N]: the unfinished output of that pixel
    float* stage0 = smem + SM::CTRL_PAD;
    float* raw0 = stage0 + (size_t)S * SM::STAGE;
    float* wall = raw0 + (size_t)RS * SM::RAW_STAGE;   // [WCH][B_TILE] resident weights

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, K = a.K;
    const size_t plane = (size_t)H * W;
    const int nchunks = K / KC;
    const int tiles_x = a.tiles_x;
    const int tiles_per_img = tiles_x * a.tiles_y;
    const int ngroups = a.num_tiles / tiles_x;   // a CTA owns whole row groups: the tiles_x tiles of R image rows, walked left to right

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(ab_full + s, XW + (WCH ? 0 : 1)); mbar_init(ab_empty + s, 1); }   // one arrival per warp
        mbar_init(w_full, 1);
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, XW); }
        for (int s = 0; s < NSLOT; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 8); }
        fence_mbar_init();
    }
    if (warp == X_MMA_WARP) tmem_alloc(tmem_slot, SM::TMEM_COLS);
    if (AFFINE) {
        for (int i = tid; i < K; i += X_THREADS) { sc_s[i] = __ldg(a.in_scale + i); sh_s[i] = __ldg(a.in_shift + i); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (WCH > 0 && tid == 0) {   // all weights of the layer, once: one bulk copy
        mbar_arrive_expect_tx(w_full, SM::W_ALL * 4);
        bulk_g2s(wall, a.wprep, SM::W_ALL * 4, w_full);
    }

    auto tile_coords = [&](int tile, int& b, int& y0, int& x0) {
        b = tile / tiles_per_img;
        const int t = tile - b * tiles_per_img;
        const int ty = t / tiles_x;
        y0 = ty * R;
        x0 = (t - ty * tiles_x) * 128;
    };

    if (warp == X_LOAD_WARP) {
        // ======================= loader: one TMA box (8 ch x TROWS rows x 128 cols) per chunk =======================
        int g = 0;
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            for (int tx = 0; tx < tiles_x; ++tx) {
                int b, y0, x0;
                tile_coords(grp * tiles_x + tx, b, y0, x0);
                for (int ch = 0; ch < nchunks; ++ch, ++g) {
                    const int rs = g % RS;
                    if (g >= RS) mbar_wait_warp(raw_empty + rs, ((g / RS) - 1) & 1);
                    if (lane == 0) {
                        TC_STAMP(0, g);
                        mbar_arrive_expect_tx(raw_full + rs, SM::RAW_STAGE * 4);
                        tma_load_3d(raw0 + (size_t)rs * SM::RAW_STAGE, &tmap, x0, y0 - 1, b * K + ch * KC, raw_full + rs);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == X_MMA_WARP) {
        // ======================= MMA issuer (one thread) =======================
        constexpr uint32_t idesc1 = BF ? make_idesc_bf16(128, 6 * N) : make_idesc(128, 6 * N);  // a_hi x [w_hi(kx0..2) ; w_lo(kx0..2)]
        constexpr uint32_t idesc2 = BF ? make_idesc_bf16(128, 3 * N) : make_idesc(128, 3 * N);  // a_lo x  w_hi(kx0..2)
        auto mma = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) { if (BF) umma_bf16(d, da, db, id, acc); else umma_tf32(d, da, db, id, acc); };
        int g = 0, it = 0;
        if (WCH > 0) mbar_wait(w_full, 0);
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            for (int tx = 0; tx < tiles_x; ++tx, ++it) {
                int b, y0, x0;
                tile_coords(grp * tiles_x + tx, b, y0, x0);
                for (int ch = 0; ch < nchunks; ++ch, ++g) {
                    const int s = g % S;
                    mbar_wait(ab_full + s, (g / S) & 1);
                    tc_fence_after();
                    if (lane == 0) TC_STAMP(4, g);
                    constexpr uint32_t LBO_A = TROWS * 128 * 16, LBO_B = 6 * N * 16, SBO = 128;
                    const uint32_t a_hi = smem_u32(stage0 + (size_t)s * SM::STAGE);
                    const uint64_t da_hi = make_desc(a_hi, LBO_A, SBO);
                    const uint64_t da_lo = da_hi + (uint64_t)(SM::A_TILE * 4 / 16);
                    const uint64_t db = WCH ? make_desc(smem_u32(wall + (size_t)ch * SM::B_TILE), LBO_B, SBO)
                                            : da_hi + (uint64_t)(2 * SM::A_TILE * 4 / 16) - ((uint64_t)(LBO_A >> 4) << 16) + ((uint64_t)(LBO_B >> 4) << 16);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (lane == 0) {
                            const int gr = it * R + r, slot = gr % NSLOT, use = gr / NSLOT;   // row slots are handed out round-robin
                            if (ch == 0 && use >= 1) {   // the slot still holds an earlier row until the epilogue has drained it
                                mbar_wait(acc_empty + slot, (use - 1) & 1);
                                tc_fence_after();
                            }
                            if (ch == 0) TC_STAMP(8, gr);
                            const uint32_t d = tmem_base + slot * SM::ROW_COLS;
#ifdef SIFNN_TC_ABLATE
                            if (!(a.ablate & 1))
#endif
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
                                const uint64_t oa = (uint64_t)((r + ky) * 128);  // 16-byte units
                                const uint64_t ob = (uint64_t)(ky * 2 * 6 * N);
                                mma(d, da_hi + oa, db + ob, idesc1, (ky == 0 && ch == 0) ? 0u : 1u);
                                mma(d + 3 * N, da_lo + oa, db + ob, idesc2, 1u);
                            }
                            if (PAD == 1) {
                                // Adjoint of the forward's replicate padding along y: the first (last) image row receives its own dy row a second
                                // time through the ky = 0 (ky = 2) weights, i.e. the flipped filter row 2 (0), on the tile row of the output row itself.
                                const int y = y0 + r;
                                if (y == 0 || y == H - 1) {
                                    const uint64_t oa = (uint64_t)((r + 1) * 128);
                                    const uint64_t ob = (uint64_t)(((y == 0) ? 2 : 0) * 2 * 6 * N);
                                    mma(d, da_hi + oa, db + ob, idesc1, 1u);
                                    mma(d + 3 * N, da_lo + oa, db + ob, idesc2, 1u);
                                    if (H == 1) {   // a one-row image is both first and last
                                        const uint64_t ob2 = 0;
                                        mma(d, da_hi + oa, db + ob2, idesc1, 1u);
                                        mma(d + 3 * N, da_lo + oa, db + ob2, idesc2, 1u);
                                    }
                                }
                            }
                            if (ch == nchunks - 1) umma_commit(acc_full + slot);   // row slot complete
                        }
                        __syncwarp();
                    }
                    if (lane == 0) { TC_STAMP(5, g); umma_commit(ab_empty + s); }                   // stage reusable once these MMAs have read it
                    __syncwarp();
                }
            }
        }
    } else if (warp >= X_XF_WARP0) {
        // ======================= transformers: raw -> BN/ReLU -> TF32 hi/lo pixel-major tiles (no halo columns) =======================
        const int xt = tid - X_XF_WARP0 * 32;
        constexpr int ITEMS = 2 * TROWS * 128;
        constexpr int NIT = ITEMS / XF_T;
        static_assert(ITEMS % XF_T == 0, "whole items per thread");
        int g = 0;
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            for (int tx = 0; tx < tiles_x; ++tx) {
                int b, y0, x0;
                tile_coords(grp * tiles_x + tx, b, y0, x0);
                for (int ch = 0; ch < nchunks; ++ch, ++g) {
                    const int s = g % S, rs = g % RS;
                    float* a_hi = stage0 + (size_t)s * SM::STAGE;
                    float* a_lo = a_hi + SM::A_TILE;
                    const float* raw = raw0 + (size_t)rs * SM::RAW_STAGE;
                    if (g >= S) mbar_wait_warp(ab_empty + s, ((g / S) - 1) & 1);
                    if (xt == 0) TC_STAMP(1, g);
                    if (WCH == 0 && xt == 0) {  // weights of this chunk: one bulk copy straight into the stage
                        mbar_arrive_expect_tx(ab_full + s, SM::B_TILE * 4);
                        bulk_g2s(a_lo + SM::A_TILE, a.wprep + (size_t)ch * SM::B_TILE, SM::B_TILE * 4, ab_full + s);
                    }
                    mbar_wait_warp(raw_full + rs, (g / RS) & 1);
                    if (xt == 0) TC_STAMP(2, g);
                    const int c0 = ch * KC;
#ifdef SIFNN_TC_ABLATE
                    if (!(a.ablate & 4))
#endif
#pragma unroll
                    for (int i = 0; i < NIT; ++i) {
                        const int item = xt + i * XF_T;       // (q, row, pixel)
                        const int px = item & 127;
                        const int rr = (item >> 7) % TROWS;
                        const int q = item / (128 * TROWS);
                        int rj = rr;
                        if (PAD == 0) rj = min(max(y0 + rr - 1, 0), H - 1) - (y0 - 1);   // replicate padding along y
                        if constexpr (!BF) {
                            float hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float t = raw[((4 * q + e) * TROWS + rj) * 128 + px];
                                if (AFFINE) t = sifnn::act_affine_relu(t, sc_s[c0 + 4 * q + e], sh_s[c0 + 4 * q + e]);
                                hi[e] = tf32_hi(t);
                                lo[e] = t - hi[e];
                            }
                            *reinterpret_cast<float4*>(a_hi + (size_t)item * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<float4*>(a_lo + (size_t)item * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        } else {   // 8 channels of one pixel -> one 16-byte unit of 8 BF16 (hi tile) and one of the residuals (lo tile)
                            uint32_t hp[4], lp[4];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                unsigned short h0, l0, h1, l1;
                                float t0 = raw[((8 * q + e) * TROWS + rj) * 128 + px], t1 = raw[((8 * q + e + 1) * TROWS + rj) * 128 + px];
                                if (AFFINE) {
                                    t0 = sifnn::act_affine_relu(t0, sc_s[c0 + 8 * q + e], sh_s[c0 + 8 * q + e]);
                                    t1 = sifnn::act_affine_relu(t1, sc_s[c0 + 8 * q + e + 1], sh_s[c0 + 8 * q + e + 1]);
                                }
                                bf16_split(t0, h0, l0);
                                bf16_split(t1, h1, l1);
                                hp[e >> 1] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                                lp[e >> 1] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                            }
                            *reinterpret_cast<uint4*>(a_hi + (size_t)item * 4) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                            *reinterpret_cast<uint4*>(a_lo + (size_t)item * 4) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                        }
                    }
                    if (xt == 0) TC_STAMP(3, g);
                    fence_proxy_async();           // this thread's st.shared -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if ((xt & 31) == 0) {          // one arrival per warp: 8 instead of 256 atomics on the barrier word
                        mbar_arrive(raw_empty + rs);
                        mbar_arrive(ab_full + s);
                    }
                }
            }
        }
    } else {
        // ======================= epilogue: TMEM -> lane-shifted sum over kx -> global (+ BatchNorm statistics) =======================
        // The eight epilogue warps are the critical path of this kernel once the operand fetch is cut threefold, so the per-row code is
        // kept lean: everything that depends only on the thread (channel base, exchange slots) is hoisted, the common path (31 of 32
        // lanes on each side) is straight-line shuffles and adds, and the edge lanes patch their neighbour terms in two short branches.
        const int quad = warp & 3, half = (warp >> 2) & 1, eg = warp >> 3;   // eg: row group of this warp
        constexpr int NH = N / 2;                    // channels of this warp: [half * NH, half * NH + NH)
        constexpr int LW = NH >= 8 ? 8 : 4;          // TMEM load width (columns)
        constexpr int NB = NH / LW;
        static_assert(NH % LW == 0 && NH % 4 == 0, "whole TMEM loads per warp");
        const int nbase = half * NH;
        float s1[STATS ? NH : 1], s2[STATS ? NH : 1];
#pragma unroll
        for (int j = 0; j < (STATS ? NH : 1); ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        const bool accum = a.accumulate != 0, has_bias = a.bias != nullptr;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + nbase;
        int it = 0, rowcnt = 0;
        for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            for (int tx = 0; tx < tiles_x; ++tx, ++it) {
                int b, y0, x0;
                tile_coords(grp * tiles_x + tx, b, y0, x0);
                const bool img_left = (x0 == 0), img_right = (x0 + 128 == W);
                const bool take_carry = (!img_left) && quad == 0;     // warp-uniform; lane 0 finishes the previous tile's last pixel
                const bool defer = (!img_right) && quad == 3;         // warp-uniform; lane 31's right neighbour lives in the next tile
                float* orow = a.out + ((size_t)b * a.O + nbase) * plane + (size_t)(y0 + eg) * W + x0 + quad * 32 + lane;
#pragma unroll 1
                for (int r = eg; r < R; r += EG, ++rowcnt, orow += EG * W) {
                    const int gr = it * R + r, slot = gr % NSLOT;
                    mbar_wait_warp(acc_full + slot, (gr / NSLOT) & 1);
                    tc_fence_after();
                    if (tid == 0) TC_STAMP(6, gr);
                    float e0[NH], e1[NH], e2[NH];
#ifdef SIFNN_TC_ABLATE
                    if (a.ablate & 8) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + slot);
                        continue;
                    }
#endif
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        const uint32_t taddr = tlane + slot * SM::ROW_COLS + nb * LW;
                        float h0[LW], h1[LW], h2[LW], c0[LW], c1[LW], c2[LW];
                        tmem_ldw<LW>(taddr, h0); tmem_ldw<LW>(taddr + N, h1); tmem_ldw<LW>(taddr + 2 * N, h2);
                        tmem_ldw<LW>(taddr + 3 * N, c0); tmem_ldw<LW>(taddr + 4 * N, c1); tmem_ldw<LW>(taddr + 5 * N, c2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < LW; ++j) { e0[nb * LW + j] = h0[j] + c0[j]; e1[nb * LW + j] = h1[j] + c1[j]; e2[nb * LW + j] = h2[j] + c2[j]; }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + slot);   // the slot may be overwritten (one arrival per warp)
                    // quadrant edges: lane 31 publishes its E_kx0 (the right neighbour's left term), lane 0 its E_kx2
                    float* xb = xch + (eg * 2 + (rowcnt & 1)) * (2 * 4 * N) + quad * N + nbase;
                    if (lane == 31) {
#pragma unroll
                        for (int j = 0; j < NH; j += 4) *reinterpret_cast<float4*>(xb + j) = make_float4(e0[j], e0[j + 1], e0[j + 2], e0[j + 3]);
                    } else if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < NH; j += 4) *reinterpret_cast<float4*>(xb + 4 * N + j) = make_float4(e2[j], e2[j + 1], e2[j + 2], e2[j + 3]);
                    }
                    // tile carries alternate between two buffers by tile parity: this tile reads what the previous one wrote R rows (>= 1 barrier) ago
                    const float* cin_e0 = carry_e0 + (((it + 1) & 1) * R + r) * N + nbase;
                    const float* cin_out = carry_out + (((it + 1) & 1) * R + r) * N + nbase;
                    float* cout_e0 = carry_e0 + ((it & 1) * R + r) * N + nbase;
                    float* cout_out = carry_out + ((it & 1) * R + r) * N + nbase;
                    asm volatile("bar.sync %0, 256;" ::"r"(1 + eg) : "memory");   // the 8 warps of this row group
                    // in-place lane shift: e0 <- left neighbour's E_kx0, e2 <- right neighbour's E_kx2.  Lane 0 (31) keeps its own value, which is
                    // exactly the replicate-padding rule at the image edge; elsewhere the edge lanes patch the term in below.
#pragma unroll
                    for (int j = 0; j < NH; ++j) {
                        e0[j] = __shfl_up_sync(0xffffffffu, e0[j], 1);
                        e2[j] = __shfl_down_sync(0xffffffffu, e2[j], 1);
                    }
                    if (lane == 0) {
                        if (quad > 0) {
#pragma unroll
                            for (int j = 0; j < NH; ++j) e0[j] = xb[j - N];
                        } else if (!img_left) {
#pragma unroll
                            for (int j = 0; j < NH; ++j) e0[j] = cin_e0[j];
                        } else if (PAD == 1) {
#pragma unroll
                            for (int j = 0; j < NH; ++j) e0[j] = 0.f;
                        }
                    } else if (lane == 31) {
                        if (quad < 3) {
#pragma unroll
                            for (int j = 0; j < NH; ++j) e2[j] = xb[4 * N + N + j];
                        } else if (!img_right || PAD == 1) {   // deferred to the next tile, or zero padding
#pragma unroll
                            for (int j = 0; j < NH; ++j) e2[j] = 0.f;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NH; ++j) e1[j] += e0[j] + e2[j];
                    float* const v = e1;
#ifdef SIFNN_TC_ABLATE
                    if (a.ablate & 2) continue;
#endif
                    if (tid == 0) TC_STAMP(7, gr);
                    if (y0 + r < H) {
                        if (defer && lane == 31) {
#pragma unroll
                            for (int j = 0; j < NH; ++j) { cout_out[j] = v[j]; cout_e0[j] = xb[j]; }   // xb[j]: this lane's own E_kx0, published above
                        } else {
                            float* op = orow;
#pragma unroll
                            for (int j = 0; j < NH; ++j, op += plane) {
                                float o = v[j];
                                if (has_bias) o += __ldg(a.bias + nbase + j);
                                if (accum) o += *op;
                                *op = o;
                                if (STATS) { s1[j] += o; s2[j] = fmaf(o, o, s2[j]); }
                            }
                        }
                        if (take_carry && lane == 0) {   // finish pixel x0 - 1 (the last pixel of the previous tile of this row)
                            float* op = orow - 1;
#pragma unroll
                            for (int j = 0; j < NH; ++j, op += plane) {
                                float o = cin_out[j] + xb[4 * N + j];   // + this pixel's own E_kx2 (published above)
                                if (has_bias) o += __ldg(a.bias + nbase + j);
                                if (accum) o += *op;
                                *op = o;
                                if (STATS) { s1[j] += o; s2[j] = fmaf(o, o, s2[j]); }
                            }
                        }
                    }
                }
            }
        }
        if (STATS && a.stats) {
#pragma unroll
            for (int j = 0; j < (STATS ? NH : 1); ++j) {
                const float t1 = sifnn::warp_sum(s1[j]), t2 = sifnn::warp_sum(s2[j]);
                if (lane == 0) {
                    atomicAdd(a.stats + half * NH + j, (double)t1);
                    atomicAdd(a.stats + a.O + half * NH + j, (double)t2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == X_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SM::TMEM_COLS);
    }
}

// Weight split layouts (produced by tc_prep_many_kernel at the end of this file), the exact shared-memory image of a stage:
// layout 0: wprep[chunk][tap][q][row][e], row < N: hi of W[n = row][c = 8*chunk + 4q + e][tap], row >= N: lo of W[n = row - N];
// layout 1: wprep[chunk][ky][q][row][e], row = (s * 3 + kx) * N + n  (hi rows of the three kx first, then the lo rows);
// layout 2: the same with BF16 pairs, wprep[chunk of 16 ch][ky][q][row][8 bf16], channel = 16*chunk + 8q + e.

template <int N, int R, int PAD, bool AFFINE, int EW, int WCH, bool BF>
int launch_tcx_w(const TcArgs& a0, cudaStream_t st) {
    using SM = TcxSmem<N, R, WCH, BF>;
    auto kern = conv3x3_tcx_kernel<N, R, PAD, AFFINE, EW, WCH, BF>;
    static sifnn::PerDeviceOnce attr_once;   // the attribute is per device: one flag per device, not one per process
    if (attr_once.first_time()) {
        SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
    }
    TcArgs a = a0;
    { const char* e = getenv("SIFNN_TC_ABLATE"); a.ablate = e ? atoi(e) : 0; }
#ifdef SIFNN_TC_TRACE
    a.trace = g_trace;
#endif
    a.tiles_x = a.W / 128;
    a.tiles_y = (a.H + R - 1) / R;
    a.num_tiles = a.B * a.tiles_x * a.tiles_y;
    CUtensorMap tmap;
    SIFNN_REQUIRE(encode_planes_map(&tmap, a.in, a.W, a.H, (long long)a.B * a.K, 128, R + 2, SM::KC),
                  "conv3x3_tc: cuTensorMapEncodeTiled is unavailable or failed");
    const int groups = a.num_tiles / a.tiles_x;
    const int grid = groups < sifnn::num_sms() ? groups : sifnn::num_sms();
    kern<<<grid, (EW + 2 + (EW == 16 ? 4 : 8)) * 32, SM::BYTES, st>>>(a, tmap);
    return sifnn::check_launch("conv3x3_tcx_kernel");
}

// SIFNN_TC_BF16=1: 3-term BF16 split in the kx-folded kernel for 16-output-channel layers (experiment; accuracy ~1e-5 instead of ~1e-6)
bool tc_bf16() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_TC_BF16"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
bool tc_use_bf16(int K, int O, int W);

// resident weights when the layer has 2 or 4 chunks of input channels (16 / 32: every 16-output-channel use in ModelB)
template <int N, int R, int PAD, bool AFFINE, int EW>
int launch_tcx(const TcArgs& a, cudaStream_t st) {
    if (N == 16 && tc_bf16()) {   // opt-in 3-term BF16 split (16-output-channel layers only so far)
        if (a.K == 16) return launch_tcx_w<N, R, PAD, AFFINE, EW, (N == 16 ? 1 : 0), (N == 16)>(a, st);
        if (a.K == 32) return launch_tcx_w<N, R, PAD, AFFINE, EW, (N == 16 ? 2 : 0), (N == 16)>(a, st);
    }
    if (N == 16 && a.K == 16) return launch_tcx_w<N, R, PAD, AFFINE, EW, (N == 16 ? 2 : 0), false>(a, st);
    if (N == 16 && a.K == 32) return launch_tcx_w<N, R, PAD, AFFINE, EW, (N == 16 ? 4 : 0), false>(a, st);
    return launch_tcx_w<N, R, PAD, AFFINE, EW, 0, false>(a, st);
}

// the kx-folded kernel serves the 128-pixel MMAs with <= 32 output channels; SIFNN_TC_KXFOLD=0 falls back to the 9-tap kernel (A/B runs)
bool tc_fold_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_TC_KXFOLD"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// measured (profiles/r1n_conv3x3_tc_kxfold.log): 16 output channels gain 8-25 %; 32 output channels only break even from 64 input channels on
// (two row slots fit the TMEM ring instead of four, and the epilogue works twice as long per row)
// experiment switch: SIFNN_TC_EPI16=1 gives the kx-folded kernel 16 epilogue warps in two row groups
bool tc_epi16() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_TC_EPI16"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
bool tc_use_bf16(int K, int O, int W) { return tc_bf16() && tc_fold_enabled() && (W % 128 == 0) && O == 16 && (K == 16 || K == 32); }
bool tc_use_fold(int K, int O, int W) { return tc_fold_enabled() && (W % 128 == 0) && (O == 16 || (O == 32 && K >= 64)); }

template <int N, int R, int MM, int PAD, bool AFFINE>
int launch_tc(const TcArgs& a0, cudaStream_t st) {
    using SM = TcSmem<N, R, MM>;
    auto kern = conv3x3_tc_kernel<N, R, MM, PAD, AFFINE>;
    static sifnn::PerDeviceOnce attr_once;   // the attribute is per device: one flag per device, not one per process
    if (attr_once.first_time()) {
        SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
    }
    TcArgs a = a0;
    a.tiles_x = (a.W + MM - 1) / MM;
    a.tiles_y = (a.H + R - 1) / R;
    a.num_tiles = a.B * a.tiles_x * a.tiles_y;
    // activation tensor as a 3-D TMA tensor {W, H, B*K planes}; box = {136 columns, R+2 rows, 8 planes}
    CUtensorMap tmap;
    SIFNN_REQUIRE(encode_planes_map(&tmap, a.in, a.W, a.H, (long long)a.B * a.K, SM::RAW_ROW, R + 2, TC_KC),
                  "conv3x3_tc: cuTensorMapEncodeTiled is unavailable or failed");
    const int grid = a.num_tiles < sifnn::num_sms() ? a.num_tiles : sifnn::num_sms();  // persistent: one CTA per SM
    kern<<<grid, TC_THREADS2, SM::BYTES, st>>>(a, tmap);
    return sifnn::check_launch("conv3x3_tc_kernel");
}

// 4 output rows per tile for the 128-pixel MMAs with <= 32 output channels (halo-row staging overhead 1.5x instead of 2x: 5-12 % faster,
// profiles/r1m_conv3x3_tc_rows2_vs_rows4.log); SIFNN_TC_R4=0 restores 2 rows for A/B measurements.
bool tc_rows4() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_TC_R4"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <int PAD, bool AFFINE>
int dispatch_tc(const TcArgs& a, cudaStream_t st) {
    if (tc_use_fold(a.K, a.O, a.W)) {
        if (a.O == 16) return tc_epi16() ? launch_tcx<16, 4, PAD, AFFINE, 16>(a, st) : launch_tcx<16, 4, PAD, AFFINE, 8>(a, st);
        return tc_epi16() ? launch_tcx<32, 2, PAD, AFFINE, 16>(a, st) : launch_tcx<32, 2, PAD, AFFINE, 8>(a, st);
    }
    if (a.W % 128 == 0) {
        switch (a.O) {
            case 16: return tc_rows4() ? launch_tc<16, 4, 128, PAD, AFFINE>(a, st) : launch_tc<16, 2, 128, PAD, AFFINE>(a, st);
            case 32: return tc_rows4() ? launch_tc<32, 4, 128, PAD, AFFINE>(a, st) : launch_tc<32, 2, 128, PAD, AFFINE>(a, st);
            case 64: return launch_tc<64, 2, 128, PAD, AFFINE>(a, st);
        }
    } else {  // 64-pixel-wide images: one image row per M = 64 MMA
        switch (a.O) {
            case 16: return launch_tc<16, 4, 64, PAD, AFFINE>(a, st);
            case 32: return launch_tc<32, 4, 64, PAD, AFFINE>(a, st);
            case 64: return launch_tc<64, 2, 64, PAD, AFFINE>(a, st);
            case 128: if constexpr (PAD == 1 && !AFFINE) return launch_tc<128, 1, 64, PAD, AFFINE>(a, st); else break;
        }
    }
    sifnn::set_error("conv3x3_tc: unsupported channel count %d", a.O);
    return SIFNN_EINVAL;
}

}  // namespace

extern "C" int sifnn_conv3x3_tc_supported(int Cin, int Cout, int H, int W) {
    // Any width that is a multiple of 4 (TMA row pitch) and at least 16: multiples of 128 use M = 128 MMAs, everything else M = 64 MMAs
    // with the columns past the image masked.  128 output channels (data gradient of ub1's first convolution) only in the M = 64
    // form: the stage would not fit otherwise.
    return (W % 4 == 0) && W >= 16 && (Cin % 8 == 0) && Cin <= 128 && (Cout == 16 || Cout == 32 || Cout == 64 || (Cout == 128 && W % 128 != 0)) && H >= 1;
}

extern "C" size_t sifnn_conv3x3_tc_wprep_bytes(int Cin, int Cout) { return (size_t)(Cin / 8) * 9 * 2 * 2 * Cout * 4 * sizeof(float); }

namespace {

constexpr int TC_PREP_MAX = 24;
struct TcPrepBatch { sifnn::TcPrepJob j[TC_PREP_MAX]; };

// every job of the batch in one launch: blockIdx.y = job, the x dimension strides over that job's elements
__global__ void __launch_bounds__(256) tc_prep_many_kernel(const __grid_constant__ TcPrepBatch batch) {
    const sifnn::TcPrepJob& J = batch.j[blockIdx.y];
    const int N = J.N;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < J.total; idx += gridDim.x * blockDim.x) {
        int n, q, t, c, lo;
        if (J.layout == 0) {
            const int e = idx & 3, row = (idx >> 2) % (2 * N);
            q = ((idx >> 2) / (2 * N)) & 1;
            t = ((idx >> 2) / (2 * N * 2)) % 9;
            const int chunk = (idx >> 2) / (2 * N * 2 * 9);
            n = row < N ? row : row - N;
            lo = row >= N;
            c = chunk * 8 + 4 * q + e;
        } else if (J.layout == 1) {
            const int e = idx & 3, row = (idx >> 2) % (6 * N);
            q = ((idx >> 2) / (6 * N)) & 1;
            const int ky = ((idx >> 2) / (6 * N * 2)) % 3, chunk = (idx >> 2) / (6 * N * 2 * 3);
            lo = row >= 3 * N;
            const int rr = lo ? row - 3 * N : row;
            t = ky * 3 + rr / N;
            n = rr % N;
            c = chunk * 8 + 4 * q + e;
        } else {
            const int e = idx & 7, row = (idx >> 3) % (6 * N);
            q = ((idx >> 3) / (6 * N)) & 1;
            const int ky = ((idx >> 3) / (6 * N * 2)) % 3, chunk = (idx >> 3) / (6 * N * 2 * 3);
            lo = row >= 3 * N;
            const int rr = lo ? row - 3 * N : row;
            t = ky * 3 + rr / N;
            n = rr % N;
            c = chunk * 16 + 8 * q + e;
        }
        const float v = __ldg(J.w + (size_t)n * J.w_so + (size_t)c * J.w_sk + (J.flip ? 8 - t : t));
        if (J.layout == 2) {
            unsigned short h, l;
            bf16_split(v, h, l);
            static_cast<unsigned short*>(J.wprep)[idx] = lo ? l : h;
        } else {
            const float hi = tf32_hi(v);
            static_cast<float*>(J.wprep)[idx] = lo ? v - hi : hi;
        }
    }
}

sifnn::TcPrepJob make_prep_job(const float* w, void* wprep, int K, int N, int w_so, int w_sk, int flip, int W) {
    sifnn::TcPrepJob j{};
    j.w = w; j.wprep = wprep; j.K = K; j.N = N; j.w_so = w_so; j.w_sk = w_sk; j.flip = flip;
    j.layout = tc_use_bf16(K, N, W) ? 2 : (tc_use_fold(K, N, W) ? 1 : 0);
    j.total = (j.layout == 2) ? (K / 16) * 3 * 2 * 6 * N * 8 : (K / 8) * 9 * 2 * 2 * N * 4;
    return j;
}

}  // namespace

namespace sifnn {

TcPrepJob tc_prep_job_fwd(const float* w, void* wprep, int Cin, int Cout, int W) { return make_prep_job(w, wprep, Cin, Cout, Cin * 9, 9, 0, W); }
TcPrepJob tc_prep_job_dgrad(const float* w, void* wprep, int Cin, int Cout, int W) { return make_prep_job(w, wprep, Cout, Cin, 9, Cin * 9, 1, W); }

int tc_prep_many(const TcPrepJob* jobs, int n, cudaStream_t st) {
    for (int i0 = 0; i0 < n; i0 += TC_PREP_MAX) {
        TcPrepBatch b{};
        const int m = n - i0 < TC_PREP_MAX ? n - i0 : TC_PREP_MAX;
        int mx = 0;
        for (int i = 0; i < m; ++i) { b.j[i] = jobs[i0 + i]; if (b.j[i].total > mx) mx = b.j[i].total; }
        if (m == 0 || mx == 0) continue;
        int bx = (mx + 256 * 4 - 1) / (256 * 4);
        if (bx > 64) bx = 64;
        tc_prep_many_kernel<<<dim3(bx, m), 256, 0, st>>>(b);
        SIFNN_TRY(check_launch("tc_prep_many_kernel"));
    }
    return 0;
}

int conv3x3_fwd_tc_prepped(const float* in, const float* in_scale, const float* in_shift, const void* wprep, const float* bias, float* out, double* stats,
                           int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
    TcArgs a{};
    a.in = in; a.in_scale = in_scale; a.in_shift = in_shift; a.wprep = static_cast<const float*>(wprep); a.bias = bias; a.out = out; a.stats = stats;
    a.B = B; a.K = Cin; a.O = Cout; a.H = H; a.W = W; a.accumulate = 0;
    return in_scale ? dispatch_tc<0, true>(a, st) : dispatch_tc<0, false>(a, st);
}

int conv3x3_dgrad_tc_main_prepped(const float* dy, const void* wprep, float* dx, int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
    TcArgs a{};
    a.in = dy; a.wprep = static_cast<const float*>(wprep); a.out = dx;
    a.B = B; a.K = Cout; a.O = Cin; a.H = H; a.W = W; a.accumulate = accumulate ? 1 : 0;
    return dispatch_tc<1, false>(a, st);
}

}  // namespace sifnn

extern "C" int sifnn_conv3x3_fwd_tc(const float* in, const float* in_scale, const float* in_shift, const float* w, const float* bias,
                                    float* out, double* stats, void* wprep, int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(in && w && out && wprep, "conv3x3_fwd_tc: null pointer");
    SIFNN_REQUIRE(Cout <= 64, "conv3x3_fwd_tc: the forward form supports Cout in {16,32,64}");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_fwd_tc: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(sifnn_conv3x3_tc_supported(Cin, Cout, H, W) && B > 0 && B <= 65535, "conv3x3_fwd_tc: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    const sifnn::TcPrepJob job = sifnn::tc_prep_job_fwd(w, wprep, Cin, Cout, W);
    SIFNN_TRY(sifnn::tc_prep_many(&job, 1, st));
    return sifnn::conv3x3_fwd_tc_prepped(in, in_scale, in_shift, wprep, bias, out, stats, B, Cin, Cout, H, W, st);
}

// Main (zero-padded, transposed) part of the data gradient; the caller adds the replicate-padding border terms.
extern "C" int sifnn_conv3x3_dgrad_tc_main(const float* dy, const float* w, float* dx, int accumulate, void* wprep, int B, int Cin, int Cout,
                                           int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dy && w && dx && wprep, "conv3x3_dgrad_tc: null pointer");
    SIFNN_REQUIRE(sifnn_conv3x3_tc_supported(Cout, Cin, H, W) && B > 0 && B <= 65535, "conv3x3_dgrad_tc: unsupported shape");
    cudaStream_t st = sifnn::as_stream(stream);
    const sifnn::TcPrepJob job = sifnn::tc_prep_job_dgrad(w, wprep, Cin, Cout, W);
    SIFNN_TRY(sifnn::tc_prep_many(&job, 1, st));
    return sifnn::conv3x3_dgrad_tc_main_prepped(dy, wprep, dx, accumulate, B, Cin, Cout, H, W, st);
}
