// Weight gradient of the 3x3 replicate-padded convolution on the tcgen05 tensor cores
// (autograd of model.py:135; loss.backward() at train_model_B_gradFTM.py:119), fp32-accurate
// through the same 3-term TF32 split as conv3x3_tc.cu.
//
//   dW[o][c][ky][kx] = sum_{b,y,x} dy[b][o][y][x] * act(in)[b][c][clamp(y+ky-1)][clamp(x+kx-1)]
//
// The reduction runs over PIXELS, so pixels are the MMA's K dimension and both operands are
// MN-major.  The only shared-memory layout the tensor core accepts for MN-major 32-bit operands
// is SWIZZLE_128B_BASE32B (tools/mn_major_probe.cu; the un-swizzled MN-major form silently
// returns zeros): one k is a 128-byte line of 32 consecutive M (N) indices, four lines form a
// 512-byte atom whose 32-byte chunks are XOR-permuted by the line index, the next four k are SBO
// bytes away and the next 32 M (N) indices LBO bytes away.  The staged tiles are therefore
// pixel-major, T[row][column][32 channels] with the chunk swizzle, a K step is 8 consecutive
// columns of one image row (two atoms, SBO = 512 B) and the M (N) groups of 32 are tile rows
// (LBO = WT * 128 B):
//
//   D[M = (ky', c)][N = (s, kx, o)] += A[(ky', c)][8 px] * B[(s, kx, o)][8 px]        per dy row and 8-column block
//
// * ky is folded into M for free: with 32 input channels per CTA the four M groups of an M = 128 MMA are the pixel
//   lines of four successive tile rows (ky' = 0..3) from the start address of input row y-1.  With 16 channels a
//   pixel line holds (row r : 16 ch | row r+1 : 16 ch), each row being stored twice, and the two groups of an M = 64
//   MMA are two tile rows apart.  ky' = 3 is a junk row block; three quarters of the MMA is useful.
// * kx is folded into N through three column-shifted copies of the dy tile (dy is the narrow operand), and the
//   replicate padding along x folds into those copies: the clamped reads x = -1 -> 0 and x = W -> W-1 mean that input
//   column 0 also meets dy[0] under kx = 0 and input column W-1 also meets dy[W-1] under kx = 2, so
//   copy0[x] = dy[x+1] (+ dy[0] at x = 0), copy1[x] = dy[x], copy2[x] = dy[x-1] (+ dy[W-1] at x = W-1); K then runs
//   over exactly the W image columns and the input tile needs no halo columns.  Replicate padding along y is a
//   clamped row index when the input tile is staged.
// * 3-term split: s = (hi, lo) of dy is stacked along N as well.  MMA 1: a_hi x [dy_hi ; dy_lo] (N = 6*Cout) writes
//   a_hi*dy_hi into columns [0, 3*Cout) and a_hi*dy_lo into [3*Cout, 6*Cout); MMA 2: a_lo x dy_hi (N = 3*Cout)
//   accumulates into [3*Cout, 6*Cout).  The big accumulator only ever sees hi*hi products.
// * Tensor-core fp32 accumulation truncates (about -6e-8 relative per accumulate step, tools/tc_accuracy_probe.py);
//   the accumulation chain of a weight gradient is thousands of steps long, so the K blocks rotate over NSETS
//   independent TMEM accumulator sets that are summed in fp32 at the end.
// * Persistent, warp-specialised: a loader warp (two TMA box loads per tile: input rows and dy rows, out-of-image
//   elements zero-filled), eight transformer warps (BatchNorm+ReLU on load, hi/lo split, shifted dy copies), one MMA
//   thread; the accumulators stay in TMEM for the whole kernel and are drained once.  Per-CTA partial sums are
//   reduced in a fixed order by wgrad_reduce (deterministic, no atomics).
#include "tc_common.cuh"

namespace {

using namespace sifnn_tc;

struct WtcArgs {
    const float* in_scale;
    const float* in_shift;
    float* partial;  // [S][O][K][9]
    int B, K, O, H, W;
    int tiles_x, tiles_y, num_tiles;
};

constexpr int WTC_LOAD_WARP = 0;
constexpr int WTC_MMA_WARP = 1;
constexpr int WTC_XF_WARP0 = 2;
#ifndef SIFNN_WTC_XFW
#define SIFNN_WTC_XFW 8
#endif
constexpr int WTC_XFW = SIFNN_WTC_XFW;              // transformer warps
constexpr int WTC_XF_THREADS = WTC_XFW * 32;
constexpr int WTC_THREADS = (WTC_XF_WARP0 + WTC_XFW) * 32;

// KC = input channels per CTA (16 or 32; M = 4*KC), N = Cout, R = dy rows per tile, WT = columns per tile
template <int KC, int N, int R, int WT>
struct WtcSmem {
    static constexpr int M = 4 * KC;
    static constexpr int Q = KC / 4, OQ = N / 4;
    static constexpr int G = 6 * N / 32;                     // 32-wide N groups of the dy tile
    static constexpr int TROWS = R + 2;
    static constexpr int DYW = WT + 8;                       // staged dy row: gx = x0-4 .. x0+WT+3
    static constexpr int A_TILE = TROWS * WT * 32;           // floats per hi (or lo) tile  [TROWS][WT][32 ch] (KC = 16: row r | row r+1)
    static constexpr int B_TILE = R * G * WT * 32;           // floats                      [R][G][WT][32]: n = (s*3 + kx)*N + o = 32 g + slot
    static constexpr int STAGE = 2 * A_TILE + B_TILE;
    static constexpr int RAW_X = KC * TROWS * WT;            // [TROWS][KC][WT]   (channels of a row WT floats apart)
    static constexpr int RAW_DY = N * R * DYW;               // [R][N][DYW]
    static constexpr int RAW_STAGE = RAW_X + RAW_DY;
    static constexpr int RAW_STAGES = 2;
    static constexpr int CTRL_FLOATS = 256;                  // barriers, TMEM slot, BatchNorm scale/shift of the KC channels
    static constexpr int BUDGET = 223 * 1024;
    static constexpr int REST = BUDGET - CTRL_FLOATS * 4 - RAW_STAGES * RAW_STAGE * 4;
    static constexpr int STAGES = (STAGE * 4 * 3 <= REST) ? 3 : 2;
    static constexpr size_t BYTES = (size_t)(STAGES * STAGE + RAW_STAGES * RAW_STAGE + CTRL_FLOATS) * 4 + 1024;  // + slack to align the base to 1 KB
    static constexpr int SET_COLS = 6 * N;                   // [a_hi*dy_hi : 3N][cross terms : 3N]
    static constexpr int NSETS = (512 / SET_COLS) >= 4 ? 4 : (512 / SET_COLS);
    static constexpr int KBLOCKS = R * (WT / 8);             // K steps (dy row, 8-column block) per tile
    static_assert(M == 64 || M == 128, "UMMA M");
    static_assert(NSETS >= 1, "at least one accumulator set");
    static_assert(STAGE * 4 * 2 <= REST, "two pipeline stages must fit shared memory");
    static_assert((A_TILE * 4) % 1024 == 0 && (B_TILE * 4) % 1024 == 0, "swizzled tiles must stay 1 KB aligned");
    static_assert((RAW_X * 4) % 128 == 0 && (RAW_STAGE * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(WT % 8 == 0 && (3 * N) % 16 == 0 && (6 * N) % 32 == 0, "MMA shape");
};

template <int KC, int N, int R, int WT, bool AFFINE>
__global__ void __launch_bounds__(WTC_THREADS, 1) wgrad_tc_kernel(const WtcArgs a, const __grid_constant__ CUtensorMap tmap_x,
                                                                 const __grid_constant__ CUtensorMap tmap_dy) {
    using SM = WtcSmem<KC, N, R, WT>;
    constexpr int S = SM::STAGES, RS = SM::RAW_STAGES, TROWS = SM::TROWS, Q = SM::Q, OQ = SM::OQ, DYW = SM::DYW, M = SM::M, G = SM::G;
    extern __shared__ __align__(128) float smem_raw[];
    float* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;   // the chunk swizzle is a function of the address bits
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* ab_full = bars;           // [3]
    uint64_t* ab_empty = bars + 4;      // [3]
    uint64_t* raw_full = bars + 8;      // [2]
    uint64_t* raw_empty = bars + 12;    // [2]
    uint64_t* acc_full = bars + 16;     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    float* sc_s = smem + 64;            // [KC <= 32]
    float* sh_s = smem + 96;            // [KC <= 32]
    float* stage0 = smem + SM::CTRL_FLOATS;
    float* raw0 = stage0 + (size_t)S * SM::STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W;
    const int c0 = blockIdx.y * KC;          // this CTA's input-channel chunk
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(ab_full + s, WTC_XF_THREADS); mbar_init(ab_empty + s, 1); }
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, WTC_XF_THREADS); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == WTC_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (AFFINE && tid < KC) { sc_s[tid] = __ldg(a.in_scale + c0 + tid); sh_s[tid] = __ldg(a.in_shift + c0 + tid); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int& b, int& y0, int& x0) {
        b = tile / tiles_per_img;
        const int t = tile - b * tiles_per_img;
        const int ty = t / a.tiles_x;
        y0 = ty * R;
        x0 = (t - ty * a.tiles_x) * WT;
    };

    if (warp == WTC_LOAD_WARP) {
        // ======================= loader: input box (WT cols x R+2 rows x KC planes) + dy box (WT+8 cols x R rows x N planes) =======================
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++g) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int rs = g % RS;
            if (g >= RS) mbar_wait(raw_empty + rs, ((g / RS) - 1) & 1);
            if (lane == 0) {
                float* dst = raw0 + (size_t)rs * SM::RAW_STAGE;
                mbar_arrive_expect_tx(raw_full + rs, SM::RAW_STAGE * 4);
                tma_load_3d(dst, &tmap_x, x0, b * a.K + c0, y0 - 1, raw_full + rs);
                tma_load_3d(dst + SM::RAW_X, &tmap_dy, x0 - 4, b * a.O, y0, raw_full + rs);
            }
            __syncwarp();
        }
    } else if (warp == WTC_MMA_WARP) {
        // ======================= MMA issuer (one thread) =======================
        constexpr uint32_t SBO = 512, LBO_A = (KC == 32 ? 1 : 2) * WT * 128, LBO_B = WT * 128;
        constexpr uint64_t SW = 1ull << 61;                           // layout type 1: SWIZZLE_128B_BASE32B
        constexpr bool SPLIT_N = (6 * N > 256);                       // Cout = 64: hi and lo halves of dy as two MMAs
        constexpr uint32_t idesc1 = make_idesc_mn(M, SPLIT_N ? 3 * N : 6 * N);
        constexpr uint32_t idesc2 = make_idesc_mn(M, 3 * N);
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++g) {
            const int s = g % S;
            mbar_wait(ab_full + s, (g / S) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(stage0 + (size_t)s * SM::STAGE);
                const uint64_t da_hi = make_desc(a_hi, LBO_A, SBO) | SW;
                const uint64_t da_lo = da_hi + (uint64_t)(SM::A_TILE * 4 / 16);
                const uint64_t db = (make_desc(a_hi, LBO_B, SBO) | SW) + (uint64_t)(2 * SM::A_TILE * 4 / 16);
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int j = 0; j < WT / 8; ++j) {
                        // one accumulator set per TILE (rotating over the sets from tile to tile): switching the accumulator costs the tensor
                        // core ~265 clocks, a follow-up MMA into the same one ~120 (tools/mma_cost_probe.cu), so the set changes 8x less often
                        // than with a per-K-block rotation while the chain per set stays total / NSETS
                        const int kb = r * (WT / 8) + j;
                        const uint32_t d = tmem_base + (uint32_t)((g % SM::NSETS) * SM::SET_COLS);
                        const uint32_t acc = (g < SM::NSETS && kb == 0) ? 0u : 1u;
                        const uint64_t oa = (uint64_t)((r * WT + 8 * j) * 8);          // 16-byte units (a pixel line is 128 B): tile row r = input row y-1 of dy row r
                        const uint64_t ob = (uint64_t)((r * G * WT + 8 * j) * 8);
                        if (SPLIT_N) {
                            umma_tf32(d, da_hi + oa, db + ob, idesc1, acc);
                            umma_tf32(d + 3 * N, da_hi + oa, db + ob + (uint64_t)((G / 2) * WT * 8), idesc1, acc);
                        } else {
                            umma_tf32(d, da_hi + oa, db + ob, idesc1, acc);
                        }
                        umma_tf32(d + 3 * N, da_lo + oa, db + ob, idesc2, 1u);
                    }
                }
                umma_commit(ab_empty + s);
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(acc_full);
        __syncwarp();
    } else {
        // ======================= transformers =======================
        // Work items are (row, 8-channel chunk, half X, column): lane bits = column & 3, X, column >> 2; the warp index and the iteration give the
        // chunk and the row, so column, X and chunk -- hence the four channels, their BatchNorm affine and every address term but the row --
        // are per-thread constants.  The slot <-> channel assignment inside a chunk is chosen for the shared-memory banks:
        //   input:  slot 8*Q2 + 4*X + e  <->  channel 8*Q2 + 2*e + X             (raw rows are [channel][16 cols]: X shifts the bank by 16)
        //   dy:     slot 8*Q2 + 4*X + e  <->  channel 8*Q2 + (e & 1) + 4*(e >> 1) + 2*X   (raw rows are [channel][24 cols]: +2 channels = +16 banks)
        // so both the scalar reads (16 columns x 2 halves = 32 banks) and the 16-byte stores (8 lanes of a wavefront = 4 swizzled chunks x 2
        // halves) are conflict-free.  The drain undoes the permutation.
        const int xt = tid - WTC_XF_WARP0 * 32;
        const int xw = xt >> 5;
        const int col = (xt & 3) | (((xt >> 3) & 3) << 2);
        const int X = (xt >> 2) & 1;
        static_assert(WT == 16, "lane mapping of the transformers assumes 16-column tiles");
        constexpr int ITEMS_A = TROWS * Q * WT;   // (row, channel quad, column)
        constexpr int NIT_A = (ITEMS_A + WTC_XF_THREADS - 1) / WTC_XF_THREADS;
        constexpr int ITEMS_B = R * OQ * WT;
        constexpr int NIT_B = (ITEMS_B + WTC_XF_THREADS - 1) / WTC_XF_THREADS;
        constexpr int QH = Q / 2, OQH = OQ / 2;   // 8-channel chunks
        const int Q2 = xw % QH, rrA0 = xw / QH;
        const int O2 = xw % OQH, rB0 = xw / OQH;
        float sc[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { sc[e] = AFFINE ? sc_s[8 * Q2 + 2 * e + X] : 1.f; sh[e] = AFFINE ? sh_s[8 * Q2 + 2 * e + X] : 0.f; }
        const int a_src = (8 * Q2 + X) * WT + col;                                       // + rj * KC * WT + 2 * e * WT
        const int a_dst = col * 32 + ((Q2 ^ (col & 3)) << 3) + 4 * X;                    // + rr * WT * 32
        const int a_dst2 = col * 32 + (((2 + Q2) ^ (col & 3)) << 3) + 4 * X;             // KC = 16: slots 16..31 of the line above
        const int b_src = (8 * O2 + 2 * X) * DYW + 4 + col;                              // + r * N * DYW + ((e & 1) + 4 * (e >> 1)) * DYW
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++g) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int s = g % S, rs = g % RS;
            float* a_hi = stage0 + (size_t)s * SM::STAGE;
            float* a_lo = a_hi + SM::A_TILE;
            float* bt = a_lo + SM::A_TILE;
            const float* rawx = raw0 + (size_t)rs * SM::RAW_STAGE;
            const float* rawd = rawx + SM::RAW_X;
            if (g >= S) mbar_wait(ab_empty + s, ((g / S) - 1) & 1);
            mbar_wait(raw_full + rs, (g / RS) & 1);
            // ---- input tile: pixel lines [row][col][32 floats], 32-byte chunks XOR-permuted by (col & 3); rows clamped onto the image
            //      (replicate padding along y)
#pragma unroll
            for (int i = 0; i < NIT_A; ++i) {
                const int rr = rrA0 + i * (WTC_XFW / QH);
                if ((ITEMS_A % WTC_XF_THREADS == 0 && WTC_XFW == 8) || rr < TROWS) {
                    const int rj = min(max(y0 + rr - 1, 0), H - 1) - (y0 - 1);
                    const float* src = rawx + rj * (KC * WT) + a_src;
                    float hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = src[2 * e * WT];
                        if (AFFINE) t = sifnn::act_affine_relu(t, sc[e], sh[e]);
                        hi[e] = tf32_hi(t);
                        lo[e] = t - hi[e];
                    }
                    const float4 vh = make_float4(hi[0], hi[1], hi[2], hi[3]), vl = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4*>(a_hi + rr * (WT * 32) + a_dst) = vh;
                    *reinterpret_cast<float4*>(a_lo + rr * (WT * 32) + a_dst) = vl;
                    if (KC == 16 && rr >= 1) {  // second copy: slots 16..31 of the line of the row above
                        *reinterpret_cast<float4*>(a_hi + (rr - 1) * (WT * 32) + a_dst2) = vh;
                        *reinterpret_cast<float4*>(a_lo + (rr - 1) * (WT * 32) + a_dst2) = vl;
                    }
                }
            }
            // ---- dy tile: [row][group][col][32 slots], n = (s*3 + kx)*N + o-slot: three column-shifted copies (replicate padding along x folded in), hi | lo
            const int gx = x0 + col;
#pragma unroll
            for (int i = 0; i < NIT_B; ++i) {
                const int r = rB0 + i * (WTC_XFW / OQH);
                if ((ITEMS_B % WTC_XF_THREADS == 0 && WTC_XFW == 8) || r < R) {
                    const float* src = rawd + r * (N * DYW) + b_src;
                    float v[3][4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float* p = src + ((e & 1) + 4 * (e >> 1)) * DYW;
                        const float dm = p[-1], dc = p[0], dp = p[1];
                        v[0][e] = (gx == 0) ? dp + dc : dp;
                        v[1][e] = dc;
                        v[2][e] = (gx == W - 1) ? dm + dc : dm;
                    }
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        float hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { hi[e] = tf32_hi(v[kx][e]); lo[e] = v[kx][e] - hi[e]; }
#pragma unroll
                        for (int sp = 0; sp < 2; ++sp) {
                            const int n0 = (sp * 3 + kx) * N;   // compile-time; + 8 * O2 + 4 * X
                            const int n = n0 + 8 * O2;
                            const int o = ((r * G + (n >> 5)) * WT + col) * 32 + ((((n & 31) >> 3) ^ (col & 3)) << 3) + 4 * X;
                            *reinterpret_cast<float4*>(bt + o) = sp ? make_float4(lo[0], lo[1], lo[2], lo[3]) : make_float4(hi[0], hi[1], hi[2], hi[3]);
                        }
                    }
                }
            }
            mbar_arrive(raw_empty + rs);
            fence_proxy_async();
            mbar_arrive(ab_full + s);
        }
    }

    // ======================= drain: TMEM -> fp32 sum over the accumulator sets -> per-CTA partial =======================
    __syncthreads();
    if (warp < 8) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        // M = 128: row m <-> TMEM lane m;  M = 64: row m <-> lane (m % 16) + 32 * (m / 16).  With m = ky' * KC + c both give
        // ky' = lane quadrant, c = lane within the quadrant.
        const int quad = warp & 3, half = warp >> 2;
        const bool ok = (quad < 3) && (lane < KC);
        const int my_tiles = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int sets_used = my_tiles < SM::NSETS ? my_tiles : SM::NSETS;
        const int cch = (lane & ~7) + 2 * (lane & 3) + ((lane >> 2) & 1);   // input slot -> channel
        constexpr int NB = 3 * N / 8;           // 8-column blocks of one accumulator half
#pragma unroll 1
        for (int nb = half; nb < NB; nb += 2) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int set = 0; set < SM::NSETS; ++set) {
                if (set >= sets_used) break;   // a CTA with fewer tiles than sets never touched the others
                float d1[8], d2[8];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + set * SM::SET_COLS + nb * 8;
                tmem_ld8(taddr, d1);
                tmem_ld8(taddr + 3 * N, d2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += d1[j] + d2[j];
            }
            if (ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = nb * 8 + j;       // n = kx * N + o
                    const int kx = n / N, os = n - kx * N;
                    const int o = (os & ~7) + (os & 1) + 4 * ((os >> 1) & 1) + 2 * ((os >> 2) & 1);   // dy slot -> channel
                    a.partial[(((size_t)blockIdx.x * a.O + o) * a.K + c0 + cch) * 9 + quad * 3 + kx] = acc[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WTC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// out[i] = sum_s partial[s][i], fixed summation order (deterministic): 32 slice-strided partial sums per output, combined in order.
__global__ void __launch_bounds__(1024) wgrad_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int S) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (i < n)
        for (int s = ty; s < S; s += 32) acc += __ldg(partial + (size_t)s * n + i);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && i < n) {
        float t = red[0][tx];
#pragma unroll
        for (int j = 1; j < 32; ++j) t += red[j][tx];
        out[i] = t;
    }
}

template <int KC, int N, int R, int WT>
int launch_wtc(const float* in, const float* dy, WtcArgs a, int& S, bool affine, cudaStream_t st) {
    using SM = WtcSmem<KC, N, R, WT>;
    auto k_aff = wgrad_tc_kernel<KC, N, R, WT, true>;
    auto k_pln = wgrad_tc_kernel<KC, N, R, WT, false>;
    static sifnn::PerDeviceOnce attr_once;   // the attribute is per device: one flag per device, not one per process
    if (attr_once.first_time()) {
        SIFNN_CUDA(cudaFuncSetAttribute(k_aff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
        SIFNN_CUDA(cudaFuncSetAttribute(k_pln, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
    }
    a.tiles_x = a.W / WT;
    a.tiles_y = (a.H + R - 1) / R;
    a.num_tiles = a.B * a.tiles_x * a.tiles_y;
    CUtensorMap tx, td;
    SIFNN_REQUIRE(encode_rows_of_planes_map(&tx, in, a.W, a.H, (long long)a.B * a.K, WT, KC, R + 2) &&
                      encode_rows_of_planes_map(&td, dy, a.W, a.H, (long long)a.B * a.O, WT + 8, N, R),
                  "conv3x3_wgrad_tc: cuTensorMapEncodeTiled is unavailable or failed");
    if (S > a.num_tiles) S = a.num_tiles;   // the caller reduces over the S partial slices actually written
    dim3 grid(S, a.K / KC);
    if (affine) k_aff<<<grid, WTC_THREADS, SM::BYTES, st>>>(a, tx, td);
    else k_pln<<<grid, WTC_THREADS, SM::BYTES, st>>>(a, tx, td);
    return sifnn::check_launch("wgrad_tc_kernel");
}

int wtc_kc(int Cin) { return Cin == 16 ? 16 : 32; }
int wtc_S(int B, int Cin, int H, int W) {
    const int chunks = Cin / wtc_kc(Cin);
    int S = sifnn::num_sms() / chunks;
    if (S < 1) S = 1;
    return S;
}

}  // namespace

extern "C" int sifnn_conv3x3_wgrad_tc_supported(int Cin, int Cout, int H, int W) {
    return (Cin == 16 || (Cin % 32 == 0 && Cin <= 256)) && (Cout == 16 || Cout == 32 || Cout == 64) && (W % 16 == 0) && W >= 16 && H >= 1;
}

extern "C" size_t sifnn_conv3x3_wgrad_tc_workspace(int B, int Cin, int Cout, int H, int W) {
    if (B <= 0 || !sifnn_conv3x3_wgrad_tc_supported(Cin, Cout, H, W)) return 0;
    return (size_t)wtc_S(B, Cin, H, W) * Cout * Cin * 9 * sizeof(float);
}

namespace sifnn {
int wgrad_tc_partials(const float* in, const float* in_scale, const float* in_shift, const float* dy, void* workspace, int B, int Cin, int Cout, int H, int W,
                      cudaStream_t st, int* slots_out) {
    SIFNN_REQUIRE(in && dy && workspace, "conv3x3_wgrad_tc: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_wgrad_tc: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(B > 0 && B <= 65535 && sifnn_conv3x3_wgrad_tc_supported(Cin, Cout, H, W), "conv3x3_wgrad_tc: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin,
                  Cout, H, W);
    SIFNN_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "conv3x3_wgrad_tc: tensors must be 16-byte aligned");
    WtcArgs a{};
    a.in_scale = in_scale; a.in_shift = in_shift; a.partial = static_cast<float*>(workspace);
    a.B = B; a.K = Cin; a.O = Cout; a.H = H; a.W = W;
    const bool affine = in_scale != nullptr;
    int S = wtc_S(B, Cin, H, W);
    const int kc = wtc_kc(Cin);
    int rc = SIFNN_EINVAL;
    if (kc == 16) {
        switch (Cout) {
            case 16: rc = launch_wtc<16, 16, 4, 16>(in, dy, a, S, affine, st); break;
            case 32: rc = launch_wtc<16, 32, 4, 16>(in, dy, a, S, affine, st); break;
            case 64: rc = launch_wtc<16, 64, 2, 16>(in, dy, a, S, affine, st); break;
        }
    } else {
        switch (Cout) {
            case 16: rc = launch_wtc<32, 16, 4, 16>(in, dy, a, S, affine, st); break;
            case 32: rc = launch_wtc<32, 32, 4, 16>(in, dy, a, S, affine, st); break;
            case 64: rc = launch_wtc<32, 64, 2, 16>(in, dy, a, S, affine, st); break;
        }
    }
    SIFNN_TRY(rc);
    *slots_out = S;
    return 0;
}
}  // namespace sifnn

extern "C" int sifnn_conv3x3_wgrad_tc(const float* in, const float* in_scale, const float* in_shift, const float* dy, float* dw, void* workspace,
                                      int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dw, "conv3x3_wgrad_tc: null pointer");
    cudaStream_t st = sifnn::as_stream(stream);
    int S = 0;
    SIFNN_TRY(sifnn::wgrad_tc_partials(in, in_scale, in_shift, dy, workspace, B, Cin, Cout, H, W, st, &S));
    const int n = Cout * Cin * 9;
    wgrad_tc_reduce_kernel<<<(n + 31) / 32, 1024, 0, st>>>(static_cast<const float*>(workspace), dw, n, S);
    return sifnn::check_launch("wgrad_tc_reduce_kernel");
}
