// Weight gradient of the 3x3 convolution on the tensor cores with 16-bit K-major operands (round 2).
//
//   dW[o][c][ky][kx] = sum over images, rows y, columns x of  dy[o][y][x] * X[c][cy(y + ky - 1)][cx(x + kx - 1)]        X = act(in), c* = clamp
//
// The reduction runs over pixels, so pixels are the MMA's K dimension (16 per instruction with 16-bit operands) and BOTH operands are K-major
// ([8-pixel octet][row][8 x 2 bytes]; a row of the MMA is one 16-byte unit, rows are contiguous).  The round-1 kernel (wgrad_tc.cu: TF32, 8 pixels
// per MMA, two MMAs per K step, 16-column tiles) is bound by the ~111-clock floor of a small tcgen05.mma (profiles/r2a_mma_probe2.log) --
// 32 MMAs per 128 pixels of a 16-channel layer.  Here the same 128 pixels take 8:
//
//   M side  X window:  rows (tile row r, split s, channel c) of THREE consecutive staged input rows -- ky is a start address into the tile
//                      (16 channels: one M = 96 window holds hi and lo; 32 channels: a hi window and a lo window accumulate into the same columns)
//   N side  dy copies: rows (kx, split s', channel o) -- three column-shifted copies of the dy row, the replicate padding along x folded into
//                      the edge elements (copy kx=0 [x = 0] += dy[0], copy kx=2 [x = W-1] += dy[W-1]); N = 6 * C_out
//   D[(r, s, c)][(kx, s', o)] += X . dy          one MMA per 16 pixels computes all nine taps and all four hi/lo products
//
// Operand formats: both BF16 hi + lo (16 significant bits, full fp32 exponent range: gradients are tiny).  FP16 for the activations would give 22 bits,
// but kind::f16 rejects mixed FP16 x BF16 operands on sm_100a (illegal instruction, tools/km_fmt_probe.py) and FP16 gradients would need a scale.  Accumulators stay in TMEM for the whole kernel, rotating over
// NSETS independent sets (the tensor core's fp32 accumulation truncates: shorter chains, and two MMA issuers can alternate tiles without ever
// sharing a set); one drain at the end writes a per-CTA partial that a fixed-order reduction sums (deterministic).
// More than 32 channels on either side are split over blockIdx.y into 32-channel blocks.
//
// Warp roles (11 warps): 0 TMA loader, 1 and 2 MMA issuers (even / odd tiles), 3..10 transformers (raw fp32 -> BatchNorm affine + ReLU -> hi / lo
// split -> operand tiles), which also drain the accumulators at the end.
//
// What bounds it (profiles/r2_ncu_full_wgrad_km_16x16x256_summary.csv, r2w/r2za_profile_wgrad.log): shared-memory bandwidth.  Per 8 x 32 tile the
// transformers read 45 KB of raw fp32 and write 69 KB of operands, the MMAs read another ~80 KB; ncu shows the shared-memory data pipe 37 % (tensor
// core reads) + 45 % (LSU) busy, DRAM traffic = algorithmic bytes, tensor pipe 32 %.  Tried and measured: pairing two dy rows per MMA (N = 192, half
// the MMAs: kept, no time change), 16 instead of 8 transformer warps (no change), 4-row tiles with four raw stages (slower: 84 -> 107 us), a raw
// ring of up to four stages where it fits (32 -> 16: 137 -> 131 us, 32 x 32 blocks: 131 -> 118 us; kept).
#include "tc_common.cuh"

#include <cstdlib>

namespace {

using namespace sifnn_tc;

constexpr int KM_WT = 32;                     // columns per tile: two K = 16 steps per row
constexpr int KM_XW = KM_WT + 4;              // raw X box width: 36 floats -> channel pitch 36 words, conflict-free 16-byte reads across channels
constexpr int KM_DW = KM_WT + 12;             // raw dy box width: x0 - 4 .. x0 + 39 (44 words pitch, conflict-free as well)
constexpr int KM_LOAD_WARP = 0, KM_MMA_WARP = 1, KM_MMA_WARP2 = 2, KM_XF_WARP0 = 3, KM_XF_WARPS = 8;
constexpr int KM_XF_THREADS = KM_XF_WARPS * 32;
constexpr int KM_THREADS = (KM_XF_WARP0 + KM_XF_WARPS) * 32;

struct KmArgs {
    const float* in_scale;
    const float* in_shift;
    float* partial;          // [gridDim.x * PAIR][O][K][9]
    int B, K, O, H, W;
    int tiles_x, tiles_y, num_tiles;
    int nco;                 // 32-channel (or 16-channel) blocks of the output channels: blockIdx.y = ci_block * nco + co_block
    int fmt_x, fmt_dy;       // 0 FP16, 1 BF16
};

template <int CI, int CO, int R>
struct KmCfg {
    static constexpr int TR = R + 2, G8 = KM_WT / 8;
    static constexpr int XR = TR * 2 * CI;                 // X rows per octet plane: (tile row, hi / lo, channel)
    static constexpr int DR = R * 6 * CO;                  // dy rows per octet plane: (row, kx, hi / lo, channel)
    static constexpr int X_TILE = G8 * XR * 16, D_TILE = G8 * DR * 16, STAGE = X_TILE + D_TILE;
    static constexpr int RAW_X = TR * CI * KM_XW * 4, RAW_D = R * CO * KM_DW * 4, RAW_STAGE = RAW_X + RAW_D;
    static constexpr int PAIR = (CO == 16) ? 2 : 1;        // dy rows per MMA: an MMA costs ~111 clocks up to N = 144 and ~130 at N = 192, so two 16-channel
                                                           // dy rows share one instruction (and one M = 128 window of FOUR input rows)
    static constexpr int N = PAIR * 6 * CO;                // 192
    static constexpr int NSETS = 2;
    static constexpr int MB = CI / 16;                     // MMAs per K step
    static constexpr int RS_FIT = (227 * 1024 - 1024 - 2 * STAGE) / RAW_STAGE;
    static constexpr int RS = RS_FIT >= 4 ? 4 : RS_FIT;    // raw (TMA) stages: the loads of the next tiles are in flight while this one is converted
    static constexpr size_t BYTES = 1024 + 2 * (size_t)STAGE + RS * (size_t)RAW_STAGE;
    static_assert(RS >= 2, "at least two raw stages");
    static_assert(CI == 16 || CI == 32, "16 or 32 input channels per CTA");
    static_assert(N <= 256 && NSETS * N <= 512, "accumulator sets must fit TMEM");
    static_assert(RAW_X % 128 == 0 && RAW_D % 128 == 0 && STAGE % 128 == 0 && X_TILE % 128 == 0, "TMA destinations / tiles stay 128-byte aligned");
    static_assert(BYTES <= 227 * 1024, "shared-memory budget");
};

__device__ __forceinline__ uint32_t km_pack(float lo_elem, float hi_elem, int fmt) {
    uint32_t r;
    if (fmt == 0) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void km_unpack(uint32_t h, float& lo_elem, float& hi_elem, int fmt) {
    if (fmt == 0) {
        asm("{\n\t.reg .f16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(lo_elem), "=f"(hi_elem) : "r"(h));
    } else {
        lo_elem = __uint_as_float(h << 16);
        hi_elem = __uint_as_float(h & 0xffff0000u);
    }
}
// eight fp32 values -> one 16-byte unit of hi parts and one of residuals
__device__ __forceinline__ void km_split8(const float* v, int fmt, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
        h[e >> 1] = km_pack(v[e], v[e + 1], fmt);
        float f0, f1;
        km_unpack(h[e >> 1], f0, f1, fmt);
        l[e >> 1] = km_pack(v[e] - f0, v[e + 1] - f1, fmt);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int CI, int CO, int R, bool AFFINE>
__global__ void __launch_bounds__(KM_THREADS, 1) wgrad_km_kernel(const KmArgs a, const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy) {
    sifnn::pdl_wait_and_trigger();   // launched with launch_pdl: every global access below comes after the previous kernel of the stream
    using C = KmCfg<CI, CO, R>;
    constexpr int TR = C::TR, G8 = C::G8, XR = C::XR, DR = C::DR, N = C::N, NSETS = C::NSETS, MB = C::MB, RS = C::RS;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* ab_full = bars;           // [2]
    uint64_t* ab_empty = bars + 2;      // [2]
    uint64_t* raw_full = bars + 4;      // [RS <= 4]
    uint64_t* raw_empty = bars + 8;     // [RS <= 4]
    uint64_t* acc_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    float* sc_s = reinterpret_cast<float*>(smem + 128);    // [CI]
    float* sh_s = sc_s + 32;
    unsigned char* stage0 = smem + 1024;
    unsigned char* raw0 = stage0 + 2 * C::STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W;
    const int ci0 = (blockIdx.y / a.nco) * CI, co0 = (blockIdx.y % a.nco) * CO;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(ab_full + s, KM_XF_WARPS); mbar_init(ab_empty + s, 1); }
        for (int s = 0; s < RS; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, KM_XF_WARPS); }
        mbar_init(acc_full, 2);
        fence_mbar_init();
    }
    if (warp == KM_MMA_WARP) tmem_alloc(tmem_slot, 512);
    if (AFFINE && tid < CI) { sc_s[tid] = __ldg(a.in_scale + ci0 + tid); sh_s[tid] = __ldg(a.in_shift + ci0 + tid); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int& b, int& y0, int& x0) {
        b = tile / tiles_per_img;
        const int t = tile - b * tiles_per_img;
        const int ty = t / a.tiles_x;
        y0 = ty * R;
        x0 = (t - ty * a.tiles_x) * KM_WT;
    };

    if (warp == KM_LOAD_WARP) {
        // ======================= loader: input box (36 cols x CI planes x R+2 rows) + dy box (44 cols x CO planes x R rows); outside the image = zeros =======================
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++g) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int rs = g % RS;
            if (g >= RS) mbar_wait(raw_empty + rs, ((g / RS) - 1) & 1);
            if (lane == 0) {
                unsigned char* dst = raw0 + (size_t)rs * C::RAW_STAGE;
                mbar_arrive_expect_tx(raw_full + rs, C::RAW_STAGE);
                tma_load_3d(dst, &tmap_x, x0, b * a.K + ci0, y0 - 1, raw_full + rs);
                tma_load_3d(dst + C::RAW_X, &tmap_dy, x0 - 4, b * a.O + co0, y0, raw_full + rs);
            }
            __syncwarp();
        }
    } else if (warp == KM_MMA_WARP || warp == KM_MMA_WARP2) {
        // ======================= MMA issuers (one thread each): even / odd tiles, disjoint accumulator sets =======================
        if (lane == 0) {
            const int mine = (warp == KM_MMA_WARP) ? 0 : 1;
            const uint32_t idesc = (1u << 4) | ((uint32_t)a.fmt_x << 7) | ((uint32_t)a.fmt_dy << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int g = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++g) {
                if ((g & 1) != mine) continue;
                const int s = g & 1;
                mbar_wait_spin(ab_full + s, (g >> 1) & 1);
                tc_fence_after();
                const uint32_t xt = smem_u32(stage0 + (size_t)s * C::STAGE), dt = xt + C::X_TILE;
                const uint32_t d = tmem_base + (uint32_t)((g % NSETS) * N);
                const bool first_use = g < NSETS;
#pragma unroll
                for (int r = 0; r < R; r += C::PAIR) {
#pragma unroll
                    for (int j = 0; j < KM_WT / 16; ++j) {
                        const uint64_t db = make_desc(dt + (uint32_t)((2 * j * DR + r * 6 * CO) * 16), DR * 16, 128);
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
                            // CI = 16: rows (r, s, c), window = 3 (4 with paired dy rows) tile rows x 32;  CI = 32: rows (s, r, c), hi window then lo window
                            const int row0 = (CI == 16) ? r * 32 : (mb * TR + r) * 32;
                            const uint64_t da = make_desc(xt + (uint32_t)((2 * j * XR + row0) * 16), XR * 16, 128);
                            umma_bf16(d, da, db, idesc, (first_use && r == 0 && j == 0 && mb == 0) ? 0u : 1u);
                        }
                    }
                }
                umma_commit(ab_empty + s);
            }
            umma_commit(acc_full);
        }
        __syncwarp();
    } else {
        // ======================= transformers: thread = (row, octet, channel), channel fastest =======================
        const int xt = tid - KM_XF_WARP0 * 32;
        constexpr int XI = TR * G8 * CI, DI = R * G8 * CO;
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++g) {
            int b, y0, x0;
            tile_coords(tile, b, y0, x0);
            const int s = g & 1;
            unsigned char* xtile = stage0 + (size_t)s * C::STAGE;
            unsigned char* dtile = xtile + C::X_TILE;
            const int rs = g % RS;
            const float* rawx = reinterpret_cast<const float*>(raw0 + (size_t)rs * C::RAW_STAGE);
            const float* rawd = reinterpret_cast<const float*>(raw0 + (size_t)rs * C::RAW_STAGE + C::RAW_X);
            if (g >= 2) mbar_wait(ab_empty + s, ((g >> 1) - 1) & 1);
            mbar_wait(raw_full + rs, (g / RS) & 1);
            // ---- input: rows clamped onto the image (replicate padding along y); BatchNorm affine + ReLU; hi / lo
#pragma unroll 2
            for (int item = xt; item < XI; item += KM_XF_THREADS) {
                const int c = item % CI, gq = (item / CI) % G8, rr = item / (CI * G8);
                const int rj = min(max(y0 + rr - 1, 0), H - 1) - (y0 - 1);
                const float* src = rawx + (rj * CI + c) * KM_XW + 8 * gq;
                const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                if (AFFINE) {
                    const float sc = sc_s[c], sh = sh_s[c];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = sifnn::act_affine_relu(v[e], sc, sh);
                }
                uint4 hi, lo;
                km_split8(v, a.fmt_x, hi, lo);
                const int row_hi = (CI == 16) ? (rr * 2) * 16 + c : rr * 32 + c;
                const int row_lo = (CI == 16) ? (rr * 2 + 1) * 16 + c : (TR + rr) * 32 + c;
                *reinterpret_cast<uint4*>(xtile + (size_t)(gq * XR + row_hi) * 16) = hi;
                *reinterpret_cast<uint4*>(xtile + (size_t)(gq * XR + row_lo) * 16) = lo;
            }
            // ---- dy: three column-shifted copies (kx = 0: x + 1, kx = 1: x, kx = 2: x - 1) with the replicate padding along x folded in; hi / lo
#pragma unroll 2
            for (int item = xt; item < DI; item += KM_XF_THREADS) {
                const int o = item % CO, gq = (item / CO) % G8, r = item / (CO * G8);
                const float* src = rawd + (r * CO + o) * KM_DW + 4 + 8 * gq;
                const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
                const float vm = src[-1], vp = src[8];
                const float v[10] = {vm, v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, vp};   // v[1 + e] = dy[x0 + 8 gq + e]
                const int gx = x0 + 8 * gq;
                float c0v[8], c2v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { c0v[e] = v[2 + e]; c2v[e] = v[e]; }
                if (gx == 0) c0v[0] += v[1];                 // X[0] also pairs with dy[0] through the clamped tap
                if (gx + 8 == W) c2v[7] += v[8];             // X[W-1] also pairs with dy[W-1]
                const size_t rowbase = (size_t)gq * DR + (size_t)(r * 6) * CO + o;
                uint4 hi, lo;
                km_split8(c0v, a.fmt_dy, hi, lo);
                *reinterpret_cast<uint4*>(dtile + (rowbase + 0 * CO) * 16) = hi;
                *reinterpret_cast<uint4*>(dtile + (rowbase + 1 * CO) * 16) = lo;
                km_split8(v + 1, a.fmt_dy, hi, lo);
                *reinterpret_cast<uint4*>(dtile + (rowbase + 2 * CO) * 16) = hi;
                *reinterpret_cast<uint4*>(dtile + (rowbase + 3 * CO) * 16) = lo;
                km_split8(c2v, a.fmt_dy, hi, lo);
                *reinterpret_cast<uint4*>(dtile + (rowbase + 4 * CO) * 16) = hi;
                *reinterpret_cast<uint4*>(dtile + (rowbase + 5 * CO) * 16) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(raw_empty + rs); mbar_arrive(ab_full + s); }
        }

        // ======================= drain: TMEM -> sum over sets and over the hi / lo blocks -> per-CTA partial =======================
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int quad = warp & 3;                         // TMEM lane quadrant of this warp = input row of the window
        const int half = (warp - KM_XF_WARP0) >> 2;        // the warps of a quadrant split the (kx, 8-channel block) list
        const int my_tiles = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int sets_used = my_tiles < NSETS ? my_tiles : NSETS;
        constexpr int NBLK = 3 * (CO / 8), PAIR = C::PAIR;
        // window row q pairs with dy row p of the MMA as tap ky = q - p; with paired dy rows a tap gets its two halves from two quadrants, which go
        // to two partial slots (2 blockIdx.x + p) that the fixed-order reduction adds
#pragma unroll
        for (int pr = 0; pr < PAIR; ++pr) {
            const int ky = quad - pr;
            if (ky < 0 || ky > 2) continue;
            for (int blk = half; blk < NBLK; blk += KM_XF_WARPS / 4) {
                const int kx = blk / (CO / 8), ob = blk - kx * (CO / 8);
                float acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
                for (int set = 0; set < NSETS; ++set) {
                    if (set >= sets_used) break;
                    float d1[8], d2[8];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + set * N + pr * 6 * CO + (kx * 2) * CO + ob * 8;
                    tmem_ld8(taddr, d1);
                    tmem_ld8(taddr + CO, d2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += d1[j] + d2[j];
                }
                if (CI == 16) {   // lanes 0..15 hold the hi rows of channel c = lane, lanes 16..31 the lo rows
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
                }
                if (CI == 32 || lane < 16) {
                    const int c = (CI == 16) ? (lane & 15) : lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        a.partial[((((size_t)blockIdx.x * PAIR + pr) * a.O + co0 + ob * 8 + j) * a.K + ci0 + c) * 9 + ky * 3 + kx] = acc[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == KM_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// out[i] = sum_s partial[s][i], fixed summation order (deterministic)
__global__ void __launch_bounds__(1024) wgrad_km_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int S) {
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

int g_km_fmt_x = 1, g_km_fmt_dy = 1;   // both BF16: kind::f16 takes ONE format per MMA for both operands on sm_100a in practice (mixed FP16 x BF16 raised an illegal instruction, tools/km_fmt_probe.py); FP16 for both is 10x more accurate but needs a per-tensor scale for the gradients

template <int CI, int CO, int R>
int launch_km(const float* in, const float* dy, KmArgs a, int S, bool affine, cudaStream_t st) {
    using C = KmCfg<CI, CO, R>;
    auto k_aff = wgrad_km_kernel<CI, CO, R, true>;
    auto k_pln = wgrad_km_kernel<CI, CO, R, false>;
    SIFNN_CUDA(cudaFuncSetAttribute(affine ? k_aff : k_pln, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES));
    a.tiles_x = a.W / KM_WT;
    a.tiles_y = a.H / R;
    a.num_tiles = a.B * a.tiles_x * a.tiles_y;
    a.nco = a.O / CO;
    CUtensorMap tx, td;
    SIFNN_REQUIRE(encode_rows_of_planes_map(&tx, in, a.W, a.H, (long long)a.B * a.K, KM_XW, CI, R + 2) &&
                      encode_rows_of_planes_map(&td, dy, a.W, a.H, (long long)a.B * a.O, KM_DW, CO, R),
                  "conv3x3_wgrad_km: cuTensorMapEncodeTiled is unavailable or failed");
    dim3 grid(S, (a.K / CI) * a.nco);
    SIFNN_CUDA(sifnn::launch_pdl(affine ? k_aff : k_pln, grid, dim3(KM_THREADS), (size_t)C::BYTES, st, a, tx, td));
    return sifnn::check_launch("wgrad_km_kernel");
}

int km_ci(int Cin) { return Cin == 16 ? 16 : 32; }
int km_co(int Cout) { return Cout == 16 ? 16 : 32; }
int km_rows16() {   // rows per tile of the 16 -> 16 configuration: 8 (two raw stages fit) or 4 (four raw stages); SIFNN_KM_R16 for A/B runs
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_KM_R16"); v = (e && atoi(e) == 4) ? 4 : 8; }
    return v;
}
int km_rows(int Cin, int Cout) { return (km_ci(Cin) == 16 && km_co(Cout) == 16) ? km_rows16() : ((km_ci(Cin) == 32 && km_co(Cout) == 32) ? 2 : 4); }
bool km_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_WGRAD_KM"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

}  // namespace

extern "C" void sifnn_conv3x3_wgrad_km_config(int fmt_x, int fmt_dy) { g_km_fmt_x = fmt_x ? 1 : 0; g_km_fmt_dy = fmt_dy ? 1 : 0; }

extern "C" int sifnn_conv3x3_wgrad_km_supported(int Cin, int Cout, int H, int W) {
    if (!km_enabled()) return 0;
    if (!((Cin == 16 || Cin % 32 == 0) && Cin <= 256 && (Cout == 16 || Cout % 32 == 0) && Cout <= 256)) return 0;
    return (W % KM_WT == 0) && W >= KM_WT && (H % km_rows(Cin, Cout) == 0);
}

extern "C" size_t sifnn_conv3x3_wgrad_km_workspace(int B, int Cin, int Cout, int H, int W) {
    if (B <= 0 || !sifnn_conv3x3_wgrad_km_supported(Cin, Cout, H, W)) return 0;
    return (size_t)2 * sifnn::num_sms() * Cout * Cin * 9 * sizeof(float);   // two partial slots per CTA when dy rows are paired (16 output channels per CTA)
}

namespace sifnn {

int wgrad_km_partials(const float* in, const float* in_scale, const float* in_shift, const float* dy, void* workspace, int B, int Cin, int Cout, int H, int W,
                      cudaStream_t st, int* slots_out) {
    SIFNN_REQUIRE(in && dy && workspace, "conv3x3_wgrad_km: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_wgrad_km: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(B > 0 && B <= 65535 && sifnn_conv3x3_wgrad_km_supported(Cin, Cout, H, W), "conv3x3_wgrad_km: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin,
                  Cout, H, W);
    SIFNN_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "conv3x3_wgrad_km: tensors must be 16-byte aligned");
    KmArgs a{};
    a.in_scale = in_scale; a.in_shift = in_shift; a.partial = static_cast<float*>(workspace);
    a.B = B; a.K = Cin; a.O = Cout; a.H = H; a.W = W;
    a.fmt_x = g_km_fmt_x; a.fmt_dy = g_km_fmt_dy;
    const bool affine = in_scale != nullptr;
    const int ci = km_ci(Cin), co = km_co(Cout);
    const int blocks = (Cin / ci) * (Cout / co);
    int S = sifnn::num_sms() / blocks;
    if (S < 1) S = 1;
    const int ntiles = B * (W / KM_WT) * (H / km_rows(Cin, Cout));
    if (S > ntiles) S = ntiles;
    int rc;
    if (ci == 16 && co == 16) rc = km_rows16() == 8 ? launch_km<16, 16, 8>(in, dy, a, S, affine, st) : launch_km<16, 16, 4>(in, dy, a, S, affine, st);
    else if (ci == 32 && co == 16) rc = launch_km<32, 16, 4>(in, dy, a, S, affine, st);
    else if (ci == 16 && co == 32) rc = launch_km<16, 32, 4>(in, dy, a, S, affine, st);
    else rc = launch_km<32, 32, 2>(in, dy, a, S, affine, st);
    SIFNN_TRY(rc);
    *slots_out = co == 16 ? 2 * S : S;
    return 0;
}

// out[i] = sum_s partial[s][i] for up to REDUCE_MAX_JOBS (layer) jobs in one launch; block -> job by a prefix table
struct ReduceTable { ReduceJob j[REDUCE_MAX_JOBS]; int first_block[REDUCE_MAX_JOBS + 1]; int njobs; };
__global__ void __launch_bounds__(1024) wgrad_reduce_many_kernel(const ReduceTable t) {
    sifnn::pdl_wait_and_trigger();
    __shared__ float red[32][33];
    int k = 0;
    while (k + 1 < t.njobs && (int)blockIdx.x >= t.first_block[k + 1]) ++k;
    const ReduceJob job = t.j[k];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = ((int)blockIdx.x - t.first_block[k]) * 32 + tx;
    float acc = 0.f;
    if (i < job.n)
        for (int s = ty; s < job.slots; s += 32) acc += __ldg(job.partial + (size_t)s * job.n + i);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && i < job.n) {
        float v = red[0][tx];
#pragma unroll
        for (int r = 1; r < 32; ++r) v += red[r][tx];
        job.out[i] = v;
    }
}

int wgrad_reduce_many(const ReduceJob* jobs, int njobs, cudaStream_t st) {
    if (njobs <= 0) return 0;
    SIFNN_REQUIRE(njobs <= REDUCE_MAX_JOBS, "wgrad_reduce_many: too many jobs (%d)", njobs);
    ReduceTable t{};
    int blocks = 0;
    for (int k = 0; k < njobs; ++k) {
        t.j[k] = jobs[k];
        t.first_block[k] = blocks;
        blocks += (jobs[k].n + 31) / 32;
    }
    t.first_block[njobs] = blocks;
    t.njobs = njobs;
    SIFNN_CUDA(launch_pdl_if(pdl_mode() != 0, wgrad_reduce_many_kernel, dim3(blocks), dim3(1024), (size_t)0, st, t));
    return check_launch("wgrad_reduce_many_kernel");
}

}  // namespace sifnn

extern "C" int sifnn_conv3x3_wgrad_km(const float* in, const float* in_scale, const float* in_shift, const float* dy, float* dw, void* workspace,
                                      int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dw, "conv3x3_wgrad_km: null pointer");
    cudaStream_t st = sifnn::as_stream(stream);
    int S = 0;
    SIFNN_TRY(sifnn::wgrad_km_partials(in, in_scale, in_shift, dy, workspace, B, Cin, Cout, H, W, st, &S));
    const int n = Cout * Cin * 9;
    wgrad_km_reduce_kernel<<<(n + 31) / 32, 1024, 0, st>>>(static_cast<const float*>(workspace), dw, n, S);
    return sifnn::check_launch("wgrad_km_reduce_kernel");
}

extern "C" int sifnn_conv3x3_wgrad_km_partials(const float* in, const float* in_scale, const float* in_shift, const float* dy, void* workspace, int B, int Cin,
                                               int Cout, int H, int W, sifnn_stream_t stream, int* slots) {
    SIFNN_REQUIRE(slots, "conv3x3_wgrad_km_partials: null pointer");
    return sifnn::wgrad_km_partials(in, in_scale, in_shift, dy, workspace, B, Cin, Cout, H, W, sifnn::as_stream(stream), slots);
}

extern "C" int sifnn_wgrad_reduce(const float* partial, float* dw, int n, int slots, sifnn_stream_t stream) {
    SIFNN_REQUIRE(partial && dw && n > 0 && slots > 0, "wgrad_reduce: bad arguments");
    const sifnn::ReduceJob job{partial, dw, n, slots};
    return sifnn::wgrad_reduce_many(&job, 1, sifnn::as_stream(stream));
}
