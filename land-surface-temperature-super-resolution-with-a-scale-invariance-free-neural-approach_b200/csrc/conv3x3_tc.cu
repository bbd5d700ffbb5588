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
#include "common.cuh"

namespace {

constexpr int TC_PITCH = 130;      // 128 pixels + 2 halo columns
constexpr int TC_KC = 8;           // channels per pipeline stage = one MMA K step (TF32: 32 bytes)
constexpr int TC_PRODUCERS = 256;  // threads staging A
constexpr int TC_THREADS = TC_PRODUCERS + 32;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8 rows x 16 B core
// matrices; SBO = byte distance between 8-row groups, LBO = byte distance between the two 16-byte K chunks.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A/B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

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
    int dbg;  // timing experiments only: 1 = skip the MMAs, 2 = issue every MMA twice
};

template <int N, int R>
struct TcSmem {
    static constexpr int TROWS = R + 2;
    static constexpr int A_TILE = 2 * TROWS * TC_PITCH * 4;  // floats per (hi or lo) tile: [2 q][TROWS*PITCH px][4]
    static constexpr int B_TILE = 9 * 2 * 2 * N * 4;          // floats: [9 taps][2 q][2N rows][4]
    static constexpr int STAGE = 2 * A_TILE + B_TILE;
    static constexpr int CTRL_FLOATS = 512;  // barriers, TMEM slot, BatchNorm scale/shift (2 KB)
    static constexpr int BUDGET = (N <= 32) ? 112 * 1024 : 200 * 1024;  // N <= 32: two CTAs per SM
    static constexpr int STAGES = (STAGE * 4 * 4 + 2048 <= BUDGET) ? 4 : ((STAGE * 4 * 3 + 2048 <= BUDGET) ? 3 : 2);
    static constexpr size_t BYTES = (size_t)STAGES * STAGE * 4 + CTRL_FLOATS * 4;
    static constexpr int TMEM_COLS = (R * 2 * N <= 32) ? 32 : (R * 2 * N <= 64) ? 64 : (R * 2 * N <= 128) ? 128 : (R * 2 * N <= 256) ? 256 : 512;
};

template <int N, int R, int PAD, bool AFFINE>
__global__ void __launch_bounds__(TC_THREADS, (N <= 32) ? 2 : 1) conv3x3_tc_kernel(const TcArgs a) {
    using SM = TcSmem<N, R>;
    constexpr int TROWS = SM::TROWS;
    constexpr int S = SM::STAGES;
    extern __shared__ __align__(128) float smem[];
    // control block at the front (2 KB), stages after it
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);       // [S]
    uint64_t* empty_bar = full_bar + 4;                           // [S]
    uint64_t* accum_bar = full_bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 9);
    float* sc_s = smem + 32;   // [<=128]
    float* sh_s = smem + 160;  // [<=128]
    float* stage0 = smem + SM::CTRL_FLOATS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, K = a.K;
    const int tiles_x = W / 128;
    const int x0 = (blockIdx.x % tiles_x) * 128;
    const int y0 = (blockIdx.x / tiles_x) * R;
    const int b = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* in_b = a.in + (size_t)b * K * plane;
    const int nchunks = K / TC_KC;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar + s, TC_PRODUCERS + 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, SM::TMEM_COLS);
    if (AFFINE) {
        for (int i = tid; i < K; i += TC_THREADS) { sc_s[i] = __ldg(a.in_scale + i); sh_s[i] = __ldg(a.in_shift + i); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // =============================== producers: stage A (and kick the weight bulk copy) ===============================
        // Each thread owns NIT fixed (channel quad, tile row, pixel) items.  All 4*NIT global loads of a chunk are issued
        // before anything is consumed, and the loads of chunk ch+1 are in flight while chunk ch is split and stored.
        constexpr int ITEMS = 2 * TROWS * TC_PITCH;
        constexpr int NIT = (ITEMS + TC_PRODUCERS - 1) / TC_PRODUCERS;
        int goff[NIT];   // pixel offset inside a channel plane, -1: zero-filled / unused item
        int cq[NIT];     // channel quad (0/1) of the item
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int it = tid + i * TC_PRODUCERS;
            goff[i] = -1; cq[i] = 0;
            if (it < ITEMS) {
                const int q = it / (TROWS * TC_PITCH);
                const int rem = it - q * (TROWS * TC_PITCH);
                const int rr = rem / TC_PITCH, px = rem - rr * TC_PITCH;
                int gy = y0 + rr - 1, gx = x0 + px - 1;
                bool ok = true;
                if (PAD == 0) {
                    gy = min(max(gy, 0), H - 1);
                    gx = min(max(gx, 0), W - 1);
                } else {
                    ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
                }
                cq[i] = q;
                goff[i] = ok ? gy * W + gx : -1;
            }
        }
        float v[NIT][4], vn[NIT][4];
        auto load_chunk = [&](int ch, float (&dst)[NIT][4]) {
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const float* src = in_b + (size_t)(ch * TC_KC + 4 * cq[i]) * plane + (goff[i] >= 0 ? goff[i] : 0);
#pragma unroll
                for (int e = 0; e < 4; ++e) dst[i][e] = goff[i] >= 0 ? __ldg(src + (size_t)e * plane) : 0.f;
            }
        };
        load_chunk(0, v);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % S;
            float* a_hi = stage0 + (size_t)s * SM::STAGE;
            float* a_lo = a_hi + SM::A_TILE;
            float* b_st = a_lo + SM::A_TILE;
            if (ch + 1 < nchunks) load_chunk(ch + 1, vn);
            if (ch >= S) mbar_wait(empty_bar + s, ((ch / S) - 1) & 1);
            if (tid == 0) {
                mbar_arrive_expect_tx(full_bar + s, SM::B_TILE * 4);
                bulk_g2s(b_st, a.wprep + (size_t)ch * SM::B_TILE, SM::B_TILE * 4, full_bar + s);
            }
            const int c0 = ch * TC_KC;
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int it = tid + i * TC_PRODUCERS;
                if (it < ITEMS) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float t = v[i][e];
                        if (AFFINE && goff[i] >= 0) t = sifnn::act_affine_relu(t, sc_s[c0 + 4 * cq[i] + e], sh_s[c0 + 4 * cq[i] + e]);
                        hi[e] = tf32_hi(t);
                        lo[e] = t - hi[e];
                    }
                    *reinterpret_cast<float4*>(a_hi + (size_t)it * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(a_lo + (size_t)it * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            fence_proxy_async();  // generic-proxy st.shared -> visible to the tensor core (async proxy)
            mbar_arrive(full_bar + s);
#pragma unroll
            for (int i = 0; i < NIT; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[i][e] = vn[i][e];
        }
    } else {
        // =============================== MMA issuer (one thread) ===============================
        constexpr uint32_t idesc1 = make_idesc(128, 2 * N);  // a_hi x [w_hi ; w_lo]
        constexpr uint32_t idesc2 = make_idesc(128, N);      // a_lo x  w_hi
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % S;
            mbar_wait(full_bar + s, (ch / S) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(stage0 + (size_t)s * SM::STAGE);
                const uint32_t a_lo = a_hi + SM::A_TILE * 4;
                const uint32_t b_st = a_lo + SM::A_TILE * 4;
                constexpr uint32_t LBO_A = TROWS * TC_PITCH * 16, LBO_B = 2 * N * 16, SBO = 128;
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int ky = t / 3, kx = t - 3 * ky;
                        const uint32_t aoff = ((r + ky) * TC_PITCH + kx) * 16;
                        const uint64_t db = make_desc(b_st + t * (2 * 2 * N * 16), LBO_B, SBO);
                        const uint32_t d = tmem_base + r * 2 * N;
                        if (a.dbg == 1) continue;
                        umma_tf32(d, make_desc(a_hi + aoff, LBO_A, SBO), db, idesc1, (ch | t) != 0);
                        umma_tf32(d, make_desc(a_lo + aoff, LBO_A, SBO), db, idesc2, 1);
                        if (a.dbg == 2) {
                            umma_tf32(d, make_desc(a_hi + aoff, LBO_A, SBO), db, idesc1, 1);
                            umma_tf32(d, make_desc(a_lo + aoff, LBO_A, SBO), db, idesc2, 1);
                        }
                    }
                }
                umma_commit(empty_bar + s);                       // stage free once these MMAs have read it
                if (ch == nchunks - 1) umma_commit(accum_bar);    // accumulators complete
            }
            __syncwarp();
        }
    }

    // =============================== epilogue: TMEM -> registers -> global ===============================
    if (warp < 8) {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;       // TMEM lane quadrant this warp may read; channel half it handles
        const int x = x0 + quad * 32 + lane;
        constexpr int NH = N / 2;
        float s1[NH], s2[NH];
#pragma unroll
        for (int j = 0; j < NH; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const int y = y0 + r;
#pragma unroll
            for (int n0 = 0; n0 < NH; n0 += 8) {
                const int n = half * NH + n0;
                float d1[8], d2[8];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + r * 2 * N + n;
                tmem_ld8(taddr, d1);
                tmem_ld8(taddr + N, d2);
                tmem_ld_wait();
                if (y < H) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = d1[j] + d2[j];
                        if (a.bias) v += __ldg(a.bias + n + j);
                        float* op = a.out + ((size_t)b * a.O + n + j) * plane + (size_t)y * W + x;
                        if (a.accumulate) v += *op;
                        *op = v;
                        s1[n0 + j] += v;
                        s2[n0 + j] = fmaf(v, v, s2[n0 + j]);
                    }
                }
            }
        }
        if (a.stats) {
            // per-channel sums over this CTA's pixels: lanes -> shuffle; 4 quadrant warps -> shared -> one atomic per channel
            float* red = stage0;  // all MMAs have completed (accum_bar), the stages are dead: [8 warps][NH][2]
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                const float t1 = sifnn::warp_sum(s1[j]), t2 = sifnn::warp_sum(s2[j]);
                if (lane == 0) { red[(warp * NH + j) * 2] = t1; red[(warp * NH + j) * 2 + 1] = t2; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.stats && tid < N) {
        constexpr int NH = N / 2;
        const int half = tid / NH, j = tid - half * NH;
        const float* red = stage0;
        double d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            d1 += (double)red[((half * 4 + q) * NH + j) * 2];
            d2 += (double)red[((half * 4 + q) * NH + j) * 2 + 1];
        }
        atomicAdd(a.stats + tid, d1);
        atomicAdd(a.stats + a.O + tid, d2);
    }
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SM::TMEM_COLS);
    }
}

// Split the weights of one layer into TF32 hi / lo parts in the exact shared-memory image of a stage:
// wprep[chunk][tap][q][row][e], row < N: hi of W[n = row][c = 8*chunk + 4q + e][tap], row >= N: lo of W[n = row - N].
__global__ void tc_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wprep, int K, int N, int w_so, int w_sk, int flip) {
    const int total = (K / 8) * 9 * 2 * 2 * N * 4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int e = idx & 3;
        const int row = (idx >> 2) % (2 * N);
        const int q = ((idx >> 2) / (2 * N)) & 1;
        const int t = ((idx >> 2) / (2 * N * 2)) % 9;
        const int chunk = (idx >> 2) / (2 * N * 2 * 9);
        const int n = row < N ? row : row - N;
        const int c = chunk * 8 + 4 * q + e;
        const float v = __ldg(w + (size_t)n * w_so + (size_t)c * w_sk + (flip ? 8 - t : t));
        const float hi = tf32_hi(v);
        wprep[idx] = row < N ? hi : v - hi;
    }
}

template <int N, int R, int PAD, bool AFFINE>
int launch_tc(const TcArgs& a, cudaStream_t st) {
    using SM = TcSmem<N, R>;
    auto kern = conv3x3_tc_kernel<N, R, PAD, AFFINE>;
    static bool attr_done = false;
    if (!attr_done) {
        SIFNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES));
        attr_done = true;
    }
    dim3 grid((a.W / 128) * ((a.H + R - 1) / R), a.B);
    kern<<<grid, TC_THREADS, SM::BYTES, st>>>(a);
    return sifnn::check_launch("conv3x3_tc_kernel");
}

template <int PAD, bool AFFINE>
int dispatch_tc(const TcArgs& a, cudaStream_t st) {
    switch (a.O) {
        case 16: return launch_tc<16, 2, PAD, AFFINE>(a, st);
        case 32: return launch_tc<32, 2, PAD, AFFINE>(a, st);
        case 64: return launch_tc<64, 2, PAD, AFFINE>(a, st);
        default: sifnn::set_error("conv3x3_tc: unsupported channel count %d", a.O); return SIFNN_EINVAL;
    }
}

}  // namespace

static int g_tc_dbg = 0;
extern "C" void sifnn_tc_debug(int mode) { g_tc_dbg = mode; }

extern "C" int sifnn_conv3x3_tc_supported(int Cin, int Cout, int H, int W) {
    return (W % 128 == 0) && (Cin % 8 == 0) && Cin <= 128 && (Cout == 16 || Cout == 32 || Cout == 64) && H >= 1;
}

extern "C" size_t sifnn_conv3x3_tc_wprep_bytes(int Cin, int Cout) { return (size_t)(Cin / 8) * 9 * 2 * 2 * Cout * 4 * sizeof(float); }

extern "C" int sifnn_conv3x3_fwd_tc(const float* in, const float* in_scale, const float* in_shift, const float* w, const float* bias,
                                    float* out, double* stats, void* wprep, int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(in && w && out && wprep, "conv3x3_fwd_tc: null pointer");
    SIFNN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_fwd_tc: in_scale/in_shift must both be set or both NULL");
    SIFNN_REQUIRE(sifnn_conv3x3_tc_supported(Cin, Cout, H, W) && B > 0 && B <= 65535, "conv3x3_fwd_tc: unsupported shape Cin=%d Cout=%d H=%d W=%d", Cin, Cout, H, W);
    cudaStream_t st = sifnn::as_stream(stream);
    const int total = (Cin / 8) * 9 * 2 * 2 * Cout * 4;
    tc_prep_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, static_cast<float*>(wprep), Cin, Cout, Cin * 9, 9, 0);
    SIFNN_TRY(sifnn::check_launch("tc_prep_weights_kernel"));
    TcArgs a{};
    a.in = in; a.in_scale = in_scale; a.in_shift = in_shift; a.wprep = static_cast<const float*>(wprep); a.bias = bias; a.out = out; a.stats = stats;
    a.B = B; a.K = Cin; a.O = Cout; a.H = H; a.W = W; a.accumulate = 0; a.dbg = g_tc_dbg;
    return in_scale ? dispatch_tc<0, true>(a, st) : dispatch_tc<0, false>(a, st);
}

// Main (zero-padded, transposed) part of the data gradient; the caller adds the replicate-padding border terms.
extern "C" int sifnn_conv3x3_dgrad_tc_main(const float* dy, const float* w, float* dx, int accumulate, void* wprep, int B, int Cin, int Cout,
                                           int H, int W, sifnn_stream_t stream) {
    SIFNN_REQUIRE(dy && w && dx && wprep, "conv3x3_dgrad_tc: null pointer");
    SIFNN_REQUIRE(sifnn_conv3x3_tc_supported(Cout, Cin, H, W) && B > 0 && B <= 65535, "conv3x3_dgrad_tc: unsupported shape");
    cudaStream_t st = sifnn::as_stream(stream);
    const int total = (Cout / 8) * 9 * 2 * 2 * Cin * 4;
    tc_prep_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, static_cast<float*>(wprep), Cout, Cin, 9, Cin * 9, 1);
    SIFNN_TRY(sifnn::check_launch("tc_prep_weights_kernel"));
    TcArgs a{};
    a.in = dy; a.wprep = static_cast<const float*>(wprep); a.out = dx;
    a.B = B; a.K = Cout; a.O = Cin; a.H = H; a.W = W; a.accumulate = accumulate ? 1 : 0; a.dbg = g_tc_dbg;
    return dispatch_tc<1, false>(a, st);
}
