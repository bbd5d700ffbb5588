// Second cost probe for small tcgen05.mma instructions (round 2): what makes an M = 128 MMA cost ~120-145 clocks in the round-1 kernels?
// Varies: operand layout of A / B in shared memory (no swizzle vs 128-byte swizzle), A from TMEM (TS form), kind (tf32 K = 8, bf16 K = 16),
// N, whether consecutive MMAs use the SAME or DIFFERENT A tiles, and the accumulator pattern.  One CTA, one issuing thread, clock64.
// Also times tcgen05.ld / tcgen05.st (drain and A-in-TMEM staging cost).
#include <cstdio>
#include <cstring>
#include <vector>
#include "../land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200/csrc/tc_common.cuh"
namespace sifnn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } unsigned long long launches() { return 0; } }
using namespace sifnn_tc;

struct Step { int dcol, a, b, n; };
struct Cfg {
    char name[96];
    int kind;   // 0 tf32 (K = 8), 1 bf16 (K = 16)
    int asrc;   // 0 smem no swizzle, 1 smem 128-byte swizzle, 2 TMEM
    int bsw;    // 0 no swizzle, 1 128-byte swizzle
    int M;
    int nsteps;
    Step s[8];
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_ts(int kind, uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc) {
    if (kind == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(1u) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(1u) : "memory");
}

constexpr int A_TILE_BYTES = 16384, B_TILE_BYTES = 32768, NA = 4, NB = 3;
constexpr int SMEM_BYTES = NA * A_TILE_BYTES + NB * B_TILE_BYTES + 1024;

template <int NS>
__global__ void probe(unsigned long long* out, const Cfg c, int count) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (NA * A_TILE_BYTES + NB * B_TILE_BYTES) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    // zero the TMEM A region (columns 384..511) so the TS form reads defined data
    {
        const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 384;
        for (int cb = 0; cb < 128; cb += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta + cb), "r"(0x3F800000u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        uint32_t d[NS], id[NS], at[NS];
        uint64_t da[NS], db[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const Step s = c.s[i];
            d[i] = tb + s.dcol;
            id[i] = c.kind ? make_idesc_bf16(c.M, s.n) : make_idesc(c.M, s.n);
            const uint32_t aaddr = smem_u32(smem + s.a * A_TILE_BYTES), baddr = smem_u32(smem + NA * A_TILE_BYTES + s.b * B_TILE_BYTES);
            da[i] = c.asrc == 1 ? desc_sw128(aaddr) : make_desc(aaddr, c.M * 16, 128);
            db[i] = c.bsw == 1 ? desc_sw128(baddr) : make_desc(baddr, 256 * 16, 128);
            at[i] = tb + 384 + s.a * 32;
        }
        const long long t0 = clock64();
        for (int it = 0; it < count; ++it) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                if (c.asrc == 2) umma_ts(c.kind, d[i], at[i], db[i], id[i]);
                else if (c.kind == 0) umma_tf32(d[i], da[i], db[i], id[i], 1u);
                else umma_bf16(d[i], da[i], db[i], id[i], 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// tcgen05.ld / st throughput: `warps` warps (4 or 8; warp w reads lane quadrant w % 4), each moves `cols` columns per repetition
template <int X, bool STORE>
__global__ void ldst_probe(unsigned long long* out, int cols, int reps) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int cb = 0; cb < cols; cb += X) {
            uint32_t v[X];
            if constexpr (STORE) {
                if constexpr (X == 8)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + cb), "r"(acc) : "memory");
                else
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + cb), "r"(acc) : "memory");
            } else {
                if constexpr (X == 8) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tb + cb));
                } else if constexpr (X == 16) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(tb + cb));
                } else {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                                   "=r"(v[30]), "=r"(v[31]) : "r"(tb + cb));
                }
            }
            if constexpr (!STORE) {
#pragma unroll
                for (int j = 0; j < X; ++j) acc ^= v[j];   // consumed after the wait below; keeps the loads alive
            }
        }
        if constexpr (STORE) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) out[2] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

static std::vector<Cfg> cfgs;
static void add(const char* name, int kind, int asrc, int bsw, std::vector<Step> st) {
    Cfg c{};
    snprintf(c.name, sizeof(c.name), "%s", name);
    c.kind = kind; c.asrc = asrc; c.bsw = bsw; c.M = 128; c.nsteps = (int)st.size();
    for (size_t i = 0; i < st.size(); ++i) c.s[i] = st[i];
    cfgs.push_back(c);
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 64);
    const char* lay[3] = {"A smem none", "A smem sw128", "A tmem"};
    for (int kind = 0; kind < 2; ++kind)
        for (int asrc = 0; asrc < 3; ++asrc)
            for (int bsw = 0; bsw < 2; ++bsw) {
                if (asrc == 0 && bsw == 1) continue;
                if (asrc == 1 && bsw == 0) continue;
                char nm[96];
                for (int N : {48, 96, 144, 192, 256}) {
                    snprintf(nm, 96, "%s %s Bsw%d N=%3d same A, same D", kind ? "bf16" : "tf32", lay[asrc], bsw, N);
                    add(nm, kind, asrc, bsw, {{0, 0, 0, N}});
                    snprintf(nm, 96, "%s %s Bsw%d N=%3d alt A0/A1, same D", kind ? "bf16" : "tf32", lay[asrc], bsw, N);
                    add(nm, kind, asrc, bsw, {{0, 0, 0, N}, {0, 1, 0, N}});
                }
                snprintf(nm, 96, "%s %s Bsw%d N=144 x3 (hi.hi, lo.hi, hi.lo) same D", kind ? "bf16" : "tf32", lay[asrc], bsw);
                add(nm, kind, asrc, bsw, {{0, 0, 0, 144}, {0, 1, 0, 144}, {0, 0, 1, 144}});
                snprintf(nm, 96, "%s %s Bsw%d N=144 alt D0/D1 per MMA, alt A", kind ? "bf16" : "tf32", lay[asrc], bsw);
                add(nm, kind, asrc, bsw, {{0, 0, 0, 144}, {144, 1, 0, 144}});
                snprintf(nm, 96, "%s %s Bsw%d N=144 3 on D0 then 3 on D1", kind ? "bf16" : "tf32", lay[asrc], bsw);
                add(nm, kind, asrc, bsw, {{0, 0, 0, 144}, {0, 1, 0, 144}, {0, 0, 1, 144}, {144, 2, 0, 144}, {144, 3, 0, 144}, {144, 2, 1, 144}});
                snprintf(nm, 96, "%s %s Bsw%d pair (192 @0 A0, 96 @96 A1)", kind ? "bf16" : "tf32", lay[asrc], bsw);
                add(nm, kind, asrc, bsw, {{0, 0, 0, 192}, {96, 1, 0, 96}});
                snprintf(nm, 96, "%s %s Bsw%d pair (96 @0 A0, 48 @48 A1)", kind ? "bf16" : "tf32", lay[asrc], bsw);
                add(nm, kind, asrc, bsw, {{0, 0, 0, 96}, {48, 1, 0, 48}});
                snprintf(nm, 96, "%s %s Bsw%d N=96 3 A tiles (3 ky) x hi/lo, same D (round-1 tcx row)", kind ? "bf16" : "tf32", lay[asrc], bsw);
                add(nm, kind, asrc, bsw, {{0, 0, 0, 96}, {48, 1, 0, 48}, {0, 2, 1, 96}, {48, 3, 1, 48}, {0, 0, 2, 96}, {48, 1, 2, 48}});
            }
    const int count = 400;
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(probe<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (c.nsteps) {
                case 1: probe<1><<<1, 128, SMEM_BYTES>>>(d, c, count); break;
                case 2: probe<2><<<1, 128, SMEM_BYTES>>>(d, c, count); break;
                case 3: probe<3><<<1, 128, SMEM_BYTES>>>(d, c, count); break;
                case 6: probe<6><<<1, 128, SMEM_BYTES>>>(d, c, count); break;
            }
        }
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        double macs = 0;
        const int K = 8 * (c.kind + 1);
        for (int i = 0; i < c.nsteps; ++i) macs += 128.0 * c.s[i].n * K;
        const double per_it = (double)h[1] / count;
        printf("%-72s: %7.1f clk / iteration (%d MMA, %6.1f each) -> %6.0f MAC/clk%s\n", c.name, per_it, c.nsteps, per_it / c.nsteps, macs / per_it,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    // TMEM load / store
    for (int warps : {4, 8}) {
        for (int cols : {96, 144, 256}) {
            const int reps = 200;
            unsigned long long h;
            ldst_probe<8, false><<<1, warps * 32>>>(d, cols, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("tcgen05.ld x8  %d warps, %3d columns each: %7.1f clk / repetition (%5.1f B/clk per CTA)\n", warps, cols, (double)h / reps, warps * 32.0 * cols * 4 * reps / h);
            ldst_probe<16, false><<<1, warps * 32>>>(d, cols, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("tcgen05.ld x16 %d warps, %3d columns each: %7.1f clk / repetition (%5.1f B/clk per CTA)\n", warps, cols, (double)h / reps, warps * 32.0 * cols * 4 * reps / h);
            if (cols % 32 == 0) {
                ldst_probe<32, false><<<1, warps * 32>>>(d, cols, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("tcgen05.ld x32 %d warps, %3d columns each: %7.1f clk / repetition (%5.1f B/clk per CTA)\n", warps, cols, (double)h / reps, warps * 32.0 * cols * 4 * reps / h);
            }
            ldst_probe<8, true><<<1, warps * 32>>>(d, cols, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("tcgen05.st x8  %d warps, %3d columns each: %7.1f clk / repetition (%5.1f B/clk per CTA)\n", warps, cols, (double)h / reps, warps * 32.0 * cols * 4 * reps / h);
            ldst_probe<16, true><<<1, warps * 32>>>(d, cols, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("tcgen05.st x16 %d warps, %3d columns each: %7.1f clk / repetition (%5.1f B/clk per CTA)\n", warps, cols, (double)h / reps, warps * 32.0 * cols * 4 * reps / h);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
