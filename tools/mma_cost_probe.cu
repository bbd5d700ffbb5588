// Cost model of small tcgen05.mma kind::tf32 instructions (M = 128, K = 8): clocks per MMA as a function of N and of the accumulator
// access pattern (same D back to back, rotating over several D slots, the [N=96 @ d, N=48 @ d+48] pair of conv3x3_tcx_kernel).
// One CTA, one issuing thread, operands in shared memory (contents irrelevant), clock64 around issue .. commit .. wait.
#include <cstdio>
#include <vector>
#include "../land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200/csrc/tc_common.cuh"
namespace sifnn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } unsigned long long launches() { return 0; } }
using namespace sifnn_tc;

// pattern: 0 same D; 1 rotate over `slots` D slots of `N` columns; 2 pair (N @ d, N/2 @ d + N/2); 3 pair rotating over slots
__global__ void probe(unsigned long long* out, int N, int pattern, int slots, int count) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += blockDim.x) smem[i] = 1.0f;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(smem), 128 * 16, 128), db = make_desc(smem_u32(smem + 4096), 256 * 16, 128);
        const uint32_t id1 = make_idesc(128, N), id2 = make_idesc(128, N / 2);
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i) {
            const uint32_t d = tb + ((pattern == 1 || pattern == 3) ? (i % slots) * N : 0);
            umma_tf32(d, da, db, id1, 1u);
            if (pattern >= 2) umma_tf32(d + N / 2, da, db, id2, 1u);
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

int main() {
    unsigned long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    struct Cfg { int N, pattern, slots; } cfgs[] = {{16, 0, 1}, {32, 0, 1}, {48, 0, 1}, {96, 0, 1}, {192, 0, 1}, {256, 0, 1}, {96, 1, 2}, {96, 1, 4}, {32, 1, 4},
                                                     {96, 2, 1}, {96, 3, 2}, {96, 3, 4}, {32, 2, 1}, {32, 3, 4}, {192, 2, 1}, {192, 3, 2}};
    const int count = 512;
    for (auto c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 65536 + 1024>>>(d, c.N, c.pattern, c.slots, count);
        cudaDeviceSynchronize();
        unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const int per = (c.pattern >= 2) ? 2 : 1;
        const double macs = (c.pattern >= 2 ? 1.5 : 1.0) * 128.0 * c.N * 8;
        printf("N=%3d pattern=%d slots=%d: issue %6.1f clk / iteration, complete %6.1f clk / iteration (%d MMA) -> %6.0f MAC/clk\n", c.N, c.pattern, c.slots,
               (double)h[0] / count, (double)h[1] / count, per, macs / ((double)h[1] / count));
    }
    return 0;
}
