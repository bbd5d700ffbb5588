// TMA latency / issue-rate / pipelining probe: one CTA, lane 0 issues N boxes {128 px, 1 row, 16 planes} (8 KB) back to back into N stages, each with
// its own mbarrier, then waits for them in order; clock64 after every issue and after every completion.
#include <cstdio>
#include <vector>
#include "../land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200/csrc/tc_common.cuh"
namespace sifnn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } unsigned long long launches() { return 0; } }
using namespace sifnn_tc;
constexpr int N = 12, STAGE = 8192;

__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tmap, int rowbase, int onebar, int poll, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[N];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < N; ++s) mbar_init(full + s, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        long long t[2 * N + 1];
        t[0] = clock64();
        if (onebar) mbar_arrive_expect_tx(full, N * STAGE);
        for (int i = 0; i < N; ++i) {
            if (!onebar) mbar_arrive_expect_tx(full + i, STAGE);
            tma_load_3d(smem + i * STAGE, &tmap, 0, rowbase + i, blockIdx.x * 16, onebar ? full : full + i);
            t[1 + i] = clock64();
        }
        for (int i = 0; i < N; ++i) {
            if (poll) {   // pure test_wait spin
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(onebar ? full : full + i)), "r"(0u) : "memory");
            } else {
                mbar_wait(onebar ? full : full + i, 0);
            }
            t[1 + N + i] = clock64();
        }
        if (blockIdx.x == 0) for (int i = 0; i < 2 * N + 1; ++i) out[i] = t[i] - t[0];
    }
}

int main() {
    const int W = 256, H = 256, planes = 148 * 16;
    float* in; cudaMalloc(&in, (size_t)planes * H * W * 4); cudaMemset(in, 0, (size_t)planes * H * W * 4);
    long long* out; cudaMalloc(&out, (2 * N + 1) * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, N * STAGE + 1024);
    CUtensorMap tm;
    if (!encode_planes_map(&tm, in, W, H, planes, 128, 1, 16)) { printf("encode failed\n"); return 1; }
    for (int grid : {1, 148})
        for (int onebar = 0; onebar < 2; ++onebar)
            for (int poll = 0; poll < 2; ++poll)
                for (int rep = 0; rep < 3; ++rep) {
                    probe<<<grid, 64, N * STAGE + 1024>>>(tm, rep * 16, onebar, poll, out);
                    cudaDeviceSynchronize();
                    long long h[2 * N + 1]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("grid %3d %s %s rep %d (rep 0 = cold rows): issue", grid, onebar ? "one barrier " : "barrier/stage", poll ? "test_wait spin" : "try_wait      ", rep);
                    for (int i = 0; i < N; ++i) printf(" %lld", h[1 + i]);
                    printf(" | done");
                    for (int i = 0; i < N; ++i) printf(" %lld", h[1 + N + i]);
                    printf("\n");
                }
    return 0;
}
