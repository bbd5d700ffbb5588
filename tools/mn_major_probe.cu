// Probe of tcgen05.mma kind::tf32 with MN-major operands.  CUTLASS (sm100_common.inl) states that the only shared-memory
// layout available to MN-major 32-bit operands is SWIZZLE_128B_BASE32B (descriptor layout type 1):
//   in 16-byte units ((8, n), (4, k)) : ((1, LBO), (8, SBO)), byte-address bits [5,7) ^= bits [7,9)
// i.e. one k (pixel) is a 128-byte line of 32 consecutive M (N) indices, 4 k lines form a 512-byte atom whose 32-byte
// chunks are XOR-permuted by the line index, the next 4 k are SBO bytes away, the next 32 M indices LBO bytes away.
// This program fills A and B in that layout with small integers and checks D = A * B^T exactly.
#include <cstdio>
#include <vector>
#include "../land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200/csrc/tc_common.cuh"
namespace sifnn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } unsigned long long launches() { return 0; } }
using namespace sifnn_tc;

__host__ __device__ inline float aval(int m, int k) { return (float)((m * 7 + k * 3) % 23 - 11); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k * 11) % 19 - 9); }
__host__ __device__ inline int elem_off(int m, int k, int lbo, int sbo) {  // float index
    const int G = m / 32, mm = m % 32, kk = k / 4, kr = k % 4;
    return (G * lbo + kk * sbo + kr * 128 + (((mm / 8) ^ kr) * 32) + (mm % 8) * 4) / 4;
}

__global__ void probe(float* out, int M, int N, int lbo, int sbo, int layout_type) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    float* A = smem;             // 32 KB
    float* Bm = smem + 8192;     // 64 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 24576; i += blockDim.x) smem[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < M * 8; i += blockDim.x) A[elem_off(i / 8, i % 8, lbo, sbo)] = aval(i / 8, i % 8);
    for (int i = tid; i < N * 8; i += blockDim.x) Bm[elem_off(i / 8, i % 8, lbo, sbo)] = bval(i / 8, i % 8);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (tid == 0) {
        const uint64_t lt = (uint64_t)layout_type << 61;
        const uint64_t da = make_desc(smem_u32(A), lbo, sbo) | lt, db = make_desc(smem_u32(Bm), lbo, sbo) | lt;
        umma_tf32(tb, da, db, make_idesc_mn(M, N), 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < N; n0 += 8) {
        float v[8];
        tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + n0, v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * N + n0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

void run(int M, int N, int lbo, int sbo, int lt) {
    float* d; cudaMalloc(&d, 128 * N * 4); cudaMemset(d, 0xff, 128 * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024);
    probe<<<1, 128, 98304 + 1024>>>(d, M, N, lbo, sbo, lt);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * N); cudaMemcpy(h.data(), d, 128 * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, zeros = 0;
    for (int m = 0; m < M; ++m) {
        const int lane = (M == 128) ? m : (m % 16) + 32 * (m / 16);
        for (int n = 0; n < N; ++n) {
            float want = 0.f;
            for (int k = 0; k < 8; ++k) want += aval(m, k) * bval(n, k);
            if (h[lane * N + n] != want) ++bad;
            if (h[lane * N + n] == 0.f) ++zeros;
        }
    }
    printf("M=%3d N=%3d lbo=%5d sbo=%4d layout=%d -> %s : mismatches %d of %d (zeros %d)\n", M, N, lbo, sbo, lt, cudaGetErrorString(e), bad, M * N, zeros);
    if (bad && bad < M * N) {
        int shown = 0;
        for (int m = 0; m < M && shown < 12; ++m) {
            const int lane = (M == 128) ? m : (m % 16) + 32 * (m / 16);
            for (int n = 0; n < N && shown < 12; ++n) {
                float want = 0.f;
                for (int k = 0; k < 8; ++k) want += aval(m, k) * bval(n, k);
                if (h[lane * N + n] != want) { printf("   m=%d n=%d got %g want %g\n", m, n, h[lane * N + n], want); ++shown; }
            }
        }
    }
    cudaFree(d);
}

int main() {
    run(128, 96, 2048, 512, 1);
    run(128, 48, 2048, 512, 1);
    run(64, 96, 2048, 512, 1);
    run(64, 48, 2048, 512, 1);
    run(128, 192, 2048, 512, 1);
    run(128, 32, 4096, 512, 1);
    run(128, 96, 512, 2048, 1);   // LBO / SBO swapped (control: must fail if the reading above is right)
    run(128, 96, 2048, 512, 0);   // no-swizzle layout type (control: known to return zeros)
    return 0;
}
