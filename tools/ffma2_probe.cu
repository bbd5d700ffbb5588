// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) throughput, and how much
// non-FMA work (LDS / integer) fits beside each.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>  // 0 scalar FFMA, 1 FFMA2, 2 FFMA + LDS mix, 3 FFMA2 + LDS mix
__global__ void __launch_bounds__(256) probe(float* sink, int iters) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 256] = 1.f; sm[threadIdx.x + 512] = 2.f; sm[threadIdx.x + 768] = 3.f;
    __syncthreads();
    float a[16]; u64 p[8];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const float m = 0.9999f, c = 1e-4f;
    const u64 m2 = pk(m, m), c2 = pk(c, c);
    float ld = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
            }
            if (MODE >= 2) {  // 4 LDS per 16 FMA (25% extra issue slots)
#pragma unroll
                for (int i = 0; i < 4; ++i) ld += sm[(threadIdx.x + 32 * (u * 4 + i) + it) & 1023];
            }
        }
    }
    float r = ld;
    for (int i = 0; i < 16; ++i) r += a[i];
    for (int i = 0; i < 8; ++i) { float2 f = *reinterpret_cast<float2*>(&p[i]); r += f.x + f.y; }
    if (r == 123.456f) sink[0] = r;
}

template <int MODE> void run(const char* name, float* sink) {
    const int blocks = 148 * 8, iters = 4000;
    probe<MODE><<<blocks, 256>>>(sink, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); probe<MODE><<<blocks, 256>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 16 * 8 * (double)iters * blocks * 256;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, best, flops / (best * 1e-3) / 1e12);
}
int main() {
    float* sink; cudaMalloc(&sink, 16);
    run<0>("FFMA", sink); run<1>("FFMA2", sink); run<2>("FFMA + 25% LDS", sink); run<3>("FFMA2 + 25% LDS (per FMA)", sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
