// Throughput of the epilogue's building blocks with W warps resident on one SM (round 2): warp shuffles, shared-memory exchange, named barriers.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters, unsigned long long* clk) {
    __shared__ float xch[32 * 24 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = threadIdx.x * 0.001f + j; acc[j] = 0.f; }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {          // 16 shuffles + 8 FFMA-pairs (the kx reduce of 8 values)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float l = __shfl_up_sync(0xffffffffu, v[j], 1), r = __shfl_down_sync(0xffffffffu, v[j] + 1.f, 1);
                acc[j] = fmaf(l, 0.5f, fmaf(r, 0.25f, acc[j]));
            }
        } else if (MODE == 1) {   // same data movement through shared memory: 2 x STS.128 per 4 values, sync, 2 x LDS.128 of the neighbours
            float* mine = xch + (warp * 32 + lane) * 8;
            *reinterpret_cast<float4*>(mine) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(mine + 4) = make_float4(v[4], v[5], v[6], v[7]);
            __syncwarp();
            const float4 a = *reinterpret_cast<const float4*>(xch + (warp * 32 + ((lane + 31) & 31)) * 8);
            const float4 b = *reinterpret_cast<const float4*>(xch + (warp * 32 + ((lane + 1) & 31)) * 8 + 4);
            acc[0] += a.x + b.x; acc[1] += a.y + b.y; acc[2] += a.z + b.z; acc[3] += a.w + b.w;
            __syncwarp();
        } else if (MODE == 2) {   // only FFMAs (16 per iteration)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], 0.5f, fmaf(v[(j + 1) & 7], 0.25f, acc[j]));
        } else if (MODE == 3) {   // named barrier of 128 threads (4 warps)
            asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp >> 2)) : "memory");
            acc[0] += 1.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = acc[j] * 0.999f;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    unsigned long long* clk; cudaMalloc(&clk, 148 * 8);
    const int iters = 2000;
    const char* names[4] = {"16 SHFL + 16 FFMA", "smem exchange (2 STS.128 + 2 LDS.128 + 2 syncwarp)", "16 FFMA only", "bar.sync 128"};
    for (int warps : {4, 8, 16, 24}) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<1, warps * 32>>>(out, iters, clk);
                if (mode == 1) probe<1><<<1, warps * 32>>>(out, iters, clk);
                if (mode == 2) probe<2><<<1, warps * 32>>>(out, iters, clk);
                if (mode == 3) probe<3><<<1, warps * 32>>>(out, iters, clk);
            }
            cudaDeviceSynchronize();
            unsigned long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
            printf("%2d warps, %-52s: %7.1f clk / iteration (all warps together)\n", warps, names[mode], (double)h / iters);
        }
    }
    return 0;
}
