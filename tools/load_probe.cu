// How fast can one SM pull 8 KB pieces of an NCHW fp32 tensor into shared memory?  (round 2: the full-fold convolution's loader issued one
// TMA box {128 px, 1 row, 16 planes} per step and the trace showed ~1340 clocks per box even when the data sat in L2.)
// Methods: TMA tensor boxes of several shapes, 16 x cp.async.bulk of 512 B, LDGSTS (cp.async 16 B) by one or two warps.
// Each CTA runs a ring of 4 stages with an mbarrier per stage and no consumer work; reports clocks per 8 KB and aggregate GB/s.
#include <cstdio>
#include <vector>
#include "../land-surface-temperature-super-resolution-with-a-scale-invariance-free-neural-approach_b200/csrc/tc_common.cuh"
namespace sifnn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } unsigned long long launches() { return 0; } }
using namespace sifnn_tc;

constexpr int RSMAX = 16, STAGE = 8192;

// mode 0: TMA tensor box (bw, bh, bp as encoded in the map); 1: 16 bulk copies of 512 B; 2: LDGSTS by `lw` warps
__global__ void __launch_bounds__(128) probe(const float* in, const __grid_constant__ CUtensorMap tmap, int mode, int lw, int W, int H, int planes, int steps, int stream,
                                             int bw, int bh, int bp, int RS, unsigned long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[RSMAX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { for (int s = 0; s < RS; ++s) mbar_init(full + s, mode == 2 ? lw * 32 : 1); fence_mbar_init(); }
    __syncthreads();
    const long long t0 = clock64();
    // walk: step i -> (image b, row r) like the convolution: CTA c owns rows [c * rows_per, ...) of the B*H rows (stream) or always row 0 (L2)
    const int nx = W / bw, ny = H / bh, np = planes / bp, total = nx * ny * np;
    if (mode != 2 ? (warp == 0) : (warp < lw)) {
        for (int i = 0; i < steps; ++i) {
            const int s = i % RS;
            if (i >= RS) mbar_wait(full + s, ((i / RS) - 1) & 1);   // previous use of this stage has landed (no consumer: reuse immediately)
            const int u = stream ? (int)(((long long)blockIdx.x * steps + i) % total) : 0;   // distinct boxes, y fastest, then x, then planes
            const int yb = u % ny, xb = (u / ny) % nx, pb = u / (ny * nx);
            const int x0 = xb * bw, r = yb * bh, p0 = pb * bp;
            unsigned char* dst = smem + s * STAGE;
            if (mode == 0) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(full + s, STAGE);
                    tma_load_3d(dst, &tmap, x0, r, p0, full + s);
                }
            } else if (mode == 1) {
                if (lane == 0) mbar_arrive_expect_tx(full + s, STAGE);
                __syncwarp();
                if (lane < 16) bulk_g2s(dst + lane * 512, in + ((size_t)(p0 + lane) * H + r) * W + x0, 512, full + s);
            } else {
                // 8 KB = 512 x 16 B; lw * 32 lanes
                for (int k = warp * 32 + lane; k < 512; k += lw * 32) {
                    const int c = k >> 5, x16 = k & 31;
                    const float* src = in + ((size_t)(p0 + c) * H + r) * W + x0 + x16 * 4;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + k * 16)), "l"(src) : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full + s)) : "memory");
            }
            __syncwarp();
        }
        for (int i = steps > RS ? steps - RS : 0; i < steps; ++i) mbar_wait(full + (i % RS), (i / RS) & 1);
    }
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    const int W = 256, H = 256, B = 32, planes = B * 16;
    float* in; cudaMalloc(&in, (size_t)planes * H * W * 4); cudaMemset(in, 0, (size_t)planes * H * W * 4);
    unsigned long long* out; cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMAX * STAGE + 1024);
    struct Box { const char* name; int bw, bh, bp; int swz; } boxes[] = {
        {"TMA {128 px, 1 row, 16 planes}", 128, 1, 16, 0}, {"TMA {128 px, 16 rows, 1 plane}", 128, 16, 1, 0}, {"TMA {128 px, 4 rows, 4 planes}", 128, 4, 4, 0},
        {"TMA {256 px, 1 row, 8 planes}", 256, 1, 8, 0}, {"TMA {256 px, 8 rows, 1 plane}", 256, 8, 1, 0}, {"TMA {32 px, 4 rows, 16 planes} sw128", 32, 4, 16, 1},
        {"TMA {32 px, 1 row, 64 planes}... (n/a)", 0, 0, 0, 0}};
    EncodeTiledFn enc = get_encode_fn();
    for (int RS : {4, 12})
    for (int stream = 0; stream < 2; ++stream) {
        const int grid = stream ? 148 : 1, steps = stream ? 200 : 400;
        printf("---- ring of %d stages, %s ----\n", RS, stream ? "148 CTAs streaming distinct rows from HBM" : "1 CTA, same 8 KB from L2");
        for (auto& bx : boxes) {
            if (!bx.bw) continue;
            CUtensorMap tm;
            const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
            const cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
            const cuuint32_t box[3] = {(cuuint32_t)bx.bw, (cuuint32_t)bx.bh, (cuuint32_t)bx.bp};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, in, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             bx.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("%-44s: encode failed %d\n", bx.name, (int)r); continue; }
            for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, RSMAX * STAGE + 1024>>>(in, tm, 0, 1, W, H, planes, steps, stream, bx.bw, bx.bh, bx.bp, RS, out);
            cudaDeviceSynchronize();
            std::vector<unsigned long long> h(grid); cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
            printf("%-44s: %7.1f clk / 8 KB per SM  (%6.0f GB/s aggregate at 1.965 GHz)\n", bx.name, (double)mx / steps, grid * 8192.0 * steps / (mx / 1.965e9) / 1e9);
        }
        CUtensorMap tm0{};
        {
            const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
            const cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
            const cuuint32_t box[3] = {128, 1, 16};
            const cuuint32_t estr[3] = {1, 1, 1};
            enc(&tm0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, in, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        struct M { const char* name; int mode, lw; } ms[] = {{"16 x cp.async.bulk 512 B", 1, 1}, {"LDGSTS 16 B, 1 warp", 2, 1}, {"LDGSTS 16 B, 2 warps", 2, 2}, {"LDGSTS 16 B, 4 warps", 2, 4}};
        for (auto& m : ms) {
            for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, RSMAX * STAGE + 1024>>>(in, tm0, m.mode, m.lw, W, H, planes, steps, stream, 128, 1, 16, RS, out);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<unsigned long long> h(grid); cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
            printf("%-44s: %7.1f clk / 8 KB per SM  (%6.0f GB/s aggregate at 1.965 GHz) %s\n", m.name, (double)mx / steps, grid * 8192.0 * steps / (mx / 1.965e9) / 1e9,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
