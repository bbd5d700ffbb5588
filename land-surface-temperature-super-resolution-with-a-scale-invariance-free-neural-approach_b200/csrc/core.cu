// Error reporting, device queries and the fp32 peak probe.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

namespace sifnn {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels launched by this library (bench.py reports it)

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED);
    return 0;
}

unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// Operand format of the 3-term split in the round-2 tensor-core convolutions, per direction (0 forward, 1 data gradient):
//   0 BF16 (K = 16 per MMA; 16 significant bits, fp32 range; per-layer error ~5e-6)
//   1 TF32 (K =  8 per MMA: twice the MMAs; 22 bits; ~4e-7)
//   2 FP16 (K = 16 per MMA; 22 bits like TF32 at the MMA count of BF16; the residual is scaled by 2^11 so it never goes subnormal; needs
//           |value| < 65504: fine for z-scored inputs / BatchNorm outputs / weights, NOT for gradients, whose magnitudes can be tiny)
// Defaults: forward FP16 -- the gradients of weights that feed a train-mode BatchNorm amplify forward rounding ~1000x, BF16 fails the per-tensor
// gradient check of tests/test_gpu_model.py there; data gradient BF16.  Environment: SIFNN_FWD_SPLIT / SIFNN_DGRAD_SPLIT = bf16 | tf32 | fp16.
static int g_split[2] = {-1, -1};
int tc_split_kind(int dgrad) {
    int& v = g_split[dgrad ? 1 : 0];
    if (v < 0) {
        v = dgrad ? 0 : 2;
        const char* e = getenv(dgrad ? "SIFNN_DGRAD_SPLIT" : "SIFNN_FWD_SPLIT");
        if (e) {
            if (e[0] == 'b' || e[0] == 'B') v = 0;
            else if (e[0] == 't' || e[0] == 'T') v = 1;
            else if (e[0] == 'f' || e[0] == 'F') v = 2;
        }
    }
    return v;
}
void tc_split_set(int fwd_kind, int dgrad_kind) { g_split[0] = fwd_kind; g_split[1] = dgrad_kind; }

int pdl_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIFNN_PDL"); v = e ? atoi(e) : 2; if (v < 0 || v > 2) v = 2; }   // default 2: see common.cuh
    return v;
}
bool pdl_enabled() { return pdl_mode() != 0; }

int num_sms() {   // of the CURRENT device (cached per device)
    static int cache[128] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) dev = 0;
    int n = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        __atomic_store_n(&cache[dev], n, __ATOMIC_RELAXED);
    }
    return n;
}

}  // namespace sifnn

extern "C" int sifnn_version(void) { return SIFNN_VERSION; }
extern "C" const char* sifnn_last_error(void) { return sifnn::g_err; }
extern "C" unsigned long long sifnn_launch_count(void) { return sifnn::launches(); }

namespace {
// 8 independent FFMA chains per thread, operands in registers: measures the fp32
// CUDA-core roof the SIMT convolutions are judged against (SURVEY section 8d asks the
// builder to measure it; it is not in MEASURED_PEAKS.json).
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* sink, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.9999f + blockIdx.x * 1e-9f, c = 1e-4f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456f) sink[0] = r;
}
}  // namespace

extern "C" int sifnn_fp32_peak_kernel(float* sink, int iters, double* flops_out, sifnn_stream_t stream) {
    SIFNN_REQUIRE(sink && iters > 0, "fp32_peak: bad arguments");
    const int blocks = sifnn::num_sms() * 8, threads = 256;
    fp32_peak_kernel<<<blocks, threads, 0, sifnn::as_stream(stream)>>>(sink, iters);
    if (flops_out) *flops_out = 2.0 * 128.0 * (double)iters * blocks * threads;
    return sifnn::check_launch("fp32_peak_kernel");
}
