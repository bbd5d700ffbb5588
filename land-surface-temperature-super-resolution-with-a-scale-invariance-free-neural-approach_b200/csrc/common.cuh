// Shared helpers for libsifnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/sifnn.h"

namespace sifnn {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int num_sms();
unsigned long long launches();
int tc_split_kind(int dgrad);               // operand format of the round-2 tensor-core convolutions: 0 BF16, 1 TF32, 2 FP16 (core.cu)
void tc_split_set(int fwd_kind, int dgrad_kind);

// Tensor-core convolution with the weight split done beforehand (csrc/conv3x3_tc.cu).  The split weights depend only on the layer, so the
// network plan prepares ALL layers of a pass in one launch (tc_prep_many) instead of one tiny launch in front of every convolution.
struct TcPrepJob { const float* w; void* wprep; int K, N, w_so, w_sk, flip, layout, total; };   // layout: 0 nine taps, 1 kx-folded, 2 kx-folded BF16
TcPrepJob tc_prep_job_fwd(const float* w, void* wprep, int Cin, int Cout, int W);
TcPrepJob tc_prep_job_dgrad(const float* w, void* wprep, int Cin, int Cout, int W);
int tc_prep_many(const TcPrepJob* jobs, int n, cudaStream_t st);
int conv3x3_fwd_tc_prepped(const float* in, const float* in_scale, const float* in_shift, const void* wprep, const float* bias, float* out, double* stats,
                           int B, int Cin, int Cout, int H, int W, cudaStream_t st);
int conv3x3_dgrad_tc_main_prepped(const float* dy, const void* wprep, float* dx, int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st);
int conv3x3_dgrad_border_cols(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H, int W, cudaStream_t st);

// BatchNorm finalize (bn_train_finalize_kernel, model.py:136 nn.BatchNorm2d in training mode) folded into the producing convolution: every CTA takes a
// ticket after its statistics atomics; the CTA that draws the last one turns the sums into scale / shift / mean / invstd and updates the running
// buffers -- one launch fewer per BatchNorm layer on the critical path of the step.  `counter` must be zero at launch; the last CTA re-arms it.
struct BnTail {
    unsigned int* counter;   // nullptr = no tail (the caller finalizes with sifnn_bn_train_finalize)
    const float* gamma;
    const float* beta;
    float* running_mean;     // may be nullptr (with running_var)
    float* running_var;
    float* scale;
    float* shift;
    float* save_mean;
    float* save_invstd;
    double n;                // elements per channel: B * H * W
};
#ifdef __CUDACC__
// Call from ALL threads of the CTA after a __syncthreads() that follows the CTA's statistics atomics, each issuing thread having executed
// __threadfence() after its atomics.  `flag` is a shared-memory int.
__device__ __forceinline__ void bn_tail_finalize(const BnTail& t, const double* stats, int C, unsigned int total_ctas, int* flag) {
    if (threadIdx.x == 0) *flag = (atomicAdd(t.counter, 1u) == total_ctas - 1u) ? 1 : 0;
    __syncthreads();
    if (*flag) {
        __threadfence();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double mean = __ldcg(stats + c) / t.n;
            double var = __ldcg(stats + C + c) / t.n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float invstd = (float)(1.0 / sqrt(var + 1e-5));
            const float sc = t.gamma[c] * invstd;
            t.scale[c] = sc;
            t.shift[c] = fmaf(-(float)mean, sc, t.beta[c]);
            t.save_mean[c] = (float)mean;
            t.save_invstd[c] = invstd;
            if (t.running_mean) {
                const double unbiased = t.n > 1.0 ? var * t.n / (t.n - 1.0) : var;
                t.running_mean[c] = (float)(0.9 * (double)t.running_mean[c] + 0.1 * mean);
                t.running_var[c] = (float)(0.9 * (double)t.running_var[c] + 0.1 * unbiased);
            }
        }
        if (threadIdx.x == 0) *t.counter = 0u;
    }
}
#endif

// Full-fold tensor-core convolution (csrc/conv3x3_ff.cu): all nine taps in the MMA's N dimension, shift-and-add epilogue, padding adjoint included.
bool conv3x3_ff_supported(int K, int O, int H, int W);
int ff_prep(const float* const* w, void* const* wprep, const int* K, const int* O, const int* w_so, const int* w_sk, const int* flip, int n, cudaStream_t st);
int conv3x3_fwd_ff_prepped(const float* in, const float* in2, int K1, const float* in_scale, const float* in_shift, const void* wprep, float* out, double* stats,
                           int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st, const BnTail* tail = nullptr);
int conv3x3_dgrad_ff_prepped(const float* dy, const void* wprep, float* dx, int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st);

// Fold + shift tensor-core convolution for widths that are multiples of 128 (csrc/conv3x3_fs.cu): ky in the MMA's N dimension, kx through shifted
// operand views, rolling row sums in the epilogue; complete data gradient (padding adjoint included).
bool conv3x3_fs_supported(int K, int O, int H, int W);
bool conv3x3_fs_fwd_preferred(int K, int O, int H, int W);   // forward only: also the 64-pixel-level shapes where the M = 64 form beats the full-fold kernel
size_t conv3x3_fs_wedge_bytes(int K, int O);
int fs_prep(const float* const* w, void* const* wprep, float* const* wedge, const int* K, const int* O, const int* w_so, const int* w_sk, const int* flip, int n,
            cudaStream_t st);
int conv3x3_fwd_fs_prepped(const float* in, const float* in2, int K1, const float* in_scale, const float* in_shift, const void* wprep, float* out, double* stats,
                           int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st, const BnTail* tail = nullptr);
int conv3x3_dgrad_fs_prepped(const float* dy, const void* wprep, const float* wedge, float* dx, int accumulate, int B, int Cin, int Cout, int H, int W, cudaStream_t st);

// Weight gradient in two halves for the network plan: `*_partials` launches only the main kernel, which leaves per-CTA partial sums
// [slots][Cout * Cin * 9] in `partial` (*slots_out slots); wgrad_reduce_many then sums the partials of MANY layers in one launch (fixed order,
// deterministic).  The public sifnn_conv3x3_wgrad_km / _tc do both for one layer.
int wgrad_km_partials(const float* in, const float* in_scale, const float* in_shift, const float* dy, void* partial, int B, int Cin, int Cout, int H, int W,
                      cudaStream_t st, int* slots_out);
int wgrad_tc_partials(const float* in, const float* in_scale, const float* in_shift, const float* dy, void* partial, int B, int Cin, int Cout, int H, int W,
                      cudaStream_t st, int* slots_out);
struct ReduceJob { const float* partial; float* out; int n; int slots; };
constexpr int REDUCE_MAX_JOBS = 20;
int wgrad_reduce_many(const ReduceJob* jobs, int njobs, cudaStream_t st);

// Programmatic dependent launch.  A kernel that starts with pdl_wait_and_trigger() may be launched with launch_pdl(): its CTAs are scheduled while
// the previous kernel of the stream is still draining (launch latency and block scheduling overlap that kernel's tail) and block in
// griddepcontrol.wait until the previous grid has completed and its memory is visible -- every global access comes after the wait, so the
// semantics are those of ordinary stream order.  griddepcontrol.launch_dependents right after it lets the NEXT kernel do the same.
// MEASURED (round 2, same-box A/B of the training step): attribute on the fs / ff / km / BatchNorm-backward kernels (~80 of 115 launches, SIFNN_PDL=1)
// 4.274 ms against 4.136 ms without -- a persistent 226 KB CTA scheduled early next to the CTAs of an HBM-bound elementwise kernel takes their
// occupancy.  Attribute on the elementwise kernels that FOLLOW a persistent kernel only (SIFNN_PDL=2, the default: BatchNorm backward,
// pooling, residual, up-sample + concat and their adjoints, the batched reduce): 4.095 against 4.121 ms -- their CTAs cannot co-reside with the
// persistent kernels they follow, so only the launch gap disappears.  SIFNN_PDL=0 turns it off.  (Kernels without the wait must keep the
// <<< >>> launch.)
bool pdl_enabled();
int pdl_mode();   // SIFNN_PDL: 0 off, 1 every converted kernel, 2 (default) only the elementwise kernels that follow a persistent kernel
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait_and_trigger() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool allow, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = allow ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    return launch_pdl_if(pdl_mode() == 1, kern, grid, block, smem, st, args...);
}
#endif

// "do once per device" (cudaFuncSetAttribute is per device; a process may drive several GPUs, possibly from several threads)
struct PerDeviceOnce {
    unsigned long long mask[2] = {0ull, 0ull};   // up to 128 devices
    bool first_time() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return true;   // unknown device: just do the work again
        const unsigned long long bit = 1ull << (dev & 63);
        const unsigned long long old = __atomic_fetch_or(&mask[dev >> 6], bit, __ATOMIC_ACQ_REL);
        return (old & bit) == 0;
    }
};

inline cudaStream_t as_stream(sifnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- cp.async (LDGSTS): global -> shared without staging registers; src_bytes < size zero-fills the rest ----
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- packed fp32 FMA (FFMA2, sm_100+): two IEEE fmas per instruction.  Same FMA-pipe throughput as two scalar
// FFMAs but half the issue slots, which is what bounds the direct convolutions (tools/ffma2_probe.cu).  ptxas folds
// a {x, x} pair into a scalar-broadcast operand, so broadcasting costs no instruction.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// same, but `volatile`: ptxas keeps the program order of these among themselves, which is how the kernels pin the
// operand-reuse-friendly (weight-stationary) issue order
__device__ __forceinline__ f32x2_t fma2_ordered(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ float act_affine_relu(float x, float sc, float sh) { return fmaxf(fmaf(x, sc, sh), 0.0f); }

}  // namespace sifnn

#define SIFNN_REQUIRE(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            sifnn::set_error(__VA_ARGS__);  \
            return SIFNN_EINVAL;            \
        }                                   \
    } while (0)

#define SIFNN_CUDA(expr)                                                         \
    do {                                                                         \
        cudaError_t e__ = (expr);                                                \
        if (e__ != cudaSuccess) {                                                \
            sifnn::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));   \
            return (int)e__;                                                     \
        }                                                                        \
    } while (0)

#define SIFNN_TRY(expr)          \
    do {                         \
        int r__ = (expr);        \
        if (r__ != 0) return r__; \
    } while (0)
