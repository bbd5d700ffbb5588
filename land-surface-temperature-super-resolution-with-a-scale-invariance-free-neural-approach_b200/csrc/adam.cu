// torch.optim.Adam(params, lr) step (train_model_B_gradFTM.py:453,121): betas (.9,.999),
// eps 1e-8, no weight decay, no amsgrad, bias-corrected -- fused over ONE flat buffer
// holding all 53 parameter tensors (282 705 floats), same operation order as
// torch's _single_tensor_adam.  The step counter lives on the device so that the whole
// training step can be replayed from a CUDA graph.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const int64_t* __restrict__ step_count, double lr,
                                                   double beta1d, double beta2d, double epsd, float grad_scale, int64_t n) {
    const double t = (double)(*step_count + 1);
    const double bc1 = 1.0 - pow(beta1d, t);
    const double bc2 = 1.0 - pow(beta2d, t);
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    // python-side scalars are doubles that torch rounds to fp32 when they meet the tensor
    const float w1 = (float)(1.0 - beta1d), beta2 = (float)beta2d, w2 = (float)(1.0 - beta2d), eps = (float)epsd;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = m[i] + w1 * (gi - m[i]);               // exp_avg.lerp_(grad, 1-beta1)
        const float vi = fmaf(w2, gi * gi, v[i] * beta2);           // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        m[i] = mi;
        v[i] = vi;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

__global__ void adam_bump_kernel(int64_t* step_count) { *step_count += 1; }

}  // namespace

extern "C" int sifnn_adam_step(float* p, const float* g, float* m, float* v, int64_t* step_count, double lr, double beta1,
                               double beta2, double eps, float grad_scale, int64_t n, sifnn_stream_t stream) {
    SIFNN_REQUIRE(p && g && m && v && step_count && n > 0, "adam_step: bad arguments");
    cudaStream_t st = sifnn::as_stream(stream);
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sifnn::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, step_count, lr, beta1, beta2, eps, grad_scale, n);
    SIFNN_TRY(sifnn::check_launch("adam_kernel"));
    adam_bump_kernel<<<1, 1, 0, st>>>(step_count);
    return sifnn::check_launch("adam_bump_kernel");
}
