/*
 * sifnn.h -- C-ABI of libsifnn_b200.so: the SIF-NN-SR ModelB hot path as hand-written
 * sm_100a CUDA kernels.
 *
 * The reference (cgranerob/Land-Surface-Temperature-Super-Resolution-...) has no FFI of
 * its own: its hot path is a stack of PyTorch library calls.  Each entry point below
 * names the reference call site (file:line, relative to the reference root) it
 * replaces.  All pointers are DEVICE pointers to fp32, NCHW-contiguous data unless the
 * parameter is documented as "host".  The caller (PyTorch) owns every buffer, including
 * workspaces; the library never allocates or frees device memory, never synchronises the
 * host, and launches everything on the given stream.
 *
 * Return value: 0 on success, otherwise a non-zero code (cudaError_t value for CUDA
 * failures, SIFNN_EINVAL for bad arguments); sifnn_last_error() returns a thread-local
 * message.  There is no CPU fallback.
 */
#ifndef SIFNN_H
#define SIFNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sifnn_stream_t; /* cudaStream_t */

#define SIFNN_EINVAL 10001
#define SIFNN_VERSION 1

int sifnn_version(void);
const char* sifnn_last_error(void);
/* number of CUDA kernels this library has launched in this process (monotonic) */
unsigned long long sifnn_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Per-op entry points
 * ---------------------------------------------------------------------------------- */

/* nn.Conv2d(k=3, padding=1, padding_mode='replicate') forward  (model.py:135,138,507,605).
 * in  (B,Cin,H,W); the consumer-side prologue optionally applies the previous layer's
 * BatchNorm+ReLU on load: v = max(in*in_scale[c] + in_shift[c], 0)  (model.py:136-137);
 * pass NULL/NULL for a plain input.   w (Cout,Cin,3,3); bias (Cout) or NULL.
 * out (B,Cout,H,W) receives the raw (pre-BatchNorm) convolution.
 * stats: NULL, or 2*Cout doubles that receive += (sum, sum of squares) of `out` per
 * channel -- the BatchNorm batch statistics (model.py:136) gathered in the epilogue. */
int sifnn_conv3x3_fwd(const float* in, const float* in_scale, const float* in_shift,
                      const float* w, const float* bias, float* out, double* stats,
                      int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Data gradient of the same convolution (autograd of model.py:135; loss.backward() at
 * train_model_B_gradFTM.py:119).  dy (B,Cout,H,W) -> dx (B,Cin,H,W), including the
 * adjoint of the replicate padding.  accumulate != 0 adds into dx. */
int sifnn_conv3x3_dgrad(const float* dy, const float* w, float* dx, int accumulate,
                        int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Tensor-core variants (tcgen05 / TMEM implicit GEMM, 3-term TF32 split: fp32-accurate).  Same semantics as
 * sifnn_conv3x3_fwd / the zero-padded main part of sifnn_conv3x3_dgrad, for shapes where
 * sifnn_conv3x3_tc_supported() != 0 (W % 128 == 0, Cin % 8 == 0, Cout in {16,32,64}).  wprep: caller-owned
 * scratch of sifnn_conv3x3_tc_wprep_bytes() bytes that receives the hi/lo-split weights of the call.
 * sifnn_conv3x3_dgrad_tc_main does NOT add the replicate-padding border terms; sifnn_conv3x3_dgrad_tc does
 * (same kernel + the border pass of the SIMT path). */
int sifnn_conv3x3_tc_supported(int Cin, int Cout, int H, int W);
size_t sifnn_conv3x3_tc_wprep_bytes(int Cin, int Cout);
int sifnn_conv3x3_fwd_tc(const float* in, const float* in_scale, const float* in_shift,
                         const float* w, const float* bias, float* out, double* stats, void* wprep,
                         int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);
int sifnn_conv3x3_dgrad_tc_main(const float* dy, const float* w, float* dx, int accumulate, void* wprep,
                                int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);
int sifnn_conv3x3_dgrad_tc(const float* dy, const float* w, float* dx, int accumulate, void* wprep,
                           int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);
/* The border pass alone: dx += adjoint-of-replicate-padding terms (used after a *_main call). */
int sifnn_conv3x3_dgrad_border(const float* dy, const float* w, float* dx,
                               int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Full-fold tensor-core convolution (csrc/conv3x3_ff.cu, round 2): the nine taps of 16 output channels are one N = 144 MMA per 16 (BF16 split)
 * or 8 (TF32 split, SIFNN_FF_TF32=1) input channels; the epilogue shifts and adds.  Same semantics as sifnn_conv3x3_fwd (no bias) and as the COMPLETE
 * sifnn_conv3x3_dgrad (the adjoint of the replicate padding is part of the epilogue: no border pass).  Shapes: W in {32,64,128,256},
 * Cin (forward) / Cout (data gradient) a multiple of 16 up to 64, the other channel count a multiple of 16 up to 128.
 * wprep: sifnn_conv3x3_tc_wprep_bytes() bytes of scratch. */
int sifnn_conv3x3_ff_supported(int Cin, int Cout, int H, int W);
/* kind: operand format of the 3-term split, 0 BF16, 1 TF32, 2 FP16 (forward only; the data gradient then uses BF16) -- sets BOTH round-2 kernels;
 * max_ctas > 0 caps gridDim.x (tests: long row strips on small inputs), 0 = one CTA per SM.  Defaults: forward FP16, data gradient BF16
 * (environment: SIFNN_FWD_SPLIT / SIFNN_DGRAD_SPLIT = bf16 | tf32 | fp16). */
void sifnn_conv3x3_ff_config(int kind, int max_ctas);
/* debug: ablation bits for pipeline studies (1 no MMAs, 2 no epilogue math, 4 no TMEM loads, 8 no transform, 16 no stores, 32 L2-resident loads); results are then wrong */
void sifnn_conv3x3_ff_debug(int ablate);
/* debug: device buffer of 11 * 256 uint64 that receives clock64 stamps of CTA (0,0) per pipeline step (tools/trace_ff.py), or NULL */
void sifnn_conv3x3_ff_trace(void* buf);
int sifnn_conv3x3_fwd_ff(const float* in, const float* in_scale, const float* in_shift, const float* w,
                         float* out, double* stats, void* wprep,
                         int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);
int sifnn_conv3x3_dgrad_ff(const float* dy, const float* w, float* dx, int accumulate, void* wprep,
                           int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Fold + shift tensor-core convolution (csrc/conv3x3_fs.cu, round 2) for widths that are multiples of 128: the three filter rows are stacked in the MMA's
 * N dimension, the three filter columns are three start addresses into one staged input row, the epilogue keeps two partial output rows in registers.
 * Same semantics as sifnn_conv3x3_fwd (no bias) and the COMPLETE sifnn_conv3x3_dgrad.  Channel limits as for the full-fold kernel.
 * wprep: sifnn_conv3x3_tc_wprep_bytes() bytes; the data gradient needs 2 * Cout * 3 * Cin * 4 bytes more (fp32 edge taps). */
int sifnn_conv3x3_fs_supported(int Cin, int Cout, int H, int W);
void sifnn_conv3x3_fs_config(int kind, int max_ctas);
/* also accept 64- and 32-pixel-wide images (M = 64 MMAs, one image row each); off by default: the full-fold kernel is faster there */
void sifnn_conv3x3_fs_narrow(int on);
/* debug: device buffer of 16 * 256 uint64 for clock64 stamps of CTA (0,0) per pipeline step (tools/trace_fs.py), or NULL */
void sifnn_conv3x3_fs_trace(void* buf);
int sifnn_conv3x3_fwd_fs(const float* in, const float* in_scale, const float* in_shift, const float* w,
                         float* out, double* stats, void* wprep,
                         int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);
int sifnn_conv3x3_dgrad_fs(const float* dy, const float* w, float* dx, int accumulate, void* wprep,
                           int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Weight gradient.  `in`/in_scale/in_shift as in sifnn_conv3x3_fwd.  dw (Cout,Cin,3,3)
 * is overwritten; dbias (Cout) or NULL.  workspace: sifnn_conv3x3_wgrad_workspace()
 * bytes of scratch (per-CTA partial sums, reduced in a fixed order -> deterministic). */
size_t sifnn_conv3x3_wgrad_workspace(int B, int Cin, int Cout, int H, int W);
int sifnn_conv3x3_wgrad(const float* in, const float* in_scale, const float* in_shift,
                        const float* dy, float* dw, float* dbias, void* workspace,
                        int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Same weight gradient on the tcgen05 tensor cores (pixels = MMA K dimension, MN-major operands, 3-term TF32 split;
 * csrc/wgrad_tc.cu).  Shapes: Cin == 16 or a multiple of 32, Cout in {16,32,64}, W a multiple of 16
 * (sifnn_conv3x3_wgrad_tc_supported); no dbias.  workspace: sifnn_conv3x3_wgrad_tc_workspace() bytes. */
int sifnn_conv3x3_wgrad_tc_supported(int Cin, int Cout, int H, int W);
size_t sifnn_conv3x3_wgrad_tc_workspace(int B, int Cin, int Cout, int H, int W);
int sifnn_conv3x3_wgrad_tc(const float* in, const float* in_scale, const float* in_shift,
                           const float* dy, float* dw, void* workspace,
                           int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Weight gradient with 16-bit K-major operands (csrc/wgrad_km.cu, round 2; autograd of nn.Conv2d at model.py:135 w.r.t. the weight): 16 pixels per
 * MMA, all nine taps and all hi/lo products of two dy rows in ONE MMA per K step (M = 4 input rows x hi/lo x channels, N = 2 rows x 3 column-shifted
 * dy copies x hi/lo x channels).  Both operands are split hi + lo in BF16 (16 significant bits, fp32 exponent range); kind::f16 takes one format for
 * both operands of an MMA (FP16 x BF16 faults), sifnn_conv3x3_wgrad_km_config(fmt_x, fmt_dy) with 0 = FP16, 1 = BF16 exists for experiments
 * (defaults 1, 1; 0, 0 is 10x more accurate but only safe for gradients of O(1) magnitude).
 * Same result contract as sifnn_conv3x3_wgrad_tc.  Shapes: Cin, Cout = 16 or a multiple of 32 (up to 256), W a multiple of 32, H a multiple of
 * 8 / 4 / 2 (rows per tile).  workspace: sifnn_conv3x3_wgrad_km_workspace() bytes.
 * Two-step form (what the network plan uses): sifnn_conv3x3_wgrad_km_partials leaves per-CTA partial sums [*slots][Cout*Cin*9] in `workspace`;
 * sifnn_wgrad_reduce sums them in a fixed order (deterministic) into dw. */
int sifnn_conv3x3_wgrad_km_supported(int Cin, int Cout, int H, int W);
size_t sifnn_conv3x3_wgrad_km_workspace(int B, int Cin, int Cout, int H, int W);
void sifnn_conv3x3_wgrad_km_config(int fmt_x, int fmt_dy);
int sifnn_conv3x3_wgrad_km_partials(const float* in, const float* in_scale, const float* in_shift, const float* dy, void* workspace,
                                    int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream, int* slots);
int sifnn_wgrad_reduce(const float* partial, float* dw, int n, int slots, sifnn_stream_t stream);
int sifnn_conv3x3_wgrad_km(const float* in, const float* in_scale, const float* in_shift,
                           const float* dy, float* dw, void* workspace,
                           int B, int Cin, int Cout, int H, int W, sifnn_stream_t stream);

/* Train-time quality metrics on the device (utils.py:548-578 psnr_skimage / ssim_skimage, called every step at
 * train_model_B_gradFTM.py:126-127 after two device->host copies): out2[0] = mean PSNR, out2[1] = mean SSIM
 * (skimage defaults: 7x7 uniform window, sample covariance, K1 .01, K2 .03) of pred against target, both (B,1,H,W),
 * data_range = max - min of the whole target batch.  workspace: sifnn_quality_workspace_bytes(B). */
size_t sifnn_quality_workspace_bytes(int B);
int sifnn_quality_psnr_ssim(const float* pred, const float* target, float* out2, void* workspace,
                            int B, int H, int W, sifnn_stream_t stream);

/* nn.BatchNorm2d training statistics -> affine (model.py:136,139,508).
 * stats (2*C doubles: sum, sumsq over n = B*H*W).  Writes scale = gamma*invstd,
 * shift = beta - mean*scale, save_mean, save_invstd (each C floats) and, if
 * running_mean != NULL, updates the running buffers with momentum 0.1 and the unbiased
 * variance.  eps = 1e-5. */
int sifnn_bn_train_finalize(const double* stats, const float* gamma, const float* beta,
                            float* running_mean, float* running_var,
                            float* scale, float* shift, float* save_mean, float* save_invstd,
                            int C, double n, sifnn_stream_t stream);

/* nn.BatchNorm2d eval: scale/shift from the running buffers. */
int sifnn_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean,
                         const float* running_var, float* scale, float* shift, int C,
                         sifnn_stream_t stream);

/* BatchNorm+ReLU backward (autograd of model.py:136-137), two passes.
 * reduce: sums (2*C doubles, zeroed by the caller) += (sum dy, sum dy*xhat), with
 *         dy = dY * (raw*scale+shift > 0), xhat = (raw-mean)*invstd.
 * apply : dx = gamma*invstd*(dy - sum_dy/n - xhat*sum_dy_xhat/n); dgamma = sum dy*xhat,
 *         dbeta = sum dy.  dx may alias dY. */
int sifnn_bn_relu_bwd_reduce(const float* dY, const float* raw, const float* scale, const float* shift,
                             const float* save_mean, const float* save_invstd, double* sums,
                             int B, int C, int HW, sifnn_stream_t stream);
int sifnn_bn_relu_bwd_apply(const float* dY, const float* raw, const float* scale, const float* shift,
                            const float* save_mean, const float* save_invstd, const float* gamma,
                            const double* sums, float* dx, float* dgamma, float* dbeta,
                            int B, int C, int HW, sifnn_stream_t stream);

/* AvgPool2d(2,2) of relu(raw*scale+shift)  (model.py:504,529) and its adjoint. */
int sifnn_act_avgpool2_fwd(const float* raw, const float* scale, const float* shift, float* out,
                           int B, int C, int H, int W, sifnn_stream_t stream);
int sifnn_avgpool2_bwd(const float* dout, float* din, int accumulate,
                       int B, int C, int H, int W, sifnn_stream_t stream);

/* ResidualConnection: out = x + relu(raw*scale+shift)  (model.py:311-312). */
int sifnn_act_residual_fwd(const float* x, const float* raw, const float* scale, const float* shift,
                           float* out, int B, int C, int HW, sifnn_stream_t stream);

/* UpBlock head: out = cat([bilinear_x2(relu(low)), relu(skip)], 1), align_corners=True
 * (model.py:207,236,247).  low (B,C1,H,W) raw + its affine; skip (B,C2,2H,2W) raw + affine. */
int sifnn_act_upcat_fwd(const float* low, const float* low_scale, const float* low_shift,
                        const float* skip, const float* skip_scale, const float* skip_shift,
                        float* out, int B, int C1, int C2, int H, int W, sifnn_stream_t stream);
/* Adjoint: dlow (B,C1,H,W) = up2^T(dout[:, :C1]); dskip (B,C2,2H,2W) = dout[:, C1:]. */
int sifnn_upcat_bwd(const float* dout, float* dlow, float* dskip,
                    int B, int C1, int C2, int H, int W, sifnn_stream_t stream);

/* Input preparation: x[:,0] = bicubic_x4(lst) (cv2.INTER_CUBIC, utils.py:163-180),
 * x[:,1] = ndvi  (torch.cat at train_model_B_gradFTM.py:94 / predict.py:101).
 * lst (B,1,h,w), ndvi (B,1,4h,4w), x (B,2,4h,4w). */
int sifnn_bicubic4_cat(const float* lst, const float* ndvi, float* x, int B, int h, int w,
                       sifnn_stream_t stream);

/* Whole-tile driver (reference predict.py:84-103).  lst_tile (Ht,Wt) Kelvin, ndvi_tile (4Ht,4Wt); patch p of the list
 * is the 64x64 LST window / 256x256 NDVI window at window coordinates (wy[p], wx[p]) (device int arrays).
 * gather : lst_out (P,1,64,64) = (lst - mean_lst)/std_lst; ndvi_out (P,1,256,256) = (clip(ndvi,-1,1) - mean_ndvi)/std_ndvi
 * scatter: out_tile[4*64*wy .. , 4*64*wx ..] = sr * std_lst + mean_lst          (predict.py:88-89,96-103) */
int sifnn_tile_gather(const float* lst_tile, const float* ndvi_tile, const int* wy, const int* wx,
                      float* lst_out, float* ndvi_out, int P, int Ht, int Wt,
                      float mean_lst, float std_lst, float mean_ndvi, float std_ndvi, sifnn_stream_t stream);
int sifnn_tile_scatter(const float* sr, const int* wy, const int* wx, float* out_tile, int P, int Ht, int Wt,
                       float mean_lst, float std_lst, sifnn_stream_t stream);

/* Fused loss forward + dLoss/dSR.
 * kind 1 = SR1 (train_model_B_predef_filters.py:111-133: Huber(downscale) + Huber(Sobel4)),
 * kind 2 = SR2 (train_model_B_gradFTM.py:99-117: Huber(downscale) + Huber(x - G_0.25(x))).
 * sr, ndvi (B,1,H,W); lst (B,1,H/4,W/4).
 * tab_ds : H x 3   per-axis ADJOINT weights of (blur mtf 0.1 + bicubic/4): tab_ds[r][i] is the weight of
 *                   SR row r in low-res row r/4-1+i (reflection folded in).  Built on the host, lives on device.
 * h12    : 12      forward taps of the same operator: D[I] = sum_t h12[t] * SR[reflect(4I-4+t)]
 * tab_lp : H x 9   per-axis adjoint weights of the reflect-padded G_0.25 blur (SR2 only, else NULL)
 * g9     : 9       taps of G_0.25 (SR2 only, else NULL)
 * losses : 3 doubles, zeroed by the caller: += (ds, percep, alpha*ds+(1-alpha)*percep)
 * dsr    : (B,1,H,W) gradient of the total loss, or NULL for loss only.
 * Requires H == W and H % 64 == 0.  The un-normalise / re-normalise pair around the down-scaling
 * cancels exactly (the weights sum to one), so mean/std are not parameters. */
int sifnn_loss_fwd_bwd(int kind, const float* sr, const float* ndvi, const float* lst,
                       const float* tab_ds, const float* h12, const float* tab_lp, const float* g9,
                       float alpha, float gamma, double* losses, float* dsr,
                       int B, int H, int W, sifnn_stream_t stream);

/* torch.optim.Adam(lr) step (train_model_B_gradFTM.py:453,121) over one flat buffer.
 * step_count: device int64, incremented by the kernel launch. grad_scale multiplies g first. */
int sifnn_adam_step(float* p, const float* g, float* m, float* v, int64_t* step_count,
                    double lr, double beta1, double beta2, double eps, float grad_scale,
                    int64_t n, sifnn_stream_t stream);

/* Measured fp32 FMA peak helper (roofline denominator for the SIMT kernels): runs a
 * register-resident FFMA loop; writes total FLOPs issued to *flops_out (host). */
int sifnn_fp32_peak_kernel(float* sink, int iters, double* flops_out, sifnn_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Whole-network entry points (ModelB_2.forward, model.py:608-645, and its autograd)
 * ---------------------------------------------------------------------------------- */

typedef struct {
    int in_channels; /* 2 */
    int down[4];     /* 16,32,64,128 */
} sifnn_modelb_cfg;

#define SIFNN_MODELB_NCONV 18
#define SIFNN_MODELB_NBN 17

/* Whether the whole-network entry points may use the tcgen05 kernels on eligible layers (default 1, or 0 when the
 * environment variable SIFNN_DISABLE_TC is set).  0 = "strict fp32": SIMT kernels everywhere. */
void sifnn_set_tensor_cores(int on);
int sifnn_get_tensor_cores(void);

/* Offsets (in floats) of every tensor inside the flat parameter buffer, in
 * module.parameters() order: conv weight, then (gamma, beta) per BatchNorm layer; outlay
 * weight, outlay bias.  w_off[18], gamma_off[17], beta_off[17]; returns total floats
 * (282705 for the shipped configuration).  bn_off[17]: offset of each layer's channel
 * block inside the flat running_mean / running_var buffers; *bn_total = sum of C. */
int64_t sifnn_modelb_param_layout(const sifnn_modelb_cfg* cfg, int64_t* w_off, int64_t* gamma_off,
                                  int64_t* beta_off, int64_t* bias_off, int64_t* bn_off, int64_t* bn_total);

/* Bytes of workspace for a chunk of B patches of H x W.  train != 0 keeps everything the
 * backward pass needs. */
size_t sifnn_modelb_workspace_bytes(const sifnn_modelb_cfg* cfg, int B, int H, int W, int train);

/* Forward.  params: flat parameter buffer; running_mean / running_var: flat BN buffers.
 * train != 0: batch statistics, running buffers updated in place (num_batches_tracked is
 * the caller's job).  x (B,in_channels,H,W) -> y (B,1,H,W). */
int sifnn_modelb_forward(const sifnn_modelb_cfg* cfg, const float* params, float* running_mean,
                         float* running_var, const float* x, float* y, void* workspace,
                         int B, int H, int W, int train, sifnn_stream_t stream);

/* Backward of the last train-mode forward that used `workspace`.  dy (B,1,H,W);
 * grads: flat buffer laid out like params, overwritten.
 * phase: 0 = whole backward, 1 = decoder half only (ub3..ub1 + outlay gradients are final
 * afterwards), 2 = encoder half (must follow phase 1).  Splitting lets the host start the
 * all-reduce of the decoder bucket while the encoder half runs. */
int sifnn_modelb_backward(const sifnn_modelb_cfg* cfg, const float* params, const float* x,
                          const float* dy, float* grads, void* workspace,
                          int B, int H, int W, int phase, sifnn_stream_t stream);

/* Offset (floats) in the flat parameter buffer where the decoder bucket starts. */
int64_t sifnn_modelb_decoder_offset(const sifnn_modelb_cfg* cfg);

#ifdef __cplusplus
}
#endif
#endif /* SIFNN_H */
