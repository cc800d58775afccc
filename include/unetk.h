/* unetk.h — C ABI of the B200-native U-Net hot path (libunetk.so, sm_100a).
 *
 * The reference (jcfszxc/jcfszxc-UNet) has no FFI of its own: its hot path is the chain of torch.nn
 * primitives inside UNetFamily/utils/unet_parts.py and utils/dice_score.py that ATen dispatches to
 * cuDNN.  Each entry point below replaces one such dispatch (cited as reference file:line) and is what
 * a binding from the reference's Python would call (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *  - Activations are bf16, NHWC ("channels_last" memory, what the reference runs on GPU:
 *    train.py:248-252,525).  A tensor argument is (pointer, ld): pointer to channel 0 of pixel (0,0,0)
 *    of the view, ld = elements between consecutive pixels.  ld > C lets producers write straight
 *    into a channel slice of a concat buffer (torch.cat of unet_parts.py:69 disappears).
 *  - Every buffer is caller-owned device memory; the library never allocates, frees or keeps pointers.
 *  - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; they are re-entrant and
 *    thread-safe (autograd calls backward from worker threads).
 *  - Return 0 on success, negative on error (-1 bad argument, -2 driver/TMA, -3 CUDA runtime);
 *    unetk_last_error() then describes it.  Unsupported shapes are errors, never a fallback.
 */
#ifndef UNETK_H_
#define UNETK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNETK_ABI_VERSION 3

int unetk_abi_version(void);
const char* unetk_last_error(void);
/* Number of CUDA kernels this library has launched (or recorded into a CUDA graph) in this process. */
int64_t unetk_launch_count(void);

/* ---- weight cache -------------------------------------------------------------------------------
 * fp32 master weights stay in PyTorch layout (state_dict compatible); the kernels read bf16 packs.
 * src is [A][B][T] fp32 (Conv2d: A=Cout,B=Cin,T=kh*kw; ConvTranspose2d: A=Cin,B=Cout,T=4).
 * dst_ab is bf16 [T][A][B], dst_ba is bf16 [T][B][A]; either may be NULL. */
int unetk_pack_weight(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, void* stream);
/* Every weight of a model in ONE launch.  table = device int64 [n][8], row = {src, dst_ab, dst_ba, A, B, T,
 * first_tile, mode} where first_tile is the running sum of unetk_pack_tiles(A, B) over the preceding rows and
 * total_tiles the sum over all rows.  T <= 9.  mode 0: the packs above; mode 1 (T = 9): the sub-pixel packs of an
 * up_conv weight (dst_ab = dst_fwd, dst_ba = dst_dgrad of unetk_pack_upconv_weight, 16*A*B elements each). */
int64_t unetk_pack_tiles(int A, int B);
int unetk_pack_weights(const int64_t* table, int n, int64_t total_tiles, void* stream);

/* ---- 3x3 convolution, padding 1, stride 1 (nn.Conv2d, unet_parts.py:24,27 / 85,88 / 103 / 119) ----
 * fwd:   y[n,h,w,co] = bias[co] + sum_{r,s,ci} x[n,h+r-1,w+s-1,ci] * w[co,ci,r,s]
 *        w_pack = bf16 [9][Cout][Cin] (dst_ab of unetk_pack_weight), bias fp32 [Cout] or NULL.
 * dgrad: dx[n,h,w,ci] = sum_{r,s,co} dy[n,h-r+1,w-s+1,co] * w[co,ci,r,s]
 *        w_pack_t = bf16 [9][Cin][Cout] (dst_ba).  accumulate != 0: dx += result (bf16 add in the TMA store;
 *        for tensors with several consumers: recurrent / residual / dense-skip variants).
 * wgrad: dw[co,ci,r,s] (fp32, PyTorch layout) = sum_{n,h,w} dy[n,h,w,co] * x[n,h+r-1,w+s-1,ci]
 *        workspace >= unetk_conv_wgrad_workspace(...) bytes; accumulate!=0 adds into dw.
 * Cin, Cout multiples of 8 (the 3-channel stem has its own entry point below). */
int unetk_conv3x3_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                      int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream);
/* fwd fused with the BatchNorm statistics pass that follows it in DoubleConv (unet_parts.py:24-25): also
 * writes sums = double[2][Cout] (per-channel sum and sum of squares of the bf16 output y), exactly what
 * unetk_bn_stats(y) would produce; partial >= unetk_conv_stats_partial_floats(Cout) floats of scratch. */
size_t unetk_conv_stats_partial_floats(int Cout);
int unetk_conv3x3_fwd_bnstats(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                              int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin,
                              int Cout, void* stream);
int unetk_conv3x3_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                        int accumulate, int N, int H, int W, int Cin, int Cout, void* stream);
/* dgrad of input channels [col0, col0 + ncols) only, written to its own NHWC view dx (ncols channels): the input of a
 * conv that reads a torch.cat (NestedUNet, UNetPP.py:80-97) is made of members whose gradients live in different
 * buffers, so each member's columns go (accumulate != 0: bf16 reduce-add) straight to that member's gradient instead
 * of through a concat-gradient buffer and one add pass per member.  w_pack_t is the whole pack [9][Cin_total][Cout]. */
int unetk_conv3x3_dgrad_cols(const void* dy, int64_t dy_ld, const void* w_pack_t, int Cin_total, int col0, void* dx,
                             int64_t dx_ld, int accumulate, int N, int H, int W, int ncols, int Cout, void* stream);
/* dgrad that also returns sums = double[2][Cin]: per-channel (sum, sum of squares) over all pixels of the bf16 dx it
 * wrote (no accumulate).  When dx is the gradient of a concat buffer [skip | up] (unet_parts.py:69), sums[0][C..2C)
 * IS the bias gradient of the ConvTranspose2d that produced `up` (unet_parts.py:56-58): unetk_sums_to_f32 copies it
 * out, which removes the column-sum pass over that gradient.  partial as for unetk_conv3x3_fwd_bnstats (Cin). */
int unetk_conv3x3_dgrad_colsum(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                               float* partial, double* sums, int N, int H, int W, int Cin, int Cout, void* stream);
/* out[i] (+)= (float)sums[i], i < n. */
int unetk_sums_to_f32(const double* sums, int n, float* out, int accumulate, void* stream);
size_t unetk_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int taps);
int unetk_conv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                        int accumulate, int N, int H, int W, int Cin, int Cout, void* workspace,
                        size_t ws_bytes, void* stream);

/* ---- 3x3 convolution, padding 1, STRIDE 2 (ResidualConv, unet_parts.py:460-462,469; ResUNet.py:32-35) ----
 * x is [N,2Ho,2Wo,Cin], y / dy are [N,Ho,Wo,Cout]:  y[n,h,w,co] = bias[co] + sum x[n,2h+r-1,2w+s-1,ci]*w[co,ci,r,s].
 * Same weight packs, workspace rule and accumulate flags as the stride-1 entry points; the dgrad runs as four
 * sub-pixel classes (1, 2, 2 and 4 taps) that store into the parity views of dx. */
int unetk_conv3x3s2_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y, int64_t y_ld,
                        float* partial, double* sums, int N, int Ho, int Wo, int Cin, int Cout, void* stream);
int unetk_conv3x3s2_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                          int accumulate, int N, int Ho, int Wo, int Cin, int Cout, void* stream);
int unetk_conv3x3s2_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                          int N, int Ho, int Wo, int Cin, int Cout, void* workspace, size_t ws_bytes,
                          void* stream);

/* ---- 1x1 convolution (nn.Conv2d k=1: unet_parts.py:77,143,153,158) on the tensor-core path ------
 * Same contracts with a single tap; w_pack = bf16 [Cout][Cin], w_pack_t = bf16 [Cin][Cout]. */
int unetk_conv1x1_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                      int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_conv1x1_fwd_bnstats(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                              int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin,
                              int Cout, void* stream);
int unetk_conv1x1_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                        int accumulate, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_conv1x1_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                        int accumulate, int N, int H, int W, int Cin, int Cout, void* workspace,
                        size_t ws_bytes, void* stream);

/* ---- ConvTranspose2d kernel 2, stride 2 (unet_parts.py:56-58 Up.up, :481 Upsample) ---------------
 * x is [N,H,W,Cin]; y is [N,2H,2W,Cout] (y_ld lets it be the upper slice of the concat buffer).
 * fwd:   y[n,2h+a,2w+b,co] = bias[co] + sum_ci x[n,h,w,ci] * w[ci,co,a,b]
 *        w_pack = bf16 [4][Cout][Cin] (dst_ba of pack with A=Cin,B=Cout,T=4).
 * dgrad: dx[n,h,w,ci] = sum_{a,b,co} dy[n,2h+a,2w+b,co] * w[ci,co,a,b];  w_pack_t = bf16 [4][Cin][Cout].
 * wgrad: dw[ci,co,a,b] fp32 = sum_{n,h,w} x[n,h,w,ci] * dy[n,2h+a,2w+b,co]. */
int unetk_convT2x2_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                       int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_convT2x2_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                         int accumulate, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_convT2x2_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                         int accumulate, int N, int H, int W, int Cin, int Cout, void* workspace,
                         size_t ws_bytes, void* stream);

/* ---- up_conv: nn.Upsample(scale_factor=2) (nearest) -> nn.Conv2d(k=3,p=1) (unet_parts.py:99-111) ---
 * Sub-pixel form: phase (qy,qx) of the 2x output grid is a 2x2-tap convolution of the LOW-resolution input whose
 * window starts at (qy-1,qx-1); 3x3 taps that read the same low-resolution pixel are pre-summed in fp32 and rounded to
 * bf16 once (qy=0: {kh0},{kh1+kh2}; qy=1: {kh0+kh1},{kh2}; same along kw).  2.25x fewer FLOPs than the conv on the
 * up-sampled tensor, which (like its gradient) is never materialised.  x is [N,H,W,Cin]; y, dy are [N,2H,2W,Cout].
 *   pack_upconv_weight: src = fp32 master [Cout][Cin][3][3] -> dst_fwd bf16 [4 (u*2+v)][4 (qy*2+qx)][Cout][Cin],
 *                       dst_dgrad bf16 [16 (q*4+u*2+v)][Cin][Cout] (either may be NULL).  unetk_pack_weights builds the
 *                       same packs for table rows whose last column is 1.
 *   fwd:    y = bias + conv; partial/sums (both or neither): fused per-channel (sum, sum sq) of the bf16 output as in
 *           conv3x3_fwd_bnstats (partial >= unetk_conv_stats_partial_floats(Cout)).  fwd_affine: eval-mode BN fold.
 *   dgrad:  dx[N,H,W,Cin] (+)= the input gradient (w_up_t = dst_dgrad).
 *   wgrad:  dw fp32 [Cout][Cin][3][3] (+)= the gradient of the 3x3 MASTER weight (the sixteen sub-filter gradients are
 *           folded back by the reduce step); workspace >= unetk_upconv_wgrad_workspace bytes. */
int unetk_pack_upconv_weight(const float* src, void* dst_fwd, void* dst_dgrad, int Cout, int Cin, void* stream);
int unetk_upconv3x3_fwd(const void* x, int64_t x_ld, const void* w_up, const float* bias, void* y, int64_t y_ld,
                        float* partial, double* sums, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_upconv3x3_fwd_affine(const void* x, int64_t x_ld, const void* w_up, const float* scale, const float* shift,
                               int relu, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_upconv3x3_dgrad(const void* dy, int64_t dy_ld, const void* w_up_t, void* dx, int64_t dx_ld, int accumulate,
                          int N, int H, int W, int Cin, int Cout, void* stream);
size_t unetk_upconv_wgrad_workspace(int N, int H, int W, int Cin, int Cout);
int unetk_upconv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                          int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream);

/* ---- stem convolution (network input, Cin <= 4; UNet.py:21 -> unet_parts.py:24) -------------------
 * Reads the fp32 image through arbitrary element strides (sn,sc,sh,sw) — NCHW or channels_last — rounds
 * operands to bf16 like autocast, writes NHWC bf16.  w is the fp32 master weight [Cout][Cin][3][3].
 * wgrad: dw fp32 [Cout][Cin][3][3]; workspace >= unetk_stem_wgrad_workspace bytes. */
int unetk_stem_conv3x3_fwd(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                           const float* bias, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                           void* stream);
/* fwd_bnstats: the same conv, plus sums = double[2][Cout] = per-channel (sum, sum of squares) of the bf16 output —
 * what unetk_bn_stats would return — from the kernel's epilogue; partial >= unetk_stem_stats_partial_floats floats. */
size_t unetk_stem_stats_partial_floats(int N, int H, int W, int Cout);
int unetk_stem_conv3x3_fwd_bnstats(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                                   const float* bias, void* y, int64_t y_ld, float* partial, double* sums, int N, int H,
                                   int W, int Cin, int Cout, void* stream);
size_t unetk_stem_wgrad_workspace(int N, int H, int W, int Cin);
int unetk_stem_conv3x3_wgrad(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const void* dy,
                             int64_t dy_ld, float* dw, int accumulate, int N, int H, int W, int Cin, int Cout,
                             void* workspace, size_t ws_bytes, void* stream);

/* ---- BatchNorm2d (+ReLU) (+MaxPool2d(2)), training and eval (unet_parts.py:25-26,28-29,42-44) -----
 * Two-phase in training because scale/shift depend on whole-batch statistics:
 *   unetk_bn_stats    : sums = double[2][C] (sum, sum of squares) of the raw conv output.
 *                       (data-parallel ranks may all-reduce `sums` here = SyncBN-exact statistics)
 *   unetk_bn_finalize : count = pixels behind `sums`; writes scale = gamma*invstd, shift = beta - mean*scale,
 *                       mean, invstd and updates running_mean/var (momentum, unbiased var) and
 *                       num_batches_tracked (+1) when those pointers are non-NULL.
 *   unetk_bn_eval_fold: eval mode, scale/shift from the running statistics.
 *   unetk_bn_apply    : out = relu?(bf16(raw*scale+shift)) [+ res]; if pooled != NULL also writes the 2x2 max-pool.
 *                       res (optional, bf16 NHWC) is the residual of Recurrent_block / RRCNN_block / ResidualConv
 *                       (unet_parts.py:128,146,475); res and pooled are mutually exclusive.
 * partial: fp32 scratch of >= unetk_chan_partial_floats(units, C) floats (units = pixels). */
size_t unetk_chan_partial_floats(int64_t units, int C);
int unetk_bn_stats(const void* x, int64_t x_ld, int64_t npix, int C, float* partial, double* sums, void* stream);
int unetk_bn_finalize(const double* sums, int C, double count, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, void* stream);
int unetk_bn_eval_fold(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                       const float* running_var, float* scale, float* shift, float* mean, float* invstd,
                       void* stream);
/* Persistent kernels size their grids for the SM count of the current device.  unetk_set_sm_limit(n > 0) makes the
 * launches of the CALLING THREAD use at most n SMs (0 restores all; returns the previous limit): a data-parallel
 * trainer leaves a few SMs to the NCCL all-reduce kernels that run beside the backward — a persistent grid of one CTA
 * per SM otherwise waits for the SMs NCCL holds and its late CTAs double the kernel's duration. */
int unetk_set_sm_limit(int n);
int unetk_device_sms(void);

/* Eval mode, BatchNorm folded into the producing conv's epilogue (north_star: "BatchNorm-fold plus ReLU fused into the
 * epilogue"; what evaluate.py:259-275 runs): the conv writes the post-BN(+ReLU) activation directly, the raw conv
 * output and the bn_apply pass do not exist.
 *   unetk_bn_eval_fold_bias    : scale = gamma*invstd, shift = beta - mean*scale + conv_bias*scale (conv_bias may be NULL)
 *   unetk_conv3x3_fwd_affine   : y = relu?(conv3x3(x, w_pack) * scale + shift), stride 1 or 2, padding 1; N, Ho, Wo = output
 *   unetk_stem_conv3x3_fwd_affine: the same for the image-input conv (fp32 image, fp32 weights [Cout][Cin][3][3])
 * scale / shift: fp32 [Cout], 16-byte aligned. */
int unetk_bn_eval_fold_bias(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                            const float* running_var, const float* conv_bias, float* scale, float* shift, void* stream);
int unetk_conv3x3_fwd_affine(const void* x, int64_t x_ld, const void* w_pack, const float* scale, const float* shift,
                             int relu, void* y, int64_t y_ld, int N, int Ho, int Wo, int Cin, int Cout, int stride,
                             void* stream);
int unetk_stem_conv3x3_fwd_affine(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                                  const float* scale, const float* shift, int relu, void* y, int64_t y_ld, int N, int H,
                                  int W, int Cin, int Cout, void* stream);
int unetk_bn_apply(const void* raw, int64_t raw_ld, const float* scale, const float* shift, const void* res,
                   int64_t res_ld, void* out, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H, int W,
                   int C, int relu, void* stream);
/* bn_apply_copies: bn_apply that also writes the activation to up to three more NHWC views (c0.. packed, NULL at the
 * end) — the node of NestedUNet that is a member of several torch.cat's (UNetPP.py:80-97) leaves the BatchNorm pass
 * in all of them at once instead of being re-read once per copy.  No residual input. */
int unetk_bn_apply_copies(const void* raw, int64_t raw_ld, const float* scale, const float* shift, void* out,
                          int64_t out_ld, void* pooled, int64_t pooled_ld, void* c0, int64_t c0_ld, void* c1, int64_t c1_ld,
                          void* c2, int64_t c2_ld, int N, int H, int W, int C, int relu, void* stream);
/* Backward of out = relu?(bn(raw)).  Incoming gradient = g1 (same resolution; may be NULL) + the scatter of
 * gp (gradient of the 2x2 max-pool of `out`; may be NULL) through the recomputed argmax (first max wins).
 *   unetk_bn_bwd_reduce: sums = double[2][C] (sum g, sum g*xhat)      (SyncBN: all-reduce here)
 *   unetk_bn_bwd_apply : dgamma/dbeta (fp32, optional, accumulate!=0 adds), coef = fp32 scratch [2][C],
 *                        draw = gradient w.r.t. the raw conv output (bf16 NHWC). count = pixels behind sums.
 *                        draw_accumulate != 0: draw += (pre-activation BN whose input has other consumers).
 *                        dconv_bias (optional, fp32 [C]): gradient of the bias of the convolution feeding this
 *                        BatchNorm (conv_block / up_conv / Recurrent_block ..., unet_parts.py:85,103,119): batch
 *                        statistics cancel a per-channel constant, so it is exactly zero and is written as such
 *                        (left untouched when accumulate != 0) instead of reducing d(raw) once more. */
int unetk_bn_bwd_reduce(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp,
                        int64_t gp_ld, const float* scale, const float* shift, const float* mean,
                        const float* invstd, float* partial, double* sums, int N, int H, int W, int C, int relu,
                        void* stream);
int unetk_bn_bwd_apply(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp,
                       int64_t gp_ld, const float* scale, const float* shift, const float* mean,
                       const float* invstd, const double* sums, double count, float* dgamma, float* dbeta,
                       int accumulate, float* coef, float* dconv_bias, void* draw, int64_t draw_ld,
                       int draw_accumulate, int N, int H, int W, int C, int relu, void* stream);
/* unetk_bn_bwd_apply for a unit whose output was relu(bn(raw)) + res (Recurrent_block / RRCNN_block, unet_parts.py:125-146):
 * the same pass also delivers d(res) = g1 into dres (dres_accumulate != 0: dres += g1, one bf16 rounding), instead of a
 * separate add pass over g1.  No fused pool. */
int unetk_bn_bwd_apply_res(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const float* scale,
                           const float* shift, const float* mean, const float* invstd, const double* sums, double count,
                           float* dgamma, float* dbeta, int accumulate, float* coef, float* dconv_bias, void* draw,
                           int64_t draw_ld, int draw_accumulate, void* dres, int64_t dres_ld, int dres_accumulate, int N,
                           int H, int W, int C, int relu, void* stream);
/* The per-channel part of unetk_bn_bwd_apply alone: dgamma/dbeta and coef = [K0[C] | K1[C]] with
 * d(raw) = scale*g + K1*raw + K0 (used by the attention gate's fused backward; C >= 1). */
int unetk_bn_bwd_coef(const double* sums, int C, double count, const float* scale, const float* mean,
                      const float* invstd, float* dgamma, float* dbeta, int accumulate, float* coef,
                      float* dconv_bias, void* stream);

/* ---- MaxPool2d(2) standalone (unet_parts.py:43; indices as F.max_pool2d(return_indices=True)) ----
 * idx (optional) is int64 [N][C][H/2][W/2] holding h*W+w of the selected input element:
 * first maximum in row-major window order, NaN always taken (last NaN wins) — bit-exact with ATen. */
int unetk_maxpool2x2_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t* idx, int N, int H, int W,
                         int C, void* stream);
int unetk_maxpool2x2_bwd(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld,
                         int accumulate, int N, int H, int W, int C, void* stream);

/* ---- SegNet-style pool / unpool pair (UNetFamily/SegNet.py:89-138: F.max_pool2d(return_indices=True) ...
 *      F.max_unpool2d(z, idx, 2, 2)) ---------------------------------------------------------------------------
 * maxpool2x2_fwd_codes: the same pooling rule as maxpool2x2_fwd; the arg-max leaves as a compact CODE, one byte per
 *   pooled element (window position 0..3, row-major), NHWC uint8 [N][H/2][W/2][C] (8-byte aligned): 1 B per element
 *   instead of the 8 B of an int64 index.
 * max_unpool2x2: out [N][2Ho][2Wo][C] <- x [N][Ho][Wo][C]: the selected window position gets the value, the other
 *   three get zero (every output element is written: no zero-fill, no scatter).  `where` is the byte codes
 *   (where_is_idx = 0) or int64 indices [N][C][Ho][Wo] holding h*(2Wo)+w as F.max_pool2d(kernel 2, stride 2) returns
 *   them (where_is_idx = 1; each index lies inside its own window).  Equal to F.max_unpool2d bit for bit.
 * max_unpool2x2_bwd: dx[pooled] (+)= dy[selected position]. */
int unetk_maxpool2x2_fwd_codes(const void* x, int64_t x_ld, void* y, int64_t y_ld, uint8_t* code, int N, int H, int W,
                               int C, void* stream);
int unetk_max_unpool2x2(const void* x, int64_t x_ld, const void* where, int where_is_idx, void* out, int64_t out_ld,
                        int N, int Ho, int Wo, int C, void* stream);
int unetk_max_unpool2x2_bwd(const void* dy, int64_t dy_ld, const void* where, int where_is_idx, void* dx,
                            int64_t dx_ld, int accumulate, int N, int Ho, int Wo, int C, void* stream);

/* ---- F.pad of the Up block (unet_parts.py:64-67) and its backward (a crop) -------------------------------------
 * dst[n,y,x,:] = src[n,y-oy,x-ox,:] where that lies inside src, else 0.  dst [N][Hd][Wd][C] / src [N][Hs][Ws][C] are
 * bf16 NHWC views with pixel strides dst_ld / src_ld.  Forward: (oy, ox) = (diffY/2, diffX/2); backward: negated. */
int unetk_shift_copy(void* dst, int64_t dst_ld, int Hd, int Wd, const void* src, int64_t src_ld, int Hs, int Ws, int oy,
                     int ox, int N, int C, void* stream);

/* ---- per-channel column sum (bias gradients of ConvTranspose2d / biased convs) -------------------- */
int unetk_colsum(const void* x, int64_t x_ld, int64_t npix, int C, float* partial, float* out, int accumulate,
                 void* stream);

/* ---- segmentation head + loss (unet_parts.py:73-79, train.py:264-278, utils/dice_score.py:13-59) ----
 * head_fwd : logits[pix] = bias + sum_c x[pix][c]*w[c] (fp32 out).  With labels != NULL also accumulates
 *            sums = double[4] {sum BCE-with-logits, sum p*y, sum p, sum y}, p = clamp(sigmoid, 1e-7, 1-1e-7).
 *            (data-parallel ranks all-reduce `sums`: the reference's dice is one ratio over the batch)
 * loss_finalize: fin = fp32[8] {loss, bce, dice, 1/npix, cA, cB, -, -}; loss = 0.5*bce + 0.5*(1-dice).
 * head_bwd : dz from (logits, labels, fin) — or from dlogits when the loss was computed outside —
 *            dx[pix][c] = dz*w[c] (bf16), dw[c] = sum dz*x[pix][c], db = sum dz.  gscale multiplies dz.
 * post_sigmoid != 0: the model ends in nn.Sigmoid (ResUNet.py:47-50, UNetPP.py:105-106): logits[] receives
 *            sigmoid(conv) and the loss / backward treat that value as the logit, as train.py:264-278 does.
 * C power of two in [8,256]; partial >= unetk_head_partial_floats(npix, C) floats. */
size_t unetk_head_partial_floats(int64_t npix, int C);
int unetk_head_fwd(const void* x, int64_t x_ld, const float* w, const float* bias, const float* labels,
                   float* logits, int post_sigmoid, int64_t npix, int C, float* partial, double* sums,
                   void* stream);
int unetk_loss_finalize(const double* sums, double npix_total, float* fin, void* stream);
int unetk_head_bwd(const void* x, int64_t x_ld, const float* w, const float* labels, const float* logits,
                   const float* fin, const float* dlogits, float gscale, int post_sigmoid, void* dx, int64_t dx_ld,
                   float* dw, float* db, int accumulate, int64_t npix, int C, float* partial, void* stream);

/* ---- last BatchNorm(+ReLU) folded into the head (DoubleConv -> OutConv, UNet.py:54; unet_parts.py:24-31,73-79) ----
 * The activation in front of OutConv, a = act(bf16(raw*scale + shift)), feeds nothing but the head and its gradient is
 * rank one (dz[pix]*w[c]), so neither a nor d(a) is materialised.
 * bn_head_fwd        : unetk_bn_apply + unetk_head_fwd in one pass over the conv output `raw` (same results).
 * bn_head_bwd_reduce : one pass over `raw`: dz[pix] (fp32, kept for the apply pass), dw/db of the head ((+)= when
 *                      accumulate), and sums = double[2][C] = the two BatchNorm backward sums of unetk_bn_bwd_reduce.
 * bn_head_bwd_apply  : d(raw) from (raw, dz) and coef = [K0 | K1] of unetk_bn_bwd_coef.
 * C = 32 or 64 (the widths in front of OutConv in the U-Net family); partial >= unetk_bn_head_partial_floats(npix, C). */
size_t unetk_bn_head_partial_floats(int64_t npix, int C);
int unetk_bn_head_fwd(const void* raw, int64_t raw_ld, const float* scale, const float* shift, int relu, const float* w,
                      const float* bias, const float* labels, float* logits, int post_sigmoid, int64_t npix, int C,
                      float* partial, double* sums, void* stream);
int unetk_bn_head_bwd_reduce(const void* raw, int64_t raw_ld, const float* scale, const float* shift, const float* mean,
                             int relu, const float* w, const float* labels, const float* logits, const float* fin,
                             const float* dlogits, float gscale, int post_sigmoid, float* dz, float* dw, float* db,
                             int accumulate, double* sums, int64_t npix, int C, float* partial, void* stream);
int unetk_bn_head_bwd_apply(const void* raw, int64_t raw_ld, const float* scale, const float* shift, int relu,
                            const float* w, const float* dz, const float* coef, void* draw, int64_t draw_ld,
                            int64_t npix, int C, void* stream);

/* ---- optimizer tail on flat fp32 buffers (train.py:107-112,299-300) -------------------------------
 * grad_clip_coef: out[0] = gscale*||g||_2, out[1] = gscale*min(1, max_norm/(out[0]+1e-6)); partial >=
 *                 unetk_sqnorm_partial_floats(n) floats.
 * rmsprop_step  : torch.optim.RMSprop (momentum, weight decay, not centered); the gradient is first
 *                 multiplied by clip[1] when clip != NULL. */
size_t unetk_sqnorm_partial_floats(int64_t n);
int unetk_grad_clip_coef(const float* g, int64_t n, float gscale, float max_norm, float* partial, float* out,
                         void* stream);
int unetk_rmsprop_step(float* p, const float* g, float* square_avg, float* momentum_buf, int64_t n, float lr,
                       float alpha, float eps, float weight_decay, float momentum, const float* clip,
                       void* stream);
/* rmsprop_step_dev: the same update with hyper = {lr, alpha, eps, weight_decay, momentum} (5 floats) read from
 *                 DEVICE memory at run time: a captured CUDA graph follows optimizer.param_groups[0]["lr"] changes
 *                 (the reference's ReduceLROnPlateau, train.py:114-122,355).  momentum_buf must not be NULL. */
int unetk_rmsprop_step_dev(float* p, const float* g, float* square_avg, float* momentum_buf, int64_t n,
                           const float* hyper, const float* clip, void* stream);

/* ---- glue of the U-Net variants (bf16 NHWC views, C multiple of 8) ---------------------------------
 * add_n: dst = [dst +] a [+ b [+ c [+ d]]] with every partial sum rounded to bf16 (a chain of bf16 tensor adds:
 *        x + x1 of Recurrent_block / RRCNN_block / ResidualConv, unet_parts.py:128,146,475; ResUNet.py:54).  With
 *        one source it is the slice copy behind NestedUNet's dense torch.cat (UNetPP.py:75-99), and with
 *        accumulate its backward.  Unused sources are NULL.
 * upsample_nearest2x (nn.Upsample(scale_factor=2), unet_parts.py:103): fwd x [N,H,W,C] -> y [N,2H,2W,C];
 *        bwd dx (+)= sum of the 2x2 block of dy.
 * upsample_bilinear2x (mode="bilinear", align_corners=True, UNetPP.py:44): same shapes, ATen's index rule. */
int unetk_add_n(void* dst, int64_t dst_ld, int accumulate, const void* a, int64_t a_ld, const void* b, int64_t b_ld,
                const void* c, int64_t c_ld, const void* d, int64_t d_ld, int64_t npix, int C, void* stream);
int unetk_upsample_nearest2x_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int N, int H, int W, int C,
                                 void* stream);
int unetk_upsample_nearest2x_bwd(const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld, int accumulate, int N, int H,
                                 int W, int C, void* stream);
int unetk_upsample_bilinear2x_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int N, int H, int W, int C,
                                  void* stream);
int unetk_upsample_bilinear2x_bwd(const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld, int accumulate, int N, int H,
                                  int W, int C, void* stream);
/* dst[i*dst_stride] (+)= src[i*src_stride], fp32, i < n: derived weight caches (a 1x1 stem kernel embedded in the
 * centre tap of the 3x3 stem kernel, RRCNN_block.Conv_1x1 on the image, unet_parts.py:143) and their gradients. */
/* Training-batch assembly on the device (train.py:200-253: a Python loop of numpy slices + np.stack + a synchronous
 * H2D copy per step in the reference).  images: fp32 pool, element strides (si_n, si_c, si_h, si_w) of the logical
 * [Nimg, C, H, W] view; labels: fp32 pool [Nimg, H, W] with strides (sl_n, sl_h, sl_w); centers: device int32 [B][3] =
 * (image index, x = index along H, y = index along W) as drawn from the reference's filtered sample map.  Writes
 * out_images fp32 [B][P][P][C] (the channels_last batch of train.py:248-252) and out_labels fp32 [B][P][P]:
 * patch b = rows [x - P/2, x + P/2) x columns [y - P/2, y + P/2).  Bit-exact copies.  labels / out_labels may both
 * be NULL (inference: evaluate.py:72-79 cuts image patches only). */
int unetk_gather_patches(const float* images, int64_t si_n, int64_t si_c, int64_t si_h, int64_t si_w, const float* labels,
                         int64_t sl_n, int64_t sl_h, int64_t sl_w, const int32_t* centers, int B, int C, int P, int H,
                         int W, float* out_images, float* out_labels, void* stream);
/* Sliding-window inference on the device (predict_full_image, evaluate.py:28-96): acc / cnt are double [H][W] maps
 * (zeroed by the caller).  tile_accumulate adds the B patch predictions logits[b] (fp32 [P][P], sigmoid applied when
 * apply_sigmoid != 0) at their top-left corners pos[b] = (y, x) in batch order and counts the coverage;
 * tile_finalize writes out = acc / cnt where cnt != 0, else 0. */
int unetk_tile_accumulate(const float* logits, const int32_t* pos, int B, int P, int H, int W, int apply_sigmoid,
                          double* acc, double* cnt, void* stream);
int unetk_tile_finalize(const double* acc, const double* cnt, int64_t n, double* out, void* stream);
int unetk_copy_f32_strided(float* dst, int64_t dst_stride, const float* src, int64_t src_stride, int64_t n,
                           int accumulate, void* stream);

/* ---- attention gate (Attention_block, unet_parts.py:149-176) -------------------------------------------
 * raw_g / raw_x: outputs of the two 1x1 GEMMs (bf16, F_int channels); (scale, shift, mean) of BN_g / BN_x from
 * unetk_bn_finalize; w_psi fp32 [F_int], b_psi fp32 [1]; BN_1 = the single-channel BatchNorm of psi.
 *   gate_fwd        s[pix] = b_psi + sum_c w_psi[c]*relu(bn_g(raw_g) + bn_x(raw_x))   (fp32 holding bf16 values)
 *                   sums = double[2] (sum s, sum s^2) for BN_1
 *   gate_apply      out = x * sigmoid(s*sc1 + sh1)                                     (F_l channels)
 *   gate_bwd_psi    dx (+)= dout*psi; dz[pix] = (sum_c dout*x) * psi*(1-psi); sums = double[2] (sum dz, sum dz*(s-mean1))
 *   gate_bwd_reduce ds = sc1*dz + K1*s + K0 (coef1 = {K0,K1} from unetk_bn_bwd_coef with C = 1);
 *                   da = ds*w_psi*[a>0]; sums_g / sums_x = double[2][F_int] for unetk_bn_bwd_coef of BN_g / BN_x;
 *                   dw_psi[c] (+)= sum ds*a_c, db_psi (+)= sum ds
 *   gate_bwd_apply  draw_g = sc_g*da + K1g*raw_g + K0g, draw_x likewise (coef_g / coef_x = [K0 | K1])
 * F_int power of two in [8,256]; partial >= unetk_gate_partial_floats(npix, F_int) floats. */
size_t unetk_gate_partial_floats(int64_t npix, int F_int);
int unetk_gate_fwd(const void* raw_g, int64_t raw_g_ld, const void* raw_x, int64_t raw_x_ld, const float* sc_g,
                   const float* sh_g, const float* sc_x, const float* sh_x, const float* w_psi, const float* b_psi,
                   float* s, float* partial, double* sums, int64_t npix, int F_int, void* stream);
int unetk_gate_apply(const void* x, int64_t x_ld, const float* s, const float* sc1, const float* sh1, void* out,
                     int64_t out_ld, int64_t npix, int F_l, void* stream);
int unetk_gate_bwd_psi(const void* dout, int64_t dout_ld, const void* x, int64_t x_ld, const float* s,
                       const float* sc1, const float* sh1, const float* mean1, void* dx, int64_t dx_ld,
                       int dx_accumulate, float* dz, float* partial, double* sums, int64_t npix, int F_l, void* stream);
int unetk_gate_bwd_reduce(const void* raw_g, int64_t raw_g_ld, const void* raw_x, int64_t raw_x_ld, const float* sc_g,
                          const float* sh_g, const float* mean_g, const float* sc_x, const float* sh_x,
                          const float* mean_x, const float* w_psi, const float* s, const float* dz, const float* sc1,
                          const float* coef1, float* partial, double* sums_g, double* sums_x, float* dw_psi,
                          float* db_psi, int accumulate, int64_t npix, int F_int, void* stream);
int unetk_gate_bwd_apply(const void* raw_g, int64_t raw_g_ld, const void* raw_x, int64_t raw_x_ld, const float* sc_g,
                         const float* sh_g, const float* sc_x, const float* sh_x, const float* w_psi, const float* s,
                         const float* dz, const float* sc1, const float* coef1, const float* coef_g,
                         const float* coef_x, void* draw_g, int64_t draw_g_ld, void* draw_x, int64_t draw_x_ld,
                         int64_t npix, int F_int, void* stream);

/* ---- n_classes > 1 and the stand-alone Dice coefficient (unet_parts.py:73-79; utils/dice_score.py:13-59) ----------
 * head_multi_fwd: OutConv with K = n_classes <= 8 outputs: logits[n][k][h][w] (fp32, NCHW as the reference returns it)
 *                 = bias[k] + sum_c x[n,h,w,c] * w[k][c];  hw = H*W, C a power of two in [8,256].
 * head_multi_bwd: given dlogits (same layout, scaled by gscale): dx (bf16 NHWC), dw[k][c], db[k] (fp32, (+)= when
 *                 accumulate != 0); partial >= unetk_head_multi_partial_floats floats.
 * dice_sums:      p, t fp32 [groups][n]; sums = double[groups][3] = (sum clamp(p,lo,hi)*t, sum clamp(p,lo,hi), sum t).
 *                 dice_coeff sums over (H,W) per leading index or over everything (reduce_batch_first);
 *                 multiclass_dice_coeff is the same call on the (batch x class)-flattened tensors.
 * dice_bwd:       dp[g][i] = gout[0] * (coef[g][0]*t + coef[g][1]) where lo <= p <= hi, else 0. */
size_t unetk_head_multi_partial_floats(int64_t npix, int C, int K);
int unetk_head_multi_fwd(const void* x, int64_t x_ld, const float* w, const float* bias, float* logits, int N,
                         int64_t hw, int C, int K, void* stream);
int unetk_head_multi_bwd(const void* x, int64_t x_ld, const float* w, const float* dlogits, float gscale, void* dx,
                         int64_t dx_ld, float* dw, float* db, int accumulate, int N, int64_t hw, int C, int K,
                         float* partial, void* stream);
size_t unetk_dice_partial_floats(int64_t groups, int64_t n);
int unetk_dice_sums(const float* p, const float* t, int64_t groups, int64_t n, float lo, float hi, float* partial,
                    double* sums, void* stream);
int unetk_dice_bwd(const float* p, const float* t, const float* coef, const float* gout, int64_t groups, int64_t n,
                   float lo, float hi, float* dp, void* stream);

/* ---- fp32 mode (BASELINE.json configs[0]: vanilla UNet fp32 forward; logits within 1e-4 of torch fp32) ----------
 * The reference without autocast computes in fp32 (UNet.py:39-55 on unet_parts.py:17-79).  Here an fp32 value is
 * carried as three bf16 terms (hi + mid + lo) and a product as six exact bf16 x bf16 terms accumulated in fp32 on the
 * tensor core.  "Split tensor": bf16 NHWC with 6C channels [hi|hi|hi|mid|mid|lo]; "split pack": bf16 [T][R][6K] with
 * the K axis [hi|mid|lo|hi|mid|hi] per input slice.  Forward only.
 *   f32_pack_split3: dst[t][r][6K] from the fp32 master src[r*sr + k*sk + t*st]; `slices` (HOST array of n_slices ints
 *                    summing to K) are the widths of the channel slices of the consumer's input (concat order).
 *   f32_stem_conv3x3: 3x3 conv of the fp32 image (Cin <= 4, element strides sn,sc,sh,sw) in fp32 FMAs -> fp32 NHWC.
 *   f32_conv3x3 / f32_convT2x2: x_split [N,H,W,Cin6], w_split [9|4][Cout][Cin6] -> y fp32 NHWC (+bias), y_ld in floats.
 *   f32_stats: sums = double[2][C] (sum, sum sq) of an fp32 NHWC tensor; partial >= f32_stats_partial_doubles doubles.
 *              Feed unetk_bn_finalize / unetk_bn_eval_fold as in the bf16 path.
 *   f32_bn_split: v = relu?(raw*scale+shift) (scale/shift NULL: identity); writes any of: split (6C-wide slice),
 *                 out_f32 (C-wide), pooled (split of MaxPool2d(2) of v, first max wins).
 *   f32_head: logits[p] = bias + sum_c x[p,c]*w[c] (OutConv with n_classes == 1). */
int unetk_f32_pack_split3(const float* src, void* dst, int64_t sr, int64_t sk, int64_t st, int R, int K, int T,
                          const int* slices, int n_slices, void* stream);
int unetk_f32_stem_conv3x3(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                           const float* bias, float* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                           void* stream);
int unetk_f32_conv3x3(const void* x_split, int64_t x_ld, const void* w_split, const float* bias, float* y,
                      int64_t y_ld, int N, int H, int W, int Cin6, int Cout, void* stream);
int unetk_f32_convT2x2(const void* x_split, int64_t x_ld, const void* w_split, const float* bias, float* y,
                       int64_t y_ld, int N, int H, int W, int Cin6, int Cout, void* stream);
size_t unetk_f32_stats_partial_doubles(int64_t npix, int C);
int unetk_f32_stats(const float* x, int64_t x_ld, int64_t npix, int C, double* partial, double* sums, void* stream);
int unetk_f32_bn_split(const float* raw, int64_t raw_ld, const float* scale, const float* shift, void* split,
                       int64_t split_ld, float* out_f32, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H,
                       int W, int C, int relu, void* stream);
int unetk_f32_head(const float* x, int64_t x_ld, const float* w, const float* bias, float* logits, int64_t npix, int C,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETK_H_ */
