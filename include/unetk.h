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

#define UNETK_ABI_VERSION 1

int unetk_abi_version(void);
const char* unetk_last_error(void);
/* Number of CUDA kernels this library has launched (or recorded into a CUDA graph) in this process. */
int64_t unetk_launch_count(void);

/* ---- weight cache -------------------------------------------------------------------------------
 * fp32 master weights stay in PyTorch layout (state_dict compatible); the kernels read bf16 packs.
 * src is [A][B][T] fp32 (Conv2d: A=Cout,B=Cin,T=kh*kw; ConvTranspose2d: A=Cin,B=Cout,T=4).
 * dst_ab is bf16 [T][A][B], dst_ba is bf16 [T][B][A]; either may be NULL. */
int unetk_pack_weight(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, void* stream);

/* ---- 3x3 convolution, padding 1, stride 1 (nn.Conv2d, unet_parts.py:24,27 / 85,88 / 103 / 119) ----
 * fwd:   y[n,h,w,co] = bias[co] + sum_{r,s,ci} x[n,h+r-1,w+s-1,ci] * w[co,ci,r,s]
 *        w_pack = bf16 [9][Cout][Cin] (dst_ab of unetk_pack_weight), bias fp32 [Cout] or NULL.
 * dgrad: dx[n,h,w,ci] = sum_{r,s,co} dy[n,h-r+1,w-s+1,co] * w[co,ci,r,s]
 *        w_pack_t = bf16 [9][Cin][Cout] (dst_ba).
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
                        int N, int H, int W, int Cin, int Cout, void* stream);
size_t unetk_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int taps);
int unetk_conv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                        int accumulate, int N, int H, int W, int Cin, int Cout, void* workspace,
                        size_t ws_bytes, void* stream);

/* ---- 1x1 convolution (nn.Conv2d k=1: unet_parts.py:77,143,153,158) on the tensor-core path ------
 * Same contracts with a single tap; w_pack = bf16 [Cout][Cin], w_pack_t = bf16 [Cin][Cout]. */
int unetk_conv1x1_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                      int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_conv1x1_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                        int N, int H, int W, int Cin, int Cout, void* stream);
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
                         int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_convT2x2_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                         int accumulate, int N, int H, int W, int Cin, int Cout, void* workspace,
                         size_t ws_bytes, void* stream);

/* ---- stem convolution (network input, Cin <= 4; UNet.py:21 -> unet_parts.py:24) -------------------
 * Reads the fp32 image through arbitrary element strides (sn,sc,sh,sw) — NCHW or channels_last — rounds
 * operands to bf16 like autocast, writes NHWC bf16.  w is the fp32 master weight [Cout][Cin][3][3].
 * wgrad: dw fp32 [Cout][Cin][3][3]; workspace >= unetk_stem_wgrad_workspace bytes. */
int unetk_stem_conv3x3_fwd(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                           const float* bias, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                           void* stream);
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
 *   unetk_bn_apply    : out = relu?(bf16(raw*scale+shift)); if pooled != NULL also writes the 2x2 max-pool.
 * partial: fp32 scratch of >= unetk_chan_partial_floats(units, C) floats (units = pixels). */
size_t unetk_chan_partial_floats(int64_t units, int C);
int unetk_bn_stats(const void* x, int64_t x_ld, int64_t npix, int C, float* partial, double* sums, void* stream);
int unetk_bn_finalize(const double* sums, int C, double count, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, void* stream);
int unetk_bn_eval_fold(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                       const float* running_var, float* scale, float* shift, float* mean, float* invstd,
                       void* stream);
int unetk_bn_apply(const void* raw, int64_t raw_ld, const float* scale, const float* shift, void* out,
                   int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H, int W, int C, int relu,
                   void* stream);
/* Backward of out = relu?(bn(raw)).  Incoming gradient = g1 (same resolution; may be NULL) + the scatter of
 * gp (gradient of the 2x2 max-pool of `out`; may be NULL) through the recomputed argmax (first max wins).
 *   unetk_bn_bwd_reduce: sums = double[2][C] (sum g, sum g*xhat)      (SyncBN: all-reduce here)
 *   unetk_bn_bwd_apply : dgamma/dbeta (fp32, optional, accumulate!=0 adds), coef = fp32 scratch [2][C],
 *                        draw = gradient w.r.t. the raw conv output (bf16 NHWC). count = pixels behind sums. */
int unetk_bn_bwd_reduce(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp,
                        int64_t gp_ld, const float* scale, const float* shift, const float* mean,
                        const float* invstd, float* partial, double* sums, int N, int H, int W, int C, int relu,
                        void* stream);
int unetk_bn_bwd_apply(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp,
                       int64_t gp_ld, const float* scale, const float* shift, const float* mean,
                       const float* invstd, const double* sums, double count, float* dgamma, float* dbeta,
                       int accumulate, float* coef, void* draw, int64_t draw_ld, int N, int H, int W, int C,
                       int relu, void* stream);

/* ---- MaxPool2d(2) standalone (unet_parts.py:43; indices as F.max_pool2d(return_indices=True)) ----
 * idx (optional) is int64 [N][C][H/2][W/2] holding h*W+w of the selected input element:
 * first maximum in row-major window order, NaN always taken (last NaN wins) — bit-exact with ATen. */
int unetk_maxpool2x2_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t* idx, int N, int H, int W,
                         int C, void* stream);
int unetk_maxpool2x2_bwd(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld,
                         int N, int H, int W, int C, void* stream);

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
 * C power of two in [8,256]; partial >= unetk_head_partial_floats(npix, C) floats. */
size_t unetk_head_partial_floats(int64_t npix, int C);
int unetk_head_fwd(const void* x, int64_t x_ld, const float* w, const float* bias, const float* labels,
                   float* logits, int64_t npix, int C, float* partial, double* sums, void* stream);
int unetk_loss_finalize(const double* sums, double npix_total, float* fin, void* stream);
int unetk_head_bwd(const void* x, int64_t x_ld, const float* w, const float* labels, const float* logits,
                   const float* fin, const float* dlogits, float gscale, void* dx, int64_t dx_ld, float* dw,
                   float* db, int accumulate, int64_t npix, int C, float* partial, void* stream);

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

/* ---- test infrastructure: tcgen05 descriptor-semantics probe (not on the product path) ---------- */
int unetk_probe_umma(const void* a, const void* b, float* d, int mode, int shift, int base_offset,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETK_H_ */
