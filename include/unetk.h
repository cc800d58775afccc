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
 *        w_pack = bf16 [4][Cout][Cin] (dst_ba of pack with A=Cin,B=Cout,T=4); Cout multiple of 64.
 * dgrad: dx[n,h,w,ci] = sum_{a,b,co} dy[n,2h+a,2w+b,co] * w[ci,co,a,b];  w_pack_t = bf16 [4][Cin][Cout].
 * wgrad: dw[ci,co,a,b] fp32 = sum_{n,h,w} x[n,h,w,ci] * dy[n,2h+a,2w+b,co]. */
int unetk_convT2x2_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                       int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_convT2x2_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                         int N, int H, int W, int Cin, int Cout, void* stream);
int unetk_convT2x2_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                         int accumulate, int N, int H, int W, int Cin, int Cout, void* workspace,
                         size_t ws_bytes, void* stream);

/* ---- test infrastructure: tcgen05 descriptor-semantics probe (not on the product path) ---------- */
int unetk_probe_umma(const void* a, const void* b, float* d, int mode, int shift, int base_offset,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETK_H_ */
