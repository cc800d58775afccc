// extern "C" boundary (include/unetk.h): argument validation + dispatch to the kernels.
#include "../../include/unetk.h"

#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "wgrad.cuh"

namespace unetk {
const char* last_error();
int probe_run(const void* a, const void* b, float* d, int mode, int shift, int bo, cudaStream_t stream);
int pack_weight_run(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, cudaStream_t stream);
// wgrad3x3.cu
size_t wgrad3x3_workspace_bytes(int N, int H, int W, int M, int Nn);
int wgrad3x3_run(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, float* dw, int accumulate, int N, int H,
                 int W, int M, int Nn, void* workspace, size_t ws_bytes, cudaStream_t stream);
// stem.cu
int stem_fwd_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                 void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s);
size_t stem_wgrad_workspace(int N, int H, int W, int Cin);
int stem_wgrad_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const void* dy, int64_t dy_ld,
                   float* dw, int accumulate, int N, int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes,
                   cudaStream_t s);
// elementwise.cu
size_t chan_partial_floats(int64_t units, int C);
int bn_stats_run(const void* x, int64_t ld, int64_t npix, int C, float* partial, double* sums, cudaStream_t s);
int bn_finalize_run(const double* sums, int C, double count, const float* gamma, const float* beta, float eps,
                    float momentum, float* rm, float* rv, long long* nbt, float* scale, float* shift, float* mean,
                    float* invstd, cudaStream_t s);
int bn_eval_fold_run(int C, const float* gamma, const float* beta, float eps, const float* rm, const float* rv,
                     float* scale, float* shift, float* mean, float* invstd, cudaStream_t s);
int colsum_run(const void* x, int64_t ld, int64_t npix, int C, float* partial, float* out, int accumulate,
               cudaStream_t s);
int bn_apply_run(const void* raw, int64_t raw_ld, const float* scale, const float* shift, void* out, int64_t out_ld,
                 void* pooled, int64_t pooled_ld, int N, int H, int W, int C, int relu, cudaStream_t s);
int maxpool_fwd_run(const void* x, int64_t x_ld, void* y, int64_t y_ld, long long* idx, int N, int H, int W, int C,
                    cudaStream_t s);
int maxpool_bwd_run(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld, int N, int H,
                    int W, int C, cudaStream_t s);
int bn_bwd_reduce_run(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                      const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                      double* sums, int N, int H, int W, int C, int relu, cudaStream_t s);
int bn_bwd_apply_run(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                     const float* scale, const float* shift, const float* mean, const float* invstd,
                     const double* sums, double count, float* dgamma, float* dbeta, int accumulate, float* coef,
                     void* draw, int64_t draw_ld, int N, int H, int W, int C, int relu, cudaStream_t s);
// loss.cu
size_t head_partial_floats(int64_t npix, int C);
int head_loss_fwd_run(const void* x, int64_t ld, const float* w, const float* bias, const float* labels, float* logits,
                      int64_t npix, int C, float* partial, double* sums, cudaStream_t s);
int loss_finalize_run(const double* sums, double npix_total, float* out, cudaStream_t s);
int head_loss_bwd_run(const void* x, int64_t ld, const float* w, const float* labels, const float* logits,
                      const float* fin, const float* dlogits, float gscale, void* dx, int64_t dx_ld, float* dw,
                      float* db, int accumulate, int64_t npix, int C, float* partial, cudaStream_t s);
// optim.cu
int sqnorm_blocks(int64_t n);
int grad_clip_coef_run(const float* g, int64_t n, float gscale, float max_norm, float* partial, float* out,
                       cudaStream_t s);
int rmsprop_run(float* p, const float* g, float* sq, float* buf, int64_t n, float lr, float alpha, float eps, float wd,
                float momentum, const float* clip, cudaStream_t s);
}  // namespace unetk

using namespace unetk;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int unetk_abi_version(void) { return UNETK_ABI_VERSION; }
const char* unetk_last_error(void) { return unetk::last_error(); }
int64_t unetk_launch_count(void) { return static_cast<int64_t>(unetk::launch_count()); }

int unetk_pack_weight(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, void* stream) {
  UNETK_CHECK(src != nullptr && A > 0 && B > 0 && T > 0, -1, "pack_weight: bad arguments");
  return pack_weight_run(src, dst_ab, dst_ba, A, B, T, S(stream));
}

static int conv_fwd_like(const void* x, int64_t x_ld, const void* w, const float* bias, void* y, int64_t y_ld,
                         int N, int H, int W, int K, int ncols, int ksize, bool dgrad, void* stream,
                         float* stats_partial = nullptr, double* stats_sums = nullptr) {
  UNETK_CHECK(x && w && y && N > 0 && H > 0 && W > 0 && K > 0 && ncols > 0, -1, "conv: bad arguments");
  ConvGemmDesc d{};
  d.a = x; d.a_ld = x_ld; d.b = w; d.out = y; d.out_ld = y_ld; d.bias = bias;
  d.stats_partial = stats_partial; d.stats_sums = stats_sums;
  d.N = N; d.H = H; d.W = W; d.K = K; d.ncols = ncols; d.q_groups = 1;
  d.a_step = 1; d.out_step = 1;
  d.taps = ksize * ksize; d.b_taps = d.taps;
  const int half = ksize / 2;
  for (int t = 0; t < d.taps; ++t) {
    const int r = t / ksize, s = t % ksize;
    d.dh[t] = static_cast<int8_t>(dgrad ? half - r : r - half);
    d.dw[t] = static_cast<int8_t>(dgrad ? half - s : s - half);
    d.btap[t] = static_cast<int8_t>(t);
  }
  return conv_gemm_run(d, S(stream));
}

int unetk_conv3x3_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y, int64_t y_ld,
                      int N, int H, int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, H, W, Cin, Cout, 3, false, stream);
}
size_t unetk_conv_stats_partial_floats(int Cout) {
  if (Cout < 8 || Cout % 8) return 0;
  return conv_gemm_stats_partial_floats(Cout);
}
int unetk_conv3x3_fwd_bnstats(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                              int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin, int Cout,
                              void* stream) {
  UNETK_CHECK(partial && sums, -1, "conv3x3_fwd_bnstats: null statistics buffers");
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, H, W, Cin, Cout, 3, false, stream, partial, sums);
}
int unetk_conv3x3_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld, int N, int H,
                        int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(dy, dy_ld, w_pack_t, nullptr, dx, dx_ld, N, H, W, Cout, Cin, 3, true, stream);
}
int unetk_conv1x1_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y, int64_t y_ld,
                      int N, int H, int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, H, W, Cin, Cout, 1, false, stream);
}
int unetk_conv1x1_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld, int N, int H,
                        int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(dy, dy_ld, w_pack_t, nullptr, dx, dx_ld, N, H, W, Cout, Cin, 1, true, stream);
}

// Orientation of the conv weight-gradient GEMM: rows of D come from the operand with >= 128 channels
// when possible (UMMA M is 128), see wgrad.cu.
static WgradDesc conv_wgrad_desc(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                                 int accumulate, int N, int H, int W, int Cin, int Cout, int ksize) {
  WgradDesc d{};
  const int taps = ksize * ksize, half = ksize / 2;
  d.N = N; d.H = H; d.W = W; d.taps = taps; d.p_step = 1; d.q_step = 1;
  d.dw = dw; d.accumulate = accumulate;
  const bool swap = (Cout < 128 && Cin >= 128);
  for (int t = 0; t < taps; ++t) {
    const int8_t oh = static_cast<int8_t>(t / ksize - half), ow = static_cast<int8_t>(t % ksize - half);
    if (!swap) { d.q_dh[t] = oh; d.q_dw[t] = ow; } else { d.p_dh[t] = oh; d.p_dw[t] = ow; }
  }
  if (!swap) {
    d.p = dy; d.p_ld = dy_ld; d.M = Cout; d.q = x; d.q_ld = x_ld; d.Nn = Cin;
    d.dw_sm = static_cast<int64_t>(Cin) * taps; d.dw_sn = taps; d.dw_st = 1;
  } else {
    d.p = x; d.p_ld = x_ld; d.M = Cin; d.q = dy; d.q_ld = dy_ld; d.Nn = Cout;
    d.dw_sm = taps; d.dw_sn = static_cast<int64_t>(Cin) * taps; d.dw_st = 1;
  }
  return d;
}

size_t unetk_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int taps) {
  if (taps == 4) {  // ConvTranspose
    WgradDesc d{};
    d.N = N; d.H = H; d.W = W; d.taps = 4; d.M = Cin; d.Nn = Cout;
    return wgrad_workspace_bytes(d);
  }
  const int ksize = taps == 9 ? 3 : 1;
  WgradDesc d = conv_wgrad_desc(nullptr, 0, nullptr, 0, nullptr, 0, N, H, W, Cin, Cout, ksize);
  size_t need = wgrad_workspace_bytes(d);
  if (taps == 9) {
    const size_t need3 = wgrad3x3_workspace_bytes(N, H, W, Cout, Cin);
    if (need3 > need) need = need3;
  }
  return need;
}

int unetk_conv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                        int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "conv3x3_wgrad: null pointer");
  // halo-reuse kernel (one activation load per filter row); shapes it does not cover use the per-tap kernel
  const int rc3 = wgrad3x3_run(dy, dy_ld, x, x_ld, dw, accumulate, N, H, W, Cout, Cin, workspace, ws_bytes, S(stream));
  if (rc3 <= 0) return rc3;
  WgradDesc d = conv_wgrad_desc(x, x_ld, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout, 3);
  return wgrad_run(d, workspace, ws_bytes, S(stream));
}
int unetk_conv1x1_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                        int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "conv1x1_wgrad: null pointer");
  WgradDesc d = conv_wgrad_desc(x, x_ld, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout, 1);
  return wgrad_run(d, workspace, ws_bytes, S(stream));
}

int unetk_convT2x2_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y, int64_t y_ld,
                       int N, int H, int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(x && w_pack && y, -1, "convT2x2_fwd: null pointer");
  ConvGemmDesc d{};
  d.a = x; d.a_ld = x_ld; d.b = w_pack; d.b_taps = 1; d.out = y; d.out_ld = y_ld; d.bias = bias;
  d.N = N; d.H = H; d.W = W; d.K = Cin; d.ncols = Cout; d.q_groups = 4;
  d.taps = 1; d.a_step = 1; d.out_step = 2;
  d.dh[0] = 0; d.dw[0] = 0; d.btap[0] = 0;
  return conv_gemm_run(d, S(stream));
}
int unetk_convT2x2_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld, int N, int H,
                         int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(dy && w_pack_t && dx, -1, "convT2x2_dgrad: null pointer");
  ConvGemmDesc d{};
  d.a = dy; d.a_ld = dy_ld; d.b = w_pack_t; d.b_taps = 4; d.out = dx; d.out_ld = dx_ld; d.bias = nullptr;
  d.N = N; d.H = H; d.W = W; d.K = Cout; d.ncols = Cin; d.q_groups = 1;
  d.taps = 4; d.a_step = 2; d.out_step = 1;
  for (int t = 0; t < 4; ++t) { d.dh[t] = static_cast<int8_t>(t >> 1); d.dw[t] = static_cast<int8_t>(t & 1); d.btap[t] = static_cast<int8_t>(t); }
  return conv_gemm_run(d, S(stream));
}
int unetk_convT2x2_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                         int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "convT2x2_wgrad: null pointer");
  WgradDesc d{};
  d.N = N; d.H = H; d.W = W; d.taps = 4; d.p_step = 1; d.q_step = 2;
  d.p = x; d.p_ld = x_ld; d.M = Cin; d.q = dy; d.q_ld = dy_ld; d.Nn = Cout;
  for (int t = 0; t < 4; ++t) { d.q_dh[t] = static_cast<int8_t>(t >> 1); d.q_dw[t] = static_cast<int8_t>(t & 1); }
  d.dw = dw; d.accumulate = accumulate;
  d.dw_sm = static_cast<int64_t>(Cout) * 4; d.dw_sn = 4; d.dw_st = 1;
  return wgrad_run(d, workspace, ws_bytes, S(stream));
}

// ------------------------------------------------------------------------------------------------ stem
int unetk_stem_conv3x3_fwd(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                           const float* bias, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                           void* stream) {
  UNETK_CHECK(x && w && y, -1, "stem_fwd: null pointer");
  return stem_fwd_run(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout, S(stream));
}
size_t unetk_stem_wgrad_workspace(int N, int H, int W, int Cin) { return stem_wgrad_workspace(N, H, W, Cin); }
int unetk_stem_conv3x3_wgrad(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const void* dy,
                             int64_t dy_ld, float* dw, int accumulate, int N, int H, int W, int Cin, int Cout,
                             void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "stem_wgrad: null pointer");
  return stem_wgrad_run(x, sn, sc, sh, sw, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout, workspace, ws_bytes,
                        S(stream));
}

// ------------------------------------------------------------------------------------------------ BN / pool
size_t unetk_chan_partial_floats(int64_t units, int C) {
  if (C < 8 || C % 8) return 0;
  return chan_partial_floats(units, C);
}
int unetk_bn_stats(const void* x, int64_t x_ld, int64_t npix, int C, float* partial, double* sums, void* stream) {
  UNETK_CHECK(x && partial && sums && npix > 0, -1, "bn_stats: bad arguments");
  return bn_stats_run(x, x_ld, npix, C, partial, sums, S(stream));
}
int unetk_bn_finalize(const double* sums, int C, double count, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, void* stream) {
  UNETK_CHECK(sums && scale && shift && mean && invstd && count > 0, -1, "bn_finalize: bad arguments");
  return bn_finalize_run(sums, C, count, gamma, beta, eps, momentum, running_mean, running_var,
                         reinterpret_cast<long long*>(num_batches_tracked), scale, shift, mean, invstd, S(stream));
}
int unetk_bn_eval_fold(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                       const float* running_var, float* scale, float* shift, float* mean, float* invstd,
                       void* stream) {
  UNETK_CHECK(running_mean && running_var && scale && shift && mean && invstd, -1, "bn_eval_fold: null pointer");
  return bn_eval_fold_run(C, gamma, beta, eps, running_mean, running_var, scale, shift, mean, invstd, S(stream));
}
int unetk_bn_apply(const void* raw, int64_t raw_ld, const float* scale, const float* shift, void* out, int64_t out_ld,
                   void* pooled, int64_t pooled_ld, int N, int H, int W, int C, int relu, void* stream) {
  UNETK_CHECK(raw && scale && shift && out, -1, "bn_apply: null pointer");
  return bn_apply_run(raw, raw_ld, scale, shift, out, out_ld, pooled, pooled_ld, N, H, W, C, relu, S(stream));
}
int unetk_bn_bwd_reduce(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                        const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                        double* sums, int N, int H, int W, int C, int relu, void* stream) {
  UNETK_CHECK(raw && scale && shift && mean && invstd && partial && sums, -1, "bn_bwd_reduce: null pointer");
  return bn_bwd_reduce_run(raw, raw_ld, g1, g1_ld, gp, gp_ld, scale, shift, mean, invstd, partial, sums, N, H, W, C,
                           relu, S(stream));
}
int unetk_bn_bwd_apply(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                       const float* scale, const float* shift, const float* mean, const float* invstd,
                       const double* sums, double count, float* dgamma, float* dbeta, int accumulate, float* coef,
                       void* draw, int64_t draw_ld, int N, int H, int W, int C, int relu, void* stream) {
  UNETK_CHECK(raw && scale && shift && mean && invstd && sums && coef && draw && count > 0, -1,
              "bn_bwd_apply: bad arguments");
  return bn_bwd_apply_run(raw, raw_ld, g1, g1_ld, gp, gp_ld, scale, shift, mean, invstd, sums, count, dgamma, dbeta,
                          accumulate, coef, draw, draw_ld, N, H, W, C, relu, S(stream));
}
int unetk_maxpool2x2_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t* idx, int N, int H, int W, int C,
                         void* stream) {
  UNETK_CHECK(x && y, -1, "maxpool_fwd: null pointer");
  return maxpool_fwd_run(x, x_ld, y, y_ld, reinterpret_cast<long long*>(idx), N, H, W, C, S(stream));
}
int unetk_maxpool2x2_bwd(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld, int N,
                         int H, int W, int C, void* stream) {
  UNETK_CHECK(x && dy && dx, -1, "maxpool_bwd: null pointer");
  return maxpool_bwd_run(x, x_ld, dy, dy_ld, dx, dx_ld, N, H, W, C, S(stream));
}
int unetk_colsum(const void* x, int64_t x_ld, int64_t npix, int C, float* partial, float* out, int accumulate,
                 void* stream) {
  UNETK_CHECK(x && partial && out && npix > 0, -1, "colsum: bad arguments");
  return colsum_run(x, x_ld, npix, C, partial, out, accumulate, S(stream));
}

// ------------------------------------------------------------------------------------------------ head / loss
size_t unetk_head_partial_floats(int64_t npix, int C) {
  if (C < 8 || C % 8) return 0;
  return head_partial_floats(npix, C);
}
int unetk_head_fwd(const void* x, int64_t x_ld, const float* w, const float* bias, const float* labels, float* logits,
                   int64_t npix, int C, float* partial, double* sums, void* stream) {
  UNETK_CHECK(x && w && logits && partial && npix > 0, -1, "head_fwd: bad arguments");
  return head_loss_fwd_run(x, x_ld, w, bias, labels, logits, npix, C, partial, sums, S(stream));
}
int unetk_loss_finalize(const double* sums, double npix_total, float* fin, void* stream) {
  UNETK_CHECK(sums && fin && npix_total > 0, -1, "loss_finalize: bad arguments");
  return loss_finalize_run(sums, npix_total, fin, S(stream));
}
int unetk_head_bwd(const void* x, int64_t x_ld, const float* w, const float* labels, const float* logits,
                   const float* fin, const float* dlogits, float gscale, void* dx, int64_t dx_ld, float* dw, float* db,
                   int accumulate, int64_t npix, int C, float* partial, void* stream) {
  UNETK_CHECK(x && w && dx && partial && npix > 0, -1, "head_bwd: bad arguments");
  return head_loss_bwd_run(x, x_ld, w, labels, logits, fin, dlogits, gscale, dx, dx_ld, dw, db, accumulate, npix, C,
                           partial, S(stream));
}

// ------------------------------------------------------------------------------------------------ optimizer
size_t unetk_sqnorm_partial_floats(int64_t n) { return static_cast<size_t>(sqnorm_blocks(n)); }
int unetk_grad_clip_coef(const float* g, int64_t n, float gscale, float max_norm, float* partial, float* out,
                         void* stream) {
  UNETK_CHECK(g && partial && out && n > 0, -1, "grad_clip_coef: bad arguments");
  return grad_clip_coef_run(g, n, gscale, max_norm, partial, out, S(stream));
}
int unetk_rmsprop_step(float* p, const float* g, float* square_avg, float* momentum_buf, int64_t n, float lr,
                       float alpha, float eps, float weight_decay, float momentum, const float* clip, void* stream) {
  UNETK_CHECK(p && g && square_avg && n > 0 && (momentum <= 0.f || momentum_buf), -1, "rmsprop_step: bad arguments");
  return rmsprop_run(p, g, square_avg, momentum_buf, n, lr, alpha, eps, weight_decay, momentum, clip, S(stream));
}

int unetk_probe_umma(const void* a, const void* b, float* d, int mode, int shift, int base_offset, void* stream) {
  return probe_run(a, b, d, mode, shift, base_offset, S(stream));
}

}  // extern "C"
