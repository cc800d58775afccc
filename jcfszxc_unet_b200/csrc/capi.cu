// extern "C" boundary (include/unetk.h): argument validation + dispatch to the kernels.
#include "../../include/unetk.h"

#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "kernels.cuh"
#include "wgrad.cuh"

using namespace unetk;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// Scratch sizes depend on grid sizes, grid sizes on the SM count the launching thread may use (unetk_set_sm_limit), and
// not monotonically (wave-aware split-K).  Every size query therefore answers for ANY limit a caller may set later:
// the maximum over "all SMs" ... "all but kMaxSmReserve".
namespace {
constexpr int kMaxSmReserve = 16;
template <class F>
size_t max_over_sm_limits(F&& f) {
  const int prev = set_sm_limit(0), real = device_sms();
  size_t m = 0;
  for (int r = 0; r <= kMaxSmReserve && r < real; ++r) {
    set_sm_limit(r == 0 ? 0 : real - r);
    const size_t v = f();
    if (v > m) m = v;
  }
  set_sm_limit(prev);
  return m;
}
}  // namespace

extern "C" {

int unetk_abi_version(void) { return UNETK_ABI_VERSION; }
const char* unetk_last_error(void) { return unetk::last_error(); }
int64_t unetk_launch_count(void) { return static_cast<int64_t>(unetk::launch_count()); }

int unetk_pack_weight(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, void* stream) {
  UNETK_CHECK(src != nullptr && A > 0 && B > 0 && T > 0, -1, "pack_weight: bad arguments");
  return pack_weight_run(src, dst_ab, dst_ba, A, B, T, S(stream));
}

int64_t unetk_pack_tiles(int A, int B) { return (A > 0 && B > 0) ? pack_tiles(A, B) : 0; }
int unetk_pack_weights(const int64_t* table, int n, int64_t total_tiles, void* stream) {
  return pack_weights_run(reinterpret_cast<const long long*>(table), n, total_tiles, S(stream));
}

static WgradDesc conv_wgrad_desc(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw,
                                 int accumulate, int N, int H, int W, int Cin, int Cout, int ksize);

static int conv_fwd_like(const void* x, int64_t x_ld, const void* w, const float* bias, void* y, int64_t y_ld,
                         int N, int H, int W, int K, int ncols, int ksize, bool dgrad, void* stream,
                         float* stats_partial = nullptr, double* stats_sums = nullptr, int accumulate = 0,
                         int a_step = 1, int out_f32 = 0, int b_rows = 0, int col0 = 0) {
  UNETK_CHECK(x && w && y && N > 0 && H > 0 && W > 0 && K > 0 && ncols > 0, -1, "conv: bad arguments");
  UNETK_CHECK(b_rows == 0 || (col0 >= 0 && col0 + ncols <= b_rows && col0 % 8 == 0), -1,
              "conv: column slice [%d, %d) of %d weight rows", col0, col0 + ncols, b_rows);
  ConvGemmDesc d{};
  d.a = x; d.a_ld = x_ld; d.out = y; d.out_ld = y_ld; d.bias = bias;
  // column slice of a wider pack: rows col0 .. col0 + ncols of every tap (the taps stay b_rows rows apart)
  d.b = static_cast<const __nv_bfloat16*>(w) + static_cast<size_t>(col0) * K;
  d.b_rows = b_rows;
  d.stats_partial = stats_partial; d.stats_sums = stats_sums;
  d.N = N; d.H = H; d.W = W; d.K = K; d.ncols = ncols; d.q_groups = 1;
  d.a_step = a_step; d.out_step = 1; d.accumulate = accumulate; d.out_f32 = out_f32;
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
// eval-mode BatchNorm (+ReLU) folded into the conv epilogue: y = relu?(conv(x) * scale + shift), stride 1 or 2
int unetk_conv3x3_fwd_affine(const void* x, int64_t x_ld, const void* w_pack, const float* scale, const float* shift,
                             int relu, void* y, int64_t y_ld, int N, int Ho, int Wo, int Cin, int Cout, int stride,
                             void* stream) {
  UNETK_CHECK(x && w_pack && y && scale && shift && N > 0 && Ho > 0 && Wo > 0 && Cin > 0 && Cout > 0, -1, "conv3x3_fwd_affine: bad arguments");
  UNETK_CHECK(stride == 1 || stride == 2, -1, "conv3x3_fwd_affine: stride %d", stride);
  ConvGemmDesc d{};
  d.a = x; d.a_ld = x_ld; d.out = y; d.out_ld = y_ld; d.bias = shift; d.scale = scale; d.relu = relu;
  d.b = w_pack;
  d.N = N; d.H = Ho; d.W = Wo; d.K = Cin; d.ncols = Cout; d.q_groups = 1;
  d.a_step = stride; d.out_step = 1;
  d.taps = 9; d.b_taps = 9;
  for (int t = 0; t < 9; ++t) {
    d.dh[t] = static_cast<int8_t>(t / 3 - 1);
    d.dw[t] = static_cast<int8_t>(t % 3 - 1);
    d.btap[t] = static_cast<int8_t>(t);
  }
  return conv_gemm_run(d, S(stream));
}
int unetk_stem_conv3x3_fwd_affine(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                                  const float* scale, const float* shift, int relu, void* y, int64_t y_ld, int N, int H,
                                  int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(x && w && y && scale && shift, -1, "stem_conv3x3_fwd_affine: null pointer");
  return stem_fwd_affine_run(x, sn, sc, sh, sw, w, scale, shift, relu, y, y_ld, N, H, W, Cin, Cout, S(stream));
}
int unetk_bn_eval_fold_bias(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                            const float* running_var, const float* conv_bias, float* scale, float* shift, void* stream) {
  UNETK_CHECK(running_mean && running_var && scale && shift && C > 0, -1, "bn_eval_fold_bias: bad arguments");
  return bn_eval_fold_bias_run(C, gamma, beta, eps, running_mean, running_var, conv_bias, scale, shift, S(stream));
}
int unetk_set_sm_limit(int n) {
  const int real = device_sms();
  if (n > 0 && n < real - kMaxSmReserve) n = real - kMaxSmReserve;   // scratch buffers were sized for at most that reserve
  return set_sm_limit(n);
}
int unetk_device_sms(void) { return device_sms(); }
size_t unetk_conv_stats_partial_floats(int Cout) {
  if (Cout < 8 || Cout % 8) return 0;
  return max_over_sm_limits([&] { return conv_gemm_stats_partial_floats(Cout); });
}
int unetk_conv3x3_fwd_bnstats(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                              int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin, int Cout,
                              void* stream) {
  UNETK_CHECK(partial && sums, -1, "conv3x3_fwd_bnstats: null statistics buffers");
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, H, W, Cin, Cout, 3, false, stream, partial, sums);
}
int unetk_conv3x3_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld, int accumulate,
                        int N, int H, int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(dy, dy_ld, w_pack_t, nullptr, dx, dx_ld, N, H, W, Cout, Cin, 3, true, stream, nullptr, nullptr,
                       accumulate);
}
int unetk_conv3x3_dgrad_cols(const void* dy, int64_t dy_ld, const void* w_pack_t, int Cin_total, int col0, void* dx,
                             int64_t dx_ld, int accumulate, int N, int H, int W, int ncols, int Cout, void* stream) {
  return conv_fwd_like(dy, dy_ld, w_pack_t, nullptr, dx, dx_ld, N, H, W, Cout, ncols, 3, true, stream, nullptr, nullptr,
                       accumulate, 1, 0, Cin_total, col0);
}
int unetk_conv3x3_dgrad_colsum(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                               float* partial, double* sums, int N, int H, int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(partial && sums, -1, "conv3x3_dgrad_colsum: null statistics buffers");
  return conv_fwd_like(dy, dy_ld, w_pack_t, nullptr, dx, dx_ld, N, H, W, Cout, Cin, 3, true, stream, partial, sums, 0);
}
int unetk_sums_to_f32(const double* sums, int n, float* out, int accumulate, void* stream) {
  UNETK_CHECK(sums && out && n > 0, -1, "sums_to_f32: bad arguments");
  return sums_to_f32_run(sums, n, out, accumulate, S(stream));
}

// ---- stride 2: forward = the same tap-GEMM with A coordinates 2*pos + (r-1, s-1) (TMA element stride 2)
int unetk_conv3x3s2_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y, int64_t y_ld,
                        float* partial, double* sums, int N, int Ho, int Wo, int Cin, int Cout, void* stream) {
  UNETK_CHECK((partial == nullptr) == (sums == nullptr), -1, "conv3x3s2_fwd: partial and sums go together");
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, Ho, Wo, Cin, Cout, 3, false, stream, partial, sums, 0, 2);
}
// dgrad: dx[2i+ph, 2j+pw] = sum over the taps (r, s) with r = ph+1 (mod 2), s = pw+1 (mod 2) of
//        dy[i + (ph+1-r)/2, j + (pw+1-s)/2] * w[:, :, r, s]  -> one tap-GEMM per sub-pixel class (ph, pw),
//        storing into the parity view of dx (out_step = 2, base offset (ph, pw)).
int unetk_conv3x3s2_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld,
                          int accumulate, int N, int Ho, int Wo, int Cin, int Cout, void* stream) {
  UNETK_CHECK(dy && w_pack_t && dx && N > 0 && Ho > 0 && Wo > 0, -1, "conv3x3s2_dgrad: bad arguments");
  for (int ph = 0; ph < 2; ++ph) {
    for (int pw = 0; pw < 2; ++pw) {
      ConvGemmDesc d{};
      d.a = dy; d.a_ld = dy_ld; d.b = w_pack_t; d.b_taps = 9; d.bias = nullptr;
      d.out = static_cast<uint8_t*>(dx) + (static_cast<int64_t>(ph) * (2 * Wo) + pw) * dx_ld * 2;
      d.out_ld = dx_ld;
      d.N = N; d.H = Ho; d.W = Wo; d.K = Cout; d.ncols = Cin; d.q_groups = 1;
      d.a_step = 1; d.out_step = 2; d.accumulate = accumulate;
      int t = 0;
      for (int r = 0; r < 3; ++r) {
        if (((ph + 1 - r) & 1) != 0) continue;
        for (int s = 0; s < 3; ++s) {
          if (((pw + 1 - s) & 1) != 0) continue;
          d.dh[t] = static_cast<int8_t>((ph + 1 - r) / 2);
          d.dw[t] = static_cast<int8_t>((pw + 1 - s) / 2);
          d.btap[t] = static_cast<int8_t>(r * 3 + s);
          ++t;
        }
      }
      d.taps = t;
      if (int rc = conv_gemm_run(d, S(stream))) return rc;
    }
  }
  return 0;
}
int unetk_conv3x3s2_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                          int N, int Ho, int Wo, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "conv3x3s2_wgrad: null pointer");
  WgradDesc d = conv_wgrad_desc(x, x_ld, dy, dy_ld, dw, accumulate, N, Ho, Wo, Cin, Cout, 3);
  if (d.p == x) d.p_step = 2; else d.q_step = 2;   // the activation operand is read at 2*pos + (r-1, s-1)
  return wgrad_run(d, workspace, ws_bytes, S(stream));
}
int unetk_conv1x1_fwd(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y, int64_t y_ld,
                      int N, int H, int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, H, W, Cin, Cout, 1, false, stream);
}
int unetk_conv1x1_fwd_bnstats(const void* x, int64_t x_ld, const void* w_pack, const float* bias, void* y,
                              int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin, int Cout,
                              void* stream) {
  UNETK_CHECK(partial && sums, -1, "conv1x1_fwd_bnstats: null statistics buffers");
  return conv_fwd_like(x, x_ld, w_pack, bias, y, y_ld, N, H, W, Cin, Cout, 1, false, stream, partial, sums);
}
int unetk_conv1x1_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld, int accumulate,
                        int N, int H, int W, int Cin, int Cout, void* stream) {
  return conv_fwd_like(dy, dy_ld, w_pack_t, nullptr, dx, dx_ld, N, H, W, Cout, Cin, 1, true, stream, nullptr, nullptr,
                       accumulate);
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

static size_t conv_wgrad_workspace_now(int N, int H, int W, int Cin, int Cout, int taps) {
  if (taps == 4) {  // ConvTranspose
    WgradDesc d{};
    d.N = N; d.H = H; d.W = W; d.taps = 4; d.M = Cin; d.Nn = Cout;
    return wgrad_workspace_bytes(d);
  }
  const int ksize = taps == 9 ? 3 : 1;
  WgradDesc d = conv_wgrad_desc(nullptr, 0, nullptr, 0, nullptr, 0, N, H, W, Cin, Cout, ksize);
  size_t need = wgrad_workspace_bytes(d);
  if (taps == 9) {
    const size_t need3 = wgrad3x3_workspace_bytes(N, H, W, Cout, Cin), need3f = wgrad3x3_workspace_bytes(N, H, W, Cin, Cout);
    if (need3 > need) need = need3;
    if (need3f > need) need = need3f;
    const size_t need2 = wgrad3x3_2sm_workspace_bytes(N, H, W, Cout, Cin);
    if (need2 > need) need = need2;
  }
  return need;
}
size_t unetk_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int taps) {
  return max_over_sm_limits([&] { return conv_wgrad_workspace_now(N, H, W, Cin, Cout, taps); });
}

int unetk_conv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                        int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "conv3x3_wgrad: null pointer");
  // halo-reuse kernel (one activation load per filter row); shapes it does not cover use the per-tap kernel
  // Cout <= 64 and more input than output channels: swapped operands (the 128 MMA rows carry input channels, the
  // narrow side goes to N where three taps share one MMA), see wgrad3x3_run
  const bool flip = Cout <= 64 && Cin > Cout && Cin >= 96;
  if (!flip) {
    // >= 256 output channels: CTA pairs, one M = 256 MMA per tap over both SMs (wgrad3x3_2sm.cu)
    const int rc2 = wgrad3x3_2sm_run(dy, dy_ld, x, x_ld, dw, accumulate, N, H, W, Cout, Cin, workspace, ws_bytes, S(stream));
    if (rc2 <= 0) return rc2;
  }
  const int rc3 = flip ? wgrad3x3_run(x, x_ld, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout, workspace, ws_bytes, S(stream), 1)
                       : wgrad3x3_run(dy, dy_ld, x, x_ld, dw, accumulate, N, H, W, Cout, Cin, workspace, ws_bytes, S(stream));
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
int unetk_convT2x2_dgrad(const void* dy, int64_t dy_ld, const void* w_pack_t, void* dx, int64_t dx_ld, int accumulate,
                         int N, int H, int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(dy && w_pack_t && dx, -1, "convT2x2_dgrad: null pointer");
  ConvGemmDesc d{};
  d.a = dy; d.a_ld = dy_ld; d.b = w_pack_t; d.b_taps = 4; d.out = dx; d.out_ld = dx_ld; d.bias = nullptr;
  d.N = N; d.H = H; d.W = W; d.K = Cout; d.ncols = Cin; d.q_groups = 1;
  d.taps = 4; d.a_step = 2; d.out_step = 1; d.accumulate = accumulate;
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

// ---- up_conv (unet_parts.py:99-111): nearest 2x + conv3x3(pad 1) in sub-pixel form — output phase q = (qy, qx) of the
// 2x grid is a 2x2-tap convolution of the LOW-resolution input whose window starts at (qy-1, qx-1), with the 3x3 taps
// that read the same low-resolution pixel pre-summed (pack.cu).  16 tap-GEMMs instead of 36: 2.25x fewer FLOPs, and the
// up-sampled tensor and its gradient never exist.  H, W = the low-resolution size; y / dy are [N,2H,2W,Cout].
static void upconv_fwd_desc(ConvGemmDesc& d, const void* x, int64_t x_ld, const void* w_up, void* y, int64_t y_ld, int N,
                            int H, int W, int Cin, int Cout) {
  d.a = x; d.a_ld = x_ld; d.b = w_up; d.b_taps = 4; d.out = y; d.out_ld = y_ld;
  d.N = N; d.H = H; d.W = W; d.K = Cin; d.ncols = Cout; d.q_groups = 4;
  d.taps = 4; d.a_step = 1; d.out_step = 2; d.q_shift = 1;
  for (int t = 0; t < 4; ++t) {
    d.dh[t] = static_cast<int8_t>((t >> 1) - 1);
    d.dw[t] = static_cast<int8_t>((t & 1) - 1);
    d.btap[t] = static_cast<int8_t>(t);
  }
}
int unetk_pack_upconv_weight(const float* src, void* dst_fwd, void* dst_dgrad, int Cout, int Cin, void* stream) {
  UNETK_CHECK(src != nullptr && Cout > 0 && Cin > 0, -1, "pack_upconv_weight: bad arguments");
  return pack_upconv_weight_run(src, dst_fwd, dst_dgrad, Cout, Cin, S(stream));
}
int unetk_upconv3x3_fwd(const void* x, int64_t x_ld, const void* w_up, const float* bias, void* y, int64_t y_ld,
                        float* partial, double* sums, int N, int H, int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(x && w_up && y && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, -1, "upconv3x3_fwd: bad arguments");
  UNETK_CHECK((partial == nullptr) == (sums == nullptr), -1, "upconv3x3_fwd: partial and sums go together");
  ConvGemmDesc d{};
  upconv_fwd_desc(d, x, x_ld, w_up, y, y_ld, N, H, W, Cin, Cout);
  d.bias = bias; d.stats_partial = partial; d.stats_sums = sums;
  return conv_gemm_run(d, S(stream));
}
int unetk_upconv3x3_fwd_affine(const void* x, int64_t x_ld, const void* w_up, const float* scale, const float* shift,
                               int relu, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(x && w_up && y && scale && shift && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, -1, "upconv3x3_fwd_affine: bad arguments");
  ConvGemmDesc d{};
  upconv_fwd_desc(d, x, x_ld, w_up, y, y_ld, N, H, W, Cin, Cout);
  d.bias = shift; d.scale = scale; d.relu = relu;
  return conv_gemm_run(d, S(stream));
}
// dx[i, j] = sum over (phase q, window tap (u, v)) of dy[2(i - (qy-1+u)) + qy, 2(j - (qx-1+v)) + qx] * w_up[q][u][v]^T:
// sixteen taps on the stride-2 view of dy (TMA element stride 2), offsets 2 - q - 2u in {2, 0, 1, -1}
int unetk_upconv3x3_dgrad(const void* dy, int64_t dy_ld, const void* w_up_t, void* dx, int64_t dx_ld, int accumulate,
                          int N, int H, int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(dy && w_up_t && dx && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, -1, "upconv3x3_dgrad: bad arguments");
  ConvGemmDesc d{};
  d.a = dy; d.a_ld = dy_ld; d.b = w_up_t; d.b_taps = 16; d.out = dx; d.out_ld = dx_ld; d.bias = nullptr;
  d.N = N; d.H = H; d.W = W; d.K = Cout; d.ncols = Cin; d.q_groups = 1;
  d.taps = 16; d.a_step = 2; d.out_step = 1; d.accumulate = accumulate;
  for (int t = 0; t < 16; ++t) {
    const int qy = t >> 3, qx = (t >> 2) & 1, u = (t >> 1) & 1, v = t & 1;
    d.dh[t] = static_cast<int8_t>(2 - qy - 2 * u);
    d.dw[t] = static_cast<int8_t>(2 - qx - 2 * v);
    d.btap[t] = static_cast<int8_t>(t);
  }
  return conv_gemm_run(d, S(stream));
}
static WgradDesc upconv_wgrad_desc(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                                   int N, int H, int W, int Cin, int Cout) {
  WgradDesc d{};
  d.N = N; d.H = H; d.W = W; d.taps = 16; d.p_step = 1; d.q_step = 2; d.fold_up = 1;
  d.p = x; d.p_ld = x_ld; d.M = Cin; d.q = dy; d.q_ld = dy_ld; d.Nn = Cout;
  for (int t = 0; t < 16; ++t) {
    const int qy = t >> 3, qx = (t >> 2) & 1, u = (t >> 1) & 1, v = t & 1;
    d.p_dh[t] = static_cast<int8_t>(qy - 1 + u); d.p_dw[t] = static_cast<int8_t>(qx - 1 + v);
    d.q_dh[t] = static_cast<int8_t>(qy);         d.q_dw[t] = static_cast<int8_t>(qx);
  }
  d.dw = dw; d.accumulate = accumulate;
  d.dw_sm = 9; d.dw_sn = static_cast<int64_t>(Cin) * 9; d.dw_st = 1;   // dw is the 3x3 master's gradient [Cout][Cin][3][3]
  return d;
}
size_t unetk_upconv_wgrad_workspace(int N, int H, int W, int Cin, int Cout) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return 0;
  return max_over_sm_limits([&] {
    const size_t a = wgrad_workspace_bytes(upconv_wgrad_desc(nullptr, 0, nullptr, 0, nullptr, 0, N, H, W, Cin, Cout));
    const size_t b = wgrad_up_workspace_bytes(N, H, W, Cin, Cout);
    return a > b ? a : b;
  });
}
int unetk_upconv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                          int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw && N > 0 && H > 0 && W > 0, -1, "upconv3x3_wgrad: bad arguments");
  // all four taps of a phase from one halo load (wgrad_up.cu); shapes it does not cover (W < 16) use the per-tap kernel
  const int rc = wgrad_up_run(x, x_ld, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout, workspace, ws_bytes, S(stream));
  if (rc <= 0) return rc;
  return wgrad_run(upconv_wgrad_desc(x, x_ld, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout), workspace, ws_bytes, S(stream));
}

// ------------------------------------------------------------------------------------------------ stem
int unetk_stem_conv3x3_fwd(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                           const float* bias, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                           void* stream) {
  UNETK_CHECK(x && w && y, -1, "stem_fwd: null pointer");
  return stem_fwd_run(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout, S(stream));
}
size_t unetk_stem_stats_partial_floats(int N, int H, int W, int Cout) {
  if (N <= 0 || H <= 0 || W <= 0 || Cout < 8 || Cout % 8) return 0;
  return max_over_sm_limits([&] { return stem_stats_partial_floats(N, H, W, Cout); });
}
int unetk_stem_conv3x3_fwd_bnstats(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                                   const float* bias, void* y, int64_t y_ld, float* partial, double* sums, int N, int H,
                                   int W, int Cin, int Cout, void* stream) {
  UNETK_CHECK(x && w && y && partial && sums, -1, "stem_fwd_bnstats: null pointer");
  return stem_fwd_stats_run(x, sn, sc, sh, sw, w, bias, y, y_ld, partial, sums, N, H, W, Cin, Cout, S(stream));
}
size_t unetk_stem_wgrad_workspace(int N, int H, int W, int Cin) {
  return max_over_sm_limits([&] { return stem_wgrad_workspace(N, H, W, Cin); });
}
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
  return max_over_sm_limits([&] { return chan_partial_floats(units, C); });
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
int unetk_bn_apply(const void* raw, int64_t raw_ld, const float* scale, const float* shift, const void* res,
                   int64_t res_ld, void* out, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H, int W, int C,
                   int relu, void* stream) {
  UNETK_CHECK(raw && scale && shift && out, -1, "bn_apply: null pointer");
  return bn_apply_run(raw, raw_ld, scale, shift, res, res_ld, out, out_ld, pooled, pooled_ld, N, H, W, C, relu,
                      S(stream));
}
int unetk_bn_apply_copies(const void* raw, int64_t raw_ld, const float* scale, const float* shift, void* out,
                          int64_t out_ld, void* pooled, int64_t pooled_ld, void* c0, int64_t c0_ld, void* c1, int64_t c1_ld,
                          void* c2, int64_t c2_ld, int N, int H, int W, int C, int relu, void* stream) {
  UNETK_CHECK(raw && scale && shift && out, -1, "bn_apply_copies: null pointer");
  void* cp[3] = {c0, c1, c2};
  const int64_t ld[3] = {c0_ld, c1_ld, c2_ld};
  int n = 0;
  while (n < 3 && cp[n] != nullptr) ++n;
  for (int k = n; k < 3; ++k) UNETK_CHECK(cp[k] == nullptr, -1, "bn_apply_copies: destinations must be packed (NULL only at the end)");
  return bn_apply_run(raw, raw_ld, scale, shift, nullptr, 0, out, out_ld, pooled, pooled_ld, N, H, W, C, relu, S(stream), n,
                      cp, ld);
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
                       float* dconv_bias, void* draw, int64_t draw_ld, int draw_accumulate, int N, int H, int W, int C,
                       int relu, void* stream) {
  UNETK_CHECK(raw && scale && shift && mean && invstd && sums && coef && draw && count > 0, -1,
              "bn_bwd_apply: bad arguments");
  return bn_bwd_apply_run(raw, raw_ld, g1, g1_ld, gp, gp_ld, scale, shift, mean, invstd, sums, count, dgamma, dbeta,
                          accumulate, coef, dconv_bias, draw, draw_ld, draw_accumulate, N, H, W, C, relu, S(stream));
}
int unetk_bn_bwd_apply_res(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const float* scale,
                           const float* shift, const float* mean, const float* invstd, const double* sums, double count,
                           float* dgamma, float* dbeta, int accumulate, float* coef, float* dconv_bias, void* draw,
                           int64_t draw_ld, int draw_accumulate, void* dres, int64_t dres_ld, int dres_accumulate, int N,
                           int H, int W, int C, int relu, void* stream) {
  UNETK_CHECK(raw && g1 && scale && shift && mean && invstd && sums && coef && draw && dres && count > 0, -1,
              "bn_bwd_apply_res: bad arguments");
  return bn_bwd_apply_run(raw, raw_ld, g1, g1_ld, nullptr, 0, scale, shift, mean, invstd, sums, count, dgamma, dbeta,
                          accumulate, coef, dconv_bias, draw, draw_ld, draw_accumulate, N, H, W, C, relu, S(stream), dres,
                          dres_ld, dres_accumulate);
}
int unetk_bn_bwd_coef(const double* sums, int C, double count, const float* scale, const float* mean,
                      const float* invstd, float* dgamma, float* dbeta, int accumulate, float* coef, float* dconv_bias,
                      void* stream) {
  UNETK_CHECK(sums && scale && mean && invstd && coef && count > 0, -1, "bn_bwd_coef: bad arguments");
  return bn_bwd_coef_run(sums, C, count, scale, mean, invstd, dgamma, dbeta, accumulate, coef, dconv_bias, S(stream));
}
int unetk_maxpool2x2_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t* idx, int N, int H, int W, int C,
                         void* stream) {
  UNETK_CHECK(x && y, -1, "maxpool_fwd: null pointer");
  return maxpool_fwd_run(x, x_ld, y, y_ld, reinterpret_cast<long long*>(idx), N, H, W, C, S(stream));
}
int unetk_shift_copy(void* dst, int64_t dst_ld, int Hd, int Wd, const void* src, int64_t src_ld, int Hs, int Ws, int oy,
                     int ox, int N, int C, void* stream) {
  UNETK_CHECK(dst && src && Hd >= 0 && Wd >= 0 && Hs >= 0 && Ws >= 0, -1, "shift_copy: bad arguments");
  return shift_copy_run(dst, dst_ld, Hd, Wd, src, src_ld, Hs, Ws, oy, ox, N, C, S(stream));
}
int unetk_maxpool2x2_fwd_codes(const void* x, int64_t x_ld, void* y, int64_t y_ld, uint8_t* code, int N, int H, int W,
                               int C, void* stream) {
  UNETK_CHECK(x && y && code, -1, "maxpool_fwd_codes: null pointer");
  return maxpool_codes_run(x, x_ld, y, y_ld, code, N, H, W, C, S(stream));
}
int unetk_max_unpool2x2(const void* x, int64_t x_ld, const void* where, int where_is_idx, void* out, int64_t out_ld,
                        int N, int Ho, int Wo, int C, void* stream) {
  UNETK_CHECK(x && where && out, -1, "max_unpool: null pointer");
  return max_unpool_run(x, x_ld, where, where_is_idx, out, out_ld, N, Ho, Wo, C, S(stream));
}
int unetk_max_unpool2x2_bwd(const void* dy, int64_t dy_ld, const void* where, int where_is_idx, void* dx,
                            int64_t dx_ld, int accumulate, int N, int Ho, int Wo, int C, void* stream) {
  UNETK_CHECK(dy && where && dx, -1, "max_unpool_bwd: null pointer");
  return max_unpool_bwd_run(dy, dy_ld, where, where_is_idx, dx, dx_ld, accumulate, N, Ho, Wo, C, S(stream));
}
int unetk_maxpool2x2_bwd(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld,
                         int accumulate, int N, int H, int W, int C, void* stream) {
  UNETK_CHECK(x && dy && dx, -1, "maxpool_bwd: null pointer");
  return maxpool_bwd_run(x, x_ld, dy, dy_ld, dx, dx_ld, accumulate, N, H, W, C, S(stream));
}
int unetk_colsum(const void* x, int64_t x_ld, int64_t npix, int C, float* partial, float* out, int accumulate,
                 void* stream) {
  UNETK_CHECK(x && partial && out && npix > 0, -1, "colsum: bad arguments");
  return colsum_run(x, x_ld, npix, C, partial, out, accumulate, S(stream));
}

// ------------------------------------------------------------------------------------------------ head / loss
size_t unetk_head_partial_floats(int64_t npix, int C) {
  if (C < 8 || C % 8) return 0;
  return max_over_sm_limits([&] { return head_partial_floats(npix, C); });
}
int unetk_head_fwd(const void* x, int64_t x_ld, const float* w, const float* bias, const float* labels, float* logits,
                   int post_sigmoid, int64_t npix, int C, float* partial, double* sums, void* stream) {
  UNETK_CHECK(x && w && logits && partial && npix > 0, -1, "head_fwd: bad arguments");
  return head_loss_fwd_run(x, x_ld, w, bias, labels, logits, post_sigmoid, npix, C, partial, sums, S(stream));
}
int unetk_loss_finalize(const double* sums, double npix_total, float* fin, void* stream) {
  UNETK_CHECK(sums && fin && npix_total > 0, -1, "loss_finalize: bad arguments");
  return loss_finalize_run(sums, npix_total, fin, S(stream));
}
int unetk_head_bwd(const void* x, int64_t x_ld, const float* w, const float* labels, const float* logits,
                   const float* fin, const float* dlogits, float gscale, int post_sigmoid, void* dx, int64_t dx_ld,
                   float* dw, float* db, int accumulate, int64_t npix, int C, float* partial, void* stream) {
  UNETK_CHECK(x && w && dx && partial && npix > 0, -1, "head_bwd: bad arguments");
  return head_loss_bwd_run(x, x_ld, w, labels, logits, fin, dlogits, gscale, post_sigmoid, dx, dx_ld, dw, db,
                           accumulate, npix, C, partial, S(stream));
}

size_t unetk_bn_head_partial_floats(int64_t npix, int C) {
  if (C < 8 || C % 8) return 0;
  return max_over_sm_limits([&] { return bn_head_partial_floats(npix, C); });
}
int unetk_bn_head_fwd(const void* raw, int64_t raw_ld, const float* scale, const float* shift, int relu, const float* w,
                      const float* bias, const float* labels, float* logits, int post_sigmoid, int64_t npix, int C,
                      float* partial, double* sums, void* stream) {
  UNETK_CHECK(raw && scale && shift && w && logits && partial && npix > 0, -1, "bn_head_fwd: bad arguments");
  return bn_head_fwd_run(raw, raw_ld, scale, shift, relu, w, bias, labels, logits, post_sigmoid, npix, C, partial, sums,
                         S(stream));
}
int unetk_bn_head_bwd_reduce(const void* raw, int64_t raw_ld, const float* scale, const float* shift, const float* mean,
                             int relu, const float* w, const float* labels, const float* logits, const float* fin,
                             const float* dlogits, float gscale, int post_sigmoid, float* dz, float* dw, float* db,
                             int accumulate, double* sums, int64_t npix, int C, float* partial, void* stream) {
  UNETK_CHECK(raw && scale && shift && mean && w && dz && sums && partial && npix > 0, -1,
              "bn_head_bwd_reduce: bad arguments");
  return bn_head_bwd_reduce_run(raw, raw_ld, scale, shift, mean, relu, w, labels, logits, fin, dlogits, gscale,
                                post_sigmoid, dz, dw, db, accumulate, sums, npix, C, partial, S(stream));
}
int unetk_bn_head_bwd_apply(const void* raw, int64_t raw_ld, const float* scale, const float* shift, int relu,
                            const float* w, const float* dz, const float* coef, void* draw, int64_t draw_ld,
                            int64_t npix, int C, void* stream) {
  UNETK_CHECK(raw && scale && shift && w && dz && coef && draw && npix > 0, -1, "bn_head_bwd_apply: bad arguments");
  return bn_head_bwd_apply_run(raw, raw_ld, scale, shift, relu, w, dz, coef, draw, draw_ld, npix, C, S(stream));
}

// ------------------------------------------------------------------------------------------------ n_classes > 1, Dice
size_t unetk_head_multi_partial_floats(int64_t npix, int C, int K) {
  if (C < 8 || C % 8 || K < 1 || K > 8) return 0;
  return max_over_sm_limits([&] { return head_multi_partial_floats(npix, C, K); });
}
int unetk_head_multi_fwd(const void* x, int64_t x_ld, const float* w, const float* bias, float* logits, int N,
                         int64_t hw, int C, int K, void* stream) {
  UNETK_CHECK(x && w && logits && N > 0 && hw > 0, -1, "head_multi_fwd: bad arguments");
  return head_multi_fwd_run(x, x_ld, w, bias, logits, N, hw, C, K, S(stream));
}
int unetk_head_multi_bwd(const void* x, int64_t x_ld, const float* w, const float* dlogits, float gscale, void* dx,
                         int64_t dx_ld, float* dw, float* db, int accumulate, int N, int64_t hw, int C, int K,
                         float* partial, void* stream) {
  UNETK_CHECK(x && w && dlogits && dx && partial && N > 0 && hw > 0, -1, "head_multi_bwd: bad arguments");
  return head_multi_bwd_run(x, x_ld, w, dlogits, gscale, dx, dx_ld, dw, db, accumulate, N, hw, C, K, partial, S(stream));
}
size_t unetk_dice_partial_floats(int64_t groups, int64_t n) {
  if (groups < 1 || n < 1) return 0;
  return max_over_sm_limits([&] { return dice_partial_floats(groups, n); });
}
int unetk_dice_sums(const float* p, const float* t, int64_t groups, int64_t n, float lo, float hi, float* partial,
                    double* sums, void* stream) {
  UNETK_CHECK(p && t && partial && sums, -1, "dice_sums: null pointer");
  return dice_sums_run(p, t, groups, n, lo, hi, partial, sums, S(stream));
}
int unetk_dice_bwd(const float* p, const float* t, const float* coef, const float* gout, int64_t groups, int64_t n,
                   float lo, float hi, float* dp, void* stream) {
  UNETK_CHECK(p && t && coef && gout && dp, -1, "dice_bwd: null pointer");
  return dice_bwd_run(p, t, coef, gout, groups, n, lo, hi, dp, S(stream));
}

// ------------------------------------------------------------------------------------------------ optimizer
size_t unetk_sqnorm_partial_floats(int64_t n) {
  return max_over_sm_limits([&] { return static_cast<size_t>(sqnorm_blocks(n)); });
}
int unetk_grad_clip_coef(const float* g, int64_t n, float gscale, float max_norm, float* partial, float* out,
                         void* stream) {
  UNETK_CHECK(g && partial && out && n > 0, -1, "grad_clip_coef: bad arguments");
  return grad_clip_coef_run(g, n, gscale, max_norm, partial, out, S(stream));
}
int unetk_rmsprop_step(float* p, const float* g, float* square_avg, float* momentum_buf, int64_t n, float lr,
                       float alpha, float eps, float weight_decay, float momentum, const float* clip, void* stream) {
  UNETK_CHECK(p && g && square_avg && n > 0 && (momentum <= 0.f || momentum_buf), -1, "rmsprop_step: bad arguments");
  return rmsprop_run(p, g, square_avg, momentum_buf, n, lr, alpha, eps, weight_decay, momentum, clip, nullptr, S(stream));
}
int unetk_rmsprop_step_dev(float* p, const float* g, float* square_avg, float* momentum_buf, int64_t n,
                           const float* hyper, const float* clip, void* stream) {
  UNETK_CHECK(p && g && square_avg && momentum_buf && hyper && n > 0, -1, "rmsprop_step_dev: bad arguments");
  return rmsprop_run(p, g, square_avg, momentum_buf, n, 0.f, 0.f, 0.f, 0.f, 1.f, clip, hyper, S(stream));
}

// ------------------------------------------------------------------------------------------------ variants glue
int unetk_add_n(void* dst, int64_t dst_ld, int accumulate, const void* a, int64_t a_ld, const void* b, int64_t b_ld,
                const void* c, int64_t c_ld, const void* d, int64_t d_ld, int64_t npix, int C, void* stream) {
  UNETK_CHECK(dst && a, -1, "add_n: null pointer");
  const void* src[4] = {a, b, c, d};
  const int64_t ld[4] = {a_ld, b_ld, c_ld, d_ld};
  int n = 1;
  while (n < 4 && src[n] != nullptr) ++n;
  for (int k = n; k < 4; ++k) UNETK_CHECK(src[k] == nullptr, -1, "add_n: sources must be packed (NULL only at the end)");
  return add_n_run(dst, dst_ld, accumulate, src, ld, n, npix, C, S(stream));
}
int unetk_upsample_nearest2x_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int N, int H, int W, int C,
                                 void* stream) {
  UNETK_CHECK(x && y, -1, "upsample_nearest2x_fwd: null pointer");
  return upsample_nearest2x_run(x, x_ld, y, y_ld, 0, 0, N, H, W, C, S(stream));
}
int unetk_upsample_nearest2x_bwd(const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld, int accumulate, int N, int H,
                                 int W, int C, void* stream) {
  UNETK_CHECK(dy && dx, -1, "upsample_nearest2x_bwd: null pointer");
  return upsample_nearest2x_run(dy, dy_ld, dx, dx_ld, 1, accumulate, N, H, W, C, S(stream));
}
int unetk_upsample_bilinear2x_fwd(const void* x, int64_t x_ld, void* y, int64_t y_ld, int N, int H, int W, int C,
                                  void* stream) {
  UNETK_CHECK(x && y, -1, "upsample_bilinear2x_fwd: null pointer");
  return upsample_bilinear2x_run(x, x_ld, y, y_ld, 0, 0, N, H, W, C, S(stream));
}
int unetk_upsample_bilinear2x_bwd(const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld, int accumulate, int N, int H,
                                  int W, int C, void* stream) {
  UNETK_CHECK(dy && dx, -1, "upsample_bilinear2x_bwd: null pointer");
  return upsample_bilinear2x_run(dy, dy_ld, dx, dx_ld, 1, accumulate, N, H, W, C, S(stream));
}
int unetk_gather_patches(const float* images, int64_t si_n, int64_t si_c, int64_t si_h, int64_t si_w, const float* labels,
                         int64_t sl_n, int64_t sl_h, int64_t sl_w, const int32_t* centers, int B, int C, int P, int H,
                         int W, float* out_images, float* out_labels, void* stream) {
  UNETK_CHECK(images && centers && out_images && ((labels == nullptr) == (out_labels == nullptr)), -1,
              "gather_patches: null pointer (labels and out_labels go together)");
  UNETK_CHECK(B > 0 && C > 0 && P > 0 && P % 2 == 0 && H >= P && W >= P, -1, "gather_patches: B=%d C=%d P=%d H=%d W=%d", B, C, P, H, W);
  return gather_patches_run(images, si_n, si_c, si_h, si_w, labels, sl_n, sl_h, sl_w, centers, B, C, P, H, W, out_images,
                            out_labels, S(stream));
}
int unetk_tile_accumulate(const float* logits, const int32_t* pos, int B, int P, int H, int W, int apply_sigmoid,
                          double* acc, double* cnt, void* stream) {
  UNETK_CHECK(logits && pos && acc && cnt && B > 0 && P > 0 && H > 0 && W > 0, -1, "tile_accumulate: bad arguments");
  return tile_accumulate_run(logits, pos, B, P, H, W, apply_sigmoid, acc, cnt, S(stream));
}
int unetk_tile_finalize(const double* acc, const double* cnt, int64_t n, double* out, void* stream) {
  UNETK_CHECK(acc && cnt && out && n > 0, -1, "tile_finalize: bad arguments");
  return tile_finalize_run(acc, cnt, n, out, S(stream));
}
int unetk_copy_f32_strided(float* dst, int64_t dst_stride, const float* src, int64_t src_stride, int64_t n,
                           int accumulate, void* stream) {
  return copy_f32_strided_run(dst, dst_stride, src, src_stride, n, accumulate, S(stream));
}

// ------------------------------------------------------------------------------------------------ attention gate
size_t unetk_gate_partial_floats(int64_t npix, int F_int) {
  return max_over_sm_limits([&] { return gate_partial_floats(npix, F_int); });
}
int unetk_gate_fwd(const void* raw_g, int64_t raw_g_ld, const void* raw_x, int64_t raw_x_ld, const float* sc_g,
                   const float* sh_g, const float* sc_x, const float* sh_x, const float* w_psi, const float* b_psi,
                   float* s, float* partial, double* sums, int64_t npix, int F_int, void* stream) {
  UNETK_CHECK(raw_g && raw_x && sc_g && sh_g && sc_x && sh_x && w_psi && s && partial && sums && npix > 0, -1,
              "gate_fwd: bad arguments");
  return gate_fwd_run(raw_g, raw_g_ld, raw_x, raw_x_ld, sc_g, sh_g, sc_x, sh_x, w_psi, b_psi, s, partial, sums, npix,
                      F_int, S(stream));
}
int unetk_gate_apply(const void* x, int64_t x_ld, const float* s, const float* sc1, const float* sh1, void* out,
                     int64_t out_ld, int64_t npix, int F_l, void* stream) {
  UNETK_CHECK(x && s && sc1 && sh1 && out && npix > 0, -1, "gate_apply: bad arguments");
  return gate_apply_run(x, x_ld, s, sc1, sh1, out, out_ld, npix, F_l, S(stream));
}
int unetk_gate_bwd_psi(const void* dout, int64_t dout_ld, const void* x, int64_t x_ld, const float* s,
                       const float* sc1, const float* sh1, const float* mean1, void* dx, int64_t dx_ld,
                       int dx_accumulate, float* dz, float* partial, double* sums, int64_t npix, int F_l, void* stream) {
  UNETK_CHECK(dout && x && s && sc1 && sh1 && mean1 && dx && dz && partial && sums && npix > 0, -1,
              "gate_bwd_psi: bad arguments");
  return gate_bwd_psi_run(dout, dout_ld, x, x_ld, s, sc1, sh1, mean1, dx, dx_ld, dx_accumulate, dz, partial, sums, npix,
                          F_l, S(stream));
}
int unetk_gate_bwd_reduce(const void* raw_g, int64_t raw_g_ld, const void* raw_x, int64_t raw_x_ld, const float* sc_g,
                          const float* sh_g, const float* mean_g, const float* sc_x, const float* sh_x,
                          const float* mean_x, const float* w_psi, const float* s, const float* dz, const float* sc1,
                          const float* coef1, float* partial, double* sums_g, double* sums_x, float* dw_psi,
                          float* db_psi, int accumulate, int64_t npix, int F_int, void* stream) {
  UNETK_CHECK(raw_g && raw_x && sc_g && sh_g && mean_g && sc_x && sh_x && mean_x && w_psi && s && dz && sc1 && coef1 &&
                  partial && sums_g && sums_x && npix > 0,
              -1, "gate_bwd_reduce: bad arguments");
  return gate_bwd_reduce_run(raw_g, raw_g_ld, raw_x, raw_x_ld, sc_g, sh_g, mean_g, sc_x, sh_x, mean_x, w_psi, s, dz, sc1,
                             coef1, partial, sums_g, sums_x, dw_psi, db_psi, accumulate, npix, F_int, S(stream));
}
int unetk_gate_bwd_apply(const void* raw_g, int64_t raw_g_ld, const void* raw_x, int64_t raw_x_ld, const float* sc_g,
                         const float* sh_g, const float* sc_x, const float* sh_x, const float* w_psi, const float* s,
                         const float* dz, const float* sc1, const float* coef1, const float* coef_g,
                         const float* coef_x, void* draw_g, int64_t draw_g_ld, void* draw_x, int64_t draw_x_ld,
                         int64_t npix, int F_int, void* stream) {
  UNETK_CHECK(raw_g && raw_x && sc_g && sh_g && sc_x && sh_x && w_psi && s && dz && sc1 && coef1 && coef_g && coef_x &&
                  draw_g && draw_x && npix > 0,
              -1, "gate_bwd_apply: bad arguments");
  return gate_bwd_apply_run(raw_g, raw_g_ld, raw_x, raw_x_ld, sc_g, sh_g, sc_x, sh_x, w_psi, s, dz, sc1, coef1, coef_g,
                            coef_x, draw_g, draw_g_ld, draw_x, draw_x_ld, npix, F_int, S(stream));
}

// ------------------------------------------------------------------------------------------------ fp32 mode
int unetk_f32_pack_split3(const float* src, void* dst, int64_t sr, int64_t sk, int64_t st, int R, int K, int T,
                          const int* slices, int n_slices, void* stream) {
  UNETK_CHECK(src && dst && slices && R > 0 && K > 0 && T > 0, -1, "f32_pack_split3: bad arguments");
  return f32_pack_split3_run(src, dst, sr, sk, st, R, K, T, slices, n_slices, S(stream));
}
int unetk_f32_stem_conv3x3(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w,
                           const float* bias, float* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                           void* stream) {
  UNETK_CHECK(x && w && y, -1, "f32_stem_conv3x3: null pointer");
  return f32_stem_run(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout, S(stream));
}
int unetk_f32_conv3x3(const void* x_split, int64_t x_ld, const void* w_split, const float* bias, float* y,
                      int64_t y_ld, int N, int H, int W, int Cin6, int Cout, void* stream) {
  return conv_fwd_like(x_split, x_ld, w_split, bias, y, y_ld, N, H, W, Cin6, Cout, 3, false, stream, nullptr, nullptr,
                       0, 1, 1);
}
int unetk_f32_convT2x2(const void* x_split, int64_t x_ld, const void* w_split, const float* bias, float* y,
                       int64_t y_ld, int N, int H, int W, int Cin6, int Cout, void* stream) {
  UNETK_CHECK(x_split && w_split && y, -1, "f32_convT2x2: null pointer");
  ConvGemmDesc d{};
  d.a = x_split; d.a_ld = x_ld; d.b = w_split; d.b_taps = 1; d.out = y; d.out_ld = y_ld; d.bias = bias;
  d.N = N; d.H = H; d.W = W; d.K = Cin6; d.ncols = Cout; d.q_groups = 4;
  d.taps = 1; d.a_step = 1; d.out_step = 2; d.out_f32 = 1;
  d.dh[0] = 0; d.dw[0] = 0; d.btap[0] = 0;
  return conv_gemm_run(d, S(stream));
}
size_t unetk_f32_stats_partial_doubles(int64_t npix, int C) {
  return max_over_sm_limits([&] { return f32_stats_partial_doubles(npix, C); });
}
int unetk_f32_stats(const float* x, int64_t x_ld, int64_t npix, int C, double* partial, double* sums, void* stream) {
  UNETK_CHECK(x && partial && sums && npix > 0, -1, "f32_stats: bad arguments");
  return f32_stats_run(x, x_ld, npix, C, partial, sums, S(stream));
}
int unetk_f32_bn_split(const float* raw, int64_t raw_ld, const float* scale, const float* shift, void* split,
                       int64_t split_ld, float* out_f32, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H,
                       int W, int C, int relu, void* stream) {
  UNETK_CHECK(raw && (split || out_f32 || pooled) && N > 0 && H > 0 && W > 0, -1, "f32_bn_split: bad arguments");
  return f32_bn_split_run(raw, raw_ld, scale, shift, split, split_ld, out_f32, out_ld, pooled, pooled_ld, N, H, W, C,
                          relu, S(stream));
}
int unetk_f32_head(const float* x, int64_t x_ld, const float* w, const float* bias, float* logits, int64_t npix, int C,
                   void* stream) {
  UNETK_CHECK(x && w && logits && npix > 0 && C > 0, -1, "f32_head: bad arguments");
  return f32_head_run(x, x_ld, w, bias, logits, npix, C, S(stream));
}

}  // extern "C"
