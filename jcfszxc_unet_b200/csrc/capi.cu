// extern "C" boundary (include/unetk.h): argument validation + dispatch to the kernels.
#include "../../include/unetk.h"

#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "wgrad.cuh"

namespace unetk {
const char* last_error();
int probe_run(const void* a, const void* b, float* d, int mode, int shift, int bo, cudaStream_t stream);
int pack_weight_run(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, cudaStream_t stream);
}  // namespace unetk

using namespace unetk;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int unetk_abi_version(void) { return UNETK_ABI_VERSION; }
const char* unetk_last_error(void) { return unetk::last_error(); }

int unetk_pack_weight(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, void* stream) {
  UNETK_CHECK(src != nullptr && A > 0 && B > 0 && T > 0, -1, "pack_weight: bad arguments");
  return pack_weight_run(src, dst_ab, dst_ba, A, B, T, S(stream));
}

static int conv_fwd_like(const void* x, int64_t x_ld, const void* w, const float* bias, void* y, int64_t y_ld,
                         int N, int H, int W, int K, int ncols, int ksize, bool dgrad, void* stream) {
  UNETK_CHECK(x && w && y && N > 0 && H > 0 && W > 0 && K > 0 && ncols > 0, -1, "conv: bad arguments");
  ConvGemmDesc d{};
  d.a = x; d.a_ld = x_ld; d.b = w; d.out = y; d.out_ld = y_ld; d.bias = bias;
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
  return wgrad_workspace_bytes(d);
}

int unetk_conv3x3_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate,
                        int N, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, void* stream) {
  UNETK_CHECK(x && dy && dw, -1, "conv3x3_wgrad: null pointer");
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
  UNETK_CHECK(Cout % 64 == 0, -1, "convT2x2_fwd: Cout=%d must be a multiple of 64", Cout);
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

int unetk_probe_umma(const void* a, const void* b, float* d, int mode, int shift, int base_offset, void* stream) {
  return probe_run(a, b, d, mode, shift, base_offset, S(stream));
}

}  // extern "C"
