// Internal descriptor of one weight-gradient launch (see wgrad.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace unetk {

struct WgradDesc {
  const void* p;  // bf16 NHWC, M channels, spatial (H*p_step, W*p_step), pixel stride p_ld  -> rows of D
  int64_t p_ld;
  const void* q;  // bf16 NHWC, Nn channels, spatial (H*q_step, W*q_step), pixel stride q_ld -> cols of D
  int64_t q_ld;
  float* dw;      // fp32 output, element (m, n, tap) at m*dw_sm + n*dw_sn + tap*dw_st
  int64_t dw_sm, dw_sn, dw_st;
  int accumulate; // add into dw instead of overwriting (shared-weight recurrences)
  int N, H, W;    // reduction grid
  int M, Nn, taps;
  int p_step, q_step;
  int8_t p_dh[16], p_dw[16], q_dh[16], q_dw[16];
  // != 0 (taps == 16): the taps are the (phase q, window tap t4) pairs of the sub-pixel up-conv, index q*4 + t4, and the
  // reduce step folds them into the nine taps of the 3x3 filter it came from (wgrad_reduce_upfold_kernel)
  int fold_up;
};

struct WgradParams {
  CUtensorMap tmP;
  CUtensorMap tmQ;
  float* partial;
  int TH, TW, tiles_h, tiles_w, pix_tiles;
  int m_tiles, n_tiles, taps, ksplit;
  int tpn;      // taps packed side by side in one N tile (1: one tap per item); BN = tpn * Nn, n_tiles = 1 when > 1
  int tgroups;  // taps / tpn
  int M, Nn;
  int p_step, q_step;
  int8_t p_dh[16], p_dw[16], q_dh[16], q_dw[16];
};

size_t wgrad_workspace_bytes(const WgradDesc& d);
int wgrad_run(const WgradDesc& d, void* workspace, size_t ws_bytes, cudaStream_t stream);

}  // namespace unetk
