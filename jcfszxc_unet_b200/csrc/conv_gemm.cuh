// Internal descriptor of one tap-GEMM launch (see conv_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace unetk {

// Host-side description; pointers are device pointers, already offset to the channel slice.
struct ConvGemmDesc {
  const void* a;      // bf16 NHWC activations, spatial (H*a_step, W*a_step), K channels, pixel stride a_ld
  int64_t a_ld;
  const void* b;      // bf16 packed weights [b_taps][q_groups*ncols][K]
  int b_taps;
  void* out;          // bf16 NHWC, spatial (H*out_step, W*out_step), pixel stride out_ld
  int64_t out_ld;
  const float* bias;  // fp32 [ncols] or null
  int N, H, W;        // grid of GEMM rows (one row per (n,h,w))
  int K;              // reduction channels per tap
  int ncols;          // output channels (per phase q)
  int q_groups;       // 1, or 4 for ConvTranspose fwd (output phase q=(dy,dx) selects weight rows + out pixel)
  int taps;
  int a_step;         // A coordinate = a_step*pos + offset (2 for ConvTranspose dgrad)
  int out_step;       // out coordinate = out_step*pos + phase (2 for ConvTranspose fwd)
  int8_t dh[9], dw[9], btap[9];
};

// Kernel parameter block (passed by value, holds the TMA descriptors).
struct ConvGemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  __nv_bfloat16* out;
  const float* bias;
  int64_t out_ld;
  int H, W, Hout, Wout;
  int TH, TW, tiles_h, tiles_w;
  int num_m_tiles, num_n_tiles, tiles_per_q, rows_per_q, ncols;
  int taps, kchunks, a_step, out_step;
  int8_t dh[9], dw[9], btap[9];
};

int conv_gemm_run(const ConvGemmDesc& d, cudaStream_t stream);

}  // namespace unetk
