// Internal descriptor of one tap-GEMM launch (see conv_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "fastdiv.cuh"

namespace unetk {

// Host-side description; pointers are device pointers, already offset to the channel slice.
struct ConvGemmDesc {
  const void* a;      // bf16 NHWC activations, spatial (H*a_step, W*a_step), K channels, pixel stride a_ld
  int64_t a_ld;
  const void* b;      // bf16 packed weights [b_taps][q_groups*ncols][K]
  int b_taps;
  int b_rows;         // rows between two taps of `b` (0 = q_groups*ncols): > ncols when `b` is a row slice of a wider pack
  void* out;          // bf16 NHWC, spatial (H*out_step, W*out_step), pixel stride out_ld
  int64_t out_ld;
  const float* bias;  // fp32 [ncols] or null
  // Eval-mode BatchNorm (+ReLU) folded into the epilogue: out = relu?(acc * scale[c] + bias[c]) with bias = the folded
  // shift (beta - mean*scale + conv_bias*scale).  scale == null: out = acc + bias as before.
  const float* scale;
  int relu;
  int N, H, W;        // grid of GEMM rows (one row per (n,h,w))
  int K;              // reduction channels per tap
  int ncols;          // output channels (per phase q)
  int q_groups;       // 1, or 4 for ConvTranspose fwd (output phase q=(dy,dx) selects weight rows + out pixel)
  int taps;             // <= 16
  int a_step;         // A coordinate = a_step*pos + offset (2 for ConvTranspose dgrad)
  int out_step;       // out coordinate = out_step*pos + phase (2 for ConvTranspose fwd)
  int8_t dh[16], dw[16], btap[16];
  // != 0: phase q = (qy, qx) reads A at dh[t] + q_shift*qy, dw[t] + q_shift*qx (the sub-pixel form of nearest-2x + conv3x3:
  // phase (qy, qx) of the output is a 2x2-tap conv of the low-resolution input whose window starts at (qy-1, qx-1))
  int q_shift;
  // Optional fused BatchNorm statistics of the bf16 output (every output pixel of every phase counted once):
  float* stats_partial;   // scratch, >= conv_gemm_stats_partial_floats(ncols) floats
  double* stats_sums;     // out: double [2][ncols] = per-channel (sum, sum of squares)
  int accumulate;         // != 0: out += result (TMA reduce-add, bf16) instead of out = result
  int out_f32;            // != 0: `out` is fp32 (same NHWC view, out_ld in floats): the fp32-accuracy path of f32path.cu
};

// Kernel parameter block (passed by value, holds the TMA descriptors).
struct ConvGemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmOut[4];  // one per output phase q (only [0] unless ConvTranspose fwd)
  const float* bias;
  const float* scale;    // non-null: affine epilogue (the AFFINE kernel instantiation)
  int relu;
  float* stats_partial;  // [gridDim][2][BN] or null
  int H, W;
  int TH, TW, tw_shift, tiles_h, tiles_w;
  int num_m_tiles, num_n_tiles, tiles_per_q, rows_per_q, ncols;
  FastDiv fd_n_tiles, fd_tiles_w, fd_tiles_h, fd_tiles_per_q;
  int taps, kchunks, a_step;
  int ksteps_last;  // 16-channel MMA steps in the last 64-channel chunk of a tap (1..4)
  int accumulate;   // != 0: epilogue uses cp.reduce.async.bulk.tensor (.add) instead of a plain store
  int l2_prefetch;  // > 0: prefetch the A box of the tile `l2_prefetch` rounds ahead into L2
  float* out_f32;   // non-null: fp32 output written straight from registers (no staging, no statistics)
  long long out_ld; // fp32 path: floats between consecutive output pixels
  int out_step;     // fp32 path: out pixel = out_step * pos + phase (2 for ConvTranspose fwd)
  int q_shift;      // A offset added per output phase: (q >> 1, q & 1) * q_shift
  int8_t dh[16], dw[16], btap[16];
};

size_t conv_gemm_stats_partial_floats(int ncols);
int conv_gemm_run(const ConvGemmDesc& d, cudaStream_t stream);

}  // namespace unetk
