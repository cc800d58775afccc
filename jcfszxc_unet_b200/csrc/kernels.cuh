// Internal entry points of the kernel translation units (one *_run per C-ABI function family).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace unetk {
const char* last_error();
int pack_weight_run(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, cudaStream_t stream);
int pack_upconv_weight_run(const float* src, void* dst_fwd, void* dst_dgrad, int Cout, int Cin, cudaStream_t stream);
// wgrad3x3.cu
size_t wgrad3x3_workspace_bytes(int N, int H, int W, int M, int Nn);
size_t wgrad_up_workspace_bytes(int N, int H, int W, int Cin, int Cout);
size_t wgrad3x3_2sm_workspace_bytes(int N, int H, int W, int M, int Nn);
int wgrad3x3_2sm_run(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, float* dw, int accumulate, int N, int H,
                     int W, int M, int Nn, void* workspace, size_t ws_bytes, cudaStream_t stream);
int wgrad_up_run(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate, int N, int H, int W,
                 int Cin, int Cout, void* workspace, size_t ws_bytes, cudaStream_t stream);
int wgrad3x3_run(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, float* dw, int accumulate, int N, int H,
                 int W, int M, int Nn, void* workspace, size_t ws_bytes, cudaStream_t stream, int flip = 0);
// stem.cu
int stem_fwd_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                 void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s);
size_t stem_stats_partial_floats(int N, int H, int W, int Cout);
int stem_fwd_stats_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                       void* y, int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin, int Cout,
                       cudaStream_t s);
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
int sums_to_f32_run(const double* sums, int n, float* out, int accumulate, cudaStream_t s);
int bn_apply_run(const void* raw, int64_t raw_ld, const float* scale, const float* shift, const void* res,
                 int64_t res_ld, void* out, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H, int W, int C,
                 int relu, cudaStream_t s, int ncopies = 0, void* const* copies = nullptr,
                 const int64_t* copies_ld = nullptr);
int maxpool_fwd_run(const void* x, int64_t x_ld, void* y, int64_t y_ld, long long* idx, int N, int H, int W, int C,
                    cudaStream_t s);
int bn_eval_fold_bias_run(int C, const float* gamma, const float* beta, float eps, const float* rm, const float* rv,
                          const float* conv_bias, float* scale, float* shift, cudaStream_t s);
int stem_fwd_affine_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* scale,
                        const float* shift, int relu, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                        cudaStream_t s);
// pad.cu
int shift_copy_run(void* dst, int64_t dst_ld, int Hd, int Wd, const void* src, int64_t src_ld, int Hs, int Ws, int oy, int ox,
                   int N, int C, cudaStream_t s);
// unpool.cu
int maxpool_codes_run(const void* x, int64_t x_ld, void* y, int64_t y_ld, uint8_t* code, int N, int H, int W, int C,
                      cudaStream_t s);
int max_unpool_run(const void* x, int64_t x_ld, const void* where, int is_idx, void* out, int64_t out_ld, int N, int Ho,
                   int Wo, int C, cudaStream_t s);
int max_unpool_bwd_run(const void* dy, int64_t dy_ld, const void* where, int is_idx, void* dx, int64_t dx_ld, int accumulate,
                       int N, int Ho, int Wo, int C, cudaStream_t s);
int maxpool_bwd_run(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld,
                    int accumulate, int N, int H, int W, int C, cudaStream_t s);
int bn_bwd_reduce_run(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                      const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                      double* sums, int N, int H, int W, int C, int relu, cudaStream_t s);
int bn_bwd_apply_run(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                     const float* scale, const float* shift, const float* mean, const float* invstd,
                     const double* sums, double count, float* dgamma, float* dbeta, int accumulate, float* coef,
                     float* dconv_bias, void* draw, int64_t draw_ld, int draw_accumulate, int N, int H, int W, int C,
                     int relu, cudaStream_t s, void* dres = nullptr, int64_t dres_ld = 0, int dres_accumulate = 0);
// loss.cu
size_t head_partial_floats(int64_t npix, int C);
int head_loss_fwd_run(const void* x, int64_t ld, const float* w, const float* bias, const float* labels, float* logits,
                      int post_sigmoid, int64_t npix, int C, float* partial, double* sums, cudaStream_t s);
int loss_finalize_run(const double* sums, double npix_total, float* out, cudaStream_t s);
int head_loss_bwd_run(const void* x, int64_t ld, const float* w, const float* labels, const float* logits,
                      const float* fin, const float* dlogits, float gscale, int post_sigmoid, void* dx, int64_t dx_ld,
                      float* dw, float* db, int accumulate, int64_t npix, int C, float* partial, cudaStream_t s);
size_t bn_head_partial_floats(int64_t npix, int C);
int bn_head_fwd_run(const void* raw, int64_t ld, const float* scale, const float* shift, int relu, const float* w,
                    const float* bias, const float* labels, float* logits, int post_sigmoid, int64_t npix, int C,
                    float* partial, double* sums, cudaStream_t s);
int bn_head_bwd_reduce_run(const void* raw, int64_t ld, const float* scale, const float* shift, const float* mean,
                           int relu, const float* w, const float* labels, const float* logits, const float* fin,
                           const float* dlogits, float gscale, int post_sigmoid, float* dz, float* dw, float* db,
                           int accumulate, double* sums, int64_t npix, int C, float* partial, cudaStream_t s);
int bn_head_bwd_apply_run(const void* raw, int64_t ld, const float* scale, const float* shift, int relu, const float* w,
                          const float* dz, const float* coef, void* draw, int64_t draw_ld, int64_t npix, int C,
                          cudaStream_t s);
// optim.cu
int sqnorm_blocks(int64_t n);
int grad_clip_coef_run(const float* g, int64_t n, float gscale, float max_norm, float* partial, float* out,
                       cudaStream_t s);
int rmsprop_run(float* p, const float* g, float* sq, float* buf, int64_t n, float lr, float alpha, float eps, float wd,
                float momentum, const float* clip, const float* hyper, cudaStream_t s);
// pack.cu
long long pack_tiles(int A, int B);
int pack_weights_run(const long long* table, int n, long long total_tiles, cudaStream_t stream);
// resample.cu
int add_n_run(void* dst, int64_t dst_ld, int accumulate, const void* const* src, const int64_t* src_ld, int nsrc,
              int64_t npix, int C, cudaStream_t s);
int upsample_nearest2x_run(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int backward, int accumulate,
                           int N, int H, int W, int C, cudaStream_t s);
int upsample_bilinear2x_run(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int backward, int accumulate,
                            int N, int H, int W, int C, cudaStream_t s);
int gather_patches_run(const float* images, int64_t si_n, int64_t si_c, int64_t si_h, int64_t si_w, const float* labels,
                       int64_t sl_n, int64_t sl_h, int64_t sl_w, const int* centers, int B, int C, int P, int H, int W,
                       float* out_images, float* out_labels, cudaStream_t s);
int tile_accumulate_run(const float* logits, const int* pos, int B, int P, int H, int W, int apply_sigmoid, double* acc,
                        double* cnt, cudaStream_t s);
int tile_finalize_run(const double* acc, const double* cnt, long long n, double* out, cudaStream_t s);
int copy_f32_strided_run(float* dst, int64_t ds, const float* src, int64_t ss, int64_t n, int accumulate,
                         cudaStream_t s);
// elementwise.cu
int bn_bwd_coef_run(const double* sums, int C, double count, const float* scale, const float* mean,
                    const float* invstd, float* dgamma, float* dbeta, int accumulate, float* coef, float* dconv_bias,
                    cudaStream_t s);
// gate.cu
size_t gate_partial_floats(int64_t npix, int F);
int gate_fwd_run(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                 const float* shg, const float* scx, const float* shx, const float* wpsi, const float* bpsi, float* s,
                 float* partial, double* sums, int64_t npix, int F, cudaStream_t st);
int gate_apply_run(const void* x, int64_t x_ld, const float* s, const float* sc1, const float* sh1, void* out,
                   int64_t out_ld, int64_t npix, int F, cudaStream_t st);
int gate_bwd_psi_run(const void* dout, int64_t dout_ld, const void* x, int64_t x_ld, const float* s, const float* sc1,
                     const float* sh1, const float* mean1, void* dx, int64_t dx_ld, int dx_accumulate, float* dz,
                     float* partial, double* sums, int64_t npix, int F, cudaStream_t st);
int gate_bwd_reduce_run(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                        const float* shg, const float* mug, const float* scx, const float* shx, const float* mux,
                        const float* wpsi, const float* s, const float* dz, const float* sc1, const float* coef1,
                        float* partial, double* sums_g, double* sums_x, float* dwpsi, float* dbpsi, int accumulate,
                        int64_t npix, int F, cudaStream_t st);
int gate_bwd_apply_run(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                       const float* shg, const float* scx, const float* shx, const float* wpsi, const float* s,
                       const float* dz, const float* sc1, const float* coef1, const float* coefg, const float* coefx,
                       void* drawg, int64_t drawg_ld, void* drawx, int64_t drawx_ld, int64_t npix, int F,
                       cudaStream_t st);
// multiclass.cu
size_t head_multi_partial_floats(int64_t npix, int C, int K);
int head_multi_fwd_run(const void* x, int64_t ld, const float* w, const float* bias, float* logits, int N, int64_t hw,
                       int C, int K, cudaStream_t s);
int head_multi_bwd_run(const void* x, int64_t ld, const float* w, const float* dlogits, float gscale, void* dx,
                       int64_t dx_ld, float* dw, float* db, int accumulate, int N, int64_t hw, int C, int K,
                       float* partial, cudaStream_t s);
size_t dice_partial_floats(int64_t groups, int64_t n);
int dice_sums_run(const float* p, const float* t, int64_t groups, int64_t n, float lo, float hi, float* partial,
                  double* sums, cudaStream_t s);
int dice_bwd_run(const float* p, const float* t, const float* coef, const float* gout, int64_t groups, int64_t n,
                 float lo, float hi, float* dp, cudaStream_t s);
// f32path.cu
int f32_pack_split3_run(const float* src, void* dst, long long sr, long long sk, long long st, int R, int K, int T,
                        const int* slices, int ns, cudaStream_t s);
int f32_stem_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                 float* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s);
size_t f32_stats_partial_doubles(long long npix, int C);
int f32_stats_run(const float* x, int64_t ld, int64_t npix, int C, double* partial, double* sums, cudaStream_t s);
int f32_bn_split_run(const float* raw, int64_t raw_ld, const float* scale, const float* shift, void* split,
                     int64_t split_ld, float* out_f32, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H,
                     int W, int C, int relu, cudaStream_t s);
int f32_head_run(const float* x, int64_t ld, const float* w, const float* bias, float* logits, int64_t npix, int C,
                 cudaStream_t s);
}  // namespace unetk
