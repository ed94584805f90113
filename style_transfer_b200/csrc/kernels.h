// Host-side launchers of the CUDA kernels (implemented in kernels_simt.cu / conv_tc.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

// Where the first convolution reads its pixels: an un-rolled image [3][H][W] addressed circularly
// (the virtual roll) so that no tile copy is ever materialised.
struct ImageView {
  const float* base;   // [3][H][W]
  int H, W;            // full image size
  int oy, ox;          // canonical coordinate of tile pixel (0,0), before wrapping
};

// Scratch for deterministic reductions, shared by all kernels of one context (single stream).
struct ReduceScratch {
  double* partials;    // kMaxReduceBlocks * 4 doubles
  unsigned* counter;   // zero between kernels
};
constexpr int kMaxReduceBlocks = 1 << 17;

enum BwdEpilogue { kEpiNone = 0, kEpiMask = 1, kEpiInj = 2 };

// ---- convolutions (NHWC activations of type T) -----------------------------------------------
template <typename T>
int conv_first_fwd(const ImageView& img, int h, int w, const float* wpack /*[27][cout]*/,
                   const float* bias, T* out, int cout, cudaStream_t s);
template <typename T>
int conv_last_bwd(const T* dz, int h, int w, int cin_of_dz, const float* wpack /*[27][cout]*/,
                  float* grad, long plane_stride, long row_stride, cudaStream_t s);
// out[p][co] = epilogue(sum_{tap,ci} in[p+tap][ci] * wpack[tap][ci][co])
template <typename T>
int conv3x3_simt(const T* in, const float* wpack, const float* bias, T* out, int h, int w, int cin,
                 int cout, bool forward, const T* mask_act, const T* inj, cudaStream_t s);

// ---- pooling -----------------------------------------------------------------------------------
template <typename T>
int pool_fwd(const T* in, T* out, int h, int w, int c, bool is_max, cudaStream_t s);
// d_in = [mask](in>0) * pool_bwd(d_out) + [inj]
template <typename T>
int pool_bwd(const T* d_out, const T* in, T* d_in, int h, int w, int c, bool is_max,
             bool apply_mask, const T* inj, cudaStream_t s);

// ---- Gram / style ------------------------------------------------------------------------------
// gram_full[C][C] (symmetric, float) = F^T F / (C*HW).  F is NHWC [hw][c] (channel_major=false) or
// [c][hw] (channel_major=true).  part: scratch for split-K partials (part_floats floats).
template <typename T>
int gram_full(const T* f, int hw, int c, bool channel_major, float* gram, float* part,
              size_t part_floats, int sm_count, cudaStream_t s);
// delta = gram - target (both symmetric, full); *loss_accum += w * 0.5 * sum_{j<=i} delta_ij^2
int gram_delta(const float* gram, const float* target, float* delta, __nv_bfloat16* delta_bf16,
               int c, double w, double* loss_accum, ReduceScratch rs, cudaStream_t s);
int sum_partials(const double* partials, int n, double* out, cudaStream_t s);
// S[p][co] = sum_ci F[p][ci] * delta[ci][co]; *sum_abs = sum |S|
template <typename T>
int style_grad(const T* f, const float* delta, T* s_out, int hw, int c, double* sum_abs,
               ReduceScratch rs, cudaStream_t s);
// inj = (accumulate ? inj : 0) + w / (*sum_abs / n + EPS) * src
template <typename T>
int inject_scaled(T* inj, const T* src, size_t n, float w, const double* sum_abs, bool accumulate,
                  cudaStream_t s);
int symmetrize_lower(const float* lower, float* full, int c, cudaStream_t s);
int extract_lower(const float* full, float* lower, int c, cudaStream_t s);

// ---- content / deep-dream ------------------------------------------------------------------------
// Target map tgt is NHWC float [Hf][Wf][C] of the whole image, read at ((ty0+y) mod Hf, (tx0+x) mod
// Wf).  stats[0] = sum c^2, stats[1] = sum |c| with c = F - target (target==nullptr: c = F).
template <typename T>
int diff_stats(const T* f, int hf, int wf, int c, const float* tgt, int Hf, int Wf, int ty0,
               int tx0, double* stats, ReduceScratch rs, cudaStream_t s);
// inj = (accumulate ? inj : 0) + w / (stats[1]/n + EPS) * (F - target);
// *loss_accum += loss_w * 0.5 * stats[0]
template <typename T>
int diff_inject(const T* f, int hf, int wf, int c, const float* tgt, int Hf, int Wf, int ty0,
                int tx0, const double* stats, float w, double loss_w, double* loss_accum, T* inj,
                bool accumulate, cudaStream_t s);

// ---- layout conversion ---------------------------------------------------------------------------
template <typename T>
int nhwc_to_nchw_f32(const T* in, float* out, int hw, int c, cudaStream_t s);
int nchw_to_nhwc_f32(const float* in, float* out, int hw, int c, cudaStream_t s);

// ---- whole-image kernels -------------------------------------------------------------------------
int unpack_grad(const float* packed, int H, int W, int roll_y, int roll_x, int nty, int ntx,
                int th, int tw, int thmax, int twmax, int world, int tiles_per_rank, float* grad,
                cudaStream_t s);
int regularizers(const float* img, int H, int W, float m0, float m1, float m2, float tv_w,
                 float tv_beta, float p_w, float p_pow, const float* aux, float aux_w, int roll_y,
                 int roll_x, double* loss_accum, float* grad, ReduceScratch rs, cudaStream_t s);
int adam_step(float* params, const float* grad, float* g1, float* g2, float* p1, float* avg_out,
              size_t n, float step_size, float b1, float b2, float bp1, float g1_corr,
              float g2_corr, float p1_corr, cudaStream_t s);
int dot_to(const float* x, const float* y, size_t n, double* out, ReduceScratch rs,
           cudaStream_t s);
int asum_to(const float* x, size_t n, double* out, ReduceScratch rs, cudaStream_t s);
int axpby(float a, const float* x, float b, float* y, size_t n, cudaStream_t s);
// y += sign * (num[0] / den_host [- sub[0]]) * x : the coefficient forms of the L-BFGS two-loop
// recursion; optionally stores the coefficient's first factor into *store.
int axpy_dev(const float* x, float* y, size_t n, const double* num, double den, const double* sub,
             double sign, double* store, cudaStream_t s);
int scale_dev(float* y, size_t n, double num, const double* den, cudaStream_t s);

}  // namespace st
