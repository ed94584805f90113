// Host-side launchers of the CUDA kernels (implemented in kernels_simt.cu / conv_tc.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

// Tiles of equal shape are evaluated as one batch: every activation is [nb][h][w][C].
constexpr int kMaxBatch = 16;

// Where the first convolution reads its pixels: an un-rolled image [3][H][W] addressed circularly
// (the virtual roll) so that no tile copy is ever materialised.
struct ImageBatch {
  const float* base;   // [3][H][W]
  int H, W;            // full image size
  int nb;              // tiles in the batch
  int oy[kMaxBatch], ox[kMaxBatch];   // canonical coordinate of tile pixel (0,0), before wrapping
};

// Per-tile origin inside the whole-image content feature map (already including the feature roll).
struct TargetOffsets {
  int ty0[kMaxBatch], tx0[kMaxBatch];
};

// Scratch for deterministic reductions, shared by all kernels of one context (single stream).
struct ReduceScratch {
  double* partials;    // kMaxReduceBlocks * 4 doubles
  unsigned* counter;   // kMaxBatch counters (one per tile of the batch), zero between kernels
};
constexpr int kMaxReduceBlocks = 1 << 17;

enum BwdEpilogue { kEpiNone = 0, kEpiMask = 1, kEpiInj = 2 };

// ---- convolutions (NHWC activations of type T) -----------------------------------------------
template <typename T>
int conv_first_fwd(const ImageBatch& img, int h, int w, const float* wpack /*[27][cout]*/,
                   const float* bias, T* out, int cout, cudaStream_t s);
// grad of tile b goes to grad + b * batch_stride as [3][h][w] with the given plane / row strides
template <typename T>
int conv_last_bwd(const T* dz, int nb, int h, int w, int cin_of_dz, const float* wpack /*[27][cout]*/,
                  float* grad, long batch_stride, long plane_stride, long row_stride,
                  cudaStream_t s);
// out[p][co] = epilogue(sum_{tap,ci} in[p+tap][ci] * wpack[tap][ci][co])
template <typename T>
int conv3x3_simt(const T* in, const float* wpack, const float* bias, T* out, int nb, int h, int w,
                 int cin, int cout, bool forward, const T* mask_act, const T* inj, cudaStream_t s);

// ---- pooling -----------------------------------------------------------------------------------
template <typename T>
int pool_fwd(const T* in, T* out, int nb, int h, int w, int c, bool is_max, cudaStream_t s);
// d_in = [mask](in>0) * pool_bwd(d_out) + [inj_scale[tile]] * [inj]
template <typename TA, typename T>
int pool_bwd(const T* d_out, const TA* in, T* d_in, int nb, int h, int w, int c, bool is_max,
             bool apply_mask, const T* inj, const float* inj_scale, cudaStream_t s);

// backward of a pooling layer from the one-byte mask the fused conv+pool kernel stored
// (conv_tc.h: conv3x3_pool_tc_pair); the ReLU mask of the pooled layer is part of the byte
template <typename T>
int pool_bwd_mask(const T* d_out, const uint8_t* mask, T* d_in, int nb, int h, int w, int c,
                  bool is_max, const T* inj, const float* inj_scale, cudaStream_t s);

// ---- Gram / style ------------------------------------------------------------------------------
// gram_full[C][C] (symmetric, float) = F^T F / (C*HW).  F is NHWC [hw][c] (channel_major=false) or
// [c][hw] (channel_major=true).  part: scratch for split-K partials (part_floats floats).
template <typename T>
int gram_full(const T* f, int hw, int c, bool channel_major, float* gram, float* part,
              size_t part_floats, int sm_count, cudaStream_t s);
// delta[b] = gram[b] - target (symmetric, full [C][C] each);
// tile_loss[b * loss_stride] += w * 0.5 * sum_{j<=i} delta_ij^2.  delta_16 (optional) receives the
// 16-bit copy for the tensor-core style GEMM: bf16, or fp16 scaled per tile by a power of two
// (half) -- eps_eff[b] is the EPS that compensates that scaling in normalize().
// (track_max: max_bits[b] = max |delta_b| even without a 16-bit copy)
int gram_delta(const float* gram, const float* target, float* delta, void* delta_16, bool half,
               unsigned* max_bits, float* eps_eff, int c, int nb, double w, double* tile_loss,
               int loss_stride, ReduceScratch rs, cudaStream_t s, bool track_max = false);
// The 16-bit copy of delta alone (second half of gram_delta): bf16, or fp16 scaled per tile by the
// power of two derived from max_bits[b]; eps_eff[b] as in gram_delta.
// With loss_part (optional): tile_loss[b * loss_stride] += w * 0.5 * sum_i loss_part[b * n_part + i],
// the partials added in a fixed order by one block per tile.
int delta_pack(const float* delta, void* delta_16, bool half, unsigned* max_bits, float* eps_eff,
               int c, int nb, const double* loss_part, int n_part, double w, double* tile_loss,
               int loss_stride, cudaStream_t s);
// out[b * out_stride] = sum(partials[b*n .. b*n+n)) in a launch-independent order; if scale is not
// null also scale[b] = w / (sum / count + EPS)
int sum_partials(const double* partials, int n, int nb, double* out, int out_stride, float* scale,
                 float w, double count, const float* eps_eff, cudaStream_t s);
// *loss_accum += sum_b tile_loss[b * stride] (tile order); clears the slots
int loss_finalize(double* tile_loss, int stride, int nb, double* loss_accum, cudaStream_t s);
// S[p][co] = sum_ci F[p][ci] * delta[ci][co]; *sum_abs = sum |S|
template <typename T>
int style_grad(const T* f, const float* delta, T* s_out, int hw, int c, double* sum_abs,
               ReduceScratch rs, cudaStream_t s);
// per tile b (n elements each): inj = (accumulate ? inj : 0) + w / (sum_abs[b*stride] / n + EPS) * src
template <typename T>
int inject_scaled(T* inj, const T* src, size_t n, int nb, float w, const double* sum_abs,
                  int stat_stride, const float* eps_eff, bool accumulate, cudaStream_t s);
int symmetrize_lower(const float* lower, float* full, int c, cudaStream_t s);
int extract_lower(const float* full, float* lower, int c, cudaStream_t s);

// ---- content / deep-dream ------------------------------------------------------------------------
// Target map tgt is NHWC float [Hf][Wf][C] of the whole image, read for tile b at
// ((ty0[b]+y) mod Hf, (tx0[b]+x) mod Wf).  stats[b*stride + 0] = sum c^2, [.. + 1] = sum |c| with
// c = F - target (target==nullptr: c = F).
template <typename T>
int diff_stats(const T* f, int nb, int hf, int wf, int c, const float* tgt, int Hf, int Wf,
               const TargetOffsets& offs, double* stats, int stat_stride, ReduceScratch rs,
               cudaStream_t s);
// inj = (accumulate ? inj : 0) + w / (stats[1]/n + EPS) * (F - target);
// tile_loss[b * loss_stride] += loss_w * 0.5 * stats[0]
template <typename TA, typename T>
int diff_inject(const TA* f, int nb, int hf, int wf, int c, const float* tgt, int Hf, int Wf,
                const TargetOffsets& offs, const double* stats, int stat_stride, float w,
                double loss_w, double* tile_loss, int loss_stride, T* inj, bool accumulate,
                cudaStream_t s);

// bits[p][c/32] = ReLU mask of act [p][c] (one bit per element, position relu_bit() of conv_tc.h):
// the fallback for blobs whose producing kernel did not write the mask itself.
template <typename T>
int relu_bits_from_act(const T* act, uint32_t* bits, size_t pixels, int c, cudaStream_t s);

// ---- layout conversion ---------------------------------------------------------------------------
template <typename T>
int nhwc_to_nchw_f32(const T* in, float* out, int hw, int c, cudaStream_t s);
int nchw_to_nhwc_f32(const float* in, float* out, int hw, int c, cudaStream_t s);

// ---- whole-image kernels -------------------------------------------------------------------------
// floats of one rank's chunk of the exchange buffer: the tiles, padded to a multiple of four, plus a
// four-float tail that carries the rank's loss as a double
size_t packed_rank_stride(int tiles_per_rank, int thmax, int twmax);
// loss_accum (optional) += the losses in the tails of the ranks' chunks, in rank order
int unpack_grad(const float* packed, int H, int W, int roll_y, int roll_x, int nty, int ntx,
                int th, int tw, int thmax, int twmax, int world, int tiles_per_rank, float* grad,
                double* loss_accum, cudaStream_t s);
int regularizers(const float* img, int H, int W, float m0, float m1, float m2, float tv_w,
                 float tv_beta, float p_w, float p_pow, const float* aux, float aux_w, int roll_y,
                 int roll_x, double* loss_accum, float* grad, ReduceScratch rs, cudaStream_t s);
// st_unpack_grad + st_regularizers in one pass over the image
int unpack_regularizers(const float* packed, int H, int W, int nty, int ntx, int th, int tw,
                        int thmax, int twmax, int world, int tiles_per_rank, const float* img,
                        float m0, float m1, float m2, float tv_w, float tv_beta, float p_w,
                        float p_pow, const float* aux, float aux_w, int roll_y, int roll_x,
                        double* loss_accum, float* grad, ReduceScratch rs, cudaStream_t s);
int adam_step(float* params, const float* grad, float* g1, float* g2, float* p1, float* avg_out,
              size_t n, float step_size, float b1, float b2, float bp1, float g1_corr,
              float g2_corr, float p1_corr, cudaStream_t s);
// One resampling pass along x (transposed = false: out[c][y][xx]) or along y (transposed = true:
// out[c][yy][x]) with per-output (first source index, count) bounds and ksize float64 weights.
int resample_pass(const float* in, float* out, int channels, int in_h, int in_w, int out_size,
                  bool along_y, const int* bounds_dev, const double* kk_dev, int ksize, cudaStream_t s);
// stats[0] = sum |avg - old|, stats[1] = sum of squared periodic forward differences; old := avg
int iter_stats(const float* avg, float* old, int H, int W, double* stats, ReduceScratch rs,
               cudaStream_t s);
// iter_stats + get_image_u8 (pic may be null) in one pass over the averaged iterate
int output_step(const float* avg, float* old, int H, int W, float m0, float m1, float m2, bool bgr,
                double* stats, uint8_t* pic, ReduceScratch rs, cudaStream_t s);
// out[y][x][k] = uint8(clip(params[c][y][x] + mean[c], 0, 255)), c = bgr ? 2 - k : k
int get_image_u8(const float* params, int H, int W, float m0, float m1, float m2, bool bgr,
                 uint8_t* out, cudaStream_t s);
int dot_to(const float* x, const float* y, size_t n, double* out, ReduceScratch rs,
           cudaStream_t s);
int asum_to(const float* x, size_t n, double* out, ReduceScratch rs, cudaStream_t s);
int axpby(float a, const float* x, float b, float* y, size_t n, cudaStream_t s);
// y += sign * (num[0] / den_host [- sub[0]]) * x : the coefficient forms of the L-BFGS two-loop
// recursion; optionally stores the coefficient's first factor into *store.
int axpy_dev(const float* x, float* y, size_t n, const double* num, double den, const double* sub,
             double sign, double* store, cudaStream_t s);
int scale_dev(float* y, size_t n, double num, const double* den, cudaStream_t s);
// L-BFGS with device-resident memory (see kernels_image.cu): the step s = -scale * H grad is written
// to ring_s[head] and added to params; the candidate pair is completed and kept / dropped on the device
int lbfgs_direction(const float* grad, size_t n, int n_corr, float* ring_s, const float* ring_y,
                    double* state, float* p_scratch, float* params, float initial_step,
                    ReduceScratch rs, cudaStream_t s);
int lbfgs_commit(const float* grad_new, const float* grad_old, size_t n, int n_corr,
                 const float* ring_s, float* ring_y, double* state, ReduceScratch rs, cudaStream_t s);

}  // namespace st
