// tcgen05 / TMA implicit-GEMM 3x3 convolution for bf16 NHWC activations (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace st {

// bf16 weight matrices of one conv layer, K-major, for both directions:
//   fwd [cout][9*cin]  : B operand of  out[p][co] = sum_k A[p][k] * Wf[co][k],  k = tap*cin + ci
//   bwd [cin][9*cout]  : same GEMM with flipped taps and in/out channels exchanged
struct TcWeights {
  void* fwd = nullptr;             // 16-bit: bf16, or fp16 when fwd_half (ST_PREC_FP16)
  __nv_bfloat16* bwd = nullptr;    // always bf16: the backward pass runs on bf16 gradients
  __nv_bfloat16* bwd_rows = nullptr;   // first layer only: [48][3*cout], the three x taps of a kernel
                                       // row side by side (conv_pix_tc.cu)
  bool fwd_half = false;
  // split-operand mode (ST_PREC_TC32, conv_tc2.cu): fp16 [rows][9 * 3K] packs, K segments
  // [Whi | Whi | Wlo] per tap, weights pre-scaled by the power of two split_scale
  void* fwd32 = nullptr;
  void* bwd32 = nullptr;
  float split_scale = 1.f;
  void* map_fwd = nullptr;     // host copies of the CUtensorMap descriptors (128 B each)
  void* map_bwd = nullptr;
};

struct TcContext {
  bool enabled = false;
  bool pair_kernel = true;     // conv_tc2.cu (cta_group::2); ST_CONV_V1=1 selects the single-CTA kernel
  // debugging switches, read once from the environment in tc_init
  bool resident_weights = true;   // ST_TC_NO_RESB=1 disables
  bool defer_scale = true;        // ST_NO_DEFER=1 disables
  bool pool_fusion = true;        // ST_NO_POOL_FUSION=1 disables
  bool pix_rows_kernel = true;    // ST_NO_PIX_ROWS=1: first-layer backward through conv_tc2.cu instead
  bool fwd_bits = true;           // ST_NO_FWD_BITS=1: ReLU bit masks made from the activations by
                                  // relu_bits_from_act instead of the forward epilogues
  bool pdl = true;                // ST_NO_PDL=1: no programmatic dependent launch of the conv kernels
  int force_bn = 0;               // ST_TC_BN=64|128|256
  int sm_count = 0;
  void* encode_fn = nullptr;   // cuTensorMapEncodeTiled, fetched through cudaGetDriverEntryPoint
};

int tc_init(TcContext& tc, int sm_count);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (kernel, device): set it once for each.
cudaError_t tc_allow_smem_impl(const void* kernel, int bytes);
template <typename K>
inline cudaError_t tc_allow_smem(K kernel, int bytes) {
  return tc_allow_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}
void tc_destroy(TcContext& tc);
int tc_pack_weights(TcContext& tc, TcWeights& w, const float* w_oihw_host, int cin, int cout,
                    bool fwd_half);
// First layer (cin = 3): only the backward pack, [16][9*cout] with rows 3..15 zero.
int tc_pack_first(TcContext& tc, TcWeights& w, const float* w_oihw_host, int cout);
void tc_free_weights(TcWeights& w);

bool tc_shape_ok(const TcContext& tc, const TcWeights& w, int cin, int cout);

template <typename T>
inline bool tc_usable(const TcContext& tc, const TcWeights& w, int cin, int cout) {
  if constexpr (sizeof(T) == 2) return tc_shape_ok(tc, w, cin, cout);
  return false;
}

// out = epilogue(conv3x3(in)) with in [nb][h][w][cin], out [nb][h][w][cout] (16-bit NHWC).
//   forward : out = max(acc + bias, 0); operands/outputs are fp16 when w.fwd_half, else bf16
//   backward: out = (mask > 0 ? acc : 0) + inj_scale[tile] * inj   (each may be null); gradients
//             and backward weights are bf16
// The ReLU mask travels as ONE BIT per element, `relu_bits` [nb][h][w][cout/32] words (relu_bit()
// gives the position of a channel inside its word): written by the forward kernel that produced the
// blob (may be null: not wanted), read by the backward kernel of the layer that consumes it (may be
// null: no ReLU).  The single-CTA kernel (ST_CONV_V1) still reads the activation `mask_act`.
int conv3x3_tc(TcContext& tc, const TcWeights& w, const void* in, void* out, int nb, int h, int wd,
               int cin, int cout, bool forward, const float* bias, const void* mask_act,
               uint32_t* relu_bits, const __nv_bfloat16* inj, const float* inj_scale,
               cudaStream_t s);
// CTA-pair (cta_group::2) kernel of conv_tc2.cu; activations are [nb][h][w][c] (a batch of tiles).
int conv3x3_tc_pair(TcContext& tc, const TcWeights& w, const void* in, void* out, int nb, int h,
                    int wd, int cin, int cout, bool forward, const float* bias, uint32_t* relu_bits,
                    const __nv_bfloat16* inj, const float* inj_scale, cudaStream_t s);
// channel c of a 32-channel chunk sits at this bit of the chunk's mask word
inline int relu_bit(int c) { return ((c & 31) >> 1) + 16 * (c & 1); }
// Split-operand mode (ST_PREC_TC32): fp32 NHWC in / out, fp16 hi + lo operands, three MMAs per product
// (fp32-class result on the tensor cores).  See conv_tc2.cu.
int tc_pack_split(TcContext& tc, TcWeights& w, const float* w_oihw_host, int cin, int cout);
int split_f32(const float* in, void* out_hi_lo, size_t pixels, int c, float scale, cudaStream_t s);
int conv3x3_tc32(TcContext& tc, const TcWeights& w, const float* in, float* out, int nb, int h,
                 int wd, int cin, int cout, bool forward, const float* bias, const float* mask_act,
                 const float* inj, float in_scale, void* split_buf, cudaStream_t s);
// Style GEMM of the split-operand mode: S_b = F_b * D_b in fp32 from the [hi | lo] planes of F and the
// fp32 delta-Gram (split on the fly with a per-tile power-of-two scale), sum |S_b| as partial sums.
int gemm_abs_tc32(TcContext& tc, const void* f_split, const float* delta, const unsigned* max_bits,
                  void* d_split, float* inv_sigma, float* s_out, int nb, int h, int w, int c,
                  double* abs_partials, int* per_tile, cudaStream_t s);
// Forward convolution + the 2x2/2 pooling layer behind it in one kernel.  pool_out [nb][ho][wo][cout]
// receives the pooled map, pool_mask (bytes, same shape) what pool_bwd_mask needs:
//   max: bits 0-1 = window position of the first maximum, bit 2 = maximum > 0
//   ave: bit d    = window input d > 0
// `out` is written only when write_full.
int conv3x3_pool_tc_pair(TcContext& tc, const TcWeights& w, const void* in, void* out,
                         void* pool_out, uint8_t* pool_mask, int nb, int h, int wd, int cin, int cout,
                         const float* bias, bool is_max, bool write_full, uint32_t* relu_bits,
                         cudaStream_t s);
// Backward of the first (3-channel) convolution on tensor cores: dz [nb][h][w][cz] bf16 -> planar f32
// gradient; tile b goes to grad + b * batch_stride.  Needs weights packed by tc_pack_first.
int conv_last_bwd_tc_pair(TcContext& tc, const TcWeights& w, const __nv_bfloat16* dz, int nb, int h,
                          int wd, int cz, float* grad, long batch_stride, long plane_stride,
                          long row_stride, cudaStream_t s);
// The same on the dedicated kernel of conv_pix_tc.cu (one MMA per kernel ROW, the x shift in the
// epilogue); needs tc_pack_first_rows.
int tc_pack_first_rows(TcContext& tc, TcWeights& w, const float* w_oihw_host, int cout);
bool conv_pix_bwd_tc_ok(const TcContext& tc, const TcWeights& w, int cz);
int conv_pix_bwd_tc(TcContext& tc, const TcWeights& w, const __nv_bfloat16* dz, int nb, int h, int wd,
                    int cz, float* grad, long batch_stride, long plane_stride, long row_stride,
                    cudaStream_t s);
// First convolution (3 -> 64 channels) on tensor cores from the planar f32 image (conv_first_tc.cu).
struct ImageBatch;
int tc_pack_first_fwd(TcContext& tc, TcWeights& w, const float* w_oihw_host, int cout, bool half,
                      const float* bias_host);
int conv_first_fwd_tc(TcContext& tc, const TcWeights& w, const ImageBatch& img, int h, int wd,
                      const float* bias, void* out, uint32_t* relu_bits, cudaStream_t s);
// Per batch tile b: S_b[p][n] = sum_c F_b[p][c] * D_b[n][c] for F [nb][h][w][c] and D [nb][c][c],
// both bf16 or both fp16 (half_in); S is written as bf16.  sum |S_b| is left as partial sums
// abs_partials[b * per_tile + i], i < *per_tile, to be added in index order; the buffer must hold
// gemm_abs_partials_needed() doubles.
int gemm_abs_tc_pair(TcContext& tc, const void* f, const void* d, bool half_in, __nv_bfloat16* s_out,
                     int nb, int h, int w, int c, double* abs_partials, int* per_tile,
                     cudaStream_t s);
size_t gemm_abs_partials_needed(int nb, int h, int w, int c);

// ST_PREC_TC32: Gram matrices of fp32 features on the tensor cores, from the [hi | lo] fp16 planes of
// split_f32 (gram_tc.cu).  gram [nb][C][C] full symmetric fp32.
bool gram_tc32_ok(const TcContext& tc, int c);
size_t gram_tc32_part_floats(int nb, int hw, int c);
int gram_tc32(TcContext& tc, const void* f_split, int nb, int hw, int c, float* part, float* gram,
              cudaStream_t s);
// G_b = F_b^T F_b / (C*hw) for 16-bit NHWC F [nb][hw][c] on tcgen05 (gram_tc.cu), finished straight
// into the style term.  part: split-K scratch of gram_tc_part_floats() floats.
bool gram_tc_ok(const TcContext& tc, int c);
size_t gram_tc_part_floats(const TcContext& tc, int nb, int hw, int c);
// delta[b] = G_b - target (full [C][C]),
// max_bits[b] (optional) = max |delta_b| as float bits, and loss_part[b * *parts_per_tile + i] =
// partial sums of delta_ij^2 over j <= i, to be added in index order (delta_pack does).  G itself is
// not stored.  loss_part must hold nb * (c / 32)^2 doubles.
int gram_tc_delta(TcContext& tc, const void* f, bool half, int nb, int hw, int c, float* part,
                  const float* target, float* delta, unsigned* max_bits, double* loss_part,
                  int* parts_per_tile, cudaStream_t s);



}  // namespace st
