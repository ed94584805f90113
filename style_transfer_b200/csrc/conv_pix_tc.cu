// Backward of the first convolution to the pixels (conv1_1: cz channels -> 3 image planes) on tcgen05.
//
//   grad[b][c][y][x] = sum_{dy, dx, k} dz[b][y + dy - 1][x + dx - 1][k] * Wb[c][(dy, dx)][k]
//
// The generic pair kernel (conv_tc2.cu, <16, 9, Pix>) spends one MMA per TAP and K = 16 step: 36
// MMAs per 256 pixels, each re-reading 4 KB of the A window from shared memory for 16 accumulator
// columns of which 3 are used -- it is bound by the shared-memory bandwidth at 2.2x the HBM time of
// its 64-channel input (profiles/r01_conv_step_ncu_final.md).  Here the three x taps of a kernel row
// share ONE MMA: the B operand holds their weights side by side (N = 48: column dx*16 + c), the A
// operand is the un-shifted window row, and the accumulator row of window pixel q holds the three
// partial sums T_dx[q][c] that pixel contributes to its left / own / right output pixel.  12 MMAs per
// tile instead of 36; the x shift moves into the epilogue:
//
//   grad[(y, x)] = T_0[(y, x - 1)] + T_1[(y, x)] + T_2[(y, x + 1)]
//
// which needs T of the two halo columns, so the M = 128 rows of a tile are the 12 x 10 pixels of the
// output tile (12 rows x 8 columns) plus its left / right halo, row m = ty*10 + txw.  With the window
// stored at pitch 10 (one TMA box of 14 x 10 pixels, as in conv_tc2.cu) the A operand of kernel row dy
// is simply the 128 CONSECUTIVE 128-byte rows starting at window row dy*10: a dense K-major tile.
//
// Single-CTA kernel (cta_group::1), persistent, two CTAs per SM, 192 threads: warp 0 TMA producer, warp 1 MMA issuer +
// TMEM owner, warps 2..5 epilogue (TMEM -> shared-memory exchange of the 9 partials -> 3 planar f32
// stores).  Weights (48 x 3*cz bf16, 18 KB at cz = 64) stay resident in shared memory.
#include <cuda.h>

#include <vector>

#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"

namespace st {

namespace {

constexpr int kPThreads = 192;
constexpr int kPH = 12, kPW = 8;                 // output pixels per tile
constexpr int kWinH = kPH + 2, kWinW = kPW + 2;  // halo window
constexpr int kWinBytes = kWinH * kWinW * 128;   // 17920 bytes loaded per 64-channel block
constexpr int kStageBytes = 19 * 1024;           // >= (2*kWinW + 128) * 128: the dy = 2 operand over-reads
constexpr int kPStages = 3;                    // 85 KB per CTA at cz = 64: two CTAs share an SM
constexpr int kBTapBytes = 48 * 128;             // one (channel block, kernel row) weight tile
constexpr int kExFloats = 128 * 9;               // exchange buffer of one tile
constexpr uint32_t kSpinP = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > kSpinP) __trap();             // protocol bug: fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// columns taddr .. taddr + 2 of this warp's 32 lanes (the x4 shape is the narrowest 32x32b load)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// K-major SWIZZLE_128B descriptor of a dense tile: 128-byte rows, 8-row groups 1024 bytes apart.  The
// start address may be any multiple of 128 bytes (profiles/r01_desc_offset.md).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// c_format f32 | A, B bf16 | N = 48 | M = 128
constexpr uint32_t kPixIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(48 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);

struct PixArgs {
  int nb, h, w, cz;
  int tiles_x, tiles_y, num_tiles;
  FastDiv div_x, div_y;
  float* pix;                       // planar f32 gradient
  long pix_batch, pix_plane, pix_row;
};

struct PixTile {
  int b, x0, y0;
};
__device__ __forceinline__ PixTile decode(const PixArgs& a, int tile) {
  PixTile t;
  const int r = (int)a.div_x.div((unsigned)tile);
  t.x0 = (tile - r * a.tiles_x) * kPW;
  t.b = (int)a.div_y.div((unsigned)r);
  t.y0 = (r - t.b * a.tiles_y) * kPH;
  return t;
}

__global__ void __launch_bounds__(kPThreads, 2)
conv_pix_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_in,
                       const __grid_constant__ CUtensorMap map_w, const PixArgs a) {
  ST_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int kcb = a.cz >> 6;                               // 64-channel blocks
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + kPStages * kStageBytes;       // [kcb][3][48 rows][128 B]
  float* ex = reinterpret_cast<float*>(b_base + kcb * 3 * kBTapBytes);     // [2][128][9]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ex + 2 * kExFloats);
  uint64_t* full = bars;
  uint64_t* empty = full + kPStages;
  uint64_t* b_full = empty + kPStages;
  uint64_t* t_full = b_full + 1;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_in), prefetch_tmap(&map_w);
    for (int i = 0; i < kPStages; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&t_full[i], 1), mbar_init(&t_empty[i], 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);               // two accumulators of 64 columns (48 used)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    if (elect_one()) {                                      // resident weights, one barrier
      mbar_expect_tx(b_full, (uint32_t)(kcb * 3 * kBTapBytes));
      for (int cb = 0; cb < kcb; ++cb)
        for (int dy = 0; dy < 3; ++dy)
          tma_load_2d(&map_w, b_full, b_base + (cb * 3 + dy) * kBTapBytes, dy * a.cz + cb * 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const PixTile t = decode(a, tile);
      for (int cb = 0; cb < kcb; ++cb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], kWinBytes);
          tma_load_4d(&map_in, &full[stage], a_base + stage * kStageBytes, cb * 64, t.x0 - 1,
                      t.y0 - 1, t.b);
        }
        __syncwarp();
        if (++stage == kPStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    int stage = 0;
    uint32_t phase = 0, it = 0;
    mbar_wait(b_full, 0);
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, use = it >> 1;
      mbar_wait(&t_empty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 64;
      for (int cb = 0; cb < kcb; ++cb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t da0 = make_desc(smem_u32(a_base + stage * kStageBytes));
        const uint64_t db0 = make_desc(smem_u32(b_base + cb * 3 * kBTapBytes));
        if (elect_one()) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma(d_tmem, da0 + (uint64_t)(dy * kWinW * 8 + k * 2),
                     db0 + (uint64_t)(dy * (kBTapBytes >> 4) + k * 2), kPixIdesc,
                     (dy | k) != 0 ? 1u : (uint32_t)(cb != 0));
          tc_commit(&empty[stage]);
          if (cb == kcb - 1) tc_commit(&t_full[buf]);
        }
        __syncwarp();
        if (++stage == kPStages) stage = 0, phase ^= 1;
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    const int q = warp & 3;                                // TMEM lane quadrant of this warp
    const int m = q * 32 + lane;                           // accumulator row = window pixel
    const int ty = (m * 205) >> 11, txw = m - ty * kWinW;  // m / 10 for m < 128
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, use = it >> 1;
      const PixTile t = decode(a, tile);
      mbar_wait(&t_full[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 64 + ((uint32_t)(q * 32) << 16);
      uint32_t r0[4], r1[4], r2[4];                        // T_dx[m][c], c = 0..2
      tmem_ld4(taddr, r0), tmem_ld4(taddr + 16, r1), tmem_ld4(taddr + 32, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
      float* e = ex + buf * kExFloats + m * 9;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        e[c] = __uint_as_float(r0[c]), e[3 + c] = __uint_as_float(r1[c]), e[6 + c] = __uint_as_float(r2[c]);
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps
      const int y = t.y0 + ty, x = t.x0 + txw - 1;
      if (m < kPH * kWinW && txw >= 1 && txw <= kPW && y < a.h && x < a.w) {
        const float* el = ex + buf * kExFloats + (m - 1) * 9;     // T of the left / own / right pixel
        float* dst = a.pix + (size_t)t.b * a.pix_batch + (size_t)y * a.pix_row + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[(size_t)c * a.pix_plane] = (el[c] + el[9 + 3 + c]) + el[18 + 6 + c];
      }
      // ex[buf] is rewritten two tiles later: every warp has passed the next tile's barrier by then
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode(const TcContext& tc, CUtensorMap* map, int rank, const void* base, const cuuint64_t* dims,
           const cuuint64_t* strides, const cuuint32_t* box) {
  cuuint32_t estride[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(tc.encode_fn)(
      map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estride,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (pixel backward) failed with CUresult " + std::to_string((int)r));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

}  // namespace

// B3[dx*16 + c][dy*cout + co] = w[co][c][2 - dy][2 - dx]: the three x taps of kernel row dy side by
// side in N, flipped like every backward-data weight.  Rows with c >= 3 stay zero.
int tc_pack_first_rows(TcContext& tc, TcWeights& w, const float* w_host, int cout) {
  if (!tc.enabled || !tc.pair_kernel || cout % 64 != 0) return ST_OK;
  std::vector<__nv_bfloat16> host((size_t)48 * 3 * cout, __float2bfloat16_rn(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < 3; ++ci)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
          host[(size_t)((2 - kx) * 16 + ci) * 3 * cout + (size_t)(2 - ky) * cout + co] =
              __float2bfloat16_rn(w_host[(((size_t)co * 3 + ci) * 3 + ky) * 3 + kx]);
  if (!w.bwd_rows) ST_CUDA(cudaMalloc((void**)&w.bwd_rows, host.size() * sizeof(__nv_bfloat16)));
  ST_CUDA(cudaMemcpy(w.bwd_rows, host.data(), host.size() * sizeof(__nv_bfloat16),
                     cudaMemcpyHostToDevice));
  return ST_OK;
}

bool conv_pix_bwd_tc_ok(const TcContext& tc, const TcWeights& w, int cz) {
  // the resident weights (18 KB per 64 channels) must fit beside the A stages
  return tc.enabled && tc.pair_kernel && tc.pix_rows_kernel && w.bwd_rows != nullptr && cz % 64 == 0 &&
         cz <= 256;
}

int conv_pix_bwd_tc(TcContext& tc, const TcWeights& w, const __nv_bfloat16* dz, int nb, int h, int wd,
                    int cz, float* grad, long batch_stride, long plane_stride, long row_stride,
                    cudaStream_t s) {
  PixArgs a{};
  a.nb = nb, a.h = h, a.w = wd, a.cz = cz;
  a.tiles_x = cdiv(wd, kPW), a.tiles_y = cdiv(h, kPH), a.num_tiles = a.tiles_x * a.tiles_y * nb;
  a.div_x = FastDiv(a.tiles_x), a.div_y = FastDiv(a.tiles_y);
  a.pix = grad, a.pix_batch = batch_stride, a.pix_plane = plane_stride, a.pix_row = row_stride;
  CUtensorMap map_in, map_w;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)cz, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)nb};
    const cuuint64_t strides[3] = {(cuuint64_t)cz * 2, (cuuint64_t)wd * cz * 2, (cuuint64_t)h * wd * cz * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kWinW, (cuuint32_t)kWinH, 1};
    int rc = encode(tc, &map_in, 4, dz, dims, strides, box);
    if (rc != ST_OK) return rc;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)3 * cz, 48};
    const cuuint64_t strides[1] = {(cuuint64_t)3 * cz * 2};
    const cuuint32_t box[2] = {64, 48};
    int rc = encode(tc, &map_w, 2, w.bwd_rows, dims, strides, box);
    if (rc != ST_OK) return rc;
  }
  const int smem_bytes = 1024 + kPStages * kStageBytes + (cz / 64) * 3 * kBTapBytes +
                         2 * kExFloats * (int)sizeof(float) + 256;
  auto kern = conv_pix_bwd_tc_kernel;
  ST_CUDA(tc_allow_smem(kern, 227 * 1024));
  // two co-resident CTAs per SM when the shared memory allows: per tile the epilogue is a serial
  // chain (accumulator wait -> TMEM load -> exchange -> barrier -> store) that a second CTA overlaps
  const int per_sm = smem_bytes <= 110 * 1024 ? 2 : 1;
  const int max_ctas = tc.sm_count * per_sm;
  const int grid = a.num_tiles < max_ctas ? a.num_tiles : max_ctas;
  TimerScope ts(s, kTimeConvSimt, 2.0 * 9 * cz * 3 * h * wd * nb);
  ST_LAUNCH(kern, grid, kPThreads, smem_bytes, s, map_in, map_w, a);
  return ST_OK;
}

}  // namespace st
