// Gram matrix G = F^T F / (C * HW) on tcgen05 tensor cores (num_utils.py:143-147, the "C^T C" of the
// north star) for bf16 NHWC features F [HW][C].
//
// GEMM view: D[i][j] = sum_p F[p][i] * F[p][j]: M = N = channels, K = pixels.  F is stored with the
// channels contiguous, i.e. both operands are "MN-major": a TMA box of 64 pixels x 64 channels
// (128-byte rows, SWIZZLE_128B) is exactly the canonical MN-major SW128 shared-memory layout
//   ((8,n),(8,k)) : ((1,LBO),(8,SBO))   [units of 16 bytes]
// with LBO = 8 KB (next 64-channel group = next box) and SBO = 1 KB (next 8 pixels).  One k-block
// (64 pixels of all C channels) is loaded ONCE and serves as A (the 128 rows of this CTA's M block)
// and as B (all C columns).
//
// The kernel is HBM-bound for C <= 256 (2*C flop per bf16 byte): the pixels are split over CTAs
// (split-K), every CTA streams its share once, accumulates a [128][C] fp32 block in TMEM and writes
// it to a partial buffer; gram_tc_finish sums the partials in split order (deterministic, no float
// atomics), scales, mirrors the lower triangle and subtracts the style target in the same pass.
//
// Warps (192 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = TMEM -> global.
#include <cuda.h>

#include <cstdlib>

#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"

namespace st {

namespace {

constexpr int kGThreads = 192;
constexpr int kBoxBytes = 64 * 128;       // 64 pixels x 64 channels bf16
constexpr uint32_t kSpinG = 1u << 24;

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
    if (spin > kSpinG) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// MN-major SWIZZLE_128B descriptor: 64-element MN groups kBoxBytes apart (LBO), 8-row K groups
// 1024 bytes apart (SBO).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kBoxBytes >> 4) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int C>
struct GramCfg {
  static constexpr int kGroups = C / 64;                              // boxes per k-block
  static constexpr int kStageBytes = (kGroups < 2 ? 2 : kGroups) * kBoxBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  // C <= 128 (16 KB stages): six stages = 97 KB, so TWO CTAs share an SM and one streams while the
  // other runs its prologue / epilogue (these layers are HBM-bound streams of ~1 MB per CTA)
  static constexpr int kStages = C <= 128 ? 6 : (kStagesRaw > 8 ? 8 : kStagesRaw);
  static constexpr int kMBlocks = C < 128 ? 1 : C / 128;
  static constexpr int kN = C > 256 ? 256 : C;                        // columns per MMA
  static constexpr int kNHalves = C / kN;
  static constexpr int kTmemCols = C < 32 ? 32 : C;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  // c_format f32 | A and B MN-major | N | M = 128; operand formats (bits 7, 10) are set at run time
  static constexpr uint32_t kIdescBase = (1u << 4) | (1u << 15) | (1u << 16) |
                                         ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

struct GramArgs {
  int hw;              // pixels
  int kb_total;        // ceil(hw / 64)
  int kb_per_split;
  int nsplit;
  int half;            // features are fp16 instead of bf16
  float* part;         // [nb][nsplit][C][C] fp32 partial sums
};

template <int C>
__global__ void __launch_bounds__(kGThreads, C <= 128 ? 2 : 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_f, const GramArgs a) {
  // PDL: barrier init / TMEM allocation overlap the previous kernel's tail; features are read below
  if (gridDim.x * gridDim.y <= (C <= 128 ? 296u : 148u))       // only when the whole grid is resident
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  using Cfg = GramCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x % Cfg::kMBlocks, split = blockIdx.x / Cfg::kMBlocks;
  const int b = blockIdx.y;                              // tile of the batch
  const int kb_begin = split * a.kb_per_split;
  const int kb_end = min(a.kb_total, kb_begin + a.kb_per_split);

  if constexpr (C == 64) {
    // rows 64..127 of the 128-row A operand read the second half of every stage: keep it zero
    for (int i = threadIdx.x; i < Cfg::kStages * (kBoxBytes / 16); i += kGThreads) {
      const int st = i / (kBoxBytes / 16), o = i % (kBoxBytes / 16);
      *reinterpret_cast<uint4*>(smem + st * Cfg::kStageBytes + kBoxBytes + o * 16) =
          make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_f);
    for (int i = 0; i < Cfg::kStages; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // warp-wide loops with warp-uniform state; only the TMA / MMA issue is elected (see conv_tc2.cu)
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* dst = smem + stage * Cfg::kStageBytes;
      if (elect_one()) {
        mbar_expect_tx(&full[stage], Cfg::kGroups * kBoxBytes);
#pragma unroll
        for (int g = 0; g < Cfg::kGroups; ++g)
          tma_load_3d(&map_f, &full[stage], dst + g * kBoxBytes, g * 64, kb * 64, b);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t fmt = a.half ? 0u : 1u;
    const uint32_t idesc = Cfg::kIdescBase | (fmt << 7) | (fmt << 10);
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t base = smem_u32(smem + stage * Cfg::kStageBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {               // 16 pixels of K per MMA: 2 KB down the box
          const uint64_t da = make_desc_mn(base + (2 * mb) * kBoxBytes + k * 2048);
#pragma unroll
          for (int nh = 0; nh < Cfg::kNHalves; ++nh) {
            const uint64_t db = make_desc_mn(base + nh * 4 * kBoxBytes + k * 2048);
            tc_mma(tmem_base + nh * 256, da, db, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
          }
        }
        tc_commit(&empty[stage]);
        if (kb == kb_end - 1) tc_commit(tfull);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) stage = 0, phase ^= 1;
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // row of this CTA's M block
    const int row = mb * 128 + m;                      // channel i
    if (kb_end > kb_begin) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    float* dst = a.part + (((size_t)b * a.nsplit + split) * C + row) * C;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc < C / 32; ++cc) {
      uint32_t r[32];
      if (kb_end > kb_begin) {
        tmem_ld32(taddr + cc * 32, r);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (row < C) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(dst + cc * 32 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// Reduction of the split partials fused with what the style term needs from the Gram matrix
// (style_transfer.py:587,591): delta = G - G_style (both triangles), per block the partial sum of
// delta^2 over j <= i (added up in block order by delta_pack: deterministic, and no ticket / fence per
// block -- a grid_reduce in each of the 32768 blocks of a C = 512 batch doubled this kernel's time)
// and max |delta| (float bits, for the fp16 scaling of the style GEMM operand).  Replaces
// a separate finalize + gram_delta pair: one launch and no round trip of G through memory per style layer.
struct GramFinish {
  const float* target;     // [C][C] symmetric
  float* delta;            // [nb][C][C]
  unsigned* max_bits;      // [nb] or null
  double* loss_part;       // [nb][blocks per tile]
};

// One block per 32 x 32 tile (ti, tj) of the matrix with tj <= ti; 32 x (32 / R) threads, R rows each.
// R = 1 on the narrow layers (few tiles but up to 64 splits: with four rows per thread their
// reduction was a chain of 256 dependent-latency loads on 48 blocks), R = 4 on the wide ones (many
// tiles, few splits: 1024-thread blocks were twice slower there).  The tile is summed over the
// splits with coalesced reads, written, and its mirror image (tj, ti) written from a shared-memory
// transpose, so both triangles are stored with full 128-byte lines (writing out[j][i] straight from
// the thread that owns (i, j) cost a 32-way scattered store per warp).
template <int R>
__global__ void __launch_bounds__(1024 / R)
gram_tc_finish_kernel(const float* __restrict__ part, int nsplit, int c, double scale,
                      const GramFinish fin) {
  ST_PDL_ENTRY();
  constexpr int kRowsPerPass = 32 / R, kWarps = 32 / R;
  __shared__ float tile[32][33];
  __shared__ double sh[kWarps];
  const int tj = blockIdx.x, ti = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const unsigned nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
  if (tj > ti) {                                       // upper-triangle tiles are written by their mirror
    if (tx == 0 && ty == 0) fin.loss_part[(size_t)b * nblk + blk] = 0.0;
    return;
  }
  const size_t cc = (size_t)c * c;
  float* out = fin.delta + (size_t)b * cc;
  double v = 0.0;
  float mx = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int il = ty + kRowsPerPass * r, i = ti * 32 + il, j = tj * 32 + tx;
    const float* p = part + (size_t)b * nsplit * cc + (size_t)i * c + j;
    double sum = 0.0;
#pragma unroll 8
    for (int s = 0; s < nsplit; ++s) sum += (double)p[(size_t)s * cc];
    const float g = (float)(sum * scale);
    const float d = g - fin.target[(size_t)i * c + j];
    tile[il][tx] = d;
    if (j <= i) {                                      // lower triangle incl. the diagonal
      mx = fmaxf(mx, fabsf(d));
      v += (double)d * d;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int il = ty + kRowsPerPass * r;
    if (ti != tj) {
      out[(size_t)(ti * 32 + il) * c + tj * 32 + tx] = tile[il][tx];
      out[(size_t)(tj * 32 + il) * c + ti * 32 + tx] = tile[tx][il];      // mirror tile
    } else {
      // diagonal tile: the lower triangle is authoritative, the upper one its mirror
      out[(size_t)(ti * 32 + il) * c + tj * 32 + tx] = tx <= il ? tile[il][tx] : tile[tx][il];
    }
  }
  if (fin.max_bits != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (tx == 0 && mx > 0.f) atomicMax(fin.max_bits + b, __float_as_uint(mx));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (tx == 0) sh[ty] = v;
  __syncthreads();
  if (ty == 0) {                                       // fixed tree over the warp sums
    double x = tx < kWarps ? sh[tx] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (tx == 0) fin.loss_part[(size_t)b * nblk + blk] = x;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Number of pixel splits per tile.  It depends on the layer shape only -- never on the batch size or
// the SM count -- so that a tile's Gram matrix has the same bits whichever rank / batch computes
// it.  The cap keeps the partial buffer (written, then re-read by the finalize kernel) at 1-8 MB per
// tile while a batch of a few tiles still fills the machine.
template <int C>
int gram_splits(int hw) {
  const int kb_total = cdiv(hw, 64);
  // wide layers: few splits -- a split's partial block is [128][C] fp32, so at C >= 256 the partials
  // written and re-read by the finish kernel outweighed the 16-bit features themselves
  int nsplit = C <= 64 ? 64 : (C <= 128 ? 32 : (C <= 256 ? 8 : 4));
  // at least 8 k-blocks (512 pixels) per split: shorter streams are all pipeline fill and drain
  nsplit = nsplit > kb_total / 8 ? kb_total / 8 : nsplit;
  nsplit = nsplit < 1 ? 1 : nsplit;
  const int kb_per_split = cdiv(kb_total, nsplit);
  return cdiv(kb_total, kb_per_split);
}

template <int C>
int launch_gram(TcContext& tc, const void* f, bool half, int nb, int hw, float* part,
                const GramFinish& fin, cudaStream_t s) {
  using Cfg = GramCfg<C>;
  CUtensorMap map_f;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)hw, (cuuint64_t)nb};
    cuuint64_t gstride[2] = {(cuuint64_t)C * 2, (cuuint64_t)hw * C * 2};
    cuuint32_t box[3] = {64, 64, 1}, estride[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(tc.encode_fn)(
        &map_f, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
        const_cast<void*>(f), gdim, gstride,
        box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (gram) failed with CUresult " + std::to_string((int)r));
      return ST_ERR_CUDA;
    }
  }
  GramArgs a{};
  a.hw = hw, a.kb_total = cdiv(hw, 64), a.part = part, a.half = half ? 1 : 0;
  a.nsplit = gram_splits<C>(hw);
  a.kb_per_split = cdiv(a.kb_total, a.nsplit);
  auto kern = gram_tc_kernel<C>;
  ST_CUDA(tc_allow_smem(kern, Cfg::kSmemBytes));
  TimerScope ts(s, kTimeGram, 2.0 * C * C * hw * nb);
  ST_LAUNCH(kern, dim3(a.nsplit * Cfg::kMBlocks, nb), kGThreads, Cfg::kSmemBytes, s, map_f, a);
  constexpr int kR = C >= 256 ? 4 : 1;
  ST_LAUNCH(gram_tc_finish_kernel<kR>, dim3(C / 32, C / 32, nb), dim3(32, 32 / kR), 0, s, part,
            a.nsplit, C, 1.0 / ((double)C * hw), fin);
  return ST_OK;
}

// ---- Gram matrix of an fp32 feature map on the tensor cores (ST_PREC_TC32) -------------------------
// The features arrive as the fp16 [hi | lo] planes of split_f32 (2C channels per pixel); the Gram
// kernel above runs on them as a 2C-channel map and delivers per split the four quadrants
// [[hi^T hi, hi^T lo], [lo^T hi, lo^T lo]], whose sum is F^T F to ~2^-22.  The tensor core truncates
// when it accumulates (conv_tc2.cu), so a split covers only kChainKb k-blocks (32 MMA steps: a bias of
// ~1e-6); the many partial blocks are added in double by the fold kernel in split order.
constexpr int kChainKb = 8;

__global__ void __launch_bounds__(256)
gram_fold_kernel(const float* __restrict__ part, int nsplit, int c, double scale, float* __restrict__ gram) {
  ST_PDL_ENTRY();
  // blockIdx.x = row i of the C x C result, blockIdx.y = tile of the batch.  The 256 threads are
  // 256 / C groups of C columns; group g adds the splits s = g, g + groups, ... (coalesced rows of
  // the four quadrants), the groups are then added in index order: deterministic.
  __shared__ double sh[256];
  const int c2 = 2 * c, i = blockIdx.x;
  const int j = threadIdx.x % c, g = threadIdx.x / c, groups = 256 / c;
  const size_t pstride = (size_t)c2 * c2;
  const float* p = part + (size_t)blockIdx.y * nsplit * pstride;
  const size_t o00 = (size_t)i * c2 + j, o01 = o00 + c, o10 = (size_t)(c + i) * c2 + j, o11 = o10 + c;
  double sum = 0.0;
#pragma unroll 4
  for (int s = g; s < nsplit; s += groups) {
    const float* q = p + (size_t)s * pstride;
    sum += ((double)q[o00] + (double)q[o01]) + ((double)q[o10] + (double)q[o11]);
  }
  sh[threadIdx.x] = sum;
  __syncthreads();
  if (g == 0) {
    for (int k = 1; k < groups; ++k) sum += sh[k * c + j];
    gram[((size_t)blockIdx.y * c + i) * c + j] = (float)(sum * scale);
  }
}

template <int C2>
int launch_gram_split(TcContext& tc, const void* f_split, int nb, int hw, float* part, float* gram,
                      cudaStream_t s) {
  using Cfg = GramCfg<C2>;
  CUtensorMap map_f;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)C2, (cuuint64_t)hw, (cuuint64_t)nb};
    cuuint64_t gstride[2] = {(cuuint64_t)C2 * 2, (cuuint64_t)hw * C2 * 2};
    cuuint32_t box[3] = {64, 64, 1}, estride[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(tc.encode_fn)(
        &map_f, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(f_split), gdim, gstride, box,
        estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (split gram) failed with CUresult " + std::to_string((int)r));
      return ST_ERR_CUDA;
    }
  }
  GramArgs a{};
  a.hw = hw, a.kb_total = cdiv(hw, 64), a.part = part, a.half = 1;
  a.kb_per_split = kChainKb;
  a.nsplit = cdiv(a.kb_total, kChainKb);
  auto kern = gram_tc_kernel<C2>;
  ST_CUDA(tc_allow_smem(kern, Cfg::kSmemBytes));
  TimerScope ts(s, kTimeGram, 2.0 * (C2 / 2) * (C2 / 2) * hw * nb);
  ST_LAUNCH(kern, dim3(a.nsplit * Cfg::kMBlocks, nb), kGThreads, Cfg::kSmemBytes, s, map_f, a);
  const int c = C2 / 2;
  ST_LAUNCH(gram_fold_kernel, dim3(c, nb), 256, 0, s, part, a.nsplit, c, 1.0 / ((double)c * hw), gram);
  return ST_OK;
}

}  // namespace

bool gram_tc32_ok(const TcContext& tc, int c) {
  return tc.enabled && tc.pair_kernel && (c == 64 || c == 128 || c == 256);
}

size_t gram_tc32_part_floats(int nb, int hw, int c) {
  return (size_t)nb * cdiv(cdiv(hw, 64), kChainKb) * (2 * c) * (2 * c);
}

// gram[b] (full symmetric [C][C] fp32) = F_b^T F_b / (C * hw) from the [hi | lo] fp16 planes
// f_split [nb][hw][2C] of the fp32 features; part: gram_tc32_part_floats() floats of scratch.
int gram_tc32(TcContext& tc, const void* f_split, int nb, int hw, int c, float* part, float* gram,
              cudaStream_t s) {
  switch (c) {
    case 64: return launch_gram_split<128>(tc, f_split, nb, hw, part, gram, s);
    case 128: return launch_gram_split<256>(tc, f_split, nb, hw, part, gram, s);
    case 256: return launch_gram_split<512>(tc, f_split, nb, hw, part, gram, s);
  }
  set_error("gram_tc32: unsupported channel count");
  return ST_ERR_INVALID;
}

bool gram_tc_ok(const TcContext& tc, int c) {
  return tc.enabled && tc.pair_kernel && (c == 64 || c == 128 || c == 256 || c == 512);
}

size_t gram_tc_part_floats(const TcContext& tc, int nb, int hw, int c) {
  int nsplit = 1;
  switch (c) {
    case 64: nsplit = gram_splits<64>(hw); break;
    case 128: nsplit = gram_splits<128>(hw); break;
    case 256: nsplit = gram_splits<256>(hw); break;
    case 512: nsplit = gram_splits<512>(hw); break;
  }
  return (size_t)nb * nsplit * c * c;
}

int gram_tc_delta(TcContext& tc, const void* f, bool half, int nb, int hw, int c, float* part,
                  const float* target, float* delta, unsigned* max_bits, double* loss_part,
                  int* parts_per_tile, cudaStream_t s) {
  *parts_per_tile = (c / 32) * (c / 32);
  ST_REQUIRE((size_t)nb * *parts_per_tile <= (size_t)kMaxReduceBlocks * 4,
             "gram_tc_delta: batch too large for the reduction scratch");
  if (max_bits != nullptr) ST_CUDA(cudaMemsetAsync(max_bits, 0, nb * sizeof(unsigned), s));
  const GramFinish fin{target, delta, max_bits, loss_part};
  switch (c) {
    case 64: return launch_gram<64>(tc, f, half, nb, hw, part, fin, s);
    case 128: return launch_gram<128>(tc, f, half, nb, hw, part, fin, s);
    case 256: return launch_gram<256>(tc, f, half, nb, hw, part, fin, s);
    case 512: return launch_gram<512>(tc, f, half, nb, hw, part, fin, s);
  }
  set_error("gram_tc_delta: unsupported channel count");
  return ST_ERR_INVALID;
}

}  // namespace st
