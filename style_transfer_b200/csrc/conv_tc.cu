// tcgen05 / TMA implicit-GEMM 3x3 convolution (sm_100a), forward and backward-data.
//
//   out[p][n] = epilogue( sum_{tap, c} in[p + off(tap)][c] * Wk[n][tap*Cin + c] )
//
// GEMM view: M = pixels (CTA tile = BH x BW = 128 pixels), N = output channels (BN per CTA tile),
// K = 9*Cin swept in k-blocks of 128 bytes of channels of one tap.  Activations are NHWC, so for
// one tap the A operand of a k-block is a [BH][BW][128 B] box of the input shifted by the tap
// offset: one TMA tiled load (zero fill outside the image gives the padding and the ragged edges)
// lands it in shared memory in exactly the K-major SWIZZLE_128B layout tcgen05.mma consumes.  The
// weights are pre-packed K-major [N][9*Cin] and loaded the same way.
//
// Warp roles (192 threads, persistent CTAs, one per SM):
//   warp 0      TMA producer  (one elected lane)
//   warp 1      MMA issuer    (one elected lane; also owns the TMEM allocation)
//   warps 2..5  epilogue      (TMEM -> registers -> bias/ReLU or mask/inject -> swizzled smem
//                              -> TMA store); the accumulator is double-buffered in TMEM so the
//                              epilogue of tile i overlaps the main loop of tile i+1.
//
// The element type T is __nv_bfloat16 (kind::f16, K=16 per MMA) or float (kind::tf32, K=8 per
// MMA); everything is expressed in 128-byte K chunks so both share the code.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"

namespace st {

namespace {

constexpr int kTileM = 128;             // pixels per CTA tile == TMEM lanes
constexpr int kKBytes = 128;            // bytes of K per k-block row (one swizzle atom row)
constexpr int kThreads = 192;
constexpr uint32_t kSpinLimit = 1u << 26;

// ---- PTX wrappers ---------------------------------------------------------------------------------
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
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.
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
    if (spin > kSpinLimit) __trap();
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::
                   "l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;"); }
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
template <bool TF32>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
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

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct TcArgs {
  int h, w, cin, cout;
  int bw_log2;              // BW = 1 << bw_log2, BH = 128 / BW
  int tiles_x, tiles_y, tiles_n;
  int forward;
  const float* bias;
  const void* mask_act;     // T*, NHWC [h][w][cout]
  const void* inj;          // T*
};

template <typename T, int BN>
struct TcCfg {
  static constexpr bool kTf32 = sizeof(T) == 4;
  static constexpr int kElemsPerKB = kKBytes / (int)sizeof(T);         // 64 bf16 / 32 tf32
  static constexpr int kABytes = kTileM * kKBytes;                     // 16 KB
  static constexpr int kBBytes = BN * kKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutGroups = BN / kElemsPerKB;                  // 128-byte channel groups
  static constexpr int kOutBytes = kOutGroups * kTileM * kKBytes;
  static constexpr int kStages = (200 * 1024 - kOutBytes) / kStageBytes > 6
                                     ? 6
                                     : (200 * 1024 - kOutBytes) / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;                             // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*align*/ + 256;
  static constexpr uint32_t kIdesc = (1u << 4) | ((kTf32 ? 2u : 1u) << 7) |
                                     ((kTf32 ? 2u : 1u) << 10) | ((uint32_t)(BN >> 3) << 17) |
                                     ((uint32_t)(kTileM >> 4) << 24);
};

template <typename T, int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_out, const TcArgs a) {
  ST_PDL_ENTRY();
  using Cfg = TcCfg<T, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* out_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_base + Cfg::kOutBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bw = 1 << a.bw_log2, bh = kTileM >> a.bw_log2;
  const int kb_per_tap = a.cin / Cfg::kElemsPerKB;
  const int num_kb = 9 * kb_per_tap;
  const int num_tiles = a.tiles_x * a.tiles_y * a.tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_in), prefetch_tmap(&map_w), prefetch_tmap(&map_out);
    for (int i = 0; i < Cfg::kStages; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tfull[i], 1), mbar_init(&tempty[i], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % a.tiles_n, m_tile = tile / a.tiles_n;
        const int x0 = (m_tile % a.tiles_x) * bw, y0 = (m_tile / a.tiles_x) * bh;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / kb_per_tap, cb = kb % kb_per_tap;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          tma_load_3d(&map_in, &full[stage], sa, cb * Cfg::kElemsPerKB, x0 + tap % 3 - 1,
                      y0 + tap / 3 - 1);
          tma_load_2d(&map_w, &full[stage], sa + Cfg::kABytes,
                      tap * a.cin + cb * Cfg::kElemsPerKB, n_tile * BN);
          if (++stage == Cfg::kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&tempty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kKBytes / 32; ++k)       // 4 MMAs of 32 bytes of K each
            tc_mma<Cfg::kTf32>(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), Cfg::kIdesc,
                               (kb | k) != 0);
          tc_commit(&empty[stage]);                    // smem slot reusable once these MMAs retire
          if (++stage == Cfg::kStages) stage = 0, phase ^= 1;
        }
        tc_commit(&tfull[buf]);                        // accumulator complete
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    const int q = warp & 3;                            // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;                       // pixel row of the tile
    const bool issuer = threadIdx.x == 64;             // first epilogue thread issues TMA stores
    const T* mask_act = static_cast<const T*>(a.mask_act);
    const T* inj = static_cast<const T*>(a.inj);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, use = it >> 1;
      const int n_tile = tile % a.tiles_n, m_tile = tile / a.tiles_n;
      const int x0 = (m_tile % a.tiles_x) * bw, y0 = (m_tile / a.tiles_x) * bh;
      const int py = y0 + (m >> a.bw_log2), px = x0 + (m & (bw - 1));
      const bool valid = py < a.h && px < a.w;
      const size_t gofs = ((size_t)py * a.w + px) * a.cout + (size_t)n_tile * BN;

      if (issuer) tma_store_wait_read();               // staging buffer free again
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (a.forward) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = fmaxf(v[i] + __ldg(a.bias + n_tile * BN + cc * 32 + i), 0.f);
        } else {
          if (mask_act != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
              if (valid) mk = Store<T>::ld4(mask_act + gofs + cc * 32 + i);
              v[i] = mk.x > 0.f ? v[i] : 0.f, v[i + 1] = mk.y > 0.f ? v[i + 1] : 0.f;
              v[i + 2] = mk.z > 0.f ? v[i + 2] : 0.f, v[i + 3] = mk.w > 0.f ? v[i + 3] : 0.f;
            }
          }
          if (inj != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
              if (valid) e = Store<T>::ld4(inj + gofs + cc * 32 + i);
              v[i] += e.x, v[i + 1] += e.y, v[i + 2] += e.z, v[i + 3] += e.w;
            }
          }
        }
        // registers -> swizzled staging tile [group][128 rows][128 B] (SWIZZLE_128B, as TMA expects)
        if constexpr (Cfg::kTf32) {
          uint8_t* row = out_base + (size_t)cc * (kTileM * kKBytes) + (size_t)m * kKBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(row + ((j ^ (m & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          uint8_t* row = out_base + (size_t)(cc >> 1) * (kTileM * kKBytes) + (size_t)m * kKBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0), pk.y = *reinterpret_cast<uint32_t*>(&p1);
            pk.z = *reinterpret_cast<uint32_t*>(&p2), pk.w = *reinterpret_cast<uint32_t*>(&p3);
            const int chunk = (cc & 1) * 4 + j;
            *reinterpret_cast<uint4*>(row + ((chunk ^ (m & 7)) << 4)) = pk;
          }
        }
      }
      // all TMEM reads of this buffer are done: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
#pragma unroll
        for (int g = 0; g < Cfg::kOutGroups; ++g)
          tma_store_3d(&map_out, out_base + (size_t)g * (kTileM * kKBytes),
                       n_tile * BN + g * Cfg::kElemsPerKB, x0, y0);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(const TcContext& tc, CUtensorMap* map, bool is_f32, int rank, void* base,
               const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t bdim[3], estride[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], bdim[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_bytes[i];
  CUresult r = reinterpret_cast<EncodeTiledFn>(tc.encode_fn)(
      map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base,
      gdim, gstride, bdim, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

template <typename T, int BN>
int launch(TcContext& tc, const CUtensorMap& map_w, const T* in, T* out, const TcArgs& base_args,
           cudaStream_t s) {
  using Cfg = TcCfg<T, BN>;
  TcArgs a = base_args;
  // pixel-tile shape: BW x BH = 128, minimising padded work for this feature-map size
  long best = -1;
  for (int l2 = 3; l2 <= 7; ++l2) {
    const int bw = 1 << l2, bh = kTileM >> l2;
    const long cost = (long)cdiv(a.w, bw) * bw * cdiv(a.h, bh) * bh;
    if (best < 0 || cost <= best) best = cost, a.bw_log2 = l2;   // prefer wide rows on ties
  }
  const int bw = 1 << a.bw_log2, bh = kTileM >> a.bw_log2;
  a.tiles_x = cdiv(a.w, bw), a.tiles_y = cdiv(a.h, bh), a.tiles_n = a.cout / BN;
  const bool f32 = sizeof(T) == 4;
  CUtensorMap map_in, map_out;
  {
    const uint64_t dims[3] = {(uint64_t)a.cin, (uint64_t)a.w, (uint64_t)a.h};
    const uint64_t strides[2] = {(uint64_t)a.cin * sizeof(T), (uint64_t)a.w * a.cin * sizeof(T)};
    const uint32_t box[3] = {(uint32_t)Cfg::kElemsPerKB, (uint32_t)bw, (uint32_t)bh};
    int rc = encode_map(tc, &map_in, f32, 3, const_cast<T*>(in), dims, strides, box);
    if (rc != ST_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.cout, (uint64_t)a.w, (uint64_t)a.h};
    const uint64_t strides[2] = {(uint64_t)a.cout * sizeof(T), (uint64_t)a.w * a.cout * sizeof(T)};
    const uint32_t box[3] = {(uint32_t)Cfg::kElemsPerKB, (uint32_t)bw, (uint32_t)bh};
    int rc = encode_map(tc, &map_out, f32, 3, out, dims, strides, box);
    if (rc != ST_OK) return rc;
  }
  auto kern = conv3x3_tc_kernel<T, BN>;
  ST_CUDA(tc_allow_smem(kern, Cfg::kSmemBytes));
  const int tiles = a.tiles_x * a.tiles_y * a.tiles_n;
  const int grid = tiles < tc.sm_count ? tiles : tc.sm_count;
  TimerScope ts(s, kTimeConvTc, 18.0 * a.cin * a.cout * a.h * a.w);
  ST_LAUNCH(kern, grid, kThreads, Cfg::kSmemBytes, s, map_in, map_w, map_out, a);
  return ST_OK;
}

template <typename T>
int pack_one(const TcContext& tc, const float* w, int cin, int cout, bool backward, T** dev,
             void** map_host) {
  // rows = output channels of this direction, K = 9 * (input channels of this direction)
  const int rows = backward ? cin : cout, kc = backward ? cout : cin;
  std::vector<T> host((size_t)rows * 9 * kc);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const float v = w[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx];
          size_t idx;
          if (!backward)
            idx = (size_t)co * 9 * cin + (size_t)(ky * 3 + kx) * cin + ci;
          else
            idx = (size_t)ci * 9 * cout + (size_t)((2 - ky) * 3 + (2 - kx)) * cout + co;
          if constexpr (sizeof(T) == 4)
            host[idx] = v;
          else
            host[idx] = __float2bfloat16_rn(v);
        }
  if (!*dev) ST_CUDA(cudaMalloc((void**)dev, host.size() * sizeof(T)));
  ST_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  if (!*map_host) *map_host = new CUtensorMap;
  const int bn = rows % 128 == 0 ? 128 : 64;
  const uint64_t dims[2] = {(uint64_t)9 * kc, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)9 * kc * sizeof(T)};
  const uint32_t box[2] = {(uint32_t)(kKBytes / sizeof(T)), (uint32_t)bn};
  return encode_map(tc, static_cast<CUtensorMap*>(*map_host), sizeof(T) == 4, 2, *dev, dims,
                    strides, box);
}

}  // namespace

cudaError_t tc_allow_smem_impl(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<const void*, uint64_t> done;      // kernel -> bit mask of devices already set
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  uint64_t& mask = done[kernel];
  if (dev < 64 && (mask >> dev) & 1) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) mask |= (uint64_t)1 << dev;
  return e;
}

int tc_init(TcContext& tc, int sm_count) {
  tc.sm_count = sm_count;
  if (getenv("ST_DISABLE_TC") != nullptr) {   // debugging aid: bf16 storage with the SIMT convolution
    tc.enabled = false;
    return ST_OK;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  ST_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return ST_ERR_CUDA;
  }
  tc.encode_fn = fn;
  tc.enabled = true;
  tc.pair_kernel = getenv("ST_CONV_V1") == nullptr;
  tc.resident_weights = getenv("ST_TC_NO_RESB") == nullptr;
  tc.defer_scale = getenv("ST_NO_DEFER") == nullptr;
  tc.pool_fusion = getenv("ST_NO_POOL_FUSION") == nullptr;
  tc.pix_rows_kernel = getenv("ST_NO_PIX_ROWS") == nullptr;
  tc.fwd_bits = getenv("ST_NO_FWD_BITS") == nullptr;
  tc.pdl = getenv("ST_NO_PDL") == nullptr;
  if (const char* f = getenv("ST_TC_BN")) tc.force_bn = atoi(f);
  return ST_OK;
}

void tc_destroy(TcContext& tc) { tc.enabled = false; }

int tc_pack_weights(TcContext& tc, TcWeights& w, const float* w_host, int cin, int cout,
                    bool fwd_half) {
  if (!tc.enabled) return ST_OK;
  w.fwd_half = fwd_half;
  int rc;
  if (fwd_half) {
    // forward weights as fp16 [cout][9*cin] (the pair kernel builds its own tensor maps)
    std::vector<__half> host((size_t)cout * 9 * cin);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < 9; ++t)
          host[(size_t)co * 9 * cin + (size_t)t * cin + ci] =
              __float2half_rn(w_host[((size_t)co * cin + ci) * 9 + t]);
    if (!w.fwd) ST_CUDA(cudaMalloc(&w.fwd, host.size() * sizeof(__half)));
    ST_CUDA(cudaMemcpy(w.fwd, host.data(), host.size() * sizeof(__half), cudaMemcpyHostToDevice));
    rc = ST_OK;
  } else {
    rc = pack_one<__nv_bfloat16>(tc, w_host, cin, cout, false,
                                 reinterpret_cast<__nv_bfloat16**>(&w.fwd), &w.map_fwd);
  }
  if (rc == ST_OK) rc = pack_one<__nv_bfloat16>(tc, w_host, cin, cout, true, &w.bwd, &w.map_bwd);
  return rc;
}

int tc_pack_first(TcContext& tc, TcWeights& w, const float* w_host, int cout) {
  if (!tc.enabled || !tc.pair_kernel) return ST_OK;
  const int rows = 16;
  std::vector<__nv_bfloat16> host((size_t)rows * 9 * cout, __float2bfloat16_rn(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < 3; ++ci)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
          host[(size_t)ci * 9 * cout + (size_t)((2 - ky) * 3 + (2 - kx)) * cout + co] =
              __float2bfloat16_rn(w_host[(((size_t)co * 3 + ci) * 3 + ky) * 3 + kx]);
  if (!w.bwd) ST_CUDA(cudaMalloc((void**)&w.bwd, host.size() * sizeof(__nv_bfloat16)));
  ST_CUDA(cudaMemcpy(w.bwd, host.data(), host.size() * sizeof(__nv_bfloat16),
                     cudaMemcpyHostToDevice));
  return ST_OK;
}

void tc_free_weights(TcWeights& w) {
  cudaFree(w.fwd), cudaFree(w.bwd), cudaFree(w.bwd_rows), cudaFree(w.fwd32), cudaFree(w.bwd32);
  delete static_cast<CUtensorMap*>(w.map_fwd);
  delete static_cast<CUtensorMap*>(w.map_bwd);
  w = TcWeights{};
}

bool tc_shape_ok(const TcContext& tc, const TcWeights& w, int cin, int cout) {
  return tc.enabled && w.fwd != nullptr && cin % 64 == 0 && cout % 64 == 0;
}

int conv3x3_tc(TcContext& tc, const TcWeights& w, const void* in_v, void* out_v, int nb, int h,
               int wd, int cin, int cout, bool forward, const float* bias, const void* mask_v,
               uint32_t* relu_bits, const __nv_bfloat16* inj, const float* inj_scale,
               cudaStream_t s) {
  if (tc.pair_kernel)
    return conv3x3_tc_pair(tc, w, in_v, out_v, nb, h, wd, cin, cout, forward, bias, relu_bits, inj,
                           inj_scale, s);
  ST_REQUIRE(nb == 1 && inj_scale == nullptr && !w.fwd_half,
             "the single-CTA convolution kernel (ST_CONV_V1) takes one bf16 tile and a pre-scaled injection");
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(in_v);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_v);
  const __nv_bfloat16* mask_act = static_cast<const __nv_bfloat16*>(mask_v);
  TcArgs a{};
  a.h = h, a.w = wd, a.cin = cin, a.cout = cout, a.forward = forward ? 1 : 0;
  a.bias = bias, a.mask_act = mask_act, a.inj = inj;
  const CUtensorMap& map_w = *static_cast<const CUtensorMap*>(forward ? w.map_fwd : w.map_bwd);
  if (cout % 128 == 0) return launch<__nv_bfloat16, 128>(tc, map_w, in, out, a, s);
  return launch<__nv_bfloat16, 64>(tc, map_w, in, out, a, s);
}

}  // namespace st
