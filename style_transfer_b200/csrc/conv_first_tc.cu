// First convolution (3 image planes -> 64 channels, conv1_1) on tcgen05 tensor cores.
//
// K = 27 is far too short for the TMA/tap-reuse pipeline of conv_tc2.cu and the input is the
// planar f32 image (read through the virtual roll), so this kernel builds the im2col operand itself:
// each of the 128 threads of a CTA gathers the 27 neighbours of one pixel, converts them to bf16
// and writes one 128-byte K-major SWIZZLE_128B row (K padded to 64 with zeros).  One elected thread
// issues four M=128 x N=64 x K=16 MMAs; the epilogue (bias + ReLU -> bf16) writes each pixel's 64
// channels as one full 128-byte line.  The kernel is bound by that 64-channel write (HBM).
// Several CTAs share an SM (24 KB of shared memory, 64 TMEM columns each) and overlap each other's
// gather / MMA / store phases; tiles are 1 row x 128 pixels so the gather is coalesced.
#include <cuda.h>

#include <vector>

#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"

namespace st {

namespace {

constexpr int kFirstThreads = 128;
constexpr int kFirstCout = 64;
constexpr uint32_t kSpinF = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
    if (spin > kSpinF) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// c_format f32 | K-major | N = 64 | M = 128; operand formats (bits 7, 10: 0 = f16, 1 = bf16) at run time
constexpr uint32_t kFirstIdescBase = (1u << 4) | ((uint32_t)(kFirstCout >> 3) << 17) |
                                     ((uint32_t)(128 >> 4) << 24);

struct FirstArgs {
  ImageBatch img;
  int h, w;                         // tile size
  int tiles_x;                      // ceil(w / 128)
  int num_tiles;                    // nb * h * tiles_x
  const void* wk;                   // [64][64] K-major, k = tap*3 + ci for the hi half, 27 + that for lo
  const float* bias;
  void* out;                        // [nb][h][w][64], bf16 or fp16 (half)
  uint32_t* bits;                   // ReLU bit mask of the output, [nb][h][w][2] words (may be null)
  int half;                         // operands and output are fp16
};

__global__ void __launch_bounds__(kFirstThreads)
conv_first_tc_kernel(const FirstArgs a) {
  __shared__ __align__(1024) uint8_t a_s[128 * 128];     // im2col rows, SWIZZLE_128B
  __shared__ __align__(1024) uint8_t b_s[64 * 128];      // weights
  __shared__ float bias_s[kFirstCout];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < kFirstCout) bias_s[tid] = a.bias[tid];

  // weights -> swizzled smem: row n, 16-byte chunk j lands at chunk j ^ (n & 7)
  for (int i = tid; i < 64 * 8; i += kFirstThreads) {
    const int n = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(b_s + n * 128 + ((j ^ (n & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(a.wk) + n * 128 + j * 16);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint64_t db = make_smem_desc(smem_u32(b_s)), da = make_smem_desc(smem_u32(a_s));
  const bool half = a.half != 0;
  const uint32_t fmt = half ? 0u : 1u;
  const uint32_t idesc = kFirstIdescBase | (fmt << 7) | (fmt << 10);
  const float* base = a.img.base;
  const size_t plane = (size_t)a.img.H * a.img.W;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int tx = tile % a.tiles_x, row_all = tile / a.tiles_x;
    const int y = row_all % a.h, b = row_all / a.h;
    const int x = tx * 128 + tid;
    // ---- gather: 27 neighbours of pixel (y, x) of tile b -> one bf16 K-major row --------------------
    float v[32];                           // k = tap * 3 + ci; k >= 27 is zero padding
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
    if (x < a.w) {
      // the tile origin arrives reduced to [0, H) x [0, W) (host) and a tile is no larger than the
      // image, so the virtual roll wraps with one conditional add / subtract (an integer modulo per
      // neighbour was a fifth of this kernel's instructions)
      const int oy = a.img.oy[b], ox = a.img.ox[b];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= a.h) continue;                      // zero padding of the TILE
        int cy = oy + yy;                                       // virtual roll of the image
        cy = cy < 0 ? cy + a.img.H : (cy >= a.img.H ? cy - a.img.H : cy);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          if (xx < 0 || xx >= a.w) continue;
          int cx = ox + xx;
          cx = cx < 0 ? cx + a.img.W : (cx >= a.img.W ? cx - a.img.W : cx);
          const float* p = base + (size_t)cy * a.img.W + cx;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) v[(ky * 3 + kx) * 3 + ci] = __ldg(p + ci * plane);
        }
      }
    }
    // Pixels span +-150 grey levels: one bf16 (8 significant bits) would quantise them to 0.5-1
    // level.  Each value enters as hi + lo (lo = bf16(v - hi)), the weights are repeated for the lo
    // half: K = 54 of the 64 padded columns, ~16 significant bits, no extra MMA.
    uint32_t kv[32];                       // 64 x 16 bit: k in [0,27) hi, [27,54) lo, rest zero
    {
      float hi[27], lo[27];
      if (half) {
#pragma unroll
        for (int k = 0; k < 27; ++k) hi[k] = __half2float(__float2half_rn(v[k])), lo[k] = v[k] - hi[k];
      } else {
#pragma unroll
        for (int k = 0; k < 27; ++k)
          hi[k] = __bfloat162float(__float2bfloat16_rn(v[k])), lo[k] = v[k] - hi[k];
      }
      float xk[64];
#pragma unroll
      for (int k = 0; k < 64; ++k) xk[k] = k < 27 ? hi[k] : (k < 54 ? lo[k - 27] : 0.f);
      // pixels are a few hundred grey levels at most: no saturation needed (pack16 clamps fp16)
      if (half) {                            // one uniform branch, not one per packed word
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const __half2 h2 = __floats2half2_rn(xk[2 * k], xk[2 * k + 1]);
          kv[k] = *reinterpret_cast<const uint32_t*>(&h2);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) kv[k] = pack16(xk[2 * k], xk[2 * k + 1], false);
      }
    }
    uint8_t* row = a_s + tid * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)            // 16-byte chunk j of the row lands at j ^ (row & 7)
      *reinterpret_cast<uint4*>(row + ((j ^ (tid & 7)) << 4)) =
          make_uint4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
    fence_proxy_async();                  // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    // ---- MMA: D[128 px][64 ch] = A[128][64] * B[64][64]^T ---------------------------------------------
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k != 0);
        tc_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: bias + ReLU -> bf16 -> a_s (free now: the MMAs have retired) -> coalesced copy --------
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld32(taddr + cc * 32, r);
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]) + bias_s[cc * 32 + i];
      // ReLU, saturation and rounding in the pack; the mask bits of the backward pass from the pairs
      uint32_t pw[16], bits = 0u;
      if (half) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pw[i] = pack16_relu(f[2 * i], f[2 * i + 1], true);
          bits |= gt2_mask<true>(pw[i], 0u) & (0x00010001u << i);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pw[i] = pack16_relu(f[2 * i], f[2 * i + 1], false);
          bits |= gt2_mask<false>(pw[i], 0u) & (0x00010001u << i);
        }
      }
      if (a.bits != nullptr && tx * 128 + tid < a.w)
        a.bits[(((size_t)b * a.h + y) * a.w + tx * 128 + tid) * 2 + cc] = bits;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(row + (((cc * 4 + j) ^ (tid & 7)) << 4)) =
            make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
    }
    tc_fence_before();
    __syncthreads();
    // the 128 pixels of the tile are 16 KB of contiguous global memory: 16-byte chunk g of the tile
    // is chunk (g & 7) of row (g >> 3)
    {
      const int valid_rows = min(128, a.w - tx * 128);
      uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(a.out) +
                                            (((size_t)b * a.h + y) * a.w + tx * 128) * kFirstCout * 2);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int g = it * kFirstThreads + tid, r = g >> 3, j = g & 7;
        if (r < valid_rows)
          dst[g] = *reinterpret_cast<const uint4*>(a_s + r * 128 + ((j ^ (r & 7)) << 4));
      }
    }
    __syncthreads();                      // a_s and TMEM are free for the next tile
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

}  // namespace

int tc_pack_first_fwd(TcContext& tc, TcWeights& w, const float* w_host, int cout, bool half) {
  if (!tc.enabled || !tc.pair_kernel || cout != kFirstCout) return ST_OK;
  std::vector<uint16_t> host((size_t)64 * 64, 0);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < 3; ++ci)
      for (int tap = 0; tap < 9; ++tap) {
        const float v = w_host[((size_t)co * 3 + ci) * 9 + tap];
        uint16_t bits;
        if (half) {
          const __half h = __float2half_rn(v);
          bits = *reinterpret_cast<const uint16_t*>(&h);
        } else {
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          bits = *reinterpret_cast<const uint16_t*>(&h);
        }
        host[(size_t)co * 64 + tap * 3 + ci] = host[(size_t)co * 64 + 27 + tap * 3 + ci] = bits;
      }
  if (!w.fwd) ST_CUDA(cudaMalloc(&w.fwd, host.size() * 2));
  ST_CUDA(cudaMemcpy(w.fwd, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
  w.fwd_half = half;
  return ST_OK;
}

int conv_first_fwd_tc(TcContext& tc, const TcWeights& w, const ImageBatch& img, int h, int wd,
                      const float* bias, void* out, uint32_t* relu_bits, cudaStream_t s) {
  ST_REQUIRE(h <= img.H && wd <= img.W, "conv_first_fwd_tc: tile larger than the image");
  FirstArgs a{};
  a.bits = relu_bits;
  a.half = w.fwd_half ? 1 : 0;
  a.img = img, a.h = h, a.w = wd, a.tiles_x = cdiv(wd, 128);
  for (int i = 0; i < img.nb; ++i) {       // tile origins reduced to the image (see the gather)
    a.img.oy[i] = ((img.oy[i] % img.H) + img.H) % img.H;
    a.img.ox[i] = ((img.ox[i] % img.W) + img.W) % img.W;
  }
  a.num_tiles = img.nb * h * a.tiles_x;
  a.wk = w.fwd, a.bias = bias, a.out = out;
  const int grid = a.num_tiles < tc.sm_count * 8 ? a.num_tiles : tc.sm_count * 8;   // 8 x 64 TMEM columns
  TimerScope ts(s, kTimeConvSimt, 18.0 * 3 * kFirstCout * h * wd * img.nb);
  ST_LAUNCH(conv_first_tc_kernel, grid, kFirstThreads, 0, s, a);
  return ST_OK;
}

}  // namespace st
