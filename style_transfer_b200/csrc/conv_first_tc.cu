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
  const void* wk;                   // [3 kx][64 n][64 k'] K-major (see tc_pack_first_fwd)
  void* out;                        // [nb][h][w][64], bf16 or fp16 (half)
  uint32_t* bits;                   // ReLU bit mask of the output, [nb][h][w][2] words (may be null)
  int half;                         // operands and output are fp16
};

// Version 2.  The first version gathered all 27 neighbours per pixel-thread and split each into
// hi + lo there (27 loads, 54 conversions, 64 bias adds: ~800 SASS instructions per pixel, issue-bound
// at 2.5x the HBM floor).  Now a thread stages only ITS OWN pixel column: the 3 x 3 (ky, ci) values,
// hi + lo, as one K-major row  k' = [9 hi | 9 lo | 1, 1 | 0 ...]  of a 130-row array (pixels x0-1 ..
// x0+128).  The three x taps are three MMA groups whose A descriptor starts one ROW later each --
// the same rows, shifted by a pixel (start address + kx * 128 bytes inside the swizzle period, the
// trick of conv_tc2.cu) -- against three weight blocks B_kx.  The bias rides in the two "1" columns
// (hi + lo of the bias in B_1), so the epilogue is ReLU + pack + mask bits only.
constexpr int kFirstRows = 136;           // 130 used, padded to a multiple of 8

__global__ void __launch_bounds__(kFirstThreads)
conv_first_tc_kernel(const FirstArgs a) {
  // PDL: the weight staging and the TMEM allocation below overlap the previous kernel's tail
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ __align__(1024) uint8_t a_s[kFirstRows * 128];   // pixel rows, SWIZZLE_128B, K' = 32 used
  __shared__ __align__(1024) uint8_t b_s[3 * 64 * 128];       // weights per x tap
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  // weights -> swizzled smem: row n of block kx, 16-byte chunk j lands at chunk j ^ (n & 7)
  for (int i = tid; i < 3 * 64 * 8; i += kFirstThreads) {
    const int n = i >> 3, j = i & 7;                      // n = kx * 64 + row
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
  const uint32_t one16 = half ? 0x3C00u : 0x3F80u;        // 1.0 in fp16 / bf16
  uint32_t phase = 0;
  asm volatile("griddepcontrol.wait;" ::: "memory");     // the image is read (and `out` written) below

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int tx = tile % a.tiles_x, row_all = tile / a.tiles_x;
    const int y = row_all % a.h, b = row_all / a.h;
    const int x0 = tx * 128;
    // ---- stage: row r of a_s = pixel column x0 + r - 1 (rows 0 and 129 are the halo) -----------------
    // thread t owns row t + 1; threads 0 and 1 also build rows 0 and 129
    const int oy = a.img.oy[b], ox = a.img.ox[b];
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1 && tid > 1) break;
      const int r = pass == 0 ? tid + 1 : (tid == 0 ? 0 : 129);
      const int xx = x0 + r - 1;
      float v[9];                                         // k = ky * 3 + ci
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = 0.f;
      if (xx >= 0 && xx < a.w) {                           // zero padding of the TILE outside
        // the tile origin arrives reduced to [0, H) x [0, W) (host) and a tile is no larger than the
        // image, so the virtual roll wraps with one conditional add / subtract
        int cx = ox + xx;
        cx = cx >= a.img.W ? cx - a.img.W : cx;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1;
          if (yy < 0 || yy >= a.h) continue;
          int cy = oy + yy;
          cy = cy < 0 ? cy + a.img.H : (cy >= a.img.H ? cy - a.img.H : cy);
          const float* p = base + (size_t)cy * a.img.W + cx;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) v[ky * 3 + ci] = __ldg(p + ci * plane);
        }
      }
      // Pixels span +-150 grey levels: one 16-bit value would quantise them to 0.1-1 level.  Each
      // enters as hi + lo (lo = r16(v - hi)), the weights are repeated for the lo half.
      uint32_t h16[9], l16[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (half) {
          const __half hh = __float2half_rn(v[k]);
          const __half ll = __float2half_rn(v[k] - __half2float(hh));
          h16[k] = *reinterpret_cast<const uint16_t*>(&hh), l16[k] = *reinterpret_cast<const uint16_t*>(&ll);
        } else {
          const __nv_bfloat16 hh = __float2bfloat16_rn(v[k]);
          const __nv_bfloat16 ll = __float2bfloat16_rn(v[k] - __bfloat162float(hh));
          h16[k] = *reinterpret_cast<const uint16_t*>(&hh), l16[k] = *reinterpret_cast<const uint16_t*>(&ll);
        }
      }
      // 32 halfs: [h0..h8 | l0..l8 | 1 1 | 0 x 12]
      uint32_t kv[16];
      kv[0] = h16[0] | (h16[1] << 16), kv[1] = h16[2] | (h16[3] << 16), kv[2] = h16[4] | (h16[5] << 16);
      kv[3] = h16[6] | (h16[7] << 16), kv[4] = h16[8] | (l16[0] << 16), kv[5] = l16[1] | (l16[2] << 16);
      kv[6] = l16[3] | (l16[4] << 16), kv[7] = l16[5] | (l16[6] << 16), kv[8] = l16[7] | (l16[8] << 16);
      kv[9] = one16 | (one16 << 16);
#pragma unroll
      for (int k = 10; k < 16; ++k) kv[k] = 0u;
      uint8_t* rowp = a_s + r * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j)            // 16-byte chunk j of the row lands at j ^ (row & 7)
        *reinterpret_cast<uint4*>(rowp + ((j ^ (r & 7)) << 4)) =
            make_uint4(kv[4 * j], kv[4 * j + 1], kv[4 * j + 2], kv[4 * j + 3]);
    }
    fence_proxy_async();                  // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    // ---- MMA: D[128 px][64 ch] = sum_kx A[rows kx .. kx+127][32] * B_kx[64][32]^T ----------------------
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc_mma(tmem_base, da + (uint64_t)(kx * 8 + ks * 2), db + (uint64_t)(kx * 512 + ks * 2), idesc,
                   (kx | ks) != 0);
        tc_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: ReLU -> 16 bit -> a_s (free now: the MMAs have retired) -> coalesced copy ------------
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint8_t* row = a_s + tid * 128;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld32(taddr + cc * 32, r);
      // ReLU, saturation and rounding in the pack; the mask bits of the backward pass from the pairs
      uint32_t pw[16], bits = 0u;
      if (half) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pw[i] = pack16_relu(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]), true);
          bits |= gt2_mask<true>(pw[i], 0u) & (0x00010001u << i);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pw[i] = pack16_relu(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]), false);
          bits |= gt2_mask<false>(pw[i], 0u) & (0x00010001u << i);
        }
      }
      if (a.bits != nullptr && x0 + tid < a.w)
        a.bits[(((size_t)b * a.h + y) * a.w + x0 + tid) * 2 + cc] = bits;
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
      const int valid_rows = min(128, a.w - x0);
      uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(a.out) +
                                            (((size_t)b * a.h + y) * a.w + x0) * kFirstCout * 2);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int g = it * kFirstThreads + tid, rr = g >> 3, j = g & 7;
        if (rr < valid_rows)
          dst[g] = *reinterpret_cast<const uint4*>(a_s + rr * 128 + ((j ^ (rr & 7)) << 4));
      }
    }
    __syncthreads();                      // a_s and TMEM are free for the next tile
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

}  // namespace

// B_kx[n][k'] (K-major, 64 halfs per row): k' = ky*3 + ci -> w[n][ci][ky][kx] for the hi half (0..8)
// and again for the lo half (9..17); block kx = 1 also carries the bias as hi + lo in k' = 18, 19
// (the A rows hold 1.0 there).
int tc_pack_first_fwd(TcContext& tc, TcWeights& w, const float* w_host, int cout, bool half,
                      const float* bias_host) {
  if (!tc.enabled || !tc.pair_kernel || cout != kFirstCout) return ST_OK;
  auto r16 = [half](float v) {
    uint16_t bits;
    if (half) {
      const __half h = __float2half_rn(v);
      bits = *reinterpret_cast<const uint16_t*>(&h);
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      bits = *reinterpret_cast<const uint16_t*>(&h);
    }
    return bits;
  };
  auto f16 = [half](uint16_t bits) {
    if (half) return __half2float(*reinterpret_cast<const __half*>(&bits));
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&bits));
  };
  std::vector<uint16_t> host((size_t)3 * 64 * 64, 0);
  for (int co = 0; co < cout; ++co) {
    for (int ci = 0; ci < 3; ++ci)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const uint16_t bits = r16(w_host[((size_t)co * 3 + ci) * 9 + ky * 3 + kx]);
          uint16_t* row = host.data() + ((size_t)kx * 64 + co) * 64;
          row[ky * 3 + ci] = row[9 + ky * 3 + ci] = bits;
        }
    const uint16_t bh = r16(bias_host[co]);
    uint16_t* row1 = host.data() + ((size_t)1 * 64 + co) * 64;
    row1[18] = bh, row1[19] = r16(bias_host[co] - f16(bh));
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
  a.wk = w.fwd, a.out = out;
  (void)bias;                               // packed into the weights (tc_pack_first_fwd)
  const int grid = a.num_tiles < tc.sm_count * 5 ? a.num_tiles : tc.sm_count * 5;   // 42 KB of smem each
  TimerScope ts(s, kTimeConvSimt, 18.0 * 3 * kFirstCout * h * wd * img.nb);
  ST_LAUNCH(conv_first_tc_kernel, grid, kFirstThreads, 0, s, a);
  return ST_OK;
}

}  // namespace st
