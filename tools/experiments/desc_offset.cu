// experiment: can ONE copy of a convolution halo window serve all nine taps of a 3x3 convolution?
//
// The window is stored as pixel rows of 128 bytes (64 x 16-bit channels) with the absolute-address
// SWIZZLE_128B pattern TMA produces (16-byte chunk j of the row at byte offset o sits at chunk
// j ^ ((o >> 7) & 7)); window pixel (y, x) lives at row y*P + x (P = pitch in pixels).  The A operand
// of tap (dy, dx) for a 16 x 8 pixel tile is then the K-major matrix whose row m is window pixel
// ((m >> 3) + dy, (m & 7) + dx): 8-row groups P*128 bytes apart (stride-byte-offset = P*128) and
// a start address (dy*P + dx)*128 bytes into the window, i.e. NOT aligned to the 1024-byte swizzle
// period unless dx == 0 and P % 8 == 0.  Questions: does tcgen05.mma accept (a) a start address that
// is a multiple of 128 but not of 1024, with or without the descriptor's base-offset field, and
// (b) a stride-byte-offset that is not a multiple of 1024?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o desc_offset desc_offset.cu && ./desc_offset
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)(base_offset & 7) << 49) | ((uint64_t)2 << 61);
}

constexpr int kWinRows = 18, kMaxPitch = 16;

// win: [18][P][64] halves (dense, un-swizzled) ; B: [64][64] halves K-major ; D: [128][64] floats
__global__ void k(const __half* win, const __half* B, float* D, int P, int dy, int dx, int bo_mode) {
  __shared__ __align__(1024) uint8_t a_s[kWinRows * kMaxPitch * 128];
  __shared__ __align__(1024) uint8_t b_s[64 * 128];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < kWinRows * P * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *(uint4*)(a_s + r * 128 + ((j ^ (r & 7)) << 4)) = *(const uint4*)((const uint8_t*)win + r * 128 + j * 16);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *(uint4*)(b_s + r * 128 + ((j ^ (r & 7)) << 4)) = *(const uint4*)((const uint8_t*)B + r * 128 + j * 16);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32
  if (tid == 0) {
    const uint32_t a_start = smem_u32(a_s) + (uint32_t)(dy * P + dx) * 128u;
    const uint32_t bo = bo_mode ? ((a_start >> 7) & 7u) : 0u;
    const uint64_t da = desc(a_start, (uint32_t)P * 128u, bo), db = desc(smem_u32(b_s), 1024u, 0u);
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t acc = kk != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                   "l"(da + (uint64_t)(kk * 2)), "l"(db + (uint64_t)(kk * 2)), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0, spin = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (++spin > (1u << 22)) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5;
  for (int cc = 0; cc < 2; ++cc) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tm + cc * 32 + ((uint32_t)(warp * 32) << 16))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[tid * 64 + cc * 32 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}

int main() {
  std::vector<__half> B(64 * 64);
  std::vector<float> Bf(64 * 64);
  srand(1);
  for (int i = 0; i < 64 * 64; ++i) {
    B[i] = __float2half((rand() % 2001 - 1000) / 911.f);
    Bf[i] = __half2float(B[i]);
  }
  __half *dW, *dB;
  float* dD;
  cudaMalloc(&dW, kWinRows * kMaxPitch * 64 * 2);
  cudaMalloc(&dB, 64 * 64 * 2);
  cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dB, B.data(), 64 * 64 * 2, cudaMemcpyHostToDevice);
  const int pitches[3] = {8, 10, 16};
  for (int pi = 0; pi < 3; ++pi) {
    const int P = pitches[pi];
    std::vector<__half> W(kWinRows * P * 64);
    std::vector<float> Wf(W.size());
    for (size_t i = 0; i < W.size(); ++i) {
      W[i] = __float2half((rand() % 2001 - 1000) / 37.f);
      Wf[i] = __half2float(W[i]);
    }
    cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode)
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          if (dx + 8 > P) continue;                       // pitch 8 has no room for an x shift
          cudaMemset(dD, 0, 128 * 64 * 4);
          k<<<1, 128>>>(dW, dB, dD, P, dy, dx, bo_mode);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("P=%2d dy=%d dx=%d base_offset=%s: CUDA error %s\n", P, dy, dx, bo_mode ? "computed" : "0",
                   cudaGetErrorString(e));
            return 1;
          }
          std::vector<float> D(128 * 64);
          cudaMemcpy(D.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
          double maxerr = 0, maxref = 0;
          int bad_rows = 0;
          for (int m = 0; m < 128; ++m) {
            const int row = ((m >> 3) + dy) * P + (m & 7) + dx;
            double rowerr = 0;
            for (int n = 0; n < 64; ++n) {
              double s = 0;
              for (int kk = 0; kk < 64; ++kk) s += (double)Wf[row * 64 + kk] * Bf[n * 64 + kk];
              rowerr = fmax(rowerr, fabs(s - D[m * 64 + n]));
              maxref = fmax(maxref, fabs(s));
            }
            if (rowerr > 1e-2 * 50) ++bad_rows;
            maxerr = fmax(maxerr, rowerr);
          }
          printf("P=%2d dy=%d dx=%d base_offset=%-8s: max abs err %.4g (max ref %.4g) bad rows %d/128 %s\n", P, dy, dx,
                 bo_mode ? "computed" : "0", maxerr, maxref, bad_rows, bad_rows == 0 ? "OK" : "MISMATCH");
        }
  }
  return 0;
}
