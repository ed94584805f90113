// Shared device/host helpers for libstyle_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

namespace st {

constexpr float kEps = 1.1920928955078125e-07f;  // float32 machine epsilon (num_utils.py:14)

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;
extern bool g_pdl;          // ST_NO_PDL=1 disables programmatic dependent launch (read once)
extern bool g_pdl_all;      // ST_PDL_ALL=1: for every kernel, not only the convolution kernels

#define ST_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      st::set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                    ":" + std::to_string(__LINE__) + ")");                                 \
      return ST_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define ST_REQUIRE(cond, msg)                      \
  do {                                             \
    if (!(cond)) {                                 \
      st::set_error(std::string("invalid: ") + msg); \
      return ST_ERR_INVALID;                       \
    }                                              \
  } while (0)

// Every kernel launch of the library goes through this so bench.py can report "gpu_launches".
#define ST_LAUNCH(kernel, grid, block, smem, strm_, ...)                                   \
  do {                                                                                     \
    cudaLaunchConfig_t cfg_ = {};                                                          \
    cfg_.gridDim = dim3(grid), cfg_.blockDim = dim3(block);                                \
    cfg_.dynamicSmemBytes = (smem), cfg_.stream = (strm_);                                 \
    cudaLaunchAttribute attr_[1];                                                          \
    attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
    attr_[0].val.programmaticStreamSerializationAllowed = st::g_pdl_all ? 1 : 0;           \
    cfg_.attrs = attr_, cfg_.numAttrs = 1;                                                 \
    st::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
    ST_CUDA(cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__));                               \
  } while (0)

// launch with the attribute decided by the caller (the convolution kernels: on unless ST_NO_PDL=1)
#define ST_LAUNCH_ATTR(kernel, grid, block, smem, strm_, pdl_, ...)                        \
  do {                                                                                     \
    cudaLaunchConfig_t cfg_ = {};                                                          \
    cfg_.gridDim = dim3(grid), cfg_.blockDim = dim3(block);                                \
    cfg_.dynamicSmemBytes = (smem), cfg_.stream = (strm_);                                 \
    cudaLaunchAttribute attr_[1];                                                          \
    attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
    attr_[0].val.programmaticStreamSerializationAllowed = (pdl_) ? 1 : 0;                  \
    cfg_.attrs = attr_, cfg_.numAttrs = 1;                                                 \
    st::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
    ST_CUDA(cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__));                               \
  } while (0)

// Programmatic dependent launch: every kernel of the library starts with ST_PDL_ENTRY(), so that any
// launch MAY carry the programmatic-stream-serialization attribute.  By default only the convolution
// kernels do (ST_LAUNCH_ATTR in conv_tc2.cu); ST_PDL_ALL=1 switches it on for every launch -- that
// setting hung the 2048^2 benchmark loop on the B200 (not the small test cases) and stays
// experimental.  launch_dependents lets
// the NEXT kernel of the stream be scheduled as soon as all blocks of this one are resident or done;
// griddepcontrol.wait blocks until the PREVIOUS kernel has completed and its writes are visible, so
// the ordering the code relies on is unchanged -- only the launch latency (and, where the wait is
// placed after a prologue, that prologue) overlaps the predecessor's tail.  A step is ~62 dependent
// launches; the gaps between them were ~0.3 ms of a 6 ms step (profiles/r02_launches_step_b.md).
// The early trigger is only issued by grids that are certainly resident as a whole (at most half the
// thread capacity of the device, kernels without large shared memory): with ST_PDL_ALL=1 an early
// trigger from a multi-wave grid (16-tile batches: 1000-4000 blocks) hung the device.
#define ST_PDL_ENTRY()                                                                        \
  do {                                                                                        \
    if ((size_t)gridDim.x * gridDim.y * gridDim.z * (blockDim.x * blockDim.y * blockDim.z) <= \
        (size_t)148 * 1024)                                                                   \
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");                         \
    asm volatile("griddepcontrol.wait;" ::: "memory");                                        \
  } while (0)

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- opt-in per-kernel timing (st_timing_* in style_b200.h) -------------------------------------
// A launcher declares `TimerScope t(stream, category, work)`; when timing is enabled the scope
// records a CUDA event on `stream` before and after the launches it covers.  `work` is the
// ALGORITHMIC work of those launches: flops for the tensor categories, bytes for the HBM ones.
enum TimingCategory {
  kTimeConvTc = 0,     // tcgen05 implicit-GEMM convolutions (flops)
  kTimeConvSimt = 1,   // the 3-channel first/last layer (any kernel) and the fp32 SIMT convolutions (flops)
  kTimePool = 2,       // pooling forward / backward (bytes)
  kTimeGram = 3,       // Gram F^T F (flops)
  kTimeStyleGrad = 4,  // delta-Gram x F (flops)
  kTimeLoss = 5,       // content / dd / gram-delta statistics and gradient injection (bytes)
  kTimeImage = 6,      // regularisers, optimizers, gradient unpack (bytes)
  kTimeCategories = 7
};
extern bool g_timing_enabled;
void timing_mark(cudaStream_t s, int category, double work, bool begin);
struct TimerScope {
  cudaStream_t s;
  int cat;
  bool on;
  TimerScope(cudaStream_t stream, int category, double work)
      : s(stream), cat(category), on(g_timing_enabled) {
    if (on) timing_mark(s, cat, work, true);
  }
  ~TimerScope() {
    if (on) timing_mark(s, cat, 0.0, false);
  }
};

// ---- storage-type traits: activations are NHWC in float or bf16 ------------------------------------
template <typename T> struct Store;
template <> struct Store<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float4 ld4(const float* p) {
    return *reinterpret_cast<const float4*>(p);
  }
  static __device__ __forceinline__ void st4(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
  }
};
template <> struct Store<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
  }
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&a);
    raw.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = raw;
  }
};

template <> struct Store<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ float4 ld4(const __half* p) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    const float2 fa = __half22float2(*reinterpret_cast<__half2*>(&raw.x));
    const float2 fb = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st4(__half* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&a);
    raw.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = raw;
  }
};

// Division of n < 2^31 by a runtime constant without the ~30-instruction integer divide:
// q = (umulhi(n, M) + n) >> s with s = ceil(log2 d), M = floor(2^32 * (2^s - d) / d) + 1.
struct FastDiv {
  unsigned d, M, s;
  FastDiv() : d(1), M(1), s(0) {}
  explicit FastDiv(unsigned div) : d(div) {
    s = 0;
    while ((1u << s) < d) ++s;
    M = (unsigned)((((unsigned long long)1 << 32) * (((unsigned long long)1 << s) - d)) / d + 1);
  }
  __device__ __forceinline__ unsigned div(unsigned n) const { return (__umulhi(n, M) + n) >> s; }
  __device__ __forceinline__ unsigned mod(unsigned n) const { return n - div(n) * d; }
};

// 8 consecutive channels at once: 16 bytes of bf16 / 32 bytes of float
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  return F8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void st8(float* p, const F8& f) {
  *reinterpret_cast<float4*>(p) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f.v[4], f.v[5], f.v[6], f.v[7]);
}
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  F8 f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f.v[2 * i] = __uint_as_float(w[i] << 16);
    f.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
  return f;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f.v[2 * i], f.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ F8 ld8(const __half* p) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  F8 f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f.v[2 * i] = t.x, f.v[2 * i + 1] = t.y;
  }
  return f;
}
__device__ __forceinline__ void st8(__half* p, const F8& f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(f.v[2 * i], f.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
// two floats <-> one 32-bit word of 16-bit storage; `half` selects fp16 (true) or bf16 (false)
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool half) {
  if (half) {
    // saturate instead of producing inf: fp16 tops out at 65504
    __half2 h = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f),
                                  fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&b);
}
__device__ __forceinline__ float2 unpack16(uint32_t w, bool half) {
  if (half) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}

// ReLU + saturation + rounding + packing of two fp32 values in one F2FP (hi lands in the upper half)
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi, bool half) {
  uint32_t d;
  if (half)
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// packed 16-bit pairs: per-half (a > b) ? 0xFFFF : 0, and the per-half maximum
template <bool HALF>
__device__ __forceinline__ uint32_t gt2_mask(uint32_t a, uint32_t b) {
  if constexpr (HALF)
    return __hgt2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  else
    return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a),
                       *reinterpret_cast<const __nv_bfloat162*>(&b));
}
template <bool HALF>
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  if constexpr (HALF) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                     *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}

// ---- deterministic grid reductions ----------------------------------------------------------------
// Each block reduces K doubles, publishes them, takes a ticket; the last block to arrive sums the
// per-block partials in index order (fixed grid => bit-reproducible) and returns true on its
// thread 0 with the totals in `v`.  `partials` must hold gridDim.x*K doubles; `*counter` must be 0
// on entry and is reset to 0 on exit.  All threads of the block must call it.
template <int K>
__device__ __forceinline__ bool grid_reduce_impl(double (&v)[K], double* partials, unsigned* counter,
                                                 unsigned block_id, unsigned num_blocks);
template <int K>
__device__ __forceinline__ bool grid_reduce(double (&v)[K], double* partials, unsigned* counter) {
  return grid_reduce_impl<K>(v, partials, counter, blockIdx.x, gridDim.x);
}
// same, for a (gridDim.x, gridDim.y) grid reduced as one
template <int K>
__device__ __forceinline__ bool grid_reduce_2d(double (&v)[K], double* partials, unsigned* counter) {
  return grid_reduce_impl<K>(v, partials, counter, blockIdx.y * gridDim.x + blockIdx.x,
                             gridDim.x * gridDim.y);
}
template <int K>
__device__ __forceinline__ bool grid_reduce_impl(double (&v)[K], double* partials, unsigned* counter,
                                                 unsigned block_id, unsigned num_blocks) {
  __shared__ double sh[K][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[k][warp] = x;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = lane < nwarps ? sh[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) partials[(size_t)block_id * K + k] = x;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    is_last = (ticket == num_blocks - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = 0.0;
    for (unsigned b = threadIdx.x; b < num_blocks; b += blockDim.x)
      x += partials[(size_t)b * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) sh[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = 0.0;
      for (int w = 0; w < nwarps; ++w) x += sh[k][w];
      v[k] = x;
    }
    *counter = 0u;
    return true;
  }
  return false;
}

__device__ __forceinline__ int wrap(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

}  // namespace st
