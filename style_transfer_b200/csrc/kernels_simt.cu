// SIMT kernels of libstyle_b200: the exact-fp32 convolution path, pooling, Gram / style / content
// losses and the layout conversions.  Activations are NHWC ([h][w][C], C contiguous) of type T
// (float in ST_PREC_FP32, bf16 in ST_PREC_BF16); all arithmetic is fp32, reductions finish in
// double.  Reference semantics: style_transfer.py:556-612 (losses), Caffe layers (see oracle/).
#include <float.h>

#include "style_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace st {

// =====================================================================================================
// 3x3 / pad 1 convolution as an implicit GEMM on CUDA cores.
//   CTA tile: 8 rows x 16 cols of pixels (M = 128) x 64 output channels, K swept in chunks of CI
//   input channels x 9 taps.  Thread (tx, ty): 4 consecutive output channels x 8 consecutive pixels
//   of one row, so each shared-memory input value feeds 3 taps x 4 channels.
// =====================================================================================================
constexpr int kTW = 16, kTH = 8, kCoT = 64;

template <typename T>
struct ConvArgs {
  const T* in;          // NHWC [nb][h][w][cin] (ignored when PLANAR)
  ImageBatch img;       // PLANAR source
  const float* wpack;   // [9][cin][cout]
  const float* bias;    // forward only
  T* out;
  const T* mask_act;    // backward: multiply by (mask_act > 0) when non-null
  const T* inj;         // backward: add when non-null
  int h, w, cin, cout;
  int forward;
};

template <typename T, int CI, bool PLANAR>
__global__ void __launch_bounds__(256) conv3x3_kernel(ConvArgs<T> a) {
  ST_PDL_ENTRY();
  __shared__ __align__(16) float in_s[CI][kTH + 2][kTW + 2];
  __shared__ __align__(16) float w_s[9][CI][kCoT];

  const int tid = threadIdx.x;
  const int tiles_x = (a.w + kTW - 1) / kTW;
  const int oy = (blockIdx.x / tiles_x) * kTH, ox = (blockIdx.x % tiles_x) * kTW;
  const int co0 = blockIdx.y * kCoT;
  const int bt = blockIdx.z;                               // tile of the batch
  const size_t in_b = (size_t)bt * a.h * a.w * a.cin, out_b = (size_t)bt * a.h * a.w * a.cout;
  const int tx = tid & 15, ty = tid >> 4;
  const int row = ty >> 1, x0 = (ty & 1) * 8;

  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;

  const int hr = tid / (kTW + 2), hc = tid % (kTW + 2);   // halo pixel owned by this thread
  const int gy = oy + hr - 1, gx = ox + hc - 1;
  const bool loader = tid < (kTH + 2) * (kTW + 2);
  const bool inside = loader && gy >= 0 && gy < a.h && gx >= 0 && gx < a.w;

  for (int ci0 = 0; ci0 < a.cin; ci0 += CI) {
    if (loader) {
      if constexpr (PLANAR) {
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          float v = 0.f;
          if (inside) {
            const int cy = wrap(a.img.oy[bt] + gy, a.img.H), cx = wrap(a.img.ox[bt] + gx, a.img.W);
            v = a.img.base[((size_t)ci * a.img.H + cy) * a.img.W + cx];
          }
          in_s[ci][hr][hc] = v;
        }
      } else {
        float v[CI];
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) v[ci] = 0.f;
        if (inside) {
          const T* p = a.in + in_b + ((size_t)gy * a.w + gx) * a.cin + ci0;
#pragma unroll
          for (int q = 0; q < CI / 4; ++q) {
            float4 f = Store<T>::ld4(p + 4 * q);
            v[4 * q] = f.x, v[4 * q + 1] = f.y, v[4 * q + 2] = f.z, v[4 * q + 3] = f.w;
          }
        }
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) in_s[ci][hr][hc] = v[ci];
      }
    }
    for (int i = tid; i < 9 * CI * (kCoT / 4); i += 256) {
      const int q = i & 15, tc = i >> 4;
      const int tap = tc / CI, ci = tc % CI;
      const float4 wv = *reinterpret_cast<const float4*>(
          a.wpack + ((size_t)tap * a.cin + ci0 + ci) * a.cout + co0 + q * 4);
      *reinterpret_cast<float4*>(&w_s[tap][ci][q * 4]) = wv;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float iv[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) iv[j] = in_s[ci][row + ky][x0 + j];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 wv = *reinterpret_cast<const float4*>(&w_s[ky * 3 + kx][ci][tx * 4]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[j][0] = fmaf(iv[j + kx], wv.x, acc[j][0]);
            acc[j][1] = fmaf(iv[j + kx], wv.y, acc[j][1]);
            acc[j][2] = fmaf(iv[j + kx], wv.z, acc[j][2]);
            acc[j][3] = fmaf(iv[j + kx], wv.w, acc[j][3]);
          }
        }
      }
    }
    __syncthreads();
  }

  const int y = oy + row;
  if (y >= a.h) return;
  const int co = co0 + tx * 4;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.forward) b = *reinterpret_cast<const float4*>(a.bias + co);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int x = ox + x0 + j;
    if (x >= a.w) break;
    const size_t o = out_b + ((size_t)y * a.w + x) * a.cout + co;
    float4 v = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    if (a.forward) {
      v.x = fmaxf(v.x + b.x, 0.f), v.y = fmaxf(v.y + b.y, 0.f);
      v.z = fmaxf(v.z + b.z, 0.f), v.w = fmaxf(v.w + b.w, 0.f);
    } else {
      if (a.mask_act) {
        const float4 m = Store<T>::ld4(a.mask_act + o);
        v.x = m.x > 0.f ? v.x : 0.f, v.y = m.y > 0.f ? v.y : 0.f;
        v.z = m.z > 0.f ? v.z : 0.f, v.w = m.w > 0.f ? v.w : 0.f;
      }
      if (a.inj) {
        const float4 g = Store<T>::ld4(a.inj + o);
        v.x += g.x, v.y += g.y, v.z += g.z, v.w += g.w;
      }
    }
    Store<T>::st4(a.out + o, v);
  }
}

template <typename T>
int conv3x3_simt(const T* in, const float* wpack, const float* bias, T* out, int nb, int h, int w,
                 int cin, int cout, bool forward, const T* mask_act, const T* inj, cudaStream_t s) {
  ST_REQUIRE(cin % 8 == 0 && cout % kCoT == 0, "conv3x3: cin must be a multiple of 8, cout of 64");
  ConvArgs<T> a{};
  a.in = in, a.wpack = wpack, a.bias = bias, a.out = out, a.mask_act = mask_act, a.inj = inj;
  a.h = h, a.w = w, a.cin = cin, a.cout = cout, a.forward = forward ? 1 : 0;
  dim3 grid(cdiv(h, kTH) * cdiv(w, kTW), cout / kCoT, nb);
  auto k = conv3x3_kernel<T, 8, false>;
  TimerScope ts(s, kTimeConvSimt, 18.0 * cin * cout * h * w * nb);
  ST_LAUNCH(k, grid, 256, 0, s, a);
  return ST_OK;
}

template <typename T>
int conv_first_fwd(const ImageBatch& img, int h, int w, const float* wpack, const float* bias,
                   T* out, int cout, cudaStream_t s) {
  ST_REQUIRE(cout % kCoT == 0, "first conv: cout must be a multiple of 64");
  ConvArgs<T> a{};
  a.img = img, a.wpack = wpack, a.bias = bias, a.out = out;
  a.h = h, a.w = w, a.cin = 3, a.cout = cout, a.forward = 1;
  dim3 grid(cdiv(h, kTH) * cdiv(w, kTW), cout / kCoT, img.nb);
  auto k = conv3x3_kernel<T, 3, true>;
  TimerScope ts(s, kTimeConvSimt, 18.0 * 3 * cout * h * w * img.nb);
  ST_LAUNCH(k, grid, 256, 0, s, a);
  return ST_OK;
}

// =====================================================================================================
// Backward of the first convolution: d(data)[ci][y][x] = sum_{tap,co} dz[p - tap][co] * W[co][ci][tap].
// wpack is the backward pack [9][cz][4] (tap already flipped, 3 input channels padded to 4).
// Thread: 4 consecutive pixels of a row x 3 image channels; co swept in chunks of 16.
// =====================================================================================================
constexpr int kLW = 32, kLH = 8, kLC = 16, kLPad = 20;

template <typename T>
__global__ void __launch_bounds__(64) conv_last_bwd_kernel(const T* __restrict__ dz, int h, int w,
                                                           int cz, const float* __restrict__ wpack,
                                                           float* __restrict__ grad,
                                                           long batch_stride, long plane,
                                                           long rstride) {
  ST_PDL_ENTRY();
  dz += (size_t)blockIdx.y * h * w * cz;
  grad += (size_t)blockIdx.y * batch_stride;
  __shared__ __align__(16) float in_s[kLH + 2][kLW + 2][kLPad];
  __shared__ __align__(16) float w_s[9][kLC][4];
  const int tid = threadIdx.x;
  const int tiles_x = (w + kLW - 1) / kLW;
  const int oy = (blockIdx.x / tiles_x) * kLH, ox = (blockIdx.x % tiles_x) * kLW;
  const int row = tid >> 3, x0 = (tid & 7) * 4;
  float acc[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;

  for (int c0 = 0; c0 < cz; c0 += kLC) {
    for (int i = tid; i < (kLH + 2) * (kLW + 2) * (kLC / 4); i += 64) {
      const int q = i & 3, pix = i >> 2;
      const int r = pix / (kLW + 2), c = pix % (kLW + 2);
      const int gy = oy + r - 1, gx = ox + c - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < h && gx >= 0 && gx < w)
        v = Store<T>::ld4(dz + ((size_t)gy * w + gx) * cz + c0 + q * 4);
      *reinterpret_cast<float4*>(&in_s[r][c][q * 4]) = v;
    }
    for (int i = tid; i < 9 * kLC; i += 64) {
      const int tap = i / kLC, co = i % kLC;
      *reinterpret_cast<float4*>(&w_s[tap][co][0]) =
          *reinterpret_cast<const float4*>(wpack + ((size_t)tap * cz + c0 + co) * 4);
    }
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int q = 0; q < kLC / 4; ++q) {
        float4 iv[6];
#pragma unroll
        for (int j = 0; j < 6; ++j)
          iv[j] = *reinterpret_cast<const float4*>(&in_s[row + ky][x0 + j][q * 4]);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(&w_s[ky * 3 + kx][q * 4 + 0][0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&w_s[ky * 3 + kx][q * 4 + 1][0]);
          const float4 w2 = *reinterpret_cast<const float4*>(&w_s[ky * 3 + kx][q * 4 + 2][0]);
          const float4 w3 = *reinterpret_cast<const float4*>(&w_s[ky * 3 + kx][q * 4 + 3][0]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = iv[j + kx];
            acc[j][0] += v.x * w0.x + v.y * w1.x + v.z * w2.x + v.w * w3.x;
            acc[j][1] += v.x * w0.y + v.y * w1.y + v.z * w2.y + v.w * w3.y;
            acc[j][2] += v.x * w0.z + v.y * w1.z + v.z * w2.z + v.w * w3.z;
          }
        }
      }
    }
    __syncthreads();
  }
  const int y = oy + row;
  if (y >= h) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = ox + x0 + j;
    if (x >= w) break;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) grad[ci * plane + (long)y * rstride + x] = acc[j][ci];
  }
}

template <typename T>
int conv_last_bwd(const T* dz, int nb, int h, int w, int cz, const float* wpack, float* grad,
                  long batch_stride, long plane_stride, long row_stride, cudaStream_t s) {
  ST_REQUIRE(cz % kLC == 0, "last conv backward: channel count must be a multiple of 16");
  auto k = conv_last_bwd_kernel<T>;
  TimerScope ts(s, kTimeConvSimt, 18.0 * 3 * cz * h * w * nb);
  ST_LAUNCH(k, dim3(cdiv(h, kLH) * cdiv(w, kLW), nb), 64, 0, s, dz, h, w, cz, wpack, grad,
            batch_stride, plane_stride, row_stride);
  return ST_OK;
}

// =====================================================================================================
// 2x2 / stride 2 pooling, ceil mode (Caffe PoolingLayer).  Thread: one output pixel x 4 channels.
// MAX: accumulator starts at -FLT_MAX, strict '>' in scan order => the first maximum wins; the
// backward pass recomputes that argmax from the stored input instead of saving a mask.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const T* __restrict__ in_all, T* __restrict__ out, int nb, int h, int w, int c,
                int ho, int wo, int is_max) {
  ST_PDL_ENTRY();
  // 32-bit index arithmetic throughout: 64-bit div/mod made these kernels instruction-bound
  const unsigned c8 = c >> 3, uwo = wo, uho = ho;
  const unsigned total = (unsigned)nb * ho * wo * c8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i % c8;
    const unsigned p = i / c8;                                // pooled pixel over the whole batch
    const unsigned row = p / uwo;
    const int x = (int)(p - row * uwo), y = (int)(row % uho);
    const T* in = in_all + (size_t)(row / uho) * ((size_t)h * w * c) + q * 8;
    F8 v[4];
    bool ok[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {                             // all four loads in flight together
      const int yy = 2 * y + (d >> 1), xx = 2 * x + (d & 1);
      ok[d] = yy < h && xx < w;
      if (ok[d]) v[d] = ld8(in + ((size_t)yy * w + xx) * c);
    }
    F8 m;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) m.v[k] = is_max ? -FLT_MAX : 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (!ok[d]) continue;
      ++cnt;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (is_max)
          m.v[k] = v[d].v[k] > m.v[k] ? v[d].v[k] : m.v[k];   // strict '>': first maximum wins
        else
          m.v[k] += v[d].v[k];
      }
    }
    if (!is_max) {
      const float inv = (float)cnt;
#pragma unroll
      for (int k = 0; k < 8; ++k) m.v[k] /= inv;
    }
    st8(out + (size_t)p * c + q * 8, m);
  }
}

template <typename TA, typename T>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const T* __restrict__ d_out, const TA* __restrict__ in_all, T* __restrict__ d_in_all,
                int nb, int h, int w, int c, int ho, int wo, int is_max, int apply_mask,
                const T* __restrict__ inj_all, const float* __restrict__ inj_scale) {
  ST_PDL_ENTRY();
  const unsigned c4 = c >> 2, uwo = wo, uho = ho;
  const unsigned total = (unsigned)nb * ho * wo * c4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i % c4;
    const unsigned p = i / c4;
    const unsigned row = p / uwo;
    const int x = (int)(p - row * uwo), y = (int)(row % uho);
    const unsigned bt = row / uho;                            // tile of the batch
    const size_t boff = (size_t)bt * ((size_t)h * w * c) + q * 4;
    const TA* in = in_all + boff;
    T* d_in = d_in_all + boff;
    const T* inj = inj_all ? inj_all + boff : nullptr;
    const float isc = inj_scale ? inj_scale[bt] : 1.f;
    const float4 g = Store<T>::ld4(d_out + (size_t)p * c + q * 4);
    float4 v[4], e[4];
    bool ok[4];
    int cnt = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int yy = 2 * y + (d >> 1), xx = 2 * x + (d & 1);
      ok[d] = yy < h && xx < w;
      const size_t o = ((size_t)yy * w + xx) * c;
      v[d] = ok[d] ? Store<TA>::ld4(in + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      e[d] = (ok[d] && inj) ? Store<T>::ld4(inj + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      cnt += ok[d];
    }
    int ax = 0, ay = 0, az = 0, aw = 0;
    if (is_max) {
      float4 m = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (!ok[d]) continue;
        if (v[d].x > m.x) m.x = v[d].x, ax = d;
        if (v[d].y > m.y) m.y = v[d].y, ay = d;
        if (v[d].z > m.z) m.z = v[d].z, az = d;
        if (v[d].w > m.w) m.w = v[d].w, aw = d;
      }
    }
    const float inv = (float)cnt;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (!ok[d]) continue;
      const int yy = 2 * y + (d >> 1), xx = 2 * x + (d & 1);
      float4 r;
      if (is_max) {
        r.x = ax == d ? g.x : 0.f, r.y = ay == d ? g.y : 0.f;
        r.z = az == d ? g.z : 0.f, r.w = aw == d ? g.w : 0.f;
      } else {
        r.x = g.x / inv, r.y = g.y / inv, r.z = g.z / inv, r.w = g.w / inv;
      }
      if (apply_mask) {
        r.x = v[d].x > 0.f ? r.x : 0.f, r.y = v[d].y > 0.f ? r.y : 0.f;
        r.z = v[d].z > 0.f ? r.z : 0.f, r.w = v[d].w > 0.f ? r.w : 0.f;
      }
      r.x += isc * e[d].x, r.y += isc * e[d].y, r.z += isc * e[d].z, r.w += isc * e[d].w;
      Store<T>::st4(d_in + ((size_t)yy * w + xx) * c, r);
    }
  }
}

// Backward of a pooling layer whose forward ran fused into the convolution (conv_tc2.cu): the input
// activations are never read, one mask byte per pooled element says where the gradient goes
// (max: bits 0-1 arg-max position, bit 2 maximum > 0) or which inputs pass the ReLU (ave: bit d).
template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_mask_kernel(const T* __restrict__ d_out, const uint8_t* __restrict__ mask,
                     T* __restrict__ d_in_all, int nb, int h, int w, int c, int ho, int wo, int is_max,
                     const T* __restrict__ inj_all, const float* __restrict__ inj_scale) {
  ST_PDL_ENTRY();
  const unsigned c8 = c >> 3, uwo = wo, uho = ho;
  const unsigned total = (unsigned)nb * ho * wo * c8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i % c8, p = i / c8;
    const unsigned row = p / uwo;
    const int x = (int)(p - row * uwo), y = (int)(row % uho);
    const unsigned bt = row / uho;
    const size_t boff = (size_t)bt * ((size_t)h * w * c) + q * 8;
    T* d_in = d_in_all + boff;
    const T* inj = inj_all ? inj_all + boff : nullptr;
    const float isc = inj_scale ? inj_scale[bt] : 1.f;
    const F8 g = ld8(d_out + (size_t)p * c + q * 8);
    const uint2 mk = *reinterpret_cast<const uint2*>(mask + (size_t)p * c + q * 8);
    int cnt = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) cnt += (2 * y + (d >> 1) < h && 2 * x + (d & 1) < w) ? 1 : 0;
    const float inv = (float)cnt;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int yy = 2 * y + (d >> 1), xx = 2 * x + (d & 1);
      if (yy >= h || xx >= w) continue;
      const size_t o = ((size_t)yy * w + xx) * c;
      F8 r;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t code = ((k < 4 ? mk.x : mk.y) >> (8 * (k & 3))) & 0xFFu;
        if (is_max)
          r.v[k] = ((code & 3u) == (uint32_t)d && (code & 4u)) ? g.v[k] : 0.f;
        else
          r.v[k] = ((code >> d) & 1u) ? g.v[k] / inv : 0.f;
      }
      if (inj) {
        const F8 e = ld8(inj + o);
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] += isc * e.v[k];
      }
      st8(d_in + o, r);
    }
  }
}

static inline int ew_grid(size_t work_items, int block) {
  size_t b = (work_items + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
int pool_fwd(const T* in, T* out, int nb, int h, int w, int c, bool is_max, cudaStream_t s) {
  ST_REQUIRE(c % 8 == 0, "pool: channels must be a multiple of 8");
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  ST_REQUIRE((size_t)nb * ho * wo * c < ((size_t)1 << 31), "pool: batch too large for 32-bit indexing");
  auto k = pool_fwd_kernel<T>;
  TimerScope ts(s, kTimePool, (double)sizeof(T) * c * nb * ((double)h * w + (double)ho * wo));
  ST_LAUNCH(k, ew_grid((size_t)nb * ho * wo * (c / 8), 256), 256, 0, s, in, out, nb, h, w, c, ho,
            wo, is_max ? 1 : 0);
  return ST_OK;
}

template <typename TA, typename T>
int pool_bwd(const T* d_out, const TA* in, T* d_in, int nb, int h, int w, int c, bool is_max,
             bool apply_mask, const T* inj, const float* inj_scale, cudaStream_t s) {
  ST_REQUIRE(c % 8 == 0, "pool: channels must be a multiple of 8");
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  ST_REQUIRE((size_t)nb * ho * wo * c < ((size_t)1 << 31), "pool: batch too large for 32-bit indexing");
  auto k = pool_bwd_kernel<TA, T>;
  TimerScope ts(s, kTimePool, (double)sizeof(T) * c * nb *
                                  ((double)h * w * (2 + (inj ? 1 : 0)) + (double)ho * wo));
  ST_LAUNCH(k, ew_grid((size_t)nb * ho * wo * (c / 4), 256), 256, 0, s, d_out, in, d_in, nb, h, w,
            c, ho, wo, is_max ? 1 : 0, apply_mask ? 1 : 0, inj, inj_scale);
  return ST_OK;
}

template <typename T>
int pool_bwd_mask(const T* d_out, const uint8_t* mask, T* d_in, int nb, int h, int w, int c,
                  bool is_max, const T* inj, const float* inj_scale, cudaStream_t s) {
  ST_REQUIRE(c % 8 == 0, "pool: channels must be a multiple of 8");
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  ST_REQUIRE((size_t)nb * ho * wo * c < ((size_t)1 << 31), "pool: batch too large for 32-bit indexing");
  auto k = pool_bwd_mask_kernel<T>;
  TimerScope ts(s, kTimePool, (double)c * nb * ((double)sizeof(T) * h * w * (1 + (inj ? 1 : 0)) +
                                                (double)(sizeof(T) + 1) * ho * wo));
  ST_LAUNCH(k, ew_grid((size_t)nb * ho * wo * (c / 8), 256), 256, 0, s, d_out, mask, d_in, nb, h, w,
            c, ho, wo, is_max ? 1 : 0, inj, inj_scale);
  return ST_OK;
}

// =====================================================================================================
// Gram matrix G = F^T F / (C*HW) (num_utils.py:143-147), split over pixels.
//   grid.x enumerates 64x64 output blocks with bi >= bj, grid.y the pixel split; partials are summed
//   in split order by gram_finalize (deterministic, no float atomics).
// =====================================================================================================
constexpr int kGP = 32, kGS = 68;

template <typename T, bool CM>
__global__ void __launch_bounds__(256) gram_partial_kernel(const T* __restrict__ f, int hw, int c,
                                                           int px_per_split,
                                                           float* __restrict__ part) {
  ST_PDL_ENTRY();
  __shared__ __align__(16) float a_s[kGP][kGS];
  __shared__ __align__(16) float b_s[kGP][kGS];
  int bi = 0, rem = blockIdx.x;
  while (rem > bi) rem -= ++bi;                 // blockIdx.x = bi*(bi+1)/2 + bj
  const int bj = rem;
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  const int p_begin = blockIdx.y * px_per_split;
  const int p_end = min(hw, p_begin + px_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int p0 = p_begin; p0 < p_end; p0 += kGP) {
    if constexpr (CM) {
      for (int i = tid; i < kGP * 64; i += 256) {
        const int ch = i / kGP, p = i % kGP;
        const bool ok = p0 + p < p_end;
        a_s[p][ch] = ok ? Store<T>::ld(f + (size_t)(bi * 64 + ch) * hw + p0 + p) : 0.f;
        b_s[p][ch] = ok ? Store<T>::ld(f + (size_t)(bj * 64 + ch) * hw + p0 + p) : 0.f;
      }
    } else {
      for (int i = tid; i < kGP * 16; i += 256) {
        const int p = i >> 4, q = i & 15;
        const bool ok = p0 + p < p_end;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const T* src = f + (size_t)(p0 + p) * c;
        *reinterpret_cast<float4*>(&a_s[p][q * 4]) = ok ? Store<T>::ld4(src + bi * 64 + q * 4) : z;
        *reinterpret_cast<float4*>(&b_s[p][q * 4]) = ok ? Store<T>::ld4(src + bj * 64 + q * 4) : z;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < kGP; ++p) {
      const float4 av = *reinterpret_cast<const float4*>(&a_s[p][ti * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&b_s[p][tj * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = part + (size_t)blockIdx.y * c * c;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + (size_t)(bi * 64 + ti * 4 + i) * c + bj * 64 + tj * 4) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

__global__ void gram_finalize_kernel(const float* __restrict__ part, int nsplit, int c,
                                     double scale, float* __restrict__ gram) {
  ST_PDL_ENTRY();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= c * c) return;
  const int i = idx / c, j = idx % c;
  const int si = (i >> 6) >= (j >> 6) ? i : j, sj = (i >> 6) >= (j >> 6) ? j : i;
  double sum = 0.0;
  for (int s = 0; s < nsplit; ++s) sum += (double)part[((size_t)s * c + si) * c + sj];
  gram[idx] = (float)(sum * scale);
}

template <typename T>
int gram_full(const T* f, int hw, int c, bool channel_major, float* gram, float* part,
              size_t part_floats, int sm_count, cudaStream_t s) {
  ST_REQUIRE(c % 64 == 0, "gram: channels must be a multiple of 64");
  const int nb = c / 64, npairs = nb * (nb + 1) / 2;
  int nsplit = cdiv(4L * sm_count, npairs);
  const int max_split = (int)(part_floats / ((size_t)c * c));
  ST_REQUIRE(max_split >= 1, "gram: partial buffer too small");
  nsplit = nsplit < 1 ? 1 : nsplit;
  nsplit = nsplit > max_split ? max_split : nsplit;
  int pps = cdiv(cdiv(hw, nsplit), kGP) * kGP;
  nsplit = cdiv(hw, pps);
  dim3 grid(npairs, nsplit);
  TimerScope ts(s, kTimeGram, 2.0 * c * c * hw);
  if (channel_major) {
    auto k = gram_partial_kernel<T, true>;
    ST_LAUNCH(k, grid, 256, 0, s, f, hw, c, pps, part);
  } else {
    auto k = gram_partial_kernel<T, false>;
    ST_LAUNCH(k, grid, 256, 0, s, f, hw, c, pps, part);
  }
  ST_LAUNCH(gram_finalize_kernel, cdiv((long)c * c, 256), 256, 0, s, part, nsplit, c,
            1.0 / ((double)c * hw), gram);
  return ST_OK;
}

// delta[b] = G[b] - G_style (symmetric); tile_loss[b] += w * 0.5 * sum_{j<=i} delta^2
// (style_transfer.py:587,591).  blockIdx.y = tile of the batch.  max_bits[b] (optional) collects
// max |delta| as float bits (atomicMax on non-negative floats is order independent).
__global__ void gram_delta_kernel(const float* __restrict__ gram, const float* __restrict__ target,
                                  float* __restrict__ delta, unsigned* max_bits, int c, double w,
                                  double* tile_loss, int loss_stride, ReduceScratch rs) {
  ST_PDL_ENTRY();
  const size_t off = (size_t)blockIdx.y * c * c;
  double v[1] = {0.0};
  float mx = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < c * c;
       idx += gridDim.x * blockDim.x) {
    const float d = gram[off + idx] - target[idx];
    delta[off + idx] = d;
    mx = fmaxf(mx, fabsf(d));
    if (idx % c <= idx / c) v[0] += (double)d * d;
  }
  if (max_bits != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits + blockIdx.y, __float_as_uint(mx));
  }
  if (grid_reduce<1>(v, rs.partials + (size_t)blockIdx.y * gridDim.x, rs.counter + blockIdx.y))
    tile_loss[(size_t)blockIdx.y * loss_stride] += w * 0.5 * v[0];
}

// The 16-bit copy of delta that is the B operand of the tensor-core style GEMM.  bf16: plain
// rounding, eps_eff[b] = EPS.  fp16 (half): delta is first multiplied by sigma_b = 2^-ceil(log2 max
// |delta_b|), an exact power-of-two scaling that keeps it inside the fp16 range; S then comes out
// scaled by sigma_b too, which normalize() cancels except in its EPS term, hence eps_eff[b] =
// EPS * sigma_b (w / (mean|S| + EPS) * S  ==  w / (mean|sigma S| + EPS sigma) * sigma S).
__global__ void __launch_bounds__(256)
delta_pack_kernel(const float* __restrict__ delta, uint16_t* __restrict__ out, unsigned* max_bits,
                  float* eps_eff, int c, int half, const double* __restrict__ loss_part, int n_part,
                  double w, double* tile_loss, int loss_stride) {
  ST_PDL_ENTRY();
  const size_t off = (size_t)blockIdx.y * c * c;
  if (loss_part != nullptr && blockIdx.x == 0) {
    // the style loss of this tile from the block partials of gram_tc_finish: thread t adds the
    // strided subsequence t, t + 256, ..., then a fixed tree -- independent of the launch
    __shared__ double sh[256];
    const double* p = loss_part + (size_t)blockIdx.y * n_part;
    double x = 0.0;
    for (int i = threadIdx.x; i < n_part; i += 256) x += p[i];
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) tile_loss[(size_t)blockIdx.y * loss_stride] += w * 0.5 * sh[0];
  }
  float sigma = 1.f;
  if (half) {
    const float mx = __uint_as_float(max_bits[blockIdx.y]);
    if (mx > 0.f && isfinite(mx)) {
      int e;
      frexpf(mx, &e);                       // mx = m * 2^e, 0.5 <= m < 1
      sigma = ldexpf(1.f, -e);              // max |delta| * sigma in [0.5, 1)
    }
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < c * c;
       idx += gridDim.x * blockDim.x) {
    const float d = delta[off + idx] * sigma;
    if (half) {
      const __half hv = __float2half_rn(d);
      out[off + idx] = *reinterpret_cast<const uint16_t*>(&hv);
    } else {
      const __nv_bfloat16 bv = __float2bfloat16_rn(d);
      out[off + idx] = *reinterpret_cast<const uint16_t*>(&bv);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) eps_eff[blockIdx.y] = kEps * sigma;
}

int gram_delta(const float* gram, const float* target, float* delta, void* delta_16, bool half,
               unsigned* max_bits, float* eps_eff, int c, int nb, double w, double* tile_loss,
               int loss_stride, ReduceScratch rs, cudaStream_t s, bool track_max) {
  const bool track = (delta_16 != nullptr && half) || (track_max && max_bits != nullptr);
  if (track) ST_CUDA(cudaMemsetAsync(max_bits, 0, nb * sizeof(unsigned), s));
  ST_LAUNCH(gram_delta_kernel, dim3(min(cdiv((long)c * c, 256), 256), nb), 256, 0, s, gram, target,
            delta, track ? max_bits : nullptr, c, w, tile_loss, loss_stride, rs);
  if (delta_16 != nullptr)
    return delta_pack(delta, delta_16, half, max_bits, eps_eff, c, nb, nullptr, 0, 0.0, nullptr, 0, s);
  return ST_OK;
}

int delta_pack(const float* delta, void* delta_16, bool half, unsigned* max_bits, float* eps_eff,
               int c, int nb, const double* loss_part, int n_part, double w, double* tile_loss,
               int loss_stride, cudaStream_t s) {
  ST_LAUNCH(delta_pack_kernel, dim3(min(cdiv((long)c * c, 256), 64), nb), 256, 0, s, delta,
            static_cast<uint16_t*>(delta_16), max_bits, eps_eff, c, half ? 1 : 0, loss_part, n_part, w,
            tile_loss, loss_stride);
  return ST_OK;
}

// out[b * out_stride] = sum of partials[b * n .. b * n + n).  One block per tile: thread t adds the
// strided subsequence t, t+256, ... and the 256 sums are combined by a fixed tree, so the result does
// not depend on the launch.  Optionally also scale[b] = w / (sum / count + EPS): the factor
// normalize() + the layer weight apply to the style gradient (num_utils.py:85-87, :591-593).
__global__ void __launch_bounds__(256)
sum_partials_kernel(const double* __restrict__ partials, int n, double* out, int out_stride,
                    float* scale, float w, double count, const float* __restrict__ eps_eff) {
  ST_PDL_ENTRY();
  __shared__ double sh[256];
  const double* p = partials + (size_t)blockIdx.x * n;
  double x = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) x += p[i];
  sh[threadIdx.x] = x;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[(size_t)blockIdx.x * out_stride] = sh[0];
    const float eps = eps_eff ? eps_eff[blockIdx.x] : kEps;
    if (scale != nullptr) scale[blockIdx.x] = w * (1.f / ((float)(sh[0] / count) + eps));
  }
}
int sum_partials(const double* partials, int n, int nb, double* out, int out_stride, float* scale,
                 float w, double count, const float* eps_eff, cudaStream_t s) {
  ST_LAUNCH(sum_partials_kernel, nb, 256, 0, s, partials, n, out, out_stride, scale, w, count,
            eps_eff);
  return ST_OK;
}

// *loss_accum += tile_loss[0] + tile_loss[stride] + ... (tile order), then clears the slots.
__global__ void loss_finalize_kernel(double* tile_loss, int stride, int nb, double* loss_accum) {
  ST_PDL_ENTRY();
  double x = 0.0;
  for (int b = 0; b < nb; ++b) x += tile_loss[(size_t)b * stride], tile_loss[(size_t)b * stride] = 0.0;
  *loss_accum += x;
}
int loss_finalize(double* tile_loss, int stride, int nb, double* loss_accum, cudaStream_t s) {
  ST_LAUNCH(loss_finalize_kernel, 1, 1, 0, s, tile_loss, stride, nb, loss_accum);
  return ST_OK;
}

__global__ void symmetrize_kernel(const float* __restrict__ src, float* __restrict__ dst, int c,
                                  int to_lower) {
  ST_PDL_ENTRY();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= c * c) return;
  const int i = idx / c, j = idx % c;
  if (to_lower)
    dst[idx] = j <= i ? src[idx] : 0.f;
  else
    dst[idx] = j <= i ? src[idx] : src[j * c + i];
}
int symmetrize_lower(const float* lower, float* full, int c, cudaStream_t s) {
  ST_LAUNCH(symmetrize_kernel, cdiv((long)c * c, 256), 256, 0, s, lower, full, c, 0);
  return ST_OK;
}
int extract_lower(const float* full, float* lower, int c, cudaStream_t s) {
  ST_LAUNCH(symmetrize_kernel, cdiv((long)c * c, 256), 256, 0, s, full, lower, c, 1);
  return ST_OK;
}

// =====================================================================================================
// Style gradient S = F * sym(dG)  ([hw][c] x [c][c], num_utils.py:60-66 / style_transfer.py:589) with
// the sum |S| that normalize() needs (num_utils.py:85-87) reduced in the same pass.
// =====================================================================================================
constexpr int kSK = 16, kSS = 68;

template <typename T>
__global__ void __launch_bounds__(256) style_grad_kernel(const T* __restrict__ f,
                                                         const float* __restrict__ delta,
                                                         T* __restrict__ s_out, int hw, int c,
                                                         double* sum_abs, ReduceScratch rs) {
  ST_PDL_ENTRY();
  __shared__ __align__(16) float f_s[kSK][kSS];    // [k][pixel]
  __shared__ __align__(16) float d_s[kSK][kSS];    // [k][out channel]
  const int tid = threadIdx.x, tp = tid >> 4, tq = tid & 15;
  const int nbx = gridDim.x / (c / 64);
  const int p0 = (blockIdx.x % nbx) * 64, co0 = (blockIdx.x / nbx) * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < c; k0 += kSK) {
    {
      const int px = tid >> 2, g = tid & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p0 + px < hw) v = Store<T>::ld4(f + (size_t)(p0 + px) * c + k0 + g * 4);
      f_s[g * 4 + 0][px] = v.x, f_s[g * 4 + 1][px] = v.y;
      f_s[g * 4 + 2][px] = v.z, f_s[g * 4 + 3][px] = v.w;
      const int k = tid >> 4, q = tid & 15;
      *reinterpret_cast<float4*>(&d_s[k][q * 4]) =
          *reinterpret_cast<const float4*>(delta + (size_t)(k0 + k) * c + co0 + q * 4);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&f_s[k][tp * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&d_s[k][tq * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
  double v[1] = {0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + tp * 4 + i;
    if (p < hw) {
      Store<T>::st4(s_out + (size_t)p * c + co0 + tq * 4,
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      v[0] += (double)(fabsf(acc[i][0]) + fabsf(acc[i][1]) + fabsf(acc[i][2]) + fabsf(acc[i][3]));
    }
  }
  if (grid_reduce<1>(v, rs.partials, rs.counter)) *sum_abs = v[0];
}

template <typename T>
int style_grad(const T* f, const float* delta, T* s_out, int hw, int c, double* sum_abs,
               ReduceScratch rs, cudaStream_t s) {
  ST_REQUIRE(c % 64 == 0, "style_grad: channels must be a multiple of 64");
  const long blocks = (long)cdiv(hw, 64) * (c / 64);
  ST_REQUIRE(blocks <= kMaxReduceBlocks, "style_grad: tile too large for the reduction scratch");
  auto k = style_grad_kernel<T>;
  TimerScope ts(s, kTimeStyleGrad, 2.0 * c * c * hw);
  ST_LAUNCH(k, (int)blocks, 256, 0, s, f, delta, s_out, hw, c, sum_abs, rs);
  return ST_OK;
}

template <typename T>
__global__ void inject_scaled_kernel(T* __restrict__ inj, const T* __restrict__ src, size_t n4,
                                     float w, const double* __restrict__ sum_abs, int stat_stride,
                                     const float* __restrict__ eps_eff, int accumulate) {
  ST_PDL_ENTRY();
  // blockIdx.y = tile of the batch; n4 = float4 groups per tile
  const float eps = eps_eff ? eps_eff[blockIdx.y] : kEps;
  const float coef =
      w * (1.f / ((float)(sum_abs[(size_t)blockIdx.y * stat_stride] / (double)(n4 * 4)) + eps));
  inj += (size_t)blockIdx.y * n4 * 4, src += (size_t)blockIdx.y * n4 * 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 sv = Store<T>::ld4(src + i * 4);
    float4 r = make_float4(coef * sv.x, coef * sv.y, coef * sv.z, coef * sv.w);
    if (accumulate) {
      const float4 o = Store<T>::ld4(inj + i * 4);
      r.x += o.x, r.y += o.y, r.z += o.z, r.w += o.w;
    }
    Store<T>::st4(inj + i * 4, r);
  }
}

template <typename T>
int inject_scaled(T* inj, const T* src, size_t n, int nb, float w, const double* sum_abs,
                  int stat_stride, const float* eps_eff, bool accumulate, cudaStream_t s) {
  ST_REQUIRE(n % 4 == 0, "inject: size must be a multiple of 4");
  auto k = inject_scaled_kernel<T>;
  ST_LAUNCH(k, dim3(ew_grid(n / 4, 256), nb), 256, 0, s, inj, src, n / 4, w, sum_abs, stat_stride,
            eps_eff, accumulate ? 1 : 0);
  return ST_OK;
}

// =====================================================================================================
// Content / Deep-Dream terms (style_transfer.py:575-580, 602-604): c = F - target slice (or F),
// loss 0.5*sum c^2, gradient c / (mean|c| + EPS).  Two passes, c is never stored.
// =====================================================================================================
// One feature row of a tile (wf pixels x c channels, contiguous in NHWC) against the matching row of
// the whole-image target: blocks walk the rows, threads walk the row in groups of 8 channels, so the
// only index arithmetic per 8 elements is one FastDiv (the first version did two 32-bit divisions
// per 4 elements and was instruction-bound at 90 % SM busy / 17 % of the HBM bandwidth,
// profiles/r01_laggards_ncu.md).
struct DiffGeom {
  int hf, wf, c8;            // tile feature map: rows, pixels per row, 8-channel groups per pixel
  int Hf, Wf;                // whole-image target map
  FastDiv div_c8;
};

template <typename T>
__device__ __forceinline__ F8 diff_row8(const T* __restrict__ frow, const float* __restrict__ trow,
                                        unsigned i, const DiffGeom& g, int tx0) {
  F8 v = ld8(frow + (size_t)i * 8);
  if (trow) {
    const unsigned x = g.div_c8.div(i), q = i - x * (unsigned)g.c8;
    int xx = tx0 + (int)x;                       // offsets are reduced on the host to [0, Wf)
    xx = xx >= g.Wf ? xx - g.Wf : xx;
    const F8 t = ld8(trow + ((size_t)xx * g.c8 + q) * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] -= t.v[e];
  }
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
diff_stats_kernel(const T* __restrict__ f, DiffGeom g, const float* __restrict__ tgt,
                  TargetOffsets offs, double* stats, int stat_stride, ReduceScratch rs) {
  ST_PDL_ENTRY();
  const int b = blockIdx.y;
  const unsigned row8 = (unsigned)g.wf * g.c8;
  f += (size_t)b * g.hf * row8 * 8;
  float sq = 0.f, ab = 0.f;
  double v[2] = {0.0, 0.0};
  int cnt = 0;
  for (int y = blockIdx.x; y < g.hf; y += gridDim.x) {
    int yy = offs.ty0[b] + y;
    yy = yy >= g.Hf ? yy - g.Hf : yy;
    const T* frow = f + (size_t)y * row8 * 8;
    const float* trow = tgt ? tgt + (size_t)yy * g.Wf * g.c8 * 8 : nullptr;
    for (unsigned i = threadIdx.x; i < row8; i += blockDim.x) {
      const F8 d = diff_row8(frow, trow, i, g, offs.tx0[b]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sq += d.v[e] * d.v[e], ab += fabsf(d.v[e]);
      if (++cnt == 32) v[0] += sq, v[1] += ab, sq = ab = 0.f, cnt = 0;
    }
  }
  v[0] += sq, v[1] += ab;
  if (grid_reduce<2>(v, rs.partials + (size_t)b * gridDim.x * 2, rs.counter + b))
    stats[(size_t)b * stat_stride] = v[0], stats[(size_t)b * stat_stride + 1] = v[1];
}

static DiffGeom diff_geom(int hf, int wf, int c, int Hf, int Wf) {
  DiffGeom g;
  g.hf = hf, g.wf = wf, g.c8 = c / 8, g.Hf = Hf, g.Wf = Wf, g.div_c8 = FastDiv(c / 8);
  return g;
}

template <typename T>
int diff_stats(const T* f, int nb, int hf, int wf, int c, const float* tgt, int Hf, int Wf,
               const TargetOffsets& offs, double* stats, int stat_stride, ReduceScratch rs,
               cudaStream_t s) {
  ST_REQUIRE(c % 8 == 0, "diff_stats: channels must be a multiple of 8");
  // rows per tile fix the grid: the summation order depends on the layer shape only
  const int gx = min(hf, kMaxReduceBlocks / (nb * 2));
  auto k = diff_stats_kernel<T>;
  TimerScope ts(s, kTimeLoss, (double)nb * hf * wf * c * (sizeof(T) + (tgt ? 4 : 0)));
  ST_LAUNCH(k, dim3(gx, nb), 256, 0, s, f, diff_geom(hf, wf, c, Hf, Wf), tgt, offs, stats,
            stat_stride, rs);
  return ST_OK;
}

// stats[b] = {sum c^2, sum |c|}; inj = (accumulate ? inj : 0) + w / (mean|c| + EPS) * c;
// tile_loss[b] += loss_w * 0.5 * sum c^2
template <typename TA, typename T>
__global__ void __launch_bounds__(256)
diff_inject_kernel(const TA* __restrict__ f, DiffGeom g, const float* __restrict__ tgt,
                   TargetOffsets offs, const double* __restrict__ stats, int stat_stride, float w,
                   double loss_w, double* tile_loss, int loss_stride, T* __restrict__ inj,
                   int accumulate) {
  ST_PDL_ENTRY();
  const int b = blockIdx.y;
  const unsigned row8 = (unsigned)g.wf * g.c8;
  f += (size_t)b * g.hf * row8 * 8, inj += (size_t)b * g.hf * row8 * 8;
  const double* st = stats + (size_t)b * stat_stride;
  const float coef = w * (1.f / ((float)(st[1] / ((double)g.hf * row8 * 8)) + kEps));
  if (blockIdx.x == 0 && threadIdx.x == 0) tile_loss[(size_t)b * loss_stride] += loss_w * 0.5 * st[0];
  for (int y = blockIdx.x; y < g.hf; y += gridDim.x) {
    int yy = offs.ty0[b] + y;
    yy = yy >= g.Hf ? yy - g.Hf : yy;
    const TA* frow = f + (size_t)y * row8 * 8;
    T* irow = inj + (size_t)y * row8 * 8;
    const float* trow = tgt ? tgt + (size_t)yy * g.Wf * g.c8 * 8 : nullptr;
    for (unsigned i = threadIdx.x; i < row8; i += blockDim.x) {
      F8 r = diff_row8(frow, trow, i, g, offs.tx0[b]);
#pragma unroll
      for (int e = 0; e < 8; ++e) r.v[e] *= coef;
      if (accumulate) {
        const F8 o = ld8(irow + (size_t)i * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] += o.v[e];
      }
      st8(irow + (size_t)i * 8, r);
    }
  }
}

template <typename TA, typename T>
int diff_inject(const TA* f, int nb, int hf, int wf, int c, const float* tgt, int Hf, int Wf,
                const TargetOffsets& offs, const double* stats, int stat_stride, float w,
                double loss_w, double* tile_loss, int loss_stride, T* inj, bool accumulate,
                cudaStream_t s) {
  ST_REQUIRE(c % 8 == 0, "diff_inject: channels must be a multiple of 8");
  auto k = diff_inject_kernel<TA, T>;
  TimerScope ts(s, kTimeLoss,
                (double)nb * hf * wf * c * (sizeof(T) * (accumulate ? 3 : 2) + (tgt ? 4 : 0)));
  ST_LAUNCH(k, dim3(hf, nb), 256, 0, s, f, diff_geom(hf, wf, c, Hf, Wf), tgt, offs, stats,
            stat_stride, w, loss_w, tile_loss, loss_stride, inj, accumulate ? 1 : 0);
  return ST_OK;
}

// One thread per 32-channel chunk: bit (e >> 1) + 16 * (e & 1) of the word = act[chunk * 32 + e] > 0.
template <typename T>
__global__ void relu_bits_kernel(const T* __restrict__ act, uint32_t* __restrict__ bits, size_t words) {
  ST_PDL_ENTRY();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words;
       i += (size_t)gridDim.x * blockDim.x) {
    uint32_t w = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const F8 v = ld8(act + i * 32 + q * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (v.v[e] > 0.f) w |= 1u << (((q * 8 + e) >> 1) + 16 * (e & 1));
    }
    bits[i] = w;
  }
}

template <typename T>
int relu_bits_from_act(const T* act, uint32_t* bits, size_t pixels, int c, cudaStream_t s) {
  ST_REQUIRE(c % 32 == 0, "relu_bits: channels must be a multiple of 32");
  const size_t words = pixels * (size_t)(c / 32);
  auto k = relu_bits_kernel<T>;
  ST_LAUNCH(k, ew_grid(words, 256), 256, 0, s, act, bits, words);
  return ST_OK;
}
template int relu_bits_from_act<__nv_bfloat16>(const __nv_bfloat16*, uint32_t*, size_t, int, cudaStream_t);
template int relu_bits_from_act<__half>(const __half*, uint32_t*, size_t, int, cudaStream_t);

// =====================================================================================================
// Layout conversion at the C-ABI boundary: NHWC (internal) <-> NCHW float32 (reference layout).
// =====================================================================================================
template <typename TI, typename TO>
__global__ void transpose_kernel(const TI* __restrict__ in, TO* __restrict__ out, long rows,
                                 long cols, int rows_on_x) {
  ST_PDL_ENTRY();
  // in [rows][cols] -> out [cols][rows]; the long dimension goes on grid.x (2^31 blocks)
  __shared__ float tile[32][33];
  const long r0 = (long)(rows_on_x ? blockIdx.x : blockIdx.y) * 32;
  const long c0 = (long)(rows_on_x ? blockIdx.y : blockIdx.x) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? Store<TI>::ld(in + r * cols + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) Store<TO>::st(out + c * rows + r, tile[threadIdx.x][i]);
  }
}

template <typename T>
int nhwc_to_nchw_f32(const T* in, float* out, int hw, int c, cudaStream_t s) {
  dim3 grid(cdiv(hw, 32), cdiv(c, 32)), block(32, 8);
  auto k = transpose_kernel<T, float>;
  ST_LAUNCH(k, grid, block, 0, s, in, out, (long)hw, (long)c, 1);
  return ST_OK;
}

int nchw_to_nhwc_f32(const float* in, float* out, int hw, int c, cudaStream_t s) {
  dim3 grid(cdiv(hw, 32), cdiv(c, 32)), block(32, 8);
  auto k = transpose_kernel<float, float>;
  ST_LAUNCH(k, grid, block, 0, s, in, out, (long)c, (long)hw, 0);
  return ST_OK;
}

// ---- explicit instantiations ---------------------------------------------------------------------
// TA = activation storage, TG = gradient storage: (float, float) in ST_PREC_FP32, (bf16, bf16) in
// ST_PREC_BF16, (fp16, bf16) in ST_PREC_FP16.
#define ST_INSTANTIATE_ACT(TA)                                                                     \
  template int conv_first_fwd<TA>(const ImageBatch&, int, int, const float*, const float*, TA*,    \
                                  int, cudaStream_t);                                              \
  template int pool_fwd<TA>(const TA*, TA*, int, int, int, int, bool, cudaStream_t);               \
  template int diff_stats<TA>(const TA*, int, int, int, int, const float*, int, int,               \
                              const TargetOffsets&, double*, int, ReduceScratch, cudaStream_t);    \
  template int nhwc_to_nchw_f32<TA>(const TA*, float*, int, int, cudaStream_t);

#define ST_INSTANTIATE_GRAD(TG)                                                                    \
  template int conv_last_bwd<TG>(const TG*, int, int, int, int, const float*, float*, long, long,  \
                                 long, cudaStream_t);                                              \
  template int pool_bwd_mask<TG>(const TG*, const uint8_t*, TG*, int, int, int, int, bool,         \
                                 const TG*, const float*, cudaStream_t);                           \
  template int inject_scaled<TG>(TG*, const TG*, size_t, int, float, const double*, int,           \
                                 const float*, bool, cudaStream_t);

#define ST_INSTANTIATE_PAIR(TA, TG)                                                                \
  template int pool_bwd<TA, TG>(const TG*, const TA*, TG*, int, int, int, int, bool, bool,         \
                                const TG*, const float*, cudaStream_t);                            \
  template int diff_inject<TA, TG>(const TA*, int, int, int, int, const float*, int, int,          \
                                   const TargetOffsets&, const double*, int, float, double,        \
                                   double*, int, TG*, bool, cudaStream_t);

// single-type SIMT kernels of the fp32 parity mode and of the bf16 fallback
#define ST_INSTANTIATE_SIMT(T)                                                                     \
  template int conv3x3_simt<T>(const T*, const float*, const float*, T*, int, int, int, int, int,  \
                               bool, const T*, const T*, cudaStream_t);                            \
  template int gram_full<T>(const T*, int, int, bool, float*, float*, size_t, int, cudaStream_t);  \
  template int style_grad<T>(const T*, const float*, T*, int, int, double*, ReduceScratch,         \
                             cudaStream_t);

ST_INSTANTIATE_ACT(float)
ST_INSTANTIATE_ACT(__nv_bfloat16)
ST_INSTANTIATE_ACT(__half)
ST_INSTANTIATE_GRAD(float)
ST_INSTANTIATE_GRAD(__nv_bfloat16)
ST_INSTANTIATE_PAIR(float, float)
ST_INSTANTIATE_PAIR(__nv_bfloat16, __nv_bfloat16)
ST_INSTANTIATE_PAIR(__half, __nv_bfloat16)
ST_INSTANTIATE_SIMT(float)
ST_INSTANTIATE_SIMT(__nv_bfloat16)

}  // namespace st
