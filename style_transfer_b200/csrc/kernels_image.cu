// Whole-image HBM-bound kernels of libstyle_b200: gradient un-packing (virtual un-roll), the fused
// TV / p-norm / aux regularisers (style_transfer.py:700-736, num_utils.py:74-82,150-162), the Adam
// step with iterate averaging (optimizers.py:26-42) and the BLAS-1 pieces of L-BFGS
// (optimizers.py:74-121).  All float32, coalesced along the image width, reductions in double.
#include <algorithm>
#include <cstdlib>

#include "style_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace st {

static inline int ew_grid(size_t work_items, int block) {
  size_t b = (work_items + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// -----------------------------------------------------------------------------------------------------
// Exchange buffer of one rank: [tiles_per_rank][3][thmax][twmax] floats, padded to a multiple of four,
// followed by a four-float tail whose first eight bytes are the rank's loss (a double).  The
// all-gather result is `world` such chunks; one collective carries gradients and losses.
// -----------------------------------------------------------------------------------------------------
size_t packed_rank_stride(int tiles_per_rank, int thmax, int twmax) {
  const size_t tiles = (size_t)tiles_per_rank * 3 * thmax * twmax;
  return (tiles + 3) / 4 * 4 + 4;
}
__device__ __forceinline__ double packed_loss_sum(const float* packed, int world, size_t rank_stride) {
  double sum = 0.0;
  for (int r = 0; r < world; ++r)
    sum += *reinterpret_cast<const double*>(packed + (size_t)(r + 1) * rank_stride - 4);
  return sum;
}

// -----------------------------------------------------------------------------------------------------
// grad[c][y][x] (un-rolled frame) = packed tile gradient at rolled position ((y+ry) mod H, (x+rx) mod W)
// -----------------------------------------------------------------------------------------------------
__global__ void unpack_grad_kernel(const float* __restrict__ packed, int H, int W, int roll_y,
                                   int roll_x, int nty, int ntx, int th, int tw, int thmax,
                                   int twmax, int world, size_t rank_stride,
                                   float* __restrict__ grad, double* loss_accum) {
  ST_PDL_ENTRY();
  // blockIdx.y = image row, blockIdx.z = plane: no per-element div/mod on 64-bit indices
  const int y = blockIdx.y, c = blockIdx.z;
  int yr = y + roll_y;                                     // host passes the roll reduced to [0, H)
  yr = yr >= H ? yr - H : yr;
  const int ty = min(yr / th, nty - 1);
  float* out = grad + ((size_t)c * H + y) * W;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
    int xr = x + roll_x;
    xr = xr >= W ? xr - W : xr;
    const int tx = min(xr / tw, ntx - 1);
    const int t = ty * ntx + tx;
    const int rank = t % world, slot = t / world;
    const size_t base = (size_t)rank * rank_stride + ((size_t)slot * 3 + c) * thmax * twmax;
    out[x] = packed[base + (size_t)(yr - ty * th) * twmax + (xr - tx * tw)];
  }
  // the ranks' losses ride in the tails of their chunks: added in rank order by one thread
  if (loss_accum != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
    *loss_accum += packed_loss_sum(packed, world, rank_stride);
}

int unpack_grad(const float* packed, int H, int W, int roll_y, int roll_x, int nty, int ntx,
                int th, int tw, int thmax, int twmax, int world, int tiles_per_rank, float* grad,
                double* loss_accum, cudaStream_t s) {
  TimerScope ts(s, kTimeImage, 8.0 * 3 * H * W);
  const int ry = ((roll_y % H) + H) % H, rx = ((roll_x % W) + W) % W;
  ST_LAUNCH(unpack_grad_kernel, dim3(cdiv(W, 256), H, 3), 256, 0, s, packed, H, W, ry, rx, nty, ntx,
            th, tw, thmax, twmax, world, packed_rank_stride(tiles_per_rank, thmax, twmax), grad,
            loss_accum);
  return ST_OK;
}

// -----------------------------------------------------------------------------------------------------
// Fused regularisers on the un-rolled image.
// -----------------------------------------------------------------------------------------------------
// |a|^(k) for small integer k by repeated multiplication (p_pow = 6 is the reference's default:
// two powf per pixel made this kernel compute-bound)
__device__ __forceinline__ float powi(float x, int k) {
  float r = 1.f;
  while (k > 0) {
    if (k & 1) r *= x;
    x *= x, k >>= 1;
  }
  return r;
}

struct TvTerm {
  float ddx, ddy, pw;   // d/d(dx), d/d(dy) contributions and g2^(beta/2)
};

template <bool BETA2 = false>
__device__ __forceinline__ TvTerm tv_term(float xc, float xr, float xd, float beta) {
  // xc = X[y][x], xr = X[y][x+1], xd = X[y+1][x]  (already divided by 127.5)
  const float dx = xc - xr, dy = xc - xd;
  const float g2 = dx * dx + dy * dy + kEps;
  TvTerm t;
  float dg;
  if (BETA2 || beta == 2.f) {
    t.pw = g2;
    dg = 1.f;
  } else if (beta == 1.f) {
    t.pw = sqrtf(g2);
    dg = 0.5f / t.pw;
  } else {
    t.pw = powf(g2, 0.5f * beta);
    dg = 0.5f * beta * powf(g2, 0.5f * beta - 1.f);
  }
  t.ddx = 2.f * dx * dg;
  t.ddy = 2.f * dy * dg;
  return t;
}

// Persistent blocks, one 32 x 8 pixel tile of one plane at a time.  The scaled pixels (img / 127.5,
// an IEEE division to match the reference) of the tile plus a one-pixel rim are staged in shared
// memory once; index arithmetic uses FastDiv (the first version spent 760 instructions per pixel on
// seven divisions and 64-bit div/mod: instruction-bound at 4 % of the HBM bandwidth).  When `packed` is given the tiled gradient is
// gathered from the all-gather buffer in the same pass (st_unpack_grad fused in).
constexpr int kRTW = 32, kRTH = 32, kRThreads = 256, kRRows = kRTH / (kRThreads / kRTW);   // 4 rows / thread

struct UnpackGeom {
  int roll_y, roll_x;          // reduced to [0, H) x [0, W)
  int nty, ntx, th, tw, thmax, twmax, world, tiles_per_rank;
  size_t rank_stride;          // floats between the chunks of two ranks (packed_rank_stride)
  FastDiv div_th, div_tw, div_world;
};
struct RegTiles {
  int tiles_x, num_tiles;
  FastDiv div_tiles_x, div_plane;    // by tiles_x and by tiles_x * tiles_y
};

// FAST: the reference's default configuration (tv_power = 2, p_power = 6 or no p-norm, no aux
// image) with every per-pixel branch on the parameters resolved at compile time -- 57 % of the
// generic kernel's executed instructions were integer / constant-load / branch overhead
// (profiles/r01_laggards_ncu.md).  Same arithmetic, same operation order.
template <bool FAST, bool PACKED>
__global__ void __launch_bounds__(kRThreads)
regularizers_kernel(const float* __restrict__ img, int H, int W, float m0, float m1, float m2,
                    float tv_w, float tv_beta, float p_w, float p_pow,
                    const float* __restrict__ aux, float aux_w, int roll_y, int roll_x,
                    double* loss_accum, float* __restrict__ grad,
                    const float* __restrict__ packed, UnpackGeom ug, RegTiles rt, ReduceScratch rs) {
  ST_PDL_ENTRY();
  __shared__ float xs[kRTH + 2][kRTW + 2];       // scaled pixels, origin (y0-1, x0-1)
  const int tx = threadIdx.x % kRTW, ty0 = threadIdx.x / kRTW;
  const int num_tiles = rt.num_tiles;
  const bool do_tv = FAST || tv_w != 0.f;
  // img / 127.5 as a multiplication by the rounded reciprocal: within 1 ulp of the reference's
  // division, and the IEEE division sequence was a fifth of this kernel's issue slots
  const float inv = 1.f / 127.5f;
  double v[1] = {0.0};
  float part = 0.f;
  int since_flush = 0;
  // persistent blocks walk the tile list: one grid-wide reduction per block, not per tile
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int c = (int)rt.div_plane.div(tile), rem = tile - c * (int)rt.div_plane.d;
    const int trow = (int)rt.div_tiles_x.div(rem);
    const int y0 = trow * kRTH, x0 = (rem - trow * rt.tiles_x) * kRTW;
    const float* pl = img + (size_t)c * H * W;
    if (do_tv) {
      // all of a thread's halo loads are issued before the first one is used (a rolled load ->
      // store loop paid one DRAM latency per element: 27 % of this kernel's stall samples,
      // profiles/r01_laggards_ncu.md)
      constexpr int kHalo = (kRTH + 2) * (kRTW + 2), kHaloIt = (kHalo + kRThreads - 1) / kRThreads;
      float hv[kHaloIt];
#pragma unroll
      for (int u = 0; u < kHaloIt; ++u) {
        const int i = threadIdx.x + u * kRThreads;
        hv[u] = 0.f;
        if (i < kHalo) {
          const int r = i / (kRTW + 2), q = i - r * (kRTW + 2);
          int yy = y0 - 1 + r, xx = x0 - 1 + q;             // periodic (num_utils.py:150-162)
          if (yy < 0 || yy >= H) yy = wrap(yy, H);
          if (xx < 0 || xx >= W) xx = wrap(xx, W);
          hv[u] = pl[(size_t)yy * W + xx];
        }
      }
      __syncthreads();                                      // previous tile's readers are done
#pragma unroll
      for (int u = 0; u < kHaloIt; ++u) {
        const int i = threadIdx.x + u * kRThreads;
        if (i < kHalo) {
          const int r = i / (kRTW + 2), q = i - r * (kRTW + 2);
          xs[r][q] = hv[u] * inv;
        }
      }
      __syncthreads();
    }
    const int x = x0 + tx;
    // column part of the gradient-tile lookup: the same for the four rows of this thread
    int xr = 0, txx = 0;
    if (PACKED) {
      xr = x + ug.roll_x;
      xr = xr >= W ? xr - W : xr;
      txx = min((int)ug.div_tw.div(xr), ug.ntx - 1);
    }
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    // the global operands of this thread's four rows are requested together, then consumed
    float raw_k[kRRows], base_k[kRRows], aux_k[kRRows];
#pragma unroll
    for (int k = 0; k < kRRows; ++k) {
      const int y = y0 + ty0 + k * (kRThreads / kRTW);
      raw_k[k] = base_k[k] = aux_k[k] = 0.f;
      if (x >= W || y >= H) continue;
      raw_k[k] = pl[(size_t)y * W + x];
      if (PACKED) {                          // fused st_unpack_grad
        int yr = y + ug.roll_y;
        yr = yr >= H ? yr - H : yr;
        const int tyy = min((int)ug.div_th.div(yr), ug.nty - 1);
        const int t = tyy * ug.ntx + txx;
        const int slot = (int)ug.div_world.div(t), rank = t - slot * ug.world;
        const size_t pb = (size_t)rank * ug.rank_stride + ((size_t)slot * 3 + c) * ug.thmax * ug.twmax;
        base_k[k] = packed[pb + (size_t)(yr - tyy * ug.th) * ug.twmax + (xr - txx * ug.tw)];
      } else {
        base_k[k] = grad[((size_t)c * H + y) * W + x];
      }
      if (!FAST && aux != nullptr) {
        const int ya = wrap(y + roll_y, H), xa = wrap(x + roll_x, W);
        aux_k[k] = aux[((size_t)c * H + ya) * W + xa];
      }
    }
#pragma unroll
    for (int k = 0; k < kRRows; ++k) {
      const int ty = ty0 + k * (kRThreads / kRTW), y = y0 + ty;
      if (x >= W || y >= H) continue;
      const size_t i = ((size_t)c * H + y) * W + x;
      const float raw = raw_k[k];
      float g = 0.f, l = 0.f;
      if (do_tv) {
        // own term and the terms of the left / upper neighbour (their d/d(dx), d/d(dy) reach here)
        const float xc = xs[ty + 1][tx + 1];
        const TvTerm t0 = tv_term<FAST>(xc, xs[ty + 1][tx + 2], xs[ty + 2][tx + 1], tv_beta);
        const TvTerm tl = tv_term<FAST>(xs[ty + 1][tx], xc, xs[ty + 2][tx], tv_beta);
        const TvTerm tu = tv_term<FAST>(xs[ty][tx + 1], xs[ty][tx + 2], xc, tv_beta);
        g += tv_w * (t0.ddx + t0.ddy - tl.ddx - tu.ddy);
        l += tv_w * t0.pw;
      }
      if (p_w != 0.f) {
        const float a = (raw + mean - 127.5f) * inv;
        const float mag = fabsf(a), sgn = a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f);
        if (!FAST && p_pow == 1.f) {
          l += p_w * mag, g += p_w * sgn;
        } else if (!FAST && p_pow == 2.f) {
          l += p_w * a * a, g += p_w * 2.f * a;
        } else if (FAST || p_pow == 6.f) {                  // the reference's default
          const float a2 = a * a, m5 = a2 * a2 * mag;
          l += p_w * m5 * mag, g += p_w * 6.f * sgn * m5;
        } else {
          const int ip = (int)p_pow;
          const float mp1 =
              ((float)ip == p_pow && ip <= 16) ? powi(mag, ip - 1) : powf(mag, p_pow - 1.f);
          l += p_w * mp1 * mag, g += p_w * p_pow * sgn * mp1;
        }
      }
      if (!FAST && aux != nullptr) {
        const float d = (raw - aux_k[k]) * inv;
        l += aux_w * 0.5f * d * d, g += aux_w * d;
      }
      grad[i] = base_k[k] + g;
      part += l;
    }
    if (++since_flush == 4) v[0] += (double)part, part = 0.f, since_flush = 0;
  }
  v[0] += (double)part;
  if (grid_reduce<1>(v, rs.partials, rs.counter))
    atomicAdd(loss_accum, PACKED ? v[0] + packed_loss_sum(packed, ug.world, ug.rank_stride) : v[0]);
}

// The default configuration (tv_power = 2, p_power = 6 or no p-norm, no aux image) fused with the
// gradient stitch, as a column-strip stencil: a thread owns four consecutive pixels of one plane and
// walks kRsRows rows downwards with the rows above / below in registers, so the image is read once
// with 16-byte accesses (+ two scalar neighbours per row that hit L1 / L2), the gradient tiles are
// gathered with 16-byte loads and the result leaves with 16-byte stores.  Same per-pixel arithmetic
// as regularizers_kernel<true, true> (which needed ~100 us for the 151 MB of a 2048^2 image: halo
// staging through shared memory with scalar accesses, 0.24 of the HBM bandwidth).
// Needs W, tile width and roll_x to be multiples of four.
constexpr int kRsRows = 16, kRsThreads = 128;
static const bool g_no_reg_strip = getenv("ST_NO_REG_STRIP") != nullptr;   // debugging switch
__global__ void __launch_bounds__(kRsThreads)
regularizers_strip_kernel(const float* __restrict__ img, int H, int W, float m0, float m1, float m2,
                          float tv_w, float p_w, double* loss_accum, float* __restrict__ grad,
                          const float* __restrict__ packed, UnpackGeom ug, ReduceScratch rs) {
  ST_PDL_ENTRY();
  const int x = (blockIdx.x * kRsThreads + threadIdx.x) * 4;
  const int c = blockIdx.z, y0 = blockIdx.y * kRsRows;
  const float inv = 1.f / 127.5f;
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
  double total = 0.0;
  if (x < W) {
    const float* pl = img + (size_t)c * H * W;
    const int xl = x == 0 ? W - 1 : x - 1, xr4 = x + 4 == W ? 0 : x + 4;
    // column part of the gradient-tile lookup (the same for every row of the strip)
    int xr = x + ug.roll_x;
    xr = xr >= W ? xr - W : xr;
    const int txx = min((int)ug.div_tw.div(xr), ug.ntx - 1);
    const int xin = xr - txx * ug.tw;
    auto ld4 = [&](int y) { return *reinterpret_cast<const float4*>(pl + (size_t)y * W + x); };
    const int yu = y0 == 0 ? H - 1 : y0 - 1;
    float4 up = ld4(yu), cur = ld4(y0);
    up.x *= inv, up.y *= inv, up.z *= inv, up.w *= inv;
    float part = 0.f;
#pragma unroll 2
    for (int r = 0; r < kRsRows; ++r) {
      const int y = y0 + r;
      if (y >= H) break;
      const int yn = y + 1 == H ? 0 : y + 1;
      const float4 nxt = ld4(yn);
      const float left = pl[(size_t)y * W + xl] * inv, right = pl[(size_t)y * W + xr4] * inv;
      // fused st_unpack_grad: this row of the strip inside its gradient tile
      int yr = y + ug.roll_y;
      yr = yr >= H ? yr - H : yr;
      const int tyy = min((int)ug.div_th.div(yr), ug.nty - 1);
      const int t = tyy * ug.ntx + txx;
      const int slot = (int)ug.div_world.div(t), rank = t - slot * ug.world;
      const size_t pb = (size_t)rank * ug.rank_stride + ((size_t)slot * 3 + c) * ug.thmax * ug.twmax;
      const float4 base =
          *reinterpret_cast<const float4*>(packed + pb + (size_t)(yr - tyy * ug.th) * ug.twmax + xin);
      const float raw[4] = {cur.x, cur.y, cur.z, cur.w};
      const float xs[6] = {left, cur.x * inv, cur.y * inv, cur.z * inv, cur.w * inv, right};
      const float dn[4] = {nxt.x * inv, nxt.y * inv, nxt.z * inv, nxt.w * inv};
      const float upv[4] = {up.x, up.y, up.z, up.w};
      const float bs[4] = {base.x, base.y, base.z, base.w};
      float out[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xc = xs[i + 1];
        // own term and the terms of the left / upper neighbour (their d/d(dx), d/d(dy) reach here)
        const TvTerm t0 = tv_term<true>(xc, xs[i + 2], dn[i], 2.f);
        const TvTerm tl = tv_term<true>(xs[i], xc, 0.f, 2.f);          // only ddx is used
        const TvTerm tu = tv_term<true>(upv[i], 0.f, xc, 2.f);         // only ddy is used
        float g = tv_w * (t0.ddx + t0.ddy - tl.ddx - tu.ddy);
        float l = tv_w * t0.pw;
        if (p_w != 0.f) {
          const float a = (raw[i] + mean - 127.5f) * inv;
          const float mag = fabsf(a), sgn = a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f);
          const float a2 = a * a, m5 = a2 * a2 * mag;
          l += p_w * m5 * mag, g += p_w * 6.f * sgn * m5;
        }
        out[i] = bs[i] + g;
        part += l;
      }
      *reinterpret_cast<float4*>(grad + ((size_t)c * H + y) * W + x) =
          make_float4(out[0], out[1], out[2], out[3]);
      up = make_float4(xs[1], xs[2], xs[3], xs[4]);
      cur = nxt;
      if ((r & 3) == 3) total += (double)part, part = 0.f;
    }
    total += (double)part;
  }
  double v[1] = {total};
  // blocks of all three planes reduce together: linear block index over (x, y, z)
  if (grid_reduce_impl<1>(v, rs.partials, rs.counter,
                          (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x,
                          gridDim.x * gridDim.y * gridDim.z))
    atomicAdd(loss_accum, v[0] + packed_loss_sum(packed, ug.world, ug.rank_stride));
}

static int launch_regularizers(const float* img, int H, int W, float m0, float m1, float m2,
                               float tv_w, float tv_beta, float p_w, float p_pow, const float* aux,
                               float aux_w, int roll_y, int roll_x, double* loss_accum, float* grad,
                               const float* packed, const UnpackGeom& ug, ReduceScratch rs,
                               cudaStream_t s) {
  RegTiles rt;
  rt.tiles_x = cdiv(W, kRTW);
  const int tiles_y = cdiv(H, kRTH);
  rt.num_tiles = 3 * rt.tiles_x * tiles_y;
  rt.div_tiles_x = FastDiv(rt.tiles_x), rt.div_plane = FastDiv(rt.tiles_x * tiles_y);
  const int num_tiles = rt.num_tiles;
  const int grid = num_tiles < 148 * 8 ? num_tiles : 148 * 8;
  TimerScope ts(s, kTimeImage, 4.0 * 3 * H * W * (3 + (aux ? 1 : 0)));
  const bool fast = tv_w != 0.f && tv_beta == 2.f && (p_w == 0.f || p_pow == 6.f) && aux == nullptr;
  auto launch = [&](auto kern) -> int {
    ST_LAUNCH(kern, grid, kRThreads, 0, s, img, H, W, m0, m1, m2, tv_w, tv_beta, p_w, p_pow, aux,
              aux_w, roll_y, roll_x, loss_accum, grad, packed, ug, rt, rs);
    return ST_OK;
  };
  if (fast && packed && W % 4 == 0 && ug.tw % 4 == 0 && ug.roll_x % 4 == 0 && W >= 8 && !g_no_reg_strip) {
    const dim3 sgrid(cdiv(W / 4, kRsThreads), cdiv(H, kRsRows), 3);
    if ((long)sgrid.x * sgrid.y * sgrid.z <= kMaxReduceBlocks) {
      ST_LAUNCH(regularizers_strip_kernel, sgrid, kRsThreads, 0, s, img, H, W, m0, m1, m2, tv_w, p_w,
                loss_accum, grad, packed, ug, rs);
      return ST_OK;
    }
  }
  if (fast) return packed ? launch(regularizers_kernel<true, true>) : launch(regularizers_kernel<true, false>);
  return packed ? launch(regularizers_kernel<false, true>) : launch(regularizers_kernel<false, false>);
}

int regularizers(const float* img, int H, int W, float m0, float m1, float m2, float tv_w,
                 float tv_beta, float p_w, float p_pow, const float* aux, float aux_w, int roll_y,
                 int roll_x, double* loss_accum, float* grad, ReduceScratch rs, cudaStream_t s) {
  return launch_regularizers(img, H, W, m0, m1, m2, tv_w, tv_beta, p_w, p_pow, aux, aux_w, roll_y,
                             roll_x, loss_accum, grad, nullptr, UnpackGeom{}, rs, s);
}

int unpack_regularizers(const float* packed, int H, int W, int nty, int ntx, int th, int tw,
                        int thmax, int twmax, int world, int tiles_per_rank, const float* img,
                        float m0, float m1, float m2, float tv_w, float tv_beta, float p_w,
                        float p_pow, const float* aux, float aux_w, int roll_y, int roll_x,
                        double* loss_accum, float* grad, ReduceScratch rs, cudaStream_t s) {
  UnpackGeom ug{((roll_y % H) + H) % H, ((roll_x % W) + W) % W, nty, ntx, th, tw, thmax, twmax,
                world, tiles_per_rank, packed_rank_stride(tiles_per_rank, thmax, twmax),
                FastDiv(th), FastDiv(tw), FastDiv(world)};
  return launch_regularizers(img, H, W, m0, m1, m2, tv_w, tv_beta, p_w, p_pow, aux, aux_w, roll_y,
                             roll_x, loss_accum, grad, packed, ug, rs, s);
}

// -----------------------------------------------------------------------------------------------------
// Adam + EWMA iterate averaging.  Operation order (and the absence of FMA contraction) follows the
// numpy expressions of optimizers.py:35-42 / average.EWMA so that results agree to the last bits.
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ewma(float value, float beta, float omb, float x) {
  return __fadd_rn(__fmul_rn(value, beta), __fmul_rn(omb, x));
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m1, float& m2, float& a1,
                                         float& avg, float neg_step, float b1, float omb1, float b2,
                                         float omb2, float bp1, float ombp1, float g1c, float g2c,
                                         float p1c) {
  m1 = ewma(m1, b1, omb1, g);
  m2 = ewma(m2, b2, omb2, __fmul_rn(g, g));
  const float step = __fdiv_rn(__fdiv_rn(m1, g1c), __fadd_rn(__fsqrt_rn(__fdiv_rn(m2, g2c)), kEps));
  p = __fadd_rn(p, __fmul_rn(neg_step, step));
  a1 = ewma(a1, bp1, ombp1, p);
  avg = __fdiv_rn(a1, p1c);
}

// float4 per thread and array (the state arrays come from the allocator: 16-byte aligned); the
// n % 4 tail is done by the first threads.  Element-wise arithmetic identical to the scalar form.
__global__ void adam_kernel(float* __restrict__ params, const float* __restrict__ grad,
                            float* __restrict__ g1, float* __restrict__ g2, float* __restrict__ p1,
                            float* __restrict__ avg, size_t n, float neg_step, float b1, float omb1,
                            float b2, float omb2, float bp1, float ombp1, float g1c, float g2c,
                            float p1c) {
  ST_PDL_ENTRY();
  const size_t n4 = n >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    float4 m1 = reinterpret_cast<float4*>(g1)[i], m2 = reinterpret_cast<float4*>(g2)[i];
    float4 p = reinterpret_cast<float4*>(params)[i], a1 = reinterpret_cast<float4*>(p1)[i], av;
    adam_one(p.x, g.x, m1.x, m2.x, a1.x, av.x, neg_step, b1, omb1, b2, omb2, bp1, ombp1, g1c, g2c, p1c);
    adam_one(p.y, g.y, m1.y, m2.y, a1.y, av.y, neg_step, b1, omb1, b2, omb2, bp1, ombp1, g1c, g2c, p1c);
    adam_one(p.z, g.z, m1.z, m2.z, a1.z, av.z, neg_step, b1, omb1, b2, omb2, bp1, ombp1, g1c, g2c, p1c);
    adam_one(p.w, g.w, m1.w, m2.w, a1.w, av.w, neg_step, b1, omb1, b2, omb2, bp1, ombp1, g1c, g2c, p1c);
    reinterpret_cast<float4*>(g1)[i] = m1, reinterpret_cast<float4*>(g2)[i] = m2;
    reinterpret_cast<float4*>(params)[i] = p, reinterpret_cast<float4*>(p1)[i] = a1;
    reinterpret_cast<float4*>(avg)[i] = av;
  }
  const size_t t = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n)
    adam_one(params[t], grad[t], g1[t], g2[t], p1[t], avg[t], neg_step, b1, omb1, b2, omb2, bp1, ombp1,
             g1c, g2c, p1c);
}

int adam_step(float* params, const float* grad, float* g1, float* g2, float* p1, float* avg_out,
              size_t n, float step_size, float b1, float b2, float bp1, float g1_corr,
              float g2_corr, float p1_corr, cudaStream_t s) {
  // (1 - beta) is formed in double and rounded once, as numpy does for a Python-float scalar.
  const float omb1 = (float)(1.0 - (double)b1), omb2 = (float)(1.0 - (double)b2);
  const float ombp1 = (float)(1.0 - (double)bp1);
  TimerScope ts(s, kTimeImage, 40.0 * n);
  const uintptr_t align = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grad) |
                          reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(g2) |
                          reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(avg_out);
  ST_REQUIRE((align & 15) == 0, "st_adam_step: arrays must be 16-byte aligned");
  ST_LAUNCH(adam_kernel, ew_grid((n + 3) / 4, 256), 256, 0, s, params, grad, g1, g2, p1, avg_out, n,
            -step_size, b1, omb1, b2, omb2, bp1, ombp1, g1_corr, g2_corr, p1_corr);
  return ST_OK;
}

// -----------------------------------------------------------------------------------------------------
// Per-iteration output step: statistics of the averaged iterate (style_transfer.py:808-815) and the
// uint8 picture (CaffeModel.get_image :378-386).
// -----------------------------------------------------------------------------------------------------
// Blocks walk the 3*H image rows (fixed grid => the block-ordered reduction is reproducible), threads
// walk a row; the right neighbour wraps inside the row, the lower neighbour is the next row (mod H).
__global__ void __launch_bounds__(256)
iter_stats_kernel(const float* __restrict__ avg, float* __restrict__ old, int H, int W, double* stats,
                  ReduceScratch rs) {
  ST_PDL_ENTRY();
  double v[2] = {0.0, 0.0};
  for (int row = blockIdx.x; row < 3 * H; row += gridDim.x) {
    const int c = row / H, y = row - c * H;
    const float* r0 = avg + (size_t)row * W;
    const float* r1 = avg + ((size_t)c * H + (y + 1 == H ? 0 : y + 1)) * W;
    float* o = old + (size_t)row * W;
    float ab = 0.f, sq = 0.f;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      const float a = r0[x], xd = a - r0[x + 1 == W ? 0 : x + 1], yd = a - r1[x];
      ab += fabsf(a - o[x]);
      sq += xd * xd + yd * yd;
      o[x] = a;
    }
    v[0] += (double)ab, v[1] += (double)sq;
  }
  if (grid_reduce<2>(v, rs.partials, rs.counter)) stats[0] = v[0], stats[1] = v[1];
}

int iter_stats(const float* avg, float* old, int H, int W, double* stats, ReduceScratch rs,
               cudaStream_t s) {
  TimerScope ts(s, kTimeImage, 4.0 * 3 * H * W * 3);
  const int grid = 3 * H < 148 * 8 ? 3 * H : 148 * 8;
  ST_LAUNCH(iter_stats_kernel, grid, 256, 0, s, avg, old, H, W, stats, rs);
  return ST_OK;
}

// One thread per pixel: three plane reads (coalesced along x), three consecutive bytes written.
// float32 add, clip, truncation toward zero -- np.uint8(np.clip(params + mean, 0, 255)) exactly.
__global__ void __launch_bounds__(256)
get_image_u8_kernel(const float* __restrict__ params, int H, int W, float m0, float m1, float m2,
                    int bgr, uint8_t* __restrict__ out) {
  ST_PDL_ENTRY();
  const size_t n = (size_t)H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float v0 = __fadd_rn(params[i], m0), v1 = __fadd_rn(params[n + i], m1),
                v2 = __fadd_rn(params[2 * n + i], m2);
    const uint8_t b0 = (uint8_t)fminf(fmaxf(v0, 0.f), 255.f), b1 = (uint8_t)fminf(fmaxf(v1, 0.f), 255.f),
                  b2 = (uint8_t)fminf(fmaxf(v2, 0.f), 255.f);
    uint8_t* o = out + i * 3;
    o[0] = bgr ? b2 : b0, o[1] = b1, o[2] = bgr ? b0 : b2;
  }
}

int get_image_u8(const float* params, int H, int W, float m0, float m1, float m2, bool bgr,
                 uint8_t* out, cudaStream_t s) {
  TimerScope ts(s, kTimeImage, 15.0 * H * W);
  ST_LAUNCH(get_image_u8_kernel, ew_grid((size_t)H * W, 256), 256, 0, s, params, H, W, m0, m1, m2,
            bgr ? 1 : 0, out);
  return ST_OK;
}

// -----------------------------------------------------------------------------------------------------
// The whole output step in ONE pass (st_output_step): statistics + old := avg + uint8 picture.  Each
// thread owns four consecutive pixels of all three planes and walks kOsRows image rows downwards, so
// that the lower neighbour of one row is the centre of the next (every array is read once: 12 B per
// element + the 3 B/pixel picture; the two separate kernels above move 20 B per element with scalar
// accesses and ran at 0.2 of the HBM bandwidth).  Needs W % 4 == 0.
// -----------------------------------------------------------------------------------------------------
constexpr int kOsRows = 8, kOsThreads = 128;
__global__ void __launch_bounds__(kOsThreads)
output_step_kernel(const float* __restrict__ avg, float* __restrict__ old, int H, int W, float m0,
                   float m1, float m2, int bgr, double* stats, uint8_t* __restrict__ pic,
                   ReduceScratch rs) {
  ST_PDL_ENTRY();
  const int x = (blockIdx.x * kOsThreads + threadIdx.x) * 4;
  const int y0 = blockIdx.y * kOsRows;
  const size_t plane = (size_t)H * W;
  float ab = 0.f, sq = 0.f;
  if (x < W) {
    const int xr = x + 4 == W ? 0 : x + 4;                      // right neighbour of the last pixel
    float4 cur[3], nxt[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) cur[c] = *reinterpret_cast<const float4*>(avg + c * plane + (size_t)y0 * W + x);
    const float mean[3] = {m0, m1, m2};
#pragma unroll 2
    for (int r = 0; r < kOsRows; ++r) {
      const int y = y0 + r;
      if (y >= H) break;
      const int yn = y + 1 == H ? 0 : y + 1;
      float right[3];
      float4 o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        nxt[c] = *reinterpret_cast<const float4*>(avg + c * plane + (size_t)yn * W + x);
        right[c] = avg[c * plane + (size_t)y * W + xr];
        o[c] = *reinterpret_cast<const float4*>(old + c * plane + (size_t)y * W + x);
      }
      uint32_t bytes[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a[5] = {cur[c].x, cur[c].y, cur[c].z, cur[c].w, right[c]};
        const float dn[4] = {nxt[c].x, nxt[c].y, nxt[c].z, nxt[c].w};
        const float ov[4] = {o[c].x, o[c].y, o[c].z, o[c].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xd = a[i] - a[i + 1], yd = a[i] - dn[i];
          ab += fabsf(a[i] - ov[i]);
          sq += xd * xd + yd * yd;
          bytes[c][i] = (uint32_t)(uint8_t)fminf(fmaxf(__fadd_rn(a[i], mean[c]), 0.f), 255.f);
        }
        *reinterpret_cast<float4*>(old + c * plane + (size_t)y * W + x) = cur[c];
        cur[c] = nxt[c];
      }
      if (pic != nullptr) {
        // picture channel k of a pixel is plane (bgr ? 2 - k : k); 4 pixels = 12 consecutive bytes
        const int p0 = bgr ? 2 : 0, p2 = bgr ? 0 : 2;
        uint32_t w3[3];
        uint8_t b[12];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          b[3 * i] = (uint8_t)bytes[p0][i], b[3 * i + 1] = (uint8_t)bytes[1][i], b[3 * i + 2] = (uint8_t)bytes[p2][i];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          w3[j] = b[4 * j] | (b[4 * j + 1] << 8) | (b[4 * j + 2] << 16) | ((uint32_t)b[4 * j + 3] << 24);
        uint32_t* dst = reinterpret_cast<uint32_t*>(pic + ((size_t)y * W + x) * 3);
        dst[0] = w3[0], dst[1] = w3[1], dst[2] = w3[2];
      }
    }
  }
  double v[2] = {(double)ab, (double)sq};
  if (grid_reduce_2d<2>(v, rs.partials, rs.counter)) stats[0] = v[0], stats[1] = v[1];
}

int output_step(const float* avg, float* old, int H, int W, float m0, float m1, float m2, bool bgr,
                double* stats, uint8_t* pic, ReduceScratch rs, cudaStream_t s) {
  if (W % 4 != 0) {                        // odd widths: the two element-wise kernels
    int rc = iter_stats(avg, old, H, W, stats, rs, s);
    if (rc == ST_OK && pic != nullptr) rc = get_image_u8(avg, H, W, m0, m1, m2, bgr, pic, s);
    return rc;
  }
  TimerScope ts(s, kTimeImage, 4.0 * 3 * H * W * 3 + 3.0 * H * W);
  const dim3 grid(cdiv(W / 4, kOsThreads), cdiv(H, kOsRows));
  ST_REQUIRE((long)grid.x * grid.y <= kMaxReduceBlocks / 2, "output_step: image too large for the reduction scratch");
  ST_LAUNCH(output_step_kernel, grid, kOsThreads, 0, s, avg, old, H, W, m0, m1, m2, bgr ? 1 : 0, stats,
            pic, rs);
  return ST_OK;
}

// -----------------------------------------------------------------------------------------------------
// Scale change: one pass of Pillow's float resampling (num_utils.resize :90-108).  One thread per
// output sample; the weights of an output index are shared by a whole row (x pass) or column (y
// pass).  double multiply and add kept separate (__dmul_rn / __dadd_rn): PIL's C loop is not
// contracted into FMAs, and the result is rounded to float32 once per pass, as PIL stores it.
// -----------------------------------------------------------------------------------------------------
template <bool ALONG_Y>
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, float* __restrict__ out, int in_h, int in_w, int out_h,
                int out_w, const int* __restrict__ bounds, const double* __restrict__ kk, int ksize) {
  ST_PDL_ENTRY();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, c = blockIdx.z;
  if (x >= out_w) return;
  const float* src = in + (size_t)c * in_h * in_w;
  const int o = ALONG_Y ? y : x;                       // index of the output sample along the pass
  const int first = bounds[2 * o], count = bounds[2 * o + 1];
  const double* k = kk + (size_t)o * ksize;
  double ss = 0.0;
  for (int i = 0; i < count; ++i) {
    const float v = ALONG_Y ? src[(size_t)(first + i) * in_w + x] : src[(size_t)y * in_w + first + i];
    ss = __dadd_rn(ss, __dmul_rn((double)v, k[i]));
  }
  out[((size_t)c * out_h + y) * out_w + x] = (float)ss;
}

int resample_pass(const float* in, float* out, int channels, int in_h, int in_w, int out_size,
                  bool along_y, const int* bounds_dev, const double* kk_dev, int ksize, cudaStream_t s) {
  const int out_h = along_y ? out_size : in_h, out_w = along_y ? in_w : out_size;
  const dim3 grid(cdiv(out_w, 256), out_h, channels);
  if (along_y) {
    ST_LAUNCH(resample_kernel<true>, grid, 256, 0, s, in, out, in_h, in_w, out_h, out_w, bounds_dev,
              kk_dev, ksize);
  } else {
    ST_LAUNCH(resample_kernel<false>, grid, 256, 0, s, in, out, in_h, in_w, out_h, out_w, bounds_dev,
              kk_dev, ksize);
  }
  return ST_OK;
}

// -----------------------------------------------------------------------------------------------------
// BLAS-1 on the device for L-BFGS.
// -----------------------------------------------------------------------------------------------------
template <bool ABS>
__global__ void dot_kernel(const float* __restrict__ x, const float* __restrict__ y, size_t n,
                           double* out, ReduceScratch rs) {
  ST_PDL_ENTRY();
  double v[1] = {0.0};
  float part = 0.f;
  int cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    part += ABS ? fabsf(x[i]) : x[i] * y[i];
    if (++cnt == 32) v[0] += part, part = 0.f, cnt = 0;
  }
  v[0] += part;
  if (grid_reduce<1>(v, rs.partials, rs.counter)) *out = v[0];
}

int dot_to(const float* x, const float* y, size_t n, double* out, ReduceScratch rs,
           cudaStream_t s) {
  auto k = dot_kernel<false>;
  ST_LAUNCH(k, ew_grid(n, 256), 256, 0, s, x, y, n, out, rs);
  return ST_OK;
}

int asum_to(const float* x, size_t n, double* out, ReduceScratch rs, cudaStream_t s) {
  auto k = dot_kernel<true>;
  ST_LAUNCH(k, ew_grid(n, 256), 256, 0, s, x, x, n, out, rs);
  return ST_OK;
}

__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y,
                             size_t n) {
  ST_PDL_ENTRY();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = b == 0.f ? a * x[i] : a * x[i] + b * y[i];
}

int axpby(float a, const float* x, float b, float* y, size_t n, cudaStream_t s) {
  ST_LAUNCH(axpby_kernel, ew_grid(n, 256), 256, 0, s, a, x, b, y, n);
  return ST_OK;
}

__global__ void axpy_dev_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n,
                                const double* __restrict__ num, double den,
                                const double* __restrict__ sub, double sign, double* store) {
  ST_PDL_ENTRY();
  const double first = num[0] / den;
  const float coef = (float)(sign * (sub ? first - sub[0] : first));
  if (store && blockIdx.x == 0 && threadIdx.x == 0) *store = first;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] += coef * x[i];
}

int axpy_dev(const float* x, float* y, size_t n, const double* num, double den, const double* sub,
             double sign, double* store, cudaStream_t s) {
  ST_LAUNCH(axpy_dev_kernel, ew_grid(n, 256), 256, 0, s, x, y, n, num, den, sub, sign, store);
  return ST_OK;
}

__global__ void scale_dev_kernel(float* __restrict__ y, size_t n, double num,
                                 const double* __restrict__ den) {
  ST_PDL_ENTRY();
  const float coef = (float)(num / den[0]);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] *= coef;
}

int scale_dev(float* y, size_t n, double num, const double* den, cudaStream_t s) {
  ST_LAUNCH(scale_dev_kernel, ew_grid(n, 256), 256, 0, s, y, n, num, den);
  return ST_OK;
}

// -----------------------------------------------------------------------------------------------------
// L-BFGS with its memory on the device (LBFGSOptimizer, optimizers.py:64-138): no host decisions.
// state (doubles): [0] count of valid pairs, [1] head = ring slot the NEXT pair is written to,
// [2] dot scratch, [3] sum|p| scratch, [4 .. 4+17) s.y per slot, [24 .. 24+16) alpha per pair.
// The ring has n_corr + 1 slots so that the candidate pair can be written before it is known
// whether it will be kept (optimizers.py:98: kept iff s.y > 1e-10); pair k (0 = newest) lives in
// slot (head - 1 - k) mod (n_corr + 1).
// -----------------------------------------------------------------------------------------------------
namespace {
constexpr int kLbSy = 4, kLbAlpha = 24;
__device__ __forceinline__ int lb_slot(const double* st, int k, int n_corr) {
  int sl = ((int)st[1] - 1 - k) % (n_corr + 1);
  return sl < 0 ? sl + n_corr + 1 : sl;
}
// pair index of iteration j: MODE 0 = first loop (newest -> oldest), 1 = second loop (oldest ->
// newest), 2 = the newest pair; < 0: nothing to do
template <int MODE>
__device__ __forceinline__ int lb_pair(const double* st, int j) {
  const int count = (int)st[0];
  if (MODE == 0) return j < count ? j : -1;
  if (MODE == 1) return count - 1 - j;
  return count > 0 ? 0 : -1;
}

// st[2] = dot(ring[slot(pair)], MODE == 2 ? the same vector : p)
template <int MODE>
__global__ void lb_dot_kernel(const float* __restrict__ ring, size_t n, int n_corr, double* st, int j,
                              const float* __restrict__ p, ReduceScratch rs) {
  ST_PDL_ENTRY();
  const int k = lb_pair<MODE>(st, j);
  if (k < 0) return;
  const float* x = ring + (size_t)lb_slot(st, k, n_corr) * n;
  const float* y = MODE == 2 ? x : p;
  double v[1] = {0.0};
  float part = 0.f;
  int cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    part += x[i] * y[i];
    if (++cnt == 32) v[0] += part, part = 0.f, cnt = 0;
  }
  v[0] += part;
  if (grid_reduce<1>(v, rs.partials, rs.counter)) st[2] = v[0];
}

// MODE 0: alpha_k = st[2] / sy_k; p -= alpha_k * y_k        (optimizers.py:110-111)
// MODE 1: beta = st[2] / sy_k;   p += (alpha_k - beta) * s_k  (:118-119)
// MODE 2: p *= sy_0 / st[2]                                    (:113-115)
template <int MODE>
__global__ void lb_update_kernel(const float* __restrict__ ring, size_t n, int n_corr, double* st,
                                 int j, float* __restrict__ p) {
  ST_PDL_ENTRY();
  const int k = lb_pair<MODE>(st, j);
  if (k < 0) return;
  const int slot = lb_slot(st, k, n_corr);
  const double dot = st[2], sy = st[kLbSy + slot];
  float coef;
  if (MODE == 0) {
    const double alpha = dot / sy;
    coef = (float)(-alpha);
    if (blockIdx.x == 0 && threadIdx.x == 0) st[kLbAlpha + k] = alpha;
  } else if (MODE == 1) {
    coef = (float)(st[kLbAlpha + k] - dot / sy);
  } else {
    coef = (float)(sy / dot);
  }
  const float* x = ring + (size_t)slot * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = MODE == 2 ? p[i] * coef : p[i] + coef * x[i];
}

// s = -scale * p into ring_s[head], params += s.  scale (:81-84): initial_step / mean|p| while the
// memory is empty, count / n_corr until it is full, then 1.  st[3] holds sum|p|.
__global__ void lb_step_kernel(const float* __restrict__ p, size_t n, int n_corr, const double* st,
                               float initial_step, float* __restrict__ ring_s,
                               float* __restrict__ params) {
  ST_PDL_ENTRY();
  const int count = (int)st[0];
  double scale = 1.0;
  if (count == 0)
    scale = (double)initial_step / (st[3] / (double)n);
  else if (count < n_corr)
    scale = (double)count / (double)n_corr;
  const float coef = (float)(-scale);
  float* s = ring_s + (size_t)((int)st[1]) * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float v = coef * p[i];
    s[i] = v;
    params[i] += v;
  }
}

// y = grad_new - grad_old into ring_y[head]; st[2] = s . y
__global__ void lb_y_kernel(const float* __restrict__ gn, const float* __restrict__ go, size_t n,
                            const float* __restrict__ ring_s, float* __restrict__ ring_y, double* st,
                            ReduceScratch rs) {
  ST_PDL_ENTRY();
  const size_t off = (size_t)((int)st[1]) * n;
  const float* s = ring_s + off;
  float* y = ring_y + off;
  double v[1] = {0.0};
  float part = 0.f;
  int cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float d = gn[i] - go[i];
    y[i] = d;
    part += s[i] * d;
    if (++cnt == 32) v[0] += part, part = 0.f, cnt = 0;
  }
  v[0] += part;
  if (grid_reduce<1>(v, rs.partials, rs.counter)) st[2] = v[0];
}

// keep the candidate pair iff s.y > 1e-10 (:98-103): advance the head, grow the count up to n_corr
__global__ void lb_commit_kernel(double* st, int n_corr) {
  ST_PDL_ENTRY();
  const double sy = st[2];
  if (sy > 1e-10) {
    const int head = (int)st[1];
    st[kLbSy + head] = sy;
    st[1] = (double)((head + 1) % (n_corr + 1));
    const int count = (int)st[0];
    st[0] = (double)(count < n_corr ? count + 1 : n_corr);
  }
}
}  // namespace

// ---- the same recursion as ONE cooperative kernel ---------------------------------------------------
// 2 * n_corr + 2 dot products and as many vector updates are ~45 dependent launches; at 512^2 each
// moves 3 MB and the launches (host enqueue + gaps) cost more than the work.  Here the blocks of one
// co-resident grid walk the phases with grid-wide barriers: block partial sums go to `partials`, and
// after the barrier every block adds them up in index order (the same double in every block).
namespace {
__device__ __forceinline__ void lb_grid_sync(unsigned* bar, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned target = (++gen) * gridDim.x;
    atomicAdd(bar, 1u);
    while (*reinterpret_cast<volatile unsigned*>(bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}
// sum over the grid of this block's `v`; every block returns the same value
__device__ __forceinline__ double lb_grid_sum(double v, double* partials, unsigned* bar, unsigned& gen) {
  __shared__ double sh[32];
  __shared__ double total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0;
    for (int w = 0; w < nwarps; ++w) x += sh[w];
    partials[blockIdx.x] = x;
  }
  lb_grid_sync(bar, gen);
  if (warp == 0) {
    double x = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) x += partials[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) total = x;
  }
  __syncthreads();
  // (every reduction below is followed by a vector update and a grid barrier before the next one
  // writes `partials`, so no second barrier is needed here)
  return total;
}
__device__ __forceinline__ double lb_dot(const float* x, const float* y, size_t n) {
  double acc = 0.0;
  float part = 0.f;
  int cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    part += x[i] * y[i];
    if (++cnt == 32) acc += part, part = 0.f, cnt = 0;
  }
  return acc + part;
}

__global__ void __launch_bounds__(256)
lbfgs_direction_coop_kernel(const float* __restrict__ grad, size_t n, int n_corr, float* ring_s,
                            const float* ring_y, double* st, float* p, float* params,
                            float initial_step, double* partials, unsigned* bar) {
  ST_PDL_ENTRY();
  unsigned gen = 0;
  const int count = (int)st[0], head = (int)st[1];
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  auto slot_of = [&](int k) {
    int sl = (head - 1 - k) % (n_corr + 1);
    return sl < 0 ? sl + n_corr + 1 : sl;
  };
  for (size_t i = tid; i < n; i += nthr) p[i] = grad[i];
  lb_grid_sync(bar, gen);
  double alpha[16];
  for (int k = 0; k < count; ++k) {                                  // optimizers.py:109-111
    const int sl = slot_of(k);
    const double dot = lb_grid_sum(lb_dot(ring_s + (size_t)sl * n, p, n), partials, bar, gen);
    alpha[k] = dot / st[kLbSy + sl];
    const float coef = (float)(-alpha[k]);
    const float* y = ring_y + (size_t)sl * n;
    for (size_t i = tid; i < n; i += nthr) p[i] += coef * y[i];
    lb_grid_sync(bar, gen);
  }
  if (count > 0) {                                                   // :113-115
    const int sl = slot_of(0);
    const float* y = ring_y + (size_t)sl * n;
    const double yy = lb_grid_sum(lb_dot(y, y, n), partials, bar, gen);
    const float coef = (float)(st[kLbSy + sl] / yy);
    for (size_t i = tid; i < n; i += nthr) p[i] *= coef;
    lb_grid_sync(bar, gen);
  }
  for (int k = count - 1; k >= 0; --k) {                             // :117-119
    const int sl = slot_of(k);
    const double dot = lb_grid_sum(lb_dot(ring_y + (size_t)sl * n, p, n), partials, bar, gen);
    const float coef = (float)(alpha[k] - dot / st[kLbSy + sl]);
    const float* sv = ring_s + (size_t)sl * n;
    for (size_t i = tid; i < n; i += nthr) p[i] += coef * sv[i];
    lb_grid_sync(bar, gen);
  }
  double scale = 1.0;                                                // :81-84
  if (count == 0) {
    double a = 0.0;
    float part = 0.f;
    int cnt = 0;
    for (size_t i = tid; i < n; i += nthr) {
      part += fabsf(p[i]);
      if (++cnt == 32) a += part, part = 0.f, cnt = 0;
    }
    scale = (double)initial_step / (lb_grid_sum(a + part, partials, bar, gen) / (double)n);
  } else if (count < n_corr) {
    scale = (double)count / (double)n_corr;
  }
  const float coef = (float)(-scale);
  float* s = ring_s + (size_t)head * n;
  for (size_t i = tid; i < n; i += nthr) {                           // :85
    const float v = coef * p[i];
    s[i] = v;
    params[i] += v;
  }
}
}  // namespace

int lbfgs_direction(const float* grad, size_t n, int n_corr, float* ring_s, const float* ring_y,
                    double* st, float* p, float* params, float initial_step, ReduceScratch rs,
                    cudaStream_t s) {
  // launch-bound regime only: on large images (4096^2: 201 MB per vector) the chain of wide
  // element-wise kernels below streams at the HBM rate and the barriers buy nothing
  static const bool no_coop = getenv("ST_LBFGS_NO_COOP") != nullptr;
  if (!no_coop && n <= ((size_t)1 << 22)) {
    // a co-resident grid: at most the number of blocks the device holds at once
    static int max_blocks = 0;
    if (max_blocks == 0) {
      int dev = 0, sms = 0, per_sm = 0;
      ST_CUDA(cudaGetDevice(&dev));
      ST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      ST_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_direction_coop_kernel, 256, 0));
      max_blocks = sms * (per_sm < 4 ? per_sm : 4);
    }
    int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)max_blocks);
    double* partials = rs.partials;
    unsigned* bar = rs.counter;          // zero between kernels (the reduction scratch contract)
    void* args[] = {(void*)&grad, (void*)&n, (void*)&n_corr, (void*)&ring_s, (void*)&ring_y, (void*)&st,
                    (void*)&p, (void*)&params, (void*)&initial_step, (void*)&partials, (void*)&bar};
    g_launches.fetch_add(1, std::memory_order_relaxed);
    ST_CUDA(cudaLaunchCooperativeKernel((const void*)lbfgs_direction_coop_kernel, dim3(grid), dim3(256),
                                        args, 0, s));
    // the barrier counter goes back to zero for the next user of the reduction scratch
    ST_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), s));
    return ST_OK;
  }
  const int grid = ew_grid(n, 256);
  ST_CUDA(cudaMemcpyAsync(p, grad, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  for (int j = 0; j < n_corr; ++j) {                         // optimizers.py:109-111
    ST_LAUNCH(lb_dot_kernel<0>, grid, 256, 0, s, ring_s, n, n_corr, st, j, p, rs);
    ST_LAUNCH(lb_update_kernel<0>, grid, 256, 0, s, ring_y, n, n_corr, st, j, p);
  }
  ST_LAUNCH(lb_dot_kernel<2>, grid, 256, 0, s, ring_y, n, n_corr, st, 0, p, rs);   // :113-115
  ST_LAUNCH(lb_update_kernel<2>, grid, 256, 0, s, ring_y, n, n_corr, st, 0, p);
  for (int j = 0; j < n_corr; ++j) {                         // :117-119
    ST_LAUNCH(lb_dot_kernel<1>, grid, 256, 0, s, ring_y, n, n_corr, st, j, p, rs);
    ST_LAUNCH(lb_update_kernel<1>, grid, 256, 0, s, ring_s, n, n_corr, st, j, p);
  }
  int rc = asum_to(p, n, st + 3, rs, s);
  if (rc != ST_OK) return rc;
  ST_LAUNCH(lb_step_kernel, grid, 256, 0, s, p, n, n_corr, st, initial_step, ring_s, params);
  return ST_OK;
}

int lbfgs_commit(const float* grad_new, const float* grad_old, size_t n, int n_corr,
                 const float* ring_s, float* ring_y, double* st, ReduceScratch rs, cudaStream_t s) {
  ST_LAUNCH(lb_y_kernel, ew_grid(n, 256), 256, 0, s, grad_new, grad_old, n, ring_s, ring_y, st, rs);
  ST_LAUNCH(lb_commit_kernel, 1, 1, 0, s, st, n_corr);
  return ST_OK;
}

}  // namespace st
