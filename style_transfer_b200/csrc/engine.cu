// libstyle_b200 engine: context, per-tile forward/backward plan and the extern "C" entry points
// declared in include/style_b200.h.  See DESIGN.md for the data layout and the kernel list.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"
#include "kernels.h"

namespace st {

static thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};
bool g_pdl = getenv("ST_NO_PDL") == nullptr;
bool g_pdl_all = getenv("ST_PDL_ALL") != nullptr && getenv("ST_NO_PDL") == nullptr;
void set_error(const std::string& msg) { g_error = msg; }

// ---- per-kernel timing ---------------------------------------------------------------------------
bool g_timing_enabled = false;
namespace {
struct TimingRec {
  cudaEvent_t begin, end;
  int cat;
  double work;
};
std::vector<TimingRec> g_timing_recs;
std::vector<cudaEvent_t> g_timing_free;
cudaEvent_t timing_event() {
  cudaEvent_t e = nullptr;
  if (!g_timing_free.empty()) {
    e = g_timing_free.back();
    g_timing_free.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
}  // namespace
void timing_mark(cudaStream_t s, int category, double work, bool begin) {
  if (begin) {
    TimingRec r{timing_event(), timing_event(), category, work};
    cudaEventRecord(r.begin, s);
    g_timing_recs.push_back(r);
  } else {
    // scopes nest like a stack; close the innermost open record of this category
    for (size_t i = g_timing_recs.size(); i-- > 0;)
      if (g_timing_recs[i].cat == category) {
        cudaEventRecord(g_timing_recs[i].end, s);
        break;
      }
  }
}

struct LayerRt {
  int kind, bottom, top, cin, cout;
  float* w_fwd = nullptr;    // [9][cin][cout]
  float* w_bwd = nullptr;    // [9][cout][cin] (taps flipped); first layer: [9][cout][4]
  float* bias = nullptr;
  TcWeights tc;              // bf16 packs for the tcgen05 path
  bool has_params = false;
  // pooling layers: byte mask written when the forward ran fused into the producing convolution
  uint8_t* pool_mask = nullptr;
  size_t pool_mask_cap = 0;
  bool pool_mask_valid = false;
};

struct BlobRt {
  int c = 0;
  int producer = -1;         // layer index, -1 for data
  int scale = 1;             // 224 // shape[1] of style_transfer.py:415-419
  bool relu = false;         // conv output (carries an in-place ReLU)
  void* act = nullptr;
  size_t act_cap = 0;        // elements
  // ReLU bit mask of `act` ([pixel][c/32] words) for the tensor-core backward pass: written by the
  // forward kernel that produced the blob; 16x smaller than the activation the backward epilogue
  // would otherwise re-read
  uint32_t* bits = nullptr;
  size_t bits_cap = 0;       // words
  bool bits_valid = false;
  void* inj = nullptr;
  size_t inj_cap = 0;
  // When the only loss term of a blob is one style term, the style GEMM writes its raw output S
  // into `inj` and the per-tile factor w / (mean|S| + EPS) is applied by the consumer's epilogue
  // (`inj_scale`, kMaxBatch floats) instead of a separate scale-and-copy pass over S.
  float* inj_scale = nullptr;
  bool inj_deferred = false;
};

struct ContentTarget {
  float* nhwc = nullptr;
  int hf = 0, wf = 0;
};

}  // namespace st

using namespace st;

struct st_ctx {
  int device = 0, precision = 0, sm_count = 0;
  void* comm = nullptr;      // ncclComm_t of st_comm_init (multi-GPU exchange), else null
  int comm_rank = 0, comm_world = 1;
  size_t esize = 4;
  std::vector<LayerRt> layers;
  std::vector<BlobRt> blobs;
  void* gbuf[2] = {nullptr, nullptr};
  void* sbuf = nullptr;
  size_t gcap = 0;           // elements of gbuf[i] / sbuf
  void* split_buf = nullptr; // ST_PREC_TC32: fp16 [hi | lo] copy of the convolution input in flight
  size_t split_cap = 0;      // elements (4 bytes each)
  float grad_scale = 1.f;    // ST_PREC_TC32: power of two applied to gradients before the fp16 split
  bool tc32_tc_gram = false;     // ST_TC32_TC_GRAM=1: tc32 Gram matrices on the tensor cores (see build_injection)
  bool tc32_simt_style = false;  // ST_TC32_SIMT_STYLE=1: tc32 style GEMM on the SIMT kernel
  // per batch tile: Gram [C][C], its difference to the target (fp32 and the bf16 copy that is the
  // B operand of the tcgen05 style GEMM)
  float *gram = nullptr, *delta = nullptr, *part = nullptr;
  void* delta_16 = nullptr;              // bf16, or per-tile scaled fp16 in ST_PREC_FP16
  float* eps_eff = nullptr;              // [kMaxBatch] EPS of normalize() under that scaling
  unsigned* delta_max = nullptr;         // [kMaxBatch] max |delta| as float bits
  double* abs_partials = nullptr;        // partial sums of |S| of the style GEMM
  size_t part_floats = 0, abs_cap = 0;
  int max_batch = 1;                     // tiles evaluated per launch (1 in fp32 mode)
  double* scalars = nullptr;             // [kMaxBatch][kStatStride] device doubles
  ReduceScratch rs{nullptr, nullptr};
  std::map<std::pair<int, int>, ContentTarget> contents;   // (content index, blob)
  std::map<std::pair<int, int>, float*> styles;            // (style index, blob) -> full [C][C]
  int n_contents = 0, n_styles = 0;
  size_t workspace_bytes = 0;
  TcContext tc;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define ST_GUARD(ctx)                                  \
  ST_REQUIRE((ctx) != nullptr, "null context");        \
  DeviceGuard guard_((ctx)->device);                   \
  if (!guard_.ok) {                                    \
    set_error("cudaSetDevice failed");                 \
    return ST_ERR_CUDA;                                \
  }

int dev_alloc(st_ctx* ctx, void** p, size_t bytes) {
  ST_CUDA(cudaMalloc(p, bytes));
  ctx->workspace_bytes += bytes;
  return ST_OK;
}

int ensure(st_ctx* ctx, void** p, size_t* cap, size_t elems, size_t esize) {
  if (*cap >= elems) return ST_OK;
  if (*p) {
    ST_CUDA(cudaDeviceSynchronize());
    ST_CUDA(cudaFree(*p));
    ctx->workspace_bytes -= *cap * esize;
    *p = nullptr;
    *cap = 0;
  }
  int rc = dev_alloc(ctx, p, elems * esize);
  if (rc != ST_OK) return rc;
  *cap = elems;
  return ST_OK;
}

inline int pooled(int n) { return (n + 1) / 2; }
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int posmod(int a, int n) {
  const int r = a % n;
  return r < 0 ? r + n : r;
}

struct Dims {
  std::vector<int> h, w;
};

Dims blob_dims(const st_ctx* ctx, int h, int w) {
  Dims d;
  d.h.assign(ctx->blobs.size(), 0);
  d.w.assign(ctx->blobs.size(), 0);
  d.h[0] = h, d.w[0] = w;
  for (const LayerRt& l : ctx->layers) {
    if (l.kind == ST_CONV3X3)
      d.h[l.top] = d.h[l.bottom], d.w[l.top] = d.w[l.bottom];
    else
      d.h[l.top] = pooled(d.h[l.bottom]), d.w[l.top] = pooled(d.w[l.bottom]);
  }
  return d;
}

int reserve_for(st_ctx* ctx, const Dims& d, int last_layer, int nb) {
  size_t gmax = 0;
  for (int i = 0; i <= last_layer; ++i) {
    const int b = ctx->layers[i].top;
    const size_t n = (size_t)nb * d.h[b] * d.w[b] * ctx->blobs[b].c;
    int rc = ensure(ctx, &ctx->blobs[b].act, &ctx->blobs[b].act_cap, n, ctx->esize);
    if (rc != ST_OK) return rc;
    gmax = std::max(gmax, n);
  }
  if (ctx->gcap < gmax) {
    size_t c0 = ctx->gcap, c1 = ctx->gcap, c2 = ctx->gcap;
    int rc = ensure(ctx, &ctx->gbuf[0], &c0, gmax, ctx->esize);
    if (rc == ST_OK) rc = ensure(ctx, &ctx->gbuf[1], &c1, gmax, ctx->esize);
    if (rc == ST_OK) rc = ensure(ctx, &ctx->sbuf, &c2, gmax, ctx->esize);
    if (rc != ST_OK) return rc;
    ctx->gcap = gmax;
  }
  if (ctx->precision == ST_PREC_TC32) {
    int rc = ensure(ctx, &ctx->split_buf, &ctx->split_cap, std::max(gmax, ctx->gcap), 4);
    if (rc != ST_OK) return rc;
  }
  return ST_OK;
}

constexpr int kStatStride = 8;   // doubles per batch tile in ctx->scalars:
                                 //   [0] sum c^2  [1] sum |c|  [2] sum |S|  [3] the tile's loss

// ---- forward -------------------------------------------------------------------------------------
// A convolution whose output feeds exactly one pooling layer runs fused with it (tensor-core path):
// the pooled map and a byte mask for the backward pass come out of the conv epilogue; the un-pooled
// map is only written when somebody else needs it (`need_full`: loss layers, requested features).
template <typename T>
int fused_pool_layer(const st_ctx* ctx, int i, int last_layer) {
  if (sizeof(T) != 2 || !ctx->tc.pool_fusion) return -1;
  const LayerRt& l = ctx->layers[i];
  if (l.bottom == 0 || i + 1 > last_layer) return -1;
  const LayerRt& p = ctx->layers[i + 1];
  if (p.kind == ST_CONV3X3 || p.bottom != l.top) return -1;
  return i + 1;
}

template <typename T>
int forward(st_ctx* ctx, const ImageBatch& view, const Dims& d, int last_layer,
            const std::vector<char>& need_full, bool for_backward, cudaStream_t s) {
  const int nb = view.nb;
  for (LayerRt& l : ctx->layers) l.pool_mask_valid = false;
  for (BlobRt& b : ctx->blobs) b.bits_valid = false;
  for (int i = 0; i <= last_layer; ++i) {
    LayerRt& l = ctx->layers[i];
    const int hb = d.h[l.bottom], wb = d.w[l.bottom];
    T* out = static_cast<T*>(ctx->blobs[l.top].act);
    int rc;
    // the ReLU bit mask of this layer's output is wanted when a convolution consumes the blob and a
    // backward pass follows (its backward epilogue applies the mask)
    uint32_t* bits_out = nullptr;
    if (for_backward && l.kind == ST_CONV3X3 && sizeof(T) == 2 && ctx->tc.enabled && ctx->tc.pair_kernel &&
        ctx->tc.fwd_bits) {
      bool wanted = false;
      for (int j = i + 1; j <= last_layer; ++j)
        wanted = wanted || (ctx->layers[j].kind == ST_CONV3X3 && ctx->layers[j].bottom == l.top);
      if (wanted) {
        BlobRt& tb = ctx->blobs[l.top];
        rc = ensure(ctx, (void**)&tb.bits, &tb.bits_cap, (size_t)nb * hb * wb * (l.cout / 32), 4);
        if (rc != ST_OK) return rc;
        bits_out = tb.bits;
      }
    }
    if (l.kind == ST_CONV3X3) {
      ST_REQUIRE(l.has_params, "conv layer has no weights (st_set_conv_params)");
      if (l.bottom == 0) {
        bool done = false;
        if constexpr (sizeof(T) == 2) {
          if (ctx->tc.enabled && ctx->tc.pair_kernel && l.tc.fwd != nullptr && l.cout == 64) {
            rc = conv_first_fwd_tc(ctx->tc, l.tc, view, hb, wb, l.bias, out, bits_out, s);
            ctx->blobs[l.top].bits_valid = bits_out != nullptr && rc == ST_OK;
            done = true;
          }
        }
        if (!done) rc = conv_first_fwd<T>(view, hb, wb, l.w_fwd, l.bias, out, l.cout, s);
      } else {
        const T* in = static_cast<const T*>(ctx->blobs[l.bottom].act);
        const int pl = tc_usable<T>(ctx->tc, l.tc, l.cin, l.cout) && ctx->tc.pair_kernel
                           ? fused_pool_layer<T>(ctx, i, last_layer) : -1;
        if (pl >= 0) {
          if constexpr (sizeof(T) == 2) {
            LayerRt& p = ctx->layers[pl];
            // other readers of the un-pooled map: loss / feature requests, or another layer (_big)
            bool full = need_full[l.top] != 0;
            for (int j = pl + 1; j <= last_layer; ++j) full = full || ctx->layers[j].bottom == l.top;
            const size_t pooled = (size_t)nb * d.h[p.top] * d.w[p.top] * l.cout;
            rc = ensure(ctx, (void**)&p.pool_mask, &p.pool_mask_cap, pooled, 1);
            if (rc == ST_OK)
              rc = conv3x3_pool_tc_pair(ctx->tc, l.tc, in, out, static_cast<T*>(ctx->blobs[p.top].act),
                                        p.pool_mask, nb, hb, wb, l.cin, l.cout, l.bias,
                                        p.kind == ST_POOL_MAX, full, full ? bits_out : nullptr, s);
            ctx->blobs[l.top].bits_valid = full && bits_out != nullptr && rc == ST_OK;
            p.pool_mask_valid = rc == ST_OK;
            ++i;                                  // the pooling layer is done
          }
        } else if (tc_usable<T>(ctx->tc, l.tc, l.cin, l.cout)) {
          rc = conv3x3_tc(ctx->tc, l.tc, in, out, nb, hb, wb, l.cin, l.cout, true, l.bias, nullptr,
                          bits_out, nullptr, nullptr, s);
          ctx->blobs[l.top].bits_valid = bits_out != nullptr && rc == ST_OK;
        } else {
          if constexpr (std::is_same<T, __half>::value) {
            set_error("invalid: ST_PREC_FP16 has no SIMT convolution (channels must be multiples of 64)");
            rc = ST_ERR_INVALID;
          } else if constexpr (std::is_same<T, float>::value) {
            if (ctx->precision == ST_PREC_TC32 && l.tc.fwd32 != nullptr)
              rc = conv3x3_tc32(ctx->tc, l.tc, in, out, nb, hb, wb, l.cin, l.cout, true, l.bias,
                                nullptr, nullptr, 1.f, ctx->split_buf, s);
            else
              rc = conv3x3_simt<T>(in, l.w_fwd, l.bias, out, nb, hb, wb, l.cin, l.cout, true, nullptr,
                                   nullptr, s);
          } else {
            rc = conv3x3_simt<T>(in, l.w_fwd, l.bias, out, nb, hb, wb, l.cin, l.cout, true, nullptr,
                                 nullptr, s);
          }
        }
      }
    } else {
      rc = pool_fwd<T>(static_cast<const T*>(ctx->blobs[l.bottom].act), out, nb, hb, wb, l.cin,
                       l.kind == ST_POOL_MAX, s);
    }
    if (rc != ST_OK) return rc;
  }
  return ST_OK;
}

// ---- loss terms -> injected gradients ----------------------------------------------------------------
// Geometry of one batch: nb tiles of h x w pixels cut out of one image.
struct BatchGeom {
  int nb;
  int start_y[kMaxBatch], start_x[kMaxBatch];   // tile origin in the rolled image (`start`, :572)
};

template <typename TA, typename T>
int build_injection(st_ctx* ctx, const st_loss_spec& sp, const Dims& d, const BatchGeom& g,
                    int froll_y, int froll_x, bool is_deepest, cudaStream_t s) {
  BlobRt& b = ctx->blobs[sp.blob];
  b.inj_deferred = false;
  const int nb = g.nb, hf = d.h[sp.blob], wf = d.w[sp.blob], c = b.c;
  const size_t n = (size_t)hf * wf * c;            // elements per tile
  int rc = ensure(ctx, &b.inj, &b.inj_cap, n * nb, ctx->esize);
  if (rc != ST_OK) return rc;
  T* inj = static_cast<T*>(b.inj);
  const TA* f = static_cast<const TA*>(b.act);
  constexpr bool kHalf = std::is_same<TA, __half>::value;
  bool accumulate = false;
  double* stats = ctx->scalars;          // per tile: [0..1] content/dd stats, [2] sum|S|, [3] loss
  double* tile_loss = ctx->scalars + 3;

  if (sp.use_content) {
    ST_REQUIRE(ctx->n_contents > 0, "content layer requested but no content features set");
    for (int ci = 0; ci < ctx->n_contents; ++ci) {
      auto it = ctx->contents.find({ci, sp.blob});
      ST_REQUIRE(it != ctx->contents.end(), "content features missing for a content layer");
      const ContentTarget& t = it->second;
      TargetOffsets offs{};
      for (int i = 0; i < nb; ++i) {
        const int s0y = g.start_y[i] / b.scale, s0x = g.start_x[i] / b.scale;
        // the reference slices [s0 : s0 + hf] out of the full map; a short slice is a shape error
        ST_REQUIRE(s0y + hf <= t.hf && s0x + wf <= t.wf,
                   "tile feature map does not fit into the content feature map at this offset");
        // reduced to [0, Hf) x [0, Wf): the kernels wrap with a single conditional subtract
        offs.ty0[i] = posmod(s0y - floordiv(froll_y, b.scale), t.hf);
        offs.tx0[i] = posmod(s0x - floordiv(froll_x, b.scale), t.wf);
      }
      rc = diff_stats<TA>(f, nb, hf, wf, c, t.nhwc, t.hf, t.wf, offs, stats, kStatStride, ctx->rs, s);
      if (rc == ST_OK)
        rc = diff_inject<TA, T>(f, nb, hf, wf, c, t.nhwc, t.hf, t.wf, offs, stats, kStatStride,
                            sp.content_weight, (double)sp.content_weight, tile_loss, kStatStride,
                            inj, accumulate, s);
      if (rc != ST_OK) return rc;
      accumulate = true;
    }
  }
  if (sp.use_style) {
    ST_REQUIRE(ctx->n_styles > 0, "style layer requested but no style Gram set");
    ST_REQUIRE((size_t)c * c <= 512 * 512, "style layer wider than 512 channels");
    for (int si = 0; si < ctx->n_styles; ++si) {
      auto it = ctx->styles.find({si, sp.blob});
      ST_REQUIRE(it != ctx->styles.end(), "style Gram missing for a style layer");
      const double w = (double)sp.style_weight / ctx->n_styles;
      bool on_tc = false;
      if constexpr (sizeof(TA) == 2) on_tc = gram_tc_ok(ctx->tc, c);
      if (on_tc) {
        if constexpr (sizeof(TA) == 2) {
          size_t cap = ctx->part_floats;
          rc = ensure(ctx, (void**)&ctx->part, &cap, gram_tc_part_floats(ctx->tc, nb, hf * wf, c), 4);
          ctx->part_floats = cap;
          // Gram contraction, then split reduction + delta + loss + max |delta| in one kernel
          int n_part = 0;
          if (rc == ST_OK)
            rc = gram_tc_delta(ctx->tc, f, kHalf, nb, hf * wf, c, ctx->part, it->second, ctx->delta,
                               kHalf ? ctx->delta_max : nullptr, ctx->rs.partials, &n_part, s);
          if (rc == ST_OK)
            rc = delta_pack(ctx->delta, ctx->delta_16, kHalf, ctx->delta_max, ctx->eps_eff, c, nb,
                            ctx->rs.partials, n_part, w, tile_loss, kStatStride, s);
        }
      } else {
        if constexpr (kHalf) {
          set_error("invalid: ST_PREC_FP16 has no SIMT Gram kernel (channels must be 64/128/256/512)");
          return ST_ERR_INVALID;
        } else {
          rc = ST_OK;
          bool gram_done = false;
          if constexpr (std::is_same<TA, float>::value) {
            if (ctx->precision == ST_PREC_TC32 && gram_tc32_ok(ctx->tc, c) && ctx->tc32_tc_gram) {
              // opt-in (ST_TC32_TC_GRAM=1): the Gram contraction on the tensor cores from the [hi | lo]
              // planes of F.  Twice as fast as the SIMT kernel, but G - G_style amplifies the ~1e-6 of
              // accumulation-truncation bias its 32-step chains leave in G (gradient error 1.2e-3 ->
              // 2.4e-3 at one 512^2 tile), so the exact double-accumulating SIMT kernel stays the default
              size_t cap = ctx->part_floats;
              rc = ensure(ctx, (void**)&ctx->part, &cap, gram_tc32_part_floats(nb, hf * wf, c), 4);
              ctx->part_floats = cap;
              if (rc == ST_OK) rc = split_f32(f, ctx->split_buf, (size_t)nb * hf * wf, c, 1.f, s);
              if (rc == ST_OK)
                rc = gram_tc32(ctx->tc, ctx->split_buf, nb, hf * wf, c, ctx->part, ctx->gram, s);
              gram_done = true;
            }
          }
          // the SIMT Gram kernel takes one tile at a time (fp32 mode: nb = 1; tc32 mode: a batch)
          for (int bi = 0; bi < nb && rc == ST_OK && !gram_done; ++bi)
            rc = gram_full<TA>(f + (size_t)bi * n, hf * wf, c, false, ctx->gram + (size_t)bi * c * c,
                               ctx->part, ctx->part_floats, ctx->sm_count, s);
        }
      }
      bool style_tc32 = false;
      if constexpr (std::is_same<TA, float>::value && std::is_same<T, float>::value)
        style_tc32 = ctx->precision == ST_PREC_TC32 && !ctx->tc32_simt_style && ctx->delta_16 != nullptr;
      if (rc == ST_OK && !on_tc)
        rc = gram_delta(ctx->gram, it->second, ctx->delta, nullptr, kHalf, ctx->delta_max,
                        ctx->eps_eff, c, nb, w, tile_loss, kStatStride, ctx->rs, s, style_tc32);
      if (rc != ST_OK) return rc;
      const float* eps_eff = on_tc ? ctx->eps_eff : nullptr;
      // the scale-and-copy pass over S can be skipped when S is this blob's whole injection and a
      // kernel epilogue (not a TMA operand load) consumes it
      const bool defer = on_tc && !sp.use_content && !sp.use_dd && ctx->n_styles == 1 && !is_deepest &&
                         ctx->tc.defer_scale;
      if (on_tc) {
        if constexpr (sizeof(TA) == 2) {
          int per_tile = 0;
          rc = ensure(ctx, (void**)&ctx->abs_partials, &ctx->abs_cap,
                      gemm_abs_partials_needed(nb, hf, wf, c), sizeof(double));
          if (rc == ST_OK && defer && b.inj_scale == nullptr)
            rc = dev_alloc(ctx, (void**)&b.inj_scale, kMaxBatch * sizeof(float));
          if (rc == ST_OK)
            rc = gemm_abs_tc_pair(ctx->tc, f, ctx->delta_16, kHalf,
                                  defer ? inj : static_cast<T*>(ctx->sbuf), nb, hf, wf, c,
                                  ctx->abs_partials, &per_tile, s);
          if (rc == ST_OK)
            rc = sum_partials(ctx->abs_partials, per_tile, nb, stats + 2, kStatStride,
                              defer ? b.inj_scale : nullptr, (float)w, (double)n, eps_eff, s);
          if (rc == ST_OK && defer) {
            b.inj_deferred = true;
            accumulate = true;
            continue;
          }
        }
      } else {
        if constexpr (std::is_same<TA, float>::value && std::is_same<T, float>::value) {
          if (style_tc32 && rc == ST_OK) {
            // tc32: S = F sym(dG) on the tensor cores from the [hi | lo] planes (still in split_buf when
            // the Gram ran there; made here for the 512-channel layers whose Gram is the SIMT one)
            int per_tile = 0;
            if (!(gram_tc32_ok(ctx->tc, c) && ctx->tc32_tc_gram))
              rc = split_f32(f, ctx->split_buf, (size_t)nb * hf * wf, c, 1.f, s);
            if (rc == ST_OK)
              rc = ensure(ctx, (void**)&ctx->abs_partials, &ctx->abs_cap,
                          gemm_abs_partials_needed(nb, hf, wf, c), sizeof(double));
            if (rc == ST_OK)
              rc = gemm_abs_tc32(ctx->tc, ctx->split_buf, ctx->delta, ctx->delta_max, ctx->delta_16,
                                 ctx->eps_eff, static_cast<float*>(ctx->sbuf), nb, hf, wf, c,
                                 ctx->abs_partials, &per_tile, s);
            if (rc == ST_OK)
              rc = sum_partials(ctx->abs_partials, per_tile, nb, stats + 2, kStatStride, nullptr,
                                (float)w, (double)n, nullptr, s);
          }
        }
        if constexpr (std::is_same<TA, T>::value) {
          for (int bi = 0; bi < nb && rc == ST_OK && !style_tc32; ++bi)
            rc = style_grad<T>(f + (size_t)bi * n, ctx->delta + (size_t)bi * c * c,
                               static_cast<T*>(ctx->sbuf) + (size_t)bi * n, hf * wf, c,
                               stats + 2 + (size_t)bi * kStatStride, ctx->rs, s);
        }
      }
      if (rc == ST_OK)
        rc = inject_scaled<T>(inj, static_cast<const T*>(ctx->sbuf), n, nb, (float)w, stats + 2,
                              kStatStride, eps_eff, accumulate, s);
      if (rc != ST_OK) return rc;
      accumulate = true;
    }
  }
  if (sp.use_dd) {
    TargetOffsets offs{};
    rc = diff_stats<TA>(f, nb, hf, wf, c, nullptr, 1, 1, offs, stats, kStatStride, ctx->rs, s);
    if (rc == ST_OK)
      rc = diff_inject<TA, T>(f, nb, hf, wf, c, nullptr, 1, 1, offs, stats, kStatStride, -sp.dd_weight,
                          -(double)sp.dd_weight, tile_loss, kStatStride, inj, accumulate, s);
    if (rc != ST_OK) return rc;
    accumulate = true;
  }
  if (!accumulate) ST_CUDA(cudaMemsetAsync(inj, 0, n * nb * ctx->esize, s));
  return ST_OK;
}

// ---- backward chain to the pixels ----------------------------------------------------------------------
template <typename TA, typename T>
int backward(st_ctx* ctx, const Dims& d, int nb, int deepest_blob, const std::vector<char>& has_inj,
             float* grad, long batch_stride, long plane, long rstride, cudaStream_t s) {
  int cur = deepest_blob, pp = 0;
  const T* g = static_cast<const T*>(ctx->blobs[cur].inj);
  while (true) {
    const LayerRt& l = ctx->layers[ctx->blobs[cur].producer];
    const int b = l.bottom;
    const int hb = d.h[b], wb = d.w[b];
    if (l.kind == ST_CONV3X3 && b == 0) {
      if constexpr (sizeof(T) == 2) {
        if (conv_pix_bwd_tc_ok(ctx->tc, l.tc, l.cout))
          return conv_pix_bwd_tc(ctx->tc, l.tc, g, nb, hb, wb, l.cout, grad, batch_stride, plane,
                                 rstride, s);
        if (ctx->tc.enabled && ctx->tc.pair_kernel && l.tc.bwd != nullptr && l.cout % 64 == 0)
          return conv_last_bwd_tc_pair(ctx->tc, l.tc, g, nb, hb, wb, l.cout, grad, batch_stride,
                                       plane, rstride, s);
      }
      return conv_last_bwd<T>(g, nb, hb, wb, l.cout, l.w_bwd, grad, batch_stride, plane, rstride, s);
    }
    ST_REQUIRE(b != 0, "a pooling layer directly on the image is not supported");
    const BlobRt& bb = ctx->blobs[b];
    const TA* mask = bb.relu ? static_cast<const TA*>(bb.act) : nullptr;
    const T* inj = has_inj[b] ? static_cast<const T*>(bb.inj) : nullptr;
    const float* inj_scale = (has_inj[b] && bb.inj_deferred) ? bb.inj_scale : nullptr;
    T* out = static_cast<T*>(ctx->gbuf[pp]);
    int rc;
    if (l.kind == ST_CONV3X3) {
      if (tc_usable<T>(ctx->tc, l.tc, l.cout, l.cin)) {
        if constexpr (sizeof(T) == 2) {
          // ReLU mask of the bottom blob as bits; made here from the activation when the kernel
          // that produced the blob did not write them
          BlobRt& bm = ctx->blobs[b];
          uint32_t* bits = nullptr;
          rc = ST_OK;
          if (bb.relu && ctx->tc.pair_kernel) {
            if (!bm.bits_valid) {
              rc = ensure(ctx, (void**)&bm.bits, &bm.bits_cap, (size_t)nb * hb * wb * (l.cin / 32), 4);
              if (rc == ST_OK)
                rc = relu_bits_from_act<TA>(static_cast<const TA*>(bm.act), bm.bits,
                                            (size_t)nb * hb * wb, l.cin, s);
              bm.bits_valid = rc == ST_OK;
            }
            bits = bm.bits;
          }
          if (rc == ST_OK)
            rc = conv3x3_tc(ctx->tc, l.tc, g, out, nb, hb, wb, l.cout, l.cin, false, nullptr, mask,
                            bits, inj, inj_scale, s);
        } else {
          rc = ST_ERR_INVALID;
        }
      } else {
        ST_REQUIRE(inj_scale == nullptr, "deferred injection scale needs the tensor-core convolution");
        if constexpr (std::is_same<T, float>::value && std::is_same<TA, float>::value) {
          if (ctx->precision == ST_PREC_TC32 && l.tc.bwd32 != nullptr)
            rc = conv3x3_tc32(ctx->tc, l.tc, g, out, nb, hb, wb, l.cout, l.cin, false, nullptr, mask,
                              inj, ctx->grad_scale, ctx->split_buf, s);
          else
            rc = conv3x3_simt<T>(g, l.w_bwd, nullptr, out, nb, hb, wb, l.cout, l.cin, false, mask, inj,
                                 s);
        } else if constexpr (std::is_same<TA, T>::value) {
          rc = conv3x3_simt<T>(g, l.w_bwd, nullptr, out, nb, hb, wb, l.cout, l.cin, false, mask, inj,
                               s);
        } else {
          set_error("invalid: ST_PREC_FP16 has no SIMT convolution (channels must be multiples of 64)");
          rc = ST_ERR_INVALID;
        }
      }
    } else if (l.pool_mask_valid) {
      rc = pool_bwd_mask<T>(g, l.pool_mask, out, nb, hb, wb, l.cin, l.kind == ST_POOL_MAX, inj,
                            inj_scale, s);
    } else {
      rc = pool_bwd<TA, T>(g, static_cast<const TA*>(bb.act), out, nb, hb, wb, l.cin,
                           l.kind == ST_POOL_MAX, bb.relu, inj, inj_scale, s);
    }
    if (rc != ST_OK) return rc;
    g = out, pp ^= 1, cur = b;
  }
}

// One batch of equally sized tiles: forward, loss terms, backward.  The gradient of tile i goes to
// grad + i * batch_stride.
template <typename TA, typename T>
int eval_batch(st_ctx* ctx, const ImageBatch& view, int h, int w, const BatchGeom& g, int froll_y,
               int froll_x, int n_specs, const st_loss_spec* specs, double* loss_accum, float* grad,
               long batch_stride, long plane, long rstride, cudaStream_t s) {
  ST_REQUIRE(n_specs > 0 && specs != nullptr, "no loss layers");
  ST_REQUIRE(h > 0 && w > 0, "empty tile");
  ST_REQUIRE(g.nb >= 1 && g.nb <= ctx->max_batch, "batch larger than the context allows");
  const int nb = (int)ctx->blobs.size();
  int deepest = -1;
  std::vector<char> has_inj(nb, 0);
  for (int i = 0; i < n_specs; ++i) {
    ST_REQUIRE(specs[i].blob > 0 && specs[i].blob < nb, "loss blob index out of range");
    ST_REQUIRE(!has_inj[specs[i].blob], "duplicate loss blob");
    has_inj[specs[i].blob] = 1;
    deepest = std::max(deepest, specs[i].blob);
  }
  // every loss blob must lie on the backward chain of the deepest one
  {
    std::vector<char> on_chain(nb, 0);
    for (int cur = deepest; cur != 0; cur = ctx->layers[ctx->blobs[cur].producer].bottom)
      on_chain[cur] = 1;
    for (int i = 0; i < n_specs; ++i)
      ST_REQUIRE(on_chain[specs[i].blob], "loss blob is not an ancestor of the deepest loss blob");
  }
  if (ctx->precision == ST_PREC_TC32) {
    // Every injected term is L1-normalised to mean |.| = its weight (num_utils.normalize), so the
    // gradient magnitude is known up front: scale it by the power of two that puts the largest
    // weight at ~2^5 before the fp16 hi/lo split (2^11 of headroom above the mean, lo stays normal)
    float wmax = 0.f;
    for (int i = 0; i < n_specs; ++i) {
      if (specs[i].use_content) wmax = std::max(wmax, std::fabs(specs[i].content_weight) * std::max(ctx->n_contents, 1));
      if (specs[i].use_style) wmax = std::max(wmax, std::fabs(specs[i].style_weight));
      if (specs[i].use_dd) wmax = std::max(wmax, std::fabs(specs[i].dd_weight));
    }
    int e = 0;
    if (wmax > 0.f) std::frexp(wmax, &e);
    ctx->grad_scale = std::ldexp(1.f, std::min(std::max(5 - e, -60), 60));
  }
  const int last_layer = ctx->blobs[deepest].producer;
  const Dims d = blob_dims(ctx, h, w);
  int rc = reserve_for(ctx, d, last_layer, g.nb);
  if (rc == ST_OK) rc = forward<TA>(ctx, view, d, last_layer, has_inj, true, s);
  for (int i = 0; i < n_specs && rc == ST_OK; ++i)
    rc = build_injection<TA, T>(ctx, specs[i], d, g, froll_y, froll_x, specs[i].blob == deepest, s);
  if (rc == ST_OK) rc = loss_finalize(ctx->scalars + 3, kStatStride, g.nb, loss_accum, s);
  if (rc == ST_OK)
    rc = backward<TA, T>(ctx, d, g.nb, deepest, has_inj, grad, batch_stride, plane, rstride, s);
  return rc;
}

int eval_batch_any(st_ctx* ctx, const ImageBatch& view, int h, int w, const BatchGeom& g,
                   int froll_y, int froll_x, int n_specs, const st_loss_spec* specs,
                   double* loss_accum, float* grad, long batch_stride, long plane, long rstride,
                   cudaStream_t s) {
  if (ctx->precision == ST_PREC_FP32 || ctx->precision == ST_PREC_TC32)
    return eval_batch<float, float>(ctx, view, h, w, g, froll_y, froll_x, n_specs, specs, loss_accum,
                                    grad, batch_stride, plane, rstride, s);
  if (ctx->precision == ST_PREC_FP16)
    return eval_batch<__half, __nv_bfloat16>(ctx, view, h, w, g, froll_y, froll_x, n_specs, specs,
                                             loss_accum, grad, batch_stride, plane, rstride, s);
  return eval_batch<__nv_bfloat16, __nv_bfloat16>(ctx, view, h, w, g, froll_y, froll_x, n_specs,
                                                  specs, loss_accum, grad, batch_stride, plane,
                                                  rstride, s);
}

struct Grid {
  int nty, ntx, th, tw, thmax, twmax;
};

Grid tile_grid(int H, int W, int tile_size) {
  Grid g;
  g.nty = (H - 1) / tile_size + 1, g.ntx = (W - 1) / tile_size + 1;
  g.th = H / g.nty, g.tw = W / g.ntx;
  g.thmax = H - (g.nty - 1) * g.th, g.twmax = W - (g.ntx - 1) * g.tw;
  return g;
}

}  // namespace

// =====================================================================================================
// extern "C"
// =====================================================================================================
extern "C" {

const char* st_last_error(void) { return g_error.c_str(); }
int st_version(void) { return 101; }

int st_timing_enable(int on) {
  g_timing_enabled = on != 0;
  return ST_OK;
}

int st_timing_read(int category, double* ms_total, double* work_total, uint64_t* n_scopes) {
  ST_REQUIRE(category >= 0 && category < kTimeCategories, "st_timing_read: unknown category");
  ST_CUDA(cudaDeviceSynchronize());
  double ms = 0.0, work = 0.0;
  uint64_t n = 0;
  for (const TimingRec& r : g_timing_recs) {
    if (r.cat != category) continue;
    float t = 0.f;
    ST_CUDA(cudaEventElapsedTime(&t, r.begin, r.end));
    ms += t, work += r.work, ++n;
  }
  if (ms_total) *ms_total = ms;
  if (work_total) *work_total = work;
  if (n_scopes) *n_scopes = n;
  return ST_OK;
}

int st_timing_reset(void) {
  ST_CUDA(cudaDeviceSynchronize());
  for (const TimingRec& r : g_timing_recs) g_timing_free.push_back(r.begin), g_timing_free.push_back(r.end);
  g_timing_recs.clear();
  return ST_OK;
}
uint64_t st_launch_count(void) { return g_launches.load(); }

int st_create(int device, int precision, int n_layers, const st_layer_desc* layers, st_ctx** out) {
  ST_REQUIRE(out != nullptr && layers != nullptr && n_layers > 0, "st_create: bad arguments");
  ST_REQUIRE(precision == ST_PREC_FP32 || precision == ST_PREC_BF16 || precision == ST_PREC_FP16 ||
                 precision == ST_PREC_TC32,
             "unknown precision");
  int ndev = 0;
  ST_CUDA(cudaGetDeviceCount(&ndev));
  ST_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
  cudaDeviceProp prop;
  ST_CUDA(cudaGetDeviceProperties(&prop, device));
  ST_REQUIRE(prop.major == 10, "libstyle_b200 is built for sm_100a (B200) only");
  st_ctx* ctx = new st_ctx();
  ctx->device = device, ctx->precision = precision, ctx->sm_count = prop.multiProcessorCount;
  ctx->esize = (precision == ST_PREC_FP32 || precision == ST_PREC_TC32) ? 4 : 2;
  ctx->tc32_tc_gram = getenv("ST_TC32_TC_GRAM") != nullptr;
  ctx->tc32_simt_style = getenv("ST_TC32_SIMT_STYLE") != nullptr;
  DeviceGuard guard(device);
  ctx->blobs.resize(n_layers + 1);
  ctx->blobs[0].c = 3;
  ctx->layers.resize(n_layers);
  for (int i = 0; i < n_layers; ++i) {
    const st_layer_desc& ld = layers[i];
    LayerRt& l = ctx->layers[i];
    l.kind = ld.kind, l.bottom = ld.bottom, l.top = i + 1, l.cin = ld.cin, l.cout = ld.cout;
    bool ok = ld.bottom >= 0 && ld.bottom <= i && ld.cin == ctx->blobs[ld.bottom].c &&
              (ld.kind == ST_CONV3X3 || ((ld.kind == ST_POOL_MAX || ld.kind == ST_POOL_AVE) &&
                                         ld.cin == ld.cout));
    if (ok && ld.kind == ST_CONV3X3)
      ok = ld.cout % 64 == 0 && (ld.bottom == 0 ? ld.cin == 3 : ld.cin % 64 == 0);
    if (!ok) {
      delete ctx;
      set_error("invalid: layer " + std::to_string(i) + " is malformed or unsupported");
      return ST_ERR_INVALID;
    }
    BlobRt& t = ctx->blobs[l.top];
    t.c = ld.cout, t.producer = i, t.relu = ld.kind == ST_CONV3X3;
    t.scale = ctx->blobs[ld.bottom].scale * (ld.kind == ST_CONV3X3 ? 1 : 2);
  }
  int rc = ST_OK;
  if (precision != ST_PREC_FP32) rc = tc_init(ctx->tc, ctx->sm_count);
  if (rc == ST_OK && (precision == ST_PREC_FP16 || precision == ST_PREC_TC32) &&
      !(ctx->tc.enabled && ctx->tc.pair_kernel)) {
    set_error("invalid: ST_PREC_FP16 / ST_PREC_TC32 need the tensor-core kernels (unset ST_DISABLE_TC / ST_CONV_V1)");
    rc = ST_ERR_INVALID;
  }
  // tiles of one shape are evaluated as a batch by the tensor-core kernels; the fp32 SIMT parity
  // path keeps the reference's one-tile-at-a-time order
  ctx->max_batch = (precision != ST_PREC_FP32 && ctx->tc.enabled && ctx->tc.pair_kernel) ? kMaxBatch : 1;
  if (const char* e = getenv("ST_MAX_BATCH")) {
    const int v = atoi(e);
    if (v >= 1 && v < ctx->max_batch) ctx->max_batch = v;
  }
  const size_t gram_floats = (size_t)ctx->max_batch * 512 * 512;
  if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->gram, gram_floats * sizeof(float));
  if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->delta, gram_floats * sizeof(float));
  if (rc == ST_OK && precision != ST_PREC_FP32) {
    // 16-bit copy of the delta-Gram: [C][C], or [C][3C] = [Dhi | Dhi | Dlo] in the tc32 mode
    rc = dev_alloc(ctx, (void**)&ctx->delta_16, gram_floats * (precision == ST_PREC_TC32 ? 6 : 2));
    if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->eps_eff, kMaxBatch * sizeof(float));
    if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->delta_max, kMaxBatch * sizeof(unsigned));
  }
  ctx->part_floats = (size_t)16 << 20;
  if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->part, ctx->part_floats * sizeof(float));
  const size_t n_scalars = (size_t)kMaxBatch * kStatStride;
  if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->scalars, n_scalars * sizeof(double));
  if (rc == ST_OK)
    rc = dev_alloc(ctx, (void**)&ctx->rs.partials, (size_t)kMaxReduceBlocks * 4 * sizeof(double));
  if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&ctx->rs.counter, kMaxBatch * sizeof(unsigned));
  if (rc == ST_OK && cudaMemset(ctx->rs.counter, 0, kMaxBatch * sizeof(unsigned)) != cudaSuccess)
    rc = ST_ERR_CUDA;
  if (rc == ST_OK && cudaMemset(ctx->scalars, 0, n_scalars * sizeof(double)) != cudaSuccess)
    rc = ST_ERR_CUDA;
  if (rc != ST_OK) {
    st_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return ST_OK;
}

int st_destroy(st_ctx* ctx) {
  if (!ctx) return ST_OK;
  DeviceGuard guard(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->comm != nullptr) st_comm_destroy(ctx);
  for (LayerRt& l : ctx->layers) {
    cudaFree(l.w_fwd), cudaFree(l.w_bwd), cudaFree(l.bias), cudaFree(l.pool_mask);
    tc_free_weights(l.tc);
  }
  for (BlobRt& b : ctx->blobs) cudaFree(b.act), cudaFree(b.inj), cudaFree(b.inj_scale), cudaFree(b.bits);
  cudaFree(ctx->gbuf[0]), cudaFree(ctx->gbuf[1]), cudaFree(ctx->sbuf), cudaFree(ctx->split_buf);
  cudaFree(ctx->gram), cudaFree(ctx->delta), cudaFree(ctx->part), cudaFree(ctx->scalars);
  cudaFree(ctx->delta_16), cudaFree(ctx->abs_partials), cudaFree(ctx->eps_eff);
  cudaFree(ctx->delta_max);
  cudaFree(ctx->rs.partials), cudaFree(ctx->rs.counter);
  for (auto& kv : ctx->contents) cudaFree(kv.second.nhwc);
  for (auto& kv : ctx->styles) cudaFree(kv.second);
  tc_destroy(ctx->tc);
  delete ctx;
  return ST_OK;
}

int st_set_conv_params(st_ctx* ctx, int layer, const float* w, const float* b) {
  ST_GUARD(ctx);
  ST_REQUIRE(layer >= 0 && layer < (int)ctx->layers.size() && w && b, "bad layer / null weights");
  LayerRt& l = ctx->layers[layer];
  ST_REQUIRE(l.kind == ST_CONV3X3, "st_set_conv_params on a pooling layer");
  const int ci_n = l.cin, co_n = l.cout;
  std::vector<float> fwd((size_t)9 * ci_n * co_n), bwd;
  const bool first = l.bottom == 0;
  bwd.assign(first ? (size_t)9 * co_n * 4 : (size_t)9 * co_n * ci_n, 0.f);
  for (int co = 0; co < co_n; ++co)
    for (int ci = 0; ci < ci_n; ++ci)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const float v = w[(((size_t)co * ci_n + ci) * 3 + ky) * 3 + kx];
          fwd[((size_t)(ky * 3 + kx) * ci_n + ci) * co_n + co] = v;
          const int tapf = (2 - ky) * 3 + (2 - kx);      // flipped tap for the transposed conv
          if (first)
            bwd[((size_t)tapf * co_n + co) * 4 + ci] = v;
          else
            bwd[((size_t)tapf * co_n + co) * ci_n + ci] = v;
        }
  ST_CUDA(cudaDeviceSynchronize());
  if (!l.w_fwd) {
    int rc = dev_alloc(ctx, (void**)&l.w_fwd, fwd.size() * sizeof(float));
    if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&l.w_bwd, bwd.size() * sizeof(float));
    if (rc == ST_OK) rc = dev_alloc(ctx, (void**)&l.bias, co_n * sizeof(float));
    if (rc != ST_OK) return rc;
  }
  ST_CUDA(cudaMemcpy(l.w_fwd, fwd.data(), fwd.size() * sizeof(float), cudaMemcpyHostToDevice));
  ST_CUDA(cudaMemcpy(l.w_bwd, bwd.data(), bwd.size() * sizeof(float), cudaMemcpyHostToDevice));
  ST_CUDA(cudaMemcpy(l.bias, b, co_n * sizeof(float), cudaMemcpyHostToDevice));
  if (ctx->precision == ST_PREC_TC32) {
    if (!first) {
      int rc = tc_pack_split(ctx->tc, l.tc, w, ci_n, co_n);
      if (rc != ST_OK) return rc;
    }
  } else if (ctx->precision != ST_PREC_FP32) {
    const bool half = ctx->precision == ST_PREC_FP16;
    int rc = first ? tc_pack_first(ctx->tc, l.tc, w, co_n)
                   : tc_pack_weights(ctx->tc, l.tc, w, ci_n, co_n, half);
    if (rc == ST_OK && first) rc = tc_pack_first_fwd(ctx->tc, l.tc, w, co_n, half, b);
    if (rc == ST_OK && first) rc = tc_pack_first_rows(ctx->tc, l.tc, w, co_n);
    if (rc != ST_OK) return rc;
  }
  l.has_params = true;
  return ST_OK;
}

int st_reserve(st_ctx* ctx, int max_h, int max_w) {
  ST_GUARD(ctx);
  ST_REQUIRE(max_h > 0 && max_w > 0, "st_reserve: empty tile");
  return reserve_for(ctx, blob_dims(ctx, max_h, max_w), (int)ctx->layers.size() - 1, 1);
}

int st_device_info(st_ctx* ctx, int* sm_count, size_t* workspace_bytes) {
  ST_REQUIRE(ctx != nullptr, "null context");
  if (sm_count) *sm_count = ctx->sm_count;
  if (workspace_bytes) *workspace_bytes = ctx->workspace_bytes;
  return ST_OK;
}

int st_clear_targets(st_ctx* ctx) {
  ST_GUARD(ctx);
  ST_CUDA(cudaDeviceSynchronize());
  for (auto& kv : ctx->contents) {
    ctx->workspace_bytes -= (size_t)kv.second.hf * kv.second.wf * ctx->blobs[kv.first.second].c * 4;
    cudaFree(kv.second.nhwc);
  }
  for (auto& kv : ctx->styles) {
    const size_t c = ctx->blobs[kv.first.second].c;
    ctx->workspace_bytes -= c * c * 4;
    cudaFree(kv.second);
  }
  ctx->contents.clear(), ctx->styles.clear();
  ctx->n_contents = ctx->n_styles = 0;
  return ST_OK;
}

int st_set_style_gram(st_ctx* ctx, int style_index, int blob, const float* gram_dev, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(style_index >= 0 && blob > 0 && blob < (int)ctx->blobs.size() && gram_dev,
             "st_set_style_gram: bad arguments");
  const int c = ctx->blobs[blob].c;
  float*& full = ctx->styles[{style_index, blob}];
  if (!full) {
    int rc = dev_alloc(ctx, (void**)&full, (size_t)c * c * sizeof(float));
    if (rc != ST_OK) return rc;
  }
  ctx->n_styles = std::max(ctx->n_styles, style_index + 1);
  return symmetrize_lower(gram_dev, full, c, (cudaStream_t)stream);
}

int st_set_content_features(st_ctx* ctx, int content_index, int blob, const float* feat_dev, int hf,
                            int wf, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(content_index >= 0 && blob > 0 && blob < (int)ctx->blobs.size() && feat_dev &&
                 hf > 0 && wf > 0,
             "st_set_content_features: bad arguments");
  const int c = ctx->blobs[blob].c;
  ContentTarget& t = ctx->contents[{content_index, blob}];
  if (t.nhwc && (t.hf != hf || t.wf != wf)) {
    ST_CUDA(cudaDeviceSynchronize());
    ctx->workspace_bytes -= (size_t)t.hf * t.wf * c * 4;
    cudaFree(t.nhwc);
    t.nhwc = nullptr;
  }
  if (!t.nhwc) {
    int rc = dev_alloc(ctx, (void**)&t.nhwc, (size_t)hf * wf * c * sizeof(float));
    if (rc != ST_OK) return rc;
  }
  t.hf = hf, t.wf = wf;
  ctx->n_contents = std::max(ctx->n_contents, content_index + 1);
  return nchw_to_nhwc_f32(feat_dev, t.nhwc, hf * wf, c, (cudaStream_t)stream);
}

int st_eval_features_tile(st_ctx* ctx, const float* img_dev, int h, int w, int n_blobs,
                          const int32_t* blob_ids, float* const* out_dev, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(img_dev && h > 0 && w > 0 && n_blobs > 0 && blob_ids && out_dev,
             "st_eval_features_tile: bad arguments");
  int deepest = 0;
  for (int i = 0; i < n_blobs; ++i) {
    ST_REQUIRE(blob_ids[i] > 0 && blob_ids[i] < (int)ctx->blobs.size(), "blob index out of range");
    deepest = std::max(deepest, blob_ids[i]);
  }
  const int last_layer = ctx->blobs[deepest].producer;
  const Dims d = blob_dims(ctx, h, w);
  int rc = reserve_for(ctx, d, last_layer, 1);
  if (rc != ST_OK) return rc;
  ImageBatch view{};
  view.base = img_dev, view.H = h, view.W = w, view.nb = 1;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<char> need_full(ctx->blobs.size(), 0);
  for (int i = 0; i < n_blobs; ++i) need_full[blob_ids[i]] = 1;
  const bool f32 = ctx->precision == ST_PREC_FP32 || ctx->precision == ST_PREC_TC32;
  rc = f32
           ? forward<float>(ctx, view, d, last_layer, need_full, false, s)
           : (ctx->precision == ST_PREC_FP16
                  ? forward<__half>(ctx, view, d, last_layer, need_full, false, s)
                  : forward<__nv_bfloat16>(ctx, view, d, last_layer, need_full, false, s));
  for (int i = 0; i < n_blobs && rc == ST_OK; ++i) {
    const int b = blob_ids[i];
    const int hw = d.h[b] * d.w[b];
    if (f32)
      rc = nhwc_to_nchw_f32<float>((const float*)ctx->blobs[b].act, out_dev[i], hw, ctx->blobs[b].c, s);
    else if (ctx->precision == ST_PREC_FP16)
      rc = nhwc_to_nchw_f32<__half>((const __half*)ctx->blobs[b].act, out_dev[i], hw, ctx->blobs[b].c,
                                    s);
    else
      rc = nhwc_to_nchw_f32<__nv_bfloat16>((const __nv_bfloat16*)ctx->blobs[b].act, out_dev[i], hw,
                                           ctx->blobs[b].c, s);
  }
  return rc;
}

int st_eval_sc_grad_tile(st_ctx* ctx, const float* img_dev, int h, int w, int start_y, int start_x,
                         int feat_roll_y, int feat_roll_x, int n_specs, const st_loss_spec* specs,
                         double* loss_accum_dev, float* grad_dev, long grad_plane_stride,
                         long grad_row_stride, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(img_dev && loss_accum_dev && grad_dev, "st_eval_sc_grad_tile: null pointer");
  ST_REQUIRE(start_y >= 0 && start_x >= 0, "negative tile origin");
  ImageBatch view{};
  view.base = img_dev, view.H = h, view.W = w, view.nb = 1;
  BatchGeom g{};
  g.nb = 1, g.start_y[0] = start_y, g.start_x[0] = start_x;
  return eval_batch_any(ctx, view, h, w, g, feat_roll_y, feat_roll_x, n_specs, specs,
                        loss_accum_dev, grad_dev, 0, grad_plane_stride, grad_row_stride,
                        (cudaStream_t)stream);
}

int st_tile_grid(int H, int W, int tile_size, int* ntiles_y, int* ntiles_x, int* tile_h_max,
                 int* tile_w_max) {
  ST_REQUIRE(H > 0 && W > 0 && tile_size > 0, "st_tile_grid: bad arguments");
  const Grid g = tile_grid(H, W, tile_size);
  if (ntiles_y) *ntiles_y = g.nty;
  if (ntiles_x) *ntiles_x = g.ntx;
  if (tile_h_max) *tile_h_max = g.thmax;
  if (tile_w_max) *tile_w_max = g.twmax;
  return ST_OK;
}

int st_eval_sc_grad_tiles(st_ctx* ctx, const float* img_dev, int H, int W, int roll_y, int roll_x,
                          int tile_size, int rank, int world, int n_specs,
                          const st_loss_spec* specs, double* loss_accum_dev,
                          float* packed_grad_dev, st_stream stream) {
  return st_eval_sc_grad_tile_range(ctx, img_dev, H, W, roll_y, roll_x, tile_size, rank, world, 0, -1,
                                    n_specs, specs, loss_accum_dev, packed_grad_dev, stream);
}

int st_eval_sc_grad_tile_range(st_ctx* ctx, const float* img_dev, int H, int W, int roll_y, int roll_x,
                               int tile_size, int rank, int world, int slot_first, int slot_count,
                               int n_specs, const st_loss_spec* specs, double* loss_accum_dev,
                               float* packed_grad_dev, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(img_dev && loss_accum_dev && packed_grad_dev, "st_eval_sc_grad_tiles: null pointer");
  ST_REQUIRE(slot_first >= 0, "st_eval_sc_grad_tile_range: negative first slot");
  ST_REQUIRE(H > 0 && W > 0 && tile_size > 0 && world > 0 && rank >= 0 && rank < world,
             "st_eval_sc_grad_tiles: bad geometry");
  const Grid g = tile_grid(H, W, tile_size);
  const long plane = (long)g.thmax * g.twmax;
  // This rank's tiles, in slot order.  Tiles of equal shape (all of them unless the grid is ragged)
  // are evaluated together, up to max_batch per launch sequence; slots of one batch must be
  // consecutive so that the gradient of batch tile i lands in slot first + i.
  struct Local {
    int sy, sx, h, w;
  };
  std::vector<Local> tiles;
  for (int t = rank; t < g.nty * g.ntx; t += world) {
    const int ty = t / g.ntx, tx = t % g.ntx;
    const int sy = ty * g.th, sx = tx * g.tw;
    tiles.push_back({sy, sx, ty == g.nty - 1 ? H - sy : g.th, tx == g.ntx - 1 ? W - sx : g.tw});
  }
  const size_t slot_end = slot_count < 0 ? tiles.size()
                                         : std::min(tiles.size(), (size_t)slot_first + (size_t)slot_count);
  for (size_t first = (size_t)slot_first; first < slot_end;) {
    size_t last = first + 1;
    while (last < slot_end && last - first < (size_t)ctx->max_batch &&
           tiles[last].h == tiles[first].h && tiles[last].w == tiles[first].w)
      ++last;
    ImageBatch view{};
    BatchGeom bg{};
    view.base = img_dev, view.H = H, view.W = W, view.nb = bg.nb = (int)(last - first);
    for (size_t i = first; i < last; ++i) {
      // rolled[y][x] = img[(y - roll_y) mod H][(x - roll_x) mod W]  (np.roll, num_utils.py:136-140)
      view.oy[i - first] = tiles[i].sy - roll_y, view.ox[i - first] = tiles[i].sx - roll_x;
      bg.start_y[i - first] = tiles[i].sy, bg.start_x[i - first] = tiles[i].sx;
    }
    int rc = eval_batch_any(ctx, view, tiles[first].h, tiles[first].w, bg, roll_y, roll_x, n_specs,
                            specs, loss_accum_dev, packed_grad_dev + first * 3 * plane, 3 * plane,
                            plane, g.twmax, (cudaStream_t)stream);
    if (rc != ST_OK) return rc;
    first = last;
  }
  return ST_OK;
}

int st_unpack_grad(const float* packed_all_dev, int H, int W, int roll_y, int roll_x, int tile_size,
                   int world, float* grad_dev, double* loss_accum_dev, st_stream stream) {
  ST_REQUIRE(packed_all_dev && grad_dev && H > 0 && W > 0 && tile_size > 0 && world > 0,
             "st_unpack_grad: bad arguments");
  const Grid g = tile_grid(H, W, tile_size);
  const int tpr = (g.nty * g.ntx + world - 1) / world;
  return unpack_grad(packed_all_dev, H, W, roll_y, roll_x, g.nty, g.ntx, g.th, g.tw, g.thmax,
                     g.twmax, world, tpr, grad_dev, loss_accum_dev, (cudaStream_t)stream);
}

size_t st_packed_floats(int H, int W, int tile_size, int world) {
  if (H <= 0 || W <= 0 || tile_size <= 0 || world <= 0) return 0;
  const Grid g = tile_grid(H, W, tile_size);
  return packed_rank_stride((g.nty * g.ntx + world - 1) / world, g.thmax, g.twmax);
}

// ---- the exchange step: NCCL all-gather on the caller's stream -------------------------------------
// libnccl is resolved at run time (the copy the process already holds -- torch's -- or the system
// one): the library has no link-time dependency on it and single-GPU use never touches it.
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
int nccl_api(NcclApi** out) {
  if (!g_nccl.handle) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      set_error(std::string("libnccl.so.2 cannot be loaded: ") + dlerror());
      return ST_ERR_STATE;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
      set_error("libnccl.so.2 lacks an expected symbol");
      return ST_ERR_STATE;
    }
    g_nccl.handle = h;
  }
  *out = &g_nccl;
  return ST_OK;
}
#define ST_NCCL(api, call)                                                                   \
  do {                                                                                       \
    ncclResult_t r_ = (call);                                                                \
    if (r_ != ncclSuccess) {                                                                 \
      set_error(std::string(#call) + ": " +                                                  \
                ((api)->GetErrorString ? (api)->GetErrorString(r_) : "NCCL error"));         \
      return ST_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)
}  // namespace

int st_comm_unique_id(void* id_out) {
  ST_REQUIRE(id_out != nullptr, "st_comm_unique_id: null pointer");
  static_assert(sizeof(ncclUniqueId) == ST_COMM_ID_BYTES, "ncclUniqueId size");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc != ST_OK) return rc;
  ST_NCCL(api, api->GetUniqueId(static_cast<ncclUniqueId*>(id_out)));
  return ST_OK;
}

int st_comm_init(st_ctx* ctx, const void* id, int rank, int world) {
  ST_GUARD(ctx);
  ST_REQUIRE(id != nullptr && world >= 1 && rank >= 0 && rank < world, "st_comm_init: bad arguments");
  ST_REQUIRE(ctx->comm == nullptr, "st_comm_init: the context already has a communicator");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc != ST_OK) return rc;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  ST_NCCL(api, api->CommInitRank(&comm, world, uid, rank));
  ctx->comm = comm, ctx->comm_rank = rank, ctx->comm_world = world;
  return ST_OK;
}

int st_comm_destroy(st_ctx* ctx) {
  ST_GUARD(ctx);
  if (ctx->comm != nullptr) {
    NcclApi* api;
    int rc = nccl_api(&api);
    if (rc != ST_OK) return rc;
    ST_CUDA(cudaDeviceSynchronize());
    ST_NCCL(api, api->CommDestroy(static_cast<ncclComm_t>(ctx->comm)));
    ctx->comm = nullptr;
  }
  return ST_OK;
}

int st_allgather_grad(st_ctx* ctx, const float* packed_local_dev, float* packed_all_dev,
                      size_t floats_per_rank, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(packed_local_dev && packed_all_dev && floats_per_rank > 0, "st_allgather_grad: bad arguments");
  ST_REQUIRE(ctx->comm != nullptr, "st_allgather_grad: no communicator (st_comm_init)");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc != ST_OK) return rc;
  ST_NCCL(api, api->AllGather(packed_local_dev, packed_all_dev, floats_per_rank, ncclFloat,
                              static_cast<ncclComm_t>(ctx->comm), (cudaStream_t)stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ST_OK;
}

int st_gram(st_ctx* ctx, const float* feat_dev, int c, int hw, float* gram_dev, st_stream stream) {
  ST_GUARD(ctx);
  ST_REQUIRE(feat_dev && gram_dev && c > 0 && c <= 512 && hw > 0, "st_gram: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = gram_full<float>(feat_dev, hw, c, true, ctx->gram, ctx->part, ctx->part_floats,
                            ctx->sm_count, s);
  if (rc == ST_OK) rc = extract_lower(ctx->gram, gram_dev, c, s);
  return rc;
}

// The image-level entry points need reduction scratch but no network context; they keep one small
// scratch per device, created on first use.
static ReduceScratch g_scratch[16];
static int global_scratch(ReduceScratch* out) {
  int dev = 0;
  ST_CUDA(cudaGetDevice(&dev));
  ST_REQUIRE(dev < 16, "device index too large");
  if (!g_scratch[dev].partials) {
    ST_CUDA(cudaMalloc((void**)&g_scratch[dev].partials,
                       (size_t)kMaxReduceBlocks * 4 * sizeof(double)));
    ST_CUDA(cudaMalloc((void**)&g_scratch[dev].counter, sizeof(unsigned)));
    ST_CUDA(cudaMemset(g_scratch[dev].counter, 0, sizeof(unsigned)));
  }
  *out = g_scratch[dev];
  return ST_OK;
}

int st_regularizers(const float* img_dev, int H, int W, const float mean[3], float tv_w,
                    float tv_beta, float p_w, float p_pow, const float* aux_dev, float aux_w,
                    int roll_y, int roll_x, double* loss_accum_dev, float* grad_dev,
                    st_stream stream) {
  ST_REQUIRE(img_dev && mean && loss_accum_dev && grad_dev && H > 0 && W > 0,
             "st_regularizers: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  if (rc != ST_OK) return rc;
  return regularizers(img_dev, H, W, mean[0], mean[1], mean[2], tv_w, tv_beta, p_w, p_pow, aux_dev,
                      aux_w, roll_y, roll_x, loss_accum_dev, grad_dev, rs, (cudaStream_t)stream);
}

int st_unpack_regularize(const float* packed_all_dev, const float* img_dev, int H, int W,
                         int roll_y, int roll_x, int tile_size, int world, const float mean[3],
                         float tv_w, float tv_beta, float p_w, float p_pow, const float* aux_dev,
                         float aux_w, double* loss_accum_dev, float* grad_dev, st_stream stream) {
  ST_REQUIRE(packed_all_dev && img_dev && mean && loss_accum_dev && grad_dev && H > 0 && W > 0 &&
                 tile_size > 0 && world > 0,
             "st_unpack_regularize: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  if (rc != ST_OK) return rc;
  const Grid g = tile_grid(H, W, tile_size);
  const int tpr = (g.nty * g.ntx + world - 1) / world;
  return unpack_regularizers(packed_all_dev, H, W, g.nty, g.ntx, g.th, g.tw, g.thmax, g.twmax, world,
                             tpr, img_dev, mean[0], mean[1], mean[2], tv_w, tv_beta, p_w, p_pow,
                             aux_dev, aux_w, roll_y, roll_x, loss_accum_dev, grad_dev, rs,
                             (cudaStream_t)stream);
}

int st_adam_step(float* params, const float* grad, float* g1, float* g2, float* p1, float* avg_out,
                 size_t n, float step_size, float b1, float b2, float bp1, float g1_corr,
                 float g2_corr, float p1_corr, st_stream stream) {
  ST_REQUIRE(params && grad && g1 && g2 && p1 && avg_out && n > 0, "st_adam_step: bad arguments");
  return adam_step(params, grad, g1, g2, p1, avg_out, n, step_size, b1, b2, bp1, g1_corr, g2_corr,
                   p1_corr, (cudaStream_t)stream);
}

namespace {
// Pillow's precompute_coeffs (libImaging/Resample.c) for the full source range, in double.
struct ResampleTable {
  int ksize = 0;
  std::vector<int> bounds;      // [out][2] = first source index, count
  std::vector<double> kk;       // [out][ksize], normalised
};
ResampleTable resample_table(int in_size, int out_size, int method) {
  const double kPi = 3.14159265358979323846;
  auto sinc = [&](double t) {
    if (t == 0.0) return 1.0;
    t *= kPi;
    return sin(t) / t;
  };
  auto filt = [&](double x) {
    if (method == 0) return (-3.0 <= x && x < 3.0) ? sinc(x) * sinc(x / 3) : 0.0;
    if (x < 0.0) x = -x;
    return x < 1.0 ? 1.0 - x : 0.0;
  };
  const double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = (method == 0 ? 3.0 : 1.0) * filterscale, ss = 1.0 / filterscale;
  ResampleTable t;
  t.ksize = (int)ceil(support) * 2 + 1;
  t.bounds.assign((size_t)out_size * 2, 0);
  t.kk.assign((size_t)out_size * t.ksize, 0.0);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double* k = &t.kk[(size_t)xx * t.ksize];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = filt((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    if (ww != 0.0)
      for (int x = 0; x < xmax; ++x) k[x] /= ww;
    t.bounds[2 * xx] = xmin, t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

int run_resample(const float* in, float* out, int channels, int in_h, int in_w, int out_size,
                 bool along_y, int method, cudaStream_t s) {
  const ResampleTable t = resample_table(along_y ? in_h : in_w, out_size, method);
  int* bounds_dev = nullptr;
  double* kk_dev = nullptr;
  ST_CUDA(cudaMallocAsync((void**)&bounds_dev, t.bounds.size() * sizeof(int), s));
  ST_CUDA(cudaMallocAsync((void**)&kk_dev, t.kk.size() * sizeof(double), s));
  // pageable sources: the copies are staged before the call returns, so the vectors may die here
  ST_CUDA(cudaMemcpyAsync(bounds_dev, t.bounds.data(), t.bounds.size() * sizeof(int),
                          cudaMemcpyHostToDevice, s));
  ST_CUDA(cudaMemcpyAsync(kk_dev, t.kk.data(), t.kk.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  int rc = resample_pass(in, out, channels, in_h, in_w, out_size, along_y, bounds_dev, kk_dev, t.ksize, s);
  ST_CUDA(cudaFreeAsync(bounds_dev, s));
  ST_CUDA(cudaFreeAsync(kk_dev, s));
  return rc;
}
}  // namespace

int st_resample_coeffs(int in_size, int out_size, int method, int* ksize, int* bounds_out,
                       double* kk_out) {
  ST_REQUIRE(in_size > 0 && out_size > 0 && (method == 0 || method == 1) && ksize,
             "st_resample_coeffs: bad arguments");
  const ResampleTable t = resample_table(in_size, out_size, method);
  *ksize = t.ksize;
  if (bounds_out) std::memcpy(bounds_out, t.bounds.data(), t.bounds.size() * sizeof(int));
  if (kk_out) std::memcpy(kk_out, t.kk.data(), t.kk.size() * sizeof(double));
  return ST_OK;
}

int st_resize_f32(const float* in_dev, int channels, int h, int w, int out_h, int out_w, int method,
                  float* out_dev, float* tmp_dev, st_stream stream) {
  ST_REQUIRE(in_dev && out_dev && channels > 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0,
             "st_resize_f32: bad arguments");
  ST_REQUIRE(method == 0 || method == 1, "st_resize_f32: method must be 0 (Lanczos) or 1 (bilinear)");
  cudaStream_t s = (cudaStream_t)stream;
  const bool need_x = out_w != w, need_y = out_h != h;
  if (!need_x && !need_y) {
    ST_CUDA(cudaMemcpyAsync(out_dev, in_dev, (size_t)channels * h * w * sizeof(float),
                            cudaMemcpyDeviceToDevice, s));
    return ST_OK;
  }
  ST_REQUIRE(!(need_x && need_y) || tmp_dev != nullptr, "st_resize_f32: tmp_dev needed for two passes");
  int rc = ST_OK;
  const float* src = in_dev;
  if (need_x) {                                         // horizontal first, like ImagingResample
    float* dst = need_y ? tmp_dev : out_dev;
    rc = run_resample(src, dst, channels, h, w, out_w, false, method, s);
    src = dst;
  }
  if (rc == ST_OK && need_y) rc = run_resample(src, out_dev, channels, h, out_w, out_h, true, method, s);
  return rc;
}

int st_iter_stats(const float* avg_dev, float* old_dev, int H, int W, double* stats_dev,
                  st_stream stream) {
  ST_REQUIRE(avg_dev && old_dev && stats_dev && H > 0 && W > 0, "st_iter_stats: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  return rc != ST_OK ? rc : iter_stats(avg_dev, old_dev, H, W, stats_dev, rs, (cudaStream_t)stream);
}

int st_output_step(const float* avg_dev, float* old_dev, int H, int W, const float mean[3], int bgr,
                   double* stats_dev, uint8_t* pic_dev, st_stream stream) {
  ST_REQUIRE(avg_dev && old_dev && mean && stats_dev && H > 0 && W > 0, "st_output_step: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  return rc != ST_OK ? rc
                     : output_step(avg_dev, old_dev, H, W, mean[0], mean[1], mean[2], bgr != 0,
                                   stats_dev, pic_dev, rs, (cudaStream_t)stream);
}

int st_get_image_u8(const float* params_dev, int H, int W, const float mean[3], int bgr,
                    uint8_t* out_dev, st_stream stream) {
  ST_REQUIRE(params_dev && mean && out_dev && H > 0 && W > 0, "st_get_image_u8: bad arguments");
  return get_image_u8(params_dev, H, W, mean[0], mean[1], mean[2], bgr != 0, out_dev,
                      (cudaStream_t)stream);
}

int st_dot(const float* x, const float* y, size_t n, double* out_dev, st_stream stream) {
  ST_REQUIRE(x && y && out_dev && n > 0, "st_dot: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  return rc != ST_OK ? rc : dot_to(x, y, n, out_dev, rs, (cudaStream_t)stream);
}

int st_asum(const float* x, size_t n, double* out_dev, st_stream stream) {
  ST_REQUIRE(x && out_dev && n > 0, "st_asum: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  return rc != ST_OK ? rc : asum_to(x, n, out_dev, rs, (cudaStream_t)stream);
}

int st_axpby(float a, const float* x, float b, float* y, size_t n, st_stream stream) {
  ST_REQUIRE(x && y && n > 0, "st_axpby: bad arguments");
  return axpby(a, x, b, y, n, (cudaStream_t)stream);
}

int st_lbfgs_inv_hv(const float* grad_dev, size_t n, int m, const float* const* s_dev,
                    const float* const* y_dev, const double* sy_host, float* p_dev,
                    double* scratch_dev, st_stream stream) {
  ST_REQUIRE(grad_dev && p_dev && scratch_dev && n > 0 && m >= 0 && m <= 16,
             "st_lbfgs_inv_hv: bad arguments");
  ST_REQUIRE(m == 0 || (s_dev && y_dev && sy_host), "st_lbfgs_inv_hv: missing curvature pairs");
  cudaStream_t s = (cudaStream_t)stream;
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  if (rc != ST_OK) return rc;
  ST_CUDA(cudaMemcpyAsync(p_dev, grad_dev, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  double* alpha = scratch_dev;            // [0..15]
  double* tmp = scratch_dev + 16;         // [16], [17]
  for (int k = m - 1; k >= 0 && rc == ST_OK; --k) {       // optimizers.py:109-111
    rc = dot_to(s_dev[k], p_dev, n, tmp, rs, s);
    if (rc == ST_OK) rc = axpy_dev(y_dev[k], p_dev, n, tmp, sy_host[k], nullptr, -1.0, alpha + k, s);
  }
  if (m > 0 && rc == ST_OK) {                               // :113-115
    rc = dot_to(y_dev[m - 1], y_dev[m - 1], n, tmp, rs, s);
    if (rc == ST_OK) rc = scale_dev(p_dev, n, sy_host[m - 1], tmp, s);
  }
  for (int k = 0; k < m && rc == ST_OK; ++k) {              // :117-119
    rc = dot_to(y_dev[k], p_dev, n, tmp, rs, s);
    // p += (alpha_k - beta) * s_k,  beta = tmp / sy_k
    if (rc == ST_OK) rc = axpy_dev(s_dev[k], p_dev, n, tmp, sy_host[k], alpha + k, -1.0, nullptr, s);
  }
  return rc;
}

int st_lbfgs_step(const float* grad_dev, size_t n, int n_corr, float* ring_s_dev,
                  const float* ring_y_dev, double* state_dev, float* scratch_dev, float* params_dev,
                  float initial_step, st_stream stream) {
  ST_REQUIRE(grad_dev && ring_s_dev && ring_y_dev && state_dev && scratch_dev && params_dev && n > 0 &&
                 n_corr >= 1 && n_corr <= 16,
             "st_lbfgs_step: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  if (rc != ST_OK) return rc;
  return lbfgs_direction(grad_dev, n, n_corr, ring_s_dev, ring_y_dev, state_dev, scratch_dev,
                         params_dev, initial_step, rs, (cudaStream_t)stream);
}

int st_lbfgs_commit(const float* grad_new_dev, const float* grad_old_dev, size_t n, int n_corr,
                    const float* ring_s_dev, float* ring_y_dev, double* state_dev, st_stream stream) {
  ST_REQUIRE(grad_new_dev && grad_old_dev && ring_s_dev && ring_y_dev && state_dev && n > 0 &&
                 n_corr >= 1 && n_corr <= 16,
             "st_lbfgs_commit: bad arguments");
  ReduceScratch rs;
  int rc = global_scratch(&rs);
  if (rc != ST_OK) return rc;
  return lbfgs_commit(grad_new_dev, grad_old_dev, n, n_corr, ring_s_dev, ring_y_dev, state_dev, rs,
                      (cudaStream_t)stream);
}

}  // extern "C"
