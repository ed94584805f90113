// tcgen05 / TMA implicit-GEMM convolution for bf16 NHWC activations, CTA-pair version (sm_100a).
//
//   out[p][n] = epilogue( sum_{tap, c} in[p + off(tap)][c] * Wk[n][tap*Cin + c] )      TAPS = 9 (3x3)
//   out[p][n] = epilogue( sum_c in[p][c] * Wk[n][c] )                                  TAPS = 1 (1x1)
//
// Why this shape.  A 128x128 single-CTA tile that re-loads its A operand for every tap needs 128
// bytes of L2->shared traffic per tensor-core cycle and SM; 148 SMs then ask the L2 for three times
// what it delivers (profiles/r01_conv_v1_l2_bound.md) and the tensor pipe idles.  This kernel cuts
// that traffic ~3.5x:
//   * two CTAs of a cluster form one tcgen05 `cta_group::2` MMA: 256 pixels x BN channels per
//     instruction, each CTA staging only its own 128 pixels of A and HALF of the weight tile B;
//   * the A operand of one 64-channel block is loaded ONCE per tile: the (16+2) x (8+2) pixel halo
//     window, one TMA box, pixel (y, x) of the window at byte (y*10 + x)*128 of the stage.  The nine
//     taps are nine shared-memory descriptors into that one copy: start address shifted by
//     (dy*10 + dx)*128 bytes, 8-row groups 10*128 = 1280 bytes apart (the stride-byte-offset of the
//     descriptor).  Neither is a multiple of the 1024-byte swizzle period, and that is fine: the
//     SWIZZLE_128B XOR is a function of the absolute shared-memory address bits for TMA and for
//     tcgen05.mma alike (descriptor base-offset field left 0; measured on B200 with
//     tools/experiments/desc_offset.cu, profiles/r01_desc_offset.md).  A moves 18*10/(16*8) = 1.4
//     times per 64-channel block instead of 9 (one load per tap) or 3.4 (three x-shifted copies, the
//     previous version of this kernel).
//
// Work split: pair tile = 32 rows x 8 columns of pixels (CTA rank r owns rows [16r, 16r+16)) x BN
// output channels; persistent pairs walk the tile list round-robin.  K loop: 64-channel blocks,
// inside each the 9 taps, inside each 4 MMAs of K = 16.
//
// Warps (384 threads per CTA):
//   0      A producer   (TMA, one lane)            both CTAs
//   1      MMA issuer   (one lane, leader CTA only) + TMEM allocation (both CTAs)
//   2      B producer   (TMA, one lane)            both CTAs
//   11     operand producer of the backward epilogue (TMA, one lane): the injected loss gradient
//          of each (tile, 64-channel group) lands in a two-stage shared-memory ring
//   3..10  epilogue     TMEM -> registers -> bias/ReLU | mask/inject | abs-sum -> swizzled smem
//                       -> TMA store; accumulators are double-buffered in TMEM.  Two warps per TMEM
//                       lane quadrant (warp % 4), each taking one 32-channel half of every
//                       64-channel group: with four warps the BN = 64 layers (36 MMAs of 32 cycles
//                       per tile) were paced by their epilogue, not by the tensor pipe
// All "full" barriers live in the leader CTA (TMA of the peer signals them through the cluster
// address); "empty" barriers are signalled in both CTAs by multicast tcgen05.commit.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"

namespace st {

namespace {

constexpr int kBH = 16, kBW = 8;          // pixels per CTA tile: 16 rows x 8 columns = 128 = TMEM lanes
constexpr int kThreads2 = 384;
constexpr int kEpiWarp0 = 3;             // first of the eight epilogue warps
constexpr int kInjWarp = 11, kSE = 2;    // operand producer warp of the backward epilogue, its stages
constexpr uint32_t kSpin = 1u << 24;
constexpr int kOutStageBytes = 128 * 128; // 128 pixels x 64 channels bf16
constexpr int kTailBytes = 3072;          // barriers (256 B) + tmem slot + bias copy (2 KB)
constexpr int kSmemBudget = 227 * 1024 - 1024 /*alignment slack*/ - kTailBytes;

// kEpiFwdPool = kEpiFwd + the 2x2/2 pooling layer that follows: the pooled map and a one-byte
// arg-max / ReLU mask per pooled element are produced from the staged output tile
// kEpiFwd32 / kEpiBwd32: the split-operand mode (ST_PREC_TC32).  The A operand is an fp16 [hi | lo]
// pair of channel planes of an fp32 activation, the weights are packed [Whi | Whi | Wlo] per tap, so
// the K loop accumulates a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in the fp32 accumulator (three MMAs per
// product, ~22 significant bits per operand); the epilogue works in fp32 and stores fp32 straight
// from registers (each thread owns 32 consecutive channels of one pixel = one 128-byte line).
enum Epilogue { kEpiFwd = 0, kEpiBwd = 1, kEpiAbs = 2, kEpiPix = 3, kEpiFwdPool = 4, kEpiFwd32 = 5,
                kEpiBwd32 = 6, kEpiAbs32 = 7 };   // kEpiAbs32: the style GEMM of the split-operand mode
constexpr int kPoolStageBytes = 32 * 128;   // 8 x 4 pooled pixels x 64 channels bf16

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// arrive on a barrier of any CTA of the cluster (cluster address).  Default semantics
// (release at CTA scope): the explicit .release.cluster form costs a MEMBAR.ALL.GPU per arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// One lane of a converged warp.  The role loops below run warp-wide with warp-uniform state and
// only the TMA / MMA / commit instructions sit under elect_one(): that keeps descriptors and
// coordinates in uniform registers (a lane-0-only loop makes ptxas re-broadcast every operand
// with R2UR before each UTCHMMA / UTMALDG: ~190 cycles per MMA, profiles/r01_conv_v2_issue.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
    if (spin > kSpin) __trap();           // protocol bug: fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA loads issued by either CTA of the pair; completion bytes go to the barrier at `bar_cluster`
// (a cluster address, the leader's barrier).
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t bar_cluster,
                                                 void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar_cluster,
                                                 void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// plain (single-CTA) load: data and completion bytes stay in this CTA
__device__ __forceinline__ void tma_load_4d_local(const CUtensorMap* map, uint64_t* bar, void* dst,
                                                  int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Arrives (once the MMAs issued so far retire) on the barrier at this smem offset in BOTH CTAs.
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 128-byte rows, 8-row groups `sbo` bytes
// apart (1024 for a dense tile; 1280 for the 10-pixel-wide halo window, see the file comment).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo = 1024) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) |
         ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// 2x2/2 max pooling of one (pooled pixel, 8-channel) item straight on the packed 16-bit values of the
// staged tile: per window position one HSET2 (mask of "greater than the running maximum": strict,
// so the first maximum in scan order wins like Caffe's), one HMNMX2 and one LOP3 per channel pair.
// The values were rounded when they were staged, so this equals comparing their fp32 images.
// Returns the four packed maxima in ow[] and the eight mask bytes (bits 0-1 = arg-max position,
// bit 2 = maximum > 0) in mk[].
template <bool HALF>
__device__ __forceinline__ void pool_max_item(const uint8_t* stage_out, int pr, int pc, int j,
                                              int rows_left, int cols_left, uint32_t (&ow)[4],
                                              uint32_t (&mk)[2]) {
  const uint32_t neg = HALF ? 0xFBFFFBFFu : 0xFF7FFF7Fu;      // most negative finite value, twice
  uint32_t best[4] = {neg, neg, neg, neg}, code[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int r = 2 * pr + (d >> 1), cx = 2 * pc + (d & 1);
    if (r >= rows_left || cx >= cols_left) continue;          // ceil mode: clipped window
    const int mm = r * 8 + cx;
    const uint4 raw = *reinterpret_cast<const uint4*>(stage_out + mm * 128 + ((j ^ (mm & 7)) << 4));
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t gt = gt2_mask<HALF>(w4[e], best[e]);
      best[e] = max2<HALF>(w4[e], best[e]);
      code[e] = (code[e] & ~gt) | ((uint32_t)(d * 0x00010001) & gt);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    code[e] |= gt2_mask<HALF>(best[e], 0u) & 0x00040004u;
    ow[e] = best[e];
  }
  // one byte per element: bytes 0 and 2 of each code word
  mk[0] = __byte_perm(code[0], code[1], 0x6420);
  mk[1] = __byte_perm(code[2], code[3], 0x6420);
}

struct Tc2Args {
  int nb, h, w, cin, cout;         // nb tiles of the batch, each [h][w]
  int in_half, out_half;           // operands (A and B) / stored output are fp16 instead of bf16
  int w_batched;                   // the B operand has one matrix per batch tile (style GEMM)
  int cin_map;                     // channels of the input tensor (0: = cin).  Split mode: 2 * real cin
  int a_wrap;                      // split mode: A channel block of K block cb is cb - a_wrap once
                                   // cb >= a_wrap (the third K segment re-reads the hi planes); else 1 << 30
  float out_scale;                 // split mode: accumulator factor (1 / (weight scale * input scale))
  const float* out_scale_tile;     // kEpiAbs32: one more factor per batch tile (1 / sigma of its delta-Gram)
  const float* mask_f32;           // kEpiBwd32: fp32 activation whose sign is the ReLU mask, may be null
  const float* inj_f32;            // kEpiBwd32: fp32 injected gradient, may be null
  float* out_f32;                  // kEpiFwd32 / kEpiBwd32: fp32 NHWC output
  int resb_bytes;                  // RESB: bytes of the resident weight block (multiple of 1024)
  int tiles_x, tiles_y, tiles_n;   // pair tiles per batch tile: 8 columns x 32 rows x BN channels
  FastDiv div_x, div_y, div_n;     // by tiles_x / tiles_y / tiles_n (decode_tile runs once per tile
                                   // in every warp role: three hardware divisions were ~100
                                   // instructions of the ~700 an epilogue warp spent per tile)
  const float* bias;               // kEpiFwd
  const uint32_t* mask_bits;       // kEpiBwd: ReLU bit mask of the output blob (see relu_bit), may be null
  uint32_t* bits_out;              // forward: ReLU bit mask of the output, [pixel][cout/32], may be null
  const __nv_bfloat16* inj;        // kEpiBwd, may be null
  const float* inj_scale;          // kEpiBwd: per batch tile factor applied to inj, may be null
  double* abs_partials;            // kEpiAbs: [pair tile][cta rank][epilogue warp]
  int pool_mode;                   // kEpiFwdPool: 1 = max, 2 = average
  int write_full;                  // kEpiFwdPool: also store the un-pooled output
  uint8_t* pool_mask;              // kEpiFwdPool: [nb][ho][wo][cout] bytes (see pool_mask_byte)
  float* pix;                      // kEpiPix: planar f32 output, channels 0..2 of the accumulator
  long pix_batch, pix_plane, pix_row;   // strides (floats) between batch tiles / planes / rows
};

struct TileCoord {
  int b, x0, y0, n_tile;
};
__device__ __forceinline__ TileCoord decode_tile(const Tc2Args& a, int tile, int rank) {
  TileCoord t;
  int m = (int)a.div_n.div((unsigned)tile);
  t.n_tile = tile - m * a.tiles_n;
  int r = (int)a.div_x.div((unsigned)m);
  t.x0 = (m - r * a.tiles_x) * kBW;
  t.b = (int)a.div_y.div((unsigned)r);
  t.y0 = (r - t.b * a.tiles_y) * (2 * kBH) + rank * kBH;
  return t;
}

// RESB: the whole weight matrix of this CTA (its BN/2 rows x all K) stays resident in shared
// memory for the life of the kernel -- for the small layers (conv1_2, conv2_1, the 3-channel
// backward) where one TMA round trip per tap costs more than the MMAs of that tap.
template <int BN, int TAPS, bool RESB = false, bool POOL = false, bool BWD = false>
struct Cfg2 {
  static constexpr int kHaloRows = TAPS == 9 ? kBH + 2 : kBH;
  static constexpr int kWinW = TAPS == 9 ? kBW + 2 : kBW;        // pixels per window row
  static constexpr int kAPitch = kWinW * 128;                    // bytes between 8-pixel row groups
  static constexpr int kALoadBytes = kHaloRows * kAPitch;        // 22.5 KB (3x3) / 16 KB (1x1)
  static constexpr int kABytes = (kALoadBytes + 1023) / 1024 * 1024;   // stages stay 1024-aligned
  static constexpr int kBBytes = (BN / 2) * 128;                 // this CTA's half of one tap's B tile
  // taps per B pipeline stage: the narrow tiles have little tensor time per tap (4 MMAs of BN/2
  // cycles), so one barrier round trip per TAP made their MMA warp issue-bound; one per kernel row
  // (three taps) amortises it.  BN = 256 keeps one tap per stage (48 KB stages would not fit).
  static constexpr int kTB = (TAPS == 9 && BN <= 128) ? 3 : 1;
  static constexpr int kBStage = kTB * kBBytes;
  // A stages: one stage is nine taps of tensor time (9 x 4 MMAs of BN/2 cycles), so the wide tile
  // needs fewer of them; the backward kernels give one up for the injected-gradient ring
  static constexpr int kSA = (BN == 256 || BWD) ? 3 : 4;
  static constexpr int kInjBytes = BWD ? kSE * kOutStageBytes : 0;
  static constexpr int kPoolBytes = POOL ? 2 * kPoolStageBytes : 0;
  static constexpr int kOutBytes = 2 * kOutStageBytes + kPoolBytes + kInjBytes;
  static constexpr int kSBRaw = (kSmemBudget - kSA * kABytes - kOutBytes) / kBStage;
  static constexpr int kSB = RESB ? 1 : (kSBRaw > 8 ? 8 : kSBRaw);
  static constexpr int kResMax = kSmemBudget - kSA * kABytes - kOutBytes;   // bytes for resident B
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;    // double-buffered accumulator
  static constexpr int kFixedBytes = kSA * kABytes + kOutBytes + 1024 + kTailBytes;
  static constexpr int kSmemBytes = kFixedBytes + kSB * kBStage;           // staged-B variant
  // instruction descriptor without the operand formats (0 = f16, 1 = bf16 at bits 7 and 10)
  static constexpr uint32_t kIdescBase = (1u << 4) | ((uint32_t)(BN >> 3) << 17) |
                                         ((uint32_t)(256 >> 4) << 24);
  static_assert(RESB || kSB >= 3, "not enough shared memory for the weight pipeline");
};

template <int BN, int TAPS, int EPI, bool RESB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_aux, const Tc2Args a) {
  // map_aux: the pooled output (kEpiFwdPool) or the injected gradient (kEpiBwd, same geometry as out)
  constexpr bool kPool = EPI == kEpiFwdPool;
  constexpr bool kFwd = EPI == kEpiFwd || EPI == kEpiFwdPool;
  constexpr bool k32 = EPI == kEpiFwd32 || EPI == kEpiBwd32 || EPI == kEpiAbs32;
  using Cfg = Cfg2<BN, TAPS, RESB, kPool, EPI == kEpiBwd>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* out_base = a_base + Cfg::kSA * Cfg::kABytes;
  uint8_t* b_base = out_base + Cfg::kOutBytes;
  const int b_bytes = RESB ? a.resb_bytes : Cfg::kSB * Cfg::kBStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::kSA;
  uint64_t* b_full = a_empty + Cfg::kSA;
  uint64_t* b_empty = b_full + Cfg::kSB;
  // accumulator buffers in TMEM: two (one tile each); the split-operand mode cycles four, one per
  // accumulation CHAIN (see the k32 epilogue)
  constexpr int kNT = k32 ? 4 : 2;
  constexpr int kTmemCols = k32 ? (4 * BN < 32 ? 32 : 4 * BN) : Cfg::kTmemCols;
  static_assert(!k32 || (BN <= 128 && !RESB), "split-operand mode: BN <= 128, staged weights");
  uint64_t* t_full = b_empty + Cfg::kSB;
  uint64_t* t_empty = t_full + kNT;
  uint64_t* e_full = t_empty + kNT;        // injected-gradient ring (backward epilogue)
  uint64_t* e_empty = e_full + kSE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_empty + kSE);
  uint8_t* inj_base = out_base + 2 * kOutStageBytes;     // kEpiBwd only (no pooling stages there)
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [cout] <= 512

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int kb_per_tap = a.cin >> 6;
  const int num_tiles = a.tiles_x * a.tiles_y * a.tiles_n * a.nb;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_in), prefetch_tmap(&map_w), prefetch_tmap(&map_out);
    for (int i = 0; i < Cfg::kSA; ++i) mbar_init(&a_full[i], 1), mbar_init(&a_empty[i], 1);
    for (int i = 0; i < Cfg::kSB; ++i) mbar_init(&b_full[i], 1), mbar_init(&b_empty[i], 1);
    for (int i = 0; i < kNT; ++i) mbar_init(&t_full[i], 1), mbar_init(&t_empty[i], 16);
    for (int i = 0; i < kSE; ++i) mbar_init(&e_full[i], 1), mbar_init(&e_empty[i], 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  if constexpr (kFwd || EPI == kEpiFwd32) {
    for (int i = threadIdx.x; i < a.cout; i += kThreads2) bias_s[i] = a.bias[i];
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, the bias
  // copy -- weights, never written by a predecessor) may run while the previous kernel of the
  // stream drains; activations are only touched below.  The next kernel may be scheduled as soon as
  // SMs free up (it blocks at its own griddepcontrol.wait until this grid has completed).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================================== A producer ============================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const TileCoord t = decode_tile(a, tile, (int)rank);
      for (int cb = 0; cb < kb_per_tap; ++cb) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        const uint32_t bar = map_to_cta(smem_u32(&a_full[stage]), 0);
        uint8_t* dst = a_base + stage * Cfg::kABytes;
        const int ca = (cb >= a.a_wrap ? cb - a.a_wrap : cb) * 64;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&a_full[stage], 2 * Cfg::kALoadBytes);
          if constexpr (TAPS == 9)
            tma_load_4d_pair(&map_in, bar, dst, ca, t.x0 - 1, t.y0 - 1, t.b);
          else
            tma_load_4d_pair(&map_in, bar, dst, ca, t.x0, t.y0, t.b);
        }
        __syncwarp();
        if (++stage == Cfg::kSA) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 2) {
    // ===================================== B producer ============================================
    if constexpr (RESB) {
      // one shot: every (channel block, tap) slab of this CTA's weight rows, one barrier
      const uint32_t bar = map_to_cta(smem_u32(&b_full[0]), 0);
      if (elect_one()) {
        if (leader) mbar_expect_tx(&b_full[0], 2 * a.resb_bytes);
        for (int cb = 0; cb < kb_per_tap; ++cb)
          for (int tap = 0; tap < TAPS; ++tap)
            tma_load_3d_pair(&map_w, bar, b_base + (cb * TAPS + tap) * Cfg::kBBytes,
                             tap * a.cin + cb * 64, (int)rank * (BN / 2), 0);
      }
      __syncwarp();
    } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const TileCoord t = decode_tile(a, tile, (int)rank);
        const int n0 = t.n_tile * BN + (int)rank * (BN / 2);
        const int wb = a.w_batched ? t.b : 0;
        for (int cb = 0; cb < kb_per_tap; ++cb) {
          for (int tap = 0; tap < TAPS; tap += Cfg::kTB) {
            mbar_wait(&b_empty[stage], phase ^ 1);
            const uint32_t bar = map_to_cta(smem_u32(&b_full[stage]), 0);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&b_full[stage], 2 * Cfg::kBStage);
#pragma unroll
              for (int j = 0; j < Cfg::kTB; ++j)
                tma_load_3d_pair(&map_w, bar, b_base + stage * Cfg::kBStage + j * Cfg::kBBytes,
                                 (tap + j) * a.cin + cb * 64, n0, wb);
            }
            __syncwarp();
            if (++stage == Cfg::kSB) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader) ===================================
    if (leader) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, it = 0;
      const uint32_t fmt = a.in_half ? 0u : 1u;
      const uint32_t idesc = Cfg::kIdescBase | (fmt << 7) | (fmt << 10);
      if constexpr (RESB) mbar_wait(&b_full[0], 0);      // the resident weights have landed
      for (int tile = pair; tile < num_tiles; tile += num_pairs, it += k32 ? 0 : 1) {
        uint32_t buf = it & 1, use = it >> 1;
        uint32_t d_tmem = tmem_base + buf * BN;
        if constexpr (!k32) {
          mbar_wait(&t_empty[buf], (use & 1) ^ 1);
          tc_fence_after();
        }
        for (int cb = 0; cb < kb_per_tap; ++cb) {
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          // Descriptors are built once per 64-channel block; the taps and the four K = 16 steps are
          // compile-time offsets of their 14-bit address fields (the whole block is unrolled).  A
          // rolled tap loop computed every descriptor from scratch: ~65 instructions (~450 cycles)
          // per tap in this one warp, against 128 cycles of tensor time per tap at BN = 64 -- the
          // small layers were bound by MMA *issue* (profiles/r01_conv_small_ncu.md).
          const uint64_t da0 = make_smem_desc(smem_u32(a_base + sa * Cfg::kABytes), Cfg::kAPitch);
          const bool last_cb = cb == kb_per_tap - 1;
          if constexpr (RESB) {
            const uint64_t db0 = make_smem_desc(smem_u32(b_base + cb * TAPS * Cfg::kBBytes));
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < TAPS; ++tap) {
                const uint64_t da = da0 + (uint64_t)(TAPS == 9 ? ((tap / 3) * Cfg::kWinW + tap % 3) * 8 : 0);
                const uint64_t db = db0 + (uint64_t)(tap * (Cfg::kBBytes >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                              (tap | k) != 0 ? 1u : (uint32_t)(cb != 0));
              }
              tc_commit_pair(&a_empty[sa]);
              if (last_cb) tc_commit_pair(&t_full[buf]);
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int tg = 0; tg < TAPS; tg += Cfg::kTB) {
              if constexpr (k32) {
                // split-operand mode: every weight stage (one kernel row: 12 MMAs) is its own
                // accumulation CHAIN into a fresh TMEM buffer; the epilogue warps add the chains up
                // in fp32 registers.  The tensor core truncates when it accumulates (measured: a
                // relative bias of ~2^-25 per MMA step, i.e. 1e-4 after the 3240 steps to conv4_2),
                // so the length of a chain, not K, sets the bias.
                buf = it & 3, use = it >> 2;
                d_tmem = tmem_base + buf * BN;
                mbar_wait(&t_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
              }
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint64_t db0 = make_smem_desc(smem_u32(b_base + sb * Cfg::kBStage));
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < Cfg::kTB; ++j) {
                  const int tap = tg + j;
                  const uint64_t da =
                      da0 + (uint64_t)(TAPS == 9 ? ((tap / 3) * Cfg::kWinW + tap % 3) * 8 : 0);
                  const uint64_t db = db0 + (uint64_t)(j * (Cfg::kBBytes >> 4));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    tc_mma_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                k32 ? (uint32_t)((j | k) != 0)
                                    : ((tap | k) != 0 ? 1u : (uint32_t)(cb != 0)));
                }
                tc_commit_pair(&b_empty[sb]);
                if constexpr (k32) tc_commit_pair(&t_full[buf]);
                if (tg + Cfg::kTB >= TAPS) {
                  tc_commit_pair(&a_empty[sa]);
                  if constexpr (!k32) {
                    if (last_cb) tc_commit_pair(&t_full[buf]);
                  }
                }
              }
              __syncwarp();
              if constexpr (k32) ++it;
              if (++sb == Cfg::kSB) sb = 0, pb ^= 1;
            }
          }
          if (++sa == Cfg::kSA) sa = 0, pa ^= 1;
        }
      }
    }
  } else if (warp == kInjWarp) {
    // ========================== operand producer of the backward epilogue ========================
    // The injected loss gradient (style / content term of the output blob) of this CTA's 128 pixels
    // x 64 channels, one TMA box per (tile, channel group) into a ring the epilogue warps drain.
    // Per-thread global loads of it (16 bytes per lane, a DRAM round trip of ~2 us under load with
    // at most one chunk in flight per warp) capped the conv1_2 backward at 1 TB/s of operand reads.
    if constexpr (EPI == kEpiBwd) {
      if (a.inj != nullptr) {
        prefetch_tmap(&map_aux);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
          const TileCoord t = decode_tile(a, tile, (int)rank);
          for (int g = 0; g < BN / 64; ++g) {
            mbar_wait(&e_empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&e_full[stage], kOutStageBytes);
              tma_load_4d_local(&map_aux, &e_full[stage], inj_base + stage * kOutStageBytes,
                                t.n_tile * BN + g * 64, t.x0, t.y0, t.b);
            }
            __syncwarp();
            if (++stage == kSE) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    const int q = warp & 3;                            // TMEM lane quadrant of this warp
    const int hsel = (warp - kEpiWarp0) >> 2;          // which 32-channel half of a 64-channel group
    const int m = q * 32 + lane;                       // pixel row of the CTA tile
    const bool issuer = threadIdx.x == kEpiWarp0 * 32;
    const bool out_half = a.out_half != 0;
    uint32_t it = 0, store_seq = 0;
    // Backward epilogue operands (ReLU mask = the forward activation, injected loss gradient): this
    // thread's pixel row of a tile, 64 bytes per 32-channel chunk and array.  They come from HBM, and
    // a load issued at the start of a tile made every tile wait a full DRAM latency (48 % of the
    // epilogue warps' samples sat on the first use, profiles/r01_conv_small_ncu.md).  Now the chunk
    // sequence is pipelined ACROSS tiles: while chunk i is processed the registers already receive
    // chunk i+1 (of this tile or the next), and the lines of the tile after that are pulled into L2.
    struct RowRef {
      size_t gofs;
      bool valid;
    };
    auto row_of = [&](int tl) {
      RowRef r{0, false};
      if (tl < num_tiles) {
        const TileCoord tc = decode_tile(a, tl, (int)rank);
        const int py = tc.y0 + (m >> 3), px = tc.x0 + (m & 7);
        r.valid = py < a.h && px < a.w;
        r.gofs = (((size_t)tc.b * a.h + py) * a.w + px) * a.cout + (size_t)tc.n_tile * BN;
      }
      return r;
    };
    uint32_t pbits = 0xFFFFFFFFu;
    int es = 0;                                            // injected-gradient ring position
    uint32_t ep = 0;
    auto prefetch = [&](const RowRef& rr, int cc) {
      if constexpr (EPI == kEpiBwd) {
        pbits = 0xFFFFFFFFu;                               // no mask: everything passes
        // volatile asm: keeps the load where it is written (see the call site in the chunk loop)
        if (rr.valid && a.mask_bits != nullptr)
          asm volatile("ld.global.nc.u32 %0, [%1];"
                       : "=r"(pbits)
                       : "l"(a.mask_bits + (rr.gofs >> 5) + cc));
      }
    };
    if constexpr (EPI == kEpiBwd) prefetch(row_of(pair), hsel);
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const uint32_t buf = it & 1, use = it >> 1;
      const TileCoord t = decode_tile(a, tile, (int)rank);
      const int n_tile = t.n_tile, x0 = t.x0, y0 = t.y0;
      const int py = y0 + (m >> 3), px = x0 + (m & 7);
      const bool valid = py < a.h && px < a.w;
      const RowRef cur{(((size_t)t.b * a.h + py) * a.w + px) * a.cout + (size_t)n_tile * BN, valid};
      RowRef nxt{0, false};
      if constexpr (EPI == kEpiBwd) nxt = row_of(tile + num_pairs);
      float abs_tile = 0.f;
      float inj_sc = 1.f;
      if constexpr (EPI == kEpiBwd) {
        if (a.inj_scale != nullptr) inj_sc = __ldg(a.inj_scale + t.b);
      }

      if constexpr (!k32) {
        mbar_wait(&t_full[buf], use & 1);
        tc_fence_after();
      }
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      if constexpr (EPI == kEpiPix) {
        // backward of the first convolution: accumulator columns 0..2 are d(loss)/d(pixel) of the
        // three image planes; written straight to the planar f32 gradient (no staging, no TMA)
        uint32_t r[16] = {0u, 0u, 0u};
        if (hsel == 0) tmem_ld16(taddr, r);            // the second warp of the quadrant only arrives
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&t_empty[buf]), 0));
        if (valid && hsel == 0) {
          float* dst = a.pix + (size_t)t.b * a.pix_batch + (size_t)py * a.pix_row + px;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) dst[(size_t)ci * a.pix_plane] = __uint_as_float(r[ci]);
        }
        continue;
      }
      if constexpr (k32) {
        // fp32 epilogue of the split-operand mode.  The MMA warp delivers one accumulation chain
        // per weight stage (3 chains per 64-channel K block); they are added here in fp32 with
        // round-to-nearest.  No staging, no TMA store: this thread's 32 channels of its pixel are one
        // 128-byte line of the NHWC fp32 output.
        constexpr int kG = BN / 64;
        float acc[kG][32];
#pragma unroll
        for (int g = 0; g < kG; ++g)
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[g][i] = 0.f;
        const int nchains = kb_per_tap * (TAPS / Cfg::kTB);
#pragma unroll 1
        for (int ch = 0; ch < nchains; ++ch, ++it) {
          const uint32_t cbuf = it & 3, cuse = it >> 2;
          mbar_wait(&t_full[cbuf], cuse & 1);
          tc_fence_after();
          const uint32_t caddr = tmem_base + cbuf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int g = 0; g < kG; ++g) {
            uint32_t r[32];
            tmem_ld32(caddr + (g * 2 + hsel) * 32, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[g][i] += __uint_as_float(r[i]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&t_empty[cbuf]), 0));
        }
        --it;                                  // the tile loop adds one
#pragma unroll
        for (int g = 0; g < kG; ++g) {
          const int cc = g * 2 + hsel;
          const size_t eofs = cur.gofs + (size_t)cc * 32;
          float* v = acc[g];
          float osc = a.out_scale;
          if constexpr (EPI == kEpiAbs32) osc *= __ldg(a.out_scale_tile + t.b);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= osc;
          if constexpr (EPI == kEpiAbs32) {
            // sum |S| over this warp's 32 rows x 32 channels, one slot per (pixel tile, 64-channel
            // group, CTA, warp) like kEpiAbs: added up later in slot order (rows outside the tile
            // hold exact zeros: their A rows were zero-filled)
            float at = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) at += fabsf(v[i]);
            double x = (double)at;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            const int m_tile = tile / a.tiles_n;
            const int g64 = n_tile * (BN / 64) + g;
            if (lane == 0)
              a.abs_partials[(((size_t)m_tile * (a.cout >> 6) + g64) * 2 + rank) * 8 + hsel * 4 + q] = x;
          } else if constexpr (EPI == kEpiFwd32) {
            const float* bs = bias_s + n_tile * BN + cc * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bs[i], 0.f);
          } else if (valid) {
            if (a.mask_f32 != nullptr) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 mk = __ldg(reinterpret_cast<const float4*>(a.mask_f32 + eofs) + i);
                v[4 * i + 0] = mk.x > 0.f ? v[4 * i + 0] : 0.f;
                v[4 * i + 1] = mk.y > 0.f ? v[4 * i + 1] : 0.f;
                v[4 * i + 2] = mk.z > 0.f ? v[4 * i + 2] : 0.f;
                v[4 * i + 3] = mk.w > 0.f ? v[4 * i + 3] : 0.f;
              }
            }
            if (a.inj_f32 != nullptr) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 ij = __ldg(reinterpret_cast<const float4*>(a.inj_f32 + eofs) + i);
                v[4 * i + 0] += ij.x, v[4 * i + 1] += ij.y, v[4 * i + 2] += ij.z, v[4 * i + 3] += ij.w;
              }
            }
          }
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.out_f32 + eofs);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        continue;
      }
#pragma unroll 1
      for (int g = 0; g < BN / 64; ++g, ++store_seq) {
        uint8_t* stage_out = out_base + (store_seq & 1) * kOutStageBytes;
        if (issuer) tma_store_wait_read<1>();          // the store that used this buffer has read it
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {
          const int hh = hsel;
          const int cc = g * 2 + hh;                   // 32-channel chunk of this warp
          uint32_t r[32];
          tmem_ld32(taddr + cc * 32, r);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if constexpr (kFwd) {
            const float* bs = bias_s + n_tile * BN + cc * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += bs[i];      // ReLU happens in the pack below
          } else if constexpr (EPI == kEpiBwd) {
            uint4 ce[4];
            const uint32_t cbits = pbits;
            if (a.inj != nullptr) {
              // this thread's pixel row of the staged tile (SWIZZLE_128B as TMA wrote it)
              mbar_wait(&e_full[es], ep);
              const uint8_t* irow = inj_base + es * kOutStageBytes + (size_t)m * 128;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                ce[i] = *reinterpret_cast<const uint4*>(irow + (((hh * 4 + i) ^ (m & 7)) << 4));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) ce[i] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t e4[4] = {ce[i].x, ce[i].y, ce[i].z, ce[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // ReLU mask of the forward activation: bit (pair index) = even channel of the pair,
                // bit (16 + pair index) = odd channel (relu_bit)
                float x0 = v[8 * i + 2 * j], x1 = v[8 * i + 2 * j + 1];
                if (!(cbits & (1u << (4 * i + j)))) x0 = 0.f;
                if (!(cbits & (0x10000u << (4 * i + j)))) x1 = 0.f;
                v[8 * i + 2 * j] = fmaf(inj_sc, __uint_as_float(e4[j] << 16), x0);
                v[8 * i + 2 * j + 1] = fmaf(inj_sc, __uint_as_float(e4[j] & 0xFFFF0000u), x1);
              }
            }
            if (a.inj != nullptr) {
              // the staged values are in registers and used: hand the ring stage back
              __syncwarp();
              if (lane == 0) mbar_arrive_local(&e_empty[es]);
              if (++es == kSE) es = 0, ep ^= 1;
            }
            // mask bits of the next chunk (this tile's or the next tile's).  Issued AFTER the math:
            // placed before it, the load shared a hardware scoreboard with the shared-memory reads
            // of the ring and every chunk waited for a DRAM round trip (ncu: 26 % of all samples on
            // the first use of the ring data, stall_long_scoreboard)
            if (cc + 2 < BN / 32)
              prefetch(cur, cc + 2);
            else
              prefetch(nxt, hsel);                   // first chunk of this warp in the next tile
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) abs_tile += fabsf(v[i]);
            {
              // sum |S| over this warp's 32 rows x 32 channels: one slot per (pixel tile, 64-channel
              // group, CTA, warp), added up later in slot order -- the grouping does not depend on
              // BN or on the tile -> CTA schedule, so the sum is reproducible
              double x = (double)abs_tile;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
              const int m_tile = tile / a.tiles_n;
              const int g64 = n_tile * (BN / 64) + g;
              if (lane == 0)
                a.abs_partials[(((size_t)m_tile * (a.cout >> 6) + g64) * 2 + rank) * 8 + hsel * 4 + q] = x;
              abs_tile = 0.f;
            }
          }
          // registers -> swizzled staging tile [128 rows][128 B] (SWIZZLE_128B, as TMA expects)
          uint8_t* row = stage_out + (size_t)m * 128;
          // one warp-uniform branch per chunk (a branch inside every pack cost the short-K
          // kernels half their speed)
          uint32_t pw[16];
          if constexpr (kFwd) {
            if (out_half) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pw[i] = pack16_relu(v[2 * i], v[2 * i + 1], true);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pw[i] = pack16_relu(v[2 * i], v[2 * i + 1], false);
            }
          } else if (out_half) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pw[i] = pack16(v[2 * i], v[2 * i + 1], true);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pw[i] = pack16(v[2 * i], v[2 * i + 1], false);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = hh * 4 + j;
            *reinterpret_cast<uint4*>(row + ((chunk ^ (m & 7)) << 4)) =
                make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
          }
          if constexpr (kFwd) {
            // ReLU bit mask of this pixel's 32 channels for the backward pass: 16x less HBM traffic
            // than re-reading the activation there (one HSET2 + one LOP3 per channel pair)
            if (a.bits_out != nullptr) {
              uint32_t bits = 0u;
              if (out_half) {
#pragma unroll
                for (int i = 0; i < 16; ++i) bits |= gt2_mask<true>(pw[i], 0u) & (0x00010001u << i);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) bits |= gt2_mask<false>(pw[i], 0u) & (0x00010001u << i);
              }
              if (valid) a.bits_out[(cur.gofs >> 5) + cc] = bits;
            }
          }
        }
        if (g == BN / 64 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp (leader)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&t_empty[buf]), 0));
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if constexpr (kPool) {
          // The staged tile holds complete 2x2 windows (tile origins are even).  Each thread pools
          // one (pooled pixel, 8-channel) item: pooled values -> a second swizzled staging tile
          // for a TMA store, one mask byte per pooled element -> global memory.
          //   max: bits 0-1 = position of the first maximum (scan order), bit 2 = maximum > 0
          //   ave: bit d = input d of the window > 0            (what the backward pass needs)
          uint8_t* pstage = out_base + 2 * kOutStageBytes + (store_seq & 1) * kPoolStageBytes;
          const int et = threadIdx.x - kEpiWarp0 * 32;           // 0..255 among the epilogue warps
          const int ho = (a.h + 1) >> 1, wo = (a.w + 1) >> 1;
          {
            const int item = et, pp = item >> 3, j = item & 7;
            const int pr = pp >> 2, pc = pp & 3;                 // pooled row / column in the tile
            const int pyo = (y0 >> 1) + pr, pxo = (x0 >> 1) + pc;
            uint32_t ow[4], mk[2] = {0u, 0u};
            if (a.pool_mode == 1) {
              if (out_half)
                pool_max_item<true>(stage_out, pr, pc, j, a.h - y0, a.w - x0, ow, mk);
              else
                pool_max_item<false>(stage_out, pr, pc, j, a.h - y0, a.w - x0, ow, mk);
            } else {
              // average pooling: fp32 sum of the staged values / clipped window size
              float sum[8];
              uint32_t code[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) sum[e] = 0.f, code[e] = 0u;
              int cnt = 0;
#pragma unroll
              for (int d = 0; d < 4; ++d) {
                const int r = 2 * pr + (d >> 1), cx = 2 * pc + (d & 1);
                if (y0 + r >= a.h || x0 + cx >= a.w) continue;    // ceil mode: clipped window
                ++cnt;
                const int mm = r * 8 + cx;
                const uint4 raw =
                    *reinterpret_cast<const uint4*>(stage_out + mm * 128 + ((j ^ (mm & 7)) << 4));
                const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t2 = unpack16(w4[e], out_half);
                  sum[2 * e] += t2.x, sum[2 * e + 1] += t2.y;
                  if (t2.x > 0.f) code[2 * e] |= 1u << d;
                  if (t2.y > 0.f) code[2 * e + 1] |= 1u << d;
                }
              }
              const float inv = (float)(cnt > 0 ? cnt : 1);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                ow[e] = pack16(sum[2 * e] / inv, sum[2 * e + 1] / inv, out_half);
#pragma unroll
              for (int e = 0; e < 8; ++e) mk[e >> 2] |= code[e] << (8 * (e & 3));
            }
            *reinterpret_cast<uint4*>(pstage + pp * 128 + ((j ^ (pp & 7)) << 4)) =
                make_uint4(ow[0], ow[1], ow[2], ow[3]);
            if (pyo < ho && pxo < wo)
              *reinterpret_cast<uint2*>(a.pool_mask + (((size_t)t.b * ho + pyo) * wo + pxo) * a.cout +
                                        n_tile * BN + g * 64 + j * 8) = make_uint2(mk[0], mk[1]);
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (issuer) {
            if (a.write_full) tma_store_4d(&map_out, stage_out, n_tile * BN + g * 64, x0, y0, t.b);
            tma_store_4d(&map_aux, pstage, n_tile * BN + g * 64, x0 >> 1, y0 >> 1, t.b);
            tma_store_commit();
          }
        } else {
          if (issuer) {
            tma_store_4d(&map_out, stage_out, n_tile * BN + g * 64, x0, y0, t.b);
            tma_store_commit();
          }
        }
      }
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_bf16_map(const TcContext& tc, CUtensorMap* map, int rank, const void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    bool half = false) {
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t bdim[4], estride[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], bdim[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_bytes[i];
  CUresult r = reinterpret_cast<EncodeTiledFn>(tc.encode_fn)(
      map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
      const_cast<void*>(base), gdim, gstride, bdim,
      estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

template <int BN, int TAPS, int EPI, bool RESB>
int launch2r(TcContext& tc, const void* in, const void* wk, int wk_rows,
             void* out, void* pool_out, Tc2Args a, cudaStream_t s) {
  using Cfg = Cfg2<BN, TAPS, RESB, EPI == kEpiFwdPool, EPI == kEpiBwd>;
  a.tiles_x = cdiv(a.w, kBW), a.tiles_y = cdiv(a.h, 2 * kBH), a.tiles_n = a.cout / BN;
  a.div_x = FastDiv(a.tiles_x), a.div_y = FastDiv(a.tiles_y), a.div_n = FastDiv(a.tiles_n);
  CUtensorMap map_in, map_out, map_w, map_pool;
  {
    const uint64_t cm = (uint64_t)(a.cin_map ? a.cin_map : a.cin);
    const uint64_t dims[4] = {cm, (uint64_t)a.w, (uint64_t)a.h, (uint64_t)a.nb};
    const uint64_t strides[3] = {cm * 2, (uint64_t)a.w * cm * 2, (uint64_t)a.h * a.w * cm * 2};
    const uint32_t box[4] = {64, (uint32_t)Cfg::kWinW, (uint32_t)Cfg::kHaloRows, 1};
    int rc = encode_bf16_map(tc, &map_in, 4, in, dims, strides, box, a.in_half != 0);
    if (rc != ST_OK) return rc;
  }
  if (a.a_wrap == 0) a.a_wrap = 1 << 30;
  constexpr bool k32 = EPI == kEpiFwd32 || EPI == kEpiBwd32 || EPI == kEpiAbs32;
  if (EPI != kEpiPix && !k32) {
    const uint64_t dims[4] = {(uint64_t)a.cout, (uint64_t)a.w, (uint64_t)a.h, (uint64_t)a.nb};
    const uint64_t strides[3] = {(uint64_t)a.cout * 2, (uint64_t)a.w * a.cout * 2,
                                 (uint64_t)a.h * a.w * a.cout * 2};
    const uint32_t box[4] = {64, (uint32_t)kBW, (uint32_t)kBH, 1};
    int rc = encode_bf16_map(tc, &map_out, 4, out, dims, strides, box, a.out_half != 0);
    if (rc != ST_OK) return rc;
  } else {
    map_out = map_in;            // never dereferenced by the pixel epilogue
  }
  map_pool = map_out;
  if (EPI == kEpiBwd && a.inj != nullptr) {
    // the injected gradient has the geometry of the output: same box, its own base address
    const uint64_t dims[4] = {(uint64_t)a.cout, (uint64_t)a.w, (uint64_t)a.h, (uint64_t)a.nb};
    const uint64_t strides[3] = {(uint64_t)a.cout * 2, (uint64_t)a.w * a.cout * 2,
                                 (uint64_t)a.h * a.w * a.cout * 2};
    const uint32_t box[4] = {64, (uint32_t)kBW, (uint32_t)kBH, 1};
    int rc = encode_bf16_map(tc, &map_pool, 4, a.inj, dims, strides, box, false);
    if (rc != ST_OK) return rc;
  }
  if (EPI == kEpiFwdPool) {
    const uint64_t ho = (a.h + 1) / 2, wo = (a.w + 1) / 2;
    const uint64_t dims[4] = {(uint64_t)a.cout, wo, ho, (uint64_t)a.nb};
    const uint64_t strides[3] = {(uint64_t)a.cout * 2, wo * a.cout * 2, ho * wo * a.cout * 2};
    const uint32_t box[4] = {64, (uint32_t)(kBW / 2), (uint32_t)(kBH / 2), 1};
    int rc = encode_bf16_map(tc, &map_pool, 4, pool_out, dims, strides, box, a.out_half != 0);
    if (rc != ST_OK) return rc;
  }
  {
    const uint64_t k = (uint64_t)TAPS * a.cin;
    const uint64_t dims[3] = {k, (uint64_t)wk_rows, (uint64_t)(a.w_batched ? a.nb : 1)};
    const uint64_t strides[2] = {k * 2, k * 2 * wk_rows};
    const uint32_t box[3] = {64, (uint32_t)(BN / 2), 1};
    int rc = encode_bf16_map(tc, &map_w, 3, wk, dims, strides, box, a.in_half != 0);
    if (rc != ST_OK) return rc;
  }
  auto kern = conv_tc2_kernel<BN, TAPS, EPI, RESB>;
  ST_CUDA(tc_allow_smem(kern, 227 * 1024));
  a.resb_bytes = RESB ? TAPS * (a.cin / 64) * Cfg::kBBytes : 0;
  const int smem_bytes = RESB ? Cfg::kFixedBytes + a.resb_bytes : Cfg::kSmemBytes;
  const int tiles = a.tiles_x * a.tiles_y * a.tiles_n * a.nb;
  const int max_pairs = tc.sm_count / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  // algorithmic flops: the pixel epilogue computes 3 of its 16 accumulator columns for real
  // (split mode: a.cin counts the three K segments; the algorithmic flops are a third of the executed)
  TimerScope ts(s, (EPI == kEpiAbs || EPI == kEpiAbs32) ? kTimeStyleGrad : (EPI == kEpiPix ? kTimeConvSimt : kTimeConvTc),
                2.0 * TAPS * (k32 ? a.cin / 3 : a.cin) * (EPI == kEpiPix ? 3 : a.cout) * a.h * a.w * a.nb);
  ST_LAUNCH_ATTR(kern, 2 * pairs, kThreads2, smem_bytes, s, g_pdl, map_in, map_w, map_out, map_pool, a);
  return ST_OK;
}

// Resident weights when one output-channel tile covers the layer, the weights are shared by the
// batch and this CTA's share of them fits beside the A stages.
template <int BN, int TAPS, int EPI>
int launch2(TcContext& tc, const void* in, const void* wk, int wk_rows, void* out, const Tc2Args& a,
            cudaStream_t s, void* pool_out = nullptr) {
  if constexpr (BN <= 128) {
    using CfgR = Cfg2<BN, TAPS, true, EPI == kEpiFwdPool, EPI == kEpiBwd>;
    const long res = (long)TAPS * (a.cin / 64) * CfgR::kBBytes;
    if (a.cout == BN && !a.w_batched && res <= CfgR::kResMax && tc.resident_weights)
      return launch2r<BN, TAPS, EPI, true>(tc, in, wk, wk_rows, out, pool_out, a, s);
  }
  return launch2r<BN, TAPS, EPI, false>(tc, in, wk, wk_rows, out, pool_out, a, s);
}

// Output-channel tile: the widest BN that still gives every CTA pair work; wide tiles halve the
// weight traffic per flop, narrow ones fill the machine on the small feature maps.
int choose_bn(const TcContext& tc, int nb, int h, int w, int cout) {
  const int px_tiles = cdiv(w, kBW) * cdiv(h, 2 * kBH) * nb, pairs = tc.sm_count / 2;
  if ((tc.force_bn == 64 || tc.force_bn == 128 || tc.force_bn == 256) && cout % tc.force_bn == 0)
    return tc.force_bn;
  for (int bn = 256; bn > 64; bn >>= 1) {
    if (cout % bn != 0) continue;
    const long tiles = (long)px_tiles * (cout / bn);
    if (tiles * 10 >= pairs * 8) return bn;     // at least ~0.8 waves of pairs
  }
  return 64;
}

template <int TAPS, int EPI>
int dispatch_bn(TcContext& tc, int bn, const void* in, const void* wk, int wk_rows, void* out,
                const Tc2Args& a, cudaStream_t s, void* pool_out = nullptr) {
  switch (bn) {
    case 256: return launch2<256, TAPS, EPI>(tc, in, wk, wk_rows, out, a, s, pool_out);
    case 128: return launch2<128, TAPS, EPI>(tc, in, wk, wk_rows, out, a, s, pool_out);
    default: return launch2<64, TAPS, EPI>(tc, in, wk, wk_rows, out, a, s, pool_out);
  }
}

}  // namespace

int conv3x3_tc_pair(TcContext& tc, const TcWeights& w, const void* in, void* out, int nb, int h,
                    int wd, int cin, int cout, bool forward, const float* bias, uint32_t* relu_bits,
                    const __nv_bfloat16* inj, const float* inj_scale, cudaStream_t s) {
  Tc2Args a{};
  a.nb = nb, a.h = h, a.w = wd, a.cin = cin, a.cout = cout;
  // forward: activations in the context's activation format; backward: gradients are always bf16
  a.in_half = a.out_half = (forward && w.fwd_half) ? 1 : 0;
  a.bias = bias, a.inj = inj, a.inj_scale = inj_scale;
  if (forward)
    a.bits_out = relu_bits;
  else
    a.mask_bits = relu_bits;
  const int bn = choose_bn(tc, nb, h, wd, cout);
  if (forward) return dispatch_bn<9, kEpiFwd>(tc, bn, in, w.fwd, cout, out, a, s);
  return dispatch_bn<9, kEpiBwd>(tc, bn, in, w.bwd, cout, out, a, s);
}

// Forward convolution fused with the 2x2/2 pooling layer that consumes it: writes the pooled map,
// the backward mask and (write_full) the un-pooled output.
int conv3x3_pool_tc_pair(TcContext& tc, const TcWeights& w, const void* in, void* out,
                         void* pool_out, uint8_t* pool_mask, int nb, int h, int wd, int cin, int cout,
                         const float* bias, bool is_max, bool write_full, uint32_t* relu_bits,
                         cudaStream_t s) {
  Tc2Args a{};
  a.nb = nb, a.h = h, a.w = wd, a.cin = cin, a.cout = cout, a.bias = bias, a.bits_out = relu_bits;
  a.in_half = a.out_half = w.fwd_half ? 1 : 0;
  a.pool_mode = is_max ? 1 : 2, a.write_full = write_full ? 1 : 0, a.pool_mask = pool_mask;
  const int bn = choose_bn(tc, nb, h, wd, cout);
  return dispatch_bn<9, kEpiFwdPool>(tc, bn, in, w.fwd, cout, out, a, s, pool_out);
}

// Backward of the first convolution (cout image planes = 3, padded to 16 accumulator columns):
// grad[b][ci][y][x] = sum_{tap, co} dz[b][p + off(tap)][co] * Wk[ci][tap*cz + co].
int conv_last_bwd_tc_pair(TcContext& tc, const TcWeights& w, const __nv_bfloat16* dz, int nb, int h,
                          int wd, int cz, float* grad, long batch_stride, long plane_stride,
                          long row_stride, cudaStream_t s) {
  Tc2Args a{};
  a.nb = nb, a.h = h, a.w = wd, a.cin = cz, a.cout = 16;
  a.pix = grad, a.pix_batch = batch_stride, a.pix_plane = plane_stride, a.pix_row = row_stride;
  return launch2<16, 9, kEpiPix>(tc, dz, w.bwd, 16, nullptr, a, s);
}

// ---- split-operand mode (ST_PREC_TC32) -------------------------------------------------------------
namespace {
// fp32 NHWC [pixels][c] -> fp16 [pixels][2c]: channels [0, c) = hi = fp16(x * scale), [c, 2c) = lo =
// fp16(x * scale - hi).  hi saturates at the largest finite fp16 (lo then carries the rest).
__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ in,
                                                        __half* __restrict__ out, size_t groups,
                                                        int c8, float scale) {
  ST_PDL_ENTRY();
  // one thread per 8 channels of one pixel
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += stride) {
    const size_t px = i / c8;
    const int cg = (int)(i - px * c8);
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(in + i * 8));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(in + i * 8) + 1);
    const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = x[2 * j] * scale, a1 = x[2 * j + 1] * scale;
      const __half2 h = __floats2half2_rn(fminf(fmaxf(a0, -65504.f), 65504.f),
                                          fminf(fmaxf(a1, -65504.f), 65504.f));
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(fminf(fmaxf(a0 - hf.x, -65504.f), 65504.f),
                                          fminf(fmaxf(a1 - hf.y, -65504.f), 65504.f));
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    __half* o = out + px * (size_t)(c8 * 16) + (size_t)cg * 8;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
}  // namespace

int split_f32(const float* in, void* out, size_t pixels, int c, float scale, cudaStream_t s) {
  ST_REQUIRE(c % 8 == 0, "split_f32: channels must be a multiple of 8");
  const size_t groups = pixels * (size_t)(c / 8);
  const int blocks = (int)std::min<size_t>((groups + 255) / 256, (size_t)148 * 16);
  TimerScope ts(s, kTimeConvSimt, 0.0);
  ST_LAUNCH(split_f32_kernel, blocks, 256, 0, s, in, static_cast<__half*>(out), groups, c / 8, scale);
  return ST_OK;
}

// fp16 split packs of one conv layer for both directions: rows [cout][9 * 3cin] with, per tap, the
// K segments [Whi | Whi | Wlo] (matching the A channel order hi, lo, hi of the kernel's wrap), weights
// pre-multiplied by the power of two *scale that brings max|w| to ~2^13 (keeps Wlo out of the fp16
// subnormals).  backward: taps flipped, in/out channels exchanged.
int tc_pack_split(TcContext& tc, TcWeights& w, const float* w_host, int cin, int cout) {
  if (!tc.enabled) return ST_OK;
  float wmax = 0.f;
  for (size_t i = 0; i < (size_t)cout * cin * 9; ++i) wmax = std::max(wmax, std::fabs(w_host[i]));
  int e = 0;
  if (wmax > 0.f) {
    std::frexp(wmax, &e);                  // wmax = m * 2^e, m in [0.5, 1)
    e = 13 - e;
  }
  const float scale = std::ldexp(1.f, e);
  w.split_scale = scale;
  for (int dir = 0; dir < 2; ++dir) {
    const int rows = dir == 0 ? cout : cin, kc = dir == 0 ? cin : cout;   // output rows, K channels
    std::vector<__half> host((size_t)rows * 9 * 3 * kc);
    for (int r = 0; r < rows; ++r)
      for (int t = 0; t < 9; ++t)
        for (int k = 0; k < kc; ++k) {
          const int co = dir == 0 ? r : k, ci = dir == 0 ? k : r;
          const int ts = dir == 0 ? t : 8 - t;                 // flipped tap for the transposed conv
          const float v = w_host[((size_t)co * cin + ci) * 9 + ts] * scale;
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          __half* row = host.data() + ((size_t)r * 9 + t) * 3 * kc;
          row[k] = hi, row[kc + k] = hi, row[2 * kc + k] = lo;
        }
    void** dst = dir == 0 ? &w.fwd32 : &w.bwd32;
    if (!*dst) ST_CUDA(cudaMalloc(dst, host.size() * sizeof(__half)));
    ST_CUDA(cudaMemcpy(*dst, host.data(), host.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  return ST_OK;
}

// out = epilogue(conv3x3(in)) on fp32 NHWC tensors through the split-operand tensor-core kernel:
//   forward : out = max(acc + bias, 0)
//   backward: out = (mask_act > 0 ? acc : 0) + inj          (mask_act / inj fp32, may be null)
// `split_buf` (2 * nb*h*w*cin fp16 elements) receives the [hi | lo] copy of `in * in_scale`.
int conv3x3_tc32(TcContext& tc, const TcWeights& w, const float* in, float* out, int nb, int h,
                 int wd, int cin, int cout, bool forward, const float* bias, const float* mask_act,
                 const float* inj, float in_scale, void* split_buf, cudaStream_t s) {
  int rc = split_f32(in, split_buf, (size_t)nb * h * wd, cin, in_scale, s);
  if (rc != ST_OK) return rc;
  Tc2Args a{};
  a.nb = nb, a.h = h, a.w = wd, a.cin = 3 * cin, a.cout = cout;
  a.cin_map = 2 * cin, a.a_wrap = 2 * (cin / 64);
  a.in_half = 1, a.out_half = 0;
  a.out_scale = 1.f / (w.split_scale * in_scale);
  a.bias = bias, a.mask_f32 = mask_act, a.inj_f32 = inj, a.out_f32 = out;
  // BN <= 128: the epilogue keeps the running sum of the chains in registers (BN / 2 per thread)
  const int bn = std::min(choose_bn(tc, nb, h, wd, cout), 128);
  if (forward) {
    if (bn == 128) return launch2r<128, 9, kEpiFwd32, false>(tc, split_buf, w.fwd32, cout, nullptr, nullptr, a, s);
    return launch2r<64, 9, kEpiFwd32, false>(tc, split_buf, w.fwd32, cout, nullptr, nullptr, a, s);
  }
  if (bn == 128) return launch2r<128, 9, kEpiBwd32, false>(tc, split_buf, w.bwd32, cout, nullptr, nullptr, a, s);
  return launch2r<64, 9, kEpiBwd32, false>(tc, split_buf, w.bwd32, cout, nullptr, nullptr, a, s);
}

// ---- style GEMM of the split-operand mode ------------------------------------------------------------
namespace {
// delta [nb][c][c] fp32 -> fp16 rows [nb][c][3c] = [Dhi | Dhi | Dlo] of delta * sigma_b, sigma_b the power
// of two that puts max |delta_b| (max_bits[b], float bits) into [2^12, 2^13); inv_sigma[b] = 1 / sigma_b
__global__ void __launch_bounds__(256)
delta_split_kernel(const float* __restrict__ delta, const unsigned* __restrict__ max_bits,
                   __half* __restrict__ out, float* __restrict__ inv_sigma, int c) {
  ST_PDL_ENTRY();
  const int b = blockIdx.y;
  const float mx = __uint_as_float(max_bits[b]);
  float sigma = 1.f;
  if (mx > 0.f && isfinite(mx)) {
    int e;
    frexpf(mx, &e);
    sigma = ldexpf(1.f, 13 - e);
  }
  const float* d = delta + (size_t)b * c * c;
  __half* o = out + (size_t)b * c * 3 * c;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < c * c; idx += gridDim.x * blockDim.x) {
    const int n = idx / c, k = idx - n * c;
    const float v = d[idx] * sigma;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    __half* row = o + (size_t)n * 3 * c;
    row[k] = hi, row[c + k] = hi, row[2 * c + k] = lo;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) inv_sigma[b] = 1.f / sigma;
}
}  // namespace

// Per batch tile b: S_b[p][n] = sum_c F_b[p][c] * D_b[n][c] in fp32 from the [hi | lo] planes of F
// (f_split [nb][h][w][2c] fp16, split_f32 with scale 1) and the fp32 delta-Gram D_b (symmetric
// [c][c]); d_split: scratch of nb * c * 3c fp16, inv_sigma: nb floats; max_bits[b] = max |D_b| as
// float bits.  sum |S_b| is left as partial sums like gemm_abs_tc_pair.
int gemm_abs_tc32(TcContext& tc, const void* f_split, const float* delta, const unsigned* max_bits,
                  void* d_split, float* inv_sigma, float* s_out, int nb, int h, int w, int c,
                  double* abs_partials, int* per_tile, cudaStream_t s) {
  ST_LAUNCH(delta_split_kernel, dim3(std::min(cdiv((long)c * c, 256), 64), nb), 256, 0, s, delta, max_bits,
            static_cast<__half*>(d_split), inv_sigma, c);
  Tc2Args a{};
  a.nb = nb, a.h = h, a.w = w, a.cin = 3 * c, a.cout = c, a.w_batched = 1;
  a.cin_map = 2 * c, a.a_wrap = 2 * (c / 64);
  a.in_half = 1, a.out_half = 0, a.out_scale = 1.f, a.out_scale_tile = inv_sigma;
  a.abs_partials = abs_partials, a.out_f32 = s_out;
  *per_tile = cdiv(w, kBW) * cdiv(h, 2 * kBH) * (c / 64) * 16;
  const int bn = std::min(choose_bn(tc, nb, h, w, c), 128);
  if (bn == 128) return launch2r<128, 1, kEpiAbs32, false>(tc, f_split, d_split, c, nullptr, nullptr, a, s);
  return launch2r<64, 1, kEpiAbs32, false>(tc, f_split, d_split, c, nullptr, nullptr, a, s);
}

// Per batch tile b: S_b[p][n] = sum_c F_b[p][c] * D_b[n][c]  (D_b symmetric bf16 [c][c]).  The sum
// of |S_b| is left as partial sums: abs_partials[b * per_tile + i], i < *per_tile.
int gemm_abs_tc_pair(TcContext& tc, const void* f, const void* d, bool half_in, __nv_bfloat16* s_out,
                     int nb, int h, int w, int c, double* abs_partials, int* per_tile,
                     cudaStream_t s) {
  Tc2Args a{};
  a.nb = nb, a.h = h, a.w = w, a.cin = c, a.cout = c, a.w_batched = 1;
  a.in_half = half_in ? 1 : 0, a.out_half = 0;          // S is a gradient: bf16
  a.abs_partials = abs_partials;
  const int bn = choose_bn(tc, nb, h, w, c);
  *per_tile = cdiv(w, kBW) * cdiv(h, 2 * kBH) * (c / 64) * 16;
  return dispatch_bn<1, kEpiAbs>(tc, bn, f, d, c, s_out, a, s);
}

size_t gemm_abs_partials_needed(int nb, int h, int w, int c) {
  return (size_t)nb * cdiv(w, kBW) * cdiv(h, 2 * kBH) * (c / 64) * 16;
}

}  // namespace st
