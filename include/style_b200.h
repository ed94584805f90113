/*
 * style_b200.h -- C ABI of libstyle_b200.so, the B200 (sm_100a) engine for the per-tile hot path
 * of crowsonkb/style_transfer.
 *
 * The reference has no FFI of its own: its hot path is Python calling pycaffe (C++/CUDA, external)
 * and SciPy BLAS.  The entry points below are what a binding for that path binds instead; each
 * cites the reference interface it replaces (file:line in the reference tree).  INTEGRATION.md
 * shows the ctypes stubs a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative st_status on failure; the message of the
 *     last failure on the calling thread is returned by st_last_error();
 *   - "dev" pointers are CUDA device pointers on the context's device, "host" pointers are
 *     ordinary host memory; all tensors are float32, C-contiguous, single image;
 *   - images / gradients are [3][h][w] (BGR, mean-subtracted: style_transfer.py:388-393), feature
 *     maps are [C][hf][wf], Gram matrices are [C][C] with only the LOWER triangle meaningful
 *     (num_utils.py:53-56, 143-147);
 *   - the caller owns every buffer it passes; the library owns its context (packed weights,
 *     activation workspace, copies of the targets).  No pointer is retained after a call returns,
 *     but work is asynchronous on `stream` (a cudaStream_t; 0 = legacy default stream): buffers
 *     must stay alive until the stream has been synchronised;
 *   - calls on one context must be serialised by the caller; contexts are independent.
 *   - scalars produced on the device (losses) are ACCUMULATED into a caller-provided device
 *     double so that a whole evaluation needs no host synchronisation.
 */
#ifndef STYLE_B200_H
#define STYLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ST_API __attribute__((visibility("default")))
#else
#define ST_API
#endif

typedef struct st_ctx st_ctx;
typedef void* st_stream;            /* cudaStream_t */

enum st_status {
  ST_OK = 0,
  ST_ERR_INVALID = -1,              /* bad argument / unsupported shape */
  ST_ERR_CUDA = -2,                 /* a CUDA runtime or driver call failed */
  ST_ERR_STATE = -3,                /* call out of order (e.g. targets not set) */
  ST_ERR_NOMEM = -4
};

/* Arithmetic of the convolution / Gram contractions.  Reductions, losses, regularisers and the
 * optimizers are float32 (double accumulators) in both modes. */
enum st_precision {
  ST_PREC_FP32 = 0,                 /* float32 SIMT everywhere: the parity mode                   */
  ST_PREC_BF16 = 1,                 /* bf16 operands, fp32 accumulate on tcgen05 tensor cores     */
  ST_PREC_FP16 = 2,                 /* tensor cores with fp16 forward activations / weights (11-bit
                                       significand, range 6e-5 .. 65504) and bf16 gradients: same
                                       speed as ST_PREC_BF16, ~4x smaller gradient error          */
  ST_PREC_TC32 = 3                  /* fp32 storage; the 3x3 convolutions and the style GEMM on the
                                       tensor cores with SPLIT fp16 operands (x = hi + lo, three MMAs
                                       per product) and CHAINED accumulation (a fresh TMEM buffer per
                                       12 MMAs, chains summed in fp32 registers: the tensor core
                                       truncates when it accumulates): the reference's fp32 results
                                       -- features within 3e-6 of ST_PREC_FP32, gradients within the
                                       same distance of the CPU oracle -- at a third of the fp16
                                       mode's tensor throughput                                   */
};

enum st_layer_kind { ST_CONV3X3 = 0, ST_POOL_MAX = 1, ST_POOL_AVE = 2 };

/* One layer of the deploy graph (vgg19.prototxt:16-60).  Blob 0 is "data"; the top of layers[i]
 * is blob i+1.  Convolutions are 3x3 / pad 1 / stride 1 followed by an in-place ReLU layer
 * (vgg19.prototxt:27-32); pools are 2x2 / stride 2, ceil mode. */
typedef struct {
  int32_t kind;                     /* st_layer_kind */
  int32_t bottom;                   /* blob index this layer reads */
  int32_t cin, cout;                /* channels (pools: cin == cout) */
} st_layer_desc;

/* Loss terms attached to one blob in eval_sc_grad_tile (style_transfer.py:569-604).  The weights
 * are the fully-multiplied factors the reference uses:
 *   content_weight = lw * content_weight[layer]        (:579-580)
 *   style_weight   = lw * style_weight[layer]          (:591-592; the library divides by n_styles)
 *   dd_weight      = lw * dd_weight[layer]             (:603-604)
 * A zero weight together with the matching flag cleared disables the term. */
typedef struct {
  int32_t blob;
  int32_t use_content, use_style, use_dd;
  float content_weight, style_weight, dd_weight;
} st_loss_spec;

ST_API const char* st_last_error(void);
ST_API int st_version(void);

/* ---- context ----------------------------------------------------------------------------
 * Replaces CaffeModel.__init__ -> caffe.Net(deploy, 1, weights=...) (style_transfer.py:359-376). */
ST_API int st_create(int device, int precision, int n_layers, const st_layer_desc* layers, st_ctx** out);
ST_API int st_destroy(st_ctx* ctx);
/* Weights of conv layer `layer` (index into layers[]): host float32 OIHW [cout][cin][3][3] and
 * bias [cout] -- the .caffemodel blobs. */
ST_API int st_set_conv_params(st_ctx* ctx, int layer, const float* w_host, const float* b_host);
/* Reserve the activation / gradient workspace for tiles up to max_h x max_w (grows if needed;
 * later calls with smaller tiles reuse it).  Replaces blobs['data'].reshape (:423, :559). */
ST_API int st_reserve(st_ctx* ctx, int max_h, int max_w);
ST_API int st_device_info(st_ctx* ctx, int* sm_count, size_t* workspace_bytes);

/* ---- targets: TileWorker SetContentsAndStyles (style_transfer.py:243-254, 309-332) ---------- */
ST_API int st_clear_targets(st_ctx* ctx);
/* Appends one StyleData entry / sets its Gram for `blob`: dev f32 [C][C], lower triangle used. */
ST_API int st_set_style_gram(st_ctx* ctx, int style_index, int blob, const float* gram_dev,
                      st_stream stream);
/* ContentData.features[layer] of the WHOLE image: dev f32 [C][hf][wf]. */
ST_API int st_set_content_features(st_ctx* ctx, int content_index, int blob, const float* feat_dev,
                            int hf, int wf, st_stream stream);

/* ---- the per-tile operator ------------------------------------------------------------------
 * CaffeModel.eval_features_tile (style_transfer.py:421-427): forward of one tile; blob_ids[i]'s
 * (post-ReLU) feature map is written to out_dev[i] as f32 [C][hf][wf].  `out_dev` is a HOST array
 * of device pointers.  The forward runs to the deepest requested blob (the reference always runs
 * to pool5; the returned maps are identical). */
ST_API int st_eval_features_tile(st_ctx* ctx, const float* img_dev, int h, int w, int n_blobs,
                          const int32_t* blob_ids, float* const* out_dev, st_stream stream);

/* CaffeModel.eval_sc_grad_tile (style_transfer.py:556-612): loss and d(loss)/d(pixels) of one
 * tile.  `img_dev` is the contiguous tile [3][h][w]; (start_y, start_x) is the tile origin in the
 * rolled image (`start`, :572); (feat_roll_y, feat_roll_x) is the pixel roll the worker applied to
 * its content features before the call (req.roll, :234) -- the library indexes the un-rolled
 * feature maps circularly instead of moving them.  grad_dev receives [3][h][w] with row stride
 * grad_row_stride and plane stride grad_plane_stride (floats); loss_accum_dev (device double) is
 * incremented by the tile's loss. */
ST_API int st_eval_sc_grad_tile(st_ctx* ctx, const float* img_dev, int h, int w, int start_y, int start_x,
                         int feat_roll_y, int feat_roll_x, int n_specs, const st_loss_spec* specs,
                         double* loss_accum_dev, float* grad_dev, long grad_plane_stride,
                         long grad_row_stride, st_stream stream);

/* CaffeModel.eval_sc_grad (style_transfer.py:614-645) for the tiles of ONE rank: the image
 * [3][H][W] is given un-rolled; the per-iteration roll (roll_y, roll_x pixels, i.e. xy*jitter_scale
 * of :784-786 with xy[0] -> x, xy[1] -> y) is applied virtually while cutting tiles.  Tiles
 * rank, rank+world, ... of the row-major tile grid (round-robin of TileWorkerPool.request,
 * :284-298) are evaluated; the gradient of local tile j is written to
 * packed_grad_dev[j][3][tile_h_max][tile_w_max].  Pass world=1, rank=0 for all tiles. */
ST_API int st_eval_sc_grad_tiles(st_ctx* ctx, const float* img_dev, int H, int W, int roll_y, int roll_x,
                          int tile_size, int rank, int world, int n_specs,
                          const st_loss_spec* specs, double* loss_accum_dev,
                          float* packed_grad_dev, st_stream stream);
/* The same for the local tiles (slots) [slot_first, slot_first + slot_count) of this rank only
 * (slot_count < 0: to the end): lets a caller overlap the upload of the image rows the later tiles
 * read with the evaluation of the earlier ones.  Tiles that share a call share kernel launches. */
ST_API int st_eval_sc_grad_tile_range(st_ctx* ctx, const float* img_dev, int H, int W, int roll_y,
                               int roll_x, int tile_size, int rank, int world, int slot_first,
                               int slot_count, int n_specs, const st_loss_spec* specs,
                               double* loss_accum_dev, float* packed_grad_dev, st_stream stream);
/* Geometry of the tile grid used above (style_transfer.py:619-631). */
ST_API int st_tile_grid(int H, int W, int tile_size, int* ntiles_y, int* ntiles_x, int* tile_h_max,
                 int* tile_w_max);
/* ---- the exchange step (style_transfer.py:635-643: the master's resp_q.get() loop) ---------------
 * Layout of one rank's chunk of the exchange buffer, st_packed_floats() floats:
 *   [tiles_per_rank][3][tile_h_max][tile_w_max] gradient tiles (what st_eval_sc_grad_tiles writes),
 *   padded to a multiple of four floats, then a four-float tail whose first eight bytes hold the
 *   rank's loss as a double (pass the tail's address as loss_accum_dev to st_eval_sc_grad_tiles).
 * One all-gather of these chunks moves gradients AND losses: a single collective per evaluation.
 * st_comm_unique_id (rank 0) / st_comm_init (every rank, same id) create the NCCL communicator of a
 * context -- libnccl.so.2 is resolved at run time, single-GPU use never needs it; st_allgather_grad
 * enqueues ncclAllGather on `stream`; packed_all_dev receives world chunks in rank order. */
#define ST_COMM_ID_BYTES 128
ST_API size_t st_packed_floats(int H, int W, int tile_size, int world);
ST_API int st_comm_unique_id(void* id_out /* ST_COMM_ID_BYTES */);
ST_API int st_comm_init(st_ctx* ctx, const void* id, int rank, int world);
ST_API int st_comm_destroy(st_ctx* ctx);
ST_API int st_allgather_grad(st_ctx* ctx, const float* packed_local_dev, float* packed_all_dev,
                             size_t floats_per_rank, st_stream stream);
/* Pastes the gradient tiles of all ranks (the all-gather result: world chunks as above; world = 1:
 * the local chunk itself) into grad_dev [3][H][W] in the UN-rolled frame (:642 + the roll-back :805)
 * and adds the ranks' losses, in rank order, to *loss_accum_dev (may be NULL). */
ST_API int st_unpack_grad(const float* packed_all_dev, int H, int W, int roll_y, int roll_x, int tile_size,
                   int world, float* grad_dev, double* loss_accum_dev, st_stream stream);

/* ---- Gram of a full feature map (preprocess_images, style_transfer.py:534; num_utils.py:143) -- */
ST_API int st_gram(st_ctx* ctx, const float* feat_dev, int c, int hw, float* gram_dev, st_stream stream);

/* ---- full-image regularisers: StyleTransfer.eval_loss_and_grad (style_transfer.py:700-736) ----
 * grad += tv_w * d tv_norm(img/127.5, tv_beta) + p_w * d p_norm((img+mean-127.5)/127.5, p_pow)
 *         + aux_w * (img - aux)/127.5 ;   loss_accum += the matching loss terms.
 * Weights already include layer_weights['data'].  TV uses periodic differences, so it is
 * invariant to the roll; `aux_dev` (may be NULL) is compared at the rolled position, as the
 * reference compares its rolled image with the un-rolled aux image (:731). */
ST_API int st_regularizers(const float* img_dev, int H, int W, const float mean[3], float tv_w,
                    float tv_beta, float p_w, float p_pow, const float* aux_dev, float aux_w,
                    int roll_y, int roll_x, double* loss_accum_dev, float* grad_dev,
                    st_stream stream);

/* st_unpack_grad followed by st_regularizers in ONE pass over the image: grad_dev is written (not
 * accumulated), the gradient tiles are gathered from the all-gather buffer on the fly, the ranks'
 * losses in the chunk tails and the regulariser terms are added to *loss_accum_dev. */
ST_API int st_unpack_regularize(const float* packed_all_dev, const float* img_dev, int H, int W,
                         int roll_y, int roll_x, int tile_size, int world, const float mean[3],
                         float tv_w, float tv_beta, float p_w, float p_pow, const float* aux_dev,
                         float aux_w, double* loss_accum_dev, float* grad_dev, st_stream stream);

/* ---- optimizers (optimizers.py) -------------------------------------------------------------
 * AdamOptimizer.update :26-42 after opfunc returned `grad`: EWMA moments (state g1,g2,p1 hold the
 * EWMA .value arrays), in-place parameter step, iterate averaging; avg_out = p1 / p1_corr.
 * g1_corr/g2_corr/p1_corr = 1 - beta^t (or 1 when bias correction is off).  All six arrays must be
 * 16-byte aligned (whole allocations are): the kernel moves float4. */
ST_API int st_adam_step(float* params, const float* grad, float* g1, float* g2, float* p1, float* avg_out,
                 size_t n, float step_size, float b1, float b2, float bp1, float g1_corr,
                 float g2_corr, float p1_corr, st_stream stream);
/* ---- per-iteration output step (style_transfer.py:808-821, :378-386) ------------------------------
 * st_iter_stats: the two statistics of StyleTransfer.transfer in one pass over the averaged iterate:
 *   stats_dev[0] = sum |avg - old|                      (update_size = stats[0] / (3*H*W), :809)
 *   stats_dev[1] = sum (x_diff^2 + y_diff^2),  x_diff = avg - roll(avg, -1, axis=-1), y_diff likewise
 *                                                       (tv_loss = sqrt(stats[1] / (3*H*W)), :813-815)
 * and old := avg (:810).  avg_dev / old_dev are f32 [3][H][W].
 * st_get_image_u8: CaffeModel.get_image (:378-386): out[y][x][k] = uint8(clip(params[c][y][x] +
 * mean[c], 0, 255)) with c = 2 - k when bgr (the model is BGR, the picture RGB), else c = k;
 * out_dev is uint8 [H][W][3]. */
ST_API int st_iter_stats(const float* avg_dev, float* old_dev, int H, int W, double* stats_dev,
                         st_stream stream);
ST_API int st_get_image_u8(const float* params_dev, int H, int W, const float mean[3], int bgr,
                           uint8_t* out_dev, st_stream stream);
/* Both of the above in ONE pass over the averaged iterate (every array is read once): the statistics,
 * old := avg, and -- unless pic_dev is NULL -- the uint8 picture of avg. */
ST_API int st_output_step(const float* avg_dev, float* old_dev, int H, int W, const float mean[3],
                          int bgr, double* stats_dev, uint8_t* pic_dev, st_stream stream);

/* ---- scale change (num_utils.resize :90-108, optimizers.py:53-61, style_transfer.py:877-881) ------
 * Per-channel float resampling of in_dev f32 [channels][h][w] to out_dev f32 [channels][out_h][out_w]
 * with Pillow's algorithm for 'F' images, which is what the reference calls: horizontal pass first
 * into a float32 intermediate, then vertical; per output sample the source samples are weighted in
 * source order with float64 coefficients (normalised Lanczos-3 or bilinear, support scaled when
 * shrinking) and float64 accumulation without fused multiply-add.  method: 0 = Lanczos, 1 = bilinear.
 * tmp_dev: scratch of channels*h*out_w floats (unused when out_w == w).  Coefficient tables are built
 * on the host per call.  Pinned: oracle.numeric.resize == PIL == the reference's num_utils.resize bit
 * for bit (CPU), and this kernel == the oracle bit for bit (B200).  The bundled command line uses it
 * for the scale change (ST_HOST_RESIZE=1: through PIL on the host, as the reference). */
ST_API int st_resize_f32(const float* in_dev, int channels, int h, int w, int out_h, int out_w,
                         int method, float* out_dev, float* tmp_dev, st_stream stream);
/* The coefficient table st_resize_f32 uses for one axis (host only, no device needed): *ksize weights
 * per output sample.  Call with bounds_out = kk_out = NULL to get *ksize, then with int
 * bounds_out[out_size][2] (first source index, count) and double kk_out[out_size][*ksize]. */
ST_API int st_resample_coeffs(int in_size, int out_size, int method, int* ksize, int* bounds_out,
                              double* kk_out);

/* LBFGSOptimizer.inv_hv :105-121 for a binding that keeps the curvature pairs as host-side lists like
 * the reference: two-loop recursion over m (<= 16) pairs.  s_dev / y_dev are HOST arrays of m device
 * pointers (oldest first), sy_host the stored s.y products; p_dev receives H*grad.  scratch_dev:
 * >= 64 doubles. */
ST_API int st_lbfgs_inv_hv(const float* grad_dev, size_t n, int m, const float* const* s_dev,
                    const float* const* y_dev, const double* sy_host, float* p_dev,
                    double* scratch_dev, st_stream stream);
/* LBFGSOptimizer.update / store_curvature_pair (optimizers.py:74-103) with the memory ON THE DEVICE:
 * no host synchronisation, no host decision.  ring_s_dev / ring_y_dev: f32 [n_corr + 1][n] (n_corr <=
 * 16); state_dev: 64 doubles, all zero = empty memory ([0] = number of valid pairs, [1] = ring slot of
 * the next pair, rest: s.y products and scratch); scratch_dev: n floats.
 * st_lbfgs_step:   s = -H grad by the two-loop recursion over the valid pairs (:105-121), scaled as
 *                  :81-84 (mean|s| = initial_step while the memory is empty, count / n_corr until it
 *                  is full), written to the free ring slot, and params += s (:85).
 * st_lbfgs_commit: after the objective was evaluated at the new params: y = grad_new - grad_old goes
 *                  to the same slot and the pair is kept iff s.y > 1e-10 (:92-103; the oldest pair
 *                  falls out once n_corr are held). */
ST_API int st_lbfgs_step(const float* grad_dev, size_t n, int n_corr, float* ring_s_dev,
                         const float* ring_y_dev, double* state_dev, float* scratch_dev,
                         float* params_dev, float initial_step, st_stream stream);
ST_API int st_lbfgs_commit(const float* grad_new_dev, const float* grad_old_dev, size_t n, int n_corr,
                           const float* ring_s_dev, float* ring_y_dev, double* state_dev,
                           st_stream stream);
/* BLAS-1 helpers used by LBFGSOptimizer.update/store_curvature_pair :74-103. */
ST_API int st_dot(const float* x, const float* y, size_t n, double* out_dev, st_stream stream);
ST_API int st_asum(const float* x, size_t n, double* out_dev, st_stream stream);
/* y = a*x + b*y elementwise (a, b host scalars). */
ST_API int st_axpby(float a, const float* x, float b, float* y, size_t n, st_stream stream);

/* ---- opt-in per-kernel timing (bench.py's roofline leg) ------------------------------------------
 * While enabled, every launcher of the categories below brackets its kernels with CUDA events on
 * the launching stream.  st_timing_read synchronises the device and returns, for one category, the
 * summed device time (ms), the summed ALGORITHMIC work (flops for the tensor categories, bytes for
 * the HBM ones) and the number of bracketed launch groups since the last st_timing_reset.
 * Process-wide, not thread-safe: enable it from the thread that drives the context. */
enum st_timing_category {
  ST_TIME_CONV_TC = 0,      /* tcgen05 implicit-GEMM convolutions, flops */
  ST_TIME_CONV_SIMT = 1,    /* first/last 3-channel layer (any kernel) + fp32-mode SIMT convolutions, flops */
  ST_TIME_POOL = 2,         /* pooling fwd/bwd, bytes */
  ST_TIME_GRAM = 3,         /* Gram F^T F, flops */
  ST_TIME_STYLE_GRAD = 4,   /* delta-Gram x F, flops */
  ST_TIME_LOSS = 5,         /* loss statistics + gradient injection, bytes */
  ST_TIME_IMAGE = 6,        /* regularisers, optimizers, gradient unpack, bytes */
  ST_TIME_CATEGORIES = 7
};
ST_API int st_timing_enable(int on);
ST_API int st_timing_read(int category, double* ms_total, double* work_total, uint64_t* n_scopes);
ST_API int st_timing_reset(void);

/* Number of kernels this library has launched on the calling process since load (bench.py's
 * "gpu_launches"). */
ST_API uint64_t st_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* STYLE_B200_H */
