/* dhg_b200 -- C ABI of the B200-native reverse-diffusion sampling engine.
 *
 * Drop-in boundary for the sampling hot path of
 * sleep3r/Diffusion-Handwriting-Generation.pytorch.  The reference has no FFI
 * layer; its boundary is the Python call surface.  Each entry
 * point below names the reference interface it stands behind.  The binding a
 * maintainer adds on the reference side is the ctypes stub in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; the message for
 *    the calling thread is dhg_last_error().  No exceptions cross the boundary.
 *  - "dev" pointers are CUDA device pointers on the ctx's device, "host"
 *    pointers are ordinary host memory.  All are BORROWED for the duration of
 *    the call (for stream-ordered calls: until the work enqueued on `stream`
 *    has completed); the library never frees or keeps them.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - one ctx per device; a ctx is not re-entrant: one host thread at a time.
 *  - stream rule: dhg_denoise / dhg_sample / dhg_sample_host all work on the
 *    plan's own buffers and captured graph.  Each call first makes its stream
 *    wait for the previous call on the same ctx (whatever stream that used) and
 *    records its own completion, so calls on different streams, or a
 *    stream-ordered call followed by dhg_sample_host, never overlap on those
 *    buffers.  The CALLER's buffers follow the usual rule: they must stay valid
 *    until the work enqueued on `stream` has completed.
 *  - there is no CPU fallback: every entry point that computes needs a CUDA
 *    device of compute capability 10.x and fails loudly otherwise.
 */
#ifndef DHG_B200_H
#define DHG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dhg_ctx dhg_ctx;

/* training_args.{att_layers_num, channels} of config.yml, read by the reference at
 * diffusion_handwriting_generation/checkpoint.py:280-286. */
typedef struct dhg_config {
  int32_t num_layers; /* att_layers_num */
  int32_t channels;   /* channels (c1); c2 = 3*c1/2, c3 = 2*c1.  Must be 128: the
                         reference hard-codes the 32-wide sigma embedding
                         (conditioning.py:9) */
} dhg_config;

#define DHG_PREC_FP32 0 /* the reference's precision (parity 1e-3 after 60 steps).  Default ("gemm" = 1): tcgen05 tensor cores on
                           split storage -- every fp32 activation / weight is a bf16 pair hi + lo and every GEMM is
                           x.w = hi.w_hi + lo.w_hi + hi.w_lo in bf16 MMAs with fp32 accumulation (~2^-17 per product), all
                           epilogue arithmetic in fp32.  "gemm" = 0: plain fp32 storage and CUDA-core fp32 FMA GEMMs. */
#define DHG_PREC_BF16 1 /* bf16 storage, tcgen05 bf16 MMA, fp32 accumulate/statistics */

#define DHG_MODE_NEW 0      /* utils/nn.py:90-112 new_diffusion_step (inference.py default) */
#define DHG_MODE_STANDARD 1 /* utils/nn.py:64-87 standard_diffusion_step */

#define DHG_NUM_STEPS 60 /* utils/nn.py:21 */

/* Message of the last failure on the calling thread ("" if none). */
const char* dhg_last_error(void);
/* ABI version of this library (bumped on any signature change). */
int32_t dhg_abi_version(void);

/* Replaces: DiffusionModel(num_layers, c1, c2, c3, drop) construction in
 * checkpoint.py:280-286 (load_model).  Needs a CUDA device. */
int32_t dhg_create(int32_t device, const dhg_config* cfg, dhg_ctx** out);
int32_t dhg_destroy(dhg_ctx* ctx);

/* The state_dict layout this ctx expects (model_final.pth, written by
 * train.py:131; 323 fp32 tensors for best_exp).  Enumerate to validate or drive a loader. */
int32_t dhg_num_weights(const dhg_ctx* ctx);
const char* dhg_weight_name(const dhg_ctx* ctx, int32_t index);
int32_t dhg_weight_ndim(const dhg_ctx* ctx, int32_t index);
int64_t dhg_weight_dim(const dhg_ctx* ctx, int32_t index, int32_t d);

/* Replaces: load_checkpoint()/load_state_dict(strict=True), checkpoint.py:92-130.
 * host_data: fp32, C-contiguous, `ndim` dims in `shape`.  Unknown name or shape
 * mismatch fails (strict). */
int32_t dhg_load_weight(dhg_ctx* ctx, const char* name, const float* host_data,
                        const int64_t* shape, int32_t ndim);
/* Optional: override the beta / alpha-bar schedule (60 fp32 each) with the values
 * the host computed (utils/nn.py:19-39, inference.py:81).  Default: computed here. */
int32_t dhg_set_schedule(dhg_ctx* ctx, const float* host_beta, const float* host_alpha_bar);
/* Strict check that every weight arrived; repacks weights for the kernels and
 * precomputes the per-step FiLM vectors for the 60 sampling noise levels.
 * Replaces model.to(device); model.eval() (checkpoint.py:295-296). */
int32_t dhg_finalize(dhg_ctx* ctx);

/* Allocates workspace and builds the launch sequence for a problem shape:
 * batch (chunk) B, T stroke points (multiple of 8, inference.py:78), L text
 * tokens, S style tokens of width 1280 (14 from StyleExtractor; text_style.py:59).
 * A later call replaces the previous plan. */
int32_t dhg_plan(dhg_ctx* ctx, int32_t B, int32_t T, int32_t L, int32_t S, int32_t precision);

/* Replaces: DiffusionModel.forward(strokes, text, sigma, style_vector) ->
 * (eps, pen_lifts, None), model.py:121-182.  All dev pointers, shapes of the
 * current plan: strokes [B,T,2] f32, text [B,L] i64, sigma [B] f32,
 * style [B,S,1280] f32 -> eps [B,T,2] f32, pen [B,T] f32.  Stream-ordered. */
int32_t dhg_denoise(dhg_ctx* ctx, const float* dev_strokes, const int64_t* dev_text,
                    const float* dev_sigma, const float* dev_style, float* dev_eps,
                    float* dev_pen, void* stream);

/* Replaces: the 60-step loop of infer(), inference.py:81-96, for `batch`
 * samples (any batch >= 1: processed in chunks of the planned B).
 * x0 [batch,T,2] f32; noise [60,batch,T,2] f32, noise[i] consumed at loop index i
 * (the draw of randn_like in utils/nn.py:86,111).  The draws are MANDATORY: the
 * library has no generator of its own (NULL fails); `seed` is reserved and
 * ignored.  text [batch,L] i64; style
 * [batch,S,1280] f32 -> out [batch,T,3] f32 = cat(x, pen_lifts of the last step).
 * Stream-ordered; the chain runs as one CUDA graph per chunk. */
int32_t dhg_sample(dhg_ctx* ctx, int32_t batch, const float* dev_x0, const float* dev_noise,
                   uint64_t seed, const int64_t* dev_text, const float* dev_style, int32_t mode,
                   float* dev_out, void* stream);
/* Same, with HOST buffers: the host->device copies of every input, the chain,
 * and the device->host copy of the result all happen inside the call, which
 * returns when `host_out` is complete.  When `batch` equals the planned B the
 * buffers are copied straight into the plan (no staging) and only the noise of
 * the first ten or so steps is waited for: the loop starts at index 59, so the
 * rest of the array travels on a second stream while those steps run (the chain
 * is then two CUDA graphs with the join between them; same kernels, same bits;
 * "host_overlap" = 0 restores the copy-everything-first path).  Pinned host
 * memory is what makes the copies asynchronous; pageable memory still works. */
int32_t dhg_sample_host(dhg_ctx* ctx, int32_t batch, const float* host_x0, const float* host_noise,
                        uint64_t seed, const int64_t* host_text, const float* host_style,
                        int32_t mode, float* host_out);

/* Stream-ordered calls cannot report bad input found on the device (a token id outside [0, 73), where the
 * reference's nn.Embedding raises IndexError): the kernel substitutes id 0 and raises a flag.  This call
 * synchronises `stream` and fails with "token id out of range" if the flag is set (and clears it).
 * dhg_sample_host checks its host ids before copying and the flag after its own synchronisation. */
int32_t dhg_check_errors(dhg_ctx* ctx, void* stream);

/* Replaces: new_diffusion_step / standard_diffusion_step (utils/nn.py:64-112)
 * with the noise draw injected: one fused streaming kernel over n = B*T*2 floats.
 * step = loop index i in [0,60).  dev_noise may be NULL (treated as zeros).
 * dev_out may alias dev_x. */
int32_t dhg_posterior_step(dhg_ctx* ctx, int32_t step, int32_t mode, const float* dev_x,
                           const float* dev_eps, const float* dev_noise, float* dev_out,
                           int64_t n, void* stream);

/* Number of kernels this library enqueued in the most recent dhg_denoise /
 * dhg_sample / dhg_sample_host call (graph nodes count individually). */
int64_t dhg_last_launch_count(const dhg_ctx* ctx);
/* Bytes of device memory held by the current plan. */
int64_t dhg_plan_bytes(const dhg_ctx* ctx);
/* Engine switches, mainly for tests and A/B measurements.  ALL of them are read when a plan is built (dhg_plan) and
 * stored in it: setting one afterwards has no effect on the current plan, its captured graphs or its launch count.
 * Per context:
 *   "gemm"  0 CUDA-core GEMM + row epilogue kernel, 1 tcgen05 GEMM with fused epilogue (default 1; in fp32 precision
 *           1 selects the split-storage tensor-core path, 0 the exact fp32 FMA path)
 *   "attn"  0 CUDA-core attention, 1 tcgen05 attention (bf16 storage, and split storage in fp32 precision; default 1)
 *   "graph" 0/1 one CUDA graph per chain in dhg_sample (default 1)
 * Process-wide (ctx may be NULL; shared by every context of the process):
 *   "host_overlap" 0/1 dhg_sample_host overlaps the noise copy with the first steps (default 1)
 *   "text_sets"    1..6 text sides of that many consecutive steps run at once on their own streams (default 2)
 *   "autotune"     0/1 time the GEMM tile configurations and the attention tile-load order at plan time (default 1)
 *   "serpentine"   0/1 alternate the row walking direction from kernel to kernel (default 1)
 *   "l2_hints"     0/1 L2 evict-first hint on streamed GEMM inputs (default 1)
 *   "tail_fusion"  0..3 chain only: 1 last fc + FiLM + skip + heads as one kernel on per-step folded tables, 2 also the
 *                  last block's conv2 / conv_skip in dot mode (no a2, skip, d1), 3 conv_skip folded onto the 3 head
 *                  channels (default 3)
 *   "head_fusion"  0/1 chain only: enc1.conv_skip(input_dense(x)) straight from x (default 1; unused while bit 0 of
 *                  "skip_fusion" is set)
 *   "skip_fusion"  bit mask 0..31, bf16 chain only: conv_skip of ConvBlock enc1 / enc2 / enc4 / dec3 / dec2 (bits 0..4) is
 *                  contracted inside the block's last GEMM (dual-operand launch, FiLM scale folded into per-step weights;
 *                  default 31)
 *   "attn_keyblock_auto" 0/1 the plan-time timing may also pick the key-block attention kernel for 128 < Tk <= 256
 *                  (default 0: that kernel sums in another order, so the BITS of a result would depend on which
 *                  variant won the timing for the batch a plan was built for; Tk > 256 always uses it)
 *   "w_resident", "specialize", "interleave", "pair", "pdl", "attn_early", "tune_bn", "tune_g", "tune_resident",
 *   "tune_pair", "tune_rev": kernel-selection overrides used by tests/test_gpu_kernel_variants.py and test_gpu_gemm.py
 *   "max_stages_a", "direct_store" (2: no epilogue at all, timing only), "split_n", "attn_max_slots", "attn_dbg": measurement switches of tools/ 
 * None of them changes results beyond the documented tolerances; all but the three fusions and "attn_keyblock_auto"
 * leave the bits unchanged.
 * Debugging aid (environment): DHG_SYNC_OPS=1 synchronises after every launch of the un-graphed paths (dhg_denoise,
 * "graph" = 0) and names the launch that faulted in dhg_last_error(). */
int32_t dhg_set_option(dhg_ctx* ctx, const char* key, int32_t value);

/* Test hook: copy a named intermediate activation of the last forward ("h1", "h2c",
 * "h2", "h3c", "h3", "att_in", "att0".., "d3", "d2", "d1", "text_act", ...) to host as
 * fp32 [B, positions, C] without halo rows.  host_out NULL: returns the element count.
 * Returns the element count, or -1 on failure.  Synchronises the device. */
int64_t dhg_debug_read(dhg_ctx* ctx, const char* name, float* host_out, int64_t capacity);

/* Test hook: run one tcgen05 GEMM (bf16 in, fused epilogue) against caller-provided
 * device buffers.  Not part of the reference-facing surface. */
int32_t dhg_debug_tc_gemm(int32_t device, const void* dev_a_bf16, int32_t lda, int32_t rows,
                          const void* dev_w_bf16, int32_t K, int32_t N, int32_t taps,
                          const float* dev_bias, void* dev_out_bf16, void* stream);

/* Test hook: the same GEMM with every fused-epilogue feature exposed (see csrc/common.cuh for the
 * order of operations).  All pointers are device pointers; null = feature off.  Row r of the flat
 * matrices maps to sample r / period, position r % period - pad_first; rows >= nvalid and (when
 * pad_first) rows with r % period == 0 are halo rows and are written as zeros. */
typedef struct dhg_debug_epilogue {
  const float* bias;       /* [N] */
  const void* rowbias;     /* bf16 [period - pad_first, rowbias_cols]: per-position term for columns < rowbias_cols */
  int32_t rowbias_cols;
  const void* res_pre;     /* bf16 [rows, res_pre_pitch] */
  int32_t res_pre_pitch;
  int32_t ln;              /* LayerNorm over N, eps 1e-6 */
  const float* gamma;      /* FiLM: gamma[b * film_bstride + n] */
  const float* beta;
  int32_t film_bstride;
  const void* res_post;    /* bf16 */
  int32_t res_post_pitch;
  int32_t res_post_up;     /* res_post lives one level down: row b*period_lo + 1 + pos/2 */
  int32_t res_post_period_lo;
  void* out_raw;           /* bf16 [rows, out_raw_pitch] */
  int32_t out_raw_pitch;
  void* out_act;           /* bf16 SiLU(value) */
  int32_t out_act_pitch;
  int32_t period, pad_first, nvalid;
  /* dot mode (instead of out_raw / out_act): dot_out[row * 4 + j] = <value or SiLU(value), dot_w[j, :]>, j < 3 */
  const float* dot_w;      /* fp32 [3, N] or NULL */
  float* dot_out;          /* fp32 [rows, 4] */
  int32_t dot_act;
  /* split I/O (the fp32-contract mode): every activation operand (A, res_pre, res_post, rowbias, outputs) is in split
   * storage: value = hi + lo, two bf16, rows laid out in groups of 32 elements (32 hi, then 32 lo: 128 bytes); pitches
   * and column counts stay in elements (multiples of 32); the caller passes K = 2 * (elements per A row), lda in bf16
   * units and W as [taps][N][K] with K grouped the same way (w_hi x 32 | w_lo x 32). */
  int32_t split_io;
  /* dual-operand mode: out = A . W[w_row_off : w_row_off + N]^T + sum_tap A2[row + tap - 1] . W2[tap]^T + bias, one
   * accumulation (taps must be 1; bias-only epilogue).  W is a stack of dual_w1_rows / N variants of [N, K]; A2 is
   * bf16 [rows, dual_K2] (pitch dual_lda2) in the same padded-row layout, W2 bf16 [3][N][dual_K2].  NULL dual_a2 = off. */
  const void* dual_a2;
  int32_t dual_lda2, dual_K2;
  const void* dual_w2;
  int32_t dual_w1_rows, w_row_off;
} dhg_debug_epilogue;
int32_t dhg_debug_tc_gemm_ex(int32_t device, const void* dev_a_bf16, int32_t lda, int32_t rows,
                             const void* dev_w_bf16, int32_t K, int32_t N, int32_t taps,
                             const dhg_debug_epilogue* epi, int32_t repeats, float* ms_per_launch,
                             void* stream);

/* Test hook: one attention launch over caller-provided bf16 row matrices (head h = columns
 * [h*D, h*D+D)); row(b, t) = b*period + pad + t.  impl 0 = CUDA-core kernel, 1 = tcgen05 kernel
 * (all keys at once for Tk <= 256, key blocks beyond), 2 = tcgen05 key-block kernel (unmasked, D = 64, Tk > 128),
 * 3 = tcgen05 kernel on split storage (fp32-contract mode: q / k / v / o in groups of 32 hi | 32 lo bf16, pitches in
 * elements, D = 64, Tk <= 256). */
typedef struct dhg_debug_attn {
  const void* q; const void* k; const void* v; void* o;
  int32_t q_pitch, k_pitch, v_pitch, o_pitch;
  int32_t q_period, q_pad, k_period, k_pad;
  int32_t B, H, D, Tq, Tk;
  int32_t q_rows, k_rows;   /* total rows of the q / k,v matrices */
  const int64_t* text;      /* [B, Tk] token ids (0 = masked key) or NULL */
} dhg_debug_attn;
/* Test / measurement hook: time the text side of a step (TextStyleEncoder + text_dense / kv of every
 * EncoderLayer) of the current plan, `sets` steps at once (1 .. the plan's text sets), as ms per step. */
int32_t dhg_debug_time_text(dhg_ctx* ctx, int32_t sets, int32_t repeats, float* ms_per_step);

int32_t dhg_debug_attention(int32_t device, const dhg_debug_attn* a, int32_t impl, int32_t repeats,
                            float* ms_per_launch, void* stream);

/* ---- StyleExtractor: the step right before the sampling path ----------------------------------
 * Replaces: StyleExtractor.forward, text_style.py:43-59 -- torchvision MobileNetV2 `features` in eval mode on the grey
 * writer image (x / 127.5 - 1, repeated to 3 channels), AvgPool2d(3, 3), AdaptiveAvgPool2d((1, 14)) -> [B, 14, 1280].
 * Weights: the `features.*` entries of a torchvision mobilenet_v2 state_dict (conv weights and BatchNorm weight / bias /
 * running_mean / running_var; other keys are ignored), fp32 host pointers; BatchNorm is folded at finalize.
 * dhg_style_extract: host_img fp32 [B, H, W] grey levels 0..255 (read_img(path, 96), utils/io.py:98-115, gives H = 96)
 * -> dev_out fp32 [B, 14, 1280] on the device; stream-ordered.  Errors: dhg_style_last_error(). */
typedef struct dhg_style dhg_style;
const char* dhg_style_last_error(void);
int32_t dhg_style_create(int32_t device, dhg_style** out);
int32_t dhg_style_destroy(dhg_style* s);
int32_t dhg_style_load_weight(dhg_style* s, const char* name, const float* host_data, const int64_t* shape, int32_t ndim);
int32_t dhg_style_finalize(dhg_style* s);
int32_t dhg_style_extract(dhg_style* s, const float* host_img, int32_t B, int32_t H, int32_t W, float* dev_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Training step, part 1: everything of train.py:26-67 around the model call and loss.backward() (SURVEY.md 8e "optional
 * train step", 8f-3).  Replaces, on device buffers:
 *   dhg_train_perturb    train.py:38-43   x_perturbed = sqrt(alpha) x + sqrt(1 - alpha) eps      (x, eps, out [B, T, 2]; alphas [B])
 *   dhg_train_loss       loss.py:5-39     losses[3] = {total, score_loss, pen_lifts_loss} and, when the pointers are
 *                                         non-NULL, d total / d score_pred [B, T, 2] and d total / d pen_lifts_pred [B, T]
 *   dhg_train_sqnorm     utils/clip_grad.py:27-46 mode "norm" (torch clip_grad_norm_): sum of squares of the flat gradient
 *   dhg_train_adam_step  scheduler.py:1-35 + torch.optim.Adam (config.yml:33-38): one update of the flat parameter buffer;
 *                        the gradient that enters is grad * (1 / world_size) * min(1, max_norm / (||grad|| / world_size + 1e-6))
 *                        with ||grad||^2 read from dev_sqnorm (NULL: no clipping), so a data-parallel caller all-reduces
 *                        (sums) the flat gradient once and never rescales it; `lr` is the scheduled rate of this step
 *                        (dhg_b200.train.InvSqrtSchedule), `step` counts from 1.
 * dev_scratch: dhg_train_scratch_doubles() doubles.  Everything is stream-ordered, nothing synchronises; reductions are
 * deterministic (fixed partial sums).  The flat gradient is an input here; dhg_trainer_backward (below) produces it.
 * Errors: dhg_train_last_error(). */
const char* dhg_train_last_error(void);
int32_t dhg_train_scratch_doubles(void);
int32_t dhg_train_perturb(int32_t device, const float* dev_x, const float* dev_alphas, const float* dev_eps, float* dev_out, int32_t B,
                          int32_t T, void* stream);
int32_t dhg_train_loss(int32_t device, const float* dev_eps, const float* dev_score_pred, const float* dev_pen_lifts,
                       const float* dev_pen_lifts_pred, const float* dev_alphas, int32_t B, int32_t T, float* dev_losses,
                       float* dev_grad_score, float* dev_grad_pen_pred, double* dev_scratch, void* stream);
int32_t dhg_train_sqnorm(int32_t device, const float* dev_grad, int64_t n, double* dev_out, double* dev_scratch, void* stream);
int32_t dhg_train_adam_step(int32_t device, float* dev_param, const float* dev_grad, float* dev_exp_avg, float* dev_exp_avg_sq, int64_t n,
                            int32_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                            const double* dev_sqnorm, double max_norm, int32_t world_size, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Forward + backward pass of the denoiser for the training step (csrc/train_step.cu).  Replaces
 *   train.py:46-55   strokes_pred, pen_lifts_pred, _ = model(x_perturbed, text, sqrt(alphas), style);  loss.backward()
 * The parameters live in ONE flat fp32 device buffer in checkpoint key order (dhg_trainer_param_info enumerates
 * name / offset / numel; Linear [out, in], Conv1d [out, in, 3] exactly as in model_final.pth) and the gradients in a
 * second buffer of the same layout -- the buffers dhg_train_sqnorm / dhg_train_adam_step work on.  Both are borrowed
 * for the life of the trainer.  A trainer is a plan for one (B, T, L); T a multiple of 8 (inference.py:78).
 *   dhg_trainer_forward   dev_x [B, T, 2] (x_perturbed), dev_text int64 [B, L] (0 = padding), dev_sigma [B]
 *                         (sqrt(alphas)), dev_style [B, 14, 1280], dev_style_keep [B, 14, 1280] or NULL: the keep mask
 *                         of text_style.py:83,92 Dropout(0.3) already divided by 0.7 (NULL = eval mode) ->
 *                         dev_score_pred [B, T, 2], dev_pen_pred [B, T] (after the sigmoid); every activation is kept
 *   dhg_trainer_backward  dev_grad_score [B, T, 2], dev_grad_pen_pred [B, T] (what dhg_train_loss writes) -> the flat
 *                         gradient buffer is zeroed and filled.  Weight gradients are summed with fp32 atomics: the
 *                         last bits depend on the order.
 * Stream-ordered, nothing synchronises, no allocation after create (graph-capturable).  fp32 on the CUDA cores.
 * Stream rule: a trainer owns its activation / gradient arenas and one internal second stream (weight gradients and skip
 * branches run there, forked from and joined back into the caller's stream inside every call), so all calls on one
 * trainer must be issued on ONE stream (or be ordered by the caller); one host thread per trainer; the parameter
 * buffer may be updated between a backward and the next forward on that same stream (dhg_train_adam_step does).
 * Errors: dhg_trainer_last_error(). */
typedef struct dhg_trainer dhg_trainer;
const char* dhg_trainer_last_error(void);
int64_t dhg_trainer_param_count(int32_t num_layers, int32_t channels);
int32_t dhg_trainer_param_info(int32_t num_layers, int32_t channels, int32_t index, char* name_out, int32_t name_cap, int64_t* offset,
                               int64_t* numel);   /* 0 ok, -1 index past the end */
int32_t dhg_trainer_create(int32_t device, int32_t num_layers, int32_t channels, int32_t B, int32_t T, int32_t L, float* dev_params,
                           float* dev_grads, dhg_trainer** out);
int32_t dhg_trainer_destroy(dhg_trainer* t);
int64_t dhg_trainer_workspace_bytes(const dhg_trainer* t);
int64_t dhg_trainer_last_launches(const dhg_trainer* t);   /* launches of the last forward (+ backward) */
int32_t dhg_trainer_set_option(const char* name, int32_t value);   /* "tiled_gemm": 1 (default) tiled fp32 GEMM on the CUDA cores, 3 tiled with
                                                                       3 x TF32 tensor-core products (same fp32 contract), 4 plain TF32 products
                                                                       (torch's allow_tf32; about 1e-3, outside the fp32 contract), 2 smallest tile
                                                                       only, 0 per-thread kernel bodies (checks); "side_stream": 1 (default) / 0
                                                                       everything on the caller's stream.  Process-wide measurement switches. */
int32_t dhg_trainer_forward(dhg_trainer* t, const float* dev_x, const int64_t* dev_text, const float* dev_sigma, const float* dev_style,
                            const float* dev_style_keep, float* dev_score_pred, float* dev_pen_pred, void* stream);
int32_t dhg_trainer_backward(dhg_trainer* t, const float* dev_grad_score, const float* dev_grad_pen_pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DHG_B200_H */
