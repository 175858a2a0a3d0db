// Kernel launch interface shared by the engine and the kernel translation units.
#pragma once
#include "common.cuh"

namespace dhg {

struct AttnParams {
  const void* q; const void* k; const void* v; void* o;   // activation dtype
  int q_pitch, k_pitch, v_pitch, o_pitch;                  // elements per row
  int q_period, q_pad;                                     // row(b, t) = b*q_period + q_pad + t
  int k_period, k_pad;
  int B, H, D, Tq, Tk;
  float scale;                                             // 1/sqrt(D)
  const int64_t* text;                                     // [B, Tk] token ids (0 = masked key) or null
  int split = 0;                                           // tcgen05 kernel: q / k / v / o are in split storage (common.cuh bfs)
};

struct HeadParams {
  int B, T;
  float* eps_out;          // [B,T,2] or null
  float* pen_out;          // pen_out[i*pen_stride + pen_offset] or null
  int pen_stride, pen_offset;
  float* x_io;             // [B,T,2] current x (null: no posterior update)
  float* x_out;            // where the updated x goes (null: in place); row stride x_out_stride
  int x_out_stride;
  const float* noise;      // [B,T,2] injected draw or null (= zero)
  int mode;                // 0 "new", 1 "standard"
  float c_eps, c_eps2, c_div, c_noise;
  // fused input_dense of the NEXT denoiser step (model.py:139 applied to the updated x): padded-row
  // [rows, C] raw and SiLU'd copies, activation dtype; null = not fused
  void* next_raw;
  void* next_act;
  const float* in_W;       // [C, 2]
  const float* in_b;       // [C]
  // second input row matrix with its own [3, C] weights (tail fusion, engine.cu: the last ConvBlock's fc + FiLM +
  // skip and the two heads collapse into  a2 . A_step + skip . H + c_step); null = single input
  const void* h2;
  const float* w2;
};

template <typename TA>
void launch_gemm_simt(const TA* A, int lda, int rows, const float* W, int K, int N, int taps,
                      float* out, cudaStream_t st);
template <typename T>
void launch_rowpost(const float* acc, int rows, int N, const Epilogue& e, cudaStream_t st);
template <typename T>
int launch_attention_simt(const AttnParams& p, cudaStream_t st);
template <typename T>
void launch_film_rows(const T* in, T* out, int rows, int C, int period, const float* gamma,
                      const float* beta, int bstride, cudaStream_t st);
template <typename T>
void launch_pool(const T* in, T* out_raw, T* out_act, int B, int Tlo, int C, cudaStream_t st);
template <typename T>
void launch_input_dense(const float* x, const float* W, const float* bias, T* out_raw, T* out_act,
                        int B, int Tn, int C, cudaStream_t st);
template <typename T>
int launch_heads_from_dots(const float* dot_a, const float* dot_b, const float* cst, int C, const HeadParams& p, cudaStream_t st);
template <typename T>
int launch_skip_from_x(const float* x, const float* M, const float* v, const float* bsk, T* out, int B, int Tn, int C, cudaStream_t st);
template <typename T>
int launch_heads_update(const T* h, int C, const float* Wo, const float* bo, const float* Wp,
                         const float* bp, const HeadParams& p, cudaStream_t st);
void launch_posterior(const float* x, const float* eps, const float* z, float* out, size_t n, int mode,
                      float c_eps, float c_eps2, float c_div, float c_noise, int num_sms, cudaStream_t st);
void launch_sigma_ffn(const float* sigma, const float* w1, const float* b1, const float* w2, const float* b2,
                      int H, float* out, int n, cudaStream_t st);
void launch_film_table(const float* sig, const float* Wc, const float* bc, int tot, float* out, int n, cudaStream_t st);
template <typename T>
void launch_embed_ln(const int64_t* ids, const float* emb, int vocab, int C, T* out, int rows, int* err, cudaStream_t st);
template <typename T>
void launch_silu_convert(const float* in, T* out, size_t n, cudaStream_t st);

// tcgen05 / TMA GEMM (gemm_tc.cu).  A: bf16 [rows, K] pitch lda; W: bf16 [taps][N][K]
// (K contiguous).  Epilogue fused.  Returns 0 on success.
struct TcGemmPlan;  // opaque: tensor maps + launch geometry, built once at plan time
// Tile configuration of a plan; -1 = the built-in rule.  bn: tile width (divides N; LayerNorm rows always use N);
// g: independent accumulators per super-tile (g * bn <= 256); resident: keep W in shared memory for the whole launch
// (if it fits); pair: 1 = cta_group::2 CTA pairs (only with streamed W and g == 1), 2 = column-split LayerNorm (a 2-CTA
// cluster per row tile, each CTA accumulates half of the row's columns double-buffered, LayerNorm statistics exchanged
// through distributed shared memory; N = 128, 256 or 384).  Every configuration computes the same bits: the K order of
// a row's dot product and the grouping of the LayerNorm partial sums do not depend on it.
struct TcTune { int bn, g, resident, pair; };
// Dual-operand GEMM: out = A . W^T + sum_tap A2[row + tap - 1] . W2[tap]^T + bias, one accumulation over two operand
// segments (the first with 1-tap weights, the second with 3-tap weights and the zero halo of the padded-row layout).
// W (the first operand's weights) may be a stack of w1_rows / N variants of [N, K]; the launch picks one through
// Epilogue::w_row_off (a multiple of N).  Used for  FiLM(fc(a2)) + conv_skip(x)  with the FiLM scale folded into fc's
// weights per sampling step (engine.cu).
struct TcDual { const bf16* A2; int lda2; int K2; const bf16* W2; int w1_rows; };
TcGemmPlan* tc_gemm_plan_create(const bf16* A, int lda, int rows, const bf16* W, int K, int N, int taps,
                                const Epilogue& e, char* err, int errlen, const TcTune* tune = nullptr, const TcDual* dual = nullptr);
void tc_gemm_plan_config(const TcGemmPlan*, TcTune* out);   // what the plan ended up with
// Walk the row tiles from the last to the first.  The engine gives every kernel the direction opposite to the kernel
// that wrote its input, so it starts on the rows that are still in the 126 MB L2 (same bits either way).
void tc_gemm_plan_set_reverse(TcGemmPlan*, int rev);
// A rows get L2 evict-first priority (only applied when every A row is read once, i.e. one column group).
void tc_gemm_plan_set_a_evict_first(TcGemmPlan*, int on);
void tc_gemm_plan_destroy(TcGemmPlan*);
void tc_gemm_set_trace(TcGemmPlan*, unsigned long long* buf, int cap);   // debug timeline of CTA 0
void tc_gemm_describe(const TcGemmPlan*, char* out, int n);
int tc_gemm_launch(const TcGemmPlan*, const Epilogue& e, cudaStream_t st);
void tc_gemm_set_option(int which, int value);   // 2 w_resident, 3 specialize, 4 interleave, 6 pdl, 7 pair; 10..13 forced TcTune {bn, g, resident, pair}

// tcgen05 attention for head depth 64 / 48 (attention_tc.cu): all keys at once for Tk <= 256, key blocks beyond.  q_rows / k_rows: total rows
// of the q / k,v row matrices (TMA bounds).
struct AttnTcPlan;
bool attn_tc_supported(const AttnParams& p);
// prefer_long: use the key-block kernel (two-pass softmax over blocks of 128 keys; always used for Tk > 256) also for
// 128 < Tk <= 256 (unmasked, head depth 64); the engine times both at plan time
AttnTcPlan* attn_tc_plan_create(const AttnParams& p, int q_rows, int k_rows, char* err, int errlen, int prefer_long = 0);
void attn_tc_plan_set_trace(AttnTcPlan*, unsigned long long* buf, int cap);   // debug timeline of CTA 0, slot 0
int attn_tc_plan_slots(const AttnTcPlan*);
bool attn_tc_plan_is_long(const AttnTcPlan*);
void attn_tc_plan_destroy(AttnTcPlan*);
void attn_tc_plan_set_reverse(AttnTcPlan*, int rev);   // work items from the last sample to the first
void attn_tc_plan_set_early_load(AttnTcPlan*, int on);   // next item's Q/K/V requested right after P V (default) or after O is stored
int attn_tc_launch(const AttnTcPlan*, cudaStream_t st);
void attn_tc_set_debug(int flags);   // timing experiments only

}  // namespace dhg
