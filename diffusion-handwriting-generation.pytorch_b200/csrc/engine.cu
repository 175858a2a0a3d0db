// Engine + C ABI (include/dhg_b200.h) for the reverse-diffusion sampling path.
//
// What is rebuilt here, and where the behaviour comes from (reference file:line):
//   DiffusionModel.forward        diffusion_handwriting_generation/model.py:121-182
//   EncoderLayer.forward          model.py:35-58
//   ConvBlock.forward             cnn.py:52-87
//   MultiHeadAttention / SDPA     attention.py:26-87,  PosEmbeddings attention.py:15-23
//   AffineTransformLayer (FiLM)   conditioning.py:16-19
//   TextStyleEncoder.forward      text_style.py:91-104
//   beta schedule / updates       utils/nn.py:19-39, 64-112
//   the 60-step loop              inference.py:81-96
//   state_dict layout             checkpoint.py:92-130, train.py:131
//
// The forward pass is compiled once per problem shape ("plan") into a flat list of
// kernel launches over pre-allocated HBM buffers; the whole 60-step chain is
// captured into one CUDA graph.  Restructurings relative to the reference (all
// algebraically identical):
//   * channels-last padded-row layout: no transposes, conv = 3 row-shifted GEMMs;
//   * the 76 FiLM linears and sigma_ffn depend only on the noise level: one
//     [60, 18560] table at load time for sampling, one tiny kernel pair for
//     dhg_denoise;
//   * positional embeddings are constants: PE @ W folds into a per-position bias
//     table of the q/k projections;
//   * LN(style_ffn(style)) and LN(emb(text)) do not depend on the step: hoisted;
//   * SiLU is applied by the producer's epilogue (raw and/or activated copy).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/dhg_b200.h"
#include "kernels.h"

using namespace dhg;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int g_opt_autotune = 1;  // time every GEMM tile configuration at plan time and keep the fastest ("autotune")
static int g_opt_l2_hints = 1;   // streamed GEMM inputs get L2 evict-first priority ("l2_hints")
static int g_opt_head_fusion = 1; // chain: enc1.conv_skip(input_dense(x)) computed from x by a K = 6 kernel ("head_fusion")
static int g_opt_tail_fusion = 3; // chain: 1 last fc + FiLM + skip + heads as one kernel on folded tables, 2 also conv2 / conv_skip of the
                                  // last block in dot mode, so neither a2 nor skip nor d1 exist ("tail_fusion")
static int g_opt_skip_fusion = 31; // chain: conv_skip folded into the last GEMM of a ConvBlock (bit i: enc1, enc2, enc4, dec3, dec2) ("skip_fusion")
static int g_opt_attn_keyblock_auto = 0; // plan-time timing may pick the key-block attention kernel for 128 < Tk <= 256 ("attn_keyblock_auto")
static int g_opt_serpentine = 1; // consumer kernels walk their rows opposite to their producer ("serpentine")
static int g_opt_host_overlap = 1; // dhg_sample_host: noise of the later steps travels while the first steps run ("host_overlap")
static int g_opt_text_sets = 2;  // text sides of this many consecutive steps run at once (dhg_set_option "text_sets")
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));     \
  } while (0)

namespace {

constexpr int kSigmaDim = 32;
constexpr int kSigmaHidden = 2048;
constexpr int kVocab = 73;
constexpr int kStyleSplit = 5;
constexpr int kStyleWidth = 1280;

struct WSpec {
  std::string name;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// Packed weight of one GEMM: `taps` K x N slabs.
struct Lin {
  int K = 0, N = 0, taps = 1;
  float* w32 = nullptr;   // [taps][K][N] exact fp32               (fp32 mode, CUDA-core GEMM)
  float* w32r = nullptr;  // [taps][K][N] bf16-rounded, as fp32    (bf16 mode, CUDA-core GEMM)
  bf16* w16 = nullptr;    // [taps][N][K] bf16, K contiguous       (bf16 mode, tcgen05 GEMM)
  bf16* w16s = nullptr;   // [taps][N][2K] bf16, K in groups of 32: (w_hi x 32 | w_lo x 32), w = w_hi + w_lo
                          // (fp32 mode on tcgen05: split storage, common.cuh bfs)
  float* bias = nullptr;  // [N]
  std::vector<float> h_w;  // host copy [taps][K][N] (for PE-folded bias tables)
  std::vector<float> h_b;
};

constexpr int kMaxTextSets = 6;

struct StepCtx {
  const float* cond;  // FiLM vectors: cond[b * bstride + off]
  int bstride;
  bool skip_input_dense;  // in_raw / in_act were already written by the previous step's head kernel
  bool fuse_next_input;   // this step's head kernel also writes in_raw / in_act of the next step
  int text_set;           // which copy of the text-side buffers this step's cross-attention reads
  bool fuse_head;         // enc1.conv_skip is computed from x (skip_from_x), the head kernel does not write in_raw
  int fuse_tail;          // 1: dec1.fc is not launched, the head kernel works on (a2, skip) with the step's folded tables;
                          // 2: dec1.conv2 / conv_skip run in dot mode and the head kernel only adds up their 3 dots per point
  int fuse_skip;          // bit mask of ConvBlocks whose conv_skip is folded into their last GEMM (Plan::opt_skip_fusion; chain only)
  int step;               // sampling step index (tail tables)
  HeadParams head;
};
// One launch (or a short fixed group of launches) of the plan; `name` is for DHG_SYNC_OPS=1 fault localisation.
struct Op {
  std::function<int(cudaStream_t, const StepCtx&)> fn;
  std::string name;
  template <typename F>
  Op(F f) : fn(std::move(f)) {}
  int operator()(cudaStream_t st, const StepCtx& sc) const { return fn(st, sc); }
};

struct Act {
  void* p = nullptr;
  int rows = 0, C = 0;
};

// Activation storage of a plan -> element type of the CUDA-core kernels: fp32, bf16, or the bfs split pairs.
template <typename F>
inline auto by_storage(int prec, bool split, F&& f) {
  if (prec == PREC_BF16) return f((bf16*)nullptr);
  if (split) return f((bfs*)nullptr);
  return f((float*)nullptr);
}
#define DHG_STORAGE(Pl, T, ...) by_storage((Pl)->prec, (Pl)->split, [&](auto* tag_) { using T = std::remove_pointer_t<decltype(tag_)>; __VA_ARGS__ })

struct EpiSpec {
  bool bias = true;
  const float* rowbias = nullptr;
  const void* rowbias16 = nullptr;
  int rowbias16_cols = 0;
  Act res_pre;
  bool ln = false;
  int film_off = -1;
  Act res_post;
  bool res_post_up = false;
  int res_post_period_lo = 0;
  Act out_raw, out_act;
  // dot mode (tail fusion level 2): 3 dot products per row instead of the stored row
  float* dot_out = nullptr;
  const float* dot_w = nullptr;   // constant [3, N] vectors, or
  bool dot_w_per_step = false;    //   ctx->tail_A + step * 3 * N at launch
  bool dot_act = false;
  std::string alt_wkey;           // the dot-mode twin uses this packed weight instead (a conv folded onto its 3 readers)
};

struct Plan {
  int B = 0, T = 0, L = 0, S = 0, SP = 0, prec = 0, gemm_impl = 0;
  bool split = false;   // fp32 precision on the tensor cores: activations stored as bfs pairs (common.cuh), every GEMM as two
                        // bf16 tcgen05 GEMMs over the split row (hi + lo against w_hi, hi against w_lo), fp32 accumulate
  bool tc() const { return gemm_impl == 1; }   // GEMMs run on tcgen05 (bf16 storage, or split storage in fp32 precision)
  int Tl[4], R[4], RT = 0, RS = 0;
  size_t esize = 4;
  std::vector<void*> allocs;
  size_t bytes = 0;
  // staging (plan-owned so the captured graph is static)
  int64_t* text = nullptr;
  float* style = nullptr;
  float* x_state = nullptr;
  float* noise = nullptr;     // [60,B,T,2], allocated on first injected-noise use
  float* out = nullptr;       // [B,T,3]
  float* sigma_in = nullptr;  // [B]
  float* sig_emb = nullptr;   // [B,32]
  float* cond_b = nullptr;    // [B,tot] (dhg_denoise)
  float* scratch = nullptr;   // fp32 accumulators of the CUDA-core GEMM path
  size_t scratch_elems = 0;
  int* err_flag = nullptr;    // set by embed_ln_kernel on a token id outside [0, vocab); read at the synchronising entry points
  // dhg_set_option switches as they were when the plan was built: the chain (and its cached graphs) only ever see these
  int opt_tail_fusion = 0, opt_head_fusion = 0, opt_skip_fusion = 0;
  // Every entry point works on the plan's own buffers and graph.  `ev_last` is recorded at the end of each call on the
  // stream it used and waited for at the start of the next one, so calls on different streams are ordered.
  cudaEvent_t ev_last = nullptr;
  // once_ops: step-independent; text_ops[set]: the part of a step that depends on sigma but not on x (TextStyleEncoder,
  // text_dense and k/v projections of every EncoderLayer); step_ops: everything that depends on x.  The text side is
  // small (192 row tiles on 148 SMs per launch), so the text sides of `text_sets` consecutive steps are run at once,
  // one stream each, before the first of those steps: their launches fill each other's last waves.
  std::vector<Op> once_ops, text_ops[kMaxTextSets], step_ops;
  int text_sets = 1;
  cudaStream_t text_stream[kMaxTextSets] = {};   // [0] unused: set 0 runs on the caller's stream
  cudaEvent_t ev_fork = nullptr, ev_text[kMaxTextSets] = {};
  Act head_in, in_raw, in_act, tail_a2, tail_skip;
  float *tail_dot_a2 = nullptr, *tail_dot_skip = nullptr;   // [R0, 4] fp32 each (tail fusion level 2)
  std::vector<TcGemmPlan*> tc_plans;
  std::vector<AttnTcPlan*> attn_plans;
  int attn_impl = 1;
  cudaStream_t cap_stream = nullptr;
  cudaGraphExec_t graphs[4] = {nullptr, nullptr, nullptr, nullptr};  // [mode*2 + has_noise]
  // dhg_sample_host: the chain as two graphs ([mode][part]; part 0 = the first head_steps steps) so that most of the noise can
  // travel from the host while part 0 runs; copy_stream carries that copy, ev_copy_go / ev_copy_done fork and join it
  cudaGraphExec_t host_graphs[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy_go = nullptr, ev_copy_done = nullptr;
  int64_t launches_once = 0, launches_step = 0, launches_text = 0;
  struct HostStage { float *x = nullptr, *noise = nullptr, *style = nullptr, *out = nullptr; int64_t* text = nullptr; int cap = 0; };
  HostStage stage;   // device copies of dhg_sample_host's host buffers
  struct Tap { Act a; int period; int pad; };
  std::map<std::string, Tap> taps;  // named activations readable through dhg_debug_read
  std::map<const void*, int> dir_of;   // walking direction of the kernel that wrote each activation (Builder::dir_of)
};

}  // namespace

struct dhg_ctx {
  int device = 0;
  int num_sms = 148;
  dhg_config cfg;
  int c1, c2, c3, d;
  std::vector<WSpec> spec;
  std::map<std::string, int> spec_index;
  std::map<std::string, std::vector<float>> raw;
  bool finalized = false;
  float beta[DHG_NUM_STEPS], abar[DHG_NUM_STEPS];
  std::map<std::string, Lin> lins;
  std::map<std::string, int> film_off;
  int film_total = 0;
  float* film_W = nullptr;  // [tot,32]
  float* film_b = nullptr;  // [tot]
  float *sff_w1 = nullptr, *sff_b1 = nullptr, *sff_w2 = nullptr, *sff_b2 = nullptr;
  float* cond60 = nullptr;  // [60, tot]
  // tail fusion (see finalize): eps|pen = a2 . tail_A[step] + skip . tail_H + tail_c[step]
  float *tail_A = nullptr, *tail_c = nullptr, *tail_H = nullptr;   // [60,3,C] | [60,3] | [3,C]
  // skip fusion: per block, fc weights with the step's FiLM scale folded in, [60 * N][K] bf16, and the matching bias
  // gamma_s * b_fc + beta_s + b_skip, [60][N] fp32 (dhg_finalize)
  std::map<std::string, bf16*> fc60_w;
  std::map<std::string, float*> fc60_bias;
  float* tail_pick = nullptr;   // [3, 32] one-hot rows: dot mode reading columns 0..2 of the folded conv_skip
  // head fusion: enc1.conv_skip(input_dense(x)) = sum_tau (x[t+tau] . head_M[tau] + head_v[tau]) + head_b
  float *head_M = nullptr, *head_v = nullptr, *head_b = nullptr;   // [3,2,C] | [3,C] | [C]
  float* emb = nullptr;     // [73, d]
  float *in_W = nullptr, *in_b = nullptr, *out_W = nullptr, *out_b = nullptr, *pen_W = nullptr, *pen_b = nullptr;
  std::vector<void*> allocs;
  Plan* plan = nullptr;
  int opt_gemm = 1, opt_graph = 1, opt_attn = 1;
  int64_t last_launches = 0;
};

// ---------------------------------------------------------------------------
// state_dict spec (mirrors oracle/dhg_oracle.py:state_dict_spec)
// ---------------------------------------------------------------------------
namespace {

void spec_lin(std::vector<WSpec>& s, const std::string& p, int din, int dout) {
  s.push_back({p + ".weight", {dout, din}});
  s.push_back({p + ".bias", {dout}});
}
void spec_conv(std::vector<WSpec>& s, const std::string& p, int din, int dout) {
  s.push_back({p + ".weight", {dout, din, 3}});
  s.push_back({p + ".bias", {dout}});
}
void spec_affine(std::vector<WSpec>& s, const std::string& p, int h) {
  spec_lin(s, p + ".gamma_emb", kSigmaDim, h);
  spec_lin(s, p + ".beta_emb", kSigmaDim, h);
}
void spec_mha(std::vector<WSpec>& s, const std::string& p, int d) {
  for (const char* n : {"wq", "wk", "wv", "dense"}) spec_lin(s, p + "." + n, d, d);
}
void spec_convblock(std::vector<WSpec>& s, const std::string& p, int din, int dout) {
  spec_affine(s, p + ".affine1", dout / 2);
  spec_affine(s, p + ".affine2", dout);
  spec_affine(s, p + ".affine3", dout);
  spec_conv(s, p + ".conv_skip", din, dout);
  spec_conv(s, p + ".conv1", din, dout / 2);
  spec_conv(s, p + ".conv2", dout / 2, dout);
  spec_lin(s, p + ".fc", dout, dout);
}
void spec_enc(std::vector<WSpec>& s, const std::string& p, int din, int dout) {
  spec_lin(s, p + ".text_dense", din, dout);
  spec_lin(s, p + ".ffn.1", dout, 2 * dout);
  spec_lin(s, p + ".ffn.3", 2 * dout, dout);
  spec_mha(s, p + ".mha", dout);
  spec_mha(s, p + ".mha2", dout);
  for (int i = 0; i < 4; ++i) spec_affine(s, p + ".affine" + std::to_string(i), dout);
}

std::vector<WSpec> build_spec(int num_layers, int ch) {
  const int c1 = ch, c2 = ch * 3 / 2, c3 = ch * 2, d = 2 * c2;
  std::vector<WSpec> s;
  spec_lin(s, "input_dense", 2, c1);
  spec_lin(s, "sigma_ffn.1", 1, kSigmaHidden);
  spec_lin(s, "sigma_ffn.3", kSigmaHidden, c1 / 4);
  spec_convblock(s, "enc1", c1, c1);
  spec_convblock(s, "enc2", c1, c2);
  spec_enc(s, "enc3", d, c2);
  spec_convblock(s, "enc4", c2, c3);
  spec_enc(s, "enc5", d, c3);
  spec_conv(s, "skip_conv1", c1, c2);
  spec_conv(s, "skip_conv2", c2, c3);
  spec_conv(s, "skip_conv3", c3, d);
  s.push_back({"text_style_model.emb.weight", {kVocab, d}});
  spec_lin(s, "text_style_model.style_ffn.1", kStyleWidth / kStyleSplit, 4 * c2);
  spec_lin(s, "text_style_model.style_ffn.3", 4 * c2, d);
  spec_lin(s, "text_style_model.text_ffn.1", d, 2 * d);
  spec_lin(s, "text_style_model.text_ffn.3", 2 * d, d);
  spec_mha(s, "text_style_model.mha", d);
  for (int i = 1; i <= 4; ++i) spec_affine(s, "text_style_model.affine" + std::to_string(i), d);
  spec_lin(s, "att_dense", 2 * c1, d);
  for (int i = 0; i < num_layers; ++i) spec_enc(s, "att_layers." + std::to_string(i), d, d);
  spec_convblock(s, "dec3", d, c3);
  spec_convblock(s, "dec2", c3, c2);
  spec_convblock(s, "dec1", c2, c1);
  spec_lin(s, "output_dense", c1, 2);
  spec_lin(s, "pen_lifts_dense.0", c1, 1);
  return s;
}

void default_schedule(float* beta, float* abar) {
  // utils/nn.py:19-39 in fp32: 0.02 + exp(linspace(log 1e-5, log 0.4, 60)); cumprod(1 - beta)
  const float lo = (float)log(1e-5), hi = (float)log(0.4);
  const int n = DHG_NUM_STEPS;
  const float step = (hi - lo) / (float)(n - 1);
  float prod = 1.f;
  for (int i = 0; i < n; ++i) {
    const float v = (i < n / 2) ? lo + step * (float)i : hi - step * (float)(n - 1 - i);
    beta[i] = 0.02f + expf(v);
    prod = prod * (1.f - beta[i]);
    abar[i] = prod;
  }
}

int dev_alloc(std::vector<void*>& track, void** p, size_t bytes, size_t* total = nullptr) {
  if (bytes == 0) bytes = 16;
  CUDA_OK(cudaMalloc(p, bytes));
  CUDA_OK(cudaMemset(*p, 0, bytes));
  track.push_back(*p);
  if (total) *total += bytes;
  return 0;
}
int dev_upload(std::vector<void*>& track, float** p, const std::vector<float>& h) {
  if (dev_alloc(track, (void**)p, h.size() * sizeof(float))) return 1;
  CUDA_OK(cudaMemcpy(*p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// Build one packed GEMM weight from >= 1 state_dict entries concatenated along N.
int make_lin(dhg_ctx* c, const std::string& key, const std::vector<std::string>& names, bool conv) {
  Lin L;
  L.taps = conv ? 3 : 1;
  int N = 0, K = -1;
  for (auto& n : names) {
    const WSpec& ws = c->spec[c->spec_index.at(n + ".weight")];
    N += (int)ws.shape[0];
    if (K < 0) K = (int)ws.shape[1];
    if (K != (int)ws.shape[1]) return fail("make_lin %s: K mismatch", key.c_str());
  }
  L.K = K;
  L.N = N;
  L.h_w.assign((size_t)L.taps * K * N, 0.f);
  L.h_b.assign(N, 0.f);
  std::vector<float> wr(L.h_w.size());
  std::vector<bf16> w16(L.h_w.size());
  int n0 = 0;
  for (auto& n : names) {
    const std::vector<float>& W = c->raw.at(n + ".weight");
    const std::vector<float>& bsrc = c->raw.at(n + ".bias");
    const int Ni = (int)bsrc.size();
    for (int nn = 0; nn < Ni; ++nn) {
      L.h_b[n0 + nn] = bsrc[nn];
      for (int k = 0; k < K; ++k)
        for (int t = 0; t < L.taps; ++t) {
          const float v = W[((size_t)nn * K + k) * L.taps + t];  // Linear [N][K], Conv1d [N][K][3]
          const bf16 hb = __float2bfloat16_rn(v);
          L.h_w[((size_t)t * K + k) * N + n0 + nn] = v;
          wr[((size_t)t * K + k) * N + n0 + nn] = __bfloat162float(hb);
          w16[((size_t)t * N + n0 + nn) * K + k] = hb;
        }
    }
    n0 += Ni;
  }
  if (dev_upload(c->allocs, &L.w32, L.h_w)) return 1;
  if (dev_upload(c->allocs, &L.w32r, wr)) return 1;
  if (dev_upload(c->allocs, &L.bias, L.h_b)) return 1;
  if (dev_alloc(c->allocs, (void**)&L.w16, w16.size() * sizeof(bf16))) return 1;
  CUDA_OK(cudaMemcpy(L.w16, w16.data(), w16.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  if (K % 32 == 0) {   // (every GEMM of the denoiser; the folded 3-channel heads weights never run in this mode)
    std::vector<bf16> ws((size_t)L.taps * N * 2 * K);
    for (int t = 0; t < L.taps; ++t)
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
          const float v = L.h_w[((size_t)t * K + k) * N + n];
          const bf16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
          const size_t g = ((size_t)t * N + n) * 2 * K + (size_t)(k / 32) * 64 + (k % 32);
          ws[g] = hi;
          ws[g + 32] = lo;
        }
    if (dev_alloc(c->allocs, (void**)&L.w16s, ws.size() * sizeof(bf16))) return 1;
    CUDA_OK(cudaMemcpy(L.w16s, ws.data(), ws.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  }
  c->lins[key] = std::move(L);
  return 0;
}

int make_convblock_lins(dhg_ctx* c, const std::string& p) {
  return make_lin(c, p + ".conv_skip", {p + ".conv_skip"}, true) || make_lin(c, p + ".conv1", {p + ".conv1"}, true) ||
         make_lin(c, p + ".conv2", {p + ".conv2"}, true) || make_lin(c, p + ".fc", {p + ".fc"}, false);
}
int make_enc_lins(dhg_ctx* c, const std::string& p) {
  return make_lin(c, p + ".text_dense", {p + ".text_dense"}, false) ||
         make_lin(c, p + ".ffn.1", {p + ".ffn.1"}, false) || make_lin(c, p + ".ffn.3", {p + ".ffn.3"}, false) ||
         make_lin(c, p + ".mha.wq", {p + ".mha.wq"}, false) ||
         make_lin(c, p + ".mha.kv", {p + ".mha.wk", p + ".mha.wv"}, false) ||
         make_lin(c, p + ".mha.dense", {p + ".mha.dense"}, false) ||
         make_lin(c, p + ".mha2.qkv", {p + ".mha2.wq", p + ".mha2.wk", p + ".mha2.wv"}, false) ||
         make_lin(c, p + ".mha2.dense", {p + ".mha2.dense"}, false);
}

void free_plan(Plan* p) {
  if (!p) return;
  for (int i = 0; i < 4; ++i)
    if (p->graphs[i]) cudaGraphExecDestroy(p->graphs[i]);
  for (auto& m : p->host_graphs)
    for (auto g : m)
      if (g) cudaGraphExecDestroy(g);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->ev_copy_go) cudaEventDestroy(p->ev_copy_go);
  if (p->ev_copy_done) cudaEventDestroy(p->ev_copy_done);
  for (auto t : p->tc_plans) tc_gemm_plan_destroy(t);
  for (auto t : p->attn_plans) attn_tc_plan_destroy(t);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  for (auto t : p->text_stream)
    if (t) cudaStreamDestroy(t);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_last) cudaEventDestroy(p->ev_last);
  for (auto e : p->ev_text)
    if (e) cudaEventDestroy(e);
  for (auto a : p->allocs) cudaFree(a);
  for (void* q : {(void*)p->stage.x, (void*)p->stage.noise, (void*)p->stage.style, (void*)p->stage.out, (void*)p->stage.text})
    if (q) cudaFree(q);
  delete p;
}

// sinusoid table of attention.py:15-23 (halves concatenated: sin | cos), fp32 op order
std::vector<float> pe_table(int len, int dim, float pos_factor) {
  const int half = dim / 2;
  const double stepd = log(10000.0) / (double)(half - 1);
  std::vector<float> t((size_t)len * dim);
  for (int j = 0; j < half; ++j) {
    const float freq = expf((float)j * (float)(-stepd));
    for (int i = 0; i < len; ++i) {
      const float ang = ((float)i * freq) * pos_factor;
      t[(size_t)i * dim + j] = (float)sin((double)ang);
      t[(size_t)i * dim + half + j] = (float)cos((double)ang);
    }
  }
  return t;
}

// rowbias[pos, n] = bias[n] + sum_k PE[pos,k] * W[k][n] for columns [n_pe0, n_pe1); bias only elsewhere.
int make_rowbias(Plan* P, const Lin& W, const std::vector<float>& pe, int len, int n_pe0, int n_pe1, float** out, void** out16) {
  std::vector<float> t((size_t)len * W.N);
  for (int i = 0; i < len; ++i)
    for (int n = 0; n < W.N; ++n) {
      double a = W.h_b[n];
      if (n >= n_pe0 && n < n_pe1)
        for (int k = 0; k < W.K; ++k) a += (double)pe[(size_t)i * W.K + k] * (double)W.h_w[(size_t)k * W.N + n];
      t[(size_t)i * W.N + n] = (float)a;
    }
  if (dev_alloc(P->allocs, (void**)out, t.size() * sizeof(float), &P->bytes)) return 1;
  CUDA_OK(cudaMemcpy(*out, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (out16 && P->split) {   // the same table in split storage (common.cuh bfs: groups of 32 hi | 32 lo), cols % 32 == 0
    const int cols = n_pe1 - n_pe0;
    std::vector<bf16> ts((size_t)len * cols * 2);
    for (int i = 0; i < len; ++i)
      for (int n = 0; n < cols; ++n) {
        const float v = t[(size_t)i * W.N + n_pe0 + n] - W.h_b[n_pe0 + n];
        const bf16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const size_t g = (size_t)i * cols * 2 + (size_t)(n / 32) * 64 + (n % 32);
        ts[g] = hi;
        ts[g + 32] = lo;
      }
    if (dev_alloc(P->allocs, out16, ts.size() * sizeof(bf16), &P->bytes)) return 1;
    CUDA_OK(cudaMemcpy(*out16, ts.data(), ts.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  } else if (out16) {   // bf16 [len, n_pe1 - n_pe0] table of the positional term alone (the bias stays an fp32 vector)
    const int cols = n_pe1 - n_pe0;
    std::vector<bf16> t16((size_t)len * cols);
    for (int i = 0; i < len; ++i)
      for (int n = 0; n < cols; ++n)
        t16[(size_t)i * cols + n] = __float2bfloat16_rn(t[(size_t)i * W.N + n_pe0 + n] - W.h_b[n_pe0 + n]);
    if (dev_alloc(P->allocs, out16, t16.size() * sizeof(bf16), &P->bytes)) return 1;
    CUDA_OK(cudaMemcpy(*out16, t16.data(), t16.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  }
  return 0;
}

struct Builder {
  dhg_ctx* c;
  Plan* P;
  std::vector<Op>* ops;
  int64_t* nlaunch;
  bool failed = false;
  bool tail_gemm = false;   // the next gemm() is dec1.fc: skipped when the step runs with the fused tail
  const EpiSpec* tail_alt = nullptr;   // the next gemm() gets a second plan with this epilogue, used when sc.fuse_tail == 2
  bool head_gemm = false;   // the next gemm() is enc1.conv_skip: replaced by skip_from_x when the step runs with the fused head
  // skip fusion: the next gemm() is a conv_skip that is not launched when its block's bit is set in
  // sc.fuse_skip / the next gemm() is that block's fc and gets a dual-operand twin that also contracts x with conv_skip's weights
  int skipfs_bit = 0;
  struct FsSpec { Act x_raw; std::string block; int bit; };
  const FsSpec* fs_alt = nullptr;

  // Walking direction of the kernel that wrote each activation (P->dir_of: buffer -> 0 first row to last, 1 last to
  // first).  A GEMM / attention launch walks its rows in the direction OPPOSITE to the producer of its input, so it
  // starts on the rows that are still in L2 and meets the evicted ones last.  Kernels without a direction switch
  // (pool, FiLM rows, heads) count as 0.
  int dir_of(const void* p) const {
    auto it = P->dir_of.find(p);
    return it == P->dir_of.end() ? 0 : it->second;
  }
  void wrote(const void* p, int dir) { if (p) P->dir_of[p] = dir; }

  Act act(int rows, int C) {
    Act a;
    a.rows = rows;
    a.C = C;
    if (dev_alloc(P->allocs, &a.p, (size_t)rows * C * P->esize, &P->bytes)) failed = true;
    return a;
  }
  // sub-view of columns [c0, c0+n) of a (pitch stays a.C): returned as pointer offset
  const void* col(const Act& a, int c0) const { return (const char*)a.p + (size_t)c0 * P->esize; }

  RowMap map_level(int l) const { return RowMap{P->Tl[l] + 1, 1, P->B * (P->Tl[l] + 1)}; }
  RowMap map_text() const { return RowMap{P->L, 0, P->RT}; }
  RowMap map_style() const { return RowMap{P->SP, 0, P->RS}; }

  // Time the launch under every tile configuration the kernel supports for this shape (tile width, interleaved
  // accumulators, resident or streamed W, CTA pairs) on the plan's own buffers and keep the fastest.  All configurations
  // compute the same bits (kernels.h TcTune), so this only moves time.  `base` is the plan of the built-in rule.
  TcGemmPlan* autotune(TcGemmPlan* base, const bf16* Ap, int lda, int rows, const bf16* Wp, int K, int N, int taps, const Epilogue& e,
                       const Epilogue& et, const std::string& wkey, const TcDual* dual = nullptr) {
    cudaStream_t st = P->cap_stream;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { fail("autotune: event"); tc_gemm_plan_destroy(base); return nullptr; }
    // false: this configuration cannot be launched (skipped); a failure that poisons the context is caught by the
    // synchronisation at the end
    auto time_plan = [&](TcGemmPlan* p, float* ms) -> bool {
      for (int i = 0; i < 2; ++i)
        if (tc_gemm_launch(p, et, st)) { cudaGetLastError(); return false; }
      cudaEventRecord(e0, st);
      const int reps = 4;
      for (int i = 0; i < reps; ++i)
        if (tc_gemm_launch(p, et, st)) { cudaGetLastError(); return false; }
      cudaEventRecord(e1, st);
      if (cudaEventSynchronize(e1) != cudaSuccess) return false;
      cudaEventElapsedTime(ms, e0, e1);
      *ms /= reps;
      return true;
    };
    TcTune best_cfg, cfg0;
    tc_gemm_plan_config(base, &cfg0);
    best_cfg = cfg0;
    float best = 0.f, t0 = 0.f;
    bool ok = time_plan(base, &best);
    t0 = best;
    TcGemmPlan* best_plan = base;
    std::vector<TcTune> seen{cfg0};
    std::vector<int> bns;
    if (e.ln) bns.push_back(N);
    else
      for (int bn : {384, 256, 192, 128, 96, 64})
        if (N % bn == 0) bns.push_back(bn);
    for (size_t bi = 0; ok && bi < bns.size(); ++bi) {
      const int bn = bns[bi];
      for (int g : {1, 2, 4}) {
        if (g * bn > 256 && g > 1) continue;
        for (int mode = 0; mode < 5; ++mode) {   // 0: resident W, 1: streamed W, 2: streamed W + CTA pairs,
          if (mode >= 2 && g != 1) continue;       // 3 / 4: column-split LayerNorm cluster with resident / streamed W
          if (mode >= 3 && !e.ln) continue;
          TcTune t{bn, g, (mode == 0 || mode == 3) ? 1 : 0, mode == 2 ? 1 : mode >= 3 ? 2 : 0};
          char buf[256];
          if (e.split_io && (g != 1 || mode >= 3)) continue;   // split I/O: no interleaved accumulators, no column-split cluster
          if (dual && mode >= 3) continue;
          TcGemmPlan* p = tc_gemm_plan_create(Ap, lda, rows, Wp, K, N, taps, e, buf, sizeof(buf), &t, dual);
          if (!p) continue;   // configuration not available for this shape
          TcTune got;
          tc_gemm_plan_config(p, &got);
          bool dup = false;
          for (auto& sn : seen) dup = dup || (sn.bn == got.bn && sn.g == got.g && sn.resident == got.resident && sn.pair == got.pair);
          float ms = 0.f;
          if (dup || !time_plan(p, &ms)) { tc_gemm_plan_destroy(p); continue; }
          seen.push_back(got);
          if (ms < best * 0.985f) {   // must win by more than the timing noise
            tc_gemm_plan_destroy(best_plan);
            best_plan = p; best = ms; best_cfg = got;
          } else {
            tc_gemm_plan_destroy(p);
          }
        }
      }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) {
      fail("autotune %s: launch failed: %s", wkey.c_str(), cudaGetErrorString(cudaGetLastError()));
      tc_gemm_plan_destroy(best_plan);
      return nullptr;
    }
    if (getenv("DHG_DESCRIBE"))
      fprintf(stderr, "autotune %-28s rows=%d K=%d N=%d taps=%d: rule {bn=%d g=%d res=%d pair=%d} %.1f us -> {bn=%d g=%d res=%d pair=%d} %.1f us (%zu tried)\n",
              wkey.c_str(), rows, K, N, taps, cfg0.bn, cfg0.g, cfg0.resident, cfg0.pair, t0 * 1e3f, best_cfg.bn, best_cfg.g,
              best_cfg.resident, best_cfg.pair, best * 1e3f, seen.size());
    return best_plan;
  }

  void gemm(const Act& A, const std::string& wkey, const EpiSpec& s, const RowMap& map) {
    if (failed) return;
    auto it = c->lins.find(wkey);
    if (it == c->lins.end()) { fail("plan: missing packed weight %s", wkey.c_str()); failed = true; return; }
    const Lin* W = &it->second;
    if (A.C != W->K) { fail("plan: gemm %s K mismatch (%d vs %d)", wkey.c_str(), A.C, W->K); failed = true; return; }
    Epilogue e;
    memset(&e, 0, sizeof(e));
    const bool tc_path = P->tc();
    const bool sio = tc_path && P->split;
    e.split_io = sio ? 1 : 0;
    e.bias = (s.bias && (tc_path || !s.rowbias)) ? W->bias : nullptr;
    e.rowbias = tc_path ? nullptr : s.rowbias;
    e.rowbias16 = tc_path ? s.rowbias16 : nullptr;
    e.rowbias16_cols = s.rowbias16_cols;
    e.res_pre = s.res_pre.p; e.res_pre_pitch = s.res_pre.C;
    e.ln = s.ln ? 1 : 0;
    e.film_planned = s.film_off >= 0 ? 1 : 0;
    e.res_post = s.res_post.p; e.res_post_pitch = s.res_post.C;
    e.res_post_up = s.res_post_up ? 1 : 0; e.res_post_period_lo = s.res_post_period_lo;
    e.out_raw = s.out_raw.p; e.out_raw_pitch = s.out_raw.C;
    e.out_act = s.out_act.p; e.out_act_pitch = s.out_act.C;
    e.map = map;
    const int rows = A.rows, N = W->N, K = W->K, taps = W->taps;
    const int film_off = s.film_off;
    const void* Ap = A.p;
    Plan* Pl = P;
    size_t need = (size_t)rows * N;
    if (need > P->scratch_elems) P->scratch_elems = need;
    TcGemmPlan* tcp = nullptr;
    // tcgen05 operands: split storage is a bf16 row of 2 K columns (groups of 32 hi | 32 lo), weights packed alike
    const bf16* Wp = sio ? W->w16s : W->w16;
    const int Kt = sio ? 2 * K : K, ldt = sio ? 2 * A.C : A.C, tapst = taps;
    if (sio && !Wp) { fail("plan: gemm %s has no split weights (K = %d is not a multiple of 32)", wkey.c_str(), K); failed = true; return; }
    if (tc_path) {
      char buf[512];
      tcp = tc_gemm_plan_create((const bf16*)Ap, ldt, rows, Wp, Kt, N, tapst, e, buf, sizeof(buf));
      if (!tcp) { fail("plan: tcgen05 gemm %s: %s", wkey.c_str(), buf); failed = true; return; }
      if (g_opt_autotune) {
        Epilogue et = e;
        if (film_off >= 0) { et.gamma = c->cond60 + film_off; et.beta = c->cond60 + film_off + N; et.film_bstride = 0; }
        tcp = autotune(tcp, (const bf16*)Ap, ldt, rows, Wp, Kt, N, tapst, e, et, wkey);
        if (!tcp) { failed = true; return; }
      }
      const int dir = g_opt_serpentine ? !dir_of(Ap) : 0;
      tc_gemm_plan_set_reverse(tcp, dir);
      tc_gemm_plan_set_a_evict_first(tcp, g_opt_l2_hints);
      wrote(e.out_raw, dir); wrote(e.out_act, dir);
      P->tc_plans.push_back(tcp);
    }
    TcGemmPlan* tcp_alt = nullptr;
    Epilogue e_alt = e;
    const EpiSpec* alt = tail_alt;
    tail_alt = nullptr;
    const bool alt_w_per_step = alt && alt->dot_w_per_step;
    int Na = N;
    if (alt && tcp) {
      e_alt.out_raw = nullptr; e_alt.out_act = nullptr;
      e_alt.res_post = nullptr; e_alt.res_post_up = 0;
      e_alt.dot_out = alt->dot_out; e_alt.dot_act = alt->dot_act ? 1 : 0; e_alt.dot_planned = 1;
      e_alt.dot_w = alt->dot_w_per_step ? c->tail_A : alt->dot_w;
      const Lin* Wa = W;
      if (!alt->alt_wkey.empty()) {
        auto ia = c->lins.find(alt->alt_wkey);
        if (ia == c->lins.end() || ia->second.K != K || ia->second.taps != taps) { fail("plan: bad dot-mode weight %s", alt->alt_wkey.c_str()); failed = true; return; }
        Wa = &ia->second;
        e_alt.bias = Wa->bias;
      }
      Na = Wa->N;
      char buf[512];
      tcp_alt = tc_gemm_plan_create((const bf16*)Ap, A.C, rows, Wa->w16, K, Na, taps, e_alt, buf, sizeof(buf));
      if (!tcp_alt) { fail("plan: tcgen05 gemm %s (dot mode): %s", wkey.c_str(), buf); failed = true; return; }
      if (g_opt_autotune) {
        Epilogue et = e_alt;
        if (film_off >= 0) { et.gamma = c->cond60 + film_off; et.beta = c->cond60 + film_off + Na; et.film_bstride = 0; }
        tcp_alt = autotune(tcp_alt, (const bf16*)Ap, A.C, rows, Wa->w16, K, Na, taps, e_alt, et, wkey + " (dot)");
        if (!tcp_alt) { failed = true; return; }
      }
      tc_gemm_plan_set_reverse(tcp_alt, g_opt_serpentine ? !dir_of(Ap) : 0);
      tc_gemm_plan_set_a_evict_first(tcp_alt, g_opt_l2_hints);
      P->tc_plans.push_back(tcp_alt);
    }
    // skip fusion twin of a ConvBlock's fc: out = a2 . (gamma_s W_fc)^T + conv_skip(x) + (gamma_s b_fc + beta_s + b_skip)
    TcGemmPlan* tcp_fs = nullptr;
    Epilogue e_fs = e;
    const FsSpec* fs = fs_alt;
    fs_alt = nullptr;
    const int fs_bit = fs ? fs->bit : 0;
    const float* fs_bias = nullptr;
    if (fs && tcp && (P->opt_skip_fusion & fs->bit)) {
      const Lin& sk = c->lins.at(fs->block + ".conv_skip");
      fs_bias = c->fc60_bias.at(fs->block);
      e_fs.res_post = nullptr; e_fs.res_post_up = 0; e_fs.film_planned = 0; e_fs.gamma = e_fs.beta = nullptr;
      e_fs.bias = fs_bias;
      TcDual dual{(const bf16*)fs->x_raw.p, fs->x_raw.C, sk.K, sk.w16, DHG_NUM_STEPS * N};
      char buf[512];
      tcp_fs = tc_gemm_plan_create((const bf16*)Ap, A.C, rows, c->fc60_w.at(fs->block), K, N, 1, e_fs, buf, sizeof(buf), nullptr, &dual);
      if (!tcp_fs) { fail("plan: tcgen05 gemm %s (skip fusion): %s", wkey.c_str(), buf); failed = true; return; }
      if (g_opt_autotune) {
        tcp_fs = autotune(tcp_fs, (const bf16*)Ap, A.C, rows, c->fc60_w.at(fs->block), K, N, 1, e_fs, e_fs, wkey + " (+skip)", &dual);
        if (!tcp_fs) { failed = true; return; }
      }
      tc_gemm_plan_set_reverse(tcp_fs, g_opt_serpentine ? !dir_of(Ap) : 0);
      tc_gemm_plan_set_a_evict_first(tcp_fs, g_opt_l2_hints);
      P->tc_plans.push_back(tcp_fs);
    }
    const int skip_bit = skipfs_bit;
    skipfs_bit = 0;
    *nlaunch += tcp ? 1 : 2;
    const bool skippable = tail_gemm, replaceable = head_gemm;
    tail_gemm = false;
    head_gemm = false;
    dhg_ctx* cc = c;
    const std::string op_name = "gemm " + wkey + " rows=" + std::to_string(rows);
    ops->push_back([=](cudaStream_t st, const StepCtx& sc) -> int {
      if (skippable && sc.fuse_tail) return 0;
      if (skip_bit & sc.fuse_skip) return 0;   // this conv_skip is contracted inside the block's last GEMM
      if (tcp_fs && (fs_bit & sc.fuse_skip)) {
        Epilogue ef = e_fs;
        ef.bias = fs_bias + (size_t)sc.step * N;
        ef.w_row_off = sc.step * N;
        return tc_gemm_launch(tcp_fs, ef, st) ? fail("tcgen05 gemm (skip fusion) launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
      }
      if (tcp_alt && sc.fuse_tail == 2) {
        Epilogue ea = e_alt;
        if (film_off >= 0) { ea.gamma = sc.cond + film_off; ea.beta = sc.cond + film_off + Na; ea.film_bstride = sc.bstride; }
        if (alt_w_per_step) ea.dot_w = cc->tail_A + (size_t)sc.step * 3 * Na;
        return tc_gemm_launch(tcp_alt, ea, st) ? fail("tcgen05 gemm (dot mode) launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
      }
      if (replaceable && sc.fuse_head) {   // enc1.conv_skip(input_dense(x)) straight from the current x
        const float* x = sc.head.x_io ? sc.head.x_io : Pl->x_state;
        return launch_skip_from_x<bf16>(x, cc->head_M, cc->head_v, cc->head_b, (bf16*)e.out_raw, Pl->B, Pl->T, N, st)
                   ? fail("skip_from_x: unsupported channel count %d", N) : 0;
      }
      Epilogue ee = e;
      if (film_off >= 0) {
        ee.gamma = sc.cond + film_off;
        ee.beta = sc.cond + film_off + N;
        ee.film_bstride = sc.bstride;
      }
      if (tcp) return tc_gemm_launch(tcp, ee, st) ? fail("tcgen05 gemm launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
      if (Pl->prec == PREC_FP32) {   // (never split storage: that only exists on the tcgen05 path)
        launch_gemm_simt<float>((const float*)Ap, K, rows, W->w32, K, N, taps, Pl->scratch, st);
        launch_rowpost<float>(Pl->scratch, rows, N, ee, st);
      } else {
        launch_gemm_simt<bf16>((const bf16*)Ap, K, rows, W->w32r, K, N, taps, Pl->scratch, st);
        launch_rowpost<bf16>(Pl->scratch, rows, N, ee, st);
      }
      return 0;
    });
    ops->back().name = op_name;
  }

  // When to request the next item's Q / K / V tiles (attention_tc.cu): right after P V helps the shapes with few slots
  // per SM and costs a little where 4-6 items are in flight anyway, so it is timed per shape like the GEMM tiles.
  float attention_time(AttnTcPlan* ap) {   // ms per launch, < 0 on failure
    cudaStream_t st = P->cap_stream;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.f;
    bool ok = true;
    float t = 0.f;
    for (int i = 0; i < 2 && ok; ++i) ok = attn_tc_launch(ap, st) == 0;
    cudaEventRecord(e0, st);
    for (int i = 0; i < 4 && ok; ++i) ok = attn_tc_launch(ap, st) == 0;
    cudaEventRecord(e1, st);
    ok = ok && cudaEventSynchronize(e1) == cudaSuccess;
    if (ok) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ok ? t / 4.f : -1.f;
  }
  // Variant of the attention launch for this shape, timed on the plan's buffers like the GEMM tiles: 0 / 1 = all keys
  // at once, next item's tiles requested after O is stored / right after P V; 2 = key-block kernel (where it applies)
  int attention_pick(AttnTcPlan* ap, AttnTcPlan* ap_long) {
    float t[3] = {-1.f, -1.f, -1.f};
    if (attn_tc_plan_is_long(ap)) return 2;
    for (int v = 0; v < 2; ++v) {
      attn_tc_plan_set_early_load(ap, v);
      t[v] = attention_time(ap);
      if (t[v] < 0.f) return -1;
    }
    if (ap_long) { t[2] = attention_time(ap_long); if (t[2] < 0.f) return -1; }
    if (getenv("DHG_DESCRIBE")) fprintf(stderr, "autotune attention: late load %.1f us, early load %.1f us, key blocks %.1f us\n", t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f);
    int best = t[1] < t[0] * 0.985f ? 1 : 0;
    if (ap_long && t[2] < t[best] * 0.985f) best = 2;
    return best;
  }

  // k / v may exist in several copies (text sets); the launch picks sc.text_set.
  void attention_sets(const void* q, int qp, const void* const* ks, const void* const* vs, int nsets, int kp, int vp,
                      const Act& o, int H, int D, int Tq, int q_period, int q_pad, int Tk, int k_period, int k_pad,
                      bool masked, int q_rows, int k_rows) {
    if (failed) return;
    std::vector<AttnParams> a(nsets);
    std::vector<AttnTcPlan*> plans(nsets, nullptr);
    int variant = 1;
    Plan* Pl = P;
    *nlaunch += 1;
    bool tc = false;
    for (int s = 0; s < nsets; ++s) {
      a[s].q = q; a[s].k = ks[s]; a[s].v = vs[s]; a[s].o = o.p;
      a[s].q_pitch = qp; a[s].k_pitch = kp; a[s].v_pitch = vp; a[s].o_pitch = o.C;
      a[s].q_period = q_period; a[s].q_pad = q_pad; a[s].k_period = k_period; a[s].k_pad = k_pad;
      a[s].B = P->B; a[s].H = H; a[s].D = D; a[s].Tq = Tq; a[s].Tk = Tk;
      a[s].scale = 1.0f / sqrtf((float)D);
      a[s].text = masked ? P->text : nullptr;
      a[s].split = P->split ? 1 : 0;
      if ((P->prec == PREC_BF16 || P->split) && P->attn_impl == 1 && attn_tc_supported(a[s])) {
        char buf[512];
        plans[s] = attn_tc_plan_create(a[s], q_rows, k_rows, buf, sizeof(buf));
        if (!plans[s]) { fail("plan: tcgen05 attention: %s", buf); failed = true; return; }
        const int dir = g_opt_serpentine ? !dir_of(q) : 0;
        attn_tc_plan_set_reverse(plans[s], dir);
        wrote(o.p, dir);
        if (g_opt_autotune) {
          if (s == 0) {
            // The key-block kernel sums its probabilities in another order than the all-keys kernel: letting a timing
            // decide between them made the BITS of a result depend on the batch a plan was built for (seen as different
            // sha1s of the same 64 prompts sharded over 1 and 2 GPUs).  Only the numerically identical variants (tile-load
            // order) are timed by default; "attn_keyblock_auto" = 1 adds the key-block kernel for 128 < Tk <= 256.
            AttnTcPlan* alt = g_opt_attn_keyblock_auto ? attn_tc_plan_create(a[s], q_rows, k_rows, buf, sizeof(buf), 1) : nullptr;
            if (alt) attn_tc_plan_set_reverse(alt, dir);
            variant = attention_pick(plans[0], alt);
            if (alt) attn_tc_plan_destroy(alt);
          }
          if (variant < 0) { fail("plan: attention autotune launch failed: %s", cudaGetErrorString(cudaGetLastError())); failed = true; return; }
          if (variant == 2 && !attn_tc_plan_is_long(plans[s])) {
            attn_tc_plan_destroy(plans[s]);
            plans[s] = attn_tc_plan_create(a[s], q_rows, k_rows, buf, sizeof(buf), 1);
            if (!plans[s]) { fail("plan: tcgen05 attention: %s", buf); failed = true; return; }
            attn_tc_plan_set_reverse(plans[s], dir);
          } else if (variant < 2) {
            attn_tc_plan_set_early_load(plans[s], variant);
          }
        }
        P->attn_plans.push_back(plans[s]);
        tc = true;
      }
    }
    ops->push_back([=](cudaStream_t st, const StepCtx& sc) -> int {
      const int set = (sc.text_set >= 0 && sc.text_set < nsets) ? sc.text_set : 0;
      if (tc) return attn_tc_launch(plans[set], st) ? fail("tcgen05 attention launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
      const int r = DHG_STORAGE(Pl, T, return launch_attention_simt<T>(a[set], st););
      return r ? fail("attention: unsupported head depth %d", a[set].D) : 0;
    });
    ops->back().name = std::string(tc ? "attention (tcgen05)" : "attention (CUDA cores)") + " Tq=" + std::to_string(Tq) + " Tk=" + std::to_string(Tk) + " H=" + std::to_string(H);
  }
  void attention(const void* q, int qp, const void* k, int kp, const void* v, int vp, const Act& o, int H, int D,
                 int Tq, int q_period, int q_pad, int Tk, int k_period, int k_pad, bool masked, int q_rows, int k_rows) {
    attention_sets(q, qp, &k, &v, 1, kp, vp, o, H, D, Tq, q_period, q_pad, Tk, k_period, k_pad, masked, q_rows, k_rows);
  }

  void film_rows(const Act& in, const Act& out, int period, int film_off) {
    if (failed) return;
    Plan* Pl = P;
    *nlaunch += 1;
    ops->push_back([=](cudaStream_t st, const StepCtx& sc) -> int {
      const float* g = sc.cond + film_off;
      const float* b = sc.cond + film_off + in.C;
      DHG_STORAGE(Pl, T, launch_film_rows<T>((const T*)in.p, (T*)out.p, in.rows, in.C, period, g, b, sc.bstride, st););
      return 0;
    });
  }

  void pool(const Act& in, const Act& out_raw, const Act& out_act, int Tlo) {
    if (failed) return;
    Plan* Pl = P;
    *nlaunch += 1;
    ops->push_back([=](cudaStream_t st, const StepCtx&) -> int {
      DHG_STORAGE(Pl, T, launch_pool<T>((const T*)in.p, (T*)out_raw.p, (T*)out_act.p, Pl->B, Tlo, in.C, st););
      return 0;
    });
  }

  int film(const std::string& name) {
    auto it = c->film_off.find(name);
    if (it == c->film_off.end()) { fail("plan: unknown FiLM layer %s", name.c_str()); failed = true; return 0; }
    return it->second;
  }

  // cnn.py:52-87
  Act conv_block(const std::string& p, const Act& in_raw, const Act& in_act, int Cout, int level, bool want_act, Act* out_act) {
    const int R = P->R[level];
    const RowMap m = map_level(level);
    Act skip = act(R, Cout), a1 = act(R, Cout / 2), a2 = act(R, Cout), out = act(R, Cout);
    const bool tail = p == "dec1" && P->prec == PREC_BF16 && P->gemm_impl == 1;
    EpiSpec d0, d2;   // dot-mode twins of conv_skip and conv2 of the last block (tail fusion level 2)
    if (tail) {
      if (dev_alloc(P->allocs, (void**)&P->tail_dot_skip, (size_t)R * 4 * sizeof(float), &P->bytes) ||
          dev_alloc(P->allocs, (void**)&P->tail_dot_a2, (size_t)R * 4 * sizeof(float), &P->bytes)) { failed = true; return Act(); }
      // conv_skip is only read through H (3 x C): fold it, W'[tau] = H . W_skip[tau] (3 of 32 output columns, K unchanged);
      // the 'dot vectors' just pick those 3 columns
      d0.dot_out = P->tail_dot_skip;
      if (P->opt_tail_fusion >= 3) { d0.dot_w = c->tail_pick; d0.alt_wkey = "dec1.conv_skip.heads"; }
      else d0.dot_w = c->tail_H;
      d2.film_off = film(p + ".affine2"); d2.dot_out = P->tail_dot_a2; d2.dot_w_per_step = true; d2.dot_act = true;
    }
    EpiSpec s0; s0.out_raw = skip;
    if (p == "enc1") head_gemm = true;
    if (tail) tail_alt = &d0;
    static const char* kFsBlocks[5] = {"enc1", "enc2", "enc4", "dec3", "dec2"};
    int fs_bit = 0;
    for (int i = 0; i < 5; ++i)
      if (p == kFsBlocks[i]) fs_bit = 1 << i;
    FsSpec fsp{in_raw, p, fs_bit};
    skipfs_bit = fs_bit;
    gemm(in_raw, p + ".conv_skip", s0, m);
    EpiSpec s1; s1.film_off = film(p + ".affine1"); s1.out_act = a1;
    gemm(in_act, p + ".conv1", s1, m);
    EpiSpec s2; s2.film_off = film(p + ".affine2"); s2.out_act = a2;
    if (tail) tail_alt = &d2;
    gemm(a1, p + ".conv2", s2, m);
    EpiSpec s3; s3.film_off = film(p + ".affine3"); s3.res_post = skip; s3.out_raw = out;
    if (want_act) { *out_act = act(R, Cout); s3.out_act = *out_act; }
    if (p == "dec1") { P->tail_a2 = a2; P->tail_skip = skip; tail_gemm = true; }
    if (fs_bit) fs_alt = &fsp;
    gemm(a2, p + ".fc", s3, m);
    return out;
  }

  // model.py:42-45, text side of an EncoderLayer: text' = FiLM0(LN(text_dense(SiLU(text)))), then the k | v projections
  // of the cross-attention (k gets the text positional embedding folded into a per-position bias row, v does not).
  Act encoder_text(const std::string& p, const Act& text_act) {
    if (failed) return Act();
    const int dm = c->lins.at(p + ".text_dense").N, L = P->L;
    const RowMap mt = map_text();
    const std::vector<float> pe_t = pe_table(L, dm, 1.0f);
    float* rb_kv = nullptr;
    void* rb16_kv = nullptr;
    if (make_rowbias(P, c->lins.at(p + ".mha.kv"), pe_t, L, 0, dm, &rb_kv, &rb16_kv)) { failed = true; return Act(); }
    Act tp = act(P->RT, dm), kv = act(P->RT, 2 * dm);
    EpiSpec s; s.ln = true; s.film_off = film(p + ".affine0"); s.out_raw = tp;
    gemm(text_act, p + ".text_dense", s, mt);
    EpiSpec skv; skv.rowbias = rb_kv; skv.rowbias16 = rb16_kv; skv.rowbias16_cols = dm; skv.out_raw = kv;
    gemm(tp, p + ".mha.kv", skv, mt);
    return kv;
  }

  // model.py:35-58, stroke side (kvs: the k | v rows written by encoder_text, one per text set)
  Act encoder_layer(const std::string& p, const Act& x, const Act* kvs, int nsets, int heads, float pos_factor, int level) {
    const int R = P->R[level], dm = x.C, Tl = P->Tl[level], L = P->L;
    const RowMap m = map_level(level);
    const int D = dm / heads;
    // positional-embedding-folded bias tables
    const std::vector<float> pe_x = pe_table(Tl, dm, pos_factor);
    float *rb_q = nullptr, *rb_qkv = nullptr;
    void *rb16_q = nullptr, *rb16_qkv = nullptr;
    if (!failed) {
      if (make_rowbias(P, c->lins.at(p + ".mha.wq"), pe_x, Tl, 0, dm, &rb_q, &rb16_q) ||
          make_rowbias(P, c->lins.at(p + ".mha2.qkv"), pe_x, Tl, 0, 2 * dm, &rb_qkv, &rb16_qkv))   // q,k get PE, v does not
        failed = true;
    }
    Act q = act(R, dm), o = act(R, dm), x2 = act(R, dm);
    Act qkv = act(R, 3 * dm), o2 = act(R, dm), x3r = act(R, dm), x3a = act(R, dm), hid = act(R, 2 * dm), out = act(R, dm);
    EpiSpec sq; sq.rowbias = rb_q; sq.rowbias16 = rb16_q; sq.rowbias16_cols = dm; sq.out_raw = q;
    gemm(x, p + ".mha.wq", sq, m);
    const void *ks[kMaxTextSets], *vs[kMaxTextSets];
    for (int s = 0; s < nsets; ++s) { ks[s] = kvs[s].p; vs[s] = col(kvs[s], dm); }
    attention_sets(q.p, dm, ks, vs, nsets, 2 * dm, 2 * dm, o, heads, D, Tl, Tl + 1, 1, L, L, 0, true, R, P->RT);
    EpiSpec sd; sd.ln = true; sd.film_off = film(p + ".affine1"); sd.res_post = x; sd.out_raw = x2;
    gemm(o, p + ".mha.dense", sd, m);
    EpiSpec sqkv; sqkv.rowbias = rb_qkv; sqkv.rowbias16 = rb16_qkv; sqkv.rowbias16_cols = 2 * dm; sqkv.out_raw = qkv;
    gemm(x2, p + ".mha2.qkv", sqkv, m);
    attention(qkv.p, 3 * dm, col(qkv, dm), 3 * dm, col(qkv, 2 * dm), 3 * dm, o2, heads, D, Tl, Tl + 1, 1, Tl, Tl + 1, 1, false, R, R);
    EpiSpec sd2; sd2.res_pre = x2; sd2.ln = true; sd2.film_off = film(p + ".affine2"); sd2.out_raw = x3r; sd2.out_act = x3a;
    gemm(o2, p + ".mha2.dense", sd2, m);
    EpiSpec sf1; sf1.out_act = hid;
    gemm(x3a, p + ".ffn.1", sf1, m);
    EpiSpec sf2; sf2.res_pre = x3r; sf2.ln = true; sf2.film_off = film(p + ".affine3"); sf2.out_raw = out;
    gemm(hid, p + ".ffn.3", sf2, m);
    return out;
  }
};

int build_plan(dhg_ctx* c, Plan* P) {
  const int B = P->B, T = P->T, L = P->L;
  const int c1 = c->c1, c2 = c->c2, c3 = c->c3, d = c->d;
  P->esize = P->prec == PREC_FP32 ? 4 : 2;
  for (int l = 0; l < 4; ++l) {
    P->Tl[l] = T >> l;
    P->R[l] = B * (P->Tl[l] + 1) + 1;
  }
  P->SP = P->S * kStyleSplit;
  P->RT = B * L;
  P->RS = B * P->SP;
  const int tot = c->film_total;

  if (dev_alloc(P->allocs, (void**)&P->text, (size_t)B * L * sizeof(int64_t), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->style, (size_t)B * P->S * kStyleWidth * sizeof(float), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->x_state, (size_t)B * T * 2 * sizeof(float), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->out, (size_t)B * T * 3 * sizeof(float), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->sigma_in, (size_t)B * sizeof(float), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->sig_emb, (size_t)B * kSigmaDim * sizeof(float), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->cond_b, (size_t)B * tot * sizeof(float), &P->bytes)) return 1;
  if (dev_alloc(P->allocs, (void**)&P->err_flag, sizeof(int), &P->bytes)) return 1;
  CUDA_OK(cudaStreamCreateWithFlags(&P->cap_stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&P->ev_last, cudaEventDisableTiming));
  P->opt_tail_fusion = g_opt_tail_fusion;
  P->opt_head_fusion = g_opt_head_fusion;
  P->opt_skip_fusion = (P->prec == PREC_BF16 && P->gemm_impl == 1) ? g_opt_skip_fusion : 0;

  // ---- hoisted, step-independent part of TextStyleEncoder (text_style.py:92-99) ----
  Builder once{c, P, &P->once_ops, &P->launches_once};
  const std::string ts = "text_style_model";
  Act style_act = once.act(P->RS, kStyleWidth / kStyleSplit);
  Act sh = once.act(P->RS, 4 * c2), s0 = once.act(P->RS, d), t0 = once.act(P->RT, d);
  {
    Plan* Pl = P;
    P->launches_once += 2;
    P->once_ops.push_back([=](cudaStream_t st, const StepCtx&) -> int {
      const size_t n = (size_t)Pl->B * Pl->S * kStyleWidth;
      DHG_STORAGE(Pl, T,
                  launch_silu_convert<T>(Pl->style, (T*)style_act.p, n, st);
                  launch_embed_ln<T>(Pl->text, c->emb, kVocab, t0.C, (T*)t0.p, Pl->RT, Pl->err_flag, st););
      return 0;
    });
  }
  EpiSpec e1; e1.out_act = sh;
  once.gemm(style_act, ts + ".style_ffn.1", e1, once.map_style());
  EpiSpec e2; e2.ln = true; e2.out_raw = s0;
  once.gemm(sh, ts + ".style_ffn.3", e2, once.map_style());
  if (once.failed) return 1;

  // ---- per-step, text side (depends on the step through FiLM only, never on x): one op list per text set ----
  std::vector<std::string> enc_names = {"enc3", "enc5"};
  for (int i = 0; i < c->cfg.num_layers; ++i) enc_names.push_back("att_layers." + std::to_string(i));
  // several sets (and streams) only where every text-side launch is a self-contained tcgen05 kernel: the CUDA-core
  // GEMM path shares one fp32 scratch buffer between launches and must stay serial
  P->text_sets = P->tc() ? std::min(std::max(g_opt_text_sets, 1), kMaxTextSets) : 1;
  if (P->text_sets > 1) {
    CUDA_OK(cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming));
    for (int i = 1; i < P->text_sets; ++i) {
      CUDA_OK(cudaStreamCreateWithFlags(&P->text_stream[i], cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&P->ev_text[i], cudaEventDisableTiming));
    }
  }
  std::vector<Act> kv_sets[kMaxTextSets];
  Act text_act0, tf0, sf0, t1a0, to0;
  for (int set = 0; set < P->text_sets; ++set) {
    int64_t other = 0;
    Builder tb{c, P, &P->text_ops[set], set == 0 ? &P->launches_text : &other};
    Act sf = tb.act(P->RS, d), tf = tb.act(P->RT, d), skv = tb.act(P->RS, 2 * d), tq = tb.act(P->RT, d);
    Act to = tb.act(P->RT, d), t1a = tb.act(P->RT, d), th = tb.act(P->RT, 2 * d), text_act = tb.act(P->RT, d);
    tb.film_rows(s0, sf, P->SP, tb.film(ts + ".affine1"));
    tb.film_rows(t0, tf, L, tb.film(ts + ".affine2"));
    { EpiSpec s; s.out_raw = skv; tb.gemm(sf, ts + ".mha.kv", s, tb.map_style()); }
    { EpiSpec s; s.out_raw = tq; tb.gemm(tf, ts + ".mha.wq", s, tb.map_text()); }
    tb.attention(tq.p, d, skv.p, 2 * d, tb.col(skv, d), 2 * d, to, 8, d / 8, L, L, 0, P->SP, P->SP, 0, false, P->RT, P->RS);
    { EpiSpec s; s.res_pre = tf; s.ln = true; s.film_off = tb.film(ts + ".affine3"); s.out_act = t1a;
      tb.gemm(to, ts + ".mha.dense", s, tb.map_text()); }
    { EpiSpec s; s.out_act = th; tb.gemm(t1a, ts + ".text_ffn.1", s, tb.map_text()); }
    { EpiSpec s; s.ln = true; s.film_off = tb.film(ts + ".affine4"); s.out_act = text_act;
      tb.gemm(th, ts + ".text_ffn.3", s, tb.map_text()); }
    for (auto& name : enc_names) kv_sets[set].push_back(tb.encoder_text(name, text_act));
    if (tb.failed) return 1;
    if (set == 0) { text_act0 = text_act; tf0 = tf; sf0 = sf; t1a0 = t1a; to0 = to; }
  }
  auto kvs_of = [&](int layer, Act* out2) { for (int s = 0; s < P->text_sets; ++s) out2[s] = kv_sets[s][layer]; };

  // ---- per-step, stroke side ----
  Builder bd{c, P, &P->step_ops, &P->launches_step};
  // input_dense (model.py:139)
  Act in_raw = bd.act(P->R[0], c1), in_act = bd.act(P->R[0], c1);
  {
    Plan* Pl = P;
    P->launches_step += 1;
    P->in_raw = in_raw; P->in_act = in_act;
    P->step_ops.push_back([=](cudaStream_t st, const StepCtx& sc) -> int {
      if (sc.skip_input_dense) return 0;
      const float* x = sc.head.x_io ? sc.head.x_io : Pl->x_state;
      DHG_STORAGE(Pl, T, launch_input_dense<T>(x, c->in_W, c->in_b, (T*)in_raw.p, (T*)in_act.p, Pl->B, Pl->T, in_raw.C, st););
      return 0;
    });
  }
  Act dummy;
  Act h1 = bd.conv_block("enc1", in_raw, in_act, c1, 0, false, &dummy);
  auto tap = [&](const char* name, const Act& a, int level) {
    if (level >= 0) P->taps[name] = Plan::Tap{a, P->Tl[level] + 1, 1};
    else P->taps[name] = Plan::Tap{a, level == -1 ? P->L : P->SP, 0};
  };
  tap("text_act", text_act0, -1); tap("tf", tf0, -1); tap("sf", sf0, -2); tap("t1a", t1a0, -1); tap("to", to0, -1);
  tap("in_raw", in_raw, 0); tap("h1", h1, 0);
  Act p1r = bd.act(P->R[1], c1), p1a = bd.act(P->R[1], c1);
  bd.pool(h1, p1r, p1a, P->Tl[1]);
  Act h2c = bd.conv_block("enc2", p1r, p1a, c2, 1, false, &dummy);
  Act kvl[kMaxTextSets];
  kvs_of(0, kvl);
  Act h2 = bd.encoder_layer("enc3", h2c, kvl, P->text_sets, 3, 4.0f, 1);
  tap("h2c", h2c, 1); tap("h2", h2, 1);
  Act p2r = bd.act(P->R[2], c2), p2a = bd.act(P->R[2], c2);
  bd.pool(h2, p2r, p2a, P->Tl[2]);
  Act h3c = bd.conv_block("enc4", p2r, p2a, c3, 2, false, &dummy);
  kvs_of(1, kvl);
  Act h3 = bd.encoder_layer("enc5", h3c, kvl, P->text_sets, 4, 2.0f, 2);
  tap("h3c", h3c, 2); tap("h3", h3, 2);
  Act p3r = bd.act(P->R[3], c3);
  bd.pool(h3, p3r, Act(), P->Tl[3]);
  Act xa = bd.act(P->R[3], d);
  { EpiSpec s; s.out_raw = xa; bd.gemm(p3r, "att_dense", s, bd.map_level(3)); }
  tap("att_in", xa, 3);
  for (int i = 0; i < c->cfg.num_layers; ++i) {
    kvs_of(2 + i, kvl);
    xa = bd.encoder_layer("att_layers." + std::to_string(i), xa, kvl, P->text_sets, 6, 1.0f, 3);
    tap(("att" + std::to_string(i)).c_str(), xa, 3);
  }
  // decoder: upsample(x) + skip_conv(h) (model.py:169-176), then ConvBlock
  Act u3r = bd.act(P->R[2], d), u3a = bd.act(P->R[2], d);
  { EpiSpec s; s.res_post = xa; s.res_post_up = true; s.res_post_period_lo = P->Tl[3] + 1; s.out_raw = u3r; s.out_act = u3a;
    bd.gemm(h3, "skip_conv3", s, bd.map_level(2)); }
  Act d3 = bd.conv_block("dec3", u3r, u3a, c3, 2, false, &dummy);
  tap("d3", d3, 2);
  Act u2r = bd.act(P->R[1], c3), u2a = bd.act(P->R[1], c3);
  { EpiSpec s; s.res_post = d3; s.res_post_up = true; s.res_post_period_lo = P->Tl[2] + 1; s.out_raw = u2r; s.out_act = u2a;
    bd.gemm(h2, "skip_conv2", s, bd.map_level(1)); }
  Act d2 = bd.conv_block("dec2", u2r, u2a, c2, 1, false, &dummy);
  tap("d2", d2, 1);
  Act u1r = bd.act(P->R[0], c2), u1a = bd.act(P->R[0], c2);
  { EpiSpec s; s.res_post = d2; s.res_post_up = true; s.res_post_period_lo = P->Tl[1] + 1; s.out_raw = u1r; s.out_act = u1a;
    bd.gemm(h1, "skip_conv1", s, bd.map_level(0)); }
  Act d1 = bd.conv_block("dec1", u1r, u1a, c1, 0, false, &dummy);
  if (bd.failed) return 1;
  tap("d1", d1, 0);
  P->head_in = d1;
  {
    Plan* Pl = P;
    P->launches_step += 1;
    P->step_ops.push_back([=](cudaStream_t st, const StepCtx& sc) -> int {
      HeadParams hp = sc.head;
      if (hp.x_io == nullptr && hp.eps_out == nullptr) return fail("head: nothing to do");
      if (sc.fuse_next_input) {
        hp.next_raw = sc.fuse_head ? nullptr : Pl->in_raw.p;   // nobody reads in_raw when enc1.conv_skip works from x
        hp.next_act = Pl->in_act.p; hp.in_W = c->in_W; hp.in_b = c->in_b;
      }
      if (sc.fuse_tail == 2) {   // the two dot-mode GEMMs of this step left a2 . tail_A[step] and skip . tail_H per point
        return launch_heads_from_dots<bf16>(Pl->tail_dot_a2, Pl->tail_dot_skip, c->tail_c + (size_t)sc.step * 3, d1.C, hp, st)
                   ? fail("head kernel: unsupported channel count %d", d1.C) : 0;
      }
      if (sc.fuse_tail) {   // eps | pen = a2 . tail_A[step] + skip . tail_H + tail_c[step]  (dhg_finalize)
        const int C = d1.C;
        const float* A = c->tail_A + (size_t)sc.step * 3 * C;
        const float* cc = c->tail_c + (size_t)sc.step * 3;
        hp.h2 = Pl->tail_skip.p;
        hp.w2 = c->tail_H;
        return launch_heads_update<bf16>((const bf16*)Pl->tail_a2.p, C, A, cc, A + 2 * C, cc + 2, hp, st)
                   ? fail("head kernel: unsupported channel count %d", C) : 0;
      }
      const int rc = DHG_STORAGE(Pl, T, return launch_heads_update<T>((const T*)d1.p, d1.C, c->out_W, c->out_b, c->pen_W, c->pen_b, hp, st););
      return rc ? fail("head kernel: unsupported channel count %d", d1.C) : 0;
    });
  }
  if (P->scratch_elems && !P->tc())
    if (dev_alloc(P->allocs, (void**)&P->scratch, P->scratch_elems * sizeof(float), &P->bytes)) return 1;
  return 0;
}

int run_ops(const std::vector<Op>& ops, cudaStream_t st, const StepCtx& sc) {
  static const bool sync_ops = getenv("DHG_SYNC_OPS") != nullptr;   // debugging aid: synchronise after every op and name the one that faulted
  for (size_t i = 0; i < ops.size(); ++i) {
    if (ops[i](st, sc)) return 1;
    if (sync_ops) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(st, &cs);
      if (cs == cudaStreamCaptureStatusNone) {
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail("op %zu (%s) faulted: %s", i, ops[i].name.c_str(), cudaGetErrorString(e));
      }
    }
  }
  return 0;
}

void head_for_step(const dhg_ctx* c, Plan* P, int i, int mode, bool has_noise, bool seeded, HeadParams* hp) {
  memset(hp, 0, sizeof(*hp));
  hp->B = P->B;
  hp->T = P->T;
  hp->x_io = P->x_state;
  hp->x_out_stride = 2;
  hp->mode = mode;
  const float beta = c->beta[i], abar = c->abar[i];
  if (mode == DHG_MODE_NEW) {
    const float anext = i > 1 ? c->abar[i - 1] : 1.0f;   // inference.py:87
    hp->c_eps = sqrtf(1.f - abar);
    hp->c_div = sqrtf(1.f - beta);
    hp->c_noise = sqrtf(1.f - anext);
  } else {
    hp->c_eps = beta;
    hp->c_eps2 = sqrtf(1.f - abar);
    hp->c_div = 1.f / sqrtf(1.f - beta);
    hp->c_noise = i > 0 ? sqrtf(beta) : 0.f;             // add_sigma=bool(i), inference.py:92
  }
  hp->noise = has_noise ? P->noise + (size_t)i * P->B * P->T * 2 : nullptr;
  (void)seeded;
  if (i == 0) {  // last iteration: emit [B,T,3] = cat(x, pen)  (inference.py:96)
    hp->x_out = P->out;
    hp->x_out_stride = 3;
    hp->pen_out = P->out;
    hp->pen_stride = 3;
    hp->pen_offset = 2;
  }
}

// Steps i_first, i_first - 1, .., i_last of the chain (default: all 60).  A partial range must start on a text-set group
// boundary ((59 - i_first) % text_sets == 0); the once-ops belong to the part that starts at step 59.
int run_chain(dhg_ctx* c, Plan* P, int mode, bool has_noise, cudaStream_t st, int i_first = DHG_NUM_STEPS - 1, int i_last = 0) {
  StepCtx sc;
  memset(&sc, 0, sizeof(sc));
  sc.cond = c->cond60;
  sc.bstride = 0;
  if ((DHG_NUM_STEPS - 1 - i_first) % P->text_sets != 0) return fail("run_chain: step %d is not a text-set boundary", i_first);
  if (i_first == DHG_NUM_STEPS - 1 && run_ops(P->once_ops, st, sc)) return 1;
  // The text side of a step does not depend on x.  Before step i with (59 - i) % n == 0 the text sides of steps
  // i, i-1, .., i-n+1 are launched together, set (j % n) for step j, set 0's on st and the others on their own streams
  // (forked from and joined back into st); the previous reader of a set, step j+n, has finished by then.
  const int n = P->text_sets;
  auto text_ctx = [&](int i) {
    StepCtx t;
    memset(&t, 0, sizeof(t));
    t.cond = c->cond60 + (size_t)i * c->film_total;
    t.bstride = 0;
    return t;
  };
  for (int i = i_first; i >= i_last; --i) {
    if ((DHG_NUM_STEPS - 1 - i) % n == 0) {
      if (n > 1) CUDA_OK(cudaEventRecord(P->ev_fork, st));
      for (int j = i; j > i - n && j >= 0; --j) {
        const int set = j % n;
        cudaStream_t ts = set == 0 ? st : P->text_stream[set];
        if (set != 0) CUDA_OK(cudaStreamWaitEvent(ts, P->ev_fork, 0));
        if (run_ops(P->text_ops[set], ts, text_ctx(j))) return 1;
        if (set != 0) CUDA_OK(cudaEventRecord(P->ev_text[set], ts));
      }
      for (int j = i; j > i - n && j >= 0; --j)
        if (j % n != 0) CUDA_OK(cudaStreamWaitEvent(st, P->ev_text[j % n], 0));
    }
    sc.cond = c->cond60 + (size_t)i * c->film_total;
    sc.skip_input_dense = i != DHG_NUM_STEPS - 1;   // written by the previous step's head kernel
    sc.fuse_next_input = i != 0;
    sc.text_set = i % n;
    sc.fuse_tail = (P->prec == PREC_BF16 && P->gemm_impl == 1) ? (P->opt_tail_fusion > 2 ? 2 : P->opt_tail_fusion) : 0;
    sc.fuse_skip = P->opt_skip_fusion;
    // enc1.conv_skip inside enc1.fc reads the raw input rows, so the head kernel keeps writing them
    sc.fuse_head = P->opt_head_fusion && P->prec == PREC_BF16 && P->gemm_impl == 1 && !(sc.fuse_skip & 1);
    sc.step = i;
    head_for_step(c, P, i, mode, has_noise, false, &sc.head);
    if (run_ops(P->step_ops, st, sc)) return 1;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_chain(dhg_ctx* c, Plan* P, int mode, bool has_noise, cudaStream_t st) {
  if (!c->opt_graph) return run_chain(c, P, mode, has_noise, st);
  const int gi = mode * 2 + (has_noise ? 1 : 0);
  if (!P->graphs[gi]) {
    cudaGraph_t g = nullptr;
    CUDA_OK(cudaStreamBeginCapture(P->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_chain(c, P, mode, has_noise, P->cap_stream);
    cudaError_t ce = cudaStreamEndCapture(P->cap_stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (ce != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&P->graphs[gi], g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail("graph instantiate failed: %s", cudaGetErrorString(ce));
  }
  CUDA_OK(cudaGraphLaunch(P->graphs[gi], st));
  return 0;
}

// First part of the chain for dhg_sample_host: a multiple of the text-set group size, about ten steps (33 ms at B = 1024:
// ample for the 160 MB of noise the remaining steps need to arrive over PCIe)
int host_head_steps(const Plan* P) {
  const int n = P->text_sets;
  return ((10 + n - 1) / n) * n;
}
int launch_chain_part(dhg_ctx* c, Plan* P, int mode, int part, cudaStream_t st) {
  const int k = host_head_steps(P);
  const int i_first = part == 0 ? DHG_NUM_STEPS - 1 : DHG_NUM_STEPS - 1 - k, i_last = part == 0 ? DHG_NUM_STEPS - k : 0;
  if (!c->opt_graph) return run_chain(c, P, mode, true, st, i_first, i_last);
  cudaGraphExec_t& ge = P->host_graphs[mode][part];
  if (!ge) {
    cudaGraph_t g = nullptr;
    CUDA_OK(cudaStreamBeginCapture(P->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_chain(c, P, mode, true, P->cap_stream, i_first, i_last);
    cudaError_t ce = cudaStreamEndCapture(P->cap_stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (ce != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail("graph instantiate failed: %s", cudaGetErrorString(ce));
  }
  CUDA_OK(cudaGraphLaunch(ge, st));
  return 0;
}

int check_ready(const dhg_ctx* c, bool need_plan) {
  if (!c) return fail("null ctx");
  if (!c->finalized) return fail("dhg_finalize has not been called");
  if (need_plan && !c->plan) return fail("dhg_plan has not been called");
  return 0;
}

// Stream rule of the plan (include/dhg_b200.h): a call first waits (on its own stream) for the previous call on this
// plan, whatever stream that one used, and leaves its own completion in ev_last.
int plan_enter(Plan* P, cudaStream_t st) {
  CUDA_OK(cudaStreamWaitEvent(st, P->ev_last, 0));
  return 0;
}
int plan_leave(Plan* P, cudaStream_t st) {
  CUDA_OK(cudaEventRecord(P->ev_last, st));
  return 0;
}
// After a synchronisation: did any kernel of the plan flag bad input (token id outside the embedding table)?
int plan_check_flags(Plan* P) {
  int flag = 0;
  CUDA_OK(cudaMemcpy(&flag, P->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    CUDA_OK(cudaMemset(P->err_flag, 0, sizeof(int)));
    return fail("text token id out of range [0, %d) (the reference's nn.Embedding raises IndexError here)", kVocab);
  }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char* dhg_last_error(void) { return g_err; }
int32_t dhg_abi_version(void) { return 1; }

int32_t dhg_create(int32_t device, const dhg_config* cfg, dhg_ctx** out) {
  if (!cfg || !out) return fail("dhg_create: null argument");
  if (cfg->channels != 128)
    return fail("dhg_create: channels must be 128 (the reference hard-codes a 32-wide sigma embedding, conditioning.py:9); got %d", cfg->channels);
  if (cfg->num_layers < 0 || cfg->num_layers > 64) return fail("dhg_create: bad num_layers %d", cfg->num_layers);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("dhg_create: no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail("dhg_create: device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail("dhg_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
  CUDA_OK(cudaSetDevice(device));
  dhg_ctx* c = new dhg_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->cfg = *cfg;
  c->c1 = cfg->channels;
  c->c2 = cfg->channels * 3 / 2;
  c->c3 = cfg->channels * 2;
  c->d = 2 * c->c2;
  c->spec = build_spec(cfg->num_layers, cfg->channels);
  for (size_t i = 0; i < c->spec.size(); ++i) c->spec_index[c->spec[i].name] = (int)i;
  default_schedule(c->beta, c->abar);
  *out = c;
  return 0;
}

int32_t dhg_destroy(dhg_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  free_plan(c->plan);
  for (auto a : c->allocs) cudaFree(a);
  delete c;
  return 0;
}

int32_t dhg_num_weights(const dhg_ctx* c) { return c ? (int32_t)c->spec.size() : 0; }
const char* dhg_weight_name(const dhg_ctx* c, int32_t i) {
  return (c && i >= 0 && i < (int)c->spec.size()) ? c->spec[i].name.c_str() : "";
}
int32_t dhg_weight_ndim(const dhg_ctx* c, int32_t i) {
  return (c && i >= 0 && i < (int)c->spec.size()) ? (int32_t)c->spec[i].shape.size() : -1;
}
int64_t dhg_weight_dim(const dhg_ctx* c, int32_t i, int32_t d) {
  if (!c || i < 0 || i >= (int)c->spec.size() || d < 0 || d >= (int)c->spec[i].shape.size()) return -1;
  return c->spec[i].shape[d];
}

int32_t dhg_load_weight(dhg_ctx* c, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  if (!c || !name || !data || !shape) return fail("dhg_load_weight: null argument");
  if (c->finalized) return fail("dhg_load_weight: ctx already finalized");
  auto it = c->spec_index.find(name);
  if (it == c->spec_index.end()) return fail("unexpected key in source state_dict: %s", name);
  const WSpec& ws = c->spec[it->second];
  bool ok = (int)ws.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = ws.shape[i] == shape[i];
  if (!ok) return fail("size mismatch for %s", name);
  c->raw[name].assign(data, data + ws.numel());
  return 0;
}

int32_t dhg_set_schedule(dhg_ctx* c, const float* beta, const float* abar) {
  if (!c || !beta || !abar) return fail("dhg_set_schedule: null argument");
  if (c->finalized) return fail("dhg_set_schedule: ctx already finalized");
  memcpy(c->beta, beta, sizeof(c->beta));
  memcpy(c->abar, abar, sizeof(c->abar));
  return 0;
}

int32_t dhg_finalize(dhg_ctx* c) {
  if (!c) return fail("null ctx");
  if (c->finalized) return 0;
  CUDA_OK(cudaSetDevice(c->device));
  std::string missing;
  int nmiss = 0;
  for (auto& ws : c->spec)
    if (!c->raw.count(ws.name)) {
      if (nmiss++ < 8) missing += (missing.empty() ? "" : ", ") + ws.name;
    }
  if (nmiss) return fail("missing keys in source state_dict (%d): %s%s", nmiss, missing.c_str(), nmiss > 8 ? ", ..." : "");

  for (const char* p : {"enc1", "enc2", "enc4", "dec3", "dec2", "dec1"})
    if (make_convblock_lins(c, p)) return 1;
  if (make_enc_lins(c, "enc3") || make_enc_lins(c, "enc5")) return 1;
  for (int i = 0; i < c->cfg.num_layers; ++i)
    if (make_enc_lins(c, "att_layers." + std::to_string(i))) return 1;
  for (const char* p : {"skip_conv1", "skip_conv2", "skip_conv3"})
    if (make_lin(c, p, {p}, true)) return 1;
  const std::string ts = "text_style_model";
  if (make_lin(c, ts + ".style_ffn.1", {ts + ".style_ffn.1"}, false) || make_lin(c, ts + ".style_ffn.3", {ts + ".style_ffn.3"}, false) ||
      make_lin(c, ts + ".text_ffn.1", {ts + ".text_ffn.1"}, false) || make_lin(c, ts + ".text_ffn.3", {ts + ".text_ffn.3"}, false) ||
      make_lin(c, ts + ".mha.wq", {ts + ".mha.wq"}, false) || make_lin(c, ts + ".mha.kv", {ts + ".mha.wk", ts + ".mha.wv"}, false) ||
      make_lin(c, ts + ".mha.dense", {ts + ".mha.dense"}, false) || make_lin(c, "att_dense", {"att_dense"}, false))
    return 1;

  // FiLM: concatenate all gamma|beta linears into one [tot,32] matrix
  std::vector<std::string> film_layers;
  for (const char* p : {"enc1", "enc2", "enc4", "dec3", "dec2", "dec1"})
    for (int i = 1; i <= 3; ++i) film_layers.push_back(std::string(p) + ".affine" + std::to_string(i));
  std::vector<std::string> encs = {"enc3", "enc5"};
  for (int i = 0; i < c->cfg.num_layers; ++i) encs.push_back("att_layers." + std::to_string(i));
  for (auto& p : encs)
    for (int i = 0; i < 4; ++i) film_layers.push_back(p + ".affine" + std::to_string(i));
  for (int i = 1; i <= 4; ++i) film_layers.push_back(ts + ".affine" + std::to_string(i));
  std::vector<float> fw, fb;
  int off = 0;
  for (auto& n : film_layers) {
    const auto& gw = c->raw.at(n + ".gamma_emb.weight");
    const auto& gb = c->raw.at(n + ".gamma_emb.bias");
    const auto& bw = c->raw.at(n + ".beta_emb.weight");
    const auto& bb = c->raw.at(n + ".beta_emb.bias");
    c->film_off[n] = off;
    fw.insert(fw.end(), gw.begin(), gw.end());
    fw.insert(fw.end(), bw.begin(), bw.end());
    fb.insert(fb.end(), gb.begin(), gb.end());
    fb.insert(fb.end(), bb.begin(), bb.end());
    off += 2 * (int)gb.size();
  }
  c->film_total = off;
  if (dev_upload(c->allocs, &c->film_W, fw) || dev_upload(c->allocs, &c->film_b, fb)) return 1;
  if (dev_upload(c->allocs, &c->sff_w1, c->raw.at("sigma_ffn.1.weight")) || dev_upload(c->allocs, &c->sff_b1, c->raw.at("sigma_ffn.1.bias")) ||
      dev_upload(c->allocs, &c->sff_w2, c->raw.at("sigma_ffn.3.weight")) || dev_upload(c->allocs, &c->sff_b2, c->raw.at("sigma_ffn.3.bias")) ||
      dev_upload(c->allocs, &c->emb, c->raw.at(ts + ".emb.weight")) ||
      dev_upload(c->allocs, &c->in_W, c->raw.at("input_dense.weight")) || dev_upload(c->allocs, &c->in_b, c->raw.at("input_dense.bias")) ||
      dev_upload(c->allocs, &c->out_W, c->raw.at("output_dense.weight")) || dev_upload(c->allocs, &c->out_b, c->raw.at("output_dense.bias")) ||
      dev_upload(c->allocs, &c->pen_W, c->raw.at("pen_lifts_dense.0.weight")) || dev_upload(c->allocs, &c->pen_b, c->raw.at("pen_lifts_dense.0.bias")))
    return 1;

  // FiLM vectors of the 60 sampling noise levels: sigma_i = sqrt(alpha_bar_i)  (inference.py:89)
  std::vector<float> sig(DHG_NUM_STEPS);
  for (int i = 0; i < DHG_NUM_STEPS; ++i) sig[i] = sqrtf(c->abar[i]);
  float *dsig = nullptr, *demb = nullptr;
  if (dev_upload(c->allocs, &dsig, sig)) return 1;
  if (dev_alloc(c->allocs, (void**)&demb, DHG_NUM_STEPS * kSigmaDim * sizeof(float))) return 1;
  if (dev_alloc(c->allocs, (void**)&c->cond60, (size_t)DHG_NUM_STEPS * c->film_total * sizeof(float))) return 1;
  launch_sigma_ffn(dsig, c->sff_w1, c->sff_b1, c->sff_w2, c->sff_b2, kSigmaHidden, demb, DHG_NUM_STEPS, 0);
  launch_film_table(demb, c->film_W, c->film_b, c->film_total, c->cond60, DHG_NUM_STEPS, 0);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  // Tail fusion.  The last ConvBlock ends with  d1 = FiLM3(fc(a2)) + skip  (cnn.py:83-87) and the heads are linear in d1
  // (model.py:179-181), so with H = [output_dense; pen_lifts_dense] (3 x C):
  //   d1 . H^T + b_H = a2 . (W_fc^T diag(gamma_s) H^T) + skip . H^T + ((b_fc * gamma_s + beta_s) . H^T + b_H)
  // gamma_s / beta_s only depend on the step: 60 tiny tables, and neither the fc GEMM nor d1 exist in the chain.
  {
    const int C = c->c1;
    const Lin& fc = c->lins.at("dec1.fc");
    const int foff = c->film_off.at("dec1.affine3");
    std::vector<float> cond((size_t)DHG_NUM_STEPS * c->film_total);
    CUDA_OK(cudaMemcpy(cond.data(), c->cond60, cond.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<float> H((size_t)3 * C), bH(3);
    const auto& ow = c->raw.at("output_dense.weight");
    const auto& pw = c->raw.at("pen_lifts_dense.0.weight");
    for (int n = 0; n < C; ++n) { H[n] = ow[n]; H[C + n] = ow[C + n]; H[2 * C + n] = pw[n]; }
    bH[0] = c->raw.at("output_dense.bias")[0]; bH[1] = c->raw.at("output_dense.bias")[1]; bH[2] = c->raw.at("pen_lifts_dense.0.bias")[0];
    std::vector<float> A((size_t)DHG_NUM_STEPS * 3 * C), cc((size_t)DHG_NUM_STEPS * 3);
    for (int s = 0; s < DHG_NUM_STEPS; ++s) {
      const float* g = cond.data() + (size_t)s * c->film_total + foff;
      const float* be = g + C;
      for (int j = 0; j < 3; ++j) {
        double acc_c = bH[j];
        for (int n = 0; n < C; ++n) acc_c += ((double)fc.h_b[n] * g[n] + be[n]) * H[(size_t)j * C + n];
        cc[(size_t)s * 3 + j] = (float)acc_c;
        for (int k = 0; k < C; ++k) {
          double a = 0.0;
          for (int n = 0; n < C; ++n) a += (double)fc.h_w[(size_t)k * C + n] * g[n] * H[(size_t)j * C + n];   // h_w: [K][N]
          A[((size_t)s * 3 + j) * C + k] = (float)a;
        }
      }
    }
    if (dev_upload(c->allocs, &c->tail_A, A) || dev_upload(c->allocs, &c->tail_c, cc) || dev_upload(c->allocs, &c->tail_H, H)) return 1;
    // dec1.conv_skip folded onto the heads: W'[tau][j][k] = sum_n H[j][n] W_skip[tau][k][n], b'[j] = sum_n H[j][n] b_skip[n]
    // (j < 3; padded to 32 output columns, the kernel's narrowest tile)
    const Lin& sk = c->lins.at("dec1.conv_skip");
    Lin F;
    F.taps = sk.taps; F.K = sk.K; F.N = 32;
    F.h_w.assign((size_t)F.taps * F.K * F.N, 0.f);
    F.h_b.assign(F.N, 0.f);
    std::vector<float> wr(F.h_w.size());
    std::vector<bf16> w16(F.h_w.size());
    for (int j = 0; j < 3; ++j) {
      double bj = 0.0;
      for (int n = 0; n < C; ++n) bj += (double)H[(size_t)j * C + n] * sk.h_b[n];
      F.h_b[j] = (float)bj;
      for (int t = 0; t < F.taps; ++t)
        for (int k = 0; k < F.K; ++k) {
          double a = 0.0;
          for (int n = 0; n < C; ++n) a += (double)H[(size_t)j * C + n] * sk.h_w[((size_t)t * sk.K + k) * sk.N + n];
          F.h_w[((size_t)t * F.K + k) * F.N + j] = (float)a;
        }
    }
    // the folded weights carry the head values directly, so their bf16 rounding would show up in eps / pen: columns
    // j + 3 hold the rounding residual of column j (the tile has 29 spare columns) and the dot vectors add both
    for (int t = 0; t < F.taps; ++t)
      for (int k = 0; k < F.K; ++k)
        for (int j = 0; j < 3; ++j) {
          const float a = F.h_w[((size_t)t * F.K + k) * F.N + j];
          F.h_w[((size_t)t * F.K + k) * F.N + j + 3] = a - __bfloat162float(__float2bfloat16_rn(a));
        }
    for (int t = 0; t < F.taps; ++t)
      for (int k = 0; k < F.K; ++k)
        for (int n = 0; n < F.N; ++n) {
          const bf16 hb = __float2bfloat16_rn(F.h_w[((size_t)t * F.K + k) * F.N + n]);
          wr[((size_t)t * F.K + k) * F.N + n] = __bfloat162float(hb);
          w16[((size_t)t * F.N + n) * F.K + k] = hb;
        }
    if (dev_upload(c->allocs, &F.w32, F.h_w) || dev_upload(c->allocs, &F.w32r, wr) || dev_upload(c->allocs, &F.bias, F.h_b)) return 1;
    if (dev_alloc(c->allocs, (void**)&F.w16, w16.size() * sizeof(bf16))) return 1;
    CUDA_OK(cudaMemcpy(F.w16, w16.data(), w16.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    c->lins["dec1.conv_skip.heads"] = std::move(F);
    std::vector<float> pick((size_t)3 * 32, 0.f);
    for (int j = 0; j < 3; ++j) { pick[(size_t)j * 32 + j] = 1.f; pick[(size_t)j * 32 + j + 3] = 1.f; }
    if (dev_upload(c->allocs, &c->tail_pick, pick)) return 1;
  }
  // Skip fusion tables.  A ConvBlock ends with  out = FiLM3(fc(a2)) + conv_skip(x)  (cnn.py:66,81-87).  In sampling the
  // FiLM vectors only depend on the step, so
  //   out = a2 . (diag(gamma_s) W_fc)^T + sum_tap x[t + tap] . W_skip[tap]^T + (gamma_s * b_fc + beta_s + b_skip)
  // is ONE accumulation over two operands: 60 scaled copies of W_fc (bf16) and bias vectors (fp32) per block.
  {
    std::vector<float> cond((size_t)DHG_NUM_STEPS * c->film_total);
    CUDA_OK(cudaMemcpy(cond.data(), c->cond60, cond.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (const char* blk : {"enc1", "enc2", "enc4", "dec3", "dec2"}) {
      const std::string p = blk;
      const Lin& fc = c->lins.at(p + ".fc");
      const Lin& sk = c->lins.at(p + ".conv_skip");
      const int N = fc.N, K = fc.K, foff = c->film_off.at(p + ".affine3");
      std::vector<bf16> w((size_t)DHG_NUM_STEPS * N * K);
      std::vector<float> bias((size_t)DHG_NUM_STEPS * N);
      for (int st = 0; st < DHG_NUM_STEPS; ++st) {
        const float* g = cond.data() + (size_t)st * c->film_total + foff;
        const float* be = g + N;
        for (int n = 0; n < N; ++n) {
          bias[(size_t)st * N + n] = g[n] * fc.h_b[n] + be[n] + sk.h_b[n];
          for (int k = 0; k < K; ++k) w[((size_t)st * N + n) * K + k] = __float2bfloat16_rn(g[n] * fc.h_w[(size_t)k * N + n]);   // h_w: [K][N]
        }
      }
      bf16* dw = nullptr;
      float* db = nullptr;
      if (dev_alloc(c->allocs, (void**)&dw, w.size() * sizeof(bf16)) || dev_upload(c->allocs, &db, bias)) return 1;
      CUDA_OK(cudaMemcpy(dw, w.data(), w.size() * sizeof(bf16), cudaMemcpyHostToDevice));
      c->fc60_w[p] = dw;
      c->fc60_bias[p] = db;
    }
  }
  // Head fusion tables (kernels_simt.cu skip_from_x_kernel)
  {
    const int C = c->c1;
    const Lin& sk = c->lins.at("enc1.conv_skip");   // h_w [3][K = C][N = C]
    const auto& iw = c->raw.at("input_dense.weight");   // [C][2]
    const auto& ib = c->raw.at("input_dense.bias");
    std::vector<float> M((size_t)3 * 2 * C), v((size_t)3 * C);
    for (int t = 0; t < 3; ++t)
      for (int n = 0; n < C; ++n) {
        double m0 = 0.0, m1 = 0.0, vv = 0.0;
        for (int k = 0; k < C; ++k) {
          const double w = sk.h_w[((size_t)t * C + k) * C + n];
          m0 += (double)iw[(size_t)k * 2] * w;
          m1 += (double)iw[(size_t)k * 2 + 1] * w;
          vv += (double)ib[k] * w;
        }
        M[((size_t)t * 2 + 0) * C + n] = (float)m0;
        M[((size_t)t * 2 + 1) * C + n] = (float)m1;
        v[(size_t)t * C + n] = (float)vv;
      }
    if (dev_upload(c->allocs, &c->head_M, M) || dev_upload(c->allocs, &c->head_v, v) || dev_upload(c->allocs, &c->head_b, sk.h_b)) return 1;
  }
  c->raw.clear();
  c->finalized = true;
  return 0;
}

int32_t dhg_plan(dhg_ctx* c, int32_t B, int32_t T, int32_t L, int32_t S, int32_t precision) {
  if (check_ready(c, false)) return 1;
  if (B < 1 || L < 1 || S < 1) return fail("dhg_plan: B, L, S must be >= 1");
  if (T < 8 || T % 8) return fail("dhg_plan: T must be a positive multiple of 8 (inference.py:78); got %d", T);
  if (precision != DHG_PREC_FP32 && precision != DHG_PREC_BF16) return fail("dhg_plan: bad precision %d", precision);
  if ((long long)B * (T + 1) + 1 >= (1LL << 24)) return fail("dhg_plan: B*T too large for one chunk (needs B*(T+1) < 2^24 rows); plan a smaller B and let dhg_sample chunk");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaDeviceSynchronize());
  free_plan(c->plan);
  c->plan = nullptr;
  Plan* P = new Plan();
  P->B = B; P->T = T; P->L = L; P->S = S; P->prec = precision;
  P->gemm_impl = c->opt_gemm;   // 1: tcgen05 GEMMs (bf16 storage, or split storage in fp32 precision); 0: CUDA-core GEMMs
  P->split = precision == DHG_PREC_FP32 && P->gemm_impl == 1;
  P->attn_impl = (precision == DHG_PREC_BF16 || P->split) ? c->opt_attn : 0;   // tcgen05 attention (bf16 or split storage)
  if (build_plan(c, P)) { free_plan(P); return 1; }
  CUDA_OK(cudaDeviceSynchronize());
  c->plan = P;
  return 0;
}

int32_t dhg_denoise(dhg_ctx* c, const float* strokes, const int64_t* text, const float* sigma, const float* style,
                    float* eps, float* pen, void* stream) {
  if (check_ready(c, true)) return 1;
  if (!strokes || !text || !sigma || !style || !eps || !pen) return fail("dhg_denoise: null argument");
  Plan* P = c->plan;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaSetDevice(c->device));
  if (plan_enter(P, st)) return 1;
  CUDA_OK(cudaMemcpyAsync(P->text, text, (size_t)P->B * P->L * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaMemcpyAsync(P->style, style, (size_t)P->B * P->S * kStyleWidth * sizeof(float), cudaMemcpyDeviceToDevice, st));
  launch_sigma_ffn(sigma, c->sff_w1, c->sff_b1, c->sff_w2, c->sff_b2, kSigmaHidden, P->sig_emb, P->B, st);
  launch_film_table(P->sig_emb, c->film_W, c->film_b, c->film_total, P->cond_b, P->B, st);
  StepCtx sc;
  memset(&sc, 0, sizeof(sc));
  sc.cond = P->cond_b;
  sc.bstride = c->film_total;
  sc.head.B = P->B;
  sc.head.T = P->T;
  sc.head.eps_out = eps;
  sc.head.pen_out = pen;
  sc.head.pen_stride = 1;
  sc.head.pen_offset = 0;
  if (run_ops(P->once_ops, st, sc)) return 1;
  if (run_ops(P->text_ops[0], st, sc)) return 1;
  // input_dense reads sc.head.x_io when set; here it must read the caller's strokes without updating them
  StepCtx sc_in = sc;
  sc_in.head.x_io = const_cast<float*>(strokes);
  for (size_t i = 0; i < P->step_ops.size(); ++i) {
    const bool is_head = (i + 1 == P->step_ops.size());
    if (P->step_ops[i](st, is_head ? sc : sc_in)) return 1;
    if (getenv("DHG_SYNC_OPS")) {
      const cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return fail("step op %zu (%s) faulted: %s", i, P->step_ops[i].name.c_str(), cudaGetErrorString(e));
    }
  }
  CUDA_OK(cudaGetLastError());
  c->last_launches = 2 + P->launches_once + P->launches_text + P->launches_step;
  return plan_leave(P, st);
}

int32_t dhg_sample(dhg_ctx* c, int32_t batch, const float* x0, const float* noise, uint64_t seed, const int64_t* text,
                   const float* style, int32_t mode, float* out, void* stream) {
  if (check_ready(c, true)) return 1;
  if (batch < 1 || !x0 || !text || !style || !out) return fail("dhg_sample: bad argument");
  if (mode != DHG_MODE_NEW && mode != DHG_MODE_STANDARD) return fail("dhg_sample: bad diffusion mode %d", mode);
  if (!noise) return fail("dhg_sample: noise is mandatory: pass the injected draws [60,batch,T,2] (the library has no generator of its own)");
  (void)seed;
  Plan* P = c->plan;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaSetDevice(c->device));
  if (plan_enter(P, st)) return 1;
  const size_t xs = (size_t)P->T * 2, ss = (size_t)P->S * kStyleWidth;
  if (!P->noise) {
    if (dev_alloc(P->allocs, (void**)&P->noise, (size_t)DHG_NUM_STEPS * P->B * xs * sizeof(float), &P->bytes)) return 1;
  }
  int64_t launches = 0;
  for (int c0 = 0; c0 < batch; c0 += P->B) {
    const int nb = batch - c0 < P->B ? batch - c0 : P->B;
    if (nb < P->B) {  // ragged last chunk: pad with all-masked text / zero inputs, results discarded
      CUDA_OK(cudaMemsetAsync(P->text, 0, (size_t)P->B * P->L * sizeof(int64_t), st));
      CUDA_OK(cudaMemsetAsync(P->style, 0, (size_t)P->B * ss * sizeof(float), st));
      CUDA_OK(cudaMemsetAsync(P->x_state, 0, (size_t)P->B * xs * sizeof(float), st));
      CUDA_OK(cudaMemsetAsync(P->noise, 0, (size_t)DHG_NUM_STEPS * P->B * xs * sizeof(float), st));
    }
    CUDA_OK(cudaMemcpyAsync(P->text, text + (size_t)c0 * P->L, (size_t)nb * P->L * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->style, style + (size_t)c0 * ss, (size_t)nb * ss * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->x_state, x0 + (size_t)c0 * xs, (size_t)nb * xs * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemcpy2DAsync(P->noise, (size_t)P->B * xs * sizeof(float), noise + (size_t)c0 * xs,
                              (size_t)batch * xs * sizeof(float), (size_t)nb * xs * sizeof(float), DHG_NUM_STEPS,
                              cudaMemcpyDeviceToDevice, st));
    if (launch_chain(c, P, mode, true, st)) return 1;
    CUDA_OK(cudaMemcpyAsync(out + (size_t)c0 * P->T * 3, P->out, (size_t)nb * P->T * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    launches += P->launches_once + (int64_t)DHG_NUM_STEPS * (P->launches_text + P->launches_step) - (DHG_NUM_STEPS - 1);   // input_dense is fused after step 1
    if (P->opt_tail_fusion && P->prec == PREC_BF16 && P->gemm_impl == 1) launches -= DHG_NUM_STEPS;   // dec1.fc lives in the head kernel
    for (int b = 0; b < 5; ++b)
      if (P->opt_skip_fusion & (1 << b)) launches -= DHG_NUM_STEPS;   // that block's conv_skip lives in its last GEMM
  }
  c->last_launches = launches;
  return plan_leave(P, st);
}

int32_t dhg_sample_host(dhg_ctx* c, int32_t batch, const float* x0, const float* noise, uint64_t seed, const int64_t* text,
                        const float* style, int32_t mode, float* out) {
  if (check_ready(c, true)) return 1;
  if (batch < 1 || !x0 || !text || !style || !out) return fail("dhg_sample_host: bad argument");
  if (!noise) return fail("dhg_sample_host: noise is mandatory: pass the injected draws [60,batch,T,2]");
  Plan* P = c->plan;
  for (size_t i = 0, n = (size_t)batch * P->L; i < n; ++i)   // host ids: checked here, before anything is copied
    if (text[i] < 0 || text[i] >= kVocab) return fail("text token id %lld at index %zu out of range [0, %d)", (long long)text[i], i, kVocab);
  CUDA_OK(cudaSetDevice(c->device));
  const size_t xs = (size_t)P->T * 2, ss = (size_t)P->S * kStyleWidth;
  if (mode != DHG_MODE_NEW && mode != DHG_MODE_STANDARD) return fail("dhg_sample_host: bad diffusion mode %d", mode);
  if (batch == P->B && g_opt_host_overlap) {
    // One chunk that fills the plan: the host buffers go straight into the plan's own buffers (no staging copy), and
    // only the noise of the first steps is waited for -- the chain starts with step 59, i.e. the END of the noise array;
    // the rest (steps 59 - k .. 0, 5/6 of it) travels on a second stream while the first k steps run.  The chain is two
    // graphs with the join between them; same kernels, same order, same bits as dhg_sample.
    cudaStream_t st = P->cap_stream;
    if (!P->copy_stream) {
      CUDA_OK(cudaStreamCreateWithFlags(&P->copy_stream, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&P->ev_copy_go, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&P->ev_copy_done, cudaEventDisableTiming));
    }
    if (!P->noise && dev_alloc(P->allocs, (void**)&P->noise, (size_t)DHG_NUM_STEPS * P->B * xs * sizeof(float), &P->bytes)) return 1;
    if (plan_enter(P, st)) return 1;
    const int k = host_head_steps(P);
    const size_t step_elems = (size_t)P->B * xs, tail_steps = (size_t)(DHG_NUM_STEPS - k);
    CUDA_OK(cudaEventRecord(P->ev_copy_go, st));   // the plan's noise buffer may still be read by an earlier call
    CUDA_OK(cudaStreamWaitEvent(P->copy_stream, P->ev_copy_go, 0));
    CUDA_OK(cudaMemcpyAsync(P->text, text, (size_t)batch * P->L * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->style, style, (size_t)batch * ss * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->x_state, x0, (size_t)batch * xs * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->noise + tail_steps * step_elems, noise + tail_steps * step_elems, (size_t)k * step_elems * sizeof(float),
                            cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(P->noise, noise, tail_steps * step_elems * sizeof(float), cudaMemcpyHostToDevice, P->copy_stream));
    CUDA_OK(cudaEventRecord(P->ev_copy_done, P->copy_stream));
    if (launch_chain_part(c, P, mode, 0, st)) return 1;
    CUDA_OK(cudaStreamWaitEvent(st, P->ev_copy_done, 0));
    if (launch_chain_part(c, P, mode, 1, st)) return 1;
    CUDA_OK(cudaMemcpyAsync(out, P->out, (size_t)batch * P->T * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    int64_t launches = P->launches_once + (int64_t)DHG_NUM_STEPS * (P->launches_text + P->launches_step) - (DHG_NUM_STEPS - 1);
    if (P->opt_tail_fusion && P->prec == PREC_BF16 && P->gemm_impl == 1) launches -= DHG_NUM_STEPS;
    for (int b = 0; b < 5; ++b)
      if (P->opt_skip_fusion & (1 << b)) launches -= DHG_NUM_STEPS;
    c->last_launches = launches;
    if (plan_leave(P, st)) return 1;
    CUDA_OK(cudaStreamSynchronize(st));
    return plan_check_flags(P);
  }
  // device staging for the host buffers: owned by the plan and kept between calls (no allocation, clearing or freeing
  // inside the call once it has been sized)
  Plan::HostStage& hs = P->stage;
  if (batch > hs.cap) {
    CUDA_OK(cudaDeviceSynchronize());
    for (void* p : {(void*)hs.x, (void*)hs.noise, (void*)hs.style, (void*)hs.out, (void*)hs.text})
      if (p) cudaFree(p);
    hs = Plan::HostStage();
    CUDA_OK(cudaMalloc((void**)&hs.x, (size_t)batch * xs * sizeof(float)));
    CUDA_OK(cudaMalloc((void**)&hs.style, (size_t)batch * ss * sizeof(float)));
    CUDA_OK(cudaMalloc((void**)&hs.text, (size_t)batch * P->L * sizeof(int64_t)));
    CUDA_OK(cudaMalloc((void**)&hs.out, (size_t)batch * P->T * 3 * sizeof(float)));
    CUDA_OK(cudaMalloc((void**)&hs.noise, (size_t)DHG_NUM_STEPS * batch * xs * sizeof(float)));
    hs.cap = batch;
  }
  const bool timing = getenv("DHG_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(P->cap_stream);
    fprintf(stderr, "dhg_sample_host: %-10s at %.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  cudaStream_t st = P->cap_stream;
  if (plan_enter(P, st)) return 1;   // the staging buffers may still be read by an earlier stream-ordered call
  CUDA_OK(cudaMemcpyAsync(hs.x, x0, (size_t)batch * xs * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(hs.style, style, (size_t)batch * ss * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(hs.text, text, (size_t)batch * P->L * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (noise) CUDA_OK(cudaMemcpyAsync(hs.noise, noise, (size_t)DHG_NUM_STEPS * batch * xs * sizeof(float), cudaMemcpyHostToDevice, st));
  lap("h2d");
  if (dhg_sample(c, batch, hs.x, noise ? hs.noise : nullptr, seed, hs.text, hs.style, mode, hs.out, st)) return 1;
  lap("chain");
  CUDA_OK(cudaMemcpyAsync(out, hs.out, (size_t)batch * P->T * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  lap("d2h");
  return plan_check_flags(P);
}

int32_t dhg_posterior_step(dhg_ctx* c, int32_t step, int32_t mode, const float* x, const float* eps, const float* noise,
                           float* out, int64_t n, void* stream) {
  if (!c) return fail("null ctx");
  if (step < 0 || step >= DHG_NUM_STEPS) return fail("dhg_posterior_step: step %d out of [0,60)", step);
  if (n < 0 || n % 4) return fail("dhg_posterior_step: n must be a non-negative multiple of 4");
  if (mode != DHG_MODE_NEW && mode != DHG_MODE_STANDARD) return fail("dhg_posterior_step: bad mode");
  if (n == 0) return 0;
  if (!x || !eps || !out) return fail("dhg_posterior_step: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  const float beta = c->beta[step], abar = c->abar[step];
  float c_eps, c_eps2 = 1.f, c_div, c_noise;
  if (mode == DHG_MODE_NEW) {
    c_eps = sqrtf(1.f - abar);
    c_div = sqrtf(1.f - beta);
    c_noise = sqrtf(1.f - (step > 1 ? c->abar[step - 1] : 1.0f));
  } else {
    c_eps = beta;
    c_eps2 = sqrtf(1.f - abar);
    c_div = 1.f / sqrtf(1.f - beta);
    c_noise = step > 0 ? sqrtf(beta) : 0.f;
  }
  launch_posterior(x, eps, noise, out, (size_t)n, mode, c_eps, c_eps2, c_div, c_noise, c->num_sms, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  c->last_launches = 1;
  return 0;
}

int32_t dhg_check_errors(dhg_ctx* c, void* stream) {
  if (check_ready(c, true)) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  CUDA_OK(cudaGetLastError());
  return plan_check_flags(c->plan);
}

int64_t dhg_last_launch_count(const dhg_ctx* c) { return c ? c->last_launches : 0; }
int64_t dhg_plan_bytes(const dhg_ctx* c) { return (c && c->plan) ? (int64_t)c->plan->bytes : 0; }

int32_t dhg_set_option(dhg_ctx* c, const char* key, int32_t value) {
  if (key && !strcmp(key, "interleave")) { tc_gemm_set_option(4, value); return 0; }
  if (key && !strcmp(key, "attn_dbg")) { attn_tc_set_debug(value); return 0; }
  if (key && !strcmp(key, "attn_early")) { attn_tc_set_debug(value ? -201 : -200); return 0; }
  if (key && !strcmp(key, "attn_max_slots")) { attn_tc_set_debug(-300 - (value < 0 ? 0 : value > 6 ? 6 : value)); return 0; }
  if (key && !strcmp(key, "pair")) { tc_gemm_set_option(7, value); return 0; }
  if (key && !strcmp(key, "pdl")) { tc_gemm_set_option(6, value); attn_tc_set_debug(value ? -101 : -100); return 0; }
  if (key && !strcmp(key, "w_resident")) { tc_gemm_set_option(2, value); return 0; }
  if (key && !strcmp(key, "text_sets")) { g_opt_text_sets = value; return 0; }
  if (key && !strcmp(key, "host_overlap")) { g_opt_host_overlap = value ? 1 : 0; return 0; }
  if (key && !strcmp(key, "autotune")) { g_opt_autotune = value ? 1 : 0; return 0; }
  if (key && !strcmp(key, "serpentine")) { g_opt_serpentine = value ? 1 : 0; return 0; }
  if (key && !strcmp(key, "attn_keyblock_auto")) { g_opt_attn_keyblock_auto = value ? 1 : 0; return 0; }
  if (key && !strcmp(key, "l2_hints")) { g_opt_l2_hints = value ? 1 : 0; return 0; }
  if (key && !strcmp(key, "tail_fusion")) { g_opt_tail_fusion = value < 0 ? 0 : value > 3 ? 3 : value; return 0; }
  if (key && !strcmp(key, "head_fusion")) { g_opt_head_fusion = value ? 1 : 0; return 0; }
  if (key && !strcmp(key, "skip_fusion")) { g_opt_skip_fusion = value & 31; return 0; }
  if (key && !strcmp(key, "tune_rev")) { tc_gemm_set_option(14, value); return 0; }
  if (key && !strcmp(key, "tune_bn")) { tc_gemm_set_option(10, value); return 0; }
  if (key && !strcmp(key, "tune_g")) { tc_gemm_set_option(11, value); return 0; }
  if (key && !strcmp(key, "tune_resident")) { tc_gemm_set_option(12, value); return 0; }
  if (key && !strcmp(key, "tune_pair")) { tc_gemm_set_option(13, value); return 0; }
  if (key && !strcmp(key, "specialize")) { tc_gemm_set_option(3, value); return 0; }
  if (key && !strcmp(key, "max_stages_a")) { tc_gemm_set_option(15, value); return 0; }
  if (key && !strcmp(key, "direct_store")) { tc_gemm_set_option(16, value); return 0; }
  if (key && !strcmp(key, "split_n")) { tc_gemm_set_option(17, value); return 0; }
  if (!c || !key) return fail("dhg_set_option: null argument");
  if (!strcmp(key, "gemm")) c->opt_gemm = value ? 1 : 0;
  else if (!strcmp(key, "graph")) c->opt_graph = value ? 1 : 0;
  else if (!strcmp(key, "attn")) c->opt_attn = value ? 1 : 0;
  else return fail("dhg_set_option: unknown key %s", key);
  return 0;
}

int64_t dhg_debug_read(dhg_ctx* c, const char* name, float* host_out, int64_t capacity) {
  if (check_ready(c, true)) return -1;
  Plan* P = c->plan;
  auto it = P->taps.find(name ? name : "");
  if (it == P->taps.end()) { fail("dhg_debug_read: unknown activation %s", name ? name : "(null)"); return -1; }
  const Act& a = it->second.a;
  const int period = it->second.period, pad = it->second.pad, per = period - pad;
  const int64_t n = (int64_t)P->B * per * a.C;
  if (!host_out) return n;
  if (capacity < n) { fail("dhg_debug_read: capacity %lld < %lld", (long long)capacity, (long long)n); return -1; }
  if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { fail("dhg_debug_read: device error"); return -1; }
  if (plan_check_flags(P)) return -1;
  std::vector<char> tmp((size_t)a.rows * a.C * P->esize);
  if (cudaMemcpy(tmp.data(), a.p, tmp.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { fail("dhg_debug_read: copy failed"); return -1; }
  for (int b = 0; b < P->B; ++b)
    for (int t = 0; t < per; ++t) {
      const size_t r = (size_t)b * period + pad + t;
      for (int ch = 0; ch < a.C; ++ch) {
        const size_t src = r * a.C + ch;
        float val;
        if (P->split) {   // groups of 32 hi | 32 lo bf16 (common.cuh bfs)
          const bf16* rowp = reinterpret_cast<const bf16*>(tmp.data()) + r * 2 * a.C + (size_t)(ch / 32) * 64 + (ch % 32);
          val = __bfloat162float(rowp[0]) + __bfloat162float(rowp[32]);
        } else if (P->esize == 4) val = reinterpret_cast<const float*>(tmp.data())[src];
        else val = __bfloat162float(reinterpret_cast<const bf16*>(tmp.data())[src]);
        host_out[((size_t)b * per + t) * a.C + ch] = val;
      }
    }
  return n;
}

int32_t dhg_debug_tc_gemm(int32_t device, const void* a, int32_t lda, int32_t rows, const void* w, int32_t K, int32_t N,
                          int32_t taps, const float* bias, void* out, void* stream) {
  CUDA_OK(cudaSetDevice(device));
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.bias = bias;
  e.out_raw = out;
  e.out_raw_pitch = N;
  e.map = RowMap{rows > 0 ? rows : 1, 0, rows};
  char buf[512];
  TcGemmPlan* p = tc_gemm_plan_create((const bf16*)a, lda, rows, (const bf16*)w, K, N, taps, e, buf, sizeof(buf));
  if (!p) return fail("dhg_debug_tc_gemm: %s", buf);
  const int rc = tc_gemm_launch(p, e, (cudaStream_t)stream);
  cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
  tc_gemm_plan_destroy(p);
  if (rc) return 1;
  if (ce != cudaSuccess) return fail("dhg_debug_tc_gemm: %s", cudaGetErrorString(ce));
  return 0;
}

int32_t dhg_debug_tc_gemm_ex(int32_t device, const void* a, int32_t lda, int32_t rows, const void* w, int32_t K, int32_t N,
                             int32_t taps, const dhg_debug_epilogue* d, int32_t repeats, float* ms_per_launch, void* stream) {
  if (!d) return fail("dhg_debug_tc_gemm_ex: null epilogue");
  CUDA_OK(cudaSetDevice(device));
  Epilogue e;
  memset(&e, 0, sizeof(e));
  e.bias = d->bias; e.rowbias16 = d->rowbias; e.rowbias16_cols = d->rowbias_cols;
  e.res_pre = d->res_pre; e.res_pre_pitch = d->res_pre_pitch;
  e.ln = d->ln;
  e.gamma = d->gamma; e.beta = d->beta; e.film_bstride = d->film_bstride; e.film_planned = d->gamma ? 1 : 0;
  e.res_post = d->res_post; e.res_post_pitch = d->res_post_pitch; e.res_post_up = d->res_post_up;
  e.res_post_period_lo = d->res_post_period_lo;
  e.out_raw = d->out_raw; e.out_raw_pitch = d->out_raw_pitch;
  e.out_act = d->out_act; e.out_act_pitch = d->out_act_pitch;
  e.dot_w = d->dot_w; e.dot_out = d->dot_out; e.dot_act = d->dot_act; e.dot_planned = d->dot_w ? 1 : 0;
  e.split_io = d->split_io;
  e.w_row_off = d->w_row_off;
  e.map = RowMap{d->period > 0 ? d->period : (rows > 0 ? rows : 1), d->pad_first, d->nvalid > 0 ? d->nvalid : rows};
  char buf[512];
  TcDual dual{(const bf16*)d->dual_a2, d->dual_lda2, d->dual_K2, (const bf16*)d->dual_w2, d->dual_w1_rows};
  TcGemmPlan* p = tc_gemm_plan_create((const bf16*)a, lda, rows, (const bf16*)w, K, N, taps, e, buf, sizeof(buf), nullptr, d->dual_a2 ? &dual : nullptr);
  if (!p) return fail("dhg_debug_tc_gemm_ex: %s", buf);
  cudaStream_t st = (cudaStream_t)stream;
  if (getenv("DHG_DESCRIBE")) { tc_gemm_describe(p, buf, sizeof(buf)); fprintf(stderr, "tc_gemm plan: %s\n", buf); }
  if (getenv("DHG_TRACE")) {   // timeline of CTA 0 for one launch, printed to stderr
    const int cap = 4096, roles = 10;   // entries per role: producer, MMA, 8 epilogue warps
    unsigned long long* tb = nullptr;
    cudaMalloc(&tb, roles * cap * sizeof(unsigned long long));
    cudaMemset(tb, 0, roles * cap * sizeof(unsigned long long));
    tc_gemm_launch(p, e, st);   // warm
    cudaStreamSynchronize(st);
    tc_gemm_set_trace(p, tb, cap);
    tc_gemm_launch(p, e, st);
    cudaStreamSynchronize(st);
    tc_gemm_set_trace(p, nullptr, 0);
    std::vector<unsigned long long> h((size_t)roles * cap);
    cudaMemcpy(h.data(), tb, (size_t)roles * cap * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(tb);
    unsigned long long t0 = ~0ull;
    for (auto x : h) if (x && (x >> 16) < t0) t0 = x >> 16;
    for (size_t i = 0; i < h.size(); ++i) {   // TR: producer, MMA and first epilogue warp (tools/trace_summary.py); TW: the other epilogue warps
      const unsigned long long x = h[i];
      const int role = (int)(i / cap);
      if (x) fprintf(stderr, "%s %llu %02x %u %d\n", role <= 2 ? "TR" : "TW", (x >> 16) - t0, (unsigned)((x >> 8) & 0xff), (unsigned)(x & 0xff), role);
    }
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int lrc = tc_gemm_launch(p, e, st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < repeats && !lrc; ++i) lrc = tc_gemm_launch(p, e, st);
  if (lrc) { tc_gemm_plan_destroy(p); return fail("dhg_debug_tc_gemm_ex: launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
  cudaEventRecord(e1, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (ce == cudaSuccess && repeats > 0) cudaEventElapsedTime(&ms, e0, e1);
  if (ms_per_launch) *ms_per_launch = repeats > 0 ? ms / repeats : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  tc_gemm_plan_destroy(p);
  if (ce != cudaSuccess) return fail("dhg_debug_tc_gemm_ex: %s", cudaGetErrorString(ce));
  return 0;
}

int32_t dhg_debug_time_text(dhg_ctx* c, int32_t sets, int32_t repeats, float* ms_per_step) {
  if (check_ready(c, true)) return 1;
  Plan* P = c->plan;
  if (sets < 1 || sets > P->text_sets || repeats < 1 || !ms_per_step) return fail("dhg_debug_time_text: bad argument");
  CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = P->cap_stream;
  StepCtx sc;
  memset(&sc, 0, sizeof(sc));
  sc.cond = c->cond60;
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  for (int r = -2; r < repeats; ++r) {
    if (r == 0) CUDA_OK(cudaEventRecord(e0, st));
    if (sets > 1) CUDA_OK(cudaEventRecord(P->ev_fork, st));
    for (int set = 0; set < sets; ++set) {
      cudaStream_t ts = set == 0 ? st : P->text_stream[set];
      if (set != 0) CUDA_OK(cudaStreamWaitEvent(ts, P->ev_fork, 0));
      if (run_ops(P->text_ops[set], ts, sc)) return 1;
      if (set != 0) CUDA_OK(cudaEventRecord(P->ev_text[set], ts));
    }
    for (int set = 1; set < sets; ++set) CUDA_OK(cudaStreamWaitEvent(st, P->ev_text[set], 0));
  }
  CUDA_OK(cudaEventRecord(e1, st));
  CUDA_OK(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_step = ms / (float)(repeats * sets);
  return 0;
}

int32_t dhg_debug_attention(int32_t device, const dhg_debug_attn* d, int32_t impl, int32_t repeats, float* ms_per_launch,
                            void* stream) {
  if (!d) return fail("dhg_debug_attention: null argument");
  CUDA_OK(cudaSetDevice(device));
  AttnParams a;
  a.q = d->q; a.k = d->k; a.v = d->v; a.o = d->o;
  a.q_pitch = d->q_pitch; a.k_pitch = d->k_pitch; a.v_pitch = d->v_pitch; a.o_pitch = d->o_pitch;
  a.q_period = d->q_period; a.q_pad = d->q_pad; a.k_period = d->k_period; a.k_pad = d->k_pad;
  a.B = d->B; a.H = d->H; a.D = d->D; a.Tq = d->Tq; a.Tk = d->Tk;
  a.scale = 1.0f / sqrtf((float)d->D);
  a.text = d->text;
  a.split = impl == 3 ? 1 : 0;   // 3: tcgen05 kernel on split storage (the fp32-contract mode); pitches in elements
  cudaStream_t st = (cudaStream_t)stream;
  AttnTcPlan* ap = nullptr;
  if (impl >= 1) {   // 2: the key-block kernel also where all keys would fit at once (128 < Tk <= 256)
    char buf[512];
    ap = attn_tc_plan_create(a, d->q_rows, d->k_rows, buf, sizeof(buf), impl == 2 ? 1 : 0);
    if (!ap) return fail("dhg_debug_attention: %s", buf);
  }
  auto launch = [&]() -> int { return ap ? attn_tc_launch(ap, st) : launch_attention_simt<bf16>(a, st); };
  if (ap && getenv("DHG_TRACE") && !attn_tc_plan_is_long(ap)) {   // timeline of CTA 0 / slot 0 for one launch, printed to stderr
    const int cap = 2048;
    unsigned long long* tb = nullptr;
    cudaMalloc(&tb, 2 * cap * sizeof(unsigned long long));
    cudaMemset(tb, 0, 2 * cap * sizeof(unsigned long long));
    launch();
    cudaStreamSynchronize(st);
    attn_tc_plan_set_trace(ap, tb, cap);
    launch();
    cudaStreamSynchronize(st);
    attn_tc_plan_set_trace(ap, nullptr, 0);
    std::vector<unsigned long long> h(2 * cap);
    cudaMemcpy(h.data(), tb, 2 * cap * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(tb);
    unsigned long long t0 = ~0ull;
    for (auto x : h) if (x && (x >> 16) < t0) t0 = x >> 16;
    fprintf(stderr, "attention plan: slots=%d\n", attn_tc_plan_slots(ap));
    for (auto x : h)
      if (x) fprintf(stderr, "ATR %llu %02x %u\n", (x >> 16) - t0, (unsigned)((x >> 8) & 0xff), (unsigned)(x & 0xff));
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = launch();
  cudaEventRecord(e0, st);
  for (int i = 0; i < repeats && !rc; ++i) rc = launch();
  cudaEventRecord(e1, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (ce == cudaSuccess && repeats > 0) cudaEventElapsedTime(&ms, e0, e1);
  if (ms_per_launch) *ms_per_launch = repeats > 0 ? ms / repeats : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ap) attn_tc_plan_destroy(ap);
  if (rc) return fail("dhg_debug_attention: unsupported shape");
  if (ce != cudaSuccess) return fail("dhg_debug_attention: %s", cudaGetErrorString(ce));
  return 0;
}

}  // extern "C"
