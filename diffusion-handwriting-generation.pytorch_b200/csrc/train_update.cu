// The update half of the reference's training step (SURVEY.md 8e "optional train step", 8f-3), fp32 CUDA kernels:
//
//   train.py:38-43   x_perturbed = sqrt(alpha) * x + sqrt(1 - alpha) * eps                  -> dhg_train_perturb
//   loss.py:5-39     score_loss + pen_lifts_loss (+ their gradients w.r.t. the predictions)   -> dhg_train_loss
//   train.py:57-61   dispatch_clip_grad(..., mode "norm") = torch clip_grad_norm_              -> dhg_train_sqnorm + the
//                    clip coefficient computed on the device inside the optimiser kernel (no host round trip)
//   train.py:63      InvSqrtScheduledOptim.step_and_update_lr (scheduler.py:1-35) around torch.optim.Adam(lr, betas,
//                    weight_decay) (config.yml:33-38)                                           -> dhg_train_adam_step
//
// and the one exchange a data-parallel step has (one all-reduce of 10,028,451 fp32 gradients, then / N): the
// caller sums the flat gradient over the ranks (NCCL through torch.distributed: plumbing), the 1/N and the clip
// coefficient are folded into the optimiser kernel, so the gradient is read exactly once after the exchange.
//
// The forward + backward pass of the denoiser that produces the flat gradient is train_step.cu (dhg_trainer_*); these
// entry points take the flat gradient as an input.
//
// All of it is HBM-bound streaming: Adam reads p, g, m, v and writes p, m, v (28 bytes per parameter: 281 MB per step for
// the reference's 10.0 M parameters), the reductions are two-stage and deterministic (fixed block partials in double).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dhg_b200.h"

namespace {

thread_local char g_terr[512] = "";
int tfail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_terr, sizeof(g_terr), fmt, ap);
  va_end(ap);
  return 1;
}
#define T_OK(call)                                                                           \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) return tfail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

constexpr int kRedBlocks = 592;   // 4 per SM: fixed, so that the partial sums (and the result) do not depend on anything else
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sm[kRedThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // sm may still be read by a previous call
  if (l == 0) sm[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = l < kRedThreads / 32 ? sm[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

// train.py:40-43.  x, eps, out: [B, T, 2]; alphas: [B] (the reference's [B, 1]).
__global__ void perturb_kernel(const float* __restrict__ x, const float* __restrict__ alphas, const float* __restrict__ eps,
                               float* __restrict__ out, int B, int T) {
  const size_t n = (size_t)B * T * 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a = alphas[i / ((size_t)T * 2)];
    out[i] = sqrtf(a) * x[i] + sqrtf(1.f - a) * eps[i];
  }
}

// loss.py:27-37 and the gradients of the total loss w.r.t. score_pred [B, T, 2] and pen_lifts_pred [B, T]:
//   score_loss = mean_{b,t} sum_c (eps - score)^2                d/dscore = -2 (eps - score) / (B T)
//   pen_loss   = mean_b( mean_t bce(p, clamp(y)) * alpha_b )     d/dp     = alpha_b / (B T) * (p - y) / max(p (1 - p), 1e-12)
// (torch's binary_cross_entropy clamps both logs at -100 and its backward uses the 1e-12 floor).
// partial: [2][kRedBlocks] doubles (score sums, pen sums).
__global__ void loss_partial_kernel(const float* __restrict__ eps, const float* __restrict__ score, const float* __restrict__ pen,
                                    const float* __restrict__ pen_pred, const float* __restrict__ alphas, int B, int T,
                                    double* __restrict__ partial, float* __restrict__ g_score, float* __restrict__ g_pen) {
  const size_t n = (size_t)B * T;
  const float inv_n = 1.f / (float)n;
  double s_score = 0.0, s_pen = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float2 e = reinterpret_cast<const float2*>(eps)[i], s = reinterpret_cast<const float2*>(score)[i];
    const float d0 = e.x - s.x, d1 = e.y - s.y;
    s_score += (double)(d0 * d0 + d1 * d1);
    const float a = alphas[i / (size_t)T];
    const float y = fminf(fmaxf(pen[i], 1e-7f), 1.f - 1e-7f);
    const float p = pen_pred[i];
    const float bce = -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f));
    s_pen += (double)(bce * a);
    if (g_score) reinterpret_cast<float2*>(g_score)[i] = make_float2(-2.f * d0 * inv_n, -2.f * d1 * inv_n);
    if (g_pen) g_pen[i] = a * inv_n * (p - y) / fmaxf((1.f - p) * p, 1e-12f);
  }
  const double t0 = block_sum(s_score);
  const double t1 = block_sum(s_pen);
  if (threadIdx.x == 0) { partial[blockIdx.x] = t0; partial[kRedBlocks + blockIdx.x] = t1; }
}
__global__ void loss_final_kernel(const double* __restrict__ partial, int B, int T, float* __restrict__ losses) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < kRedBlocks; i += blockDim.x) { a += partial[i]; b += partial[kRedBlocks + i]; }
  a = block_sum(a);
  b = block_sum(b);
  if (threadIdx.x == 0) {
    const double n = (double)B * (double)T;
    const float sc = (float)(a / n), pl = (float)(b / n);
    losses[0] = sc + pl;   // loss.py:39
    losses[1] = sc;
    losses[2] = pl;
  }
}

// sum of squares of the flat gradient (clip_grad_norm_, norm_type 2): fixed partials in double, then one block
__global__ void sqnorm_partial_kernel(const float* __restrict__ g, size_t n, double* __restrict__ partial) {
  double s = 0.0;
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float t = g[(n4 << 2) + threadIdx.x]; s += (double)t * t; }
  const double t = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void sqnorm_final_kernel(const double* __restrict__ partial, double* __restrict__ out) {
  double a = 0.0;
  for (int i = threadIdx.x; i < kRedBlocks; i += blockDim.x) a += partial[i];
  a = block_sum(a);
  if (threadIdx.x == 0) out[0] = a;
}

// torch.optim.Adam (single tensor, amsgrad off, maximize off), same operation order as torch/optim/adam.py
// _single_tensor_adam: g += wd * p; m.lerp_(g, 1 - b1); v = v * b2 + (1 - b2) g g; denom = sqrt(v) / sqrt(bc2) + eps;
// p -= (lr / bc1) * m / denom.  The gradient that enters is g_sum * inv_world * clip, clip = min(1, max_norm /
// (||g_sum|| * inv_world + 1e-6)) (clip_grad_norm_), read from the device-side sum of squares.
struct AdamArgs {
  float lr_over_bc1, bc2_sqrt, w1, beta2, w2, eps, weight_decay, max_norm, inv_world;   // w1 = 1 - beta1, w2 = 1 - beta2 (rounded from double like torch's scalars)
};
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t n, AdamArgs a, const double* __restrict__ sqnorm) {
  float gs = a.inv_world;
  if (sqnorm) {
    const float total = (float)(sqrt(sqnorm[0]) * (double)a.inv_world);
    const float coef = a.max_norm / (total + 1e-6f);
    gs *= fminf(coef, 1.f);
  }
  const float w1 = a.w1, w2 = a.w2;
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * gs;
    if (a.weight_decay != 0.f) gg = gg + a.weight_decay * pp;
    mm = mm + w1 * (gg - mm);
    vv = vv * a.beta2 + w2 * gg * gg;
    const float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
    pp = pp - a.lr_over_bc1 * (mm / denom);
  };
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    one(P.x, G.x, M.x, V.x); one(P.y, G.y, M.y, V.y); one(P.z, G.z, M.z, V.z); one(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = (n4 << 2) + threadIdx.x;
    one(p[i], g[i], m[i], v[i]);
  }
}

int grid_for(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  return (int)(b < 1 ? 1 : b > 148 * 16 ? 148 * 16 : b);
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

const char* dhg_train_last_error(void) { return g_terr; }

int32_t dhg_train_perturb(int32_t device, const float* x, const float* alphas, const float* eps, float* out, int32_t B, int32_t T,
                          void* stream) {
  if (!x || !alphas || !eps || !out || B < 1 || T < 1) return tfail("dhg_train_perturb: bad argument");
  T_OK(cudaSetDevice(device));
  perturb_kernel<<<grid_for((size_t)B * T * 2, 256), 256, 0, (cudaStream_t)stream>>>(x, alphas, eps, out, B, T);
  T_OK(cudaGetLastError());
  return 0;
}

int32_t dhg_train_scratch_doubles(void) { return 2 * kRedBlocks; }

int32_t dhg_train_loss(int32_t device, const float* eps, const float* score_pred, const float* pen_lifts, const float* pen_lifts_pred,
                       const float* alphas, int32_t B, int32_t T, float* dev_losses, float* dev_grad_score, float* dev_grad_pen_pred,
                       double* dev_scratch, void* stream) {
  if (!eps || !score_pred || !pen_lifts || !pen_lifts_pred || !alphas || !dev_losses || !dev_scratch || B < 1 || T < 1)
    return tfail("dhg_train_loss: bad argument");
  if ((reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(score_pred) | reinterpret_cast<uintptr_t>(dev_grad_score)) & 7)
    return tfail("dhg_train_loss: eps, score_pred and grad_score must be 8-byte aligned");
  T_OK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  loss_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(eps, score_pred, pen_lifts, pen_lifts_pred, alphas, B, T, dev_scratch,
                                                         dev_grad_score, dev_grad_pen_pred);
  loss_final_kernel<<<1, kRedThreads, 0, st>>>(dev_scratch, B, T, dev_losses);
  T_OK(cudaGetLastError());
  return 0;
}

int32_t dhg_train_sqnorm(int32_t device, const float* dev_grad, int64_t n, double* dev_out, double* dev_scratch, void* stream) {
  if (!dev_grad || n < 1 || !dev_out || !dev_scratch) return tfail("dhg_train_sqnorm: bad argument");
  if (!aligned16(dev_grad)) return tfail("dhg_train_sqnorm: the gradient buffer must be 16-byte aligned");
  T_OK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  sqnorm_partial_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(dev_grad, (size_t)n, dev_scratch);
  sqnorm_final_kernel<<<1, kRedThreads, 0, st>>>(dev_scratch, dev_out);
  T_OK(cudaGetLastError());
  return 0;
}

int32_t dhg_train_adam_step(int32_t device, float* dev_param, const float* dev_grad, float* dev_exp_avg, float* dev_exp_avg_sq, int64_t n,
                            int32_t step, double lr, double beta1, double beta2, double eps, double weight_decay, const double* dev_sqnorm,
                            double max_norm, int32_t world_size, void* stream) {
  if (!dev_param || !dev_grad || !dev_exp_avg || !dev_exp_avg_sq || n < 1 || step < 1 || world_size < 1)
    return tfail("dhg_train_adam_step: bad argument");
  if (!aligned16(dev_param) || !aligned16(dev_grad) || !aligned16(dev_exp_avg) || !aligned16(dev_exp_avg_sq))
    return tfail("dhg_train_adam_step: the flat buffers must be 16-byte aligned");
  if (dev_sqnorm && !(max_norm > 0.0)) return tfail("dhg_train_adam_step: max_norm must be positive when clipping");
  T_OK(cudaSetDevice(device));
  // bias corrections like torch (python floats = doubles), rounded once
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  AdamArgs a;
  a.lr_over_bc1 = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.w1 = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.w2 = (float)(1.0 - beta2); a.eps = (float)eps; a.weight_decay = (float)weight_decay;
  a.max_norm = (float)max_norm; a.inv_world = 1.0f / (float)world_size;
  adam_kernel<<<grid_for((size_t)n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(dev_param, dev_grad, dev_exp_avg, dev_exp_avg_sq, (size_t)n,
                                                                                a, dev_sqnorm);
  T_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
