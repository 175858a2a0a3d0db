// StyleExtractor on the GPU: the step right before the sampling path.
//
// Reference: diffusion_handwriting_generation/text_style.py:11-59 -- a torchvision MobileNetV2 feature stack
// (eval mode, BatchNorm with running statistics) on the grey writer image repeated to 3 channels, AvgPool2d(3, 3),
// AdaptiveAvgPool2d((1, 14)), -> [B, 14, 1280].  The network itself is torchvision's (not in /root/reference):
// torchvision/models/mobilenetv2.py, inverted residual setting t,c,n,s = (1,16,1,1) (6,24,2,2) (6,32,3,2) (6,64,4,2)
// (6,96,3,1) (6,160,3,2) (6,320,1,1), first conv 3->32 stride 2, last conv 320->1280, ReLU6, BN eps 1e-5.
//
// This runs once per prompt, before the 60-step chain, on a 96-pixel-high image: it is a few hundred microseconds of
// work, so the kernels are plain fp32 CUDA-core code (exactness against torchvision matters here, throughput does not):
//   * BatchNorm is folded into each convolution at load time (w' = w * gamma / sqrt(var + eps), b' = beta - mean * ...);
//   * the 3 identical input channels are folded into one (w'[o] = sum_c w[o, c]);
//   * activations are channels-last [B, H, W, C] (C padded to a multiple of 16), so every 1x1 convolution is a row-major
//     GEMM over pixels (the fp32 GEMM of kernels_simt.cu) with a fused bias / ReLU6 / residual pass, the depthwise 3x3
//     convolutions are a 9-tap stencil per channel, and the two poolings are one kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/dhg_b200.h"
#include "kernels.h"

using namespace dhg;

namespace {

thread_local char g_serr[512] = "";
int sfail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_serr, sizeof(g_serr), fmt, ap);
  va_end(ap);
  return 1;
}
#define S_OK(call)                                                                           \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) return sfail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

constexpr int kFeat = 1280, kStyleTokens = 14;
inline int pad16(int c) { return (c + 15) & ~15; }

// features[0]: 3x3 stride-2 conv on the (folded) single input channel + bias + ReLU6.  img: [B, H, W] grey 0..255,
// normalised here like text_style.py:51 (x / 127.5 - 1); zero padding applies to the normalised image.
__global__ void first_conv_kernel(const float* __restrict__ img, const float* __restrict__ w /*[9][32]*/, const float* __restrict__ b,
                                  float* __restrict__ out, int B, int H, int W, int Ho, int Wo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Ho * Wo * 32;
  if (i >= total) return;
  const int o = (int)(i & 31);
  size_t pix = i >> 5;
  const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
  float acc = b[o];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = 2 * y - 1 + ky, xx = 2 * x - 1 + kx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float v = img[((size_t)n * H + yy) * W + xx] / 127.5f - 1.0f;
        acc = fmaf(v, w[(ky * 3 + kx) * 32 + o], acc);
      }
    }
  out[i] = fminf(fmaxf(acc, 0.f), 6.f);
}

// Depthwise 3x3 (stride 1 or 2, pad 1) + bias + ReLU6, channels-last.  4 channels per thread.
__global__ void depthwise_kernel(const float* __restrict__ in, const float* __restrict__ w /*[9][Cp]*/, const float* __restrict__ b,
                                 float* __restrict__ out, int B, int H, int W, int Ho, int Wo, int Cp, int stride) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = Cp >> 2;
  const size_t total = (size_t)B * Ho * Wo * c4;
  if (i >= total) return;
  const int c = (int)(i % c4) * 4;
  size_t pix = i / c4;
  const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
  float4 acc = *reinterpret_cast<const float4*>(b + c);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = stride * y - 1 + ky, xx = stride * x - 1 + kx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + yy) * W + xx) * Cp + c);
        const float4 k = *reinterpret_cast<const float4*>(w + (size_t)(ky * 3 + kx) * Cp + c);
        acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
        acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
      }
    }
  acc.x = fminf(fmaxf(acc.x, 0.f), 6.f); acc.y = fminf(fmaxf(acc.y, 0.f), 6.f);
  acc.z = fminf(fmaxf(acc.z, 0.f), 6.f); acc.w = fminf(fmaxf(acc.w, 0.f), 6.f);
  *reinterpret_cast<float4*>(out + pix * Cp + c) = acc;
}

// After the pointwise GEMM: + bias, optional ReLU6, optional residual (the block's input), written with the padded
// channel pitch (padding channels are written as zeros so that they never contribute to the next GEMM).
__global__ void pointwise_post_kernel(const float* __restrict__ acc /*[rows, N]*/, const float* __restrict__ b, const float* __restrict__ res,
                                      float* __restrict__ out /*[rows, Np]*/, size_t rows, int N, int Np, int relu6) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (size_t)Np) return;
  const int n = (int)(i % Np);
  const size_t r = i / Np;
  float v = 0.f;
  if (n < N) {
    v = acc[r * N + n] + b[n];
    if (relu6) v = fminf(fmaxf(v, 0.f), 6.f);
    if (res) v += res[i];
  }
  out[i] = v;
}

// AvgPool2d(3, 3) then AdaptiveAvgPool2d((1, 14)) (text_style.py:55-56), then [B, 1280, 14] -> [B, 14, 1280].
__global__ void style_pool_kernel(const float* __restrict__ f /*[B, Hf, Wf, 1280]*/, float* __restrict__ out /*[B, 14, 1280]*/, int B, int Hf, int Wf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * kStyleTokens * kFeat) return;
  const int c = (int)(i % kFeat), j = (int)((i / kFeat) % kStyleTokens), n = (int)(i / ((size_t)kFeat * kStyleTokens));
  const int Hp = Hf / 3, Wp = Wf / 3;
  const int x0 = (j * Wp) / kStyleTokens, x1 = ((j + 1) * Wp + kStyleTokens - 1) / kStyleTokens;   // adaptive bin [floor, ceil)
  float s = 0.f;
  for (int py = 0; py < Hp; ++py)
    for (int px = x0; px < x1; ++px) {
      float a = 0.f;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) a += f[(((size_t)n * Hf + 3 * py + dy) * Wf + 3 * px + dx) * kFeat + c];
      s += a / 9.0f;
    }
  out[i] = s / (float)(Hp * (x1 - x0));
}

struct Conv {          // one convolution with its BatchNorm folded in
  int kind;            // 0 first 3x3 (1 folded input channel), 1 pointwise, 2 depthwise 3x3
  int cin, cout, stride;
  bool relu6, residual;
  std::string conv_key, bn_key;
  float *w = nullptr, *b = nullptr;   // device: kind 0 [9][32]; kind 1 [Kp][N]; kind 2 [9][Cp]
};

}  // namespace

struct dhg_style {
  int device = 0;
  std::map<std::string, std::vector<float>> raw;
  std::map<std::string, std::vector<int64_t>> shapes;
  std::vector<Conv> convs;
  std::vector<int> block_first;   // index of the first conv of every inverted-residual block (for the residual source)
  bool finalized = false;
  std::vector<void*> allocs;
  float *buf[3] = {nullptr, nullptr, nullptr}, *scratch = nullptr, *img = nullptr;
  size_t cap_act = 0, cap_scratch = 0, cap_img = 0;
};

namespace {

void build_convs(dhg_style* s) {
  auto add = [&](int kind, int cin, int cout, int stride, bool relu6, bool residual, const std::string& ck, const std::string& bk) {
    Conv c;
    c.kind = kind; c.cin = cin; c.cout = cout; c.stride = stride; c.relu6 = relu6; c.residual = residual; c.conv_key = ck; c.bn_key = bk;
    s->convs.push_back(c);
  };
  add(0, 3, 32, 2, true, false, "features.0.0", "features.0.1");
  static const int cfg[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2}, {6, 32, 3, 2}, {6, 64, 4, 2}, {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};
  int in = 32, idx = 1;
  for (auto& g : cfg)
    for (int r = 0; r < g[2]; ++r, ++idx) {
      const int t = g[0], out = g[1], stride = r == 0 ? g[3] : 1, hidden = in * t;
      const std::string p = "features." + std::to_string(idx) + ".conv.";
      const bool res = stride == 1 && in == out;
      s->block_first.push_back((int)s->convs.size());
      int k = 0;
      if (t != 1) { add(1, in, hidden, 1, true, false, p + "0.0", p + "0.1"); k = 1; }
      add(2, hidden, hidden, stride, true, false, p + std::to_string(k) + ".0", p + std::to_string(k) + ".1");
      add(1, hidden, out, 1, false, res, p + std::to_string(k + 1), p + std::to_string(k + 2));
      in = out;
    }
  s->block_first.push_back((int)s->convs.size());
  add(1, 320, kFeat, 1, true, false, "features.18.0", "features.18.1");
}

int upload(dhg_style* s, float** d, const std::vector<float>& h) {
  S_OK(cudaMalloc((void**)d, h.size() * sizeof(float)));
  s->allocs.push_back(*d);
  S_OK(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

extern "C" {

const char* dhg_style_last_error(void) { return g_serr; }

int32_t dhg_style_create(int32_t device, dhg_style** out) {
  if (!out) return sfail("dhg_style_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return sfail("dhg_style_create: no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return sfail("dhg_style_create: device %d out of range", device);
  dhg_style* s = new dhg_style();
  s->device = device;
  build_convs(s);
  *out = s;
  return 0;
}

int32_t dhg_style_destroy(dhg_style* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  for (void* p : s->allocs) cudaFree(p);
  for (float* p : {s->buf[0], s->buf[1], s->buf[2], s->scratch, s->img})
    if (p) cudaFree(p);
  delete s;
  return 0;
}

int32_t dhg_style_load_weight(dhg_style* s, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  if (!s || !name || !data || (!shape && ndim > 0)) return sfail("dhg_style_load_weight: null argument");
  if (s->finalized) return sfail("dhg_style_load_weight: already finalized");
  int64_t n = 1;
  std::vector<int64_t> sh;
  for (int i = 0; i < ndim; ++i) { n *= shape[i]; sh.push_back(shape[i]); }
  s->raw[name].assign(data, data + n);
  s->shapes[name] = sh;
  return 0;
}

int32_t dhg_style_finalize(dhg_style* s) {
  if (!s) return sfail("null style extractor");
  if (s->finalized) return 0;
  S_OK(cudaSetDevice(s->device));
  for (Conv& c : s->convs) {
    const std::string wk = c.conv_key + ".weight";
    for (const std::string& k : {wk, c.bn_key + ".weight", c.bn_key + ".bias", c.bn_key + ".running_mean", c.bn_key + ".running_var"})
      if (!s->raw.count(k)) return sfail("missing key in MobileNetV2 state_dict: %s", k.c_str());
    const std::vector<float>&W = s->raw[wk], &g = s->raw[c.bn_key + ".weight"], &be = s->raw[c.bn_key + ".bias"],
                            &mu = s->raw[c.bn_key + ".running_mean"], &var = s->raw[c.bn_key + ".running_var"];
    const int per_out = c.kind == 0 ? 27 : c.kind == 1 ? c.cin : 9;
    if ((int)g.size() != c.cout || (int64_t)W.size() != (int64_t)c.cout * per_out) return sfail("size mismatch for %s", wk.c_str());
    std::vector<float> scale(c.cout), bias(c.cout);
    for (int o = 0; o < c.cout; ++o) {
      scale[o] = g[o] / sqrtf(var[o] + 1e-5f);
      bias[o] = be[o] - mu[o] * scale[o];
    }
    std::vector<float> w;
    if (c.kind == 0) {          // [32][3][3][3] -> [9][32], the 3 (identical) input channels summed
      w.assign(9 * 32, 0.f);
      for (int o = 0; o < 32; ++o)
        for (int k = 0; k < 9; ++k) {
          double a = 0.0;
          for (int ci = 0; ci < 3; ++ci) a += W[((size_t)o * 3 + ci) * 9 + k];
          w[(size_t)k * 32 + o] = (float)a * scale[o];
        }
    } else if (c.kind == 1) {   // [N][K] -> [Kp][N] (K padded with zero rows)
      const int Kp = pad16(c.cin);
      w.assign((size_t)Kp * c.cout, 0.f);
      for (int o = 0; o < c.cout; ++o)
        for (int k = 0; k < c.cin; ++k) w[(size_t)k * c.cout + o] = W[(size_t)o * c.cin + k] * scale[o];
    } else {                    // [C][1][3][3] -> [9][Cp]
      const int Cp = pad16(c.cout);
      w.assign((size_t)9 * Cp, 0.f);
      for (int o = 0; o < c.cout; ++o)
        for (int k = 0; k < 9; ++k) w[(size_t)k * Cp + o] = W[(size_t)o * 9 + k] * scale[o];
      bias.resize(Cp, 0.f);
    }
    if (upload(s, &c.w, w) || upload(s, &c.b, bias)) return 1;
  }
  s->raw.clear();
  s->finalized = true;
  return 0;
}

int32_t dhg_style_extract(dhg_style* s, const float* host_img, int32_t B, int32_t H, int32_t W, float* dev_out, void* stream) {
  if (!s || !host_img || !dev_out) return sfail("dhg_style_extract: null argument");
  if (!s->finalized) return sfail("dhg_style_extract: dhg_style_finalize has not been called");
  if (B < 1 || H < 65 || W < 65) return sfail("dhg_style_extract: the image must be at least 65 x 65 (a 3 x 3 feature map after the /32 stack); got %d x %d", H, W);
  S_OK(cudaSetDevice(s->device));
  cudaStream_t st = (cudaStream_t)stream;
  // buffer sizes: the largest activation is the expanded 96-channel map at half resolution
  auto out_dim = [](int n, int stride) { return (n + 2 - 3) / stride + 1; };
  size_t max_act = 0, max_scr = 0;
  {
    int h = out_dim(H, 2), w = out_dim(W, 2);
    max_act = (size_t)B * h * w * 32;
    for (size_t i = 1; i < s->convs.size(); ++i) {
      const Conv& c = s->convs[i];
      if (c.kind == 2) { h = out_dim(h, c.stride); w = out_dim(w, c.stride); }
      const size_t px = (size_t)B * h * w;
      max_act = std::max(max_act, px * (size_t)pad16(c.cout));
      if (c.kind == 1) max_scr = std::max(max_scr, px * (size_t)c.cout);
    }
  }
  if (max_act > s->cap_act || max_scr > s->cap_scratch || (size_t)B * H * W > s->cap_img) {
    S_OK(cudaDeviceSynchronize());
    for (float** p : {&s->buf[0], &s->buf[1], &s->buf[2], &s->scratch, &s->img}) { if (*p) cudaFree(*p); *p = nullptr; }
    for (int i = 0; i < 3; ++i) S_OK(cudaMalloc((void**)&s->buf[i], max_act * sizeof(float)));
    S_OK(cudaMalloc((void**)&s->scratch, max_scr * sizeof(float)));
    S_OK(cudaMalloc((void**)&s->img, (size_t)B * H * W * sizeof(float)));
    s->cap_act = max_act; s->cap_scratch = max_scr; s->cap_img = (size_t)B * H * W;
  }
  S_OK(cudaMemcpyAsync(s->img, host_img, (size_t)B * H * W * sizeof(float), cudaMemcpyHostToDevice, st));
  int h = out_dim(H, 2), w = out_dim(W, 2);
  {
    const size_t total = (size_t)B * h * w * 32;
    first_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s->img, s->convs[0].w, s->convs[0].b, s->buf[0], B, H, W, h, w);
  }
  // cur: the block's running activation; the block input is kept for the residual (buffers rotate)
  int cur = 0, block_in = 0, next_block = 0;
  for (size_t i = 1; i < s->convs.size(); ++i) {
    const Conv& c = s->convs[i];
    if (next_block < (int)s->block_first.size() && (int)i == s->block_first[next_block]) { block_in = cur; ++next_block; }
    int dst = 0;
    while (dst == cur || dst == block_in) ++dst;   // a free buffer (3 buffers: current, block input, destination)
    if (c.kind == 2) {
      const int Cp = pad16(c.cout), ho = out_dim(h, c.stride), wo = out_dim(w, c.stride);
      const size_t total = (size_t)B * ho * wo * (Cp / 4);
      depthwise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s->buf[cur], c.w, c.b, s->buf[dst], B, h, w, ho, wo, Cp, c.stride);
      h = ho; w = wo;
    } else {
      const size_t rows = (size_t)B * h * w;
      const int Kp = pad16(c.cin), Np = pad16(c.cout);
      launch_gemm_simt<float>(s->buf[cur], Kp, (int)rows, c.w, Kp, c.cout, 1, s->scratch, st);
      const size_t total = rows * (size_t)Np;
      pointwise_post_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s->scratch, c.b, c.residual ? s->buf[block_in] : nullptr, s->buf[dst],
                                                                            rows, c.cout, Np, c.relu6 ? 1 : 0);
    }
    cur = dst;
  }
  if (h < 3 || w / 3 < 1) return sfail("dhg_style_extract: feature map %d x %d too small for AvgPool2d(3, 3)", h, w);
  const size_t total = (size_t)B * kStyleTokens * kFeat;
  style_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s->buf[cur], dev_out, B, h, w);
  S_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
