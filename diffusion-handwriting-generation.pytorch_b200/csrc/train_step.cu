// Forward + backward pass of the denoiser for the reference's training step (BASELINE configs[3]):
//
//   train.py:46-55   strokes_pred, pen_lifts_pred, _ = model(x_perturbed, text, sqrt(alphas), style); loss.backward()
//
// i.e. DiffusionModel.forward in training form (model.py:121-182 with every activation kept) and the gradient of the
// loss with respect to all 323 parameter tensors, written into ONE flat fp32 buffer in checkpoint key order -- the
// buffer dhg_train_sqnorm / dhg_train_adam_step (train_update.cu) and the data-parallel all-reduce work on.  Together
// with dhg_train_perturb / dhg_train_loss this is the whole of TrainingLoop.train_step on the device.
//
// Arithmetic is fp32 on the CUDA cores, like the reference's training (fp32 parameters, fp32 autograd): every
// contraction -- Linear, the three taps of a k3 Conv1d, Q K^T, P V, and all their data / weight gradients -- is ONE
// strided batched GEMM (`Bmm`: C[z] (+)= alpha sum_tap A[z](i + shift_tap, :) B_tap[z] (+ bias), arbitrary element
// strides, so transposes, head splitting and the row shift of a conv tap are pointer arithmetic and never a copy);
// weight gradients are cut into batch items along the row axis and reduced with fp32 vector atomics.  On the GPU the
// GEMM is a shared-memory tiled kernel (64..128 x 64..128 tiles, register prefetch of the next k slab, 16-byte
// operand loads and stores where the strides allow); its products can also run on the tensor cores as 3 x TF32 or
// plain TF32 (mma.sync; dhg_trainer_set_option "tiled_gemm" 3 / 4 -- measured in DESIGN.md 4.10, not the default).
// Everything else (SiLU, FiLM, LayerNorm, softmax + padding mask, pooling, upsampling, embedding, PE add) is a
// per-element or warp-per-row kernel.  The plan is a tape built once for (B, T, L): forward = the tape, backward = zero
// the gradient arena, then the tape in reverse; weight / bias gradients and the forward's skip branches run on a
// second stream (event fork / join).  Everything is stream-ordered; nothing synchronises with the host; the caller
// may capture forward + loss + backward + update in a CUDA graph.
//
// What this is NOT: a tcgen05 path (the sampling path's tcgen05 GEMMs are bf16 / split-bf16 forward kernels with
// fused epilogues; their backward twins are not written).  DESIGN.md 4.10 has the measured step time beside the
// reference's eager step on the same GPU.
//
// The same source also builds as plain C++ (-DDHG_HOSTSIM, g++ -fopenmp): every kernel body is a functor over a flat
// index, so the host build runs the identical bodies in a loop.  That build exists ONLY for tests/ (gradient check
// against torch autograd without a GPU, tests/test_train_step_hostsim.py); it is never part of libdhg_b200.so and the
// product has no CPU path.
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#ifndef DHG_HOSTSIM
#include <cuda_runtime.h>
#define TS_FN __device__ __forceinline__
typedef cudaStream_t ts_stream;
#else
#define TS_FN inline
typedef void* ts_stream;
#endif

#include "../../include/dhg_b200.h"

namespace {

thread_local char g_serr[512] = "";
int sfail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_serr, sizeof(g_serr), fmt, ap);
  va_end(ap);
  return 1;
}

// GPU build, dhg_trainer_set_option("tiled_gemm", v):
//   1 (default) shared-memory tiled GEMM, fp32 FMAs on the CUDA cores; warp-per-row LayerNorm / softmax
//   2 the same with the smallest tile only (measurement)
//   3 tiled GEMM with 3 x TF32 tensor-core products (mma.sync; same fp32 contract; measured: no faster)
//   4 tiled GEMM with plain TF32 products (one mma per product: what torch's allow_tf32 does; ~1e-3 relative, NOT the fp32 contract)
//   0 the per-thread bodies the host build runs (GEMM and per-row kernels)
int g_use_tiled = 1;
int g_side_stream = 1;   // "side_stream": weight / bias gradients on a second stream beside the data-gradient chain

// ------------------------------------------------------------------------------------------------------------------
// launch layer
// ------------------------------------------------------------------------------------------------------------------
#ifndef DHG_HOSTSIM
template <class F>
__global__ void __launch_bounds__(256) ts_kernel(long n, F f) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) f(i);
}
// one warp per row (LayerNorm, softmax): f.warp(row, lane) does the same arithmetic as f(row) with lane-strided columns
template <class F>
__global__ void __launch_bounds__(256) ts_row_kernel(long rows, F f) {
  const int lane = threadIdx.x & 31;
  for (long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long)gridDim.x * 8) f.warp(r, lane);
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
TS_FN void ts_atomic_add(float* p, float v) { atomicAdd(p, v); }
#else
TS_FN void ts_atomic_add(float* p, float v) {
#pragma omp atomic
  *p += v;
}
#endif

struct Launcher {
  ts_stream st = nullptr;
  long launches = 0;
  int grid_cap = 148 * 16;
  // Weight / bias gradients depend only on what the backward has already produced and nothing downstream reads them:
  // they run on a second stream beside the data-gradient chain (most launches here are a single partial wave).
  // side_begin(): the side stream waits for everything enqueued so far and becomes the launch stream; side_end():
  // back to the caller's stream; side_join(): the caller's stream waits for the side stream (end of the backward).
#ifndef DHG_HOSTSIM
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t>* events = nullptr;
  size_t next_event = 0;
  bool forked = false;
  bool side_ok() const { return side && events && next_event < events->size(); }
  void side_begin() {
    if (!side_ok()) return;
    cudaEvent_t e = (*events)[next_event++];
    cudaEventRecord(e, st);
    cudaStreamWaitEvent(side, e, 0);
    std::swap(st, side);
    forked = true;
    in_side = true;
  }
  void side_end() {
    if (!in_side) return;
    std::swap(st, side);
    in_side = false;
  }
  void side_join() {
    if (!forked || !side_ok()) return;
    cudaEvent_t e = (*events)[next_event++];
    cudaEventRecord(e, side);
    cudaStreamWaitEvent(st, e, 0);
    forked = false;
  }
  bool in_side = false;
#else
  void side_begin() {}
  void side_end() {}
  void side_join() {}
#endif
  template <class F>
  void run(long n, const F& f) {
    if (n <= 0) return;
#ifndef DHG_HOSTSIM
    long blocks = (n + 255) / 256;
    if (blocks > grid_cap) blocks = grid_cap;
    ts_kernel<F><<<(unsigned)blocks, 256, 0, st>>>(n, f);
#else
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) f(i);
#endif
    ++launches;
  }
  // per-row kernels: a warp per row on the GPU (the per-thread body with "tiled_gemm" 0 and in the host build)
  template <class F>
  void run_rows(long rows, const F& f) {
    if (rows <= 0) return;
#ifndef DHG_HOSTSIM
    if (g_use_tiled) {
      long blocks = (rows + 7) / 8;
      if (blocks > grid_cap) blocks = grid_cap;
      ts_row_kernel<F><<<(unsigned)blocks, 256, 0, st>>>(rows, f);
      ++launches;
      return;
    }
#endif
    run(rows, f);
  }
  void zero(void* p, size_t bytes) {
    if (!bytes) return;
#ifndef DHG_HOSTSIM
    cudaMemsetAsync(p, 0, bytes, st);
#else
    memset(p, 0, bytes);
#endif
  }
  void copy(void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
#ifndef DHG_HOSTSIM
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
#else
    memcpy(dst, src, bytes);
#endif
  }
};

void* dev_alloc(size_t bytes) {
  if (!bytes) bytes = 256;
#ifndef DHG_HOSTSIM
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  return p;
#else
  return aligned_alloc(256, (bytes + 255) / 256 * 256);
#endif
}
void dev_free(void* p) {
  if (!p) return;
#ifndef DHG_HOSTSIM
  cudaFree(p);
#else
  free(p);
#endif
}
void to_dev(void* dst, const void* src, size_t bytes) {
#ifndef DHG_HOSTSIM
  cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
#else
  memcpy(dst, src, bytes);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// the one contraction:  C[z](i, j) (+)= alpha * sum_tap sum_k A[z](i + shift_tap, k) B_tap[z](k, j) (+ bias[j]),
// z = z1 * Z2 + z2.  taps = 1 for everything but a k3 convolution, where tap t reads row i + shift0 + t * dshift of the
// same batch item (= sample) and rows outside [0, M) are the zero padding (cnn.py:32-47: padding 1).
// ------------------------------------------------------------------------------------------------------------------
struct Bmm {
  const float* A = nullptr;
  const float* B = nullptr;
  float* C = nullptr;
  const float* bias = nullptr;   // only with a plain store
  int M = 0, N = 0, K = 0, Z1 = 1, Z2 = 1;
  long sAz1 = 0, sAz2 = 0, sAi = 0, sAk = 0;
  long sBz1 = 0, sBz2 = 0, sBk = 0, sBj = 0;
  long sCz1 = 0, sCz2 = 0, sCi = 0, sCj = 0;
  int taps = 1, shift0 = 0, dshift = 0;
  long sBtap = 0;
  float alpha = 1.f;
  int mode = 0;   // 0: C = v   1: C += v (one writer per element)   2: atomicAdd(C, v) (batch items share C)
};

TS_FN void bmm_store(const Bmm& p, float* C, int i, int j, float acc) {
  float v = p.alpha * acc;
  float* c = C + (long)i * p.sCi + (long)j * p.sCj;
  if (p.mode == 0) *c = p.bias ? v + p.bias[j] : v;
  else if (p.mode == 1) *c += v;
  else ts_atomic_add(c, v);
}

// one thread = one 4 x 4 block of C, operands straight from memory (the body the host build runs; on the GPU the
// reference the tiled kernel is checked against)
struct BmmBody {
  Bmm p;
  int tm, tn;
  TS_FN void operator()(long gid) const {
    const int tj = (int)(gid % tn);
    long r = gid / tn;
    const int ti = (int)(r % tm);
    const int z = (int)(r / tm);
    const int z1 = z / p.Z2, z2 = z % p.Z2;
    const float* A = p.A + z1 * p.sAz1 + z2 * p.sAz2;
    const float* B = p.B + z1 * p.sBz1 + z2 * p.sBz2;
    float* C = p.C + z1 * p.sCz1 + z2 * p.sCz2;
    const int i0 = ti * 4, j0 = tj * 4;
    float acc[4][4];
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int t = 0; t < p.taps; ++t) {
      const int sh = p.shift0 + t * p.dshift;
      const float* Bt = B + t * p.sBtap;
      for (int k = 0; k < p.K; ++k) {
        float av[4], bv[4];
        for (int a = 0; a < 4; ++a) {
          const int row = i0 + a + sh;
          av[a] = (i0 + a < p.M && row >= 0 && row < p.M) ? A[(long)row * p.sAi + (long)k * p.sAk] : 0.f;
        }
        for (int b = 0; b < 4; ++b) bv[b] = (j0 + b < p.N) ? Bt[(long)k * p.sBk + (long)(j0 + b) * p.sBj] : 0.f;
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
    }
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        if (i0 + a < p.M && j0 + b < p.N) bmm_store(p, C, i0 + a, j0 + b, acc[a][b]);
  }
};

#ifndef DHG_HOSTSIM
// BM x BN tile of C per CTA of 256 threads, TM x TN per thread (in 4-wide groups BM/2 resp. BN/2 apart, so that the
// shared-memory reads of a warp are two broadcasts and one conflict-free 16-lane run), 16-deep k slabs: the next slab
// travels from memory into registers while the current one is multiplied.  The slab loads walk whichever of the two
// operand axes is contiguous -- as 16-byte vectors when the strides and the base allow it -- and the tile is stored /
// accumulated / atomically added as 16-byte vectors when C's rows are contiguous (red.global.add.v4.f32 for the
// weight gradients).
constexpr int kBK = 16;
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ bool al4(long v) { return (v & 3) == 0; }
__device__ __forceinline__ bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }
// TC = true: the products run on the tensor cores as 3 x TF32 (mma.sync m16n8k8: a = a_hi + a_lo, b = b_hi + b_lo in TF32,
// a_lo b_hi + a_hi b_lo + a_hi b_hi accumulated in fp32 -- the dropped a_lo b_lo term is 2^-22 relative), which keeps the
// fp32 contract of the training step; (TM, TN) is then the warp grid (TM x TN = 8 warps, each (BM / TM) x (BN / TN)).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
template <int BM, int BN, int TM, int TN, int TCT>   // TCT: 0 CUDA cores, 3 / 1 tensor cores with that many TF32 products per fp32 product
__global__ void __launch_bounds__(256, (BM * BN <= 4096) ? 4 : 2) ts_bmm_tiled(Bmm p) {
  constexpr bool TC = TCT != 0;
  static_assert(TC ? (TM * TN == 8) : ((BM / TM) * (BN / TN) == 256), "256 threads");
  constexpr int NX = TC ? 1 : BN / TN, LA = BM * kBK / 256, LB = BN * kBK / 256, GA = TC ? 1 : TM / 4, GB = TC ? 1 : TN / 4;
  constexpr int PAD = TC ? 8 : 4;   // TC: a fragment's (k, row) pairs fall on 32 different banks with a pitch of 8 mod 32
  constexpr int WM = TC ? BM / TM : 16, WN = TC ? BN / TN : 8, MT = WM / 16, NT = WN / 8;
  __shared__ __align__(16) float As[kBK][BM + PAD];
  __shared__ __align__(16) float Bs[kBK][BN + PAD];
  const int z = blockIdx.z, z1 = z / p.Z2, z2 = z % p.Z2;
  const float* __restrict__ A = p.A + z1 * p.sAz1 + z2 * p.sAz2;
  const float* __restrict__ B = p.B + z1 * p.sBz1 + z2 * p.sBz2;
  float* C = p.C + z1 * p.sCz1 + z2 * p.sCz2;
  const int bi = blockIdx.y * BM, bj = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid % NX, ty = tid / NX;
  // operand walk: 0 scalar along i / j, 1 scalar along k, 2 vector along k, 3 vector along i / j
  int am = p.sAk == 1 ? 1 : 0, bm = (p.sBk == 1 && p.sBj != 1) ? 1 : 0;
  if (am == 1 && al4(p.K) && al4(p.sAi) && al16(A)) am = 2;
  if (am == 0 && p.sAi == 1 && p.taps == 1 && p.shift0 == 0 && al4(p.M) && al4(p.sAk) && al16(A)) am = 3;
  if (bm == 1 && al4(p.K) && al4(p.sBj) && al4(p.sBtap) && al16(B)) bm = 2;
  if (bm == 0 && p.sBj == 1 && al4(p.N) && al4(p.sBk) && al4(p.sBtap) && al16(B)) bm = 3;
  const int nk = (p.K + kBK - 1) / kBK, ns = nk * p.taps;
  float acc[TC ? 1 : TM][TC ? 1 : TN];
  float accm[TC ? MT : 1][TC ? NT : 1][4];
#pragma unroll
  for (int a = 0; a < (TC ? 1 : TM); ++a)
#pragma unroll
    for (int b = 0; b < (TC ? 1 : TN); ++b) acc[a][b] = 0.f;
#pragma unroll
  for (int a = 0; a < (TC ? MT : 1); ++a)
#pragma unroll
    for (int b = 0; b < (TC ? NT : 1); ++b) accm[a][b][0] = accm[a][b][1] = accm[a][b][2] = accm[a][b][3] = 0.f;
  const int lane = tid & 31, wid = tid >> 5, gq = lane >> 2, tq = lane & 3;
  const int wm0 = TC ? (wid / TN) * WM : 0, wn0 = TC ? (wid % TN) * WN : 0;
  float ra[LA], rb[LB];
  auto fetch = [&](int s) {
    const int tap = s / nk, k0 = (s - tap * nk) * kBK, sh = p.shift0 + tap * p.dshift;
    const float* __restrict__ Bt = B + tap * p.sBtap;
    if (am >= 2) {
#pragma unroll
      for (int r = 0; r < LA / 4; ++r) {
        const int e = tid + 256 * r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (am == 2) {
          const int gi = bi + (e >> 2), gk = k0 + 4 * (e & 3), row = gi + sh;
          if (gi < p.M && gk < p.K && row >= 0 && row < p.M) v = *reinterpret_cast<const float4*>(A + (long)row * p.sAi + gk);
        } else {
          const int gi = bi + 4 * (e % (BM / 4)), gk = k0 + e / (BM / 4);
          if (gi < p.M && gk < p.K) v = *reinterpret_cast<const float4*>(A + gi + (long)gk * p.sAk);
        }
        ra[4 * r] = v.x; ra[4 * r + 1] = v.y; ra[4 * r + 2] = v.z; ra[4 * r + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < LA; ++r) {
        const int e = tid + 256 * r;
        int kk, ii;
        if (am == 1) { kk = e & 15; ii = e >> 4; } else { ii = e % BM; kk = e / BM; }
        const int gi = bi + ii, gk = k0 + kk, row = gi + sh;
        ra[r] = (gi < p.M && gk < p.K && row >= 0 && row < p.M) ? A[(long)row * p.sAi + (long)gk * p.sAk] : 0.f;
      }
    }
    if (bm >= 2) {
#pragma unroll
      for (int r = 0; r < LB / 4; ++r) {
        const int e = tid + 256 * r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bm == 2) {
          const int gj = bj + (e >> 2), gk = k0 + 4 * (e & 3);
          if (gj < p.N && gk < p.K) v = *reinterpret_cast<const float4*>(Bt + gk + (long)gj * p.sBj);
        } else {
          const int gj = bj + 4 * (e % (BN / 4)), gk = k0 + e / (BN / 4);
          if (gj < p.N && gk < p.K) v = *reinterpret_cast<const float4*>(Bt + (long)gk * p.sBk + gj);
        }
        rb[4 * r] = v.x; rb[4 * r + 1] = v.y; rb[4 * r + 2] = v.z; rb[4 * r + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < LB; ++r) {
        const int e = tid + 256 * r;
        int kb, jj;
        if (bm == 1) { kb = e & 15; jj = e >> 4; } else { jj = e % BN; kb = e / BN; }
        const int gj = bj + jj, gkb = k0 + kb;
        rb[r] = (gj < p.N && gkb < p.K) ? Bt[(long)gkb * p.sBk + (long)gj * p.sBj] : 0.f;
      }
    }
  };
  auto stash = [&]() {
    if (am == 2) {
#pragma unroll
      for (int r = 0; r < LA / 4; ++r) {
        const int e = tid + 256 * r;
#pragma unroll
        for (int q = 0; q < 4; ++q) As[4 * (e & 3) + q][e >> 2] = ra[4 * r + q];
      }
    } else if (am == 3) {
#pragma unroll
      for (int r = 0; r < LA / 4; ++r) {
        const int e = tid + 256 * r;
        *reinterpret_cast<float4*>(&As[e / (BM / 4)][4 * (e % (BM / 4))]) = make_float4(ra[4 * r], ra[4 * r + 1], ra[4 * r + 2], ra[4 * r + 3]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < LA; ++r) {
        const int e = tid + 256 * r;
        if (am == 1) As[e & 15][e >> 4] = ra[r]; else As[e / BM][e % BM] = ra[r];
      }
    }
    if (bm == 2) {
#pragma unroll
      for (int r = 0; r < LB / 4; ++r) {
        const int e = tid + 256 * r;
#pragma unroll
        for (int q = 0; q < 4; ++q) Bs[4 * (e & 3) + q][e >> 2] = rb[4 * r + q];
      }
    } else if (bm == 3) {
#pragma unroll
      for (int r = 0; r < LB / 4; ++r) {
        const int e = tid + 256 * r;
        *reinterpret_cast<float4*>(&Bs[e / (BN / 4)][4 * (e % (BN / 4))]) = make_float4(rb[4 * r], rb[4 * r + 1], rb[4 * r + 2], rb[4 * r + 3]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < LB; ++r) {
        const int e = tid + 256 * r;
        if (bm == 1) Bs[e & 15][e >> 4] = rb[r]; else Bs[e / BN][e % BN] = rb[r];
      }
    }
  };
  if (ns > 0) fetch(0);
  for (int s = 0; s < ns; ++s) {
    stash();
    __syncthreads();
    if (s + 1 < ns) fetch(s + 1);
    if constexpr (TC) {
#pragma unroll
      for (int kb = 0; kb < kBK; kb += 8) {
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          split_tf32(Bs[kb + tq][wn0 + nt * 8 + gq], bh[nt][0], bl[nt][0]);
          split_tf32(Bs[kb + tq + 4][wn0 + nt * 8 + gq], bh[nt][1], bl[nt][1]);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t ah[4], al[4];
          const int r0 = wm0 + mt * 16 + gq;
          split_tf32(As[kb + tq][r0], ah[0], al[0]);
          split_tf32(As[kb + tq][r0 + 8], ah[1], al[1]);
          split_tf32(As[kb + tq + 4][r0], ah[2], al[2]);
          split_tf32(As[kb + tq + 4][r0 + 8], ah[3], al[3]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            if constexpr (TCT == 3) {
              mma_tf32(accm[mt][nt], al, bh[nt][0], bh[nt][1]);
              mma_tf32(accm[mt][nt], ah, bl[nt][0], bl[nt][1]);
            }
            mma_tf32(accm[mt][nt], ah, bh[nt][0], bh[nt][1]);
          }
        }
      }
    } else {
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int g = 0; g < GA; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[kk][g * (BM / GA) + ty * 4]);
        av[4 * g] = v.x; av[4 * g + 1] = v.y; av[4 * g + 2] = v.z; av[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < GB; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][g * (BN / GB) + tx * 4]);
        bv[4 * g] = v.x; bv[4 * g + 1] = v.y; bv[4 * g + 2] = v.z; bv[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    }
    __syncthreads();
  }
  if constexpr (TC) {   // fragment layout of the accumulator: rows gq, gq + 8; columns 2 tq, 2 tq + 1
    const bool cvec2 = p.sCj == 1 && (p.N & 1) == 0 && (p.sCi & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 7) == 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = bi + wm0 + mt * 16 + gq + 8 * h;
        if (i >= p.M) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int j = bj + wn0 + nt * 8 + 2 * tq;
          const float v0 = accm[mt][nt][2 * h], v1 = accm[mt][nt][2 * h + 1];
          if (cvec2) {
            if (j >= p.N) continue;
            float* c = C + (long)i * p.sCi + j;
            float2 v = make_float2(p.alpha * v0, p.alpha * v1);
            if (p.mode == 0) {
              if (p.bias) { v.x += p.bias[j]; v.y += p.bias[j + 1]; }
              *reinterpret_cast<float2*>(c) = v;
            } else if (p.mode == 1) {
              const float2 o = *reinterpret_cast<const float2*>(c);
              *reinterpret_cast<float2*>(c) = make_float2(o.x + v.x, o.y + v.y);
            } else {
              red_add_v2(c, v.x, v.y);
            }
          } else {
            if (j < p.N) bmm_store(p, C, i, j, v0);
            if (j + 1 < p.N) bmm_store(p, C, i, j + 1, v1);
          }
        }
      }
    return;
  }
  const bool cvec = p.sCj == 1 && al4(p.N) && al4(p.sCi) && al16(C);
#pragma unroll
  for (int a = 0; a < TM; ++a) {
    const int i = bi + (a / 4) * (BM / GA) + ty * 4 + (a & 3);
    if (i >= p.M) continue;
#pragma unroll
    for (int g = 0; g < GB; ++g) {
      const int j = bj + g * (BN / GB) + tx * 4;
      if (cvec) {
        if (j >= p.N) continue;
        float* c = C + (long)i * p.sCi + j;
        float4 v = make_float4(p.alpha * acc[a][4 * g], p.alpha * acc[a][4 * g + 1], p.alpha * acc[a][4 * g + 2], p.alpha * acc[a][4 * g + 3]);
        if (p.mode == 0) {
          if (p.bias) { v.x += p.bias[j]; v.y += p.bias[j + 1]; v.z += p.bias[j + 2]; v.w += p.bias[j + 3]; }
          *reinterpret_cast<float4*>(c) = v;
        } else if (p.mode == 1) {
          const float4 o = *reinterpret_cast<const float4*>(c);
          *reinterpret_cast<float4*>(c) = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
        } else {
          red_add_v4(c, v.x, v.y, v.z, v.w);
        }
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (j + b < p.N) bmm_store(p, C, i, j + b, acc[a][4 * g + b]);
      }
    }
  }
}

// tile choice: predicted time = waves over the 148 SMs x work of a tile / relative speed of the tile shape
template <int BM, int BN, int TM, int TN, int TC>
void launch_tiled(Launcher& L, const Bmm& p) {
  const long gx = (p.N + BN - 1) / BN, gy = (p.M + BM - 1) / BM, gz = (long)p.Z1 * p.Z2;
  ts_bmm_tiled<BM, BN, TM, TN, TC><<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), 256, 0, L.st>>>(p);
  ++L.launches;
}
int pick_tile(const Bmm& p) {
  static const int bm[4] = {128, 128, 64, 64}, bn[4] = {128, 64, 128, 64}, per_sm[4] = {2, 2, 2, 4};
  static const double speed[4] = {1.0, 0.85, 0.85, 0.62};
  int best = 3;
  double best_t = 1e300;
  for (int c = 0; c < 4; ++c) {
    const long ctas = (long)((p.N + bn[c] - 1) / bn[c]) * ((p.M + bm[c] - 1) / bm[c]) * p.Z1 * p.Z2;
    const long slots = 148L * per_sm[c];
    const double waves = (double)((ctas + slots - 1) / slots);
    // a partial last wave still costs a full tile time, but fewer CTAs per SM run faster: count it at 0.6 + 0.4 * fill
    const long rem = ctas % slots;
    const double last = rem ? 0.6 + 0.4 * (double)rem / slots : 1.0;
    const double t = (waves - 1.0 + last) * bm[c] * bn[c] * per_sm[c] / speed[c];
    if (t < best_t) { best_t = t; best = c; }
  }
  return best;
}
#endif

void run_bmm(Launcher& L, const Bmm& p) {
  if (p.M <= 0 || p.N <= 0 || p.Z1 <= 0 || p.Z2 <= 0) return;
  if (p.K <= 0 && p.mode != 0) return;
#ifndef DHG_HOSTSIM
  if (g_use_tiled && (long)p.Z1 * p.Z2 <= 65535 && (p.M + 63) / 64 <= 65535) {
    const int tile = g_use_tiled == 2 ? 3 : pick_tile(p);
    if (g_use_tiled == 3) {          // 3 x TF32 on the tensor cores
      switch (tile) {
        case 0: launch_tiled<128, 128, 2, 4, 3>(L, p); break;
        case 1: launch_tiled<128, 64, 4, 2, 3>(L, p); break;
        case 2: launch_tiled<64, 128, 2, 4, 3>(L, p); break;
        default: launch_tiled<64, 64, 2, 4, 3>(L, p); break;
      }
    } else if (g_use_tiled == 4) {   // plain TF32
      switch (tile) {
        case 0: launch_tiled<128, 128, 2, 4, 1>(L, p); break;
        case 1: launch_tiled<128, 64, 4, 2, 1>(L, p); break;
        case 2: launch_tiled<64, 128, 2, 4, 1>(L, p); break;
        default: launch_tiled<64, 64, 2, 4, 1>(L, p); break;
      }
    } else {                         // fp32 FMAs on the CUDA cores
      switch (tile) {
        case 0: launch_tiled<128, 128, 8, 8, 0>(L, p); break;
        case 1: launch_tiled<128, 64, 8, 4, 0>(L, p); break;
        case 2: launch_tiled<64, 128, 4, 8, 0>(L, p); break;
        default: launch_tiled<64, 64, 4, 4, 0>(L, p); break;
      }
    }
    return;
  }
#endif
#ifdef DHG_HOSTSIM
  // tools/train_gemm_shapes.py: list the contractions of a plan (shape, batch, taps, store mode) without computing them
  static const bool log_only = getenv("DHG_TRAINER_LOG_BMM") != nullptr;
  if (log_only) {
    printf("BMM %d %d %d %d %d %d\n", p.M, p.N, p.K, p.Z1 * p.Z2, p.taps, p.mode);
    ++L.launches;
    return;
  }
#endif
  BmmBody f;
  f.p = p;
  f.tm = (p.M + 3) / 4;
  f.tn = (p.N + 3) / 4;
  L.run((long)f.tm * f.tn * p.Z1 * p.Z2, f);
}

// ------------------------------------------------------------------------------------------------------------------
// per-element / per-row kernels (forward and backward of each)
// ------------------------------------------------------------------------------------------------------------------
TS_FN float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// 16-byte versions of the streaming kernels (used when the channel count is a multiple of 4: every activation but sigma [B, 1])
struct alignas(16) f4 { float x, y, z, w; };
TS_FN f4 ld4(const float* p) { return *reinterpret_cast<const f4*>(p); }
TS_FN void st4(float* p, const f4& v) { *reinterpret_cast<f4*>(p) = v; }
TS_FN float silu_(float v) { return v * sigmoidf_(v); }
TS_FN float dsilu_(float v) { const float s = sigmoidf_(v); return s * (1.f + v * (1.f - s)); }
struct Silu4Fwd { const float* x; float* y; TS_FN void operator()(long i) const { const f4 v = ld4(x + 4 * i); st4(y + 4 * i, f4{silu_(v.x), silu_(v.y), silu_(v.z), silu_(v.w)}); } };
struct Silu4Bwd {
  const float* x; const float* gy; float* gx;
  TS_FN void operator()(long i) const {
    const f4 v = ld4(x + 4 * i), g = ld4(gy + 4 * i), o = ld4(gx + 4 * i);
    st4(gx + 4 * i, f4{o.x + g.x * dsilu_(v.x), o.y + g.y * dsilu_(v.y), o.z + g.z * dsilu_(v.z), o.w + g.w * dsilu_(v.w)});
  }
};
struct Add4Fwd { const float* a; const float* b; float* y; TS_FN void operator()(long i) const { const f4 u = ld4(a + 4 * i), v = ld4(b + 4 * i); st4(y + 4 * i, f4{u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w}); } };
struct Acc4Bwd { const float* gy; float* gx; TS_FN void operator()(long i) const { const f4 g = ld4(gy + 4 * i), o = ld4(gx + 4 * i); st4(gx + 4 * i, f4{o.x + g.x, o.y + g.y, o.z + g.z, o.w + g.w}); } };
struct Add4Bwd { const float* gy; float* ga; float* gb; TS_FN void operator()(long i) const { if (ga) Acc4Bwd{gy, ga}(i); if (gb) Acc4Bwd{gy, gb}(i); } };
struct AddPe4Fwd {
  const float* x; const float* pe; float* y; int C, period;
  TS_FN void operator()(long i) const { const long e = 4 * i, r = e / C; const f4 u = ld4(x + e), v = ld4(pe + (r % period) * C + (e % C)); st4(y + e, f4{u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w}); }
};
struct Film4Fwd {
  const float* x; const float* gam; const float* bet; float* y; int C, period;
  TS_FN void operator()(long i) const {
    const long e = 4 * i, o = (e / ((long)C * period)) * C + (e % C);
    const f4 u = ld4(x + e), g = ld4(gam + o), b = ld4(bet + o);
    st4(y + e, f4{u.x * g.x + b.x, u.y * g.y + b.y, u.z * g.z + b.z, u.w * g.w + b.w});
  }
};
struct Film4BwdX {
  const float* gy; const float* gam; float* gx; int C, period;
  TS_FN void operator()(long i) const {
    const long e = 4 * i, o = (e / ((long)C * period)) * C + (e % C);
    const f4 d = ld4(gy + e), g = ld4(gam + o), a = ld4(gx + e);
    st4(gx + e, f4{a.x + d.x * g.x, a.y + d.y * g.y, a.z + d.z * g.z, a.w + d.w * g.w});
  }
};
struct SiluFwd { const float* x; float* y; TS_FN void operator()(long i) const { const float v = x[i]; y[i] = v * sigmoidf_(v); } };
struct SiluBwd {
  const float* x; const float* gy; float* gx;
  TS_FN void operator()(long i) const { const float v = x[i], s = sigmoidf_(v); gx[i] += gy[i] * (s * (1.f + v * (1.f - s))); }
};
struct SigmoidFwd { const float* x; float* y; TS_FN void operator()(long i) const { y[i] = sigmoidf_(x[i]); } };
struct SigmoidBwd { const float* y; const float* gy; float* gx; TS_FN void operator()(long i) const { const float s = y[i]; gx[i] += gy[i] * s * (1.f - s); } };
struct MulFwd { const float* x; const float* m; float* y; TS_FN void operator()(long i) const { y[i] = m ? x[i] * m[i] : x[i]; } };
struct AddFwd { const float* a; const float* b; float* y; TS_FN void operator()(long i) const { y[i] = a[i] + b[i]; } };
struct AddBwd { const float* gy; float* ga; float* gb; TS_FN void operator()(long i) const { const float g = gy[i]; if (ga) ga[i] += g; if (gb) gb[i] += g; } };
// y[r, c] = x[r, c] + pe[(r % period), c]   (attention.py:15-23 tables, constant)
struct AddPeFwd { const float* x; const float* pe; float* y; int C, period; TS_FN void operator()(long i) const { const long r = i / C; y[i] = x[i] + pe[(r % period) * C + (i % C)]; } };
struct AccBwd { const float* gy; float* gx; TS_FN void operator()(long i) const { gx[i] += gy[i]; } };
// FiLM (conditioning.py:16-19): y[r, c] = x[r, c] * gamma[b, c] + beta[b, c], b = r / period
struct FilmFwd {
  const float* x; const float* gam; const float* bet; float* y; int C, period;
  TS_FN void operator()(long i) const { const long b = i / ((long)C * period); const int c = (int)(i % C); y[i] = x[i] * gam[b * C + c] + bet[b * C + c]; }
};
struct FilmBwdX {
  const float* gy; const float* gam; float* gx; int C, period;
  TS_FN void operator()(long i) const { const long b = i / ((long)C * period); gx[i] += gy[i] * gam[b * C + (i % C)]; }
};
struct FilmBwdCond {   // one thread per (b, 32-row chunk, c): partial sums over the chunk, atomics into [B, C]
  const float* gy; const float* x; float* ggam; float* gbet; int C, period, chunks;
  TS_FN void operator()(long i) const {
    const int c = (int)(i % C); const long q = i / C; const int ch = (int)(q % chunks); const long b = q / chunks;
    const int t0 = ch * 32, t1 = t0 + 32 < period ? t0 + 32 : period;
    const float* g = gy + b * period * (long)C + c; const float* xv = x + b * period * (long)C + c;
    float sg = 0.f, sb = 0.f;
    for (int t = t0; t < t1; ++t) { const float d = g[(long)t * C]; sg = fmaf(d, xv[(long)t * C], sg); sb += d; }
    ts_atomic_add(ggam + b * C + c, sg); ts_atomic_add(gbet + b * C + c, sb);
  }
};
// LayerNorm(eps 1e-6, no affine) (model.py:25, text_style.py:80); one thread per row
struct LnFwd {
  const float* x; float* y; float* rstd; int C;
  TS_FN void operator()(long r) const {
    const float* xr = x + r * C; float* yr = y + r * C;
    float m = 0.f; for (int c = 0; c < C; ++c) m += xr[c]; m /= C;
    float v = 0.f; for (int c = 0; c < C; ++c) { const float d = xr[c] - m; v = fmaf(d, d, v); } v /= C;
    const float rs = 1.f / sqrtf(v + 1e-6f); rstd[r] = rs;
    for (int c = 0; c < C; ++c) yr[c] = (xr[c] - m) * rs;
  }
#ifndef DHG_HOSTSIM
  __device__ void warp(long r, int lane) const {
    const float* xr = x + r * C; float* yr = y + r * C;
    float m = 0.f; for (int c = lane; c < C; c += 32) m += xr[c]; m = warp_sum(m) / C;
    float v = 0.f; for (int c = lane; c < C; c += 32) { const float d = xr[c] - m; v = fmaf(d, d, v); } v = warp_sum(v) / C;
    const float rs = 1.f / sqrtf(v + 1e-6f); if (lane == 0) rstd[r] = rs;
    for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - m) * rs;
  }
#endif
};
struct LnBwd {
  const float* y; const float* gy; const float* rstd; float* gx; int C;
  TS_FN void operator()(long r) const {
    const float* yr = y + r * C; const float* gr = gy + r * C; float* gxr = gx + r * C;
    float m1 = 0.f, m2 = 0.f; for (int c = 0; c < C; ++c) { m1 += gr[c]; m2 = fmaf(gr[c], yr[c], m2); } m1 /= C; m2 /= C;
    const float rs = rstd[r];
    for (int c = 0; c < C; ++c) gxr[c] += rs * (gr[c] - m1 - yr[c] * m2);
  }
#ifndef DHG_HOSTSIM
  __device__ void warp(long r, int lane) const {
    const float* yr = y + r * C; const float* gr = gy + r * C; float* gxr = gx + r * C;
    float m1 = 0.f, m2 = 0.f; for (int c = lane; c < C; c += 32) { m1 += gr[c]; m2 = fmaf(gr[c], yr[c], m2); }
    m1 = warp_sum(m1) / C; m2 = warp_sum(m2) / C;
    const float rs = rstd[r];
    for (int c = lane; c < C; c += 32) gxr[c] += rs * (gr[c] - m1 - yr[c] * m2);
  }
#endif
};
// softmax over the keys of one (b, h, query) row, with the additive -1e9 padding mask (attention.py:43, utils/nn.py:178-191)
struct SoftmaxFwd {
  const float* s; float* p; const int64_t* ids; int Tk; long rows_per_b;
  TS_FN void operator()(long r) const {
    const float* sr = s + r * Tk; float* pr = p + r * Tk;
    const int64_t* id = ids ? ids + (r / rows_per_b) * Tk : nullptr;
    float mx = -INFINITY;
    for (int j = 0; j < Tk; ++j) { const float v = sr[j] + ((id && id[j] == 0) ? -1e9f : 0.f); pr[j] = v; mx = fmaxf(mx, v); }
    float sum = 0.f; for (int j = 0; j < Tk; ++j) { const float e = expf(pr[j] - mx); pr[j] = e; sum += e; }
    const float inv = 1.f / sum; for (int j = 0; j < Tk; ++j) pr[j] *= inv;
  }
#ifndef DHG_HOSTSIM
  __device__ void warp(long r, int lane) const {
    const float* sr = s + r * Tk; float* pr = p + r * Tk;
    const int64_t* id = ids ? ids + (r / rows_per_b) * Tk : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < Tk; j += 32) { const float v = sr[j] + ((id && id[j] == 0) ? -1e9f : 0.f); pr[j] = v; mx = fmaxf(mx, v); }
    mx = warp_max(mx);
    float sum = 0.f; for (int j = lane; j < Tk; j += 32) { const float e = expf(pr[j] - mx); pr[j] = e; sum += e; }
    const float inv = 1.f / warp_sum(sum); for (int j = lane; j < Tk; j += 32) pr[j] *= inv;
  }
#endif
};
struct SoftmaxBwd {
  const float* p; const float* gp; float* gs; int Tk;
  TS_FN void operator()(long r) const {
    const float* pr = p + r * Tk; const float* gr = gp + r * Tk; float* go = gs + r * Tk;
    float d = 0.f; for (int j = 0; j < Tk; ++j) d = fmaf(pr[j], gr[j], d);
    for (int j = 0; j < Tk; ++j) go[j] += pr[j] * (gr[j] - d);
  }
#ifndef DHG_HOSTSIM
  __device__ void warp(long r, int lane) const {
    const float* pr = p + r * Tk; const float* gr = gp + r * Tk; float* go = gs + r * Tk;
    float d = 0.f; for (int j = lane; j < Tk; j += 32) d = fmaf(pr[j], gr[j], d);
    d = warp_sum(d);
    for (int j = lane; j < Tk; j += 32) go[j] += pr[j] * (gr[j] - d);
  }
#endif
};
// AvgPool1d(2) / nearest x2 over the rows of each sample (model.py:93-99, 163)
struct PoolFwd { const float* x; float* y; int C; TS_FN void operator()(long i) const { const long r = i / C; const int c = (int)(i % C); y[i] = 0.5f * (x[(2 * r) * C + c] + x[(2 * r + 1) * C + c]); } };
struct PoolBwd { const float* gy; float* gx; int C; TS_FN void operator()(long i) const { const long r = i / C; const int c = (int)(i % C); gx[i] += 0.5f * gy[(r >> 1) * C + c]; } };
struct UpFwd { const float* x; float* y; int C; TS_FN void operator()(long i) const { const long r = i / C; const int c = (int)(i % C); y[i] = x[(r >> 1) * C + c]; } };
struct UpBwd { const float* gy; float* gx; int C; TS_FN void operator()(long i) const { const long r = i / C; const int c = (int)(i % C); gx[i] += gy[(2 * r) * C + c] + gy[(2 * r + 1) * C + c]; } };
// nn.Embedding(73, d) (text_style.py:70).  Ids are validated by the caller (DenoiserTrainer raises IndexError like the
// reference); the clamp only keeps a bad id from reading outside the table.
TS_FN long embed_row(int64_t id) { return id < 0 ? 0 : id > 72 ? 72 : (long)id; }
struct EmbedFwd { const int64_t* ids; const float* E; float* y; int C; TS_FN void operator()(long i) const { y[i] = E[embed_row(ids[i / C]) * C + (i % C)]; } };
struct EmbedBwd { const int64_t* ids; const float* gy; float* gE; int C; TS_FN void operator()(long i) const { ts_atomic_add(gE + embed_row(ids[i / C]) * C + (i % C), gy[i]); } };
struct ConvWPack {   // W [n][k][tap] -> wf [tap][k][n], wd [tap][n][k]
  const float* W; float* wf; float* wd; int N, K;
  TS_FN void operator()(long i) const {
    const int tap = (int)(i % 3); const long e = i / 3; const int k = (int)(e % K), n = (int)(e / K);
    const float v = W[i];
    wf[((long)tap * K + k) * N + n] = v; wd[((long)tap * N + n) * K + k] = v;
  }
};
struct ConvWFold { const float* tmp; float* gW; long nk; TS_FN void operator()(long i) const { const long tap = i % 3, e = i / 3; gW[i] = tmp[tap * nk + e]; } };
// column sums of a [rows, N] matrix in chunks of 64 rows (bias gradients)
struct ColSum {
  const float* g; float* out; int rows, N;
  TS_FN void operator()(long i) const {
    const int j = (int)(i % N); const long ch = i / N; const long r0 = ch * 64; const long r1 = r0 + 64 < rows ? r0 + 64 : rows;
    float s = 0.f; for (long r = r0; r < r1; ++r) s += g[r * N + j];
    ts_atomic_add(out + j, s);
  }
};

// ------------------------------------------------------------------------------------------------------------------
// parameter layout: checkpoint key order (checkpoint.py:256-297, train.py:131)
// ------------------------------------------------------------------------------------------------------------------
struct PSpec { std::string name; long off, numel; };
struct Layout {
  std::vector<PSpec> items;
  std::map<std::string, int> index;
  long total = 0;
  void add(const std::string& n, long numel) { index[n] = (int)items.size(); items.push_back({n, total, numel}); total += numel; }
  void lin(const std::string& p, int in, int out) { add(p + ".weight", (long)in * out); add(p + ".bias", out); }
  void conv(const std::string& p, int in, int out) { add(p + ".weight", (long)in * out * 3); add(p + ".bias", out); }
  void affine(const std::string& p, int h) { lin(p + ".gamma_emb", 32, h); lin(p + ".beta_emb", 32, h); }
  void mha(const std::string& p, int d) { for (const char* n : {"wq", "wk", "wv", "dense"}) lin(p + "." + n, d, d); }
  void conv_block(const std::string& p, int in, int out) {
    affine(p + ".affine1", out / 2); affine(p + ".affine2", out); affine(p + ".affine3", out);
    conv(p + ".conv_skip", in, out); conv(p + ".conv1", in, out / 2); conv(p + ".conv2", out / 2, out); lin(p + ".fc", out, out);
  }
  void enc_layer(const std::string& p, int in, int out) {
    lin(p + ".text_dense", in, out); lin(p + ".ffn.1", out, 2 * out); lin(p + ".ffn.3", 2 * out, out);
    mha(p + ".mha", out); mha(p + ".mha2", out);
    for (int i = 0; i < 4; ++i) affine(p + ".affine" + std::to_string(i), out);
  }
  void build(int num_layers, int ch) {
    const int c1 = ch, c2 = ch * 3 / 2, c3 = ch * 2, d = 2 * c2;
    lin("input_dense", 2, c1); lin("sigma_ffn.1", 1, 2048); lin("sigma_ffn.3", 2048, c1 / 4);
    conv_block("enc1", c1, c1); conv_block("enc2", c1, c2); enc_layer("enc3", d, c2);
    conv_block("enc4", c2, c3); enc_layer("enc5", d, c3);
    conv("skip_conv1", c1, c2); conv("skip_conv2", c2, c3); conv("skip_conv3", c3, d);
    add("text_style_model.emb.weight", 73L * d);
    lin("text_style_model.style_ffn.1", 256, 4 * c2); lin("text_style_model.style_ffn.3", 4 * c2, d);
    lin("text_style_model.text_ffn.1", d, 2 * d); lin("text_style_model.text_ffn.3", 2 * d, d);
    mha("text_style_model.mha", d);
    for (int i = 1; i <= 4; ++i) affine("text_style_model.affine" + std::to_string(i), d);
    lin("att_dense", 2 * c1, d);
    for (int i = 0; i < num_layers; ++i) enc_layer("att_layers." + std::to_string(i), d, d);
    conv_block("dec3", d, c3); conv_block("dec2", c3, c2); conv_block("dec1", c2, c1);
    lin("output_dense", c1, 2); lin("pen_lifts_dense.0", c1, 1);
  }
};

// ------------------------------------------------------------------------------------------------------------------
// tape
// ------------------------------------------------------------------------------------------------------------------
struct Ten {
  float* v = nullptr;
  float* g = nullptr;   // null: no gradient wanted (inputs)
  int rows = 0, C = 0, period = 1;   // rows = samples * period
  bool ng = false;                    // an input (or a pure function of inputs): no gradient is ever needed
  long n() const { return (long)rows * C; }
};
struct OpRec { std::function<void(Launcher&)> fwd, bwd; };

}  // namespace

struct dhg_trainer {
  int device = 0, num_layers = 2, ch = 128, B = 0, T = 0, L = 0;
  int c1, c2, c3, d;
  Layout lay;
  float* params = nullptr;
  float* grads = nullptr;
  // arenas: values, gradients (zeroed at the start of every backward), constants
  bool dry = true;
  size_t v_need = 0, g_need = 0, v_used = 0, g_used = 0;
  float* v_arena = nullptr;
  float* g_arena = nullptr;
  std::vector<void*> consts;
  std::vector<OpRec> tape;
#ifndef DHG_HOSTSIM
  cudaStream_t side_stream = nullptr;
  std::vector<cudaEvent_t> side_events;
#endif
  std::string err;
  // plan-owned copies of the inputs, and the outputs
  Ten in_x, in_sigma, in_style, style_keep, out_score, out_pen;
  int64_t* text = nullptr;
  bool have_keep = false;
  long last_launches = 0;

  float* P(const std::string& n) { auto it = lay.index.find(n); if (it == lay.index.end()) { err = "unknown parameter " + n; return params; } return params + lay.items[it->second].off; }
  float* G(const std::string& n) { auto it = lay.index.find(n); if (it == lay.index.end()) { err = "unknown parameter " + n; return grads; } return grads + lay.items[it->second].off; }

  Ten make(int rows, int C, int period, bool grad = true) {
    Ten t; t.rows = rows; t.C = C; t.period = period;
    const size_t n = ((size_t)rows * C + 63) / 64 * 64;
    if (dry) { v_need += n; if (grad) g_need += n; return t; }
    t.v = v_arena + v_used; v_used += n;
    if (grad) { t.g = g_arena + g_used; g_used += n; }
    return t;
  }
  void rec(std::function<void(Launcher&)> f, std::function<void(Launcher&)> b) { if (!dry) tape.push_back({std::move(f), std::move(b)}); }

  // ---- ops ----
  Ten unary_silu(const Ten& x) {
    Ten y = make(x.rows, x.C, x.period, !x.ng);
    y.ng = x.ng;
    const bool v4 = x.C % 4 == 0;
    rec([=](Launcher& L) { if (v4) L.run(x.n() / 4, Silu4Fwd{x.v, y.v}); else L.run(x.n(), SiluFwd{x.v, y.v}); },
        [=](Launcher& L) { if (!x.g) return; if (v4) L.run(x.n() / 4, Silu4Bwd{x.v, y.g, x.g}); else L.run(x.n(), SiluBwd{x.v, y.g, x.g}); });
    return y;
  }
  Ten sigmoid(const Ten& x) {
    Ten y = make(x.rows, x.C, x.period);
    rec([=](Launcher& L) { L.run(x.n(), SigmoidFwd{x.v, y.v}); }, [=](Launcher& L) { if (x.g) L.run(x.n(), SigmoidBwd{y.v, y.g, x.g}); });
    return y;
  }
  Ten add(const Ten& a, const Ten& b, bool join = false) {   // join: an operand was produced on the second stream
    Ten y = make(a.rows, a.C, a.period);
    const bool v4 = a.C % 4 == 0;
    rec([=](Launcher& L) { if (join) L.side_join(); if (v4) L.run(a.n() / 4, Add4Fwd{a.v, b.v, y.v}); else L.run(a.n(), AddFwd{a.v, b.v, y.v}); },
        [=](Launcher& L) { if (v4) L.run(a.n() / 4, Add4Bwd{y.g, a.g, b.g}); else L.run(a.n(), AddBwd{y.g, a.g, b.g}); });
    return y;
  }
  Ten add_pe(const Ten& x, const float* pe) {
    Ten y = make(x.rows, x.C, x.period);
    const bool v4 = x.C % 4 == 0;
    rec([=](Launcher& L) { if (v4) L.run(x.n() / 4, AddPe4Fwd{x.v, pe, y.v, x.C, x.period}); else L.run(x.n(), AddPeFwd{x.v, pe, y.v, x.C, x.period}); },
        [=](Launcher& L) { if (!x.g) return; if (v4) L.run(x.n() / 4, Acc4Bwd{y.g, x.g}); else L.run(x.n(), AccBwd{y.g, x.g}); });
    return y;
  }
  Ten pool(const Ten& x) {
    Ten y = make(x.rows / 2, x.C, x.period / 2);
    rec([=](Launcher& L) { L.run(y.n(), PoolFwd{x.v, y.v, x.C}); }, [=](Launcher& L) { if (x.g) L.run(x.n(), PoolBwd{y.g, x.g, x.C}); });
    return y;
  }
  Ten up(const Ten& x) {
    Ten y = make(x.rows * 2, x.C, x.period * 2);
    rec([=](Launcher& L) { L.run(y.n(), UpFwd{x.v, y.v, x.C}); }, [=](Launcher& L) { if (x.g) L.run(x.n(), UpBwd{y.g, x.g, x.C}); });
    return y;
  }
  Ten ln(const Ten& x) {
    Ten y = make(x.rows, x.C, x.period);
    Ten rs = make(x.rows, 1, x.period, false);
    rec([=](Launcher& L) { L.run_rows(x.rows, LnFwd{x.v, y.v, rs.v, x.C}); }, [=](Launcher& L) { if (x.g) L.run_rows(x.rows, LnBwd{y.v, y.g, rs.v, x.g, x.C}); });
    return y;
  }
  // weight-gradient batching of a Linear: the rows are cut into Z equal runs of whole samples, about 512 rows each (one
  // batch item per run, reduced into dW with atomics); small matrices are one item
  static void wgrad_split(const Ten& x, int& Z, int& rpz) {
    Z = 1; rpz = x.rows;
    if (x.rows <= 512 || x.rows % x.period) return;
    const int nb = x.rows / x.period;
    int g = 1;
    for (int c = 1; c <= nb; ++c) if (nb % c == 0 && (long)c * x.period <= 512) g = c;
    Z = nb / g; rpz = g * x.period;
  }
  // nn.Linear: W [N, K], b [N]
  Ten linear(const Ten& x, const std::string& name, int N) {
    const int K = x.C;
    Ten y = make(x.rows, N, x.period);
    const float* W = P(name + ".weight"); const float* b = P(name + ".bias");
    float* gW = G(name + ".weight"); float* gb = G(name + ".bias");
    const long have = lay.index.count(name + ".weight") ? lay.items[lay.index[name + ".weight"]].numel : -1;
    if (have != (long)N * K) err = "shape mismatch at " + name;
    rec([=](Launcher& L) {
          Bmm p; p.A = x.v; p.B = W; p.C = y.v; p.bias = b; p.M = x.rows; p.N = N; p.K = K;
          p.sAi = K; p.sAk = 1; p.sBk = 1; p.sBj = K; p.sCi = N; p.sCj = 1;
          run_bmm(L, p);
        },
        [=](Launcher& L) {
          if (x.g) {   // dx += dy W
            Bmm p; p.A = y.g; p.B = W; p.C = x.g; p.M = x.rows; p.N = K; p.K = N; p.mode = 1;
            p.sAi = N; p.sAk = 1; p.sBk = K; p.sBj = 1; p.sCi = K; p.sCj = 1;
            run_bmm(L, p);
          }
          L.side_begin();
          int Z, rpz; wgrad_split(x, Z, rpz);
          Bmm q; q.A = y.g; q.B = x.v; q.C = gW; q.M = N; q.N = K; q.K = rpz; q.Z1 = Z; q.mode = 2;   // dW[n, k] += sum_r dy[r, n] x[r, k]  (k fastest: coalesced atomics)
          q.sAz1 = (long)rpz * N; q.sAi = 1; q.sAk = N; q.sBz1 = (long)rpz * K; q.sBk = K; q.sBj = 1; q.sCz1 = 0; q.sCi = K; q.sCj = 1;
          run_bmm(L, q);
          L.run((long)N * ((x.rows + 63) / 64), ColSum{y.g, gb, x.rows, N});
          L.side_end();
        });
    return y;
  }
  // nn.Conv1d(k = 3, padding 1) in channels-last form (cnn.py:32-47): W [N, K, 3]; tap j reads row t + j - 1 of the same sample
  // side = true: the forward launch goes to the second stream (a branch that is consumed later: conv_skip of a ConvBlock,
  // the U-Net's skip convolutions); the consumer is an add(..., join = true)
  Ten conv3(const Ten& x, const std::string& name, int N, bool side = false) {
    const int K = x.C, Tn = x.period, nb = x.rows / x.period;
    Ten y = make(x.rows, N, x.period);
    const float* W = P(name + ".weight"); const float* b = P(name + ".bias");
    float* gW = G(name + ".weight"); float* gb = G(name + ".bias");
    const long have = lay.index.count(name + ".weight") ? lay.items[lay.index[name + ".weight"]].numel : -1;
    if (have != (long)N * K * 3) err = "shape mismatch at " + name;
    auto range = [=](int tap, int& lo, int& hi) { lo = tap == 0 ? 1 : 0; hi = tap == 2 ? Tn - 1 : Tn; };
    Ten wtmp = make(3 * N, K, 1);   // only its gradient half is used: zeroed with the arena at the start of every backward
    // the checkpoint layout [n][k][tap] has no contiguous axis for either contraction (stride 3 / 3K): repacked once per
    // forward into [tap][k][n] (forward: n contiguous) and [tap][n][k] (data gradient: k contiguous); 2 x 12 bytes per weight
    Ten wf = make(3 * K, N, 1, false), wd = make(3 * N, K, 1, false);
    rec([=](Launcher& L) {   // y[t, n] = b[n] + sum_tap sum_k x[t + tap - 1, k] W[n, k, tap], one batch item per sample
          if (side) L.side_begin();
          L.run(3L * N * K, ConvWPack{W, wf.v, wd.v, N, K});
          Bmm p; p.A = x.v; p.B = wf.v; p.C = y.v; p.bias = b; p.M = Tn; p.N = N; p.K = K; p.Z1 = nb;
          p.taps = 3; p.shift0 = -1; p.dshift = 1; p.sBtap = (long)K * N;
          p.sAz1 = (long)Tn * K; p.sAi = K; p.sAk = 1; p.sBk = N; p.sBj = 1; p.sCz1 = (long)Tn * N; p.sCi = N; p.sCj = 1;
          run_bmm(L, p);
          if (side) L.side_end();
        },
        [=](Launcher& L) {
          if (x.g) {   // dx[t, k] += sum_tap sum_n dy[t - (tap - 1), n] W[n, k, tap]
            Bmm p; p.A = y.g; p.B = wd.v; p.C = x.g; p.M = Tn; p.N = K; p.K = N; p.Z1 = nb; p.mode = 1;
            p.taps = 3; p.shift0 = 1; p.dshift = -1; p.sBtap = (long)N * K;
            p.sAz1 = (long)Tn * N; p.sAi = N; p.sAk = 1; p.sBk = K; p.sBj = 1; p.sCz1 = (long)Tn * K; p.sCi = K; p.sCj = 1;
            run_bmm(L, p);
          }
          L.side_begin();
          for (int tap = 0; tap < 3; ++tap) {   // scratch[tap][n][k] += sum_b sum_t dy[t, n] x[t + tap - 1, k]  (k contiguous: vector atomics)
            int lo, hi; range(tap, lo, hi);
            Bmm q; q.A = y.g + (long)lo * N; q.B = x.v + (long)(lo + tap - 1) * K; q.C = wtmp.g + (long)tap * N * K; q.M = N; q.N = K; q.K = hi - lo; q.Z1 = nb; q.mode = 2;
            q.sAz1 = (long)Tn * N; q.sAi = 1; q.sAk = N; q.sBz1 = (long)Tn * K; q.sBk = K; q.sBj = 1; q.sCz1 = 0; q.sCi = K; q.sCj = 1;
            run_bmm(L, q);
          }
          L.run(3L * N * K, ConvWFold{wtmp.g, gW, (long)N * K});   // dW[n, k, tap] = scratch[tap][n][k]
          L.run((long)N * ((x.rows + 63) / 64), ColSum{y.g, gb, x.rows, N});
          L.side_end();
        });
    return y;
  }
  // AffineTransformLayer (conditioning.py:5-19): gamma / beta = Linear(32, C)(sigma embedding) per sample
  Ten film(const Ten& x, const Ten& sig, const std::string& name) {
    Ten gam = linear(sig, name + ".gamma_emb", x.C);
    Ten bet = linear(sig, name + ".beta_emb", x.C);
    Ten y = make(x.rows, x.C, x.period);
    const int nb = x.rows / x.period;
    const bool v4 = x.C % 4 == 0;
    rec([=](Launcher& L) { if (v4) L.run(x.n() / 4, Film4Fwd{x.v, gam.v, bet.v, y.v, x.C, x.period}); else L.run(x.n(), FilmFwd{x.v, gam.v, bet.v, y.v, x.C, x.period}); },
        [=](Launcher& L) {
          if (x.g) { if (v4) L.run(x.n() / 4, Film4BwdX{y.g, gam.v, x.g, x.C, x.period}); else L.run(x.n(), FilmBwdX{y.g, gam.v, x.g, x.C, x.period}); }
          const int chunks = (x.period + 31) / 32;
          L.run((long)nb * chunks * x.C, FilmBwdCond{y.g, x.v, gam.g, bet.g, x.C, x.period, chunks});
        });
    return y;
  }
  // ff_network(act_before=True) (utils/nn.py:145-175): SiLU -> Linear -> SiLU -> Linear
  Ten ffn(const Ten& x, const std::string& name, int hidden, int out) {
    return linear(unary_silu(linear(unary_silu(x), name + ".1", hidden)), name + ".3", out);
  }
  // MultiHeadAttention.forward (attention.py:49-87) with scaled_dp_attn (attention.py:26-46); mask_ids: key token ids or null
  Ten mha(const std::string& name, const Ten& q_in, const Ten& k_in, const Ten& v_in, int heads, const int64_t* mask_ids) {
    const int dm = q_in.C, D = dm / heads, Tq = q_in.period, Tk = k_in.period, nb = q_in.rows / q_in.period;
    Ten q = linear(q_in, name + ".wq", dm), k = linear(k_in, name + ".wk", dm), v = linear(v_in, name + ".wv", dm);
    Ten s = make(nb * heads * Tq, Tk, heads * Tq), pr = make(nb * heads * Tq, Tk, heads * Tq), o = make(q_in.rows, dm, Tq);
    const float scale = 1.f / sqrtf((float)D);
    const long sQ = (long)Tq * dm, sK = (long)Tk * dm, sS1 = (long)heads * Tq * Tk, sS2 = (long)Tq * Tk;
    rec([=](Launcher& L) {
          Bmm a; a.A = q.v; a.B = k.v; a.C = s.v; a.M = Tq; a.N = Tk; a.K = D; a.Z1 = nb; a.Z2 = heads; a.alpha = scale;   // S = scale Q K^T
          a.sAz1 = sQ; a.sAz2 = D; a.sAi = dm; a.sAk = 1; a.sBz1 = sK; a.sBz2 = D; a.sBk = 1; a.sBj = dm; a.sCz1 = sS1; a.sCz2 = sS2; a.sCi = Tk; a.sCj = 1;
          run_bmm(L, a);
          L.run_rows(s.rows, SoftmaxFwd{s.v, pr.v, mask_ids, Tk, (long)heads * Tq});
          Bmm c; c.A = pr.v; c.B = v.v; c.C = o.v; c.M = Tq; c.N = D; c.K = Tk; c.Z1 = nb; c.Z2 = heads;                    // O = P V
          c.sAz1 = sS1; c.sAz2 = sS2; c.sAi = Tk; c.sAk = 1; c.sBz1 = sK; c.sBz2 = D; c.sBk = dm; c.sBj = 1; c.sCz1 = sQ; c.sCz2 = D; c.sCi = dm; c.sCj = 1;
          run_bmm(L, c);
        },
        [=](Launcher& L) {
          Bmm a; a.A = pr.v; a.B = o.g; a.C = v.g; a.M = Tk; a.N = D; a.K = Tq; a.Z1 = nb; a.Z2 = heads; a.mode = 1;          // dV += P^T dO
          a.sAz1 = sS1; a.sAz2 = sS2; a.sAi = 1; a.sAk = Tk; a.sBz1 = sQ; a.sBz2 = D; a.sBk = dm; a.sBj = 1; a.sCz1 = sK; a.sCz2 = D; a.sCi = dm; a.sCj = 1;
          run_bmm(L, a);
          Bmm b; b.A = o.g; b.B = v.v; b.C = pr.g; b.M = Tq; b.N = Tk; b.K = D; b.Z1 = nb; b.Z2 = heads; b.mode = 1;          // dP += dO V^T
          b.sAz1 = sQ; b.sAz2 = D; b.sAi = dm; b.sAk = 1; b.sBz1 = sK; b.sBz2 = D; b.sBk = 1; b.sBj = dm; b.sCz1 = sS1; b.sCz2 = sS2; b.sCi = Tk; b.sCj = 1;
          run_bmm(L, b);
          L.run_rows(s.rows, SoftmaxBwd{pr.v, pr.g, s.g, Tk});
          Bmm c; c.A = s.g; c.B = k.v; c.C = q.g; c.M = Tq; c.N = D; c.K = Tk; c.Z1 = nb; c.Z2 = heads; c.mode = 1; c.alpha = scale;   // dQ += scale dS K
          c.sAz1 = sS1; c.sAz2 = sS2; c.sAi = Tk; c.sAk = 1; c.sBz1 = sK; c.sBz2 = D; c.sBk = dm; c.sBj = 1; c.sCz1 = sQ; c.sCz2 = D; c.sCi = dm; c.sCj = 1;
          run_bmm(L, c);
          Bmm e; e.A = s.g; e.B = q.v; e.C = k.g; e.M = Tk; e.N = D; e.K = Tq; e.Z1 = nb; e.Z2 = heads; e.mode = 1; e.alpha = scale;   // dK += scale dS^T Q
          e.sAz1 = sS1; e.sAz2 = sS2; e.sAi = 1; e.sAk = Tk; e.sBz1 = sQ; e.sBz2 = D; e.sBk = dm; e.sBj = 1; e.sCz1 = sK; e.sCz2 = D; e.sCi = dm; e.sCj = 1;
          run_bmm(L, e);
        });
    return linear(o, name + ".dense", dm);
  }
  // ConvBlock.forward (cnn.py:52-87)
  Ten conv_block(const std::string& p, const Ten& x, const Ten& sig, int out) {
    Ten skip = conv3(x, p + ".conv_skip", out, true);
    Ten y = film(conv3(unary_silu(x), p + ".conv1", out / 2), sig, p + ".affine1");
    y = film(conv3(unary_silu(y), p + ".conv2", out), sig, p + ".affine2");
    y = film(linear(unary_silu(y), p + ".fc", out), sig, p + ".affine3");
    return add(y, skip, true);
  }
  // attention.py:15-23: halves concatenated (sin | cos), fp32 like the reference
  const float* pos_table(int length, int dim, float pos_factor) {
    if (dry) return nullptr;
    const int half = dim / 2;
    std::vector<float> h((size_t)length * dim);
    const double step = log(10000.0) / (half - 1);
    for (int t = 0; t < length; ++t)
      for (int j = 0; j < half; ++j) {
        const float freq = expf((float)j * (float)(-step));
        const float ang = (float)t * freq * pos_factor;
        h[(size_t)t * dim + j] = sinf(ang);
        h[(size_t)t * dim + half + j] = cosf(ang);
      }
    float* dptr = (float*)dev_alloc(h.size() * sizeof(float));
    if (!dptr) { err = "out of memory (position table)"; return nullptr; }
    consts.push_back(dptr);
    to_dev(dptr, h.data(), h.size() * sizeof(float));
    return dptr;
  }
  // EncoderLayer.forward (model.py:35-58)
  Ten encoder_layer(const std::string& p, const Ten& x, const Ten& text_in, const Ten& sig, int heads, float pos_factor) {
    const int dm = x.C;
    Ten t = film(ln(linear(unary_silu(text_in), p + ".text_dense", dm)), sig, p + ".affine0");
    Ten t_pe = add_pe(t, pos_table(t.period, dm, 1.f));
    const float* xpos = pos_table(x.period, dm, pos_factor);
    Ten x_pe = add_pe(x, xpos);
    Ten x2 = mha(p + ".mha", x_pe, t_pe, t, heads, text);                 // v carries no PE; padding mask on the text keys
    x2 = add(film(ln(x2), sig, p + ".affine1"), x);                        // no residual inside this LayerNorm
    Ten x2_pe = add_pe(x2, xpos);
    Ten x3 = mha(p + ".mha2", x2_pe, x2_pe, x2, heads, nullptr);
    x3 = film(ln(add(x2, x3)), sig, p + ".affine2");
    Ten x4 = add(ffn(x3, p + ".ffn", 2 * dm, dm), x3);
    return film(ln(x4), sig, p + ".affine3");
  }
  // TextStyleEncoder.forward (text_style.py:91-104); the Dropout(0.3) on the style vectors is the caller's keep mask
  Ten text_style(const Ten& sig) {
    const std::string p = "text_style_model";
    Ten sdrop = make(in_style.rows, in_style.C, in_style.period, false);
    sdrop.ng = true;
    {
      const Ten a = in_style, m = style_keep; const bool* hk = &have_keep;
      rec([=](Launcher& L) { L.run(a.n(), MulFwd{a.v, *hk ? m.v : nullptr, sdrop.v}); }, [=](Launcher&) {});
    }
    Ten s = sdrop; s.rows = B * 70; s.C = 256; s.period = 70;              // reshape_up(style, 5): a plain view (utils/nn.py:115-127)
    s = film(ln(ffn(s, p + ".style_ffn", 4 * c2, d)), sig, p + ".affine1");
    Ten t = make(B * L, d, L);
    {
      const int64_t* ids = text; const float* E = P(p + ".emb.weight"); float* gE = G(p + ".emb.weight"); const int dd = d;
      rec([=](Launcher& L_) { L_.run(t.n(), EmbedFwd{ids, E, t.v, dd}); }, [=](Launcher& L_) { L_.run(t.n(), EmbedBwd{ids, t.g, gE, dd}); });
    }
    t = film(ln(t), sig, p + ".affine2");
    Ten m = mha(p + ".mha", t, s, s, 8, nullptr);                          // no mask here
    t = film(ln(add(t, m)), sig, p + ".affine3");
    return film(ln(ffn(t, p + ".text_ffn", 2 * d, d)), sig, p + ".affine4");   // no residual
  }
  // DiffusionModel.forward (model.py:121-182)
  void build_model() {
    tape.clear();
    v_used = g_used = 0;
    in_x = make(B * T, 2, T, false);
    in_sigma = make(B, 1, 1, false);
    in_style = make(B * 14, 1280, 14, false);
    style_keep = make(B * 14, 1280, 14, false);
    in_x.ng = in_sigma.ng = in_style.ng = style_keep.ng = true;
    Ten sig = ffn(in_sigma, "sigma_ffn", 2048, c1 / 4);
    Ten text_t = text_style(sig);
    Ten x = linear(in_x, "input_dense", c1);
    Ten h1 = conv_block("enc1", x, sig, c1);
    Ten sk1 = conv3(h1, "skip_conv1", c2, true);   // model.py:169-181 adds these in the decoder; they only need h_k
    Ten h2 = conv_block("enc2", pool(h1), sig, c2);
    h2 = encoder_layer("enc3", h2, text_t, sig, 3, 4.f);
    Ten sk2 = conv3(h2, "skip_conv2", c3, true);
    Ten h3 = conv_block("enc4", pool(h2), sig, c3);
    h3 = encoder_layer("enc5", h3, text_t, sig, 4, 2.f);
    Ten sk3 = conv3(h3, "skip_conv3", d, true);
    x = linear(pool(h3), "att_dense", d);
    for (int i = 0; i < num_layers; ++i) x = encoder_layer("att_layers." + std::to_string(i), x, text_t, sig, 6, 1.f);
    x = conv_block("dec3", add(up(x), sk3, true), sig, c3);
    x = conv_block("dec2", add(up(x), sk2, true), sig, c2);
    x = conv_block("dec1", add(up(x), sk1, true), sig, c1);
    out_score = linear(x, "output_dense", 2);
    out_pen = sigmoid(linear(x, "pen_lifts_dense.0", 1));
  }
};

extern "C" {

const char* dhg_trainer_last_error(void) { return g_serr; }

int64_t dhg_trainer_param_count(int32_t num_layers, int32_t channels) {
  if (num_layers < 0 || channels < 8 || channels % 8) return -1;
  Layout l;
  l.build(num_layers, channels);
  return l.total;
}

int32_t dhg_trainer_param_info(int32_t num_layers, int32_t channels, int32_t index, char* name_out, int32_t name_cap, int64_t* offset,
                               int64_t* numel) {
  if (num_layers < 0 || channels < 8 || channels % 8) return sfail("dhg_trainer_param_info: bad model size");
  Layout l;
  l.build(num_layers, channels);
  if (index < 0 || index >= (int)l.items.size()) return -1;   // past the end: not an error, the caller's loop stops here
  if (name_out && name_cap > 0) snprintf(name_out, (size_t)name_cap, "%s", l.items[index].name.c_str());
  if (offset) *offset = l.items[index].off;
  if (numel) *numel = l.items[index].numel;
  return 0;
}

int32_t dhg_trainer_create(int32_t device, int32_t num_layers, int32_t channels, int32_t B, int32_t T, int32_t L, float* dev_params,
                           float* dev_grads, dhg_trainer** out) {
  if (!out) return sfail("dhg_trainer_create: out is NULL");
  *out = nullptr;
  if (!dev_params || !dev_grads) return sfail("dhg_trainer_create: parameter / gradient buffer is NULL");
  if (num_layers < 0 || channels < 8 || channels % 8) return sfail("dhg_trainer_create: bad model size");
  if (B < 1 || L < 1 || T < 8 || T % 8) return sfail("dhg_trainer_create: need B >= 1, L >= 1 and T a positive multiple of 8 (got B=%d T=%d L=%d)", B, T, L);
#ifndef DHG_HOSTSIM
  if (cudaSetDevice(device) != cudaSuccess) return sfail("dhg_trainer_create: cudaSetDevice(%d) failed", device);
#endif
  dhg_trainer* t = new dhg_trainer();
  t->device = device; t->num_layers = num_layers; t->ch = channels; t->B = B; t->T = T; t->L = L;
  t->c1 = channels; t->c2 = channels * 3 / 2; t->c3 = channels * 2; t->d = 2 * t->c2;
  t->lay.build(num_layers, channels);
  t->params = dev_params; t->grads = dev_grads;
  t->dry = true;
  t->build_model();
  if (!t->err.empty()) { sfail("dhg_trainer_create: %s", t->err.c_str()); delete t; return 1; }
  t->v_arena = (float*)dev_alloc(t->v_need * sizeof(float));
  t->g_arena = (float*)dev_alloc(t->g_need * sizeof(float));
  t->text = (int64_t*)dev_alloc((size_t)B * L * sizeof(int64_t));
  if (!t->v_arena || !t->g_arena || !t->text) {
    sfail("dhg_trainer_create: out of device memory (%.1f MB of activations + gradients)", (t->v_need + t->g_need) * 4.0 / 1e6);
    dev_free(t->v_arena); dev_free(t->g_arena); dev_free(t->text); delete t; return 1;
  }
  t->dry = false;
  t->build_model();
#ifndef DHG_HOSTSIM
  if (cudaStreamCreateWithFlags(&t->side_stream, cudaStreamNonBlocking) == cudaSuccess) {
    t->side_events.resize(2 * t->tape.size() + 4);
    for (auto& e : t->side_events)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) t->err = "cudaEventCreate failed";
  } else {
    t->side_stream = nullptr;
  }
#endif
  if (!t->err.empty()) { sfail("dhg_trainer_create: %s", t->err.c_str()); dhg_trainer_destroy(t); return 1; }
  *out = t;
  return 0;
}

int32_t dhg_trainer_destroy(dhg_trainer* t) {
  if (!t) return 0;
#ifndef DHG_HOSTSIM
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
#endif
  dev_free(t->v_arena); dev_free(t->g_arena); dev_free(t->text);
  for (void* p : t->consts) dev_free(p);
#ifndef DHG_HOSTSIM
  for (auto e : t->side_events) if (e) cudaEventDestroy(e);
  if (t->side_stream) cudaStreamDestroy(t->side_stream);
#endif
  delete t;
  return 0;
}

int64_t dhg_trainer_workspace_bytes(const dhg_trainer* t) { return t ? (int64_t)((t->v_need + t->g_need) * sizeof(float)) : 0; }
int64_t dhg_trainer_last_launches(const dhg_trainer* t) { return t ? t->last_launches : 0; }

int32_t dhg_trainer_set_option(const char* name, int32_t value) {
  if (name && !strcmp(name, "tiled_gemm")) { g_use_tiled = value; return 0; }
  if (name && !strcmp(name, "side_stream")) { g_side_stream = value; return 0; }
  return sfail("dhg_trainer_set_option: unknown option");
}

int32_t dhg_trainer_forward(dhg_trainer* t, const float* dev_x, const int64_t* dev_text, const float* dev_sigma, const float* dev_style,
                            const float* dev_style_keep, float* dev_score_pred, float* dev_pen_pred, void* stream) {
  if (!t) return sfail("dhg_trainer_forward: trainer is NULL");
  if (!dev_x || !dev_text || !dev_sigma || !dev_style) return sfail("dhg_trainer_forward: NULL input");
#ifndef DHG_HOSTSIM
  if (cudaSetDevice(t->device) != cudaSuccess) return sfail("dhg_trainer_forward: cudaSetDevice failed");
#endif
  Launcher L; L.st = (ts_stream)stream;
  L.copy(t->in_x.v, dev_x, (size_t)t->B * t->T * 2 * sizeof(float));
  L.copy(t->text, dev_text, (size_t)t->B * t->L * sizeof(int64_t));
  L.copy(t->in_sigma.v, dev_sigma, (size_t)t->B * sizeof(float));
  L.copy(t->in_style.v, dev_style, (size_t)t->B * 14 * 1280 * sizeof(float));
  t->have_keep = dev_style_keep != nullptr;
  if (dev_style_keep) L.copy(t->style_keep.v, dev_style_keep, (size_t)t->B * 14 * 1280 * sizeof(float));
#ifndef DHG_HOSTSIM
  if (g_side_stream && t->side_stream) { L.side = t->side_stream; L.events = &t->side_events; }
#endif
  for (auto& op : t->tape) op.fwd(L);
  L.side_join();
  if (dev_score_pred) L.copy(dev_score_pred, t->out_score.v, (size_t)t->B * t->T * 2 * sizeof(float));
  if (dev_pen_pred) L.copy(dev_pen_pred, t->out_pen.v, (size_t)t->B * t->T * sizeof(float));
  t->last_launches = L.launches;
#ifndef DHG_HOSTSIM
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sfail("dhg_trainer_forward: %s", cudaGetErrorString(e));
#endif
  return 0;
}

int32_t dhg_trainer_backward(dhg_trainer* t, const float* dev_grad_score, const float* dev_grad_pen_pred, void* stream) {
  if (!t) return sfail("dhg_trainer_backward: trainer is NULL");
  if (!dev_grad_score || !dev_grad_pen_pred) return sfail("dhg_trainer_backward: NULL gradient");
#ifndef DHG_HOSTSIM
  if (cudaSetDevice(t->device) != cudaSuccess) return sfail("dhg_trainer_backward: cudaSetDevice failed");
#endif
  Launcher L; L.st = (ts_stream)stream;
  L.zero(t->g_arena, t->g_need * sizeof(float));
  L.zero(t->grads, (size_t)t->lay.total * sizeof(float));
  L.copy(t->out_score.g, dev_grad_score, (size_t)t->B * t->T * 2 * sizeof(float));
  L.copy(t->out_pen.g, dev_grad_pen_pred, (size_t)t->B * t->T * sizeof(float));
#ifndef DHG_HOSTSIM
  if (g_side_stream && t->side_stream) { L.side = t->side_stream; L.events = &t->side_events; }
#endif
  for (auto it = t->tape.rbegin(); it != t->tape.rend(); ++it) it->bwd(L);
  L.side_join();
  t->last_launches += L.launches;
#ifndef DHG_HOSTSIM
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sfail("dhg_trainer_backward: %s", cudaGetErrorString(e));
#endif
  return 0;
}

}  // extern "C"
