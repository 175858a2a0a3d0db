// Shared types for the sm_100a reverse-diffusion engine.
//
// Layout convention (DESIGN.md "Data layout in HBM"): every activation is a
// channels-last row matrix [rows, C] with C contiguous.  Stroke-level tensors at
// pyramid level l (T_l = T >> l) use the PADDED ROW layout
//     row(b, t) = b * (T_l + 1) + 1 + t,      row(b, -1) = b * (T_l + 1) is all-zero,
// plus one trailing zero row, so a k=3 'same' conv1d (reference cnn.py:32-47) is
// three row-shifted GEMMs over the flat matrix with the zero halo already in
// memory.  Text rows ([B*L, C]) and style rows ([B*70, C]) carry no padding.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dhg {

typedef __nv_bfloat16 bf16;

enum Precision { PREC_FP32 = 0, PREC_BF16 = 1 };

// Split storage of an fp32 value for the fp32-contract mode on tensor cores: x ~ hi + lo with hi = bf16(x),
// lo = bf16(x - hi), 16 mantissa bits together (relative error <= 2^-17).  Rows are laid out in GROUPS of 32 elements:
// 128 bytes = the 32 hi halves (64 B) followed by the 32 lo halves (64 B).  A row of C elements is therefore a bf16 row
// of 2C numbers in which every 64-wide k-block of the tensor-core GEMM holds the hi and the lo parts of 32 elements, and
// with the weights stored the same way (w_hi x 32 | w_lo x 32 per group) one A tile and one W tile per k-block give
//   x . w = hi . w_hi + lo . w_hi + hi . w_lo      (+ lo . w_lo, dropped: 2^-18)
// as three pairs of tcgen05.mma k-steps into the same fp32 accumulator (gemm_tc.cu, mma_kblock).
// `bfs` is the 4-byte element TAG of such a row for the CUDA-core kernels: pointer arithmetic in elements works as for
// float, but the element is never dereferenced -- the accessors below find the two halves from the ADDRESS (row bases
// are 128-byte aligned: cudaMalloc bases, C a multiple of 32, column offsets in multiples of 4 within a group).
struct __align__(4) bfs {
  uint32_t tag;
};
__device__ __forceinline__ const char* bfs_hi_ptr(const void* p) {   // address of the hi half of element p; lo is 64 bytes on
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  return reinterpret_cast<const char*>((a & ~(uintptr_t)127) + ((a & 127) >> 1));
}
__device__ __forceinline__ float bf16_bits_to_f(uint32_t h) { return __uint_as_float(h << 16); }
__device__ __forceinline__ uint32_t bf16_pair_hi(float x0, float x1, float& r0, float& r1) {   // packs bf16(x), returns the remainders
  const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
  r0 = x0 - __bfloat162float(h0);
  r1 = x1 - __bfloat162float(h1);
  return (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
}
__device__ __forceinline__ uint32_t bf16_pair(float x0, float x1) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x0)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x1)) << 16);
}
// two packed bf16 (low, high half-word) of the hi word + of the lo word -> the two fp32 values
__device__ __forceinline__ void split_unpack2(uint32_t wh, uint32_t wl, float& a, float& b) {
  a = __uint_as_float(wh << 16) + __uint_as_float(wl << 16);
  b = __uint_as_float(wh & 0xffff0000u) + __uint_as_float(wl & 0xffff0000u);
}

// How a flat row index maps to (sample, position): rows are grouped in periods
// of `period` rows per sample; if `pad_first`, the first row of each period is
// the zero halo row.  Rows >= nvalid (= B * period) are trailing padding.
struct RowMap {
  int period;
  int pad_first;
  int nvalid;
};

// Epilogue applied to one GEMM output row (see rowpost_kernel / the tcgen05
// epilogue).  Order: acc + bias|rowbias -> + res_pre -> LayerNorm -> FiLM ->
// + res_post -> zero if halo row -> store raw and/or SiLU'd copy.
struct Epilogue {
  const float* bias;       // [N] or null
  const float* rowbias;    // [period - pad_first, N] indexed by position (bias folded in) or null  (CUDA-core path)
  const void* rowbias16;   // bf16 [period - pad_first, rowbias16_cols]: per-position term added on top of `bias`
  int rowbias16_cols;      //   for output columns < rowbias16_cols (tcgen05 path), or null
  const void* res_pre;     // activation dtype, same row index, or null
  int res_pre_pitch;
  int ln;                  // LayerNorm over the N outputs, eps 1e-6, no affine
  const float* gamma;      // FiLM: gamma[b * film_bstride + n]; null = no FiLM
  const float* beta;
  int film_bstride;        // 0 = one vector shared by the whole batch (sampling)
  int film_planned;        // plan-time flag: gamma/beta will be supplied at launch
  const void* res_post;    // activation dtype or null
  int res_post_pitch;
  int res_post_up;         // 1: res_post lives one pyramid level down: row b*(period_lo)+1+pos/2
  int res_post_period_lo;
  void* out_raw;           // activation dtype or null
  int out_raw_pitch;
  void* out_act;           // SiLU(value), activation dtype or null
  int out_act_pitch;
  // dot mode (tcgen05 path, instead of out_raw / out_act): the row is not stored; its 3 dot products with dot_w[0..2]
  // (fp32 [3, N]) go to dot_out[row * 4 + j] (fp32).  dot_act: the dots are taken of SiLU(value).  Used where the only
  // reader of the row is a linear map onto 3 channels (the heads behind the last ConvBlock, engine.cu tail fusion).
  const float* dot_w;
  float* dot_out;
  int dot_act;
  int dot_planned;         // plan-time flag like film_planned: dot_w is supplied at launch
  int w_row_off;           // tcgen05 path, launch time: first row of the weight variant to use (TcDual stacks; 0 otherwise)
  int split_io;            // tcgen05 path: every activation operand (A, residuals, per-position rows, outputs) is `bfs`
                           // split storage (groups of 32 hi | 32 lo); pitches stay in elements.  The fp32-contract mode.
  RowMap map;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// the same through a correctly rounded reciprocal instead of a full division (1 ulp apart; half the instructions):
// the tcgen05 epilogue of the fp32-contract mode
__device__ __forceinline__ float silu_rcp(float x) { return x * __frcp_rn(1.0f + __expf(-x)); }
// SiLU of a value that is stored in T right away: for bf16 the one-MUFU form h + h * tanh.approx(h), h = x / 2
// (2^-11 relative, far below the bf16 rounding; 3 instructions instead of the ~15 of an fp32 division); fp32 keeps
// the exact form (the fp32 mode is the parity mode).
template <typename T> __device__ __forceinline__ float silu_out(float x) { return silu_f(x); }
template <> __device__ __forceinline__ float silu_out<__nv_bfloat16>(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive activation elements <-> float4
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <> __device__ __forceinline__ float4 load4<bfs>(const bfs* p) {
  const char* h = bfs_hi_ptr(p);
  const uint2 uh = *reinterpret_cast<const uint2*>(h), ul = *reinterpret_cast<const uint2*>(h + 64);
  float4 v;
  split_unpack2(uh.x, ul.x, v.x, v.y);
  split_unpack2(uh.y, ul.y, v.z, v.w);
  return v;
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<bfs>(bfs* p, float4 v) {
  char* h = const_cast<char*>(bfs_hi_ptr(p));
  float r0, r1, r2, r3;
  uint2 uh, ul;
  uh.x = bf16_pair_hi(v.x, v.y, r0, r1);
  uh.y = bf16_pair_hi(v.z, v.w, r2, r3);
  ul.x = bf16_pair(r0, r1);
  ul.y = bf16_pair(r2, r3);
  *reinterpret_cast<uint2*>(h) = uh;
  *reinterpret_cast<uint2*>(h + 64) = ul;
}
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// 8 consecutive activation elements <-> float[8] (one 16-byte access for bf16)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* v);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <> __device__ __forceinline__ void load8<bfs>(const bfs* p, float* v) {
  const char* h = bfs_hi_ptr(p);
  const uint4 uh = *reinterpret_cast<const uint4*>(h), ul = *reinterpret_cast<const uint4*>(h + 64);
  split_unpack2(uh.x, ul.x, v[0], v[1]);
  split_unpack2(uh.y, ul.y, v[2], v[3]);
  split_unpack2(uh.z, ul.z, v[4], v[5]);
  split_unpack2(uh.w, ul.w, v[6], v[7]);
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* v);
template <> __device__ __forceinline__ void store8<bfs>(bfs* p, const float* v) {
  char* h = const_cast<char*>(bfs_hi_ptr(p));
  float r[8];
  uint4 uh, ul;
  uh.x = bf16_pair_hi(v[0], v[1], r[0], r[1]);
  uh.y = bf16_pair_hi(v[2], v[3], r[2], r[3]);
  uh.z = bf16_pair_hi(v[4], v[5], r[4], r[5]);
  uh.w = bf16_pair_hi(v[6], v[7], r[6], r[7]);
  ul.x = bf16_pair(r[0], r[1]);
  ul.y = bf16_pair(r[2], r[3]);
  ul.z = bf16_pair(r[4], r[5]);
  ul.w = bf16_pair(r[6], r[7]);
  *reinterpret_cast<uint4*>(h) = uh;
  *reinterpret_cast<uint4*>(h + 64) = ul;
}
template <> __device__ __forceinline__ void store8<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace dhg
