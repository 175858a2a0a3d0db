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
// lo = bf16(x - hi), 16 mantissa bits together (relative error <= 2^-17).  A row of C such elements IS a bf16 row of
// 2C elements (hi_0 lo_0 hi_1 lo_1 ..), so it feeds the bf16 tensor-core GEMM directly: with the weights split the
// same way, x . w = (hi + lo) . w_hi + hi . w_lo (+ lo . w_lo, dropped: 2^-18) is two bf16 GEMMs over that row, one
// against (w_hi, w_hi)-interleaved and one against (w_lo, 0)-interleaved weights, accumulated in fp32 (gemm_tc.cu).
struct __align__(4) bfs {
  __nv_bfloat16 hi, lo;
};

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
                           // split storage; pitches stay in elements.  The fp32-contract mode (engine.cu).
  RowMap map;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
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

__device__ __forceinline__ uint32_t split_pack(float x) {   // -> hi in the low half-word, lo in the high one
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  return (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
}
__device__ __forceinline__ float split_unpack(uint32_t w) {
  return __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<bfs>(bfs v) { return __bfloat162float(v.hi) + __bfloat162float(v.lo); }
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ bfs from_f<bfs>(float v) {
  bfs r;
  r.hi = __float2bfloat16_rn(v);
  r.lo = __float2bfloat16_rn(v - __bfloat162float(r.hi));
  return r;
}
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
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  return make_float4(split_unpack(u.x), split_unpack(u.y), split_unpack(u.z), split_unpack(u.w));
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<bfs>(bfs* p, float4 v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(split_pack(v.x), split_pack(v.y), split_pack(v.z), split_pack(v.w));
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
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  v[0] = split_unpack(a.x); v[1] = split_unpack(a.y); v[2] = split_unpack(a.z); v[3] = split_unpack(a.w);
  v[4] = split_unpack(b.x); v[5] = split_unpack(b.y); v[6] = split_unpack(b.z); v[7] = split_unpack(b.w);
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* v);
template <> __device__ __forceinline__ void store8<bfs>(bfs* p, const float* v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(split_pack(v[0]), split_pack(v[1]), split_pack(v[2]), split_pack(v[3]));
  *reinterpret_cast<uint4*>(p + 4) = make_uint4(split_pack(v[4]), split_pack(v[5]), split_pack(v[6]), split_pack(v[7]));
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
