// CUDA-core kernels of the engine: the exact-fp32 GEMM used by the fp32
// precision mode (and at plan time for the PE-folded bias tables), the row-wise
// epilogue kernel, fp32 attention, and the small elementwise / gather kernels
// that sit between GEMMs.  All sm_100a, no library calls.
#include "kernels.h"

namespace dhg {

// ---------------------------------------------------------------------------
// fp32-accumulate tiled GEMM with row-shifted taps:
//   acc[m, n] = sum_{tap<taps} sum_{k<K} A[m + tap - taps/2, k] * W[tap][k][n]
// A: [rows, K] (pitch lda), rows outside [0, rows) read as zero (the conv halo
// between samples is a real zero row in memory, see common.cuh).
// W: fp32 [taps][K][N], N contiguous.  Output: fp32 scratch [rows, N].
// Tile 128x64x16, 256 threads, 8x4 outputs per thread, register prefetch.
// ---------------------------------------------------------------------------
constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;

template <typename TA>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, int lda, int rows,
                                                        const float* __restrict__ W, int K, int N,
                                                        int taps, float* __restrict__ out) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;
  // A loader: row tid/2, 8 consecutive k at (tid&1)*8
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  // B loader: k = tid/16, n = (tid&15)*4
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  const int ksteps = K / SG_BK;
  const int total = taps * ksteps;
  const int shift0 = -(taps / 2);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra0, ra1, rb;
  auto gload = [&](int it) {
    const int tap = it / ksteps, k0 = (it - tap * ksteps) * SG_BK;
    const int r = m0 + a_row + tap + shift0;
    if (r >= 0 && r < rows) {
      const TA* p = A + (size_t)r * lda + k0 + a_k;
      ra0 = load4<TA>(p);
      ra1 = load4<TA>(p + 4);
    } else {
      ra0 = make_float4(0.f, 0.f, 0.f, 0.f);
      ra1 = ra0;
    }
    const int n = n0 + b_n;
    if (n < N)
      rb = *reinterpret_cast<const float4*>(W + ((size_t)tap * K + k0 + b_k) * N + n);
    else
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
    As[buf][a_k + 0][a_row] = ra0.x;
    As[buf][a_k + 1][a_row] = ra0.y;
    As[buf][a_k + 2][a_row] = ra0.z;
    As[buf][a_k + 3][a_row] = ra0.w;
    As[buf][a_k + 4][a_row] = ra1.x;
    As[buf][a_k + 5][a_row] = ra1.y;
    As[buf][a_k + 6][a_row] = ra1.z;
    As[buf][a_k + 7][a_row] = ra1.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) gload(it + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (it + 1 < total) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int n = n0 + tx * 4;
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + ty * 8 + i;
      if (r < rows)
        *reinterpret_cast<float4*>(out + (size_t)r * N + n) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

template <typename TA>
void launch_gemm_simt(const TA* A, int lda, int rows, const float* W, int K, int N, int taps,
                      float* out, cudaStream_t st) {
  dim3 grid((rows + SG_BM - 1) / SG_BM, (N + SG_BN - 1) / SG_BN);
  gemm_simt_kernel<TA><<<grid, 256, 0, st>>>(A, lda, rows, W, K, N, taps, out);
}
template void launch_gemm_simt<float>(const float*, int, int, const float*, int, int, int, float*,
                                      cudaStream_t);
template void launch_gemm_simt<bf16>(const bf16*, int, int, const float*, int, int, int, float*,
                                     cudaStream_t);

// ---------------------------------------------------------------------------
// Row-wise epilogue over an fp32 accumulator matrix: one warp per row.
// ---------------------------------------------------------------------------
constexpr int RP_MAXCHUNK = 9;  // N <= 1152

template <typename T>
__global__ void __launch_bounds__(256) rowpost_kernel(const float* __restrict__ acc, int rows, int N,
                                                      Epilogue e) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int r = warp;
  const int b = r / e.map.period;
  const int j = r - b * e.map.period;
  const bool is_pad = (r >= e.map.nvalid) || (e.map.pad_first && j == 0);
  const int pos = j - e.map.pad_first;
  T* oraw = e.out_raw ? reinterpret_cast<T*>(e.out_raw) + (size_t)r * e.out_raw_pitch : nullptr;
  T* oact = e.out_act ? reinterpret_cast<T*>(e.out_act) + (size_t)r * e.out_act_pitch : nullptr;
  const int nchunk = (N + 127) >> 7;
  if (is_pad) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < nchunk; ++c) {
      const int n = c * 128 + lane * 4;
      if (n < N) {
        if (oraw) store4<T>(oraw + n, z);
        if (oact) store4<T>(oact + n, z);
      }
    }
    return;
  }
  float4 v[RP_MAXCHUNK];
  const float* arow = acc + (size_t)r * N;
  const float* bias = e.rowbias ? e.rowbias + (size_t)pos * N : e.bias;
  const T* rpre = e.res_pre ? reinterpret_cast<const T*>(e.res_pre) + (size_t)r * e.res_pre_pitch : nullptr;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < RP_MAXCHUNK; ++c) {
    const int n = c * 128 + lane * 4;
    if (c < nchunk && n < N) {
      float4 x = *reinterpret_cast<const float4*>(arow + n);
      if (bias) {
        const float4 bb = *reinterpret_cast<const float4*>(bias + n);
        x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
      }
      if (rpre) {
        const float4 rr = load4<T>(rpre + n);
        x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
      }
      v[c] = x;
      s += (x.x + x.y) + (x.z + x.w);
    } else {
      v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (e.ln) {
    mean = warp_sum(s) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < RP_MAXCHUNK; ++c) {
      const int n = c * 128 + lane * 4;
      if (c < nchunk && n < N) {
        const float dx = v[c].x - mean, dy = v[c].y - mean, dz = v[c].z - mean, dw = v[c].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    rstd = rsqrtf(warp_sum(q) / (float)N + 1e-6f);
  }
  const float* gam = e.gamma ? e.gamma + (size_t)b * e.film_bstride : nullptr;
  const float* bet = e.gamma ? e.beta + (size_t)b * e.film_bstride : nullptr;
  const T* rpost = nullptr;
  if (e.res_post) {
    const size_t rr = e.res_post_up ? (size_t)b * e.res_post_period_lo + 1 + (pos >> 1) : (size_t)r;
    rpost = reinterpret_cast<const T*>(e.res_post) + rr * e.res_post_pitch;
  }
#pragma unroll
  for (int c = 0; c < RP_MAXCHUNK; ++c) {
    const int n = c * 128 + lane * 4;
    if (c < nchunk && n < N) {
      float4 x = v[c];
      if (e.ln) {
        x.x = (x.x - mean) * rstd; x.y = (x.y - mean) * rstd;
        x.z = (x.z - mean) * rstd; x.w = (x.w - mean) * rstd;
      }
      if (gam) {
        const float4 g = *reinterpret_cast<const float4*>(gam + n);
        const float4 bb = *reinterpret_cast<const float4*>(bet + n);
        x.x = fmaf(x.x, g.x, bb.x); x.y = fmaf(x.y, g.y, bb.y);
        x.z = fmaf(x.z, g.z, bb.z); x.w = fmaf(x.w, g.w, bb.w);
      }
      if (rpost) {
        const float4 rr = load4<T>(rpost + n);
        x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
      }
      if (oraw) store4<T>(oraw + n, x);
      if (oact) store4<T>(oact + n, make_float4(silu_f(x.x), silu_f(x.y), silu_f(x.z), silu_f(x.w)));
    }
  }
}

template <typename T>
void launch_rowpost(const float* acc, int rows, int N, const Epilogue& e, cudaStream_t st) {
  const int wpb = 8;
  rowpost_kernel<T><<<(rows + wpb - 1) / wpb, wpb * 32, 0, st>>>(acc, rows, N, e);
}
template void launch_rowpost<float>(const float*, int, int, const Epilogue&, cudaStream_t);
template void launch_rowpost<bf16>(const float*, int, int, const Epilogue&, cudaStream_t);

// ---------------------------------------------------------------------------
// fp32 attention, one thread per query row, keys/values streamed through shared
// memory in tiles of 64 with an online softmax.  Reference: attention.py:26-46
// (SDPA with additive -1e9 mask) and :78-85 (head split by column slices).
// ---------------------------------------------------------------------------
constexpr int AT_TK = 64;

template <typename T, int D>
__global__ void __launch_bounds__(128) attention_simt_kernel(AttnParams p) {
  __shared__ __align__(16) float Ks[AT_TK][D];
  __shared__ __align__(16) float Vs[AT_TK][D];
  __shared__ float Ms[AT_TK];
  const int b = blockIdx.z, h = blockIdx.y;
  const int tq = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = tq < p.Tq;
  const T* Q = reinterpret_cast<const T*>(p.q);
  const T* Kp = reinterpret_cast<const T*>(p.k);
  const T* Vp = reinterpret_cast<const T*>(p.v);
  float q[D], o[D];
  if (active) {
    const T* qr = Q + ((size_t)b * p.q_period + p.q_pad + tq) * p.q_pitch + h * D;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      const float4 x = load4<T>(qr + d);
      q[d] = x.x * p.scale; q[d + 1] = x.y * p.scale; q[d + 2] = x.z * p.scale; q[d + 3] = x.w * p.scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;

  for (int k0 = 0; k0 < p.Tk; k0 += AT_TK) {
    const int nk = min(AT_TK, p.Tk - k0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nk * (D / 4); idx += blockDim.x) {
      const int j = idx / (D / 4), d = (idx - j * (D / 4)) * 4;
      const size_t row = (size_t)b * p.k_period + p.k_pad + k0 + j;
      *reinterpret_cast<float4*>(&Ks[j][d]) = load4<T>(Kp + row * p.k_pitch + h * D + d);
      *reinterpret_cast<float4*>(&Vs[j][d]) = load4<T>(Vp + row * p.v_pitch + h * D + d);
    }
    for (int j = threadIdx.x; j < nk; j += blockDim.x)
      Ms[j] = (p.text && p.text[(size_t)b * p.Tk + k0 + j] == 0) ? -1e9f : 0.f;
    __syncthreads();
    if (!active) continue;
    for (int j0 = 0; j0 < nk; j0 += 8) {
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        if (j < nk) {
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < D; d += 4) {
            const float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
            a = fmaf(q[d], kk.x, a); a = fmaf(q[d + 1], kk.y, a);
            a = fmaf(q[d + 2], kk.z, a); a = fmaf(q[d + 3], kk.w, a);
          }
          s[jj] = a + Ms[j];
          cmax = fmaxf(cmax, s[jj]);
        } else {
          s[jj] = -INFINITY;
        }
      }
      const float mnew = fmaxf(mrun, cmax);
      const float corr = __expf(mrun - mnew);  // exp(-inf) = 0 on the first chunk
      lrun *= corr;
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        if (j < nk) {
          const float pj = __expf(s[jj] - mnew);
          lrun += pj;
#pragma unroll
          for (int d = 0; d < D; d += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
            o[d] = fmaf(pj, vv.x, o[d]); o[d + 1] = fmaf(pj, vv.y, o[d + 1]);
            o[d + 2] = fmaf(pj, vv.z, o[d + 2]); o[d + 3] = fmaf(pj, vv.w, o[d + 3]);
          }
        }
      }
      mrun = mnew;
    }
  }
  if (active) {
    const float inv = 1.f / lrun;
    T* orow = reinterpret_cast<T*>(p.o) + ((size_t)b * p.q_period + p.q_pad + tq) * p.o_pitch + h * D;
#pragma unroll
    for (int d = 0; d < D; d += 4)
      store4<T>(orow + d, make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv));
  }
}

template <typename T>
int launch_attention_simt(const AttnParams& p, cudaStream_t st) {
  const int threads = p.Tq <= 32 ? 32 : (p.Tq <= 64 ? 64 : 128);
  dim3 grid((p.Tq + threads - 1) / threads, p.H, p.B);
  if (p.D == 64)
    attention_simt_kernel<T, 64><<<grid, threads, 0, st>>>(p);
  else if (p.D == 48)
    attention_simt_kernel<T, 48><<<grid, threads, 0, st>>>(p);
  else
    return 1;
  return 0;
}
template int launch_attention_simt<float>(const AttnParams&, cudaStream_t);
template int launch_attention_simt<bf16>(const AttnParams&, cudaStream_t);
template int launch_attention_simt<bfs>(const AttnParams&, cudaStream_t);

// ---------------------------------------------------------------------------
// FiLM over rows: out[r, c] = in[r, c] * gamma[b, c] + beta[b, c]   (conditioning.py:19)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) film_rows_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int C,
                                                        int period, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int bstride) {
  // 8 channels (16 bytes of bf16) per thread, grid-stride, 32-bit index arithmetic (rows * C / 8 < 2^31 by the plan limit)
  const uint32_t c8 = (uint32_t)C >> 3, total = (uint32_t)rows * c8;
  for (uint32_t i8 = blockIdx.x * blockDim.x + threadIdx.x; i8 < total; i8 += gridDim.x * blockDim.x) {
    const uint32_t r = i8 / c8, c = (i8 - r * c8) * 8;
    const uint32_t b = bstride ? r / (uint32_t)period : 0u;
    float x[8];
    load8<T>(in + (size_t)r * C + c, x);
    const float4* g = reinterpret_cast<const float4*>(gamma + (size_t)b * bstride + c);
    const float4* bb = reinterpret_cast<const float4*>(beta + (size_t)b * bstride + c);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 gq = g[q], bq = bb[q];
      x[q * 4 + 0] = fmaf(x[q * 4 + 0], gq.x, bq.x); x[q * 4 + 1] = fmaf(x[q * 4 + 1], gq.y, bq.y);
      x[q * 4 + 2] = fmaf(x[q * 4 + 2], gq.z, bq.z); x[q * 4 + 3] = fmaf(x[q * 4 + 3], gq.w, bq.w);
    }
    store8<T>(out + (size_t)r * C + c, x);
  }
}
template <typename T>
void launch_film_rows(const T* in, T* out, int rows, int C, int period, const float* gamma,
                      const float* beta, int bstride, cudaStream_t st) {
  const size_t n8 = (size_t)rows * (C >> 3);
  size_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  film_rows_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(in, out, rows, C, period, gamma, beta, bstride);
}
template void launch_film_rows<float>(const float*, float*, int, int, int, const float*, const float*, int, cudaStream_t);
template void launch_film_rows<bf16>(const bf16*, bf16*, int, int, int, const float*, const float*, int, cudaStream_t);
template void launch_film_rows<bfs>(const bfs*, bfs*, int, int, int, const float*, const float*, int, cudaStream_t);

// ---------------------------------------------------------------------------
// AvgPool1d(2) over T in the padded row layout (model.py:93): level l -> l+1.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_kernel(const T* __restrict__ in, T* __restrict__ out_raw, T* __restrict__ out_act,
                                                   int B, int Tlo, int C) {
  // 32-bit index arithmetic: B * Tlo * C / 8 < 2^31 by the plan limit (B * (T + 1) < 2^24, C <= 384)
  const uint32_t c8 = (uint32_t)C >> 3;
  const uint32_t total = (uint32_t)B * (uint32_t)Tlo * c8;
  for (uint32_t i8 = blockIdx.x * blockDim.x + threadIdx.x; i8 < total; i8 += gridDim.x * blockDim.x) {
    const uint32_t bt = i8 / c8, c = (i8 - bt * c8) * 8;
    const uint32_t b = bt / (uint32_t)Tlo, t = bt - b * (uint32_t)Tlo;
    const size_t rin = (size_t)b * (2 * Tlo + 1) + 1 + 2 * t;
    const size_t rout = (size_t)b * (Tlo + 1) + 1 + t;
    float x0[8], x1[8], m[8];
    load8<T>(in + rin * C + c, x0);
    load8<T>(in + (rin + 1) * C + c, x1);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = 0.5f * (x0[k] + x1[k]);
    if (out_raw) store8<T>(out_raw + rout * C + c, m);
    if (out_act) {
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = silu_out<T>(m[k]);
      store8<T>(out_act + rout * C + c, m);
    }
  }
}
template <typename T>
void launch_pool(const T* in, T* out_raw, T* out_act, int B, int Tlo, int C, cudaStream_t st) {
  const size_t n8 = (size_t)B * Tlo * (C >> 3);
  size_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pool_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(in, out_raw, out_act, B, Tlo, C);
}
template void launch_pool<float>(const float*, float*, float*, int, int, int, cudaStream_t);
template void launch_pool<bf16>(const bf16*, bf16*, bf16*, int, int, int, cudaStream_t);
template void launch_pool<bfs>(const bfs*, bfs*, bfs*, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------
// input_dense: Linear(2, C) on the fp32 strokes (model.py:82,139) -> padded rows,
// raw + SiLU copies.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void input_dense_kernel(const float* __restrict__ x, const float* __restrict__ W /*[C,2]*/,
                                   const float* __restrict__ bias, T* __restrict__ out_raw,
                                   T* __restrict__ out_act, int B, int Tn, int C) {
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = C >> 2;
  if (i4 >= (size_t)B * Tn * c4) return;
  const int c = (int)(i4 % c4) * 4;
  const size_t bt = i4 / c4;
  const int t = (int)(bt % Tn), b = (int)(bt / Tn);
  const float2 s = *reinterpret_cast<const float2*>(x + bt * 2);
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = fmaf(s.y, W[(c + i) * 2 + 1], fmaf(s.x, W[(c + i) * 2], bias[c + i]));
  const size_t r = (size_t)b * (Tn + 1) + 1 + t;
  store4<T>(out_raw + r * C + c, make_float4(v[0], v[1], v[2], v[3]));
  store4<T>(out_act + r * C + c, make_float4(silu_out<T>(v[0]), silu_out<T>(v[1]), silu_out<T>(v[2]), silu_out<T>(v[3])));
}
template <typename T>
void launch_input_dense(const float* x, const float* W, const float* bias, T* out_raw, T* out_act,
                        int B, int Tn, int C, cudaStream_t st) {
  const size_t n4 = (size_t)B * Tn * (C >> 2);
  input_dense_kernel<T><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, W, bias, out_raw, out_act, B, Tn, C);
}
template void launch_input_dense<float>(const float*, const float*, const float*, float*, float*, int, int, int, cudaStream_t);
template void launch_input_dense<bf16>(const float*, const float*, const float*, bf16*, bf16*, int, int, int, cudaStream_t);
template void launch_input_dense<bfs>(const float*, const float*, const float*, bfs*, bfs*, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------
// Head fusion: enc1.conv_skip(input_dense(x)) straight from x.  input_dense is linear (model.py:139) and conv_skip
// sees its raw output (cnn.py:66), so
//   skip[t] = sum over the taps tau in {-1,0,1} with 0 <= t+tau < T of (x[t+tau] . M_tau + v_tau) + b_skip,
//   M_tau = W_in^T W_skip,tau (2 x C),  v_tau = b_in . W_skip,tau   (tables from dhg_finalize, fp32)
// (positions outside the line are zero rows of the padded layout, NOT input_dense(0), hence the per-tap v_tau).
// 16 lanes per stroke point, 8 channels per lane held in registers, grid-stride; writes the padded-row [rows, C] matrix.
// ---------------------------------------------------------------------------
constexpr int kSkipPts = 512;   // stroke points per block
template <typename T_>
__global__ void __launch_bounds__(256) skip_from_x_kernel(const float* __restrict__ x, const float* __restrict__ M /*[3][2][C]*/,
                                                          const float* __restrict__ v /*[3][C]*/, const float* __restrict__ bsk,
                                                          T_* __restrict__ out, int B, int Tn, int C) {
  // A block owns kSkipPts consecutive points: their x (plus one neighbour on each side) is staged in shared memory with
  // one coalesced burst, after that the block only computes and stores (the output is the whole traffic: C bf16 per point).
  __shared__ float2 xs[kSkipPts + 2];
  const uint32_t npts = (uint32_t)B * (uint32_t)Tn, T = (uint32_t)Tn;
  const uint32_t p0 = blockIdx.x * kSkipPts;
  for (uint32_t j = threadIdx.x; j < kSkipPts + 2; j += blockDim.x) {
    const int64_t i = (int64_t)p0 + j - 1;
    xs[j] = (i >= 0 && i < (int64_t)npts) ? *reinterpret_cast<const float2*>(x + (size_t)i * 2) : make_float2(0.f, 0.f);
  }
  const int sub = threadIdx.x & 15, pl = threadIdx.x >> 4;   // channel octet, point lane (16 points in flight per pass)
  const int c0 = sub * 8;
  float m[3][2][8], v0[8], v2[8], ball[8];   // ball: bias + all three per-tap constants (the interior case)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v0[k] = v[c0 + k];
    v2[k] = v[2 * C + c0 + k];
    ball[k] = bsk[c0 + k] + v[C + c0 + k] + v0[k] + v2[k];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      m[t][0][k] = M[(t * 2 + 0) * C + c0 + k];
      m[t][1][k] = M[(t * 2 + 1) * C + c0 + k];
    }
  }
  __syncthreads();
  uint32_t i = p0 + pl;
  uint32_t b = i / T, t = i - b * T;   // one division per thread; then t advances by 16 per pass
  for (int j = pl; j < kSkipPts && i < npts; j += 16, i += 16) {
    const bool has_l = t > 0, has_r = t + 1 < T;
    const float2 xc = xs[j + 1];
    const float2 xl = has_l ? xs[j] : make_float2(0.f, 0.f), xr = has_r ? xs[j + 2] : make_float2(0.f, 0.f);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = ball[k];
      a = fmaf(xc.x, m[1][0][k], a); a = fmaf(xc.y, m[1][1][k], a);
      a = fmaf(xl.x, m[0][0][k], a); a = fmaf(xl.y, m[0][1][k], a);
      a = fmaf(xr.x, m[2][0][k], a); a = fmaf(xr.y, m[2][1][k], a);
      o[k] = a;
    }
    if (!(has_l && has_r)) {   // first / last position of a line (2 of T): that tap reads a zero row, not input_dense(0)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!has_l) o[k] -= v0[k];
        if (!has_r) o[k] -= v2[k];
      }
    }
    store8<T_>(out + ((size_t)i + b + 1) * C + c0, o);
    t += 16;
    while (t >= T) { t -= T; ++b; }
  }
}
template <typename T>
int launch_skip_from_x(const float* x, const float* M, const float* v, const float* bsk, T* out, int B, int Tn, int C, cudaStream_t st) {
  if (C != 128) return 1;
  const size_t npts = (size_t)B * Tn;
  skip_from_x_kernel<T><<<(unsigned)((npts + kSkipPts - 1) / kSkipPts), 256, 0, st>>>(x, M, v, bsk, out, B, Tn, C);
  return 0;
}
template int launch_skip_from_x<bf16>(const float*, const float*, const float*, const float*, bf16*, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------
// Output heads + fused posterior update (+ fused input_dense of the next step).
// 16 lanes per stroke point (each lane owns 8 of the C = 128 channels), 2 points per warp, grid-stride.
//   eps = Linear(C,2)(h), pen = sigmoid(Linear(C,1)(h))          model.py:179-181
//   x  <- posterior(x, eps, z)                                   utils/nn.py:84-87,110-112
//   next step: in = Linear(2,C)(x) -> raw and SiLU'd rows        model.py:139 (+ cnn.py:25)
// ---------------------------------------------------------------------------
#ifndef DHG_HEADS_MINB
#define DHG_HEADS_MINB 4
#endif
template <typename T_>
__global__ void __launch_bounds__(256, DHG_HEADS_MINB) heads_update_kernel(const T_* __restrict__ h, int C,
                                                              const float* __restrict__ Wo /*[2,C]*/,
                                                              const float* __restrict__ bo,
                                                              const float* __restrict__ Wp /*[1,C]*/,
                                                              const float* __restrict__ bp, HeadParams p) {
  // 16 lanes per stroke point (each lane owns 8 of the C = 128 channels: one 16-byte access per row for bf16),
  // 2 points per warp, grid-stride.  The kernel is a stream (read h, write the next step's two input rows), so what
  // matters is bytes in flight per SM: the six weight vectors live in shared memory instead of 48 registers (4 blocks
  // per SM), and the row of the NEXT point is requested before the current one is worked on.
  __shared__ __align__(16) float ws[9][128];   // w0 | w1 | wp | in_W[:,0] | in_W[:,1] | in_b | second input's w0 | w1 | wp
  const bool next_in = p.next_raw != nullptr || p.next_act != nullptr;
  const bool two = p.h2 != nullptr;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    ws[0][i] = Wo[i]; ws[1][i] = Wo[C + i]; ws[2][i] = Wp[i];
    ws[3][i] = next_in ? p.in_W[i * 2] : 0.f;       // input_dense rows (next step)
    ws[4][i] = next_in ? p.in_W[i * 2 + 1] : 0.f;
    ws[5][i] = next_in ? p.in_b[i] : 0.f;
    ws[6][i] = two ? p.w2[i] : 0.f; ws[7][i] = two ? p.w2[C + i] : 0.f; ws[8][i] = two ? p.w2[2 * C + i] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 15, grp = lane >> 4;
  const int c0 = sub * 8;
  const float b0 = bo[0], b1 = bo[1], bpv = bp[0];
  // 32-bit point indices (the plan guarantees B * (T + 1) < 2^24); the (sample, position) of my point advances by a
  // fixed stride per iteration, so the row index is kept incrementally instead of dividing by T every time
  const uint32_t npts = (uint32_t)p.B * (uint32_t)p.T, T = (uint32_t)p.T;
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t stride = ((gridDim.x * blockDim.x) >> 5) * 2;
  const uint32_t stride_b = stride / T, stride_t = stride - stride_b * T;
  float xv[8], xn[8], sv[8], sn[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { xn[k] = 0.f; sn[k] = 0.f; }
  float2 xx = make_float2(0.f, 0.f), zz = make_float2(0.f, 0.f), xx_n = xx, zz_n = zz;
  const T_* h2 = reinterpret_cast<const T_*>(p.h2);
  // point i = b * T + t lives in row b * (T + 1) + 1 + t = i + b + 1
  auto fetch = [&](uint32_t i, uint32_t b, float* v, float2& x2, float2& z2) {
    load8<T_>(h + (size_t)(i + b + 1) * C + c0, v);
    if (two) load8<T_>(h2 + (size_t)(i + b + 1) * C + c0, sn);
    if (p.x_io) x2 = *reinterpret_cast<const float2*>(p.x_io + (size_t)i * 2);
    if (p.noise) z2 = *reinterpret_cast<const float2*>(p.noise + (size_t)i * 2);
  };
  uint32_t i = warp0 * 2 + grp;
  uint32_t b = i / T, t = i - b * T;           // one division per thread
  uint32_t i_n = i, b_n = b, t_n = t;          // the point after this one
  auto advance = [&](uint32_t& ii, uint32_t& bb, uint32_t& tt) {
    ii += stride; bb += stride_b; tt += stride_t;
    if (tt >= T) { tt -= T; ++bb; }
  };
  if (i < npts) fetch(i, b, xn, xx_n, zz_n);
  for (; i - grp < npts; advance(i, b, t)) {   // warp-uniform trip count (the shuffles below need both halves)
#pragma unroll
    for (int k = 0; k < 8; ++k) { xv[k] = xn[k]; sv[k] = sn[k]; }
    xx = xx_n; zz = zz_n;
    advance(i_n, b_n, t_n);
    if (i_n < npts) fetch(i_n, b_n, xn, xx_n, zz_n);
    const bool ok = i < npts;
    const size_t row = (size_t)i + b + 1;
    float e0 = 0.f, e1 = 0.f, pl = 0.f;
    {
      const float4 *w0 = reinterpret_cast<const float4*>(&ws[0][c0]), *w1 = reinterpret_cast<const float4*>(&ws[1][c0]),
                   *wp = reinterpret_cast<const float4*>(&ws[2][c0]);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 a = w0[q], b = w1[q], c = wp[q];
        const float aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          e0 = fmaf(xv[q * 4 + k], aw[k], e0);
          e1 = fmaf(xv[q * 4 + k], bw[k], e1);
          pl = fmaf(xv[q * 4 + k], cw[k], pl);
        }
      }
      if (two) {
        const float4 *u0 = reinterpret_cast<const float4*>(&ws[6][c0]), *u1 = reinterpret_cast<const float4*>(&ws[7][c0]),
                     *up = reinterpret_cast<const float4*>(&ws[8][c0]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 a = u0[q], b = u1[q], c = up[q];
          const float aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            e0 = fmaf(sv[q * 4 + k], aw[k], e0);
            e1 = fmaf(sv[q * 4 + k], bw[k], e1);
            pl = fmaf(sv[q * 4 + k], cw[k], pl);
          }
        }
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      e0 += __shfl_xor_sync(0xffffffffu, e0, o);
      e1 += __shfl_xor_sync(0xffffffffu, e1, o);
      pl += __shfl_xor_sync(0xffffffffu, pl, o);
    }
    e0 += b0; e1 += b1; pl += bpv;
    if (!ok) continue;
    float y0 = 0.f, y1 = 0.f;
    if (p.x_io) {
      if (p.mode == 0) {  // "new": (x - sqrt(1-abar) eps) / sqrt(1-beta) + z sqrt(1-abar_next)
        y0 = (xx.x - p.c_eps * e0) / p.c_div + zz.x * p.c_noise;
        y1 = (xx.y - p.c_eps * e1) / p.c_div + zz.y * p.c_noise;
      } else {            // "standard": (1/sqrt(1-beta)) (x - beta eps / sqrt(1-abar)) + sqrt(beta) z
        y0 = p.c_div * (xx.x - p.c_eps * e0 / p.c_eps2) + p.c_noise * zz.x;
        y1 = p.c_div * (xx.y - p.c_eps * e1 / p.c_eps2) + p.c_noise * zz.y;
      }
    }
    if (sub == 0) {
      if (p.eps_out) { p.eps_out[i * 2] = e0; p.eps_out[i * 2 + 1] = e1; }
      if (p.pen_out) p.pen_out[i * p.pen_stride + p.pen_offset] = 1.f / (1.f + expf(-pl));
      if (p.x_io) {
        float* xo = p.x_out ? p.x_out : p.x_io;
        xo[i * p.x_out_stride] = y0;
        xo[i * p.x_out_stride + 1] = y1;
      }
    }
    if (next_in) {   // input_dense of the next step on the updated point
      float v[8];
      const float4 *i0 = reinterpret_cast<const float4*>(&ws[3][c0]), *i1 = reinterpret_cast<const float4*>(&ws[4][c0]),
                   *ib = reinterpret_cast<const float4*>(&ws[5][c0]);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 a = i0[q], b = i1[q], c = ib[q];
        v[q * 4 + 0] = fmaf(y1, b.x, fmaf(y0, a.x, c.x));
        v[q * 4 + 1] = fmaf(y1, b.y, fmaf(y0, a.y, c.y));
        v[q * 4 + 2] = fmaf(y1, b.z, fmaf(y0, a.z, c.z));
        v[q * 4 + 3] = fmaf(y1, b.w, fmaf(y0, a.w, c.w));
      }
      if (p.next_raw) store8<T_>(reinterpret_cast<T_*>(p.next_raw) + row * C + c0, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = silu_out<T_>(v[k]);
      if (p.next_act) store8<T_>(reinterpret_cast<T_*>(p.next_act) + row * C + c0, v);
    }
  }
}
// Tail fusion, level 2: the three head values of a point are already sitting in two [rows, 4] fp32 arrays (dot mode of
// the last ConvBlock's conv2 and conv_skip GEMMs): eps | pen = dot_a[row] + dot_b[row] + c.  What is left is the
// posterior update and input_dense of the next step.  A warp owns 32 consecutive points: lane i does the scalar work of
// point i (coalesced 16-byte / 8-byte loads, one update per lane instead of one per 16 lanes), then the warp writes the
// 32 next-step input rows one after the other, lane = 4 of the C = 128 channels (one 256-byte row per store
// instruction): ~30 warp instructions per point instead of ~75, the kernel is a store stream (C * 2 B per point and copy).
template <typename T_>
__global__ void __launch_bounds__(256) heads_from_dots_kernel(const float4* __restrict__ dot_a, const float4* __restrict__ dot_b,
                                                              const float* __restrict__ cst /*[3]*/, int C, HeadParams p) {
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 4;
  float iw0[4], iw1[4], ib[4];
  const bool next_in = p.next_raw != nullptr || p.next_act != nullptr;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    iw0[k] = next_in ? p.in_W[(c0 + k) * 2] : 0.f;
    iw1[k] = next_in ? p.in_W[(c0 + k) * 2 + 1] : 0.f;
    ib[k] = next_in ? p.in_b[c0 + k] : 0.f;
  }
  const float b0 = cst[0], b1 = cst[1], bpv = cst[2];
  const float r_div = 1.f / p.c_div, r_eps2 = p.c_eps2 != 0.f ? 1.f / p.c_eps2 : 0.f;
  const uint32_t npts = (uint32_t)p.B * (uint32_t)p.T, T = (uint32_t)p.T;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < npts; base += warps * 32u) {
    const uint32_t i = base + lane;
    const bool ok = i < npts;
    const uint32_t ii = ok ? i : npts - 1;
    const uint32_t b = ii / T;
    const uint32_t row = ii + b + 1;   // point i = b * T + t lives in row b * (T + 1) + 1 + t
    const float4 da = dot_a[row], db = dot_b[row];
    float2 xx = make_float2(0.f, 0.f), zz = xx;
    if (p.x_io) xx = *reinterpret_cast<const float2*>(p.x_io + (size_t)ii * 2);
    if (p.noise) zz = *reinterpret_cast<const float2*>(p.noise + (size_t)ii * 2);
    const float e0 = da.x + db.x + b0, e1 = da.y + db.y + b1, pl = da.z + db.z + bpv;
    float y0 = 0.f, y1 = 0.f;
    if (p.x_io) {
      // bf16 chain only: the two divisions of the update formulas as multiplications by reciprocals taken once per thread
      if (p.mode == 0) {
        y0 = fmaf(zz.x, p.c_noise, (xx.x - p.c_eps * e0) * r_div);
        y1 = fmaf(zz.y, p.c_noise, (xx.y - p.c_eps * e1) * r_div);
      } else {
        y0 = fmaf(p.c_noise, zz.x, p.c_div * (xx.x - p.c_eps * e0 * r_eps2));
        y1 = fmaf(p.c_noise, zz.y, p.c_div * (xx.y - p.c_eps * e1 * r_eps2));
      }
    }
    if (ok) {
      if (p.eps_out) *reinterpret_cast<float2*>(p.eps_out + (size_t)i * 2) = make_float2(e0, e1);
      if (p.pen_out) p.pen_out[(size_t)i * p.pen_stride + p.pen_offset] = 1.f / (1.f + expf(-pl));
      if (p.x_io) {
        float* xo = p.x_out ? p.x_out : p.x_io;
        xo[(size_t)i * p.x_out_stride] = y0;
        xo[(size_t)i * p.x_out_stride + 1] = y1;
      }
    }
    if (next_in) {
      const uint32_t nvalid = npts - base < 32u ? npts - base : 32u;   // warp-uniform
      T_* raw = reinterpret_cast<T_*>(p.next_raw);
      T_* act = reinterpret_cast<T_*>(p.next_act);
#pragma unroll 4
      for (uint32_t j = 0; j < nvalid; ++j) {
        const float yj0 = __shfl_sync(0xffffffffu, y0, j), yj1 = __shfl_sync(0xffffffffu, y1, j);
        const size_t off = (size_t)__shfl_sync(0xffffffffu, row, j) * C + c0;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = fmaf(yj1, iw1[k], fmaf(yj0, iw0[k], ib[k]));
        if (raw) store4<T_>(raw + off, make_float4(v[0], v[1], v[2], v[3]));
        if (act) store4<T_>(act + off, make_float4(silu_out<T_>(v[0]), silu_out<T_>(v[1]), silu_out<T_>(v[2]), silu_out<T_>(v[3])));
      }
    }
  }
}
template <typename T>
int launch_heads_from_dots(const float* dot_a, const float* dot_b, const float* cst, int C, const HeadParams& p, cudaStream_t st) {
  if (C != 128) return 1;
  const size_t nw = ((size_t)p.B * p.T + 31) / 32;   // one warp per 32 points
  size_t blocks = (nw + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  heads_from_dots_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(dot_a), reinterpret_cast<const float4*>(dot_b), cst, C, p);
  return 0;
}
template int launch_heads_from_dots<bf16>(const float*, const float*, const float*, int, const HeadParams&, cudaStream_t);

template <typename T>
int launch_heads_update(const T* h, int C, const float* Wo, const float* bo, const float* Wp,
                        const float* bp, const HeadParams& p, cudaStream_t st) {
  if (C != 128) return 1;
  const size_t nw = ((size_t)p.B * p.T + 1) / 2;
  size_t blocks = (nw + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;   // two waves of 4 resident blocks per SM
  heads_update_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(h, C, Wo, bo, Wp, bp, p);
  return 0;
}
template int launch_heads_update<float>(const float*, int, const float*, const float*, const float*, const float*, const HeadParams&, cudaStream_t);
template int launch_heads_update<bf16>(const bf16*, int, const float*, const float*, const float*, const float*, const HeadParams&, cudaStream_t);
template int launch_heads_update<bfs>(const bfs*, int, const float*, const float*, const float*, const float*, const HeadParams&, cudaStream_t);

// ---------------------------------------------------------------------------
// Standalone posterior update (drop-in for utils/nn.py:64-112 with injected z):
// pure streaming kernel, float4 vectorised, grid-stride.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) posterior_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                                        const float4* __restrict__ z, float4* __restrict__ out,
                                                        size_t n4, int mode, float c_eps, float c_eps2,
                                                        float c_div, float c_noise) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = x[i], e = eps[i];
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (z) n = z[i];
    float4 y;
    if (mode == 0) {
      y.x = (a.x - c_eps * e.x) / c_div + n.x * c_noise;
      y.y = (a.y - c_eps * e.y) / c_div + n.y * c_noise;
      y.z = (a.z - c_eps * e.z) / c_div + n.z * c_noise;
      y.w = (a.w - c_eps * e.w) / c_div + n.w * c_noise;
    } else {
      y.x = c_div * (a.x - c_eps * e.x / c_eps2) + c_noise * n.x;
      y.y = c_div * (a.y - c_eps * e.y / c_eps2) + c_noise * n.y;
      y.z = c_div * (a.z - c_eps * e.z / c_eps2) + c_noise * n.z;
      y.w = c_div * (a.w - c_eps * e.w / c_eps2) + c_noise * n.w;
    }
    out[i] = y;
  }
}
void launch_posterior(const float* x, const float* eps, const float* z, float* out, size_t n, int mode,
                      float c_eps, float c_eps2, float c_div, float c_noise, int num_sms, cudaStream_t st) {
  const size_t n4 = n / 4;  // n = B*T*2, T % 8 == 0
  size_t blocks = (n4 + 255) / 256;
  const size_t cap = (size_t)num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  posterior_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(eps),
                                                     reinterpret_cast<const float4*>(z), reinterpret_cast<float4*>(out),
                                                     n4, mode, c_eps, c_eps2, c_div, c_noise);
}

// ---------------------------------------------------------------------------
// Conditioning: sigma_ffn (utils/nn.py:145-175; SiLU first) and all 38 FiLM
// linears (conditioning.py:17-18) for a set of noise levels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sigma_ffn_kernel(const float* __restrict__ sigma, const float* __restrict__ w1 /*[H,1]*/,
                                                        const float* __restrict__ b1, const float* __restrict__ w2 /*[32,H]*/,
                                                        const float* __restrict__ b2, int H, float* __restrict__ out /*[n,32]*/) {
  extern __shared__ float hid[];
  const float s = silu_f(sigma[blockIdx.x]);
  for (int i = threadIdx.x; i < H; i += blockDim.x) hid[i] = silu_f(fmaf(s, w1[i], b1[i]));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < 32; o += blockDim.x >> 5) {
    float a = 0.f;
    for (int i = lane; i < H; i += 32) a = fmaf(hid[i], w2[(size_t)o * H + i], a);
    a = warp_sum(a);
    if (lane == 0) out[blockIdx.x * 32 + o] = a + b2[o];
  }
}
void launch_sigma_ffn(const float* sigma, const float* w1, const float* b1, const float* w2, const float* b2,
                      int H, float* out, int n, cudaStream_t st) {
  sigma_ffn_kernel<<<n, 256, H * sizeof(float), st>>>(sigma, w1, b1, w2, b2, H, out);
}

__global__ void __launch_bounds__(256) film_table_kernel(const float* __restrict__ sig /*[n,32]*/, const float* __restrict__ Wc /*[tot,32]*/,
                                                         const float* __restrict__ bc, int tot, float* __restrict__ out /*[n,tot]*/) {
  __shared__ float s[32];
  if (threadIdx.x < 32) s[threadIdx.x] = sig[blockIdx.y * 32 + threadIdx.x];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= tot) return;
  const float4* w = reinterpret_cast<const float4*>(Wc + (size_t)j * 32);
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 x = w[i];
    a = fmaf(x.x, s[i * 4], a); a = fmaf(x.y, s[i * 4 + 1], a);
    a = fmaf(x.z, s[i * 4 + 2], a); a = fmaf(x.w, s[i * 4 + 3], a);
  }
  out[(size_t)blockIdx.y * tot + j] = a + bc[j];
}
void launch_film_table(const float* sig, const float* Wc, const float* bc, int tot, float* out, int n, cudaStream_t st) {
  dim3 grid((tot + 255) / 256, n);
  film_table_kernel<<<grid, 256, 0, st>>>(sig, Wc, bc, tot, out);
}

// ---------------------------------------------------------------------------
// Text embedding gather + LayerNorm (text_style.py:98-99), one warp per token.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) embed_ln_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb,
                                                       int vocab, int C, T* __restrict__ out, int rows, int* __restrict__ err) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  long long id = ids[r];
  if (id < 0 || id >= vocab) { if (lane == 0) atomicExch(err, 1); id = 0; }
  const float* e = emb + (size_t)id * C;
  float4 v[4];
  float s = 0.f;
  const int nchunk = (C + 127) >> 7;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int n = c * 128 + lane * 4;
    if (c < nchunk && n < C) { v[c] = *reinterpret_cast<const float4*>(e + n); s += (v[c].x + v[c].y) + (v[c].z + v[c].w); }
    else v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int n = c * 128 + lane * 4;
    if (c < nchunk && n < C) {
      const float dx = v[c].x - mean, dy = v[c].y - mean, dz = v[c].z - mean, dw = v[c].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-6f);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int n = c * 128 + lane * 4;
    if (c < nchunk && n < C)
      store4<T>(out + (size_t)r * C + n, make_float4((v[c].x - mean) * rstd, (v[c].y - mean) * rstd,
                                                     (v[c].z - mean) * rstd, (v[c].w - mean) * rstd));
  }
}
template <typename T>
void launch_embed_ln(const int64_t* ids, const float* emb, int vocab, int C, T* out, int rows, int* err, cudaStream_t st) {
  embed_ln_kernel<T><<<(rows + 7) / 8, 256, 0, st>>>(ids, emb, vocab, C, out, rows, err);
}
template void launch_embed_ln<float>(const int64_t*, const float*, int, int, float*, int, int*, cudaStream_t);
template void launch_embed_ln<bf16>(const int64_t*, const float*, int, int, bf16*, int, int*, cudaStream_t);
template void launch_embed_ln<bfs>(const int64_t*, const float*, int, int, bfs*, int, int*, cudaStream_t);

// SiLU + dtype conversion over a flat fp32 array (style vector -> style_ffn input;
// reshape_up(style, 5) is a pure reinterpretation of the contiguous buffer).
template <typename T>
__global__ void silu_convert_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = reinterpret_cast<const float4*>(in)[i];
  store4<T>(out + i * 4, make_float4(silu_out<T>(x.x), silu_out<T>(x.y), silu_out<T>(x.z), silu_out<T>(x.w)));
}
template <typename T>
void launch_silu_convert(const float* in, T* out, size_t n, cudaStream_t st) {
  const size_t n4 = n / 4;
  silu_convert_kernel<T><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(in, out, n4);
}
template void launch_silu_convert<float>(const float*, float*, size_t, cudaStream_t);
template void launch_silu_convert<bf16>(const float*, bf16*, size_t, cudaStream_t);
template void launch_silu_convert<bfs>(const float*, bfs*, size_t, cudaStream_t);

}  // namespace dhg
