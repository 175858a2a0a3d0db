// tcgen05 / TMEM / TMA GEMM with the fused row epilogue, sm_100a.
//
//   out[m, n] = epilogue( sum_{tap<taps} sum_{k<K} A[m + tap - taps/2, k] * W[tap][n][k] )
//
// One CTA computes one 128 x BN output tile (BN <= 256, or 384 = two N=192 MMAs when a
// LayerNorm needs the whole 384-wide row).  Warp roles (192 threads):
//   warps 0-3  epilogue: thread = one accumulator row (TMEM lane); tcgen05.ld 32 columns at a time
//   warp  4    TMA producer: A tile [128 rows x 64 k] and W tile [BN x 64 k] per k-block, SWIZZLE_128B,
//              row-shifted A coordinates implement the conv taps, TMA zero-fill implements every edge
//   warp  5    TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, bf16 x bf16 -> fp32 in TMEM)
// Pipelines: smem full/empty mbarriers between TMA and MMA, one tmem_full barrier to the epilogue.
// Several CTAs are resident per SM (smem <= ~100 KB, TMEM <= 256 columns each for BN <= 256), so one
// CTA's epilogue overlaps another CTA's MMA main loop.
//
// The epilogue is the one documented in common.cuh (bias|rowbias, residual, LayerNorm, FiLM, residual,
// halo-row zeroing, raw and/or SiLU'd bf16 stores).  LayerNorm: pass 1 adds bias/residual, writes the
// row back to TMEM and accumulates shifted sums; pass 2 normalises.  Reference ops fused here:
// Linear/Conv1d (cnn.py:32-49, attention.py:58-61), LayerNorm (model.py:25), AffineTransformLayer
// (conditioning.py:16-19), SiLU (cnn.py:25), residual adds and nearest upsample (model.py:169-176).
#include <cuda.h>
#include <stdio.h>

#include "kernels.h"

namespace dhg {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;   // bf16 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int TC_THREADS = 192;
constexpr uint32_t kSpinLimit = 1u << 27;

struct TcShape {
  int rows, K, N, taps;
  int BN;         // tile width
  int stages;
  int kb_per_tap; // ceil(K / 64)
  int umma_n;     // N of one tcgen05.mma (BN, or 192 when BN == 384)
  int n_umma;     // MMAs per k-step along N (1 or 2)
  uint32_t idesc;
  int tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) {
      printf("tc_gemm: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;             // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: 8 rows x 128 B per swizzle atom
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  __syncwarp();   // .sync.aligned: the warp must be converged
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  __syncwarp();
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 consecutive bf16 (64 B) <-> fp32 registers
__device__ __forceinline__ void add_bf16x32(const bf16* p, float* v) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = q[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
      v[i * 8 + j * 2] += f.x;
      v[i * 8 + j * 2 + 1] += f.y;
    }
  }
}
template <bool kSilu>
__device__ __forceinline__ void store_bf16x32(bf16* p, const float* v) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = v[i * 8 + j * 2], b = v[i * 8 + j * 2 + 1];
      if (kSilu) { a = silu_f(a); b = silu_f(b); }
      const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    q[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_w,
                                                             const TcShape sh, const Epilogue e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: 1024-aligned tiles, then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 2;           // 16 KB
  const uint32_t b_bytes = (uint32_t)sh.BN * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)sh.stages * stage_bytes);
  uint64_t* full_bar = bars;                 // [stages]
  uint64_t* empty_bar = bars + sh.stages;    // [stages]
  uint64_t* tmem_full_bar = bars + 2 * sh.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * sh.stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * sh.BN;
  const int num_kb = sh.taps * sh.kb_per_tap;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < sh.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)sh.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int half_taps = sh.taps / 2;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % sh.stages;
        const uint32_t phase = (uint32_t)(kb / sh.stages) & 1u;
        mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1u);
        const int tap = kb / sh.kb_per_tap, kk = (kb - tap * sh.kb_per_tap) * TC_BK;
        const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_dst = a_dst + a_bytes;
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, stage_bytes);
        tma_load_2d(a_dst, &map_a, fb, kk, m0 + tap - half_taps);
        for (int j = 0; j < sh.n_umma; ++j)
          tma_load_2d(b_dst + (uint32_t)j * sh.umma_n * TC_BK * 2, &map_w, fb, kk, tap * sh.N + n0 + j * sh.umma_n);
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % sh.stages;
        const uint32_t phase = (uint32_t)(kb / sh.stages) & 1u;
        mbar_wait(smem_u32(&full_bar[s]), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t adesc = umma_desc_sw128(a_addr + k * 32);
          for (int j = 0; j < sh.n_umma; ++j) {
            const uint64_t bdesc = umma_desc_sw128(b_addr + (uint32_t)j * sh.umma_n * TC_BK * 2 + k * 32);
            umma_bf16(tmem_base + (uint32_t)j * sh.umma_n, adesc, bdesc, sh.idesc, (kb | k) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));        // frees the smem stage when these MMAs retire
      }
      umma_commit(smem_u32(tmem_full_bar));          // accumulator complete
    }
  } else {
    // ===== epilogue: warps 0..3, thread = row =====
    mbar_wait(smem_u32(tmem_full_bar), 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool in_range = m < sh.rows;
    const int mm = in_range ? m : 0;
    const int b = mm / e.map.period;
    const int j = mm - b * e.map.period;
    const bool is_pad = (mm >= e.map.nvalid) || (e.map.pad_first && j == 0);
    const int pos = is_pad ? 0 : j - e.map.pad_first;
    const int N = sh.N;
    const float* biasp = e.rowbias ? e.rowbias + (size_t)pos * N : e.bias;
    const bf16* rpre = e.res_pre ? reinterpret_cast<const bf16*>(e.res_pre) + (size_t)mm * e.res_pre_pitch : nullptr;
    const float* gam = e.gamma ? e.gamma + (size_t)b * e.film_bstride : nullptr;
    const float* bet = e.gamma ? e.beta + (size_t)b * e.film_bstride : nullptr;
    const bf16* rpost = nullptr;
    if (e.res_post) {
      const size_t rr = e.res_post_up ? (size_t)b * e.res_post_period_lo + 1 + (pos >> 1) : (size_t)mm;
      rpost = reinterpret_cast<const bf16*>(e.res_post) + rr * e.res_post_pitch;
    }
    bf16* oraw = e.out_raw ? reinterpret_cast<bf16*>(e.out_raw) + (size_t)mm * e.out_raw_pitch : nullptr;
    bf16* oact = e.out_act ? reinterpret_cast<bf16*>(e.out_act) + (size_t)mm * e.out_act_pitch : nullptr;
    const bool live = in_range && !is_pad;
    float v[32];
    float mean = 0.f, rstd = 1.f;
    if (e.ln) {
      // pass 1: x = acc + bias + res_pre -> back to TMEM; shifted sums for mean / variance
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      for (int c0 = 0; c0 < sh.BN; c0 += 32) {
        tmem_ld32(trow + c0, v);
        const int n = n0 + c0;
        if (biasp) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(biasp + n + i));
            v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
          }
        }
        if (rpre && live) add_bf16x32(rpre + n, v);
        if (c0 == 0) shift = v[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float dlt = v[i] - shift;
          s1 += dlt;
          s2 = fmaf(dlt, dlt, s2);
        }
        tmem_st32(trow + c0, v);
      }
      const float inv_n = 1.f / (float)sh.BN;
      const float dm = s1 * inv_n;
      mean = shift + dm;
      rstd = rsqrtf(fmaxf(s2 * inv_n - dm * dm, 0.f) + 1e-6f);
    }
    for (int c0 = 0; c0 < sh.BN; c0 += 32) {
      const int n = n0 + c0;
      tmem_ld32(trow + c0, v);   // warp-collective: executed by all lanes, also for dead rows
      if (n >= N) continue;      // tile column tail (BN does not divide N): nothing to store
      if (!in_range) continue;
      if (!live) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      } else {
        if (e.ln) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd;
        } else {
          if (biasp) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(biasp + n + i));
              v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
            }
          }
          if (rpre) add_bf16x32(rpre + n, v);
        }
        if (gam) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gam + n + i));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bet + n + i));
            v[i] = fmaf(v[i], g.x, bb.x); v[i + 1] = fmaf(v[i + 1], g.y, bb.y);
            v[i + 2] = fmaf(v[i + 2], g.z, bb.z); v[i + 3] = fmaf(v[i + 3], g.w, bb.w);
          }
        }
        if (rpost) add_bf16x32(rpost + n, v);
      }
      if (oraw) store_bf16x32<false>(oraw + n, v);
      if (oact) store_bf16x32<true>(oact + n, v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)sh.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 row-major tensor [rows, cols] with row pitch `pitch` elements; box = [box_rows, 64 cols], SWIZZLE_128B.
bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows, char* err, int errlen) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu box_rows=%u", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch, box_rows);
    return false;
  }
  return true;
}

}  // namespace

struct TcGemmPlan {
  CUtensorMap map_a, map_w;
  TcShape sh;
  dim3 grid;
  size_t smem;
};

TcGemmPlan* tc_gemm_plan_create(const bf16* A, int lda, int rows, const bf16* W, int K, int N, int taps, const Epilogue& e,
                                char* err, int errlen) {
  if (rows <= 0 || K % 8 || N % 32 || (taps != 1 && taps != 3)) {
    snprintf(err, errlen, "unsupported shape rows=%d K=%d N=%d taps=%d", rows, K, N, taps);
    return nullptr;
  }
  int BN = 0;
  if (e.ln) {
    if (!(N <= 256 || N == 384)) { snprintf(err, errlen, "LayerNorm epilogue needs N <= 256 or N == 384, got %d", N); return nullptr; }
    BN = N;
  } else {
    for (int cand : {256, 192, 128, 96, 64, 32})
      if (N % cand == 0) { BN = cand; break; }
  }
  if (BN == 0 || BN % 16) { snprintf(err, errlen, "no tile width for N=%d", N); return nullptr; }
  TcGemmPlan* p = new TcGemmPlan();
  TcShape& sh = p->sh;
  sh.rows = rows; sh.K = K; sh.N = N; sh.taps = taps; sh.BN = BN;
  sh.kb_per_tap = (K + TC_BK - 1) / TC_BK;
  sh.n_umma = BN > 256 ? 2 : 1;
  sh.umma_n = BN / sh.n_umma;
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format BF16 (1) [7,10) | b_format BF16 (1) [10,13) |
  // a,b K-major (0) | N>>3 [17,23) | M>>4 [24,29)
  sh.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(sh.umma_n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  sh.tmem_cols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (size_t)BN * TC_BK * 2;
  const int num_kb = taps * sh.kb_per_tap;
  int stages = BN <= 256 ? 2 : 3;   // <= ~100 KB so that >= 2 CTAs share an SM (BN=384: one CTA)
  if (stages > num_kb) stages = num_kb;
  sh.stages = stages;
  p->smem = stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
  p->grid = dim3((rows + TC_BM - 1) / TC_BM, N / BN);
  if (!make_map(&p->map_a, A, (uint64_t)rows, (uint64_t)K, (uint64_t)lda, TC_BM, err, errlen) ||
      !make_map(&p->map_w, W, (uint64_t)taps * N, (uint64_t)K, (uint64_t)K, (uint32_t)sh.umma_n, err, errlen)) {
    delete p;
    return nullptr;
  }
  cudaError_t ce = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (ce != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); delete p; return nullptr; }
  return p;
}

void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }

int tc_gemm_launch(const TcGemmPlan* p, const Epilogue& e, cudaStream_t st) {
  tc_gemm_kernel<<<p->grid, TC_THREADS, p->smem, st>>>(p->map_a, p->map_w, p->sh, e);
  return 0;
}

}  // namespace dhg
