// Fused tensor-core attention for head depth 64 or 48, sm_100a (tcgen05 / TMEM / TMA).
//
//   O[b, tq, h, :] = softmax_j( scale * <Q[b,tq,h,:], K[b,j,h,:]> + mask[b,j] ) @ V[b,j,h,:]
//
// Reference: scaled_dp_attn (attention.py:26-46: SDPA with an additive -1e9 float mask on padded text
// keys) and the head split/merge of MultiHeadAttention.forward (attention.py:78-85), which in the
// channels-last row layout is just a column slice [h*64, h*64+64) of the q / k / v / o row matrices.
//
// One CTA = one (128-query tile, head, sample).  All keys of the sample (Tk <= 256) are handled in one
// shot, so there is no online-softmax rescaling:
//   warp 4      TMA loads Q [128 x 64], K [N x 64], V [N x 64] (SWIZZLE_128B; N = Tk rounded up to 16),
//               S = Q K^T   : tcgen05.mma, both operands K-major, fp32 scores in TMEM columns [0, N)
//               O = P V     : tcgen05.mma, A = P (bf16, K-major, written to smem by the softmax
//                             threads), B = V used as an MN-major operand straight from its TMA tile
//   warps 0-3   thread = query row: max and exp2 over the score row read from TMEM, P -> smem (the
//               swizzled K-major layout the MMA expects), then O row from TMEM, 1/sum, bf16 store.
// P overlays the Q and K tiles (dead once S is complete) and O overlays S in TMEM, so a CTA needs
// <= 90 KB of smem and <= 256 TMEM columns and several CTAs share an SM, overlapping each other's
// load / MMA / softmax phases.
#include "kernels.h"
#include "tc_common.cuh"

namespace dhg {

using namespace tc;

namespace {

constexpr int AT_THREADS = 160;

struct AttnTcShape {
  int N;          // keys padded to a multiple of 16 (UMMA N of the score MMA, K extent of the PV MMA)
  int nblk;       // 64-key blocks of P
  int nchunk;     // 32-column chunks of the score row
  int tmem_cols;  // power of two >= max(64, nchunk * 32)
  uint32_t idesc_s, idesc_o;
  uint32_t off_k, off_v, off_mask, off_bar;
  float scale_log2;   // scale * log2(e)
};

__global__ void __launch_bounds__(AT_THREADS) attn_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                             const __grid_constant__ CUtensorMap map_k,
                                                             const __grid_constant__ CUtensorMap map_v,
                                                             const AttnTcShape sh, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;               // [128 x 64] bf16, later P block 0
  uint8_t* k_s = smem + sh.off_k;    // [N x 64]
  uint8_t* v_s = smem + sh.off_v;    // [N x 64]
  float* mask_s = reinterpret_cast<float*>(smem + sh.off_mask);   // additive mask in log2 units, per key
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sh.off_bar);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_p = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int N = sh.N;

  if (warp == 4) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
      mbar_init(smem_u32(bar_load), 1);
      mbar_init(smem_u32(bar_s), 1);
      mbar_init(smem_u32(bar_p), 128);
      mbar_init(smem_u32(bar_o), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)sh.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    // additive key mask (attention.py:44: mask * -1e9), pre-multiplied by log2(e); keys >= Tk never count
    for (int j = threadIdx.x; j < sh.nchunk * 32; j += 128) {
      float mv = 0.f;
      if (j < p.Tk && p.text && p.text[(size_t)b * p.Tk + j] == 0) mv = -1e9f * 1.4426950408889634f;
      mask_s[j] = mv;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // the whole warp runs this (uniform) sequence; one elected lane issues TMA / tcgen05
    const bool leader = elect_one();
    const uint32_t lb = smem_u32(bar_load);
    if (leader) {
      mbar_expect_tx(lb, (uint32_t)(128 * 128 + 2 * N * 128));
      // 64-column boxes starting at the head's first column; for D = 48 the last 16 columns belong to
      // the next head (or are zero-filled past the matrix) and are never touched by the MMAs
      tma_load_2d(smem_u32(q_s), &map_q, lb, h * p.D, b * p.q_period + p.q_pad + qt * 128);
      tma_load_2d(smem_u32(k_s), &map_k, lb, h * p.D, b * p.k_period + p.k_pad);
      tma_load_2d(smem_u32(v_s), &map_v, lb, h * p.D, b * p.k_period + p.k_pad);
    }
    mbar_wait(lb, 0);
    tc_fence_after();
    // S = Q K^T
    const uint32_t qlo = umma_desc_lo(smem_u32(q_s)), klo = umma_desc_lo(smem_u32(k_s));
    if (leader) {
      for (int k = 0; k < (p.D >> 4); ++k)
        umma_bf16(tmem_base, umma_desc_make(qlo + 2 * k, kDescHiSw128), umma_desc_make(klo + 2 * k, kDescHiSw128), sh.idesc_s, k ? 1u : 0u);
      umma_commit(smem_u32(bar_s));
    }
    __syncwarp();
    // O = P V   (P written by the softmax threads over the Q/K tiles)
    mbar_wait(smem_u32(bar_p), 0);
    tc_fence_after();
    const int nk = N >> 4;
    const uint32_t vlo = umma_desc_lo(smem_u32(v_s)) & ~(1u << 16);   // MN-major: LBO field set below
    const uint32_t vlo_mn = vlo | ((1024u >> 4) << 16);
    if (leader) {
      for (int kk = 0; kk < nk; ++kk) {
        const uint32_t alo = qlo + (uint32_t)(kk >> 2) * (16384u >> 4) + (uint32_t)(kk & 3) * 2u;
        umma_bf16(tmem_base, umma_desc_make(alo, kDescHiSw128), umma_desc_make(vlo_mn + (uint32_t)kk * (2048u >> 4), kDescHiSw128),
                  sh.idesc_o, kk ? 1u : 0u);
      }
      umma_commit(smem_u32(bar_o));
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;           // query row inside the tile = TMEM lane
    const int tq = qt * 128 + r;
    const uint32_t trow = tmem_base + (((uint32_t)(warp * 32)) << 16);
    float v[32];
    mbar_wait(smem_u32(bar_s), 0);
    tc_fence_after();
    // pass 1: row maximum of scale*s + mask (log2 units)
    float mx = -INFINITY;
    for (int c = 0; c < sh.nchunk; ++c) {
      tmem_ld32(trow + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = c * 32 + i;
        const float t = fmaf(v[i], sh.scale_log2, mask_s[j]);
        if (j < p.Tk) mx = fmaxf(mx, t);
      }
    }
    // pass 2: p = 2^(t - max), row sum, bf16 P into the swizzled K-major operand layout
    float sum = 0.f;
    for (int c = 0; c < sh.nchunk; ++c) {
      tmem_ld32(trow + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = c * 32 + i;
        const float t = fmaf(v[i], sh.scale_log2, mask_s[j]) - mx;
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
        v[i] = j < p.Tk ? e : 0.f;
        sum += v[i];
      }
      uint8_t* blk = q_s + (size_t)(c >> 1) * 16384 + (size_t)r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (c & 1) * 4 + g;   // 16-byte chunk (8 keys) inside the 64-key block row
        const uint4 u = make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
        *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) = u;
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA
    mbar_arrive(smem_u32(bar_p));
    // O row
    mbar_wait(smem_u32(bar_o), 0);
    tc_fence_after();
    const float inv = 1.f / sum;
    bf16* orow = reinterpret_cast<bf16*>(p.o) + ((size_t)b * p.q_period + p.q_pad + tq) * p.o_pitch + h * p.D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(trow + c * 32, v);
      if (tq < p.Tq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c * 32 + g * 8 >= p.D) break;
          const uint4 u = make_uint4(pack_bf16x2(v[g * 8] * inv, v[g * 8 + 1] * inv), pack_bf16x2(v[g * 8 + 2] * inv, v[g * 8 + 3] * inv),
                                     pack_bf16x2(v[g * 8 + 4] * inv, v[g * 8 + 5] * inv), pack_bf16x2(v[g * 8 + 6] * inv, v[g * 8 + 7] * inv));
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)sh.tmem_cols) : "memory");
  }
}

}  // namespace

struct AttnTcPlan {
  CUtensorMap map_q, map_k, map_v;
  AttnTcShape sh;
  AttnParams p;
  dim3 grid;
  size_t smem;
};

bool attn_tc_supported(const AttnParams& p) {
  return (p.D == 64 || p.D == 48) && p.Tk >= 1 && p.Tk <= 256 && p.q_pitch % 8 == 0 && p.k_pitch % 8 == 0 && p.v_pitch % 8 == 0 &&
         p.o_pitch % 8 == 0;
}

AttnTcPlan* attn_tc_plan_create(const AttnParams& p, int q_rows, int k_rows, char* err, int errlen) {
  if (!attn_tc_supported(p)) { snprintf(err, errlen, "attention shape not supported by the tcgen05 kernel (D=%d Tk=%d)", p.D, p.Tk); return nullptr; }
  AttnTcPlan* a = new AttnTcPlan();
  a->p = p;
  AttnTcShape& sh = a->sh;
  sh.N = (p.Tk + 15) & ~15;
  sh.nblk = (sh.N + 63) / 64;
  sh.nchunk = (sh.N + 31) / 32;
  int cols = sh.nchunk * 32 < 64 ? 64 : sh.nchunk * 32;
  sh.tmem_cols = cols <= 64 ? 64 : cols <= 128 ? 128 : 256;
  // c_format F32 [4,6) | a,b BF16 [7,10),[10,13) | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  sh.idesc_s = base | ((uint32_t)(sh.N >> 3) << 17);
  sh.idesc_o = base | (1u << 16) | ((uint32_t)(p.D >> 3) << 17);   // B = V is MN-major, N = head depth
  sh.scale_log2 = p.scale * 1.4426950408889634f;
  const uint32_t kv_bytes = (uint32_t)sh.N * 128u;
  uint32_t pq = 16384u + kv_bytes;                    // Q | K
  if (pq < (uint32_t)sh.nblk * 16384u) pq = (uint32_t)sh.nblk * 16384u;   // overlaid by P
  sh.off_k = 16384u;
  sh.off_v = pq;
  sh.off_mask = pq + kv_bytes;
  sh.off_bar = sh.off_mask + (uint32_t)sh.nchunk * 32u * 4u;
  size_t smem = sh.off_bar + 64 + 1024;
  // keep (CTAs per SM) * tmem_cols <= 512 so that no CTA ever waits in tcgen05.alloc
  const size_t min_smem = (size_t)(227 * 1024) / (512 / sh.tmem_cols) - 1024;
  const size_t floor_smem = (size_t)(227 * 1024) / ((512 / sh.tmem_cols) + 1) + 1;
  if (smem < floor_smem) smem = floor_smem < min_smem ? floor_smem : min_smem;
  a->smem = smem;
  a->grid = dim3((p.Tq + 127) / 128, p.H, p.B);
  const uint64_t qcols = (uint64_t)p.H * p.D, kcols = (uint64_t)p.H * p.D;
  if (!make_map(&a->map_q, p.q, (uint64_t)q_rows, qcols, (uint64_t)p.q_pitch, 128, err, errlen) ||
      !make_map(&a->map_k, p.k, (uint64_t)k_rows, kcols, (uint64_t)p.k_pitch, (uint32_t)sh.N, err, errlen) ||
      !make_map(&a->map_v, p.v, (uint64_t)k_rows, kcols, (uint64_t)p.v_pitch, (uint32_t)sh.N, err, errlen)) {
    delete a;
    return nullptr;
  }
  cudaError_t ce = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (ce != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); delete a; return nullptr; }
  return a;
}

void attn_tc_plan_destroy(AttnTcPlan* a) { delete a; }

int attn_tc_launch(const AttnTcPlan* a, cudaStream_t st) {
  attn_tc_kernel<<<a->grid, AT_THREADS, a->smem, st>>>(a->map_q, a->map_k, a->map_v, a->sh, a->p);
  return 0;
}

}  // namespace dhg
