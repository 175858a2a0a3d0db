// Fused tensor-core attention for head depth 64 or 48, sm_100a (tcgen05 / TMEM / TMA).  Persistent.
//
//   O[b, tq, h, :] = softmax_j( scale * <Q[b,tq,h,:], K[b,j,h,:]> + mask[b,j] ) @ V[b,j,h,:]
//
// Reference: scaled_dp_attn (attention.py:26-46: SDPA with an additive -1e9 float mask on padded text
// keys) and the head split/merge of MultiHeadAttention.forward (attention.py:78-85), which in the
// channels-last row layout is just a column slice [h*D, h*D+D) of the q / k / v / o row matrices.
//
// Work item = one (128-query tile, head, sample).  All keys of the sample (Tk <= 256) are handled in one
// shot, so there is no online-softmax rescaling.  One CTA per SM owns NS "slots"; a slot is a softmax
// warpgroup (4 warps, thread = query row) with its own smem tiles, mbarriers and TMEM columns, and works
// through its share of the items one after the other:
//   control warp  (one per slot) per item:
//                 TMA loads Q [128 x 64], K [N x 64], V [N x 64] (SWIZZLE_128B; N = Tk rounded up to 16),
//                 S = Q K^T : tcgen05.mma, both operands K-major, fp32 scores in the slot's TMEM columns,
//                 O = P V   : tcgen05.mma, A = P (bf16, K-major, written to smem by the softmax threads),
//                             B = V used as an MN-major operand straight from its TMA tile
//   softmax group max and exp2 over the score row read from TMEM, P -> smem (the swizzled K-major layout
//                 the MMA expects), then the O row from TMEM, 1/sum, bf16 store.
// P overlays the Q and K tiles (dead once S is complete) and O overlays S in TMEM.  While one slot is in
// its softmax, the others are loading or multiplying, and nothing is allocated or initialised per item.
#include "kernels.h"
#include "tc_common.cuh"

namespace dhg {

using namespace tc;

namespace {

constexpr int AT_MAX_SLOTS = 6;

struct AttnTcShape {
  int N;          // keys padded to a multiple of 16 (UMMA N of the score MMA, K extent of the PV MMA)
  int nblk;       // 64-key blocks of P
  int nchunk;     // 32-column chunks of the score row
  int tmem_cols;  // TMEM columns per slot: max(64, nchunk * 32)
  int NS;         // slots per CTA
  int halves;     // softmax warps per TMEM lane quarter (1 or 2): the score columns are split between them
  int QT;         // query tiles per (sample, head)
  int items;      // B * H * QT
  int rev;        // walk the items from the last to the first
  int early;      // request the next item's tiles right after P V (else: after O has been stored)
  uint32_t idesc_s, idesc_o;
  uint32_t off_k, off_v, off_mask, off_xchg, slot_bytes, off_bar;
  uint32_t kv_bytes;  // bytes of one [N keys x 64 bf16] K or V block
  uint32_t off_pl;    // split storage: offset of the P_lo tile behind the P_hi tile
  float scale_log2;   // scale * log2(e)
  int dbg;            // timing experiments: 1 skip max pass, 2 skip exp, 4 skip P stores, 8 skip O stores, 16 skip PV MMAs, 32 skip S MMAs
  unsigned long long* trace;   // debug: CTA 0, slot 0 appends (clock << 16 | code << 8 | item) events: role 0 control warp, 1 first softmax warp
  int trace_cap;
};

#define DHG_ATR(role, code, it)                                                                             \
  do {                                                                                                      \
    if (sh.trace && blockIdx.x == 0 && lane == 0 && tr_n < (unsigned)sh.trace_cap)                          \
      sh.trace[(size_t)(role) * sh.trace_cap + tr_n++] =                                                    \
          ((unsigned long long)clock64() << 16) | ((unsigned long long)(code) << 8) | (unsigned)((it) & 0xff); \
  } while (0)

__device__ __forceinline__ void item_coords(const AttnTcShape& sh, const AttnParams& p, int item, int& qt, int& h, int& b) {
  if (sh.rev) item = sh.items - 1 - item;
  qt = item % sh.QT;
  const int bh = item / sh.QT;
  h = bh % p.H;
  b = bh / p.H;
}

// SPLIT: the fp32-contract mode.  q / k / v / o are in split storage (common.cuh bfs: per 32 elements 32 bf16 hi
// halves, then 32 lo halves), so a head's 64 elements are TWO 64-wide bf16 k-blocks, each (hi x 32 | lo x 32), for Q and K
// alike: S = Q K^T is the six-k-step pattern of the split GEMM (hi.hi, lo.hi, hi.lo) per k-block.  P is written as a
// bf16 hi tile and a lo tile; V's two blocks are two MN-major B operands of 64 columns each, (v_hi x 32 | v_lo x 32), so
// O' = (P_hi + P_lo) V' has 128 fp32 columns and O[d] = O'[hi column of d] + O'[lo column of d]; the row is normalised in
// fp32 and stored split again.
// MAXS: slots per CTA this instance may be launched with.  The block is 160 * slots threads, so the 2-slot shapes
// (level-1 self-attention: 208 score columns per item) get a 200-register budget instead of the 64 of a 960-thread block.
template <bool SPLIT, int MAXS>
__global__ void __launch_bounds__(160 * MAXS, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                            const __grid_constant__ CUtensorMap map_k,
                                                                            const __grid_constant__ CUtensorMap map_v,
                                                                            const AttnTcShape sh, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sh.off_bar);   // per slot: load, s, p, o, free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 * AT_MAX_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = sh.N, NS = sh.NS;
  const int stride = (int)gridDim.x * NS;   // item stride of one slot

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see gemm_tc.cu
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
      for (int s = 0; s < NS; ++s) {
        mbar_init(smem_u32(&bars[s * 5 + 0]), 1);
        mbar_init(smem_u32(&bars[s * 5 + 1]), 1);
        mbar_init(smem_u32(&bars[s * 5 + 2]), 128 * sh.halves);
        mbar_init(smem_u32(&bars[s * 5 + 3]), 1);
        mbar_init(smem_u32(&bars[s * 5 + 4]), 128 * sh.halves);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // q / k / v come from the previous kernel of the chain

  if (warp < NS) {
    // ===== control warp of slot `warp`: one elected lane issues TMA / tcgen05 for the slot's items in turn =====
    const bool leader = elect_one();
    const int s = warp;
    uint8_t* q_s = smem + (size_t)s * sh.slot_bytes;
    uint64_t* sb = bars + s * 5;
    const uint32_t tm = tmem_base + (uint32_t)(s * sh.tmem_cols);
    const uint32_t qlo = umma_desc_lo(smem_u32(q_s)), klo = umma_desc_lo(smem_u32(q_s + sh.off_k));
    const uint32_t vlo_mn = (umma_desc_lo(smem_u32(q_s + sh.off_v)) & ~(1u << 16)) | ((1024u >> 4) << 16);   // MN-major: LBO field
    const uint32_t vlo_mn2 = (umma_desc_lo(smem_u32(q_s + sh.off_v + sh.kv_bytes)) & ~(1u << 16)) | ((1024u >> 4) << 16);   // SPLIT: V block 1
    const int nk = N >> 4;
    uint32_t par = 0;
    unsigned tr_n = 0;
    const bool tr_on = s == 0;
    int tr_i = 0;
    const uint32_t lb = smem_u32(&sb[0]);
    auto issue_load = [&](int it) {
      int qt, h, b;
      item_coords(sh, p, it, qt, h, b);
      if (SPLIT) {   // two 64-wide bf16 blocks per operand: bf16 columns [128 h, 128 h + 64) and [128 h + 64, 128 h + 128)
        mbar_expect_tx(lb, (uint32_t)(2 * 128 * 128 + 4 * N * 128));
        for (int j = 0; j < 2; ++j) {
          tma_load_2d(smem_u32(q_s + j * 16384), &map_q, lb, h * 128 + j * 64, b * p.q_period + p.q_pad + qt * 128);
          tma_load_2d(smem_u32(q_s + sh.off_k + j * sh.kv_bytes), &map_k, lb, h * 128 + j * 64, b * p.k_period + p.k_pad);
          tma_load_2d(smem_u32(q_s + sh.off_v + j * sh.kv_bytes), &map_v, lb, h * 128 + j * 64, b * p.k_period + p.k_pad);
        }
        return;
      }
      mbar_expect_tx(lb, (uint32_t)(128 * 128 + 2 * N * 128));
      // 64-column boxes starting at the head's first column; for D = 48 the last 16 columns belong to
      // the next head (or are zero-filled past the matrix) and are never touched by the MMAs
      tma_load_2d(smem_u32(q_s), &map_q, lb, h * p.D, b * p.q_period + p.q_pad + qt * 128);
      tma_load_2d(smem_u32(q_s + sh.off_k), &map_k, lb, h * p.D, b * p.k_period + p.k_pad);
      tma_load_2d(smem_u32(q_s + sh.off_v), &map_v, lb, h * p.D, b * p.k_period + p.k_pad);
    };
    // The smem tiles of a slot are free as soon as its P V product has completed (bar_o); the TMEM columns only when
    // the softmax threads have read O out (bar_free).  So the next item's Q / K / V are requested right after bar_o and
    // travel while O is normalised and stored; only the next S = Q K^T waits for bar_free.
    if (sh.early && (int)blockIdx.x * NS + s < sh.items && leader) issue_load((int)blockIdx.x * NS + s);
    for (int item = (int)blockIdx.x * NS + s; item < sh.items; item += stride, par ^= 1u) {
      if (!sh.early) {   // load only when the slot is completely free
        mbar_wait(smem_u32(&sb[4]), par ^ 1u);
        if (leader) issue_load(item);
      }
      mbar_wait(lb, par);
      if (tr_on) DHG_ATR(0, 0x02, tr_i);
      if (sh.early) mbar_wait(smem_u32(&sb[4]), par ^ 1u);   // TMEM columns free (previous item's O has been read out)
      tc_fence_after();
      if (tr_on) DHG_ATR(0, 0x03, tr_i);
      // S = Q K^T
      if (leader) {
        if (SPLIT) {   // per k-block (A k-step, B k-step): (0,0) (1,1) hi.hi | (2,0) (3,1) lo.hi | (0,2) (1,3) hi.lo
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              const uint32_t ka = (uint32_t)(k < 4 ? k : k - 4) * 2u, kb = (uint32_t)(k < 2 ? k : k - 2) * 2u;
              umma_bf16(tm, umma_desc_make(qlo + (uint32_t)j * (16384u >> 4) + ka, kDescHiSw128),
                        umma_desc_make(klo + (uint32_t)j * (sh.kv_bytes >> 4) + kb, kDescHiSw128), sh.idesc_s, (j | k) ? 1u : 0u);
            }
        } else {
          for (int k = 0; k < ((sh.dbg & 32) ? 1 : (p.D >> 4)); ++k)
            umma_bf16(tm, umma_desc_make(qlo + 2 * k, kDescHiSw128), umma_desc_make(klo + 2 * k, kDescHiSw128), sh.idesc_s, k ? 1u : 0u);
        }
        umma_commit(smem_u32(&sb[1]));
      }
      __syncwarp();
      // O = P V   (P written by the softmax threads over the Q/K tiles)
      mbar_wait(smem_u32(&sb[2]), par);
      tc_fence_after();
      if (tr_on) DHG_ATR(0, 0x05, tr_i);
      if (leader) {
        for (int kk = 0; kk < ((sh.dbg & 16) ? 1 : nk); ++kk) {
          const uint32_t alo = qlo + (uint32_t)(kk >> 2) * (16384u >> 4) + (uint32_t)(kk & 3) * 2u;
          if (SPLIT) {   // (P_hi + P_lo) against both V blocks: O' columns [0, 64) and [64, 128)
            const uint32_t alo2 = alo + (sh.off_pl >> 4);
            const uint32_t v0 = vlo_mn + (uint32_t)kk * (2048u >> 4), v1 = vlo_mn2 + (uint32_t)kk * (2048u >> 4);
            umma_bf16(tm, umma_desc_make(alo, kDescHiSw128), umma_desc_make(v0, kDescHiSw128), sh.idesc_o, kk ? 1u : 0u);
            umma_bf16(tm, umma_desc_make(alo2, kDescHiSw128), umma_desc_make(v0, kDescHiSw128), sh.idesc_o, 1u);
            umma_bf16(tm + 64u, umma_desc_make(alo, kDescHiSw128), umma_desc_make(v1, kDescHiSw128), sh.idesc_o, kk ? 1u : 0u);
            umma_bf16(tm + 64u, umma_desc_make(alo2, kDescHiSw128), umma_desc_make(v1, kDescHiSw128), sh.idesc_o, 1u);
          } else {
            umma_bf16(tm, umma_desc_make(alo, kDescHiSw128), umma_desc_make(vlo_mn + (uint32_t)kk * (2048u >> 4), kDescHiSw128),
                      sh.idesc_o, kk ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&sb[3]));
      }
      __syncwarp();
      if (sh.early && item + stride < sh.items) {
        mbar_wait(smem_u32(&sb[3]), par);   // P V done: Q / K(P) / V tiles are dead
        if (tr_on) DHG_ATR(0, 0x07, tr_i);
        if (leader) issue_load(item + stride);
        __syncwarp();
      }
      ++tr_i;
    }
  } else {
    // ===== softmax group of slot `slot`: thread = (query row, column half) =====
    // a warp can only touch the TMEM lane quarter (warp id % 4): that fixes which 32 query rows it owns
    const int wps = 4 * sh.halves;                       // softmax warps per slot
    const int slot = (warp - NS) / wps, wis = (warp - NS) - slot * wps;
    const int wq = warp & 3, half = wis >> 2;
    if (slot < NS) {
      uint8_t* q_s = smem + (size_t)slot * sh.slot_bytes;
      float* mask_s = reinterpret_cast<float*>(q_s + sh.off_mask);   // additive mask in log2 units, per key
      float* xchg = reinterpret_cast<float*>(q_s + sh.off_xchg);     // [2 halves][128 rows] max, then [2][128] sum
      uint64_t* sb = bars + slot * 5;
      const int r = wq * 32 + lane;             // query row inside the tile = TMEM lane
      const int gtid = wis * 32 + lane, gthreads = 32 * wps;
      const uint32_t trow = tmem_base + (uint32_t)(slot * sh.tmem_cols) + (((uint32_t)(wq * 32)) << 16);
      const bool masked = p.text != nullptr;
      const int c_lo = sh.halves == 2 ? (half ? (sh.nchunk + 1) / 2 : 0) : 0;
      const int c_hi = sh.halves == 2 ? (half ? sh.nchunk : (sh.nchunk + 1) / 2) : sh.nchunk;
      const int oc_lo = sh.halves == 2 ? half : 0, oc_hi = sh.halves == 2 ? half + 1 : 2;   // 32-column chunks of O
      uint32_t par = 0;
      unsigned tr_n = 0;
      const bool tr_on = slot == 0 && wis == 0;
      int tr_i = 0;
      float v[32];
      for (int item = (int)blockIdx.x * NS + slot; item < sh.items; item += stride, par ^= 1u, ++tr_i) {
        int qt, h, b;
        item_coords(sh, p, item, qt, h, b);
        const int tq = qt * 128 + r;
        if (tr_on) DHG_ATR(1, 0x11, tr_i);
        if (masked) {
          // additive key mask (attention.py:44: mask * -1e9), pre-multiplied by log2(e)
          for (int j = gtid; j < sh.nchunk * 32; j += gthreads)
            mask_s[j] = (j < p.Tk && p.text[(size_t)b * p.Tk + j] == 0) ? -1e9f * 1.4426950408889634f : 0.f;
          asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(gthreads) : "memory");
        }
        mbar_wait(smem_u32(&sb[1]), par);
        tc_fence_after();
        if (tr_on) DHG_ATR(1, 0x12, tr_i);
        // pass 1: row maximum (log2 units); keys >= Tk never count.  Unmasked: max over the raw scores, scaled once
        // (scale > 0); only the last chunk can contain padding keys.
        const int c_full = p.Tk >> 5;   // chunks [0, c_full) hold real keys only
        float mx = (sh.dbg & 1) ? 0.f : -INFINITY;
        for (int c = c_lo; c < ((sh.dbg & 1) ? c_lo : c_hi); ++c) {
          tmem_ld32(trow + c * 32, v);
          if (masked) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int j = c * 32 + i;
              const float t = fmaf(v[i], sh.scale_log2, mask_s[j]);
              if (j < p.Tk) mx = fmaxf(mx, t);
            }
          } else if (c < c_full) {   // four independent chains: the 32-deep dependent one paced the whole pass
            float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]), m2 = fmaxf(v[4], v[5]), m3 = fmaxf(v[6], v[7]);
#pragma unroll
            for (int i = 8; i < 32; i += 8) {
              m0 = fmaxf(m0, fmaxf(v[i], v[i + 1]));
              m1 = fmaxf(m1, fmaxf(v[i + 2], v[i + 3]));
              m2 = fmaxf(m2, fmaxf(v[i + 4], v[i + 5]));
              m3 = fmaxf(m3, fmaxf(v[i + 6], v[i + 7]));
            }
            mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < p.Tk) mx = fmaxf(mx, v[i]);
          }
        }
        if (!masked) mx *= sh.scale_log2;
        if (sh.halves == 2) {
          xchg[half * 128 + r] = mx;
          asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(gthreads) : "memory");
          mx = fmaxf(mx, xchg[(half ^ 1) * 128 + r]);
        }
        if (tr_on) DHG_ATR(1, 0x13, tr_i);
        // pass 2: p = 2^(t - max), row sum, bf16 P into the swizzled K-major operand layout
        float sum = 0.f;
        const float nmx = -mx;
        for (int c = c_lo; c < c_hi; ++c) {
          tmem_ld32(trow + c * 32, v);
          if (masked || c >= c_full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int j = c * 32 + i;
              float t = fmaf(v[i], sh.scale_log2, nmx);
              if (masked) t += mask_s[j];
              float e = t;
              if (!(sh.dbg & 2)) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
              v[i] = j < p.Tk ? e : 0.f;
              sum += v[i];
            }
          } else {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // independent partial sums (see pass 1)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = fmaf(v[i], sh.scale_log2, nmx);
              float e = t;
              if (!(sh.dbg & 2)) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
              v[i] = e;
              if ((i & 3) == 0) s0 += e; else if ((i & 3) == 1) s1 += e; else if ((i & 3) == 2) s2 += e; else s3 += e;
            }
            sum += (s0 + s1) + (s2 + s3);
          }
          const uint32_t blk = smem_u32(q_s) + (uint32_t)(c >> 1) * 16384u + (uint32_t)r * 128u;
#pragma unroll
          for (int g = 0; g < ((sh.dbg & 4) ? 0 : 4); ++g) {
            const int chunk = (c & 1) * 4 + g;   // 16-byte chunk (8 keys) inside the 64-key block row
            if (SPLIT) {   // P = P_hi + P_lo, two bf16 tiles with the same layout
              uint32_t wh[4], wl[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                float r0, r1;
                wh[k2] = bf16_pair_hi(v[g * 8 + 2 * k2], v[g * 8 + 2 * k2 + 1], r0, r1);
                wl[k2] = bf16_pair(r0, r1);
              }
              sts128(blk + ((chunk ^ (r & 7)) << 4), make_uint4(wh[0], wh[1], wh[2], wh[3]));
              sts128(blk + sh.off_pl + ((chunk ^ (r & 7)) << 4), make_uint4(wl[0], wl[1], wl[2], wl[3]));
            } else {
              sts128(blk + ((chunk ^ (r & 7)) << 4),
                     make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7])));
            }
          }
        }
        if (sh.halves == 2) xchg[256 + half * 128 + r] = sum;
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA
        mbar_arrive(smem_u32(&sb[2]));
        if (tr_on) DHG_ATR(1, 0x14, tr_i);
        // O row
        mbar_wait(smem_u32(&sb[3]), par);
        tc_fence_after();
        if (tr_on) DHG_ATR(1, 0x15, tr_i);
        if (sh.halves == 2) sum += xchg[256 + (half ^ 1) * 128 + r];   // written before the partner's bar_p arrive, which precedes bar_o
        const float inv = 1.f / sum;
        if (SPLIT) {   // O' = [sum p v_hi (0..31) | sum p v_lo (0..31) | v_hi (32..63) | v_lo (32..63)]: 2 groups of 32 elements
          char* og = reinterpret_cast<char*>(p.o) + (((size_t)b * p.q_period + p.q_pad + tq) * p.o_pitch + (size_t)h * 64) * 4;
          float w[32];
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            tmem_ld32(trow + g * 64, v);
            tmem_ld32(trow + g * 64 + 32, w);
            if (tq < p.Tq) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  const int i = q4 * 8 + 2 * k2;
                  float r0, r1;
                  wh[k2] = bf16_pair_hi((v[i] + w[i]) * inv, (v[i + 1] + w[i + 1]) * inv, r0, r1);
                  wl[k2] = bf16_pair(r0, r1);
                }
                *reinterpret_cast<uint4*>(og + g * 128 + q4 * 16) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(og + g * 128 + 64 + q4 * 16) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
              }
            }
          }
        }
        bf16* orow = reinterpret_cast<bf16*>(p.o) + ((size_t)b * p.q_period + p.q_pad + tq) * p.o_pitch + h * p.D;
        const bool wide = (p.o_pitch & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.o) | (uintptr_t)(h * p.D * 2)) & 31) == 0 && (p.D & 15) == 0;
        for (int c = oc_lo; c < (SPLIT ? oc_lo : oc_hi); ++c) {
          tmem_ld32(trow + c * 32, v);
          if (tq < p.Tq && !(sh.dbg & 8)) {
            if (wide) {   // 32-byte stores: whole sectors, half the store instructions
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                if (c * 32 + g * 16 >= p.D) break;
                uint32_t w8[8];
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) w8[k2] = pack_bf16x2(v[g * 16 + 2 * k2] * inv, v[g * 16 + 2 * k2 + 1] * inv);
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + c * 32 + g * 16), "r"(w8[0]), "r"(w8[1]),
                             "r"(w8[2]), "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                             : "memory");
              }
              continue;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (c * 32 + g * 8 >= p.D) break;
              const uint4 u = make_uint4(pack_bf16x2(v[g * 8] * inv, v[g * 8 + 1] * inv), pack_bf16x2(v[g * 8 + 2] * inv, v[g * 8 + 3] * inv),
                                         pack_bf16x2(v[g * 8 + 4] * inv, v[g * 8 + 5] * inv), pack_bf16x2(v[g * 8 + 6] * inv, v[g * 8 + 7] * inv));
              *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = u;
            }
          }
        }
        // the slot's smem tiles and TMEM columns may be overwritten by the next item.  (Handing the slot back BEFORE the
        // stores, with the packed row held in registers, was measured: -2 % with 2 slots, spills with 6: not kept.)
        tc_fence_before();
        mbar_arrive(smem_u32(&sb[4]));
        if (tr_on) DHG_ATR(1, 0x16, tr_i);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Long key sequences (Tk > 256: the self-attentions of long lines, e.g. T = 1200 -> 600 / 300 keys).  Same work item
// (128-query tile, head, sample), but the keys are walked in blocks of 128 and the softmax is taken in TWO PASSES over
// the blocks, so nothing is ever rescaled:
//   pass A   S_j = Q K_j^T for every block j, row maximum over all keys (the score MMAs are cheap: D = 64);
//   pass B   S_j again, p = 2^(s - max) -> row sum and bf16 P_j in shared memory, O += P_j V_j accumulated in TMEM.
// Per slot: Q tile, two K block buffers (the next block travels while this one is multiplied and read), one V block,
// one P tile; TMEM: 128 score columns + 64 output columns.  Two slots per CTA work on different items, so one slot's
// loads and MMAs hide behind the other's softmax.  Unmasked, head depth 64 (self-attention only needs that).
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATL_KB = 128;      // keys per block
constexpr int ATL_SLOTS = 2;
constexpr int ATL_TMEM_SLOT = 192;   // 128 score + 64 output columns

struct AttnLongShape {
  int nkb;       // key blocks
  int QT, items, rev;
  uint32_t idesc_s, idesc_o;
  uint32_t off_k, off_v, off_p, slot_bytes, off_bar;
  float scale_log2;
};

__global__ void __launch_bounds__(ATL_SLOTS * 160, 1) attn_tc_long_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                          const __grid_constant__ CUtensorMap map_k,
                                                                          const __grid_constant__ CUtensorMap map_v,
                                                                          const AttnLongShape sh, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // per slot: 0 full_q, 1-2 full_k[2], 3 full_v, 4 bar_s, 5 s_free, 6 bar_p, 7 bar_o, 8 slot_free
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sh.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9 * ATL_SLOTS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NS = ATL_SLOTS;
  const int stride = (int)gridDim.x * NS;
  const int nkb = sh.nkb;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
      for (int s = 0; s < NS; ++s) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[s * 9 + i]), 1);
        mbar_init(smem_u32(&bars[s * 9 + 5]), 128);
        mbar_init(smem_u32(&bars[s * 9 + 6]), 128);
        mbar_init(smem_u32(&bars[s * 9 + 7]), 1);
        mbar_init(smem_u32(&bars[s * 9 + 8]), 128);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < NS) {
    // ===== control warp of slot `warp` =====
    const bool leader = elect_one();
    const int s = warp;
    uint8_t* q_s = smem + (size_t)s * sh.slot_bytes;
    uint64_t* sb = bars + s * 9;
    const uint32_t tm_s = tmem_base + (uint32_t)(s * ATL_TMEM_SLOT), tm_o = tm_s + 128u;
    const uint32_t qlo = umma_desc_lo(smem_u32(q_s));
    const uint32_t klo[2] = {umma_desc_lo(smem_u32(q_s + sh.off_k)), umma_desc_lo(smem_u32(q_s + sh.off_k + 16384u))};
    const uint32_t vlo_mn = (umma_desc_lo(smem_u32(q_s + sh.off_v)) & ~(1u << 16)) | ((1024u >> 4) << 16);   // MN-major: LBO field
    const uint32_t plo = umma_desc_lo(smem_u32(q_s + sh.off_p));
    uint32_t ph_q = 0, ph_k[2] = {0, 0}, ph_v = 0, ph_sf = 0, ph_p = 0, ph_o = 0, ph_free = 0;
    for (int item = (int)blockIdx.x * NS + s; item < sh.items; item += stride) {
      int it2 = sh.rev ? sh.items - 1 - item : item;
      const int qt = it2 % sh.QT, bh = it2 / sh.QT, h = bh % p.H, b = bh / p.H;
      const int krow0 = b * p.k_period + p.k_pad;
      auto kload = [&](int jblk, int buf) {
        mbar_expect_tx(smem_u32(&sb[1 + buf]), 16384u);
        tma_load_2d(smem_u32(q_s + sh.off_k + (uint32_t)buf * 16384u), &map_k, smem_u32(&sb[1 + buf]), h * 64, krow0 + jblk * ATL_KB);
      };
      mbar_wait(smem_u32(&sb[8]), ph_free ^ 1u);   // the previous item's O has been read out: every tile and TMEM column of the slot is free
      ph_free ^= 1u;
      if (leader) {
        mbar_expect_tx(smem_u32(&sb[0]), 16384u);
        tma_load_2d(smem_u32(q_s), &map_q, smem_u32(&sb[0]), h * 64, b * p.q_period + p.q_pad + qt * 128);
        kload(0, 0);
      }
      mbar_wait(smem_u32(&sb[0]), ph_q);
      ph_q ^= 1u;
      const int nsteps = 2 * nkb;
      for (int g = 0; g < nsteps; ++g) {
        const int buf = g & 1, j = g < nkb ? g : g - nkb;
        const bool pass_b = g >= nkb;
        if (g > 0) {   // the softmax group has read S of step g-1 (so that MMA is complete and its K buffer is free)
          mbar_wait(smem_u32(&sb[5]), ph_sf);
          ph_sf ^= 1u;
        }
        if (leader) {
          if (g + 1 < nsteps) kload(g + 1 < nkb ? g + 1 : g + 1 - nkb, buf ^ 1);
          if (pass_b) {   // V block (its buffer is free: the previous P V product has completed, see the wait below)
            mbar_expect_tx(smem_u32(&sb[3]), 16384u);
            tma_load_2d(smem_u32(q_s + sh.off_v), &map_v, smem_u32(&sb[3]), h * 64, krow0 + j * ATL_KB);
          }
        }
        mbar_wait(smem_u32(&sb[1 + buf]), ph_k[buf]);
        ph_k[buf] ^= 1u;
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tm_s, umma_desc_make(qlo + 2 * k, kDescHiSw128), umma_desc_make(klo[buf] + 2 * k, kDescHiSw128), sh.idesc_s, k ? 1u : 0u);
          umma_commit(smem_u32(&sb[4]));
        }
        __syncwarp();
        if (pass_b) {
          mbar_wait(smem_u32(&sb[3]), ph_v);
          ph_v ^= 1u;
          mbar_wait(smem_u32(&sb[6]), ph_p);   // P_j written
          ph_p ^= 1u;
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < ATL_KB / 16; ++kk) {
              const uint32_t alo = plo + (uint32_t)(kk >> 2) * (16384u >> 4) + (uint32_t)(kk & 3) * 2u;
              umma_bf16(tm_o, umma_desc_make(alo, kDescHiSw128), umma_desc_make(vlo_mn + (uint32_t)kk * (2048u >> 4), kDescHiSw128),
                        sh.idesc_o, (j | kk) ? 1u : 0u);
            }
            umma_commit(smem_u32(&sb[7]));
          }
          __syncwarp();
          mbar_wait(smem_u32(&sb[7]), ph_o);   // P V done: the P tile and the V block may be overwritten
          ph_o ^= 1u;
        }
      }
      // the last step's S is read by the softmax group before it arrives on s_free one more time: consume that phase
      mbar_wait(smem_u32(&sb[5]), ph_sf);
      ph_sf ^= 1u;
    }
  } else {
    // ===== softmax group of slot `slot`: thread = query row =====
    const int slot = (warp - NS) >> 2, wq = warp & 3;
    uint8_t* q_s = smem + (size_t)slot * sh.slot_bytes;
    uint64_t* sb = bars + slot * 9;
    const int r = wq * 32 + lane;
    const uint32_t lane_sel = ((uint32_t)(wq * 32)) << 16;
    const uint32_t tm_s = tmem_base + (uint32_t)(slot * ATL_TMEM_SLOT) + lane_sel, tm_o = tm_s + 128u;
    const uint32_t p_sa = smem_u32(q_s + sh.off_p);
    uint32_t ph_s = 0, ph_o = 0;
    float v[32];
    for (int item = (int)blockIdx.x * NS + slot; item < sh.items; item += stride) {
      int it2 = sh.rev ? sh.items - 1 - item : item;
      const int qt = it2 % sh.QT, bh = it2 / sh.QT, h = bh % p.H, b = bh / p.H;
      const int tq = qt * 128 + r;
      // pass A: row maximum of the raw scores (scale > 0: scaled once at the end)
      float mx = -INFINITY;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(smem_u32(&sb[4]), ph_s);
        ph_s ^= 1u;
        tc_fence_after();
        const int k0 = j * ATL_KB;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (k0 + c * 32 >= p.Tk) break;
          tmem_ld32(tm_s + c * 32, v);
          if (k0 + c * 32 + 32 <= p.Tk) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (k0 + c * 32 + i < p.Tk) mx = fmaxf(mx, v[i]);
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&sb[5]));
      }
      const float nmx = -mx * sh.scale_log2;
      // pass B: p = 2^(s * scale_log2 - max), row sum, bf16 P block in the swizzled K-major operand layout
      float sum = 0.f;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(smem_u32(&sb[4]), ph_s);
        ph_s ^= 1u;
        tc_fence_after();
        const int k0 = j * ATL_KB;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const bool any = k0 + c * 32 < p.Tk;
          if (any) tmem_ld32(tm_s + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float t = fmaf(v[i], sh.scale_log2, nmx);
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
            v[i] = (any && k0 + c * 32 + i < p.Tk) ? e : 0.f;
            sum += v[i];
          }
          const uint32_t blk = p_sa + (uint32_t)(c >> 1) * 16384u + (uint32_t)r * 128u;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = (c & 1) * 4 + g;
            sts128(blk + ((chunk ^ (r & 7)) << 4),
                   make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                              pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7])));
          }
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(smem_u32(&sb[6]));   // P_j ready
        mbar_arrive(smem_u32(&sb[5]));   // S read
      }
      // O: nkb completions of bar_o per item, the last one is the finished accumulator
      ph_o ^= (uint32_t)((nkb - 1) & 1);
      mbar_wait(smem_u32(&sb[7]), ph_o);
      ph_o ^= 1u;
      tc_fence_after();
      const float inv = 1.f / sum;
      bf16* orow = reinterpret_cast<bf16*>(p.o) + ((size_t)b * p.q_period + p.q_pad + tq) * p.o_pitch + h * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(tm_o + c * 32, v);
        if (tq < p.Tq) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) =
                make_uint4(pack_bf16x2(v[g * 8] * inv, v[g * 8 + 1] * inv), pack_bf16x2(v[g * 8 + 2] * inv, v[g * 8 + 3] * inv),
                           pack_bf16x2(v[g * 8 + 4] * inv, v[g * 8 + 5] * inv), pack_bf16x2(v[g * 8 + 6] * inv, v[g * 8 + 7] * inv));
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sb[8]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

int g_attn_dbg = 0, g_attn_halves = 1, g_attn_pdl = 1, g_attn_early = 1, g_attn_max_slots = AT_MAX_SLOTS;
void attn_tc_set_debug(int v) { if (v <= -300 && v >= -306) g_attn_max_slots = -300 - v > 0 ? -300 - v : AT_MAX_SLOTS; else if (v == -200 || v == -201) g_attn_early = v == -201; else if (v >= 0) g_attn_dbg = v; else if (v <= -100) g_attn_pdl = v == -101; else g_attn_halves = -v; }

using AttnTcFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnTcShape, const AttnParams);
// the instance with the smallest thread bound (= largest register budget) that covers the block
static AttnTcFn attn_tc_instance(bool split, int threads) {
  if (threads <= 160 * 2) return split ? attn_tc_kernel<true, 2> : attn_tc_kernel<false, 2>;
  if (threads <= 160 * 4) return split ? attn_tc_kernel<true, 4> : attn_tc_kernel<false, 4>;
  return split ? attn_tc_kernel<true, AT_MAX_SLOTS> : attn_tc_kernel<false, AT_MAX_SLOTS>;
}

struct AttnTcPlan {
  AttnTcFn fn = nullptr;
  CUtensorMap map_q, map_k, map_v;
  bool long_keys = false;   // Tk > 256: attn_tc_long_kernel (key blocks, two-pass softmax)
  AttnLongShape lsh;
  AttnTcShape sh;
  AttnParams p;
  dim3 grid;
  int threads;
  size_t smem;
  int pdl;   // programmatic dependent launch, as the option stood when the plan was built
};

static bool attn_tc_long_capable(const AttnParams& p) { return p.D == 64 && p.Tk > ATL_KB && p.text == nullptr && !p.split; }
static bool attn_tc_long(const AttnParams& p) { return attn_tc_long_capable(p) && p.Tk > 256; }
bool attn_tc_supported(const AttnParams& p) {
  if (p.split)   // split storage: head depth 64 only (a head = two groups of 32 elements), all keys at once
    return p.D == 64 && p.Tk >= 1 && p.Tk <= 256 && p.q_pitch % 32 == 0 && p.k_pitch % 32 == 0 && p.v_pitch % 32 == 0 && p.o_pitch % 32 == 0;
  return (((p.D == 64 || p.D == 48) && p.Tk >= 1 && p.Tk <= 256) || attn_tc_long(p)) && p.q_pitch % 8 == 0 && p.k_pitch % 8 == 0 &&
         p.v_pitch % 8 == 0 && p.o_pitch % 8 == 0;
}

AttnTcPlan* attn_tc_plan_create(const AttnParams& p, int q_rows, int k_rows, char* err, int errlen, int prefer_long) {
  if (!attn_tc_supported(p)) { snprintf(err, errlen, "attention shape not supported by the tcgen05 kernel (D=%d Tk=%d)", p.D, p.Tk); return nullptr; }
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  AttnTcPlan* a = new AttnTcPlan();
  a->p = p;
  a->pdl = g_attn_pdl;
  if (prefer_long && !attn_tc_long_capable(p)) { snprintf(err, errlen, "the key-block kernel needs unmasked keys, head depth 64 and Tk > 128"); delete a; return nullptr; }
  if (attn_tc_long(p) || prefer_long) {
    a->long_keys = true;
    AttnLongShape& l = a->lsh;
    l.nkb = (p.Tk + ATL_KB - 1) / ATL_KB;
    l.QT = (p.Tq + 127) / 128;
    l.items = p.B * p.H * l.QT;
    l.rev = 0;
    const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    l.idesc_s = base | ((uint32_t)(ATL_KB >> 3) << 17);
    l.idesc_o = base | (1u << 16) | ((uint32_t)(64 >> 3) << 17);   // B = V is MN-major, N = head depth
    l.scale_log2 = p.scale * 1.4426950408889634f;
    l.off_k = 16384u; l.off_v = 16384u + 2u * 16384u; l.off_p = l.off_v + 16384u;
    l.slot_bytes = l.off_p + 32768u;
    l.off_bar = ATL_SLOTS * l.slot_bytes;
    a->smem = l.off_bar + (9 * ATL_SLOTS + 2) * 8 + 1024;
    const int ctas = (l.items + ATL_SLOTS - 1) / ATL_SLOTS;
    a->grid = dim3(ctas < num_sms ? ctas : num_sms);
    a->threads = ATL_SLOTS * 160;
    const uint64_t cols = (uint64_t)p.H * p.D;
    if (!make_map(&a->map_q, p.q, (uint64_t)q_rows, cols, (uint64_t)p.q_pitch, 128, err, errlen) ||
        !make_map(&a->map_k, p.k, (uint64_t)k_rows, cols, (uint64_t)p.k_pitch, ATL_KB, err, errlen) ||
        !make_map(&a->map_v, p.v, (uint64_t)k_rows, cols, (uint64_t)p.v_pitch, ATL_KB, err, errlen)) {
      delete a;
      return nullptr;
    }
    cudaError_t ce = cudaFuncSetAttribute(attn_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); delete a; return nullptr; }
    return a;
  }
  AttnTcShape& sh = a->sh;
  sh.N = (p.Tk + 15) & ~15;
  sh.nblk = (sh.N + 63) / 64;
  sh.nchunk = (sh.N + 31) / 32;
  int cols = sh.nchunk * 32 < 64 ? 64 : sh.nchunk * 32;
  if (p.split && cols < 128) cols = 128;   // O' has 128 columns (hi and lo column of every output element)
  sh.tmem_cols = cols <= 64 ? 64 : (cols + 31) & ~31;   // the CTA owns all 512 columns and cuts them itself: any multiple of a 32-column chunk (text/style: 96 -> 5 slots instead of 4)
  sh.QT = (p.Tq + 127) / 128;
  sh.items = p.B * p.H * sh.QT;
  sh.rev = 0;
  sh.early = g_attn_early;
  // c_format F32 [4,6) | a,b BF16 [7,10),[10,13) | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
  sh.idesc_s = base | ((uint32_t)(sh.N >> 3) << 17);
  sh.idesc_o = base | (1u << 16) | ((uint32_t)((p.split ? 64 : p.D) >> 3) << 17);   // B = V is MN-major, N = head depth (split: one 64-column block)
  sh.scale_log2 = p.scale * 1.4426950408889634f;
  sh.dbg = g_attn_dbg;
  sh.trace = nullptr; sh.trace_cap = 0;
  const uint32_t kv_bytes = (uint32_t)sh.N * 128u;
  const uint32_t nb = p.split ? 2u : 1u;              // 64-wide bf16 blocks per operand (split storage: hi|lo groups)
  sh.kv_bytes = kv_bytes;
  sh.off_pl = (uint32_t)sh.nblk * 16384u;             // P_lo behind P_hi (split)
  uint32_t pq = nb * (16384u + kv_bytes);             // Q | K
  if (pq < nb * (uint32_t)sh.nblk * 16384u) pq = nb * (uint32_t)sh.nblk * 16384u;   // overlaid by P (split: P_hi | P_lo)
  sh.off_k = nb * 16384u;
  sh.off_v = pq;
  sh.off_mask = pq + nb * kv_bytes;
  sh.off_xchg = sh.off_mask + (uint32_t)sh.nchunk * 32u * 4u;
  sh.slot_bytes = (sh.off_xchg + 4u * 128u * 4u + 1023u) & ~1023u;
  sh.halves = (g_attn_halves == 2 && sh.nchunk >= 4) ? 2 : 1;   // splitting the columns did not pay off (measured): off by default
  // slots per CTA: bounded by TMEM (512 columns), shared memory and the 1024-thread limit
  int ns = 512 / sh.tmem_cols;
  while (ns > 1 && (32 + 128 * sh.halves) * ns > 160 * AT_MAX_SLOTS) --ns;
  const size_t bar_bytes = (5 * AT_MAX_SLOTS + 2) * 8;
  while (ns > 1 && (size_t)ns * sh.slot_bytes + bar_bytes + 1024 > (size_t)227 * 1024) --ns;
  if (ns > AT_MAX_SLOTS) ns = AT_MAX_SLOTS;
  if (ns > g_attn_max_slots) ns = g_attn_max_slots;   // experiment: fewer slots = more registers per thread
  while (ns > 1 && (sh.items + ns - 1) / ns < num_sms) --ns;   // small problems: spread over the SMs first
  sh.NS = ns;
  sh.off_bar = (uint32_t)ns * sh.slot_bytes;
  a->smem = sh.off_bar + bar_bytes + 1024;
  if (a->smem > (size_t)227 * 1024) { snprintf(err, errlen, "attention tile does not fit in shared memory (Tk=%d)", p.Tk); delete a; return nullptr; }
  const int ctas = (sh.items + ns - 1) / ns;
  a->grid = dim3(ctas < num_sms ? ctas : num_sms);
  a->threads = (32 + 128 * sh.halves) * ns;
  const uint64_t em = p.split ? 2 : 1;   // bf16 numbers per element (split storage: the row matrices are viewed as bf16 [rows, 2 * pitch])
  const uint64_t qcols = em * p.H * p.D, kcols = em * p.H * p.D;
  if (!make_map(&a->map_q, p.q, (uint64_t)q_rows, qcols, em * p.q_pitch, 128, err, errlen) ||
      !make_map(&a->map_k, p.k, (uint64_t)k_rows, kcols, em * p.k_pitch, (uint32_t)sh.N, err, errlen) ||
      !make_map(&a->map_v, p.v, (uint64_t)k_rows, kcols, em * p.v_pitch, (uint32_t)sh.N, err, errlen)) {
    delete a;
    return nullptr;
  }
  a->fn = attn_tc_instance(p.split != 0, a->threads);
  cudaError_t ce = cudaFuncSetAttribute(a->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (ce != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); delete a; return nullptr; }
  return a;
}

void attn_tc_plan_destroy(AttnTcPlan* a) { delete a; }
bool attn_tc_plan_is_long(const AttnTcPlan* a) { return a->long_keys; }
void attn_tc_plan_set_reverse(AttnTcPlan* a, int rev) { a->sh.rev = rev ? 1 : 0; a->lsh.rev = rev ? 1 : 0; }
void attn_tc_plan_set_early_load(AttnTcPlan* a, int on) { a->sh.early = on ? 1 : 0; }
void attn_tc_plan_set_trace(AttnTcPlan* a, unsigned long long* buf, int cap) { a->sh.trace = buf; a->sh.trace_cap = cap; }
int attn_tc_plan_slots(const AttnTcPlan* a) { return a->long_keys ? ATL_SLOTS : a->sh.NS; }

int attn_tc_launch(const AttnTcPlan* a, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = a->grid;
  cfg.blockDim = dim3(a->threads);
  cfg.dynamicSmemBytes = a->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a->pdl ? 1 : 0;
  if (a->long_keys) return cudaLaunchKernelEx(&cfg, attn_tc_long_kernel, a->map_q, a->map_k, a->map_v, a->lsh, a->p) == cudaSuccess ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, a->fn, a->map_q, a->map_k, a->map_v, a->sh, a->p) == cudaSuccess ? 0 : 1;
}

}  // namespace dhg
