// tcgen05 / TMEM / TMA GEMM with the fused row epilogue, sm_100a.  Persistent, warp-specialised.
//
//   out[m, n] = epilogue( sum_{tap<taps} sum_{k<K} A[m + tap - taps/2, k] * W[tap][n][k] )
//
// One CTA per SM loops over 128 x BN output tiles (BN <= 256, or 384 = two N=192 MMAs when a
// LayerNorm needs the whole 384-wide row).  Warp roles (320 threads):
//   warp  0    TMA producer: A tile [128 rows x 64 k] and W tile [BN x 64 k] per k-block, SWIZZLE_128B;
//              row-shifted A coordinates implement the conv taps, TMA zero-fill implements every edge.
//              Runs ahead across tiles through a ring of smem stages.
//   warp  1    TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, bf16 x bf16 -> fp32 in
//              TMEM).  Two accumulator stages (2 x 256 columns) when BN <= 256, so the MMAs of tile
//              i+1 overlap the epilogue of tile i.
//   warps 2+   epilogue (16 warps): warp (q, part) owns TMEM lanes [32q, 32q+32) and one part of the tile's
//              32-column chunks; thread = one accumulator row.  Residual / per-position-bias rows are
//              prefetched with cp.async into a per-warp swizzled ring (coalesced global reads), the
//              bf16 outputs go through a per-warp swizzled staging tile and leave as coalesced 16-byte
//              stores.  Bias / FiLM vectors live in shared memory for the whole kernel.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> epilogue).
//
// The epilogue is the one documented in common.cuh (bias|rowbias, residual, LayerNorm, FiLM, residual,
// halo-row zeroing, raw and/or SiLU'd bf16 stores).  LayerNorm: pass 1 adds bias/residual, writes the
// row back to TMEM and accumulates shifted sums per column half, the halves are merged with Chan's
// formula through shared memory; pass 2 normalises.  Reference ops fused here: Linear/Conv1d
// (cnn.py:32-49, attention.py:58-61), LayerNorm (model.py:25), AffineTransformLayer
// (conditioning.py:16-19), SiLU (cnn.py:25), residual adds and nearest upsample (model.py:169-176).
#include "kernels.h"
#include "tc_common.cuh"

namespace dhg {

using namespace tc;

namespace {

constexpr int TC_BM = 128;
#ifndef DHG_EPI_PARTS
#define DHG_EPI_PARTS 2
#endif
constexpr int EPI_PARTS = DHG_EPI_PARTS;              // epilogue warps per TMEM lane quarter (column split)
constexpr int EPI_WARPS = 4 * EPI_PARTS;
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;
constexpr int AUX_DEPTH = 8 / EPI_PARTS;              // aux ring slots per epilogue warp
constexpr int AUX_SLOT_BYTES = 2048;                  // 32 rows x 32 bf16
constexpr int AUX_RING_BYTES = AUX_DEPTH * AUX_SLOT_BYTES;
constexpr int OUT_STAGE_BYTES = (DHG_EPI_PARTS == 4 ? 2048 : 4096);  // per epilogue warp: 2 x (32 rows x 32 bf16), alternating between TMA stores
constexpr int TMEM_COLS = 512;

enum { AUX_NONE = 0, AUX_RES_PRE = 1, AUX_RES_POST = 2, AUX_RES_POST_UP = 3, AUX_ROWBIAS = 4 };

struct TcShape {
  int rows, K, N, taps;   // taps: 1 linear, 3 conv (row-shifted A tile)
  int base_taps;   // == taps (row shifts of the A operand)
  int kb2;         // dual-operand mode: k-blocks of the SECOND A matrix (3-tap weights, rows shifted -1..+1) that follow the
                   // kb_per_tap k-blocks of the first one (1 tap) into the same accumulators; 0 = single operand
  uint32_t a2_tx_bytes;   // bytes of one 130-row tile of the second A matrix
  int sio;         // split I/O (Epilogue::split_io): activations are bfs pairs; K, lda and the A map are in bf16 units (2 per element)
  int BN;          // tile width
  int n_groups;    // N / BN
  int m_tiles;
  int G;           // row tiles per super-tile (independent accumulators interleaved by the MMA warp)
  int m_super;     // ceil(m_tiles / G)
  int rev;         // walk the super-tiles from the last row to the first (see tc_gemm_plan_set_reverse)
  int direct_store;   // bf16 outputs: each thread stores its 32 columns as two 32-byte st.global (no smem staging, no TMA store)
  int a_evict_first;   // A rows are read once by this launch: give them L2 evict-first priority
  int split;           // column-split LayerNorm: 2-CTA cluster per row tile, rank = column half
  int dot_n;           // dot mode: 3 * N floats of dot vectors in shared memory (0 = normal stores)
  int stages_a, stages_w;   // A ring / W ring depth (W ring unused when w_resident)
  int w_resident;  // all W tiles of this CTA's column group stay in smem for the whole kernel
  int sticky;      // each CTA works on one column group only
  int out_bufs;    // staging tiles per epilogue warp for the TMA stores (2, or 1 when shared memory is tight)
  int pair;        // cta_group::2: the two CTAs of a cluster share every MMA (M = 256: 128 rows each) and each loads half of W
  int grp_cta0[13];   // sticky: first CTA of each column group (n_groups + 1 entries)
  int kb_per_tap;  // ceil(K / 64)
  int umma_n;      // N of one tcgen05.mma (BN, or 192 when BN == 384)
  int n_umma;      // MMAs per k-step along N (1 or 2)
  uint32_t idesc;
  int acc_stages;  // 1 or 2 TMEM accumulator stages
  int aux_kind;
  int vec_bias_n;  // floats of bias staged in smem (0 or N)
  int film_n;      // floats of gamma / beta staged in smem (0 or N)
  uint32_t a_stage_bytes, a_tx_bytes, w_tile_bytes, off_w, off_aux, off_out, off_vec, off_ln, off_bar;
  unsigned long long* trace;   // debug: CTA 0 appends (clock << 16 | code << 8 | tile) events; trace[0] = count
  int trace_cap;
};

// fire-and-forget store into the role's private region (role 0 producer, 1 MMA, 2 first epilogue warp): no atomics,
// no round trip, so the traced warp is barely perturbed
#define DHG_TR(code, tile)                                                                        \
  do {                                                                                            \
    if (sh.trace && blockIdx.x == 0 && lane == 0 && tr_n < (unsigned)sh.trace_cap)                \
      sh.trace[(size_t)tr_role * sh.trace_cap + tr_n++] =                                         \
          ((unsigned long long)clock64() << 16) | ((unsigned long long)(code) << 8) | (unsigned)((tile) & 0xff); \
  } while (0)

// per-chunk points (tcgen05.ld / math / store phases of tools/trace_summary.py): only in a -DDHG_TRACE_FINE build, they
// cost ~5 executed instructions each inside the hottest loop of the library even when no trace is taken
#ifdef DHG_TRACE_FINE
#define DHG_TR_FINE(code, tile) DHG_TR(code, tile)
#else
#define DHG_TR_FINE(code, tile) do { } while (0)
#endif

// m / period and m % period for 0 <= m < 2^24 (checked at plan time) without the ~40-instruction integer division:
// float estimate, exact after one correction step either way.
__device__ __forceinline__ void fast_divmod(int m, int period, float inv_period, int& q, int& r) {
  q = __float2int_rz(__int2float_rz(m) * inv_period);
  r = m - q * period;
  if (r < 0) { r += period; --q; }
  if (r >= period) { r -= period; ++q; }
}

// One k-block of the MMA warp's main loop, specialised on the tap count and on the number of interleaved accumulators so
// that everything inside it is straight-line code (the single issuing warp is latency-critical: every instruction
// between two tcgen05.mma shows up in the tile time).  `first` = 0: the first MMA overwrites the accumulators.
// Resident W: tap t of this k-block is tile w_tile0 + t * w_tap_stride of the resident set.
template <int TAPS, int G, bool PAIR, bool SIO = false>
__device__ __forceinline__ void mma_kblock(const TcShape& sh, const bool leader, const uint32_t acc, const uint32_t a_ring_addr,
                                           const uint32_t w_addr, uint64_t* full_a, uint64_t* empty_a, uint64_t* full_w, uint64_t* empty_w,
                                           uint32_t& sa, uint32_t& pa, uint32_t& sw, uint32_t& pw, const uint32_t first, const int w_tile0,
                                           const int w_tap_stride, const int it, unsigned& tr_n, const int tr_role, const int lane) {
  // Split I/O: the 64-wide k-block holds the hi (k-steps 0, 1) and lo (k-steps 2, 3) halves of 32 elements, in the A
  // tile and in the W tile alike; x . w = hi . w_hi + lo . w_hi + hi . w_lo is six k-steps (a-step, b-step) instead of four.
  constexpr int NK = SIO ? 6 : TC_BK / 16;
  const uint32_t nb2 = (uint32_t)(PAIR ? sh.umma_n / 2 : sh.umma_n) * TC_BK * 2;   // byte offset of the second N half (BN = 384)
  const bool two_n = sh.n_umma == 2;
  const bool resident = sh.w_resident != 0;
  uint32_t alo[G], slot[G];
#pragma unroll
  for (int sub = 0; sub < G; ++sub) {
    mbar_wait(smem_u32(&full_a[sa]), pa);
    slot[sub] = sa;
    alo[sub] = umma_desc_lo(a_ring_addr + sa * sh.a_stage_bytes);
    if (++sa == (uint32_t)sh.stages_a) { sa = 0; pa ^= 1u; }
  }
  tc_fence_after();
  DHG_TR(0x21, it);
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
    uint32_t b_addr;
    if (resident) {
      b_addr = w_addr + (uint32_t)(w_tile0 + tap * w_tap_stride) * sh.w_tile_bytes;
    } else {
      mbar_wait(smem_u32(&full_w[sw]), pw);
      tc_fence_after();
      b_addr = w_addr + sw * sh.w_tile_bytes;
    }
    const uint32_t blo = umma_desc_lo(b_addr);
    if (leader) {
      // shifted tap: logical row r of the A operand is physical row r + tap of the 130-row tile (the 128B
      // swizzle is a function of the absolute smem address, so a +128 B start address just works)
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        // (A k-step, B k-step) in 16-byte descriptor units: plain 0..3 | split: (0,0) (1,1) (2,0) (3,1) (0,2) (1,3)
        const uint32_t ka = SIO ? (uint32_t)(k < 4 ? k : k - 4) * 2u : (uint32_t)k * 2u;
        const uint32_t kb = SIO ? (uint32_t)(k < 2 ? k : k < 4 ? k - 2 : k - 2) * 2u : (uint32_t)k * 2u;
        const uint32_t accum = (tap | k) ? 1u : first;
        if (!two_n) {
#pragma unroll
          for (int sub = 0; sub < G; ++sub) {
            if (PAIR)
              umma_bf16_pair(acc, umma_desc_make(alo[0] + tap * 8 + ka, kDescHiSw128), umma_desc_make(blo + kb, kDescHiSw128), sh.idesc, accum);
            else
              umma_bf16(acc + (uint32_t)(sub * sh.BN), umma_desc_make(alo[sub] + tap * 8 + ka, kDescHiSw128),
                        umma_desc_make(blo + kb, kDescHiSw128), sh.idesc, accum);
          }
        } else {
          const uint32_t blo2 = umma_desc_lo(b_addr + nb2);
          if (PAIR) {
            umma_bf16_pair(acc, umma_desc_make(alo[0] + tap * 8 + ka, kDescHiSw128), umma_desc_make(blo + kb, kDescHiSw128), sh.idesc, accum);
            umma_bf16_pair(acc + (uint32_t)sh.umma_n, umma_desc_make(alo[0] + tap * 8 + ka, kDescHiSw128),
                           umma_desc_make(blo2 + kb, kDescHiSw128), sh.idesc, accum);
          } else {
            umma_bf16(acc, umma_desc_make(alo[0] + tap * 8 + ka, kDescHiSw128), umma_desc_make(blo + kb, kDescHiSw128), sh.idesc, accum);
            umma_bf16(acc + (uint32_t)sh.umma_n, umma_desc_make(alo[0] + tap * 8 + ka, kDescHiSw128),
                      umma_desc_make(blo2 + kb, kDescHiSw128), sh.idesc, accum);
          }
        }
      }
      if (!resident) { if (PAIR) umma_commit_pair(smem_u32(&empty_w[sw])); else umma_commit(smem_u32(&empty_w[sw])); }   // frees the W slot (in both CTAs of a pair) when these MMAs retire
    }
    __syncwarp();
    if (!resident && ++sw == (uint32_t)sh.stages_w) { sw = 0; pw ^= 1u; }
  }
  if (leader) {
#pragma unroll
    for (int sub = 0; sub < G; ++sub) { if (PAIR) umma_commit_pair(smem_u32(&empty_a[slot[sub]])); else umma_commit(smem_u32(&empty_a[slot[sub]])); }   // frees the A slots
  }
  __syncwarp();
}

// The MMA warp's main loop.  DUAL: the accumulation runs over two operand segments (TcShape::kb2): kb_per_tap k-blocks
// of the first A matrix against 1-tap weights, then kb2 k-blocks of the second A matrix against 3-tap weights.
template <int TAPS, int G, bool PAIR, bool DUAL = false, bool SIO = false>
__device__ __forceinline__ void mma_issue_loop(const TcShape& sh, const bool leader, const uint32_t tmem_base,
                                               const uint32_t a_ring_addr, const uint32_t w_addr, uint64_t* full_a,
                                               uint64_t* empty_a, uint64_t* full_w, uint64_t* empty_w,
                                               uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, const int t_first,
                                               const int t_end, const int t_step, unsigned& tr_n, const int tr_role,
                                               const int lane) {
  uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
  int it = 0;
  for (int t = t_first; t < t_end; t += t_step, ++it) {
    const int as = sh.acc_stages == 2 ? (it & 1) : 0;
    const uint32_t use = sh.acc_stages == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
    mbar_wait(smem_u32(&tmem_empty_bar[as]), (use & 1u) ^ 1u);
    tc_fence_after();
    DHG_TR(0x20, it);
    const uint32_t acc = tmem_base + (uint32_t)as * 256u;
    for (int kbi = 0; kbi < sh.kb_per_tap; ++kbi)
      mma_kblock<TAPS, G, PAIR, SIO>(sh, leader, acc, a_ring_addr, w_addr, full_a, empty_a, full_w, empty_w, sa, pa, sw, pw, kbi == 0 ? 0u : 1u, kbi,
                                     sh.kb_per_tap, it, tr_n, tr_role, lane);
    if (DUAL)
      for (int kbi = 0; kbi < sh.kb2; ++kbi)
        mma_kblock<3, G, PAIR>(sh, leader, acc, a_ring_addr, w_addr, full_a, empty_a, full_w, empty_w, sa, pa, sw, pw, 1u,
                               sh.taps * sh.kb_per_tap + kbi, sh.kb2, it, tr_n, tr_role, lane);
    if (leader) { if (PAIR) umma_commit_pair(smem_u32(&tmem_full_bar[as])); else umma_commit(smem_u32(&tmem_full_bar[as])); }   // accumulators complete
    __syncwarp();
    DHG_TR(0x22, it);
  }
}

// Template parameters fix the epilogue variant at compile time (-1 = read the flag at run time: the
// generic instance).  kLN: LayerNorm; kAUX: AUX_* kind; kFILM: 0 none, 1 vectors shared by the batch
// (smem), 2 per-sample vectors (global loads); kOUT: 1 raw, 2 SiLU'd, 3 both.
template <int kLN, int kAUX, int kFILM, int kOUT, bool kPAIR, bool kSPLIT = false, bool kSIO = false, bool kDUAL = false>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_w,
                                                                const __grid_constant__ CUtensorMap map_oraw,
                                                                const __grid_constant__ CUtensorMap map_oact,
                                                                const __grid_constant__ CUtensorMap map_a2,
                                                                const __grid_constant__ CUtensorMap map_w2,
                                                                const TcShape sh, const Epilogue e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sh.off_bar);
  uint64_t* full_a = bars;                     // [8]
  uint64_t* empty_a = bars + 8;                // [8]
  uint64_t* full_w = bars + 16;                // [8]
  uint64_t* empty_w = bars + 24;               // [8]
  uint64_t* w_all_bar = bars + 32;             // resident W landed
  uint64_t* tmem_full_bar = bars + 33;         // [2]
  uint64_t* tmem_empty_bar = bars + 35;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);
  uint64_t* stats_bar = bars + 40;             // [2] split mode: the peer's LayerNorm partial statistics have landed
  float* bias_s = reinterpret_cast<float*>(smem + sh.off_vec);
  float* gamma_s = bias_s + sh.vec_bias_n;
  float* betap_s = gamma_s + sh.film_n;
  float* dot_s = betap_s + sh.film_n;          // [3][N] dot mode

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned tr_n = 0;
  const int tr_role = warp;   // 0 producer, 1 MMA, 2.. epilogue warps (the per-chunk points only from the first one)
  // work assignment: tile `it` of this CTA -> (m tile, column group)
  // sticky: the CTAs [grp_cta0[g], grp_cta0[g+1]) work on column group g only (more CTAs for the groups whose
  // epilogue also has per-position bias rows to fetch), so that group's W tiles can stay resident
  // pair mode: the two CTAs of a cluster work on two consecutive row tiles of the same column group
  // split mode (LayerNorm rows wider than one double-buffered accumulator): the two CTAs of a cluster work on the SAME
  // row tile, rank r on column half r; they only meet in the LayerNorm statistics (see the epilogue)
  const uint32_t cta_rank = kPAIR ? cluster_ctarank() : 0u;
  const uint32_t split_rank = kSPLIT ? cluster_ctarank() : 0u;
  constexpr bool kCLUSTER = kPAIR || kSPLIT;
  const int tiles_per_super = kPAIR ? 2 : sh.G;
  int my_group = kSPLIT ? (int)split_rank : 0;
  int t_first = kCLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, t_step = kCLUSTER ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const bool bound_group = sh.sticky || kSPLIT;   // this CTA works on one column group only
  if (sh.sticky) {
    while (my_group + 1 < sh.n_groups && (int)blockIdx.x >= sh.grp_cta0[my_group + 1]) ++my_group;
    t_first = (int)blockIdx.x - sh.grp_cta0[my_group];
    t_step = sh.grp_cta0[my_group + 1] - sh.grp_cta0[my_group];
  }
  const int t_end = bound_group ? sh.m_super : sh.m_super * sh.n_groups;   // work items = (super-tile, column group)
  const bool one_group = bound_group || sh.n_groups == 1;   // work item index == super-tile index (no division needed)
  uint8_t* a_ring = smem;
  uint8_t* w_base = smem + sh.off_w;
  const bool film = e.gamma != nullptr;
  const bool film_s = film && e.film_bstride == 0 && sh.film_n > 0;   // FiLM vectors shared by the batch -> smem
  const bool fold_bias = film_s && !e.ln && sh.vec_bias_n > 0;        // (acc + b) * g + beta = acc * g + (b * g + beta)

  // Programmatic dependent launch: let the next kernel of the chain start its own prologue (barrier init, TMEM
  // allocation, weight loads) on SMs this grid has already left; everything that reads or writes activations
  // waits for the previous kernel below (griddepcontrol.wait).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_oraw)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_oact)) : "memory");
    if (kDUAL) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a2)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w2)) : "memory");
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(smem_u32(&full_a[s]), 1);
      mbar_init(smem_u32(&empty_a[s]), 1);
      mbar_init(smem_u32(&full_w[s]), 1);
      mbar_init(smem_u32(&empty_w[s]), 1);
    }
    mbar_init(smem_u32(w_all_bar), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), EPI_WARPS * (kPAIR ? 2 : 1));   // pair: rank 0 collects both CTAs' epilogues
    }
    if (kSPLIT)
      for (int a = 0; a < 2; ++a) mbar_init(smem_u32(&stats_bar[a]), EPI_WARPS * 32);   // every epilogue thread of the peer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (kPAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int n = threadIdx.x; n < sh.vec_bias_n; n += TC_THREADS) bias_s[n] = __ldg(e.bias + n);
  if (film_s) {
    for (int n = threadIdx.x; n < sh.film_n; n += TC_THREADS) {
      const float g = __ldg(e.gamma + n), b = __ldg(e.beta + n);
      gamma_s[n] = g;
      betap_s[n] = fold_bias ? fmaf(__ldg(e.bias + n), g, b) : b;
    }
  }
  for (int n = threadIdx.x; n < sh.dot_n; n += TC_THREADS) dot_s[n] = __ldg(e.dot_w + n);
  tc_fence_before();
  __syncthreads();
  if (kCLUSTER) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the (uniform) loop, one elected lane issues the copies =====
    const bool leader = elect_one();
    const int a_row_off = sh.base_taps == 3 ? -1 : 0;   // 3 taps: one 130-row tile starting one row early
    if (sh.w_resident) {
      const int n0 = my_group * sh.BN;
      const uint32_t wb = smem_u32(w_all_bar);
      if (leader) mbar_expect_tx(wb, (uint32_t)(sh.taps * sh.kb_per_tap + 3 * sh.kb2) * sh.w_tile_bytes);
      for (int tap = 0; tap < sh.taps; ++tap)
        for (int kbi = 0; kbi < sh.kb_per_tap; ++kbi) {
          const uint32_t dst = smem_u32(w_base + (size_t)(tap * sh.kb_per_tap + kbi) * sh.w_tile_bytes);
          if (leader)
            for (int j = 0; j < sh.n_umma; ++j)
              tma_load_2d(dst + (uint32_t)j * sh.umma_n * TC_BK * 2, &map_w, wb, kbi * TC_BK, e.w_row_off + tap * sh.N + n0 + j * sh.umma_n);
        }
      for (int tap = 0; tap < 3; ++tap)   // second operand's weights (dual mode)
        for (int kbi = 0; kbi < sh.kb2; ++kbi) {
          const uint32_t dst = smem_u32(w_base + (size_t)(sh.taps * sh.kb_per_tap + tap * sh.kb2 + kbi) * sh.w_tile_bytes);
          if (leader)
            for (int j = 0; j < sh.n_umma; ++j)
              tma_load_2d(dst + (uint32_t)j * sh.umma_n * TC_BK * 2, &map_w2, wb, kbi * TC_BK, tap * sh.N + n0 + j * sh.umma_n);
        }
    }
    // weights are constants; the activations are produced by the previous kernel of the chain
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ring positions are kept incrementally (no integer division in these latency-critical loops)
    uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
    const uint64_t pol_a = l2_policy_evict_first();
    for (int t = t_first; t < t_end; t += t_step) {
      const int mts_f = one_group ? t : t / sh.n_groups;
      const int ng = bound_group ? my_group : t - mts_f * sh.n_groups;
      const int mts = sh.rev ? sh.m_super - 1 - mts_f : mts_f;
      const int n0 = ng * sh.BN;
      for (int kbi = 0; kbi < sh.kb_per_tap + sh.kb2; ++kbi) {
        const bool s2 = kbi >= sh.kb_per_tap;   // dual mode: k-block of the second operand
        const int kk = (s2 ? kbi - sh.kb_per_tap : kbi) * TC_BK;
        const CUtensorMap* ma = s2 ? &map_a2 : &map_a;
        const CUtensorMap* mw = s2 ? &map_w2 : &map_w;
        const int roff = s2 ? -1 : a_row_off;
        const uint32_t a_tx = s2 ? sh.a2_tx_bytes : sh.a_tx_bytes;
        const int ntaps = s2 ? 3 : sh.taps;
        const int w_row0 = s2 ? 0 : e.w_row_off;
        for (int sub = 0; sub < sh.G; ++sub) {   // the G row tiles of this super-tile share every W tile
          mbar_wait(smem_u32(&empty_a[sa]), pa ^ 1u);
          DHG_TR(0x10, sa);
          const uint32_t fb = smem_u32(&full_a[sa]);
          if (leader) {
            const int row0 = (mts * tiles_per_super + sub + (int)cta_rank) * TC_BM + roff;
            if (kPAIR) {   // both CTAs' bytes are counted on rank 0's barrier
              if (cta_rank == 0) mbar_expect_tx(fb, 2 * a_tx);
              if (sh.a_evict_first) tma_load_2d_pair_hint(smem_u32(a_ring + (size_t)sa * sh.a_stage_bytes), ma, fb, kk, row0, pol_a);
              else tma_load_2d_pair(smem_u32(a_ring + (size_t)sa * sh.a_stage_bytes), ma, fb, kk, row0);
            } else {
              mbar_expect_tx(fb, a_tx);
              if (sh.a_evict_first) tma_load_2d_hint(smem_u32(a_ring + (size_t)sa * sh.a_stage_bytes), ma, fb, kk, row0, pol_a);
              else tma_load_2d(smem_u32(a_ring + (size_t)sa * sh.a_stage_bytes), ma, fb, kk, row0);
            }
          }
          if (++sa == (uint32_t)sh.stages_a) { sa = 0; pa ^= 1u; }
        }
        if (!sh.w_resident) {
          for (int tap = 0; tap < ntaps; ++tap) {
            mbar_wait(smem_u32(&empty_w[sw]), pw ^ 1u);
            const uint32_t fw = smem_u32(&full_w[sw]);
            const uint32_t dst = smem_u32(w_base + (size_t)sw * sh.w_tile_bytes);
            if (leader) {
              if (kPAIR) {   // each CTA of the pair loads its half of the N rows of every MMA's B operand
                const int hn = sh.umma_n >> 1;
                if (cta_rank == 0) mbar_expect_tx(fw, 2 * sh.w_tile_bytes);
                for (int j = 0; j < sh.n_umma; ++j)
                  tma_load_2d_pair(dst + (uint32_t)j * hn * TC_BK * 2, mw, fw, kk, w_row0 + tap * sh.N + n0 + j * sh.umma_n + (int)cta_rank * hn);
              } else {
                mbar_expect_tx(fw, sh.w_tile_bytes);
                for (int j = 0; j < sh.n_umma; ++j)
                  tma_load_2d(dst + (uint32_t)j * sh.umma_n * TC_BK * 2, mw, fw, kk, w_row0 + tap * sh.N + n0 + j * sh.umma_n);
              }
            }
            if (++sw == (uint32_t)sh.stages_w) { sw = 0; pw ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues tcgen05 =====
    // Accumulating MMAs into one TMEM tile do not issue back to back (measured ~130 cycles apart at N = 128
    // against a 64-cycle floor), so a super-tile keeps G independent accumulators (G row tiles x the same W
    // tile) and interleaves them.
    const bool leader = elect_one();
    if (sh.w_resident) {
      mbar_wait(smem_u32(w_all_bar), 0);
      tc_fence_after();
    }
    const uint32_t a_ring_addr = smem_u32(a_ring), w_addr = smem_u32(w_base);
#define DHG_MMA_LOOP(TAPS, GG)                                                                                          \
  mma_issue_loop<TAPS, GG, false>(sh, leader, tmem_base, a_ring_addr, w_addr, full_a, empty_a, full_w, empty_w, tmem_full_bar, \
                           tmem_empty_bar, t_first, t_end, t_step, tr_n, tr_role, lane)
#define DHG_MMA_LOOP_PAIR(TAPS)                                                                                         \
  mma_issue_loop<TAPS, 1, true>(sh, leader, tmem_base, a_ring_addr, w_addr, full_a, empty_a, full_w, empty_w, tmem_full_bar, \
                           tmem_empty_bar, t_first, t_end, t_step, tr_n, tr_role, lane)
#define DHG_MMA_LOOP_SIO(TAPS, PP)                                                                                       \
  mma_issue_loop<TAPS, 1, PP, false, true>(sh, leader, tmem_base, a_ring_addr, w_addr, full_a, empty_a, full_w, empty_w, tmem_full_bar, \
                           tmem_empty_bar, t_first, t_end, t_step, tr_n, tr_role, lane)
#define DHG_MMA_LOOP_DUAL(GG, PP)                                                                                        \
  mma_issue_loop<1, GG, PP, true>(sh, leader, tmem_base, a_ring_addr, w_addr, full_a, empty_a, full_w, empty_w, tmem_full_bar, \
                           tmem_empty_bar, t_first, t_end, t_step, tr_n, tr_role, lane)
    if (kDUAL) {   // two operand segments into the same accumulators (1-tap first operand, 3-tap second operand)
      if (kPAIR) { if (cta_rank == 0) DHG_MMA_LOOP_DUAL(1, true); }
      else if (sh.G == 4) DHG_MMA_LOOP_DUAL(4, false); else if (sh.G == 2) DHG_MMA_LOOP_DUAL(2, false); else DHG_MMA_LOOP_DUAL(1, false);
    } else if (kPAIR) {
      if (cta_rank == 0) {   // rank 0 issues for both CTAs
        if (kSIO) { if (sh.taps == 3) DHG_MMA_LOOP_SIO(3, true); else DHG_MMA_LOOP_SIO(1, true); }
        else if (sh.taps == 3) DHG_MMA_LOOP_PAIR(3); else DHG_MMA_LOOP_PAIR(1);
      }
    } else if (kSIO) {   // split I/O: six k-steps per k-block (hi.w_hi, lo.w_hi, hi.w_lo); no interleaving (G = 1)
      if (sh.taps == 3) DHG_MMA_LOOP_SIO(3, false); else DHG_MMA_LOOP_SIO(1, false);
    } else if (sh.taps == 3) {
      if (sh.G == 4) DHG_MMA_LOOP(3, 4); else if (sh.G == 2) DHG_MMA_LOOP(3, 2); else DHG_MMA_LOOP(3, 1);
    } else {
      if (sh.G == 4) DHG_MMA_LOOP(1, 4); else if (sh.G == 2) DHG_MMA_LOOP(1, 2); else DHG_MMA_LOOP(1, 1);
    }
#undef DHG_MMA_LOOP
#undef DHG_MMA_LOOP_PAIR
#undef DHG_MMA_LOOP_DUAL
#undef DHG_MMA_LOOP_SIO
  } else {
    // ===== epilogue warps; warp (q, part): TMEM lanes [32q, 32q+32), one contiguous part of the column chunks =====
    asm volatile("griddepcontrol.wait;" ::: "memory");   // residual rows are read / outputs written only after the previous kernel
    const bool ln = kLN >= 0 ? (kLN != 0) : (e.ln != 0);
    const int aux_kind = kAUX >= 0 ? kAUX : sh.aux_kind;
    const int film_mode = kFILM >= 0 ? kFILM : (film ? (film_s ? 1 : 2) : 0);
    const int out_mode = kOUT >= 0 ? kOUT : (sh.dot_n ? 4 : ((e.out_raw ? 1 : 0) | (e.out_act ? 2 : 0)));   // 4: dot mode
    const bool has_bias = kLN >= 0 ? true : sh.vec_bias_n > 0;   // every specialised variant has a bias vector (checked at plan time)
    const int ew = warp - 2, q = warp & 3, part = ew >> 2;
    const int nch = sh.BN >> 5;
    const int c_lo = (part * nch) / EPI_PARTS;          // my contiguous range of 32-column chunks (may be empty)
    const int c_hi = ((part + 1) * nch) / EPI_PARTS;
    const int my_nch = c_hi - c_lo;
    const uint32_t aux_ring = smem_u32(smem + sh.off_aux + (size_t)ew * AUX_RING_BYTES);
    const uint32_t out_st = smem_u32(smem + sh.off_out + (size_t)ew * (sh.out_bufs == 2 ? OUT_STAGE_BYTES : 2048));
    const uint32_t bias_sa = smem_u32(bias_s), gamma_sa = smem_u32(gamma_s), betap_sa = smem_u32(betap_s);
    float2* ln_s = reinterpret_cast<float2*>(smem + sh.off_ln);   // [2 parity][128 rows][EPI_PARTS (x2 in split mode)] {mean, M2} of each column part
    const int aux_ncols = aux_kind == AUX_ROWBIAS ? e.rowbias16_cols : 0x7fffffff;   // aux only for columns below this
    // split I/O: a 32-column chunk of bfs elements is 128 bytes per row: slots twice as large, half as many
    constexpr int aux_depth = kSIO ? AUX_DEPTH / 2 : AUX_DEPTH;
    constexpr uint32_t aux_slot_bytes = kSIO ? 2 * AUX_SLOT_BYTES : AUX_SLOT_BYTES;
    constexpr int esz = kSIO ? 4 : 2;   // bytes per activation element
    const char* aux_base = nullptr;
    size_t aux_pitch_bytes = 0;
    if (aux_kind == AUX_RES_PRE) { aux_base = (const char*)e.res_pre; aux_pitch_bytes = (size_t)e.res_pre_pitch * esz; }
    else if (aux_kind == AUX_RES_POST || aux_kind == AUX_RES_POST_UP) { aux_base = (const char*)e.res_post; aux_pitch_bytes = (size_t)e.res_post_pitch * esz; }
    else if (aux_kind == AUX_ROWBIAS) { aux_base = (const char*)e.rowbias16; aux_pitch_bytes = (size_t)e.rowbias16_cols * esz; }
    const bool aux_in_pass1 = ln && aux_kind == AUX_RES_PRE;
    const int r_tile = q * 32 + lane;   // my accumulator row inside the tile
    const uint32_t lane_sel = ((uint32_t)(q * 32)) << 16;
    // swizzled staging offsets: thread-per-row side (my row = lane) and coalesced side (4 lanes per row)
    const uint32_t st_row = (uint32_t)lane * 64u, st_sw = (uint32_t)((lane >> 1) & 3);
    const int period = e.map.period, pad_first = e.map.pad_first, nvalid = e.map.nvalid;
    const float inv_period = 1.0f / (float)period;

    // aux rows are prefetched as one flat sequence of 32x32 chunks across tiles (tile it, chunk ci) ->
    // flat index it * my_nch + ci, ring slot = flat % depth, so the loads for the next tile are already in
    // flight while this tile is being finished.
    uint32_t aux_issued = 0, aux_consumed = 0, st_flip = 0;
    int iss_t = t_first, iss_sub = 0, iss_ci = 0;   // (super-tile, row tile, chunk) of the next flat chunk to issue
    int iss_ng = 0;                                 // cached per issue tile: the column group,
    uint32_t iss_off[4] = {0, 0, 0, 0};             //   (gathered rows) source offsets in 16-byte units of the 4 rows this lane copies; ~0u: none
    int iss_row0 = 0;                               //   and (same-row residuals) the first row of my 32-row slab
    const bool aux_same_row = aux_kind == AUX_RES_PRE || aux_kind == AUX_RES_POST;
    const bool aux_col_limited = aux_kind == AUX_ROWBIAS;
    auto issue_aux_flat = [&]() {
      const uint32_t f = aux_issued++;
      const int ft = iss_t, fsub = iss_sub, ci = iss_ci;
      if (ft < t_end) {
        if (ci == 0) {   // first chunk of a row tile: where does my row's residual / bias row live?
          const int fmts_f = one_group ? ft : ft / sh.n_groups;
          iss_ng = bound_group ? my_group : ft - fmts_f * sh.n_groups;
          const int fmts = sh.rev ? sh.m_super - 1 - fmts_f : fmts_f;
          const int fm = (fmts * tiles_per_super + fsub + (int)cta_rank) * TC_BM + r_tile;
          const bool f_in = fm < sh.rows;
          if (aux_same_row) {
            iss_row0 = fm - lane;
          } else {
            const int fmm = f_in ? fm : 0;
            int fb, fj;
            fast_divmod(fmm, period, inv_period, fb, fj);
            const bool f_pad = (fmm >= nvalid) || (pad_first && fj == 0);
            const int fpos = f_pad ? 0 : fj - pad_first;
            const bool f_live = f_in && !f_pad;
            int my_src;
            if (aux_kind == AUX_RES_POST_UP) my_src = f_live ? fb * e.res_post_period_lo + 1 + (fpos >> 1) : -1;
            else my_src = f_live ? fpos : -1;
            if (!kSIO) {   // the rows this lane copies (i * 8 + lane / 4), once per row tile instead of once per chunk
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int src = __shfl_sync(0xffffffffu, my_src, i * 8 + (lane >> 2));
                iss_off[i] = src < 0 ? ~0u : (uint32_t)(((size_t)src * aux_pitch_bytes) >> 4);
              }
            } else {
              iss_off[0] = (uint32_t)my_src;
            }
          }
        }
        const int aux_src = kSIO ? (int)iss_off[0] : 0, fng = iss_ng;
        const int col0 = fng * sh.BN + (c_lo + ci) * 32;
        const uint32_t slot = aux_ring + (f % (uint32_t)aux_depth) * aux_slot_bytes;
        if (kSIO) {   // 32 rows x 128 bytes: lane -> (row, one of 8 16-byte pieces), 128B-row xor swizzle
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3), piece = lane & 7;
            int src;
            if (aux_same_row) { src = iss_row0 + rr; if (src >= sh.rows) src = -1; }
            else src = __shfl_sync(0xffffffffu, aux_src, rr);
            if (aux_same_row || !aux_col_limited || col0 < aux_ncols) {
              const char* gp = aux_base + (size_t)(src < 0 ? 0 : src) * aux_pitch_bytes + (size_t)col0 * 4 + piece * 16;
              cp_async16(slot + rr * 128 + ((piece ^ (rr & 7)) << 4), gp, src < 0 ? 0u : 16u);
            }
          }
        } else if (aux_same_row) {   // residual rows = my own rows: lane -> (row, 16-byte piece) directly
          const char* gp0 = aux_base + (size_t)col0 * 2 + (lane & 3) * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2);
            const int src = iss_row0 + rr;
            const bool ok = src < sh.rows;
            cp_async16(slot + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4), gp0 + (size_t)(ok ? src : 0) * aux_pitch_bytes, ok ? 16u : 0u);
          }
        } else if (!aux_col_limited || col0 < aux_ncols) {
          const char* gp0 = aux_base + (size_t)col0 * 2 + (lane & 3) * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2), piece = lane & 3;
            const bool ok = iss_off[i] != ~0u;
            cp_async16(slot + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4), gp0 + ((size_t)(ok ? iss_off[i] : 0u) << 4), ok ? 16u : 0u);
          }
        }
      }
      if (++iss_ci == my_nch) {
        iss_ci = 0;
        if (++iss_sub == sh.G) { iss_sub = 0; iss_t += t_step; }
      }
      cp_async_commit();
    };
    // wait for the oldest aux chunk, add my row of it to v, refill its slot with the next flat chunk.
    // Note: cp.async groups and bulk (TMA store) groups are counted on the same scoreboard (both waits are DEPBAR.LE SB0
    // in SASS; tools/micro/depbar_share.cu: a bulk wait_group.read behind a fresh cp.async takes 1500 instead of 200
    // cycles), so the store's `bulk_wait_read<1>` also waits for all but one of the refills in flight.  Refilling only
    // after the store's wait was measured and is NOT faster (q|k|v 90.5 -> 93.4 us, LayerNorm GEMMs +1..+4 us): the
    // launch is not bound by this wait.
    auto consume_aux = [&](float* v, int col0) {
      if (ew == 0) DHG_TR_FINE(0x38, 0);
      cp_async_wait<aux_depth - 1>();
      __syncwarp();
      if (ew == 0) DHG_TR_FINE(0x39, 0);
      const uint32_t slot = aux_ring + (aux_consumed % (uint32_t)aux_depth) * aux_slot_bytes;
      ++aux_consumed;
      if (kSIO) {
        if (!aux_col_limited || col0 < aux_ncols) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {   // 16-byte pieces 0..3 of the row: hi halves of 8 elements each; 4..7: the lo halves
            const uint4 uh = lds128_v(slot + (uint32_t)lane * 128u + (uint32_t)((p ^ (lane & 7)) << 4));
            const uint4 ul = lds128_v(slot + (uint32_t)lane * 128u + (uint32_t)(((p + 4) ^ (lane & 7)) << 4));
            float a, b;
            split_unpack2(uh.x, ul.x, a, b); v[p * 8] += a; v[p * 8 + 1] += b;
            split_unpack2(uh.y, ul.y, a, b); v[p * 8 + 2] += a; v[p * 8 + 3] += b;
            split_unpack2(uh.z, ul.z, a, b); v[p * 8 + 4] += a; v[p * 8 + 5] += b;
            split_unpack2(uh.w, ul.w, a, b); v[p * 8 + 6] += a; v[p * 8 + 7] += b;
          }
        }
      } else if (!aux_col_limited || col0 < aux_ncols) {
        uint4 u[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) u[p] = lds128_v(slot + st_row + ((p ^ st_sw) << 4));
#pragma unroll
        for (int p = 0; p < 4; ++p) add_bf16x8(u[p], v + p * 8);
      }
      __syncwarp();
      if (ew == 0) DHG_TR_FINE(0x3a, 0);
      issue_aux_flat();
      if (ew == 0) DHG_TR_FINE(0x3b, 0);
    };
    const bool aux_active = aux_kind != AUX_NONE && my_nch > 0;
    if (aux_active)
      for (int i = 0; i < aux_depth; ++i) issue_aux_flat();

    int it = 0, tile_no = 0;
    for (int t = t_first; t < t_end; t += t_step, ++it) {
     const int mts_f = one_group ? t : t / sh.n_groups;
     const int ng = bound_group ? my_group : t - mts_f * sh.n_groups;
     const int mts = sh.rev ? sh.m_super - 1 - mts_f : mts_f;
     const int n0 = ng * sh.BN;
     const int as = sh.acc_stages == 2 ? (it & 1) : 0;
     const uint32_t use = sh.acc_stages == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
     DHG_TR(0x30, it);
     mbar_wait(smem_u32(&tmem_full_bar[as]), use & 1u);
     tc_fence_after();
     DHG_TR(0x31, it);
     if (sh.direct_store == 2) {   // measurement only ("direct_store" = 2): no epilogue at all, the accumulators go straight back
       tc_fence_before();
       __syncwarp();
       if (lane == 0) { if (cta_rank) mbar_arrive_remote(smem_u32(&tmem_empty_bar[as]), 0); else mbar_arrive(smem_u32(&tmem_empty_bar[as])); }
       tile_no += sh.G;
       continue;
     }
     for (int sub = 0; sub < sh.G; ++sub, ++tile_no) {
      const int m0 = (mts * tiles_per_super + sub + (int)cta_rank) * TC_BM;
      const bool last_sub = sub == sh.G - 1;
      const uint32_t trow = tmem_base + (uint32_t)as * 256u + (uint32_t)(sub * sh.BN) + lane_sel;

      // ---- per-row bookkeeping ----
      const int m = m0 + r_tile;
      const bool in_range = m < sh.rows;
      const int mm = in_range ? m : 0;
      int b, j;
      fast_divmod(mm, period, inv_period, b, j);
      const bool is_pad = (mm >= nvalid) || (pad_first && j == 0);
      const bool live = in_range && !is_pad;

      // bf16 store of my 32 values: thread-per-row into the SWIZZLE_64B staging tile, then ONE TMA store of the
      // [32 rows x 32 columns] chunk (rows past the end of the matrix are clipped by the tensor map)
      // Buffer reuse: right after committing store k, lane 0 waits until store k-1 has been read out of its buffer
      // (the buffer store k+1 will use); the other lanes learn about it through the next warp barrier they pass:
      // the one in front of the next chunk's tcgen05.ld, or `sync_first` for a second store of the same chunk.
      auto store_chunk = [&](const CUtensorMap* omap, int col0, const float* v, bool act, bool sync_first) {
        if (kSIO) {   // a group of 32 elements = 64 B of hi halves + 64 B of lo halves per row: one [32 rows x 32 bf16]
                      // SWIZZLE_64B tile each (hi tile at column 2 * col0, lo tile 32 columns on); exact SiLU, fp32 remainders
          float r[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t buf = out_st + ((st_flip && sh.out_bufs == 2) ? 2048u : 0u);
            st_flip ^= 1u;
            if (sync_first || half) __syncwarp();   // lane 0 has waited for this buffer's previous store to be read out
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              uint32_t w[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const int i = p * 8 + k2 * 2;
                if (half == 0) {
                  const float a = act ? silu_rcp(v[i]) : v[i], c = act ? silu_rcp(v[i + 1]) : v[i + 1];
                  w[k2] = bf16_pair_hi(a, c, r[i], r[i + 1]);
                } else {
                  w[k2] = bf16_pair(r[i], r[i + 1]);
                }
              }
              sts128(buf + st_row + ((p ^ st_sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(omap, buf, col0 * 2 + half * 32, m0 + q * 32);
              bulk_commit();
              if (sh.out_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
            }
          }
          return;
        }
        if (sh.direct_store) {
          if (in_range) {
            bf16* orow = reinterpret_cast<bf16*>(act ? e.out_act : e.out_raw) + (size_t)m * (size_t)(act ? e.out_act_pitch : e.out_raw_pitch) + col0;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t w8[8];
#pragma unroll
              for (int k2 = 0; k2 < 8; ++k2) {
                float a = v[g * 16 + k2 * 2], c = v[g * 16 + k2 * 2 + 1];
                if (act) silu_fast2(a, c);
                w8[k2] = pack_bf16x2(a, c);
              }
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + g * 16), "r"(w8[0]), "r"(w8[1]), "r"(w8[2]),
                           "r"(w8[3]), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7])
                           : "memory");
            }
          }
          return;
        }
        const uint32_t buf = out_st + ((st_flip && sh.out_bufs == 2) ? 2048u : 0u);
        st_flip ^= 1u;
        if (sync_first) __syncwarp();
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          uint32_t w[4];
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            float a = v[p * 8 + k2 * 2], c = v[p * 8 + k2 * 2 + 1];
            if (act) silu_fast2(a, c);
            w[k2] = pack_bf16x2(a, c);
          }
          sts128(buf + st_row + ((p ^ st_sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(omap, buf, col0, m0 + q * 32);
          bulk_commit();
          if (sh.out_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
      };

      float v[32];
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;   // dot mode: my partial dot products of this row
      float mean = 0.f, rstd = 1.f;
      if (ln) {
        // pass 1: x = acc + bias + res_pre -> back to TMEM; shifted sums per STATISTICS GROUP of columns.  A row always
        // has the same groups whatever the tile configuration: a quarter of the row when the row has a multiple of 4
        // chunks (so the plain / paired kernel, 2 groups per column part, and the column-split kernel, 1 group per
        // part and CTA, merge the same partial sums in the same order and give the same bits), else one per part.
        const bool quarters = !kSPLIT && (nch % (2 * EPI_PARTS)) == 0;   // plain / paired kernel with a row of 4k chunks
        const int gpp = quarters ? 2 : 1;                                 // groups per part
        const int g_nch = my_nch / gpp;                                                    // chunks per group
        float shift0 = 0.f, s1a = 0.f, s2a = 0.f, shift1 = 0.f, s1b = 0.f, s2b = 0.f;
        for (int ci = 0; ci < my_nch; ++ci) {
          const int c = c_lo + ci;
          tmem_ld32(trow + c * 32, v);
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = lds128f(bias_sa + (uint32_t)(n0 + c * 32 + i) * 4u);
              add2(v[i], v[i + 1], bb.x, bb.y); add2(v[i + 2], v[i + 3], bb.z, bb.w);
            }
          }
          if (aux_in_pass1) consume_aux(v, n0 + c * 32);
          const bool second = gpp == 2 && ci >= g_nch;
          if (ci == 0) shift0 = v[0];
          if (gpp == 2 && ci == g_nch) shift1 = v[0];
          const float shift = second ? shift1 : shift0;
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float dlt = v[i] - shift;
            t1 += dlt;
            t2 = fmaf(dlt, dlt, t2);
          }
          if (second) { s1b += t1; s2b += t2; } else { s1a += t1; s2a += t2; }
          tmem_st32(trow + c * 32, v);
        }
        // merge the groups (Chan's parallel mean / M2 update) through shared memory, in column order; in split mode
        // the row's other half lives in the peer CTA: each side also writes its groups into the other's table (DSMEM)
        // and announces them on the other's stats barrier.  Tables alternate with the tile parity: the peer can only
        // be one tile ahead, because it needs my statistics to finish a tile.
        constexpr int NG = 2 * EPI_PARTS;   // table entries per row (unused ones are skipped)
        float2* st = ln_s + ((size_t)(tile_no & 1) * TC_BM + r_tile) * NG;
        const int n_groups_row = kSPLIT ? 2 * EPI_PARTS : EPI_PARTS * gpp;
        const int my_slot = kSPLIT ? (int)split_rank * EPI_PARTS + part : part * gpp;
        if (my_nch > 0) {
          const float n_g = (float)(g_nch * 32);
          const float2 ga = make_float2(shift0 + s1a / n_g, fmaxf(s2a - s1a * s1a / n_g, 0.f));
          st[my_slot] = ga;
          if (gpp == 2) st[my_slot + 1] = make_float2(shift1 + s1b / n_g, fmaxf(s2b - s1b * s1b / n_g, 0.f));
          if (kSPLIT) st_shared_remote_f2(smem_u32(&st[my_slot]), split_rank ^ 1u, ga.x, ga.y);
        }
        if (kSPLIT) mbar_arrive_remote(smem_u32(&stats_bar[tile_no & 1]), split_rank ^ 1u);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * EPI_PARTS) : "memory");
        if (kSPLIT) mbar_wait_cluster(smem_u32(&stats_bar[tile_no & 1]), (uint32_t)(tile_no >> 1) & 1u);
        float n_acc = 0.f, m2 = 0.f;
        mean = 0.f;
        for (int pp = 0; pp < n_groups_row; ++pp) {
          // group width: uniform in split / quarters mode, else the (possibly uneven) column part pp
          const int cnt = (kSPLIT || quarters) ? (nch / (EPI_PARTS * gpp)) * 32
                                               : (((pp + 1) * nch) / EPI_PARTS - (pp * nch) / EPI_PARTS) * 32;
          if (cnt > 0) {
            const float2 o = st[pp];
            const float n_p = (float)cnt, n_new = n_acc + n_p;
            const float delta = o.x - mean;
            mean += delta * (n_p / n_new);
            m2 += o.y + delta * delta * (n_acc * n_p / n_new);
            n_acc = n_new;
          }
        }
        const float n_t = (float)(sh.BN * (kSPLIT ? 2 : 1));
        rstd = rsqrtf(m2 / n_t + 1e-6f);
      }
      const float nmr = -mean * rstd;
      const float* gam_g = nullptr;
      const float* bet_g = nullptr;
      if (film_mode == 2) {   // per-sample FiLM vectors (dhg_denoise with per-sample sigma): global loads
        // halo / trailing rows are zeroed below, but their "sample" index can be one past the batch (the trailing
        // zero row of the padded layout): they read sample 0's vectors instead of running off the [B, tot] table
        const size_t bs = live ? (size_t)b : 0;
        gam_g = e.gamma + bs * e.film_bstride;
        bet_g = e.beta + bs * e.film_bstride;
      }
      if (my_nch == 0 && last_sub) {   // no column chunk for this warp in this tile shape: just release the accumulators
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (cta_rank) mbar_arrive_remote(smem_u32(&tmem_empty_bar[as]), 0); else mbar_arrive(smem_u32(&tmem_empty_bar[as])); }
      }
      for (int ci = 0; ci < my_nch; ++ci) {
        const int c = c_lo + ci;
        const int n = n0 + c * 32;
        if (ew == 0) DHG_TR_FINE(0x33, ci);
        tmem_ld32(trow + c * 32, v);
        if (ew == 0) DHG_TR_FINE(0x34, ci);
        if (ci == my_nch - 1 && last_sub) {   // my last TMEM read of this super-tile: hand the accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (cta_rank) mbar_arrive_remote(smem_u32(&tmem_empty_bar[as]), 0); else mbar_arrive(smem_u32(&tmem_empty_bar[as])); }
          DHG_TR(0x37, it);
        }
        if (ln) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) fma2(v[i], v[i + 1], rstd, rstd, nmr, nmr);
        } else {
          if (has_bias && !(film_mode == 1)) {   // film_mode 1 without LN: bias is folded into beta'
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bb = lds128f(bias_sa + (uint32_t)(n + i) * 4u);
              add2(v[i], v[i + 1], bb.x, bb.y); add2(v[i + 2], v[i + 3], bb.z, bb.w);
            }
          }
          if (aux_kind == AUX_ROWBIAS || aux_kind == AUX_RES_PRE) consume_aux(v, n);
        }
        if (film_mode == 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g = lds128f(gamma_sa + (uint32_t)(n + i) * 4u);
            const float4 bb = lds128f(betap_sa + (uint32_t)(n + i) * 4u);
            fma2(v[i], v[i + 1], g.x, g.y, bb.x, bb.y);
            fma2(v[i + 2], v[i + 3], g.z, g.w, bb.z, bb.w);
          }
        } else if (film_mode == 2) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gam_g + n + i));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bet_g + n + i));
            v[i] = fmaf(v[i], g.x, bb.x); v[i + 1] = fmaf(v[i + 1], g.y, bb.y);
            v[i + 2] = fmaf(v[i + 2], g.z, bb.z); v[i + 3] = fmaf(v[i + 3], g.w, bb.w);
          }
        }
        if (aux_kind == AUX_RES_POST || aux_kind == AUX_RES_POST_UP) consume_aux(v, n);
        if (!live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (ew == 0) DHG_TR_FINE(0x35, ci);
        if (out_mode == 4) {   // dot mode: 3 partial dot products of my columns, the row itself is not stored
          const uint32_t dsa = smem_u32(dot_s) + (uint32_t)n * 4u;
          const uint32_t nb = (uint32_t)sh.N * 4u;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float x0 = v[i], x1 = v[i + 1], x2 = v[i + 2], x3 = v[i + 3];
            if (e.dot_act) { x0 = silu_fast(x0); x1 = silu_fast(x1); x2 = silu_fast(x2); x3 = silu_fast(x3); }
            const float4 w0 = lds128f(dsa + (uint32_t)i * 4u), w1 = lds128f(dsa + nb + (uint32_t)i * 4u),
                         w2 = lds128f(dsa + 2u * nb + (uint32_t)i * 4u);
            d0 = fmaf(x0, w0.x, d0); d0 = fmaf(x1, w0.y, d0); d0 = fmaf(x2, w0.z, d0); d0 = fmaf(x3, w0.w, d0);
            d1 = fmaf(x0, w1.x, d1); d1 = fmaf(x1, w1.y, d1); d1 = fmaf(x2, w1.z, d1); d1 = fmaf(x3, w1.w, d1);
            d2 = fmaf(x0, w2.x, d2); d2 = fmaf(x1, w2.y, d2); d2 = fmaf(x2, w2.z, d2); d2 = fmaf(x3, w2.w, d2);
          }
        } else {
          if (out_mode & 1) store_chunk(&map_oraw, n, v, false, false);
          if (out_mode & 2) store_chunk(&map_oact, n, v, true, (out_mode & 1) != 0);
          if (ew == 0) DHG_TR_FINE(0x36, ci);
        }
      }
      if (out_mode == 4) {   // add up the column parts of the row through shared memory (tables alternate with the tile parity)
        float4* ds = reinterpret_cast<float4*>(smem + sh.off_ln) + (size_t)(tile_no & 1) * TC_BM * (EPI_PARTS - 1) + (size_t)r_tile * (EPI_PARTS - 1);
        if (part > 0) ds[part - 1] = make_float4(d0, d1, d2, 0.f);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * EPI_PARTS) : "memory");
        if (part == 0 && in_range) {
#pragma unroll
          for (int pp = 0; pp < EPI_PARTS - 1; ++pp) { const float4 o = ds[pp]; d0 += o.x; d1 += o.y; d2 += o.z; }
          *reinterpret_cast<float4*>(e.dot_out + (size_t)m * 4) = make_float4(d0, d1, d2, 0.f);
        }
      }
     }
     DHG_TR(0x32, it);
    }
    if (aux_active) cp_async_wait<0>();
    if (lane == 0) bulk_wait<0>();   // all my output tiles have left shared memory and are written
  }
  tc_fence_before();
  __syncthreads();
  if (kCLUSTER) cluster_sync_all();   // the peer may still be signalling my barriers / reading or writing my smem
  if (warp == 1) {
    tc_fence_after();
    if (kPAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace

// experiment switches (dhg_set_option: "w_resident", "specialize", "interleave")
int g_opt_w_resident = 1, g_opt_specialize = 1, g_opt_interleave = 1, g_opt_pdl = 1, g_opt_pair = 1;
int g_opt_direct_store = 0;   // experiment: register -> global 32-byte stores instead of smem staging + TMA store
int g_opt_split_n = 0;        // experiment: two half-width MMAs per k-step (independent accumulators) for 128 <= BN <= 256
int g_opt_max_stages_a = 0;   // > 0: cap the A ring of resident-W plans (experiment: how much prefetch depth does a launch need?)
// forced tile configuration for plans created without an explicit TcTune (tests sweep these through dhg_set_option)
static TcTune g_tune_default = {-1, -1, -1, -1};
static int g_tune_rev = 0;   // default walking direction of new plans (tests)
void tc_gemm_set_option(int which, int value) {
  if (which == 14) g_tune_rev = value ? 1 : 0;
  else if (which == 10) g_tune_default.bn = value;
  else if (which == 11) g_tune_default.g = value;
  else if (which == 12) g_tune_default.resident = value;
  else if (which == 13) g_tune_default.pair = value;
  else if (which == 2) g_opt_w_resident = value;
  else if (which == 4) g_opt_interleave = value;
  else if (which == 6) g_opt_pdl = value;
  else if (which == 7) g_opt_pair = value;
  else if (which == 3) g_opt_specialize = value;
  else if (which == 15) g_opt_max_stages_a = value;
  else if (which == 16) g_opt_direct_store = value;
  else if (which == 17) g_opt_split_n = value;
}

typedef void (*TcKernFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcShape, const Epilogue);
// fn_pair: the cta_group::2 build, fn_split: column-split LayerNorm (cluster launches only), fn_sio / fn_sio_pair: split I/O
struct TcKernEntry { int ln, aux, film, out; TcKernFn fn, fn_pair, fn_split, fn_sio, fn_sio_pair; };
#define DHG_TC_K(ln, aux, film, out) {ln, aux, film, out, tc_gemm_kernel<ln, aux, film, out, false>, tc_gemm_kernel<ln, aux, film, out, true>, nullptr, \
                                      tc_gemm_kernel<ln, aux, film, out, false, false, true>, tc_gemm_kernel<ln, aux, film, out, true, false, true>}
#define DHG_TC_KD(ln, aux, film, out) {ln, aux, film, out, tc_gemm_kernel<ln, aux, film, out, false>, tc_gemm_kernel<ln, aux, film, out, true>, nullptr, nullptr, nullptr}
#define DHG_TC_KL(aux, film, out) {1, aux, film, out, tc_gemm_kernel<1, aux, film, out, false>, tc_gemm_kernel<1, aux, film, out, true>, tc_gemm_kernel<1, aux, film, out, false, true>, \
                                   tc_gemm_kernel<1, aux, film, out, false, false, true>, tc_gemm_kernel<1, aux, film, out, true, false, true>}
// Every epilogue variant the denoiser plan uses (engine.cu), film = 1 (sampling: one FiLM vector per step);
// anything else (per-sample FiLM in dhg_denoise, test-only combinations) runs the generic instance.
static const TcKernEntry kTcKernels[] = {
    DHG_TC_K(0, AUX_NONE, 0, 1),          // conv_skip, att_dense, wq / kv of the text-style MHA
    DHG_TC_K(0, AUX_NONE, 0, 2),          // ffn.1, text_ffn.1, style_ffn.1
    DHG_TC_K(0, AUX_NONE, 1, 2),          // conv1, conv2
    DHG_TC_K(0, AUX_RES_POST, 1, 1),      // fc + skip
    DHG_TC_K(0, AUX_RES_POST_UP, 0, 3),   // skip_conv_k + upsample
    DHG_TC_K(0, AUX_ROWBIAS, 0, 1),       // q / kv / qkv projections with the PE-folded bias table
    DHG_TC_KD(0, AUX_NONE, 0, 4),         // dec1.conv_skip in dot mode (tail fusion; bf16 only)
    DHG_TC_KD(0, AUX_NONE, 1, 4),         // dec1.conv2 in dot mode
    DHG_TC_KL(AUX_NONE, 1, 1),          // text_dense
    DHG_TC_KL(AUX_NONE, 0, 1),          // style_ffn.3
    DHG_TC_KL(AUX_NONE, 1, 2),          // text_ffn.3
    DHG_TC_KL(AUX_RES_POST, 1, 1),      // mha.dense
    DHG_TC_KL(AUX_RES_PRE, 1, 3),       // mha2.dense
    DHG_TC_KL(AUX_RES_PRE, 1, 1),       // ffn.3
    DHG_TC_KL(AUX_RES_PRE, 1, 2),       // text-style mha.dense
    {-1, -1, -1, -1, tc_gemm_kernel<-1, -1, -1, -1, false>, tc_gemm_kernel<-1, -1, -1, -1, true>, tc_gemm_kernel<-1, -1, -1, -1, false, true>,
     tc_gemm_kernel<-1, -1, -1, -1, false, false, true>, tc_gemm_kernel<-1, -1, -1, -1, true, false, true>},   // generic
};
// dual-operand mode (engine.cu: conv_skip folded into the last GEMM of a ConvBlock): bias-only epilogue, raw (or raw + SiLU'd) output
static const TcKernFn kTcKernelsDual[2][2] = {
    {tc_gemm_kernel<0, AUX_NONE, 0, 1, false, false, false, true>, tc_gemm_kernel<0, AUX_NONE, 0, 1, true, false, false, true>},
    {tc_gemm_kernel<0, AUX_NONE, 0, 3, false, false, false, true>, tc_gemm_kernel<0, AUX_NONE, 0, 3, true, false, false, true>}};
static TcKernFn pick_kernel(int ln, int aux, int film, int out, int cluster_mode) {   // 0 plain, 1 CTA pair, 2 column-split LayerNorm, 3 / 4 split I/O plain / pair
  const int n = (int)(sizeof(kTcKernels) / sizeof(kTcKernels[0]));
  auto of = [&](const TcKernEntry& k) { return cluster_mode == 4 ? k.fn_sio_pair : cluster_mode == 3 ? k.fn_sio : cluster_mode == 2 ? k.fn_split : cluster_mode == 1 ? k.fn_pair : k.fn; };
  if (g_opt_specialize)
    for (int i = 0; i < n - 1; ++i)
      if (kTcKernels[i].ln == ln && kTcKernels[i].aux == aux && kTcKernels[i].film == film && kTcKernels[i].out == out && of(kTcKernels[i]))
        return of(kTcKernels[i]);
  return of(kTcKernels[n - 1]);
}

struct TcGemmPlan {
  CUtensorMap map_a, map_w, map_oraw, map_oact, map_a2, map_w2;
  TcShape sh;
  dim3 grid;
  size_t smem;
  TcKernFn fn_shared;    // FiLM (if any) with one vector for the batch
  TcKernFn fn_generic;   // per-sample FiLM or anything unusual
  int pdl;               // programmatic dependent launch, as the option stood when the plan was built
};

TcGemmPlan* tc_gemm_plan_create(const bf16* A, int lda, int rows, const bf16* W, int K, int N, int taps, const Epilogue& e,
                                char* err, int errlen, const TcTune* tune, const TcDual* dual) {
  const TcTune tn = tune ? *tune : g_tune_default;
  if (rows >= (1 << 24)) { snprintf(err, errlen, "rows = %d: the epilogue's row arithmetic needs rows < 2^24 (plan a smaller chunk)", rows); return nullptr; }
  const bool sio = e.split_io != 0;   // K, lda in bf16 units (2 per activation element: groups of 32 hi | 32 lo)
  if (rows <= 0 || K % 8 || N % 32 || (taps != 1 && taps != 3) || (sio && K % 64)) {
    snprintf(err, errlen, "unsupported shape rows=%d K=%d N=%d taps=%d%s", rows, K, N, taps, sio ? " (split I/O)" : "");
    return nullptr;
  }
  const int base_taps = taps;
  if (dual && (sio || taps != 1 || e.ln || e.rowbias16 || e.res_pre || e.res_post || e.film_planned || !e.bias || dual->K2 % 8 || !dual->A2 || !dual->W2 ||
               dual->w1_rows < N)) {
    snprintf(err, errlen, "dual-operand mode needs a 1-tap first operand and a bias-only epilogue");
    return nullptr;
  }
  int BN = 0;
  // pair == 2: column-split LayerNorm, a 2-CTA cluster per row tile, each CTA accumulates N/2 columns double-buffered
  const bool split = tn.pair == 2;
  if (split && sio) { snprintf(err, errlen, "column-split LayerNorm is not built for split I/O"); return nullptr; }
  if (split && !(e.ln && N % 128 == 0 && N / 2 <= 256)) { snprintf(err, errlen, "column-split mode does not fit: it needs a LayerNorm epilogue with N = 128, 256 or 384 (N=%d)", N); return nullptr; }
  if (e.ln) {
    if (!((N <= 256 && N >= 64) || N == 384)) { snprintf(err, errlen, "LayerNorm epilogue needs 64 <= N <= 256 or N == 384, got %d", N); return nullptr; }
    BN = split ? N / 2 : N;
  } else {
    for (int cand : {256, 192, 128, 96, 64})
      if (N % cand == 0) { BN = cand; break; }
    if (tn.bn > 0) {
      if (N % tn.bn || tn.bn % 32 || (tn.bn > 256 && tn.bn != 384)) { snprintf(err, errlen, "tile width %d does not fit N=%d", tn.bn, N); return nullptr; }
      BN = tn.bn;
    }
  }
  const bool dot = e.dot_planned || e.dot_w != nullptr;
  if (dot && sio) { snprintf(err, errlen, "dot mode is not built for split I/O"); return nullptr; }
  if (dot && (e.ln || e.out_raw || e.out_act || N > 256 || !e.dot_out)) { snprintf(err, errlen, "dot mode needs N <= 256, no LayerNorm, no stored outputs and a dot_out buffer"); return nullptr; }
  if (dot) {
    if (tn.bn > 0 && tn.bn != N) { snprintf(err, errlen, "dot mode does not fit a tile width below N"); return nullptr; }
    BN = N;   // a row's dot products are formed inside one CTA
  }
  if (BN == 0 || BN % 32) { snprintf(err, errlen, "no tile width for N=%d", N); return nullptr; }
  int aux_kind = AUX_NONE, naux = 0;
  if (e.rowbias16) { aux_kind = AUX_ROWBIAS; ++naux; }
  if (e.res_pre) { aux_kind = AUX_RES_PRE; ++naux; }
  if (e.res_post) { aux_kind = e.res_post_up ? AUX_RES_POST_UP : AUX_RES_POST; ++naux; }
  if (naux > 1) { snprintf(err, errlen, "at most one of rowbias / res_pre / res_post per GEMM"); return nullptr; }
  if (aux_kind == AUX_ROWBIAS && e.ln) { snprintf(err, errlen, "rowbias with LayerNorm is not supported"); return nullptr; }
  if (e.rowbias16 && (e.rowbias16_cols % 32 || e.rowbias16_cols > N)) { snprintf(err, errlen, "rowbias16_cols must be a multiple of 32 and <= N"); return nullptr; }
  if ((e.res_pre && e.res_pre_pitch % 8) || (e.res_post && e.res_post_pitch % 8) || (e.out_raw && e.out_raw_pitch % 8) ||
      (e.out_act && e.out_act_pitch % 8)) {
    snprintf(err, errlen, "row pitches must be multiples of 8 elements");
    return nullptr;
  }
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);

  TcGemmPlan* p = new TcGemmPlan();
  p->pdl = g_opt_pdl;
  TcShape& sh = p->sh;
  sh.rows = rows; sh.K = K; sh.N = N; sh.taps = taps; sh.BN = BN;
  sh.base_taps = base_taps; sh.sio = sio ? 1 : 0;
  sh.kb2 = dual ? (dual->K2 + TC_BK - 1) / TC_BK : 0;
  sh.a2_tx_bytes = (uint32_t)(TC_BM + 2) * TC_BK * 2;
  sh.n_groups = N / BN;
  const int m_tiles = (rows + TC_BM - 1) / TC_BM;
  sh.m_tiles = m_tiles;
  sh.kb_per_tap = (K + TC_BK - 1) / TC_BK;
  // BN > 256: two MMAs per k-step by necessity.  "split_n": also for 128 <= BN <= 256 with one row tile per accumulator
  // stage (G = 1), so that consecutive MMAs accumulate into different TMEM tiles (dependent accumulation is the slow case)
  sh.n_umma = (BN > 256 || (g_opt_split_n && BN > 128 && (BN / 2) % 16 == 0 && !split && !sio)) ? 2 : 1;   // BN > 128: always G = 1
  sh.umma_n = BN / sh.n_umma;
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format BF16 (1) [7,10) | b_format BF16 (1) [10,13) |
  // a,b K-major (0) | N>>3 [17,23) | M>>4 [24,29)
  sh.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(sh.umma_n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  sh.acc_stages = BN <= 256 ? 2 : 1;
  sh.aux_kind = aux_kind;
  sh.trace = nullptr; sh.trace_cap = 0;
  sh.rev = g_tune_rev;
  // direct stores need 32-byte aligned rows: pitches in multiples of 16 elements, 32-byte aligned bases (not in split-I/O mode)
  sh.direct_store = g_opt_direct_store == 2 ? 2 : (g_opt_direct_store && !sio && !dot &&
                     (!e.out_raw || (e.out_raw_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(e.out_raw) & 31) == 0)) &&
                     (!e.out_act || (e.out_act_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(e.out_act) & 31) == 0))) ? 1 : 0;
  sh.a_evict_first = 0;
  sh.vec_bias_n = e.bias ? N : 0;
  sh.film_n = e.film_planned ? N : 0;
  // independent accumulators per super-tile: as many as fit in one 256-column TMEM group, at most 4
  int G = 1;
  if (g_opt_interleave && BN <= 128 && !split && !sio) G = BN <= 64 ? 4 : 2;
  if (sio && tn.g > 1) { snprintf(err, errlen, "split I/O does not fit interleaved accumulators"); delete p; return nullptr; }
  if (split && tn.g > 1) { snprintf(err, errlen, "column-split mode does not fit interleaved accumulators"); delete p; return nullptr; }
  if (tn.g > 0) {
    if ((tn.g != 1 && tn.g != 2 && tn.g != 4) || (tn.g > 1 && tn.g * BN > 256)) { snprintf(err, errlen, "%d interleaved accumulators of %d columns do not fit a 256-column TMEM group", tn.g, BN); delete p; return nullptr; }
    G = tn.g;
  }
  while (G > 1 && (m_tiles + G - 1) / G * sh.n_groups < num_sms) G >>= 1;   // keep every SM busy on small problems
  sh.G = G;
  sh.m_super = (m_tiles + G - 1) / G;
  const int a_rows = base_taps == 3 ? TC_BM + 2 : TC_BM;
  sh.a_tx_bytes = (uint32_t)a_rows * TC_BK * 2;
  sh.a_stage_bytes = ((dual ? sh.a2_tx_bytes : sh.a_tx_bytes) + 1023u) & ~1023u;
  sh.w_tile_bytes = (uint32_t)BN * TC_BK * 2;
  // smem carve-up: A ring | W ring or resident W | aux rings | out staging | vectors | LN exchange | barriers
  sh.dot_n = dot ? 3 * N : 0;
  // LayerNorm: [2 tile parities][128 rows][2 * EPI_PARTS groups] {mean, M2}; dot mode: [2][128][EPI_PARTS - 1] float4 partial dots
  const size_t ln_bytes = e.ln ? (size_t)2 * TC_BM * (2 * EPI_PARTS) * 8 : dot ? (size_t)2 * TC_BM * (EPI_PARTS - 1) * 16 : 0;
  size_t fixed = (aux_kind != AUX_NONE ? (size_t)EPI_WARPS * AUX_RING_BYTES : 0) + (size_t)EPI_WARPS * OUT_STAGE_BYTES +
                 (size_t)(sh.vec_bias_n + 2 * sh.film_n + sh.dot_n) * 4 + 16 + ln_bytes + 64 * 8;
  size_t budget = 227 * 1024 - 1024 - fixed;
  sh.out_bufs = OUT_STAGE_BYTES >= 4096 ? 2 : 1;
  if (sh.out_bufs == 2 && 2 * (size_t)(sh.a_stage_bytes + BN * TC_BK * 2) > budget) {   // not even two A + two W stages: give up the second store tile
    sh.out_bufs = 1;
    fixed -= (size_t)EPI_WARPS * OUT_STAGE_BYTES / 2;
    budget += (size_t)EPI_WARPS * OUT_STAGE_BYTES / 2;
  }
  const size_t w_all = (size_t)(taps * sh.kb_per_tap + 3 * sh.kb2) * sh.w_tile_bytes;
  const int ring_taps = dual ? 3 : taps;   // W tiles per k-block the streamed ring must hold
  const int min_a = G > 1 ? 2 * G : 3;
  const int want_resident = tn.resident >= 0 ? tn.resident : g_opt_w_resident;
  sh.w_resident = (want_resident && w_all + (size_t)min_a * sh.a_stage_bytes <= budget && sh.m_super * sh.n_groups > num_sms) ? 1 : 0;
  sh.sticky = (sh.w_resident && sh.n_groups > 1 && !split) ? 1 : 0;
  sh.split = split ? 1 : 0;
  // W does not fit: pair the CTAs of a cluster (cta_group::2) so that each SM only ingests half of every W tile
  // Measured (profiles/): pairing pays when the MMA / W-stream phase dominates the tile (K*taps >= 512 and a light
  // epilogue, or the single-buffered 384-wide LayerNorm rows); with g_opt_pair == 2 every non-resident GEMM is paired.
  const int out_mode_plan = dot ? 4 : ((e.out_raw ? 1 : 0) | (e.out_act ? 2 : 0));
  const bool pair_pays = BN == 384 || (taps * K + (dual ? 3 * dual->K2 : 0) >= 512 && out_mode_plan != 3 && !(e.ln && N <= 192));
  const bool want_pair = split ? false : tn.pair >= 0 ? tn.pair != 0 : (g_opt_pair && (g_opt_pair == 2 || pair_pays));
  sh.pair = (!sh.w_resident && want_pair && G == 1 && sh.umma_n % 16 == 0 &&
             (m_tiles + 1) / 2 * sh.n_groups >= num_sms / 2) ? 1 : 0;
  if (sh.pair) {
    sh.m_super = (m_tiles + 1) / 2;
    sh.w_tile_bytes = (uint32_t)(BN / 2) * TC_BK * 2;
    // idesc: M = 256 across the pair
    sh.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(sh.umma_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  }
  size_t w_bytes;
  if (sh.w_resident) {
    int sa = (int)((budget - w_all) / sh.a_stage_bytes);
    sh.stages_a = sa > 8 ? 8 : sa;
    if (g_opt_max_stages_a > 0 && sh.stages_a > g_opt_max_stages_a && g_opt_max_stages_a >= G) sh.stages_a = g_opt_max_stages_a;
    sh.stages_w = 1;
    w_bytes = w_all;
  } else {
    // per k-block: G A tiles and `taps` W tiles
    int sa = (int)(budget / ((size_t)G * sh.a_stage_bytes + (size_t)ring_taps * sh.w_tile_bytes)) * G;
    if (sa < 2 * G) sa = G > 1 ? 2 * G : 2;
    if (sa > 8) sa = 8;
    if ((size_t)sa * sh.a_stage_bytes + 2 * (size_t)sh.w_tile_bytes > budget) { snprintf(err, errlen, "not enough shared memory for BN=%d", BN); delete p; return nullptr; }
    int sw = (int)((budget - (size_t)sa * sh.a_stage_bytes) / sh.w_tile_bytes);
    sh.stages_a = sa;
    sh.stages_w = sw > 8 ? 8 : sw;
    w_bytes = (size_t)sh.stages_w * sh.w_tile_bytes;
  }
  if (sh.stages_a < G) { snprintf(err, errlen, "A ring too small for %d interleaved tiles", G); delete p; return nullptr; }
  uint32_t off = (uint32_t)sh.stages_a * sh.a_stage_bytes;
  sh.off_w = off; off += (uint32_t)w_bytes;
  sh.off_aux = off; off += aux_kind != AUX_NONE ? EPI_WARPS * AUX_RING_BYTES : 0;
  sh.off_out = off; off += EPI_WARPS * (sh.out_bufs == 2 ? OUT_STAGE_BYTES : 2048);
  sh.off_vec = off; off += (uint32_t)(sh.vec_bias_n + 2 * sh.film_n + sh.dot_n) * 4;
  off = (off + 15u) & ~15u;
  sh.off_ln = off; off += (uint32_t)ln_bytes;
  sh.off_bar = off; off += 64 * 8;
  p->smem = off + 1024;
  int grid = sh.m_super * sh.n_groups < num_sms ? sh.m_super * sh.n_groups : num_sms;
  if (sh.pair) {
    const int pairs = sh.m_super * sh.n_groups < num_sms / 2 ? sh.m_super * sh.n_groups : num_sms / 2;
    grid = 2 * pairs;
  }
  if (sh.split) grid = 2 * (sh.m_super < num_sms / 2 ? sh.m_super : num_sms / 2);   // one cluster per row tile in flight
  if (sh.sticky) {
    if (sh.n_groups > 12) { snprintf(err, errlen, "too many column groups"); delete p; return nullptr; }
    // CTAs per group in proportion to the estimated per-tile epilogue cost (groups with per-position bias rows: 1.35)
    double wsum = 0, w[12];
    for (int g = 0; g < sh.n_groups; ++g) { w[g] = (aux_kind == AUX_ROWBIAS && g * BN < e.rowbias16_cols) ? 1.35 : 1.0; wsum += w[g]; }
    int used = 0;
    sh.grp_cta0[0] = 0;
    for (int g = 0; g < sh.n_groups; ++g) {
      int n = g == sh.n_groups - 1 ? num_sms - used : (int)(num_sms * w[g] / wsum + 0.5);
      if (n < 1) n = 1;
      used += n;
      sh.grp_cta0[g + 1] = used;
    }
    grid = used;
  }
  p->grid = dim3(grid);
  if (!make_map(&p->map_a, A, (uint64_t)rows, (uint64_t)K, (uint64_t)lda, (uint32_t)a_rows, err, errlen) ||
      !make_map(&p->map_w, W, dual ? (uint64_t)dual->w1_rows : (uint64_t)taps * N, (uint64_t)K, (uint64_t)K, (uint32_t)(sh.pair ? sh.umma_n / 2 : sh.umma_n), err, errlen)) {
    delete p;
    return nullptr;
  }
  p->map_a2 = p->map_a; p->map_w2 = p->map_w;   // placeholders (never used) unless this is a dual-operand plan
  if (dual && (!make_map(&p->map_a2, dual->A2, (uint64_t)rows, (uint64_t)dual->K2, (uint64_t)dual->lda2, (uint32_t)(TC_BM + 2), err, errlen) ||
               !make_map(&p->map_w2, dual->W2, (uint64_t)3 * N, (uint64_t)dual->K2, (uint64_t)dual->K2, (uint32_t)(sh.pair ? sh.umma_n / 2 : sh.umma_n), err, errlen))) {
    delete p;
    return nullptr;
  }
  // output tensor maps: [32 rows x 32 columns] SWIZZLE_64B boxes for the epilogue's TMA stores
  p->map_oraw = p->map_a; p->map_oact = p->map_a;   // placeholders for absent outputs (never used)
  // (split I/O: a row of N bfs elements is 2N bf16; a 32-column chunk leaves as two such boxes)
  const uint64_t om = sio ? 2 : 1;
  const uint32_t obox = 32;
  if ((e.out_raw && !make_map(&p->map_oraw, e.out_raw, (uint64_t)rows, om * N, om * e.out_raw_pitch, 32, err, errlen, obox)) ||
      (e.out_act && !make_map(&p->map_oact, e.out_act, (uint64_t)rows, om * N, om * e.out_act_pitch, 32, err, errlen, obox))) {
    delete p;
    return nullptr;
  }
  const int out_mode = dot ? 4 : ((e.out_raw ? 1 : 0) | (e.out_act ? 2 : 0));
  const int cluster_mode = sio ? (sh.pair ? 4 : 3) : sh.split ? 2 : sh.pair ? 1 : 0;
  p->fn_shared = e.bias ? pick_kernel(e.ln ? 1 : 0, aux_kind, e.film_planned ? 1 : 0, out_mode, cluster_mode)
                        : pick_kernel(-1, -1, -1, -1, cluster_mode);   // the specialised variants assume a bias vector
  p->fn_generic = pick_kernel(-1, -1, -1, -1, cluster_mode);
  if (dual) {
    if (out_mode != 1 && out_mode != 3) { snprintf(err, errlen, "dual-operand mode stores a raw (or raw + SiLU'd) output"); delete p; return nullptr; }
    if (sh.split) { snprintf(err, errlen, "dual-operand mode does not fit the column-split cluster"); delete p; return nullptr; }
    p->fn_shared = p->fn_generic = kTcKernelsDual[out_mode == 3 ? 1 : 0][sh.pair ? 1 : 0];
  }
  for (TcKernFn fn : {p->fn_shared, p->fn_generic}) {
    cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); delete p; return nullptr; }
  }
  return p;
}

void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }
void tc_gemm_set_trace(TcGemmPlan* p, unsigned long long* buf, int cap) { p->sh.trace = buf; p->sh.trace_cap = cap; }
void tc_gemm_plan_set_reverse(TcGemmPlan* p, int rev) { p->sh.rev = rev ? 1 : 0; }
void tc_gemm_plan_set_a_evict_first(TcGemmPlan* p, int on) { p->sh.a_evict_first = (on && p->sh.n_groups == 1 && !p->sh.split) ? 1 : 0; }
void tc_gemm_plan_config(const TcGemmPlan* p, TcTune* out) {
  out->bn = p->sh.BN; out->g = p->sh.G; out->resident = p->sh.w_resident; out->pair = p->sh.split ? 2 : p->sh.pair;
}
void tc_gemm_describe(const TcGemmPlan* p, char* out, int n) {
  snprintf(out, n, "pair=%d BN=%d groups=%d m_tiles=%d G=%d stages_a=%d stages_w=%d resident=%d sticky=%d acc_stages=%d grid=%d smem=%zu", p->sh.pair, p->sh.BN, p->sh.n_groups,
           p->sh.m_tiles, p->sh.G, p->sh.stages_a, p->sh.stages_w, p->sh.w_resident, p->sh.sticky, p->sh.acc_stages, (int)p->grid.x, p->smem);
}

int tc_gemm_launch(const TcGemmPlan* p, const Epilogue& e, cudaStream_t st) {
  // the specialised instance assumes FiLM vectors shared by the batch (or no FiLM at all)
  const bool shared_ok = e.gamma ? (e.film_bstride == 0 && p->sh.film_n > 0) : (p->sh.film_n == 0);
  TcKernFn fn = shared_ok ? p->fn_shared : p->fn_generic;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = p->grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p->pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p->sh.pair || p->sh.split) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, fn, p->map_a, p->map_w, p->map_oraw, p->map_oact, p->map_a2, p->map_w2, p->sh, e) == cudaSuccess ? 0 : 1;
}

}  // namespace dhg
