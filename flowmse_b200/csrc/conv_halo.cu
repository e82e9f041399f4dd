// 3x3 implicit-GEMM convolution, "halo" variant: persistent CTAs, one TMA halo load per 64-channel chunk.
//
// Same math and operand formats as conv_gemm.cu (fp16 hi/lo split, three tcgen05 MMAs per K step, fp32 accumulate;
// replaces the ResBlock / pyramid-head nn.Conv2d calls, /root/reference/flowmse/backbones/ncsnpp_utils/layerspp.py:259-270,
// ncsnpp.py:347-366).  What changes is the data movement, which is what bounds the per-tap kernel: it re-fetches the
// 128-pixel A tile from L2 for each of the 9 filter taps (1152 KB of L2->smem traffic per 128x128x1152 tile, 2x the
// MMA time at the measured L2->SM rate).  Here:
//  * The output tile is TH x TW = 16 x 8 pixels.  For each 64-channel chunk ONE 5-D TMA box {64 ch, 10, 18} brings the
//    (TH+2) x (TW+2) halo into shared memory (hi and lo planes), rows of 128 B with the 128-byte swizzle.
//  * The 9 taps are 9 row-shifted VIEWS of that halo: the UMMA shared-memory descriptor starts at halo row
//    (dy+1)*10 + (dx+1) with an 8-row-group stride (SBO) of 10 rows = 1280 B.  The hardware applies the 128-byte
//    swizzle as a function of the absolute shared-memory address (measured with tools/umma_probe.cu: any 128-byte row
//    start and any SBO read back exactly, base_offset = 0), so a view that is not 1024-byte aligned is still read
//    consistently with what TMA wrote.  A traffic drops from 9 x 128 to 180 rows per chunk (6.4x).
//  * Weights (B operand) stream per (tap, chunk) K block as before, through their own ring of stages.
//  * Persistent grid (one CTA per SM), static round-robin tile schedule, TMEM accumulators double-buffered
//    (2 buffers x {main, correction} x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//  * Warp roles: warp 0 TMA producer (A halos + B blocks), warp 1 MMA issuer, warp 2 TMEM allocator,
//    warps 4..11 epilogue (two per TMEM lane quadrant; the XF variant: warps 4..7, one per quadrant, then 8 transform
//    warps - 512 threads with 128 registers each).
//  * PAIR variant (Cout tiles of 128): two CTAs of a cluster (one TPC) work on two M tiles with the same weights as ONE
//    tcgen05.mma.cta_group::2 of shape 256 x 128 x 16.  Each CTA stages its own A halo and only HALF of the B block
//    (64 of the 128 weight rows); the tensor cores read the other half from the peer's shared memory.  With the halo
//    the weight stream is what is left of the L2->SM traffic, so this halves the remaining bytes per MMA.
//    Protocol: both CTAs' TMA loads complete on the LEADER's full barriers (cta_group::2 form, leader arms the
//    transaction count for both), the leader's MMA warp issues for the pair and multicasts its commits to both CTAs'
//    empty / accumulator-full barriers, both epilogues arrive on the leader's accumulator-empty barrier.
//  * XF variant (fused operand preparation): 8 extra "transform" warps build the A operand tile in shared memory
//    themselves.  Per 64-channel chunk they read the fp32 NHWC activations of the halo (virtual concat of two sources),
//    apply GroupNorm (per-channel scale / shift table assembled from the producers' quad statistics) + SiLU, split to
//    fp16 hi/lo and store the rows at the SAME swizzled addresses TMA would have written (16-byte column j of row r goes
//    to r*128 + ((j ^ (r & 7)) << 4)); fence.proxy.async + mbarrier arrive hands the stage to the MMA warp.  The
//    standalone prep kernel (read 4 B + write 4 B per element, 20 % of an evaluation) is gone for these layers and the
//    operand values are bit-identical to it.  The shortcut operand (raw x, 1 tap) only needs the 128 centre rows.
//    Either operand may still come through TMA (after a resampling prep): the producer and the transform warps both
//    arrive on every stage (count 9), whoever owns the chunk does the work.
#include "flowse_internal.h"
#include "operand.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cstdlib>

namespace flowse {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int TH = 16, TW = 8;
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int A_ROWS = HALO_W * HALO_H;                  // 180 rows of 128 B
constexpr int A_PLANE_BYTES = A_ROWS * 128;              // 23040 (what TMA delivers per plane)
constexpr int A_PLANE_STRIDE = 23552;                    // next multiple of 1024 (swizzle alignment of the lo plane)
constexpr int A_STAGE_BYTES = 2 * A_PLANE_STRIDE;
constexpr int A_STAGES = 2;
constexpr int A_SBO = HALO_W * 128;                      // 8-row group g = tile row g: one halo row further
constexpr int NUM_THREADS = 384;
constexpr int kEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;
constexpr int kXfWarps = 8;                              // transform warps of the XF variant
// Fused-operand variant: ONE epilogue warp per TMEM lane quadrant instead of two.  16 warps = 512 threads give every
// thread 128 registers (640 threads: 102 -> 96), which the transform warps use for their loads in flight; a tile's
// epilogue still fits behind the MMAs of the next tile.  Same-box A/B, ms per sampler call: 21.45-21.53 -> 21.19-21.42 at
// B = 1, 143.2-143.5 -> 140.9-141.8 at B = 8.  Two alternating register sets for the transform (the loads of the batch
// after the next one in flight) on top of it measured the same (21.26-21.36 / 141.7-142.0) and were not kept.
#ifndef XF_EPI_WARPS
#define XF_EPI_WARPS 4
#endif
constexpr int kEpiWarpsXf = XF_EPI_WARPS;
constexpr int NUM_THREADS_XF = (kFirstEpiWarp + kEpiWarpsXf + kXfWarps) * 32;
constexpr int kXfMaxC = 512;                             // channels of a fused operand (scale / shift table in smem)
// How the transform warps read the fp32 activations.  Every byte is used once per CTA and the L1 data array is the same
// SRAM the tensor core streams its operands from, so loads that do not allocate there looked attractive; measured (same
// box, two runs each, ms per sampler call): ld.global.nc 20.89 / 20.93, .cg (L2 only) 21.28 / 21.35, .nc.L1::no_allocate
// 21.13 / 21.12, .cs 20.97 / 20.81 - the 128-byte L1 lines serve the neighbouring lanes' 32-byte pieces.  Default kept.
#ifndef XF_LDMODE
#define XF_LDMODE 0
#endif
__device__ __forceinline__ float4 xf_load(const float4* g) {
#if XF_LDMODE == 0
  return __ldg(g);                                  // ld.global.nc: allocates in L1
#elif XF_LDMODE == 1
  return __ldcg(g);                                 // ld.global.cg: cached in L2 only
#elif XF_LDMODE == 2
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g));
  return v;
#else
  return __ldcs(g);                                 // ld.global.cs: streaming (evict first)
#endif
}

// Experiment switches: which threads keep the clock reads around their barrier waits (the FLOWSE_CONV_DBG=1 counters).
// Same-box runs, ms per sampler call: both kept 21.00 / 21.09 / 21.02, producer without 21.00 / 21.04 / 20.99, issuer
// without 20.96 / 20.91 / 21.00, and every counter compiled out 21.36 / 21.31 / 21.32 against 21.02 / 20.99 / 21.05 - the
// kernel's timing moves by +-1 % with changes of this kind, so the counters stay in.
#ifndef HALO_DBG_PROD
#define HALO_DBG_PROD 1
#endif
#ifndef HALO_DBG_MMA
#define HALO_DBG_MMA 1
#endif
#define HALO_TW(on, acc, ...) do { if (on) { const long long c0_ = clock64(); __VA_ARGS__; acc += clock64() - c0_; } else { __VA_ARGS__; } } while (0)
// rows per straight-line batch of a transform thread.  Same-box, 512 threads / 128 registers, ms per sampler call at B = 1:
// 6 rows 20.92-20.98, 3 rows 21.42-21.47, 2 rows 21.62-21.86.
#ifndef XF_NB
#define XF_NB 6
#endif

template <int BN, int NMAIN, bool PAIR, int EPW = kEpiWarps>
struct HCfg {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;               // weight rows staged by this CTA
  static constexpr int B_PLANE = B_ROWS * 128;
  static constexpr int B_STAGE_BYTES = 2 * B_PLANE;
  static constexpr int B_STAGES = (BN >= 128) ? (PAIR ? 6 : 3) : 8;
  // FUSED (one main accumulator, single CTA): the hi and lo weight planes of a stage are contiguous in shared memory, so
  // A_hi x [B_hi ; B_lo]^T is ONE MMA of N = 2*BN whose accumulator is [main | correction]; A_lo x B_hi^T follows with
  // N = BN into the correction columns.  A_hi is fetched from shared memory once instead of twice: 20 KB instead of
  // 24 KB of operand reads per K step.  The measured limiter of this kernel is exactly that: the issuer thread never
  // waits for data, but a 128x128x16 MMA retires every ~80 cycles instead of 64 because its 8 KB of operands plus the
  // TMA writes exceed the 128 B/cycle of the shared-memory port (tools/dbg_halo.py, profiles/README.md).
  static constexpr bool FUSED = (NMAIN == 1) && !PAIR;
  static constexpr int SLOT_COLS = FUSED ? BN : ((BN < 32) ? 32 : BN);
  static constexpr int NSLOT = NMAIN + 1;                          // hi*hi chains + one correction accumulator
  static constexpr int NBUF = (NSLOT * SLOT_COLS * 2 <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS_RAW = NBUF * NSLOT * SLOT_COLS;
  static constexpr int TMEM_COLS = PAIR ? 512 : TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64
                                   : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int CH = 16;
  static constexpr int STG_STRIDE = CH + 4;
  static constexpr int STG_BYTES = EPW * 32 * STG_STRIDE * 4;
  static constexpr int COLS_PER_WARP = (BN >= 32 && EPW == 8) ? BN / 2 : BN;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + STG_BYTES + 1024;
};

struct HaloParams {
  int H, W, tiles_w, tiles_h, n_tiles, num_items;   // items = tiles (single CTA) or pairs of M tiles (PAIR)
  int nchunk_main, nchunk_sc;
  int Cout, ldc;
  float wscale_inv;
  const float* bias;
  int bias_bstride;
  const float* residual;
  float* out;
  int div_sqrt2;
  double* qstats;
  long long* dbg;   // optional per-CTA wait-cycle counters (FLOWSE_CONV_DBG=1): 16 per CTA
};

// fused operand sources (XF variant), device view of FusedOperand
struct XfOperand { const float* s1; const float* s2; int C1, C2; };
struct XfParams {
  XfOperand a, x;                        // main / shortcut operand; s1 == nullptr: that operand comes through TMA
  const double* qs1; const double* qs2;  // quad statistics of a.s1 / a.s2
  const float* gamma; const float* beta; // GroupNorm affine of the main operand (over C1 + C2 channels)
  int silu;
  unsigned long long* overflow;
};

struct TileCoord { int b, h0, w0, n0; };

// work item -> output tile of this CTA.  PAIR: item = (pair of adjacent M tiles, N tile); CTA rank r takes M tile 2*pm + r.
template <int BN, bool PAIR>
__device__ __forceinline__ TileCoord decode_tile(const HaloParams& p, int item, int rank) {
  TileCoord t;
  const int nt = item % p.n_tiles;
  int m = item / p.n_tiles;
  if (PAIR) m = 2 * m + rank;
  const int per_img = p.tiles_w * p.tiles_h;
  t.b = m / per_img;
  m -= t.b * per_img;
  t.h0 = (m / p.tiles_w) * TH;
  t.w0 = (m % p.tiles_w) * TW;
  t.n0 = nt * BN;
  return t;
}

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// TMA loads of a CTA pair: data lands in the executing CTA, the transaction bytes complete on `bar`, which may live in
// the peer (leader) CTA
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once the issuing thread's MMAs have retired) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN, int NMAIN, bool PAIR, bool XF>
__global__ void __launch_bounds__(XF ? NUM_THREADS_XF : NUM_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmW, const HaloParams p, const XfParams xf) {
  static_assert(!XF || (!PAIR && NMAIN == 1 && BN == 128), "the fused-operand variant exists for the default tile only");
  constexpr int EPW = XF ? kEpiWarpsXf : kEpiWarps;           // epilogue warps
  using C = HCfg<BN, NMAIN, PAIR, EPW>;
  extern __shared__ uint8_t smem_raw[];
  constexpr int NBARS = 2 * A_STAGES + 2 * C::B_STAGES + 2 * C::NBUF;
  __shared__ uint64_t bars[NBARS];
  __shared__ uint32_t tmem_slot_var;
  __shared__ float s_qs[2][EPW][C::COLS_PER_WARP / 4 > 0 ? C::COLS_PER_WARP / 4 : 1][2];
  __shared__ __align__(16) float s_xsc[XF ? kXfMaxC : 4], s_xsh[XF ? kXfMaxC : 4];   // GroupNorm scale / shift per channel

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (A_STAGES + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * A_STAGES + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * A_STAGES + C::B_STAGES + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (2 * A_STAGES + 2 * C::B_STAGES + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (2 * A_STAGES + 2 * C::B_STAGES + C::NBUF + s); };
  auto sA = [&](int s) { return smem_base + static_cast<uint32_t>(s) * A_STAGE_BYTES; };
  auto sB = [&](int s) { return smem_base + A_STAGES * A_STAGE_BYTES + static_cast<uint32_t>(s) * C::B_STAGE_BYTES; };
  const uint32_t tmem_slot = ptx::smem_u32(&tmem_slot_var);
  volatile uint32_t* tmem_slot_ptr = &tmem_slot_var;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 GroupNorm table (XF), 4.. epilogue (8 warps; XF: 4),
  // then 8 operand-transform warps (XF).  (Putting
  // the MMA-issuing warp last - the issue arbiter favours high warp indices - was measured and makes no difference.)
  constexpr int w_tma = 0, w_mma = 1, w_alloc = 2, w_epi0 = kFirstEpiWarp, w_xf0 = kFirstEpiWarp + EPW;
  const bool is_xf = XF && warp >= w_xf0;
  const bool is_epi = warp >= w_epi0 && warp < w_epi0 + EPW;
  const int nchunks = p.nchunk_main + p.nchunk_sc;
  const int rank = PAIR ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int item0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int kCtas = PAIR ? 2 : 1;

  if (warp == w_tma && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
    // XF: the producer thread and every transform warp arrive on each A stage
    for (int s = 0; s < A_STAGES; ++s) { ptx::mbar_init(a_full(s), XF ? 1 + kXfWarps : 1); ptx::mbar_init(a_empty(s), 1); }
    for (int s = 0; s < C::B_STAGES; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1); }
    for (int s = 0; s < C::NBUF; ++s) { ptx::mbar_init(t_full(s), 1); ptx::mbar_init(t_empty(s), EPW * kCtas); }
    ptx::fence_mbar_init();
  }
  if (warp == w_alloc) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else { ptx::tmem_alloc(tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  pdl_wait();                 // barriers, TMEM and tensor maps were set up while the previous kernel drained

  if (warp == w_tma) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      long long w_pa = 0, w_pb = 0;            // cycles the producer waited for a free A / B stage
      // PAIR: the full barriers that count are the leader's; this CTA's loads complete there
      auto full_addr = [&](uint32_t local) { return PAIR ? ptx::map_to_cta(local, 0) : local; };
      auto issue_A = [&](int item, int c) {
        const TileCoord t = decode_tile<BN, PAIR>(p, item, rank);
        HALO_TW(HALO_DBG_PROD, w_pa, ptx::mbar_wait(a_empty(as), aph ^ 1u));
        const bool main = c < p.nchunk_main;
        if (XF && (main ? xf.a.s1 : xf.x.s1) != nullptr) {
          ptx::mbar_arrive(a_full(as));              // the transform warps fill this stage
          if (++as == A_STAGES) { as = 0; aph ^= 1u; }
          return;
        }
        if (rank == 0) ptx::mbar_expect_tx(a_full(as), 2 * A_PLANE_BYTES * kCtas);
        const CUtensorMap* m = main ? &tmA : &tmX;
        const int ch = main ? c : c - p.nchunk_main;
        const uint32_t bar = full_addr(a_full(as));
        if constexpr (PAIR) {
          tma_load_5d_pair(m, bar, sA(as), ch * BK, t.w0 - 1, t.h0 - 1, t.b, 0);
          tma_load_5d_pair(m, bar, sA(as) + A_PLANE_STRIDE, ch * BK, t.w0 - 1, t.h0 - 1, t.b, 1);
        } else {
          ptx::tma_load_5d(m, bar, sA(as), ch * BK, t.w0 - 1, t.h0 - 1, t.b, 0);
          ptx::tma_load_5d(m, bar, sA(as) + A_PLANE_STRIDE, ch * BK, t.w0 - 1, t.h0 - 1, t.b, 1);
        }
        if (++as == A_STAGES) { as = 0; aph ^= 1u; }
      };
      int item = item0;
      if (item < p.num_items) issue_A(item, 0);
      for (; item < p.num_items; item += item_stride) {
        const TileCoord t = decode_tile<BN, PAIR>(p, item, rank);
        const int brow = t.n0 + rank * C::B_ROWS;       // PAIR: this CTA stages its half of the weight rows
        for (int c = 0; c < nchunks; ++c) {
          const bool main = c < p.nchunk_main;
          const int ntap = main ? 9 : 1;
          const int pre = ntap > 2 ? 2 : ntap - 1;       // prefetch the next halo while this chunk's taps stream
          for (int tp = 0; tp < ntap; ++tp) {
            const int kb = main ? tp * p.nchunk_main + c : 9 * p.nchunk_main + (c - p.nchunk_main);
            HALO_TW(HALO_DBG_PROD, w_pb, ptx::mbar_wait(b_empty(bs), bph ^ 1u));
            if (rank == 0) ptx::mbar_expect_tx(b_full(bs), C::B_STAGE_BYTES * kCtas);
            const uint32_t bar = full_addr(b_full(bs));
            if constexpr (PAIR) {
              tma_load_3d_pair(&tmW, bar, sB(bs), kb * BK, brow, 0);
              tma_load_3d_pair(&tmW, bar, sB(bs) + C::B_PLANE, kb * BK, brow, 1);
            } else {
              ptx::tma_load_3d(&tmW, bar, sB(bs), kb * BK, brow, 0);
              ptx::tma_load_3d(&tmW, bar, sB(bs) + C::B_PLANE, kb * BK, brow, 1);
            }
            if (++bs == C::B_STAGES) { bs = 0; bph ^= 1u; }
            if (tp == pre) {
              if (c + 1 < nchunks) issue_A(item, c + 1);
              else if (item + item_stride < p.num_items) issue_A(item + item_stride, 0);
            }
          }
        }
      }
      if (p.dbg) { p.dbg[blockIdx.x * 16 + 4] = w_pa; p.dbg[blockIdx.x * 16 + 5] = w_pb; }
      if constexpr (PAIR) {
        // drain: every multicast commit aimed at this CTA's empty barriers has landed before the CTA may exit
        for (int i = 0; i < C::B_STAGES; ++i) {
          ptx::mbar_wait(b_empty(bs), bph ^ 1u);
          if (++bs == C::B_STAGES) { bs = 0; bph ^= 1u; }
        }
        for (int i = 0; i < A_STAGES; ++i) {
          ptx::mbar_wait(a_empty(as), aph ^ 1u);
          if (++as == A_STAGES) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == w_mma) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BM * kCtas, BN);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc_flag) {
        if constexpr (PAIR) mma_f16_ss_pair(d, da, db, idesc, acc_flag);
        else ptx::mma_f16_ss(d, da, db, idesc, acc_flag);
      };
      auto commit = [&](uint32_t bar) {
        if constexpr (PAIR) mma_commit_pair(bar);
        else ptx::mma_commit(bar);
      };
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      long long w_t = 0, w_a = 0, w_b = 0;     // cycles the issuer waited for TMEM / A halo / B block
      const long long t_begin = clock64();
      for (int item = item0; item < p.num_items; item += item_stride, ++it) {
        const int buf = it % C::NBUF;
        const uint32_t use = static_cast<uint32_t>(it / C::NBUF);
        HALO_TW(HALO_DBG_MMA, w_t, ptx::mbar_wait(t_empty(buf), (use & 1u) ^ 1u));   // epilogue has drained this accumulator buffer
        ptx::tc_fence_after();
        const uint32_t acc = tmem_acc + static_cast<uint32_t>(buf * C::NSLOT * C::SLOT_COLS);
        const uint32_t d_corr = acc + static_cast<uint32_t>(NMAIN * C::SLOT_COLS);
        int ks = 0;
        for (int c = 0; c < nchunks; ++c) {
          const bool main = c < p.nchunk_main;
          const int ntap = main ? 9 : 1;
          HALO_TW(HALO_DBG_MMA, w_a, ptx::mbar_wait(a_full(as), aph));
          for (int tp = 0; tp < ntap; ++tp) {
            HALO_TW(HALO_DBG_MMA, w_b, ptx::mbar_wait(b_full(bs), bph));
            ptx::tc_fence_after();
            // view of the halo for this tap: rows shifted by (dy+1) halo rows and (dx+1) pixels
            const int shift = main ? (tp / 3) * HALO_W + (tp % 3) : HALO_W + 1;
            const uint32_t a_hi = sA(as) + static_cast<uint32_t>(shift) * 128u;
            const uint32_t a_lo = a_hi + A_PLANE_STRIDE;
            const uint32_t b_hi = sB(bs);
            const uint32_t b_lo = b_hi + C::B_PLANE;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k, ++ks) {
              const uint32_t koff = static_cast<uint32_t>(k * UMMA_K * 2);
              const uint64_t dA_hi = make_desc_sw128(a_hi + koff, A_SBO);
              const uint64_t dA_lo = make_desc_sw128(a_lo + koff, A_SBO);
              const uint64_t dB_hi = make_desc_sw128(b_hi + koff, 1024);
              const uint64_t dB_lo = make_desc_sw128(b_lo + koff, 1024);
              if constexpr (C::FUSED) {
                constexpr uint32_t idesc2 = ptx::make_idesc_f16(BM, 2 * BN);
                ptx::mma_f16_ss(acc, dA_hi, dB_hi, idesc2, ks > 0 ? 1u : 0u);     // [main | corr] += A_hi x [B_hi ; B_lo]^T
                ptx::mma_f16_ss(d_corr, dA_lo, dB_hi, idesc, 1u);                 // corr += A_lo x B_hi^T
              } else {
                const uint32_t d_main = acc + static_cast<uint32_t>((ks % NMAIN) * C::SLOT_COLS);
                mma(d_main, dA_hi, dB_hi, ks >= NMAIN ? 1u : 0u);
                mma(d_corr, dA_hi, dB_lo, ks > 0 ? 1u : 0u);
                mma(d_corr, dA_lo, dB_hi, 1u);
              }
            }
            commit(b_empty(bs));
            if (++bs == C::B_STAGES) { bs = 0; bph ^= 1u; }
          }
          commit(a_empty(as));
          if (++as == A_STAGES) { as = 0; aph ^= 1u; }
        }
        commit(t_full(buf));
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin; p.dbg[blockIdx.x * 16 + 1] = w_t;
        p.dbg[blockIdx.x * 16 + 2] = w_a; p.dbg[blockIdx.x * 16 + 3] = w_b;
      }
    }
  } else if (XF && warp == 3) {
    // ------------------------------------------------------------------ GroupNorm table builder (XF only)
    // Per-channel scale / shift (rstd*gamma, beta - mean*rstd*gamma) of the main operand for every batch element this CTA
    // meets, same arithmetic as the standalone prep (kernels_gn.cu: fp64 group sums -> fp32 mean / rstd).  A warp of its
    // own: the fp64 divide / sqrt are subroutine calls, which would force the transform warps to spill the activations they
    // hold in flight, and the table is ready while their first loads are still on the way.
    if (xf.a.s1 != nullptr && xf.gamma != nullptr) {
      const int Ca = xf.a.C1 + xf.a.C2;
      const int cpg = Ca / kGroups;
      int cur_b = -1;
      for (int item = item0; item < p.num_items; item += item_stride) {
        const int b = decode_tile<BN, PAIR>(p, item, rank).b;
        if (b == cur_b) continue;
        cur_b = b;
        float ga[kXfMaxC / 32], be[kXfMaxC / 32];       // requested before the statistics: one global round trip in all
#pragma unroll
        for (int i = 0; i < kXfMaxC / 32; ++i) {
          const int c = lane + 32 * i;
          ga[i] = c < Ca ? __ldg(xf.gamma + c) : 0.f; be[i] = c < Ca ? __ldg(xf.beta + c) : 0.f;
        }
        double su = 0.0, sq = 0.0;                      // lane = group
        {
          const int qpg = cpg >> 2, q1 = xf.a.C1 >> 2;
          for (int jj = 0; jj < qpg; ++jj) {
            const int qd = lane * qpg + jj;
#pragma unroll
            for (int r = 0; r < kStatReplicas; ++r) {
              const double2 v = (qd < q1)
                  ? reinterpret_cast<const double2*>(qstat_slot(xf.qs1, b, r, q1))[qd]
                  : reinterpret_cast<const double2*>(qstat_slot(xf.qs2, b, r, xf.a.C2 >> 2))[qd - q1];
              su += v.x; sq += v.y;
            }
          }
        }
        const double n = static_cast<double>(p.H) * p.W * cpg;
        const double mean_d = su / n;
        double var = sq / n - mean_d * mean_d;
        if (var < 0.0) var = 0.0;
        const float mean = static_cast<float>(mean_d);
        const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kGnEps)));
        named_bar_sync(2, kXfWarps * 32 + 32);          // the transform warps no longer read the previous table
#pragma unroll
        for (int i = 0; i < kXfMaxC / 32; ++i) {
          const int c = lane + 32 * i;
          const int g = (c < Ca ? c : 0) / cpg;
          const float m = __shfl_sync(0xffffffffu, mean, g), rs = __shfl_sync(0xffffffffu, rstd, g);
          if (c < Ca) {
            const float sc = rs * ga[i];
            s_xsc[c] = sc;
            s_xsh[c] = fmaf(-m, sc, be[i]);
          }
        }
        named_bar_sync(3, kXfWarps * 32 + 32);          // table complete (bar.sync orders the shared-memory writes)
      }
    }
  } else if (is_xf) {
    // ------------------------------------------------------------------ operand transform (warps 12..19, XF only)
    // A thread owns one 16-byte column (8 channels) of every 32nd operand row: 6 rows of a 3x3 chunk (180 halo rows), 4 of
    // a 1x1 shortcut chunk (its 128 centre rows).  Work unit = a BATCH of 3 row slots (two batches per chunk), prepared as
    // straight-line code so that the three rows' dependency chains (FFMA -> EX2 -> RCP -> pack -> unpack -> pack) overlap:
    // with only two transform warps per scheduler there is no other latency hiding.  (Measured alternatives, all slower:
    // one rolled row per trip - 12 % of the cycles issuing, the rest dependency stalls; six rows unrolled with every
    // variant inlined - 80 KB of SASS, instruction-cache bound.)  The loads of the next batch are issued as soon as a
    // batch's registers are free, before the wait for the next operand stage.
    const int xt = static_cast<int>(threadIdx.x) - w_xf0 * 32;            // 0..255
    const int j = xt & 7;                      // 16-byte column of the 128-byte operand row: channels 8j .. 8j+7 of the chunk
    const int r0 = xt >> 3;                    // rows r0 + 32 i
    const bool norm_a = xf.a.s1 != nullptr && xf.gamma != nullptr;
    const int n_my_items = item0 < p.num_items ? (p.num_items - item0 + item_stride - 1) / item_stride : 0;
    constexpr int NB = XF_NB;                  // row slots per batch
    constexpr int HB = 6 / NB;                 // batches per chunk
    struct Cur { int it, c, hb; TileCoord t; };               // work item ordinal, chunk, batch + the item's tile
    auto chunk_fused = [&](int c) { return (c < p.nchunk_main ? xf.a.s1 : xf.x.s1) != nullptr; };
    auto advance = [&](Cur& k) {
      if (++k.hb == HB) {
        k.hb = 0;
        if (++k.c == nchunks) {
          k.c = 0;
          if (++k.it < n_my_items) k.t = decode_tile<BN, PAIR>(p, item0 + k.it * item_stride, rank);
        }
      }
    };
    float4 v0[NB], v1[NB];
    int meta[NB];                              // bits 0..15 operand row, bit 16 row exists, bit 17 pixel inside the image
    auto fetch = [&](const Cur& k) {
      const bool live = k.it < n_my_items && chunk_fused(k.c);
      const bool main = k.c < p.nchunk_main;
      const XfOperand& src = main ? xf.a : xf.x;
      const int cg = (main ? k.c : k.c - p.nchunk_main) * BK;
      const float* base = nullptr; int ld = 0;
      if (live) {
        if (cg < src.C1) { base = src.s1 + cg; ld = src.C1; } else { base = src.s2 + (cg - src.C1); ld = src.C2; }
        base += static_cast<size_t>(k.t.b) * p.H * p.W * ld + j * 8;
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int slot = r0 + 32 * (NB * k.hb + i);            // 0..191
        const int r = main ? slot : ((slot >> 3) + 1) * HALO_W + (slot & 7) + 1;
        const bool exists = live && (main ? slot < A_ROWS : slot < BM);
        const int hy = r / HALO_W, hx = r - hy * HALO_W;
        const int h = k.t.h0 - 1 + hy, w = k.t.w0 - 1 + hx;
        const bool inb = exists && h >= 0 && h < p.H && w >= 0 && w < p.W;
        meta[i] = r | (exists ? (1 << 16) : 0) | (inb ? (1 << 17) : 0);
        v0[i] = make_float4(0.f, 0.f, 0.f, 0.f); v1[i] = v0[i];
        if (inb) {
          const float4* g = reinterpret_cast<const float4*>(base + (static_cast<size_t>(h) * p.W + w) * ld);
          v0[i] = xf_load(g); v1[i] = xf_load(g + 1);
        }
      }
    };
    Cur P{0, 0, 0, TileCoord{0, 0, 0, 0}};
    if (n_my_items > 0) P.t = decode_tile<BN, PAIR>(p, item0, rank);
    fetch(P);
    int as = 0;
    uint32_t aph = 0;
    int cur_b = -1;
    float vmax = 0.f;
    long long w_xe = 0, w_xp = 0;              // debug: cycles waiting for a free stage / preparing batches
#pragma unroll 1
    while (P.it < n_my_items) {
      const bool main = P.c < p.nchunk_main;
      const bool fused = chunk_fused(P.c);
      if (P.hb == 0) {
        if (norm_a && P.c == 0 && P.t.b != cur_b) {
          // the scale / shift table of this batch element is built by warp 3 (above): release the old one, wait for the new
          named_bar_sync(2, kXfWarps * 32 + 32);
          named_bar_sync(3, kXfWarps * 32 + 32);
          cur_b = P.t.b;
        }
        const long long tc0 = p.dbg ? clock64() : 0;
        ptx::mbar_wait(a_empty(as), aph ^ 1u);
        if (p.dbg) w_xe += clock64() - tc0;
      }
      const long long tc1 = p.dbg ? clock64() : 0;
      if (fused) {
        const bool norm = main && norm_a;
        float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
        if (norm) {
          const float* tsc = s_xsc + P.c * BK + j * 8;
          const float* tsh = s_xsh + P.c * BK + j * 8;
          sc0 = *reinterpret_cast<const float4*>(tsc); sc1 = *reinterpret_cast<const float4*>(tsc + 4);
          sh0 = *reinterpret_cast<const float4*>(tsh); sh1 = *reinterpret_cast<const float4*>(tsh + 4);
        }
        const int silu = norm ? xf.silu : 0;
        const uint32_t stage = sA(as);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          // rows outside the image hold zeros (zero padding of the ACTIVATED tensor); the affine of a shortcut chunk is 1, 0
          float4 a0 = v0[i], a1 = v1[i];
          if (norm) { a0 = norm_act(a0, sc0, sh0, silu); a1 = norm_act(a1, sc1, sh1, silu); }
          const bool inb = (meta[i] >> 17) & 1;
          if (!inb) { a0 = make_float4(0.f, 0.f, 0.f, 0.f); a1 = a0; }
          uint2 h0, l0, h1, l1;
          split4(a0, h0, l0); split4(a1, h1, l1);
          vmax = amax4(a0, amax4(a1, vmax));
          if ((meta[i] >> 16) & 1) {
            const int r = meta[i] & 0xffff;
            const uint32_t dst = stage + static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(j ^ (r & 7)) << 4);
            ptx::st_shared_v4(dst, pack8(h0, h1));
            ptx::st_shared_v4(dst + A_PLANE_STRIDE, pack8(l0, l1));
          }
        }
      }
      const bool last = P.hb == HB - 1;
      advance(P);
      fetch(P);                                // the registers are free again: the next batch's loads fly during the hand-over
      if (last) {                              // second batch of the chunk done: hand the stage to the MMA warp
        if (fused) ptx::fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_full(as));
        if (++as == A_STAGES) { as = 0; aph ^= 1u; }
      }
      if (p.dbg) w_xp += clock64() - tc1;
    }
    if (p.dbg && xt == 0) { p.dbg[blockIdx.x * 16 + 6] = w_xe; p.dbg[blockIdx.x * 16 + 7] = w_xp; }
    if (vmax > kHalfMax && xf.overflow) atomicAdd(xf.overflow, 1ull);
  } else if (is_epi) {
    // ------------------------------------------------------------------ epilogue (EPW warps)
    // Two warps per TMEM lane quadrant, each owning half of the tile's columns (one warp and all columns in the fused-operand
    // variant), CH columns per pass: TMEM -> registers
    // (slots summed in IEEE fp32) -> padded smem staging tile -> row-contiguous float4 residual loads / output stores.
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;
    const int half = e >> 2;
    constexpr int CH = C::CH;
    constexpr int LPR = CH / 4;
    constexpr int RPI = 32 / LPR;
    constexpr int NIT = 32 / RPI;
    constexpr int NCHUNK = C::COLS_PER_WARP / CH;
    const bool active = (BN >= 32) || (half == 0);
    float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - ptx::smem_u32(smem_raw)) +
                                          A_STAGES * A_STAGE_BYTES + C::B_STAGES * C::B_STAGE_BYTES) +
                 e * 32 * C::STG_STRIDE;
    const int sub_row = lane / LPR;
    const int cj = (lane % LPR) * 4;
    const int col_base = half * C::COLS_PER_WARP;
    const float post = p.div_sqrt2 ? 0.70710678118654752440f : 1.0f;
    const int etid = static_cast<int>(threadIdx.x) - kFirstEpiWarp * 32;
    int it = 0;
    for (int item = item0; item < p.num_items; item += item_stride, ++it) {
      const TileCoord t = decode_tile<BN, PAIR>(p, item, rank);
      const int tile = PAIR ? 2 * item + rank : item;         // spreads the statistics atomics over the replicas
      const int buf = it % C::NBUF;
      const uint32_t use = static_cast<uint32_t>(it / C::NBUF);
      // the accumulator-empty barrier that counts is the leader's
      const uint32_t t_empty_bar = (PAIR && rank != 0) ? ptx::map_to_cta(t_empty(buf), 0) : t_empty(buf);
      auto release_tmem = [&]() {
        if constexpr (PAIR) { if (rank != 0) ptx::mbar_arrive_cluster(t_empty_bar); else ptx::mbar_arrive(t_empty_bar); }
        else ptx::mbar_arrive(t_empty_bar);
      };
      const float* brow = p.bias + static_cast<size_t>(t.b) * p.bias_bstride;
      long long off[NIT];
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int row = q * 32 + i * RPI + sub_row;
        const int h = t.h0 + row / TW, w = t.w0 + row % TW;
        off[i] = static_cast<long long>((static_cast<size_t>(t.b) * p.H + h) * p.W + w) * p.ldc + t.n0 + col_base + cj;
      }
      float4 res[NIT];
      auto load_res = [&](int c0) {
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.residual && t.n0 + col_base + c0 + cj < p.Cout)
            res[i] = __ldg(reinterpret_cast<const float4*>(p.residual + off[i] + c0));
        }
      };
      if (active) load_res(0);

      ptx::mbar_wait(t_full(buf), use & 1u);
      ptx::tc_fence_after();
      const uint32_t acc = tmem_acc + static_cast<uint32_t>(buf * C::NSLOT * C::SLOT_COLS);

      if (active) {
#pragma unroll 1
        for (int ci = 0; ci < NCHUNK; ++ci) {
          const int c0 = ci * CH;
          uint32_t r[CH], r2[CH];
          const uint32_t taddr = acc + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(col_base + c0);
          ptx::tmem_ld_32x32b_x16(taddr, r);
          ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(C::SLOT_COLS), r2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          if constexpr (NMAIN == 3) {
            uint32_t r3[CH];
            ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(2 * C::SLOT_COLS), r2);
            ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(3 * C::SLOT_COLS), r3);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CH; ++j)
              r[j] = __float_as_uint((__uint_as_float(r[j]) + __uint_as_float(r2[j])) + __uint_as_float(r3[j]));
          }
          if (ci == NCHUNK - 1) {
            // all TMEM reads of this warp for this tile are complete: hand the buffer back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) release_tmem();
          } else {
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < CH; j += 4)
            *reinterpret_cast<float4*>(stg + lane * C::STG_STRIDE + j) =
                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                            __uint_as_float(r[j + 3]));
          __syncwarp();
          const int n = t.n0 + col_base + c0 + cj;
          float4 cur[NIT];
#pragma unroll
          for (int i = 0; i < NIT; ++i) cur[i] = res[i];
          if (ci + 1 < NCHUNK) load_res(c0 + CH);
          float qs_s = 0.f, qs_q = 0.f;
          if (n < p.Cout) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(brow + n));
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
              const float4 a = *reinterpret_cast<const float4*>(stg + (i * RPI + sub_row) * C::STG_STRIDE + cj);
              float4 v;
              v.x = (a.x * p.wscale_inv + bv.x + cur[i].x) * post;
              v.y = (a.y * p.wscale_inv + bv.y + cur[i].y) * post;
              v.z = (a.z * p.wscale_inv + bv.z + cur[i].z) * post;
              v.w = (a.w * p.wscale_inv + bv.w + cur[i].w) * post;
              *reinterpret_cast<float4*>(p.out + off[i] + c0) = v;
              qs_s += (v.x + v.y) + (v.z + v.w);
              qs_q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
          }
          if (p.qstats) {
            // lanes sharing lane % LPR hold the same channel quad: fold the row-lanes, park the warp's partial in smem
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
              qs_s += __shfl_xor_sync(0xffffffffu, qs_s, o);
              qs_q += __shfl_xor_sync(0xffffffffu, qs_q, o);
            }
            if (sub_row == 0) {
              s_qs[it & 1][e][(c0 + cj) >> 2][0] = qs_s;
              s_qs[it & 1][e][(c0 + cj) >> 2][1] = qs_q;
            }
          }
          __syncwarp();              // staging tile is rewritten by the next pass
        }
      } else {
        // inactive half (BN < 32): it has waited for t_full like everyone else, so its arrival belongs to this phase
        __syncwarp();
        if (lane == 0) release_tmem();
      }
      if (p.qstats) {
        // fold the four quadrant warps of each column half in a fixed order, then one fp64 atomic pair per quad
        named_bar_sync(1, EPW * 32);
        constexpr int QPW = C::COLS_PER_WARP / 4;                 // quads per warp
        constexpr int NQ = ((BN >= 32 && EPW == 8) ? 2 : 1) * QPW;   // quads per tile
        if (etid < NQ) {
          const int hf = etid / QPW, qd = etid % QPW;
          const int n = t.n0 + hf * C::COLS_PER_WARP + qd * 4;
          if (n < p.Cout) {
            float as = 0.f, aq = 0.f;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) { as += s_qs[it & 1][hf * 4 + w4][qd][0]; aq += s_qs[it & 1][hf * 4 + w4][qd][1]; }
            double* dst = qstat_slot(p.qstats, t.b, tile, p.Cout >> 2) + static_cast<size_t>(n >> 2) * 2;
            atomicAdd(dst, static_cast<double>(as));
            atomicAdd(dst + 1, static_cast<double>(aq));
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();      // neither CTA's shared memory / TMEM goes away while the peer may touch it
  if (warp == w_alloc) {
    ptx::tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_acc, C::TMEM_COLS);
    else ptx::tmem_dealloc(tmem_acc, C::TMEM_COLS);
  }
}

// activations [2][B][H][W][C] fp16 -> 5-D map, box {64 ch, TW+2, TH+2, 1, 1}
bool make_halo_map(CUtensorMap* m, const __half* base, int B, int H, int W, int C, std::string* err) {
  EncodeTiledFn enc = tensor_map_encoder(err);
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)B * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(halo B=%d H=%d W=%d C=%d) failed: %d", B, H, W, C, (int)r);
      *err = buf;
    }
    return false;
  }
  return true;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int NMAIN, bool PAIR, bool XF = false>
int launch_halo(const ConvGemmArgs& a, cudaStream_t s, std::string* err) {
  constexpr int EPW = XF ? kEpiWarpsXf : kEpiWarps;           // epilogue warps
  using C = HCfg<BN, NMAIN, PAIR, EPW>;
  static PerDevice<bool> attr_done(false);
  bool& attr_set = attr_done.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, NMAIN, PAIR, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(halo): ") + cudaGetErrorString(e); return 1; }
    attr_set = true;
  }
  HaloParams p{};
  p.H = a.H; p.W = a.W;
  p.tiles_w = a.W / TW; p.tiles_h = a.H / TH;
  p.n_tiles = (a.Cout + BN - 1) / BN;
  const int m_tiles = a.B * p.tiles_w * p.tiles_h;
  p.num_items = (PAIR ? m_tiles / 2 : m_tiles) * p.n_tiles;
  p.nchunk_main = a.Cin / BK;
  p.Cout = a.Cout; p.ldc = a.ldc; p.wscale_inv = a.wscale_inv;
  p.bias = a.bias; p.bias_bstride = a.bias_bstride; p.residual = a.residual; p.out = a.out;
  p.div_sqrt2 = a.div_sqrt2; p.qstats = a.qstats;
  const bool has_x = a.X != nullptr || a.fX.s1 != nullptr;
  p.nchunk_sc = has_x ? a.Cin2 / BK : 0;
  const int K = 9 * a.Cin + (has_x ? a.Cin2 : 0);
  XfParams xf{};
  if (XF) {
    xf.a = XfOperand{a.fA.s1, a.fA.s2, a.fA.C1, a.fA.s2 ? a.fA.C2 : 0};
    xf.x = XfOperand{a.fX.s1, a.fX.s2, a.fX.C1, a.fX.s2 ? a.fX.C2 : 0};
    xf.qs1 = a.fA.qs1; xf.qs2 = a.fA.qs2; xf.gamma = a.fA.gamma; xf.beta = a.fA.beta; xf.silu = a.fA.silu;
    xf.overflow = a.overflow;
    auto bad = [&](const char* m) { if (err) *err = std::string("conv_halo (fused operand): ") + m; return 1; };
    if (a.fA.s1 && (xf.a.C1 + xf.a.C2 != a.Cin || xf.a.C1 % BK || a.Cin > kXfMaxC)) return bad("main operand channels");
    if (a.fA.s1 && a.fA.gamma && (!a.fA.beta || !a.fA.qs1 || (xf.a.C2 && !a.fA.qs2))) return bad("GroupNorm parameters / statistics missing");
    if (a.fX.s1 && (xf.x.C1 + xf.x.C2 != a.Cin2 || xf.x.C1 % BK)) return bad("shortcut operand channels");
    if (!a.fA.s1 && !a.A) return bad("no main operand");
  }
  CUtensorMap tmA, tmX, tmW;
  if (!make_weight_map(&tmW, a.Wp, a.Npad, K, C::B_ROWS, err)) return 1;
  if (a.A && !(XF && a.fA.s1)) { if (!make_halo_map(&tmA, a.A, a.B, a.H, a.W, a.Cin, err)) return 1; }
  else tmA = tmW;                                      // unused: that operand is produced in the kernel
  if (a.X && !(XF && a.fX.s1)) { if (!make_halo_map(&tmX, a.X, a.B, a.H, a.W, a.Cin2, err)) return 1; }
  else tmX = tmA;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(XF ? NUM_THREADS_XF : NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = s;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl_mode() == 1 || pdl_mode() == 4) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (PAIR) {
    const int clusters = std::min(p.num_items, num_sms() / 2);
    cfg.gridDim = dim3(2 * clusters);
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  } else {
    cfg.gridDim = dim3(std::min(p.num_items, num_sms()));
  }
  static const bool dbg = getenv("FLOWSE_CONV_DBG") != nullptr;
  long long* dbuf = nullptr;
  const size_t nctas = cfg.gridDim.x;
  if (dbg) { cudaMalloc(&dbuf, nctas * 16 * sizeof(long long)); cudaMemset(dbuf, 0, nctas * 16 * sizeof(long long)); p.dbg = dbuf; }
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, NMAIN, PAIR, XF>, tmA, tmX, tmW, p, xf);
  ++launch_counter();
  if (dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> h(nctas * 16);
    cudaMemcpy(h.data(), dbuf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbuf);
    double sum[16] = {0}; int nl = 0;
    for (size_t c = 0; c < nctas; ++c) {
      if (h[c * 16] > 0) ++nl;
      for (int k = 0; k < 16; ++k) sum[k] += static_cast<double>(h[c * 16 + k]);
    }
    if (nl == 0) nl = 1;
    const double kb_per_cta = static_cast<double>(p.num_items) * (9.0 * p.nchunk_main + p.nchunk_sc) / nl;
    fprintf(stderr, "[halo dbg%s] items=%d issuing ctas=%d kblocks/cta=%.0f (mma floor %.0f cyc) | issuer loop %.0f cyc: wait tmem %.0f, A %.0f, B %.0f | producer wait: A-free %.0f, B-free %.0f | transform: stage wait %.0f, prepare %.0f\n",
            XF ? " XF" : "", p.num_items, nl, kb_per_cta, kb_per_cta * 768.0 * (BN / 128.0), sum[0] / nl, sum[1] / nl, sum[2] / nl, sum[3] / nl,
            sum[4] / nctas, sum[5] / nctas, sum[6] / nctas, sum[7] / nctas);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("conv_halo launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace

bool conv_halo_supported(const ConvGemmArgs& a) {
  return a.ntaps == 9 && a.H % TH == 0 && a.W % TW == 0 && a.Cin % BK == 0 && ((!a.X && !a.fX.s1) || a.Cin2 % BK == 0) &&
         (a.Npad % 128 == 0 || a.Npad == 16);
}

// variant: 1 = one main accumulator (double-buffered TMEM), 3 = three rotating main accumulators,
//          2 = CTA pairs (cta_group::2, one main accumulator, double-buffered TMEM); falls back to 1 when unavailable
int launch_conv_halo(const ConvGemmArgs& a, int variant, cudaStream_t s, std::string* err) {
  if (!conv_halo_supported(a)) { if (err) *err = "conv_halo: unsupported shape"; return 1; }
  if (a.fA.s1 || a.fX.s1) {
    if (a.Npad % 128 != 0) { if (err) *err = "conv_halo: fused operands need Cout tiles of 128"; return 1; }
    return launch_halo<128, 1, false, true>(a, s, err);
  }
  if (a.Npad % 128 == 0) {
    const int m_tiles = a.B * (a.W / TW) * (a.H / TH);
    if (variant == 2 && m_tiles % 2 == 0) return launch_halo<128, 1, true>(a, s, err);
    return variant == 3 ? launch_halo<128, 3, false>(a, s, err) : launch_halo<128, 1, false>(a, s, err);
  }
  return variant == 3 ? launch_halo<16, 3, false>(a, s, err) : launch_halo<16, 1, false>(a, s, err);
}

}  // namespace flowse
