// Implicit-GEMM convolution for the NCSN++ ResBlocks on Blackwell tensor cores (sm_100a).
//
//   out[b,h,w,n] = epilogue( sum_{tap,c} A[b, h+dy(tap), w+dx(tap), c] * Wt[n, tap, c]  (+ 1x1 shortcut blocks) )
//
// Replaces the cuDNN/oneDNN convolutions behind nn.Conv2d in the reference's ResnetBlockBigGANpp
// (/root/reference/flowmse/backbones/ncsnpp_utils/layerspp.py:259-270; conv constructors layers.py:100-124).
//
// Design (B200-first, not a translation of any library kernel):
//  * Activations are NHWC.  One CTA owns a 128-pixel (TH x TW) x BN-channel output tile; the accumulator lives in
//    TMEM (128 lanes x BN fp32 columns).
//  * No im2col: for every filter tap the A tile is ONE 5-D TMA box {64 ch, TW, TH, 1, 1} fetched at the shifted
//    coordinate (w0+dx, h0+dy); TMA's out-of-bounds zero fill IS the conv padding.  TMA writes the tile with the
//    128-byte swizzle, which is exactly the K-major SWIZZLE_128B layout tcgen05.mma consumes.
//  * fp32 parity on fp16 tensor cores: every operand is carried as an exact hi/lo fp16 pair (x = hi + lo, 22 bits
//    of significand); per K step three MMAs are issued, A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, accumulating in fp32.
//    The dropped lo*lo term is ~2^-22 relative.  Weights are pre-scaled by a power of two per conv so the lo parts
//    stay in the fp16 normal range; the scale is undone in the epilogue.
//  * The ResBlock's 1x1 shortcut conv (Conv_2) is folded into the second 3x3 conv as extra K blocks read from a
//    second tensor map, so (x_shortcut + h)/sqrt(2) costs no extra pass.
//  * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//    warps 2..5 = epilogue (TMEM -> registers -> bias/residual/scale -> global).
#include "flowse_internal.h"
#include "ptx.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

namespace flowse {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // fp16 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM lane quadrant)
constexpr int kEpiWarps = 8;
constexpr int kMainSlots = 3;
constexpr int kNumSlots = kMainSlots + 1;

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN >= 128) ? 3 : 5;
  // Accumulator slots in TMEM: the tensor core adds into its fp32 accumulator with round-toward-zero, one rounding
  // per MMA, so the bias grows with the chain length.  The hi*hi products rotate over kMainSlots accumulators and
  // the (2^-11 smaller) hi*lo + lo*hi corrections get their own; the epilogue sums the slots in IEEE fp32.
  static constexpr int SLOT_COLS = (BN < 32) ? 32 : BN;
  static constexpr int TMEM_COLS = kNumSlots * SLOT_COLS;
  // epilogue staging: 8 warps x 32 rows x (CH + 4) floats, so that global loads/stores are row-contiguous
  static constexpr int CH = 16;                             // accumulator columns handled per epilogue pass
  static constexpr int STG_STRIDE = CH + 4;                 // floats; +4 keeps both access phases conflict-free
  static constexpr int STG_BYTES = kEpiWarps * 32 * STG_STRIDE * 4;
  static constexpr int COLS_PER_WARP = (BN >= 32) ? BN / 2 : BN;   // each quadrant's columns are split over 2 warps
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024;
  // cluster split-K: fp32 tile [128][BN + 4] parked in the operand stages (+4 floats: conflict-free row-wise float4 stores)
  static constexpr int RED_STRIDE = BN + 4;
  static_assert(BM * RED_STRIDE * 4 <= STAGES * STAGE_BYTES, "reduction tile must fit in the operand stages");
};

struct GemmParams {
  int H, W, TW, TH, tiles_w, tiles_h;
  int nchunk_main, ntaps, nchunk_sc;
  int Cout, ldc;
  float wscale_inv;
  const float* bias;
  int bias_bstride;
  const float* residual;
  float* out;
  int div_sqrt2;
  long long* dbg;   // optional per-CTA phase timestamps (FLOWSE_CONV_DBG=1), else null
  // split-K (low-resolution layers: few output tiles, long K): blockIdx.z owns a contiguous range of K blocks and
  // writes its raw partial tile to partial[z][pixel][ldc]; splitk_reduce_kernel applies the epilogue.
  double* qstats;    // optional quad statistics of the output: [B][Cout/4] x {sum, sumsq}, accumulated with fp64 atomics
  int ksplit;
  float* partial;
  long long partial_plane;   // elements per split = B*H*W*ldc
  // Cluster reduction (cluster_s > 1): the ksplit CTAs of an output tile form ONE thread-block cluster (1,1,S).  Each
  // parks its raw fp32 accumulator tile in its own shared memory (the operand stages are free by then); after a
  // cluster barrier CTA z sums rows [z*128/S, (z+1)*128/S) of all S tiles through distributed shared memory
  // (ld.shared::cluster, fixed z order = the same deterministic summation as the two-pass path) and applies the
  // epilogue.  No partial planes in global memory, no second kernel.
  int cluster_s;
};

struct CtaTile { int b, h0, w0, n0; };

// Split-K cluster reduction of one thread: CTA z of an S-CTA cluster owns rows [z*128/S, (z+1)*128/S) of the tile; warp rg
// takes every 8th of them, i.e. 16/S rows, so a thread always has 16 remote float4 loads (rows x S peers), all issued
// before the first is consumed.  Summation order over the peers is fixed (z = 0, 1, ...).
template <int S, int RED_STRIDE>
__device__ __forceinline__ void cluster_reduce_rows(const GemmParams& p, const CtaTile& ct, uint32_t red0, int z, int rg,
                                                    int c4, float& qs_s, float& qs_q) {
  constexpr int ROWS = BM / S / kEpiWarps;
  static_assert(ROWS * S == 16, "16 remote loads per thread");
  const float postf = p.div_sqrt2 ? 0.70710678118654752440f : 1.0f;
  const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(ct.b) * p.bias_bstride + ct.n0) + c4);
  long long o[ROWS];
  float4 rr[ROWS], pv[ROWS][S];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int r = z * (BM / S) + rg + i * kEpiWarps;
    const int th = r / p.TW, tw = r - th * p.TW;
    const int h = ct.h0 + th, w = ct.w0 + tw;
    o[i] = (h < p.H && w < p.W)
               ? static_cast<long long>((static_cast<size_t>(ct.b) * p.H + h) * p.W + w) * p.ldc + ct.n0 + 4 * c4 : -1;
    rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o[i] >= 0) {
      if (p.residual) rr[i] = __ldg(reinterpret_cast<const float4*>(p.residual + o[i]));
      if (p.partial) {
        // partial planes in global memory (L2-resident): an SM reads L2 several times faster than a peer's shared memory
#pragma unroll
        for (int zz = 0; zz < S; ++zz)
          pv[i][zz] = __ldcg(reinterpret_cast<const float4*>(p.partial + zz * p.partial_plane + o[i]));
      } else {
        const uint32_t local = red0 + static_cast<uint32_t>((r * RED_STRIDE + 4 * c4) * 4);
#pragma unroll
        for (int zz = 0; zz < S; ++zz) pv[i][zz] = ptx::ld_dsmem_f4(ptx::map_to_cta(local, static_cast<uint32_t>(zz)));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    if (o[i] < 0) continue;
    float4 acc = pv[i][0];
#pragma unroll
    for (int zz = 1; zz < S; ++zz) { acc.x += pv[i][zz].x; acc.y += pv[i][zz].y; acc.z += pv[i][zz].z; acc.w += pv[i][zz].w; }
    float4 v;
    v.x = (acc.x + bv.x + rr[i].x) * postf; v.y = (acc.y + bv.y + rr[i].y) * postf;
    v.z = (acc.z + bv.z + rr[i].z) * postf; v.w = (acc.w + bv.w + rr[i].w) * postf;
    *reinterpret_cast<float4*>(p.out + o[i]) = v;
    qs_s += (v.x + v.y) + (v.z + v.w);
    qs_q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX,
                         const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * C::STAGES + 1];     // full[STAGES], empty[STAGES], tmem_full
  __shared__ uint32_t tmem_slot_var;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * C::STAGES);
  const uint32_t tmem_slot = ptx::smem_u32(&tmem_slot_var);
  volatile uint32_t* tmem_slot_ptr = &tmem_slot_var;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  auto stamp = [&](int slot) { if (p.dbg) p.dbg[static_cast<size_t>(cta_lin) * 8 + slot] = static_cast<long long>(ptx::globaltimer_ns()); };
  if (threadIdx.x == 0) stamp(0);

  // tile coordinates
  int m_tile = blockIdx.x;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int b = m_tile / tiles_per_img;
  m_tile -= b * tiles_per_img;
  const int h0 = (m_tile / p.tiles_w) * p.TH;
  const int w0 = (m_tile % p.tiles_w) * p.TW;
  const int n0 = blockIdx.y * BN;

  const int nkb_main = p.ntaps * p.nchunk_main;
  const int nkb_total = nkb_main + p.nchunk_sc;
  const int kb_begin = (p.ksplit > 1) ? static_cast<int>((static_cast<long long>(blockIdx.z) * nkb_total) / p.ksplit) : 0;
  const int kb_end = (p.ksplit > 1) ? static_cast<int>((static_cast<long long>(blockIdx.z + 1) * nkb_total) / p.ksplit)
                                    : nkb_total;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  pdl_wait();                 // barriers, TMEM and tensor maps were set up while the previous kernel drained
  if (threadIdx.x == 0) stamp(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sA_hi = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sA_lo = sA_hi + A_BYTES;
        const uint32_t sB_hi = sA_lo + A_BYTES;
        const uint32_t sB_lo = sB_hi + C::B_BYTES;
        ptx::mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
        if (kb < nkb_main) {
          const int tap = kb / p.nchunk_main;
          const int ch = kb - tap * p.nchunk_main;
          int dy = 0, dx = 0;
          if (p.ntaps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
          ptx::tma_load_5d(&tmA, full_bar(stage), sA_hi, ch * BK, w0 + dx, h0 + dy, b, 0);
          ptx::tma_load_5d(&tmA, full_bar(stage), sA_lo, ch * BK, w0 + dx, h0 + dy, b, 1);
        } else {
          const int ch = kb - nkb_main;
          ptx::tma_load_5d(&tmX, full_bar(stage), sA_hi, ch * BK, w0, h0, b, 0);
          ptx::tma_load_5d(&tmX, full_bar(stage), sA_lo, ch * BK, w0, h0, b, 1);
        }
        ptx::tma_load_3d(&tmW, full_bar(stage), sB_hi, kb * BK, n0, 0);
        ptx::tma_load_3d(&tmW, full_bar(stage), sB_lo, kb * BK, n0, 1);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        if (kb == kb_begin) stamp(2);
        const uint32_t sA_hi = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sA_lo = sA_hi + A_BYTES;
        const uint32_t sB_hi = sA_lo + A_BYTES;
        const uint32_t sB_lo = sB_hi + C::B_BYTES;
        const uint64_t dA_hi = ptx::make_smem_desc_sw128(sA_hi);
        const uint64_t dA_lo = ptx::make_smem_desc_sw128(sA_lo);
        const uint64_t dB_hi = ptx::make_smem_desc_sw128(sB_hi);
        const uint64_t dB_lo = ptx::make_smem_desc_sw128(sB_lo);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t koff = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);   // 32 B per K step
          const int ks = (kb - kb_begin) * (BK / UMMA_K) + k;   // K step index local to this CTA
          const uint32_t d_main = tmem_acc + static_cast<uint32_t>((ks % kMainSlots) * C::SLOT_COLS);
          const uint32_t d_corr = tmem_acc + static_cast<uint32_t>(kMainSlots * C::SLOT_COLS);
          ptx::mma_f16_ss(d_main, dA_hi + koff, dB_hi + koff, idesc, ks >= kMainSlots ? 1u : 0u);
          ptx::mma_f16_ss(d_corr, dA_hi + koff, dB_lo + koff, idesc, ks > 0 ? 1u : 0u);
          ptx::mma_f16_ss(d_corr, dA_lo + koff, dB_hi + koff, idesc, 1u);
        }
        ptx::mma_commit(empty_bar(stage));          // frees the smem stage when these MMAs retire
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      ptx::mma_commit(tmem_full_bar);               // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quadrant, each owning half of the tile's columns, CH columns per pass.
    // Phase A: TMEM -> registers (lane = tile row), the accumulator slots summed in fp32, written to a padded smem
    //          staging tile.  Phase B: the warp re-reads the tile row-wise so that every global access (residual,
    //          output) is a contiguous 64-byte run per row: LPR lanes cover one row's CH columns, RPI rows per pass.
    // Everything that does not depend on the accumulator (pixel offsets, bias, first residual block) is done
    // before waiting for the MMAs, i.e. overlapped with the main loop.
    const int e = warp - 2;
    const int q = warp & 3;                          // TMEM lane quadrant this warp may access
    const int half = e >> 2;
    constexpr int CH = C::CH;
    constexpr int LPR = CH / 4;                      // lanes per row (float4 each)
    constexpr int RPI = 32 / LPR;                    // rows per pass
    constexpr int NIT = 32 / RPI;                    // passes over the 32 rows
    constexpr int NCHUNK = C::COLS_PER_WARP / CH;
    const bool active = (BN >= 32) || (half == 0);
    float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - ptx::smem_u32(smem_raw)) +
                                          C::STAGES * C::STAGE_BYTES) + e * 32 * C::STG_STRIDE;
    const float* brow = p.bias + static_cast<size_t>(b) * p.bias_bstride;
    const int sub_row = lane / LPR;
    const int cj = (lane % LPR) * 4;
    const int col_base = half * C::COLS_PER_WARP;
    long long off[NIT];                              // element offset of (pixel, n0 + col_base + cj), -1 if outside
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int row = q * 32 + it * RPI + sub_row;
      const int th = row / p.TW;
      const int tw = row - th * p.TW;
      const int h = h0 + th, w = w0 + tw;
      off[it] = (h < p.H && w < p.W)
                    ? static_cast<long long>((static_cast<size_t>(b) * p.H + h) * p.W + w) * p.ldc + n0 + col_base + cj
                    : -1;
    }
    float4 res[NIT];
    auto load_res = [&](int c0) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.residual && !p.partial && p.cluster_s <= 1 && off[it] >= 0 && n0 + col_base + c0 + cj < p.Cout)
          res[it] = __ldg(reinterpret_cast<const float4*>(p.residual + off[it] + c0));
      }
    };
    if (active) load_res(0);
    const float post = p.div_sqrt2 ? 0.70710678118654752440f : 1.0f;

    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    if (threadIdx.x == 64) stamp(3);

    if (active) {
#pragma unroll 1
      for (int ci = 0; ci < NCHUNK; ++ci) {
        const int c0 = ci * CH;
        uint32_t r[CH], r2[CH];
        const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(col_base + c0);
        ptx::tmem_ld_32x32b_x16(taddr, r);
        ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(C::SLOT_COLS), r2);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CH; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        uint32_t r3[CH];
        ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(2 * C::SLOT_COLS), r2);
        ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(3 * C::SLOT_COLS), r3);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CH; ++j)
          r[j] = __float_as_uint((__uint_as_float(r[j]) + __uint_as_float(r2[j])) + __uint_as_float(r3[j]));
        if (p.cluster_s > 1 && !p.partial) {
          // cluster reduction through DSMEM: park the raw tile (row = TMEM lane) in the now idle operand stages
          float* red = reinterpret_cast<float*>(smem_raw + (smem_base - ptx::smem_u32(smem_raw))) +
                       (q * 32 + lane) * C::RED_STRIDE + col_base + c0;
#pragma unroll
          for (int j = 0; j < CH; j += 4)
            *reinterpret_cast<float4*>(red + j) =
                make_float4(__uint_as_float(r[j]) * p.wscale_inv, __uint_as_float(r[j + 1]) * p.wscale_inv,
                            __uint_as_float(r[j + 2]) * p.wscale_inv, __uint_as_float(r[j + 3]) * p.wscale_inv);
          continue;
        }
        __syncwarp();                                // previous pass has finished reading the staging tile
#pragma unroll
        for (int j = 0; j < CH; j += 4)
          *reinterpret_cast<float4*>(stg + lane * C::STG_STRIDE + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                          __uint_as_float(r[j + 3]));
        __syncwarp();
        const int n = n0 + col_base + c0 + cj;
        float4 cur[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) cur[it] = res[it];
        if (ci + 1 < NCHUNK) load_res(c0 + CH);        // next pass's residual block in flight during this one
        if (p.partial) {
          if (n < p.Cout) {
            float* dst = p.partial + static_cast<long long>(blockIdx.z) * p.partial_plane;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
              if (off[it] >= 0) {
                const float4 a = *reinterpret_cast<const float4*>(stg + (it * RPI + sub_row) * C::STG_STRIDE + cj);
                *reinterpret_cast<float4*>(dst + off[it] + c0) =
                    make_float4(a.x * p.wscale_inv, a.y * p.wscale_inv, a.z * p.wscale_inv, a.w * p.wscale_inv);
              }
            }
          }
        } else if (n < p.Cout) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(brow + n));
          float qs_s = 0.f, qs_q = 0.f;
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            if (off[it] >= 0) {
              const float4 a = *reinterpret_cast<const float4*>(stg + (it * RPI + sub_row) * C::STG_STRIDE + cj);
              float4 v;
              v.x = (a.x * p.wscale_inv + bv.x + cur[it].x) * post;
              v.y = (a.y * p.wscale_inv + bv.y + cur[it].y) * post;
              v.z = (a.z * p.wscale_inv + bv.z + cur[it].z) * post;
              v.w = (a.w * p.wscale_inv + bv.w + cur[it].w) * post;
              *reinterpret_cast<float4*>(p.out + off[it] + c0) = v;
              qs_s += (v.x + v.y) + (v.z + v.w);
              qs_q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
          }
          if (p.qstats) {
            // lanes sharing lane % LPR hold the same channel quad: fold the RPI row-lanes, then one fp64 atomic pair
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
              qs_s += __shfl_xor_sync(0xffffffffu, qs_s, o);
              qs_q += __shfl_xor_sync(0xffffffffu, qs_q, o);
            }
            if (sub_row == 0) {
              double* dst = qstat_slot(p.qstats, b, blockIdx.x, p.Cout >> 2) + static_cast<size_t>(n >> 2) * 2;
              atomicAdd(dst, static_cast<double>(qs_s));
              atomicAdd(dst + 1, static_cast<double>(qs_q));
            }
          }
        }
      }
    }
  }

  if (p.cluster_s > 1) {
    // ---------------------------------------------------------------- split-K reduction through distributed smem
    __shared__ float s_red[kEpiWarps][32][2];
    const int S = p.cluster_s;
    __syncwarp();
    ptx::cluster_sync_all();                                    // every CTA's tile is parked and visible cluster-wide
    if (threadIdx.x == 64) stamp(6);
    if (warp >= 2) {
      const int etid = threadIdx.x - 64;
      const int z = static_cast<int>(ptx::cluster_ctarank());
      const int ncol4 = min(BN, p.Cout - n0) >> 2;         // float4 columns of this tile
      const int c4 = etid & 31, rg = etid >> 5;            // fixed column quad per thread, 8 row groups
      const uint32_t red0 = smem_base;
      float qs_s = 0.f, qs_q = 0.f;
      if (c4 < ncol4) {
        const CtaTile ct{b, h0, w0, n0};
        switch (S) {
          case 2: cluster_reduce_rows<2, C::RED_STRIDE>(p, ct, red0, z, rg, c4, qs_s, qs_q); break;
          case 4: cluster_reduce_rows<4, C::RED_STRIDE>(p, ct, red0, z, rg, c4, qs_s, qs_q); break;
          case 8: cluster_reduce_rows<8, C::RED_STRIDE>(p, ct, red0, z, rg, c4, qs_s, qs_q); break;
          default: cluster_reduce_rows<16, C::RED_STRIDE>(p, ct, red0, z, rg, c4, qs_s, qs_q); break;
        }
      }
      if (p.qstats) {
        s_red[rg][c4][0] = qs_s; s_red[rg][c4][1] = qs_q;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (etid < ncol4) {
          float as = 0.f, aq = 0.f;
#pragma unroll
          for (int g = 0; g < kEpiWarps; ++g) { as += s_red[g][etid][0]; aq += s_red[g][etid][1]; }
          const int tile = blockIdx.y * gridDim.x + blockIdx.x;
          double* dst = qstat_slot(p.qstats, b, tile * S + z, p.Cout >> 2) + static_cast<size_t>((n0 >> 2) + etid) * 2;
          atomicAdd(dst, static_cast<double>(as));
          atomicAdd(dst + 1, static_cast<double>(aq));
        }
      }
    }
    __syncwarp();
    if (!p.partial) ptx::cluster_sync_all();                    // no CTA's shared memory goes away while a peer still reads it
  }

  if (threadIdx.x == 64) stamp(4);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, C::TMEM_COLS);
  }
  if (threadIdx.x == 32) stamp(5);
}

// out = epilogue(sum_z partial[z]) for split-K launches; fixed summation order (deterministic).  grid = (blocks, B).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int S, long long plane, const float* __restrict__ bias,
                     int bias_bstride, const float* __restrict__ residual, float post, float* __restrict__ out,
                     int n4_per_batch, int ldc4, double* __restrict__ qstats) {
  __shared__ float ss[256], sq[256];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  float ls = 0.f, lq = 0.f;
  if (j < n4_per_batch) {
    const long long i = static_cast<long long>(b) * n4_per_batch + j;
    float4 acc = __ldcg(reinterpret_cast<const float4*>(partial) + i);
    for (int z = 1; z < S; ++z) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + z * plane) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int n = (j % ldc4) * 4;
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + static_cast<size_t>(b) * bias_bstride + n));
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (residual) r = __ldg(reinterpret_cast<const float4*>(residual) + i);
    float4 v;
    v.x = (acc.x + bv.x + r.x) * post; v.y = (acc.y + bv.y + r.y) * post;
    v.z = (acc.z + bv.z + r.z) * post; v.w = (acc.w + bv.w + r.w) * post;
    reinterpret_cast<float4*>(out)[i] = v;
    ls = (v.x + v.y) + (v.z + v.w);
    lq = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (qstats) {            // block covers 256 / ldc4 pixels x ldc4 quads (256 % ldc4 == 0, rows block-aligned)
    ss[threadIdx.x] = ls; sq[threadIdx.x] = lq;
    __syncthreads();
    if (threadIdx.x < ldc4) {
      float as = 0.f, aq = 0.f;
      for (int t = threadIdx.x; t < 256; t += ldc4) { as += ss[t]; aq += sq[t]; }
      double* dst = qstat_slot(qstats, b, blockIdx.x, ldc4) + static_cast<size_t>(threadIdx.x) * 2;
      atomicAdd(dst, static_cast<double>(as));
      atomicAdd(dst + 1, static_cast<double>(aq));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Debug / cross-check kernel: same operands, plain fp32 SIMT loops.
// ------------------------------------------------------------------------------------------------
__global__ void conv_gemm_simt_kernel(const __half* __restrict__ A, const __half* __restrict__ X,
                                      const __half* __restrict__ Wp, int Cin, int Cin2, int ntaps, int Npad, int K,
                                      GemmParams p, int B) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t planeA = static_cast<size_t>(B) * p.H * p.W * Cin;
  const size_t planeX = static_cast<size_t>(B) * p.H * p.W * Cin2;
  const size_t planeW = static_cast<size_t>(Npad) * K;
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  const size_t pix = blockIdx.x;
  const int b = pix / (static_cast<size_t>(p.H) * p.W);
  const int hw = pix - static_cast<size_t>(b) * p.H * p.W;
  const int h = hw / p.W, w = hw % p.W;
  if (n >= p.Cout) return;
  const __half* wh = Wp + static_cast<size_t>(n) * K;
  const __half* wl = wh + planeW;
  float acc = 0.f;
  for (int tap = 0; tap < ntaps; ++tap) {
    int dy = 0, dx = 0;
    if (ntaps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
    const int hh = h + dy, ww = w + dx;
    if (hh < 0 || hh >= p.H || ww < 0 || ww >= p.W) continue;
    const __half* ah = A + ((static_cast<size_t>(b) * p.H + hh) * p.W + ww) * Cin;
    const __half* al = ah + planeA;
    for (int c = 0; c < Cin; ++c) {
      const float a = __half2float(ah[c]) + __half2float(al[c]);
      const float wv = __half2float(wh[tap * Cin + c]) + __half2float(wl[tap * Cin + c]);
      acc = fmaf(a, wv, acc);
    }
  }
  if (Cin2 > 0) {
    const __half* xh = X + pix * Cin2;
    const __half* xl = xh + planeX;
    for (int c = 0; c < Cin2; ++c) {
      const float a = __half2float(xh[c]) + __half2float(xl[c]);
      const float wv = __half2float(wh[ntaps * Cin + c]) + __half2float(wl[ntaps * Cin + c]);
      acc = fmaf(a, wv, acc);
    }
  }
  float v = acc * p.wscale_inv + p.bias[static_cast<size_t>(b) * p.bias_bstride + n];
  if (p.residual) v += p.residual[pix * p.ldc + n];
  if (p.div_sqrt2) v = __fdiv_rn(v, kSqrt2);
  p.out[pix * p.ldc + n] = v;
  if (p.qstats) {      // debug path: one fp64 atomic pair per element
    double* dst = qstat_slot(p.qstats, b, static_cast<int>(pix), p.Cout >> 2) + static_cast<size_t>(n >> 2) * 2;
    atomicAdd(dst, static_cast<double>(v));
    atomicAdd(dst + 1, static_cast<double>(v) * static_cast<double>(v));
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
// activations [2][B][H][W][C] fp16 -> 5-D map, box {64, TW, TH, 1, 1}
bool make_act_map(CUtensorMap* m, const __half* base, int B, int H, int W, int C, int TW, int TH, std::string* err) {
  EncodeTiledFn enc = tensor_map_encoder(err);
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)B * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(act B=%d H=%d W=%d C=%d TW=%d TH=%d) failed: %d", B, H, W, C,
               TW, TH, (int)r);
      *err = buf;
    }
    return false;
  }
  return true;
}

void choose_tile(int H, int W, int& TW, int& TH) {
  long best = LONG_MAX;
  TW = 128; TH = 1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = BM / tw;
    const long area = static_cast<long>((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (area < best) { best = area; TW = tw; TH = th; }
  }
}

GemmParams make_params(const ConvGemmArgs& a) {
  GemmParams p{};
  p.H = a.H; p.W = a.W;
  choose_tile(a.H, a.W, p.TW, p.TH);
  p.tiles_w = (a.W + p.TW - 1) / p.TW;
  p.tiles_h = (a.H + p.TH - 1) / p.TH;
  p.nchunk_main = a.Cin / BK;
  p.ntaps = a.ntaps;
  p.nchunk_sc = a.X ? a.Cin2 / BK : 0;
  p.Cout = a.Cout; p.ldc = a.ldc;
  p.wscale_inv = a.wscale_inv;
  p.bias = a.bias; p.bias_bstride = a.bias_bstride;
  p.residual = a.residual;
  p.out = a.out;
  p.div_sqrt2 = a.div_sqrt2;
  p.dbg = nullptr;
  p.ksplit = 1; p.partial = nullptr; p.partial_plane = 0; p.cluster_s = 0;
  p.qstats = a.qstats;
  return p;
}

bool check_args(const ConvGemmArgs& a, std::string* err) {
  auto fail = [&](const char* m) { if (err) *err = std::string("conv_gemm: ") + m; return false; };
  if (a.Cin <= 0 || a.Cin % BK) return fail("Cin must be a positive multiple of 64");
  if (a.X && (a.Cin2 <= 0 || a.Cin2 % BK)) return fail("Cin2 must be a positive multiple of 64");
  if (a.ntaps != 1 && a.ntaps != 9) return fail("ntaps must be 1 or 9");
  if (a.Cout % 4 || a.ldc % 4) return fail("Cout and ldc must be multiples of 4");
  if (a.Cout > a.Npad) return fail("Cout > Npad");
  if (!a.bias) return fail("bias is required");
  if (a.B <= 0 || a.H <= 0 || a.W <= 0) return fail("empty problem");
  return true;
}

// Number of 16-CTA clusters (1,1,16; a non-portable size) of this kernel the device can run concurrently, 0 when the size
// is not available.  Queried once per instantiation.  Clusters of up to 8 CTAs are always schedulable.
template <int BN>
int max_clusters16() {
  static PerDevice<int> cache(-1);
  int& cached = cache.get();
  if (cached >= 0) return cached;
  using C = Cfg<BN>;
  cached = 0;
  if (getenv("FLOWSE_CLUSTER16") && getenv("FLOWSE_CLUSTER16")[0] == '0') return cached;
  if (cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<BN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return cached;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, 1, 16); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 16;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, conv_gemm_tcgen05_kernel<BN>, &cfg);
  if (e != cudaSuccess) cudaGetLastError();
  cached = (e == cudaSuccess) ? n : 0;
  return cached;
}

template <int BN>
int launch_bn(const ConvGemmArgs& a, cudaStream_t s, std::string* err) {
  using C = Cfg<BN>;
  static PerDevice<bool> attr_done(false);
  bool& attr_set = attr_done.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_set = true;
  }
  GemmParams p = make_params(a);
  const int K = a.ntaps * a.Cin + (a.X ? a.Cin2 : 0);
  CUtensorMap tmA, tmX, tmW;
  if (!make_act_map(&tmA, a.A, a.B, a.H, a.W, a.Cin, p.TW, p.TH, err)) return 1;
  if (a.X) { if (!make_act_map(&tmX, a.X, a.B, a.H, a.W, a.Cin2, p.TW, p.TH, err)) return 1; }
  else tmX = tmA;
  if (!make_weight_map(&tmW, a.Wp, a.Npad, K, BN, err)) return 1;
  dim3 grid(a.B * p.tiles_w * p.tiles_h, (a.Cout + BN - 1) / BN);
  // split-K when the output has too few tiles to fill the GPU (needs ldc == Cout so partial planes are dense)
  const int tiles = grid.x * grid.y;
  const int nkb_total = a.ntaps * (a.Cin / BK) + (a.X ? a.Cin2 / BK : 0);
  int S = 1;
  if (a.splitk_scratch && tiles <= 74 && a.ldc == a.Cout && (!a.qstats || 256 % (a.ldc / 4) == 0)) {
    S = std::min(148 / tiles, nkb_total / 4);
    const long long plane = static_cast<long long>(a.B) * a.H * a.W * a.ldc;
    while (S > 1 && static_cast<size_t>(S) * plane > a.splitk_scratch_elems) --S;
    if (S < 2) S = 1;
    if (S > 1) { p.ksplit = S; p.partial = a.splitk_scratch; p.partial_plane = plane; grid.z = S; }
  }
  // Cluster reduction: the S CTAs of a tile become one cluster (1,1,S) and reduce through distributed shared memory.
  // S is capped by the cluster size the device can co-schedule with this kernel's shared-memory footprint (16 is a
  // non-portable size; 8 is always available).  FLOWSE_SPLITK=2pass keeps the partial-plane + reduce-kernel path.
  // Two ways for the cluster to exchange its partial tiles: "dsmem" (default) parks them in shared memory and reads the
  // peers' copies (ld.shared::cluster); "l2" writes them to the L2-resident scratch planes and reads those back after the
  // cluster barrier.  Measured on B200 (tools/dbg_conv.py, grid (4,2,8), K = 2304): dsmem parks in 1.5 us and reduces in
  // 4.0 us (distributed shared memory moves ~20 B/cycle/SM), l2 parks in 3.6 us and reduces in 2.0 us - the same 5.5 us;
  // end to end dsmem is 0.5 % faster (22.84 vs 22.96 ms per sampler call) and needs no scratch.  Both beat the
  // two-pass path (FLOWSE_SPLITK=2pass: partial planes + splitk_reduce_kernel, 23.28 ms).
  static const int splitk_mode = [] {
    const char* e = getenv("FLOWSE_SPLITK");
    if (e && !strcmp(e, "2pass")) return 0;
    if (e && !strcmp(e, "dsmem")) return 1;
    if (e && !strcmp(e, "l2")) return 2;
    return 1;
  }();
  bool cluster_reduce = false;
  if (splitk_mode >= 1 && tiles <= 74 && nkb_total >= 8) {
    const int max_cluster = (tiles <= max_clusters16<BN>()) ? 16 : 8;
    int Sc = std::min(std::min(148 / tiles, nkb_total / 4), max_cluster);
    while (Sc & (Sc - 1)) Sc &= Sc - 1;                  // cluster sizes: powers of two
    if (Sc >= 2) {
      cluster_reduce = true;
      S = Sc;
      p.ksplit = S; p.cluster_s = S; p.partial = nullptr; p.partial_plane = 0; grid.z = S;
      const long long plane = static_cast<long long>(a.B) * a.H * a.W * a.ldc;
      if (splitk_mode == 2 && a.splitk_scratch && a.ldc == a.Cout &&
          static_cast<size_t>(S) * plane <= a.splitk_scratch_elems) {
        p.partial = a.splitk_scratch; p.partial_plane = plane;
      }
    }
  }
  static const bool dbg = getenv("FLOWSE_CONV_DBG") != nullptr;
  long long* dbuf = nullptr;
  const size_t ncta = static_cast<size_t>(grid.x) * grid.y * grid.z;
  if (dbg) { cudaMalloc(&dbuf, ncta * 8 * sizeof(long long)); cudaMemset(dbuf, 0, ncta * 8 * sizeof(long long)); p.dbg = dbuf; }
  if (cluster_reduce) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = static_cast<unsigned>(S);
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_mode() >= 1 ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    ++launch_counter();
    cudaError_t le = cudaLaunchKernelEx(&cfg, conv_gemm_tcgen05_kernel<BN>, tmA, tmX, tmW, p);
    if (le != cudaSuccess) { if (err) *err = std::string("conv_gemm cluster launch: ") + cudaGetErrorString(le); return 1; }
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_mode() >= 1 ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ++launch_counter();
    cudaLaunchKernelEx(&cfg, conv_gemm_tcgen05_kernel<BN>, tmA, tmX, tmW, p);
  }
  if (dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> h(ncta * 8);
    cudaMemcpy(h.data(), dbuf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbuf);
    double ph[5] = {0, 0, 0, 0, 0};
    long long tmin = h[0], tmax = 0;
    for (size_t c = 0; c < ncta; ++c) {
      for (int k = 0; k < 5; ++k) ph[k] += static_cast<double>(h[c * 8 + k + 1] - h[c * 8 + k]);
      tmin = std::min(tmin, h[c * 8]); tmax = std::max(tmax, h[c * 8 + 5]);
    }
    double park = 0.0;
    if (cluster_reduce) for (size_t c = 0; c < ncta; ++c) park += static_cast<double>(h[c * 8 + 6] - h[c * 8 + 3]);
    fprintf(stderr, "[conv dbg] grid=(%u,%u,%u) kb=%d  setup %.2f us | first-data %.2f | mainloop %.2f | epilogue %.2f (park+cluster barrier %.2f) | teardown %.2f | kernel span %.2f us\n",
            grid.x, grid.y, grid.z, p.ntaps * p.nchunk_main + p.nchunk_sc, ph[0] / ncta / 1e3, ph[1] / ncta / 1e3, ph[2] / ncta / 1e3,
            ph[3] / ncta / 1e3, park / ncta / 1e3, ph[4] / ncta / 1e3, (tmax - tmin) / 1e3);
  }
  if (S > 1 && !cluster_reduce) {
    const int n4b = a.H * a.W * a.ldc / 4;
    const bool can_stats = a.qstats && (256 % (a.ldc / 4) == 0);
    dim3 rgrid((n4b + 255) / 256, a.B);
    launch_k(splitk_reduce_kernel, rgrid, dim3(256), 0, s, static_cast<const float*>(p.partial), S, p.partial_plane,
             a.bias, a.bias_bstride, a.residual, a.div_sqrt2 ? 0.70710678118654752440f : 1.0f, a.out, n4b, a.ldc / 4,
             can_stats ? a.qstats : static_cast<double*>(nullptr));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("conv_gemm launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Tensor-map helpers shared by the two conv kernels (declared in flowse_internal.h)
// ------------------------------------------------------------------------------------------------
EncodeTiledFn tensor_map_encoder(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
    if (err) *err = std::string("cuTensorMapEncodeTiled not available: ") + cudaGetErrorString(e);
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// weights [2][Npad][K] fp16 -> 3-D map, box {64, rows, 1}
bool make_weight_map(CUtensorMap* m, const __half* base, int Npad, int K, int rows, std::string* err) {
  EncodeTiledFn enc = tensor_map_encoder(err);
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Npad, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Npad * K * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(weights Npad=%d K=%d rows=%d) failed: %d", Npad, K, rows, (int)r);
      *err = buf;
    }
    return false;
  }
  return true;
}

int launch_conv_gemm(const ConvGemmArgs& a, cudaStream_t s, std::string* err) {
  if (!check_args(a, err)) return 1;
  if (a.Npad % 128 == 0) return launch_bn<128>(a, s, err);
  if (a.Npad % 16 == 0) return launch_bn<16>(a, s, err);
  if (err) *err = "conv_gemm: Npad must be a multiple of 16";
  return 1;
}

int launch_conv_gemm_simt(const ConvGemmArgs& a, cudaStream_t s, std::string* err) {
  if (!check_args(a, err)) return 1;
  GemmParams p = make_params(a);
  const int K = a.ntaps * a.Cin + (a.X ? a.Cin2 : 0);
  const int threads = std::min(128, ((a.Cout + 31) / 32) * 32);
  dim3 grid(static_cast<unsigned>(static_cast<size_t>(a.B) * a.H * a.W), (a.Cout + threads - 1) / threads);
  launch_k(conv_gemm_simt_kernel, grid, dim3(threads), 0, s, a.A, a.X, a.Wp, a.Cin, a.X ? a.Cin2 : 0, a.ntaps, a.Npad, K, p,
           a.B);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("conv_gemm_simt launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

// fp32 -> (hi, lo) fp16 pair, host side (round-to-nearest-even both times).
static inline void split_half_host(float v, __half* hi, __half* lo) {
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn(v - __half2float(h));
}

int pack_conv_weights_host(const float* w_main, int Cout, int Cin, int ntaps, const float* w_sc, int Cin2, int Npad,
                           __half* out_hi, __half* out_lo) {
  const int K = ntaps * Cin + (w_sc ? Cin2 : 0);
  float amax = 0.f;
  const size_t n_main = static_cast<size_t>(Cout) * Cin * ntaps;
  for (size_t i = 0; i < n_main; ++i) amax = std::max(amax, std::fabs(w_main[i]));
  if (w_sc) for (size_t i = 0; i < static_cast<size_t>(Cout) * Cin2; ++i) amax = std::max(amax, std::fabs(w_sc[i]));
  // scale so that max|w| lands in [2^13, 2^14): hi stays finite, lo parts stay normal down to 2^-14/2^13 of max
  int e = 0;
  if (amax > 0.f && std::isfinite(amax)) {
    int ex;
    std::frexp(amax, &ex);      // amax = m * 2^ex, m in [0.5, 1)
    e = 14 - ex;
  }
  const float scale = std::ldexp(1.0f, e);
  std::memset(out_hi, 0, sizeof(__half) * static_cast<size_t>(Npad) * K);
  std::memset(out_lo, 0, sizeof(__half) * static_cast<size_t>(Npad) * K);
  for (int n = 0; n < Cout; ++n) {
    __half* rh = out_hi + static_cast<size_t>(n) * K;
    __half* rl = out_lo + static_cast<size_t>(n) * K;
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < ntaps; ++t) {
        // PyTorch layout [Cout][Cin][kh][kw], tap = kh*3+kw  ->  K index tap*Cin + c
        const float v = w_main[(static_cast<size_t>(n) * Cin + c) * ntaps + t] * scale;
        split_half_host(v, rh + t * Cin + c, rl + t * Cin + c);
      }
    if (w_sc)
      for (int c = 0; c < Cin2; ++c) {
        const float v = w_sc[static_cast<size_t>(n) * Cin2 + c] * scale;
        split_half_host(v, rh + ntaps * Cin + c, rl + ntaps * Cin + c);
      }
  }
  return e;
}

}  // namespace flowse
