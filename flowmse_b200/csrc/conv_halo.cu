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
//    warps 4..11 epilogue (two per TMEM lane quadrant).
#include "flowse_internal.h"
#include "ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>

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

template <int BN, int NMAIN>
struct HCfg {
  static constexpr int B_PLANE = BN * 128;
  static constexpr int B_STAGE_BYTES = 2 * B_PLANE;
  static constexpr int B_STAGES = (BN >= 128) ? 3 : 8;
  static constexpr int SLOT_COLS = (BN < 32) ? 32 : BN;
  static constexpr int NSLOT = NMAIN + 1;                          // hi*hi chains + one correction accumulator
  static constexpr int NBUF = (NSLOT * SLOT_COLS * 2 <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS_RAW = NBUF * NSLOT * SLOT_COLS;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int CH = 16;
  static constexpr int STG_STRIDE = CH + 4;
  static constexpr int STG_BYTES = kEpiWarps * 32 * STG_STRIDE * 4;
  static constexpr int COLS_PER_WARP = (BN >= 32) ? BN / 2 : BN;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + STG_BYTES + 1024;
};

struct HaloParams {
  int H, W, tiles_w, tiles_h, n_tiles, num_tiles;
  int nchunk_main, nchunk_sc;
  int Cout, ldc;
  float wscale_inv;
  const float* bias;
  int bias_bstride;
  const float* residual;
  float* out;
  int div_sqrt2;
  double* qstats;
};

struct TileCoord { int b, h0, w0, n0; };

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const HaloParams& p, int tile) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  int m = tile / p.n_tiles;
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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN, int NMAIN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmW, const HaloParams p) {
  using C = HCfg<BN, NMAIN>;
  extern __shared__ uint8_t smem_raw[];
  constexpr int NBARS = 2 * A_STAGES + 2 * C::B_STAGES + 2 * C::NBUF;
  __shared__ uint64_t bars[NBARS];
  __shared__ uint32_t tmem_slot_var;
  __shared__ float s_qs[2][kEpiWarps][C::COLS_PER_WARP / 4 > 0 ? C::COLS_PER_WARP / 4 : 1][2];

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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunks = p.nchunk_main + p.nchunk_sc;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
    for (int s = 0; s < A_STAGES; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), 1); }
    for (int s = 0; s < C::B_STAGES; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1); }
    for (int s = 0; s < C::NBUF; ++s) { ptx::mbar_init(t_full(s), 1); ptx::mbar_init(t_empty(s), kEpiWarps); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto issue_A = [&](int tile, int c) {
        const TileCoord t = decode_tile<BN>(p, tile);
        ptx::mbar_wait(a_empty(as), aph ^ 1u);
        ptx::mbar_expect_tx(a_full(as), 2 * A_PLANE_BYTES);
        const bool main = c < p.nchunk_main;
        const CUtensorMap* m = main ? &tmA : &tmX;
        const int ch = main ? c : c - p.nchunk_main;
        ptx::tma_load_5d(m, a_full(as), sA(as), ch * BK, t.w0 - 1, t.h0 - 1, t.b, 0);
        ptx::tma_load_5d(m, a_full(as), sA(as) + A_PLANE_STRIDE, ch * BK, t.w0 - 1, t.h0 - 1, t.b, 1);
        if (++as == A_STAGES) { as = 0; aph ^= 1u; }
      };
      int tile = blockIdx.x;
      if (tile < p.num_tiles) issue_A(tile, 0);
      for (; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<BN>(p, tile);
        for (int c = 0; c < nchunks; ++c) {
          const bool main = c < p.nchunk_main;
          const int ntap = main ? 9 : 1;
          const int pre = ntap > 2 ? 2 : ntap - 1;       // prefetch the next halo while this chunk's taps stream
          for (int tp = 0; tp < ntap; ++tp) {
            const int kb = main ? tp * p.nchunk_main + c : 9 * p.nchunk_main + (c - p.nchunk_main);
            ptx::mbar_wait(b_empty(bs), bph ^ 1u);
            ptx::mbar_expect_tx(b_full(bs), C::B_STAGE_BYTES);
            ptx::tma_load_3d(&tmW, b_full(bs), sB(bs), kb * BK, t.n0, 0);
            ptx::tma_load_3d(&tmW, b_full(bs), sB(bs) + C::B_PLANE, kb * BK, t.n0, 1);
            if (++bs == C::B_STAGES) { bs = 0; bph ^= 1u; }
            if (tp == pre) {
              if (c + 1 < nchunks) issue_A(tile, c + 1);
              else if (tile + static_cast<int>(gridDim.x) < p.num_tiles) issue_A(tile + gridDim.x, 0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it % C::NBUF;
        const uint32_t use = static_cast<uint32_t>(it / C::NBUF);
        ptx::mbar_wait(t_empty(buf), (use & 1u) ^ 1u);        // epilogue has drained this accumulator buffer
        ptx::tc_fence_after();
        const uint32_t acc = tmem_acc + static_cast<uint32_t>(buf * C::NSLOT * C::SLOT_COLS);
        const uint32_t d_corr = acc + static_cast<uint32_t>(NMAIN * C::SLOT_COLS);
        int ks = 0;
        for (int c = 0; c < nchunks; ++c) {
          const bool main = c < p.nchunk_main;
          const int ntap = main ? 9 : 1;
          ptx::mbar_wait(a_full(as), aph);
          for (int tp = 0; tp < ntap; ++tp) {
            ptx::mbar_wait(b_full(bs), bph);
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
              const uint32_t d_main = acc + static_cast<uint32_t>((ks % NMAIN) * C::SLOT_COLS);
              ptx::mma_f16_ss(d_main, dA_hi, dB_hi, idesc, ks >= NMAIN ? 1u : 0u);
              ptx::mma_f16_ss(d_corr, dA_hi, dB_lo, idesc, ks > 0 ? 1u : 0u);
              ptx::mma_f16_ss(d_corr, dA_lo, dB_hi, idesc, 1u);
            }
            ptx::mma_commit(b_empty(bs));
            if (++bs == C::B_STAGES) { bs = 0; bph ^= 1u; }
          }
          ptx::mma_commit(a_empty(as));
          if (++as == A_STAGES) { as = 0; aph ^= 1u; }
        }
        ptx::mma_commit(t_full(buf));
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    // Two warps per TMEM lane quadrant, each owning half of the tile's columns, CH columns per pass: TMEM -> registers
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
    const int etid = threadIdx.x - kFirstEpiWarp * 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile<BN>(p, tile);
      const int buf = it % C::NBUF;
      const uint32_t use = static_cast<uint32_t>(it / C::NBUF);
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
            if (lane == 0) ptx::mbar_arrive(t_empty(buf));
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
        if (lane == 0) ptx::mbar_arrive(t_empty(buf));
      }
      if (p.qstats) {
        // fold the four quadrant warps of each column half in a fixed order, then one fp64 atomic pair per quad
        named_bar_sync(1, kEpiWarps * 32);
        constexpr int QPW = C::COLS_PER_WARP / 4;                 // quads per warp
        constexpr int NQ = (BN >= 32 ? 2 : 1) * QPW;              // quads per tile
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
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, C::TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn(std::string* err) {
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

// activations [2][B][H][W][C] fp16 -> 5-D map, box {64 ch, TW+2, TH+2, 1, 1}
bool make_halo_map(CUtensorMap* m, const __half* base, int B, int H, int W, int C, std::string* err) {
  EncodeTiledFn enc = encode_fn(err);
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

bool make_w_map(CUtensorMap* m, const __half* base, int Npad, int K, int BN, std::string* err) {
  EncodeTiledFn enc = encode_fn(err);
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Npad, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Npad * K * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(halo weights) failed: " + std::to_string((int)r);
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

template <int BN, int NMAIN>
int launch_halo(const ConvGemmArgs& a, cudaStream_t s, std::string* err) {
  using C = HCfg<BN, NMAIN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, NMAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(halo): ") + cudaGetErrorString(e); return 1; }
    attr_set = true;
  }
  HaloParams p{};
  p.H = a.H; p.W = a.W;
  p.tiles_w = a.W / TW; p.tiles_h = a.H / TH;
  p.n_tiles = (a.Cout + BN - 1) / BN;
  p.num_tiles = a.B * p.tiles_w * p.tiles_h * p.n_tiles;
  p.nchunk_main = a.Cin / BK;
  p.nchunk_sc = a.X ? a.Cin2 / BK : 0;
  p.Cout = a.Cout; p.ldc = a.ldc; p.wscale_inv = a.wscale_inv;
  p.bias = a.bias; p.bias_bstride = a.bias_bstride; p.residual = a.residual; p.out = a.out;
  p.div_sqrt2 = a.div_sqrt2; p.qstats = a.qstats;
  const int K = 9 * a.Cin + (a.X ? a.Cin2 : 0);
  CUtensorMap tmA, tmX, tmW;
  if (!make_halo_map(&tmA, a.A, a.B, a.H, a.W, a.Cin, err)) return 1;
  if (a.X) { if (!make_halo_map(&tmX, a.X, a.B, a.H, a.W, a.Cin2, err)) return 1; }
  else tmX = tmA;
  if (!make_w_map(&tmW, a.Wp, a.Npad, K, BN, err)) return 1;
  const int grid = std::min(p.num_tiles, num_sms());
  conv_halo_kernel<BN, NMAIN><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(tmA, tmX, tmW, p);
  ++launch_counter();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("conv_halo launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace

bool conv_halo_supported(const ConvGemmArgs& a) {
  return a.ntaps == 9 && a.H % TH == 0 && a.W % TW == 0 && a.Cin % BK == 0 && (!a.X || a.Cin2 % BK == 0) &&
         (a.Npad % 128 == 0 || a.Npad == 16);
}

int launch_conv_halo(const ConvGemmArgs& a, int nmain, cudaStream_t s, std::string* err) {
  if (!conv_halo_supported(a)) { if (err) *err = "conv_halo: unsupported shape"; return 1; }
  if (a.Npad % 128 == 0) return nmain == 3 ? launch_halo<128, 3>(a, s, err) : launch_halo<128, 1>(a, s, err);
  return nmain == 3 ? launch_halo<16, 3>(a, s, err) : launch_halo<16, 1>(a, s, err);
}

}  // namespace flowse
