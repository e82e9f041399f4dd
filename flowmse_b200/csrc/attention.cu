// Single-head spatial self-attention block of NCSN++ (AttnBlockpp, /root/reference/flowmse/backbones/ncsnpp_utils/
// layerspp.py:62-91; NIN: layers.py:546-555) in TWO kernels instead of six:
//
//   attn_qkv_kernel   h = GroupNorm(x);  q, k, v = NIN_0..2(h)              (layerspp.py:76-79)
//   attn_core_kernel  w = softmax(q k^T C^-1/2);  h = w v;  out = (x + NIN_3(h)) / sqrt(2)   (layerspp.py:81-91)
//                     + quad statistics of `out` for the next GroupNorm
//
// 1.63 GFLOP per network evaluation (0.15 % of the FLOPs): exact fp32 SIMT FMAs, no tensor cores.  What the old path
// (four GEMM launches + a softmax kernel + a prep kernel, scores round-tripping through global memory) paid for was
// launches and latency; here a CTA owns R tokens (query rows; R = 8, 4 or 2, the largest that still gives the grid about
// one CTA per SM - both kernels are FMA-issue bound, so at B = 1 halving R halves the time) and keeps everything of those
// rows on chip.
//
// Data movement is arranged so that NO operand tile is staged in shared memory:
//   * the per-CTA operands that every thread needs (the 8 normalised rows, the 8 query rows, the 8 probability rows, the
//     8 context rows) live in shared memory TRANSPOSED, [channel or key][8 rows], and are read as two broadcast LDS.128;
//   * the streamed operands (weights W[c][n], keys kT[c][l], values v[l][c]) are read straight from global memory / L2
//     with the thread index on the contiguous axis (fully coalesced, each byte used once per CTA).  For that the QKV
//     kernel writes the keys TRANSPOSED, kT[b][c][l].
// Per streamed element a warp issues 8 (16, 24) FMAs against 1 coalesced global load and 2 broadcast shared loads, i.e.
// the kernels are bound by the FMA pipe, not by shared memory.  Softmax: one warp per query row, warp-shuffle max / sum.
#include "flowse_internal.h"

#include <algorithm>
#include <cmath>

namespace flowse {

namespace {

constexpr int kC = 256;          // channels of every attention block of the default config
constexpr int kRMax = 8;         // most tokens (query rows) per CTA
constexpr int kThreads = 256;
// Tunables, measured at B = 1 (same box, attention per evaluation): 16 loads per batch and >= 96 CTAs (4 rows per CTA at
// 512 tokens) 0.200 ms; 8 loads per batch, >= 200 CTAs (2 rows per CTA, two CTAs per SM) 0.219-0.223 ms.
#ifndef ATTN_U
#define ATTN_U 16
#endif
#ifndef ATTN_MIN_CTAS
#define ATTN_MIN_CTAS 96
#endif
constexpr int kStreamU = ATTN_U; // 16-byte loads per thread and batch of the streaming loops (2 batches in flight)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[r][0..3] += h[r] * w for the R rows of a CTA; col points at the R row values of one shared-memory operand column
template <int R>
__device__ __forceinline__ void fma_rows(float (&acc)[R][4], const float* col, const float4 w) {
  float h[R];
  if constexpr (R == 8) {
    const float4 h0 = *reinterpret_cast<const float4*>(col), h1 = *reinterpret_cast<const float4*>(col + 4);
    h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
  } else if constexpr (R == 4) {
    const float4 h0 = *reinterpret_cast<const float4*>(col);
    h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w;
  } else {
    static_assert(R == 2, "R in {8, 4, 2}");
    const float2 h0 = *reinterpret_cast<const float2*>(col);
    h[0] = h0.x; h[1] = h0.y;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    acc[r][0] = fmaf(h[r], w.x, acc[r][0]); acc[r][1] = fmaf(h[r], w.y, acc[r][1]);
    acc[r][2] = fmaf(h[r], w.z, acc[r][2]); acc[r][3] = fmaf(h[r], w.w, acc[r][3]);
  }
}

// The streaming loop all four GEMM phases share: acc[R][4] += sum_i colT[i0 + i*istep][0..R-1] (x) g[i0 + i*istep][0..3] for
// i < n.  `g` rows are `gstride` floats apart in global memory (this thread's 4 contiguous floats), `colT` is the
// [index][R rows] shared-memory operand.  U rows are requested one batch ahead of the batch being consumed, so 2U 16-byte
// loads per thread are in flight: with 256 threads that is 64 KB per SM, enough to cover the L2 latency at full rate
// (the first version waited for every batch and ran 20x slower than its FMA count).
template <int U, int R>
__device__ __forceinline__ void stream_fma(float (&acc)[R][4], const float* __restrict__ g, size_t gstride,
                                           const float* __restrict__ colT, int i0, int istep, int n, bool live) {
  if (!live) return;                     // a thread outside the operand contributes nothing (its accumulators stay 0)
  float4 cur[U], nxt[U];
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](float4 (&dst)[U], int base) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u;
      dst[u] = (i < n) ? __ldg(reinterpret_cast<const float4*>(g + static_cast<size_t>(i0 + i * istep) * gstride)) : zero;
    }
  };
  auto consume = [&](const float4 (&src)[U], int base) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u;
      if (i < n) {
        fma_rows<R>(acc, colT + static_cast<size_t>(i0 + i * istep) * R, src[u]);
      }
    }
  };
  load(cur, 0);
#pragma unroll 1
  for (int base = 0; base < n; base += 2 * U) {
    load(nxt, base + U);
    consume(cur, base);
    load(cur, base + 2 * U);
    consume(nxt, base + U);
  }
}

struct QkvK {
  const float* x;            // [B][L][C] NHWC block input
  const double* qs;          // quad statistics of x
  const float* gamma; const float* beta;
  const float* wqkv;         // [C][3C]: row = input channel, columns [q | k | v] (NIN W is [in, out])
  const float* bqkv;         // [3C]
  float* q;                  // [B][L][C]
  float* kT;                 // [B][C][L]
  float* v;                  // [B][L][C]
  int L;
};

constexpr int kQkvThreads = 384;      // 192 column quads of [q | k | v] x 2 halves of the input channels

// grid (ceil(L / R), B): thread (cq, half) owns output columns 4cq .. 4cq+3 of the 768 for the CTA's R tokens and sums the
// input channels of its half; the two halves are folded through shared memory.
template <int kR>
__global__ void __launch_bounds__(kQkvThreads)
attn_qkv_kernel(const QkvK k) {
  __shared__ __align__(16) float hT[kC][kR];        // normalised rows, transposed
  __shared__ __align__(16) float red[192][kR][4];   // partial sums of the upper channel half
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kR;
  const int L = k.L;
  // the CTA's 8 x 256 input block, requested before the statistics are assembled (two dependent round trips overlap)
  float xr[kR];
  float ga = 0.f, be = 0.f;
  if (tid < kC) {
#pragma unroll
    for (int r = 0; r < kR; ++r)
      xr[r] = (l0 + r < L) ? __ldg(k.x + (static_cast<size_t>(b) * L + l0 + r) * kC + tid) : 0.f;
    ga = __ldg(k.gamma + tid); be = __ldg(k.beta + tid);
  }
  if (tid < kGroups) {                               // 32 groups of 8 channels = 2 quads
    double su = 0.0, sq = 0.0;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int r = 0; r < kStatReplicas; ++r) {
        const double2 v = reinterpret_cast<const double2*>(qstat_slot(k.qs, b, r, kC / 4))[tid * 2 + j];
        su += v.x; sq += v.y;
      }
    const double n = static_cast<double>(L) * (kC / kGroups);
    const double mean = su / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kGnEps)));
  }
  __syncthreads();
  if (tid < kC) {
    const int g = tid / (kC / kGroups);
    const float sc = s_rstd[g] * ga;
    const float sh = fmaf(-s_mean[g], sc, be);
#pragma unroll
    for (int r = 0; r < kR; ++r) hT[tid][r] = fmaf(xr[r], sc, sh);
  }
  __syncthreads();
  const int cq = tid % 192, half = tid / 192;
  float acc[kR][4];
#pragma unroll
  for (int r = 0; r < kR; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
  stream_fma<kStreamU, kR>(acc, k.wqkv + 4 * cq, 3 * kC, &hT[0][0], half * (kC / 2), 1, kC / 2, true);
  if (half == 1) {
#pragma unroll
    for (int r = 0; r < kR; ++r) *reinterpret_cast<float4*>(&red[cq][r][0]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  if (half == 1) return;
  const float4 bias = __ldg(reinterpret_cast<const float4*>(k.bqkv) + cq);
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const float4 o = *reinterpret_cast<const float4*>(&red[cq][r][0]);
    acc[r][0] += o.x + bias.x; acc[r][1] += o.y + bias.y; acc[r][2] += o.z + bias.z; acc[r][3] += o.w + bias.w;
  }
  const int which = (4 * cq) / kC, n0 = (4 * cq) % kC;     // 0 q, 1 k, 2 v
  if (which != 1) {
    float* dst = (which == 0 ? k.q : k.v) + (static_cast<size_t>(b) * L + l0) * kC + n0;
#pragma unroll
    for (int r = 0; r < kR; ++r)
      if (l0 + r < L) *reinterpret_cast<float4*>(dst + static_cast<size_t>(r) * kC) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  } else {
    // keys transposed: kT[b][n][l0 .. l0+R-1]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* kt = k.kT + (static_cast<size_t>(b) * kC + n0 + i) * L + l0;
      if (l0 + kR <= L) {
        if constexpr (kR == 2) {
          *reinterpret_cast<float2*>(kt) = make_float2(acc[0][i], acc[1][i]);
        } else {
#pragma unroll
          for (int r4 = 0; r4 < kR; r4 += 4)
            *reinterpret_cast<float4*>(kt + r4) = make_float4(acc[r4][i], acc[r4 + 1][i], acc[r4 + 2][i], acc[r4 + 3][i]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < kR; ++r) if (l0 + r < L) kt[r] = acc[r][i];
      }
    }
  }
}

struct CoreK {
  const float* x;            // [B][L][C] block input (residual)
  const float* q; const float* kT; const float* v;
  const float* w3;           // [C][C] NIN_3.W ([in, out])
  const float* b3;           // [C]
  float* out;                // [B][L][C]
  double* qstats;            // optional quad statistics of out (zeroed buffer)
  int L, Lpad;               // Lpad: L rounded up to a multiple of 512 (keys per scores pass)
  float scale;               // C^-1/2
};

constexpr int kKeysPerPass = 512;      // 128 key quads x 2 channel halves = 256 threads

// grid (ceil(L / R), B), 256 threads.  Dynamic shared memory: S [R][Lpad] scores, Pt [Lpad][R] probabilities (transposed),
// qT / oT [C][R] query rows, later the context rows, red [4][R][C] partial sums of the thread groups.
template <int kR>
__global__ void __launch_bounds__(kThreads)
attn_core_kernel(const CoreK k) {
  extern __shared__ __align__(16) float sm[];
  float* S = sm;                                     // [kR][Lpad]
  float* Pt = S + static_cast<size_t>(kR) * k.Lpad;  // [Lpad][kR]
  float* qT = Pt + static_cast<size_t>(k.Lpad) * kR; // [kC][kR]
  float* red = qT + kC * kR;                         // [4][kR][kC]
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kR;
  const int L = k.L;
  // ---- query rows, transposed
#pragma unroll
  for (int r = 0; r < kR; ++r)
    qT[tid * kR + r] = (l0 + r < L) ? __ldg(k.q + (static_cast<size_t>(b) * L + l0 + r) * kC + tid) : 0.f;
  __syncthreads();
  // ---- scores: thread (kq, half) owns keys 4kq .. 4kq+3 of each 512-key pass for all 8 rows and sums its half of the
  //      channels; kT[c][key] is contiguous in the keys
  {
    const int kq = tid & 127, half = tid >> 7;
    const float* kTb = k.kT + static_cast<size_t>(b) * kC * L;
    for (int key0 = 0; key0 < L; key0 += kKeysPerPass) {
      const int key = key0 + 4 * kq;
      const bool live = key < L;                     // L is a multiple of 4: the whole quad is inside or outside
      float acc[kR][4];
#pragma unroll
      for (int r = 0; r < kR; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
      stream_fma<kStreamU, kR>(acc, kTb + key, L, qT, half * (kC / 2), 1, kC / 2, live);
      if (half == 1) {
#pragma unroll
        for (int r = 0; r < kR; ++r)
          *reinterpret_cast<float4*>(S + static_cast<size_t>(r) * k.Lpad + key) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
      __syncthreads();
      if (half == 0) {
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          float4* dst = reinterpret_cast<float4*>(S + static_cast<size_t>(r) * k.Lpad + key);
          const float4 o = *dst;
          *dst = live ? make_float4((acc[r][0] + o.x) * k.scale, (acc[r][1] + o.y) * k.scale, (acc[r][2] + o.z) * k.scale,
                                    (acc[r][3] + o.w) * k.scale)
                      : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
      }
    }
  }
  __syncthreads();
  // ---- softmax over the keys: warp r owns row r (layerspp.py:84)
  if (warp < kR) {
    const float* row = S + static_cast<size_t>(warp) * k.Lpad;
    float m = -INFINITY;
    for (int i = lane; i < L; i += 32) m = fmaxf(m, row[i]);
    m = warp_max(m);
    float sum = 0.f;
    for (int i = lane; i < L; i += 32) sum += expf(row[i] - m);
    sum = warp_sum(sum);
    for (int i = lane; i < L; i += 32) Pt[static_cast<size_t>(i) * kR + warp] = __fdiv_rn(expf(row[i] - m), sum);
  }
  __syncthreads();
  // ---- context rows h = P v: thread (cq, kg) owns channels 4cq .. 4cq+3 and every 4th key starting at kg
  const int cq = tid & 63, grp = tid >> 6;
  {
    float acc[kR][4];
#pragma unroll
    for (int r = 0; r < kR; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
    const int nkeys = (L - grp + 3) / 4;             // keys grp, grp + 4, ... < L
    stream_fma<kStreamU, kR>(acc, k.v + static_cast<size_t>(b) * L * kC + 4 * cq, kC, Pt, grp, 4, nkeys, true);
#pragma unroll
    for (int r = 0; r < kR; ++r)
      *reinterpret_cast<float4*>(red + (static_cast<size_t>(grp) * kR + r) * kC + 4 * cq) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  // fold the 4 key groups (fixed order) and park the context rows transposed (the query rows are no longer needed)
  float* oT = qT;
#pragma unroll
  for (int r = 0; r < kR; ++r)
    oT[tid * kR + r] = (red[(0 * kR + r) * kC + tid] + red[(1 * kR + r) * kC + tid]) + (red[(2 * kR + r) * kC + tid] + red[(3 * kR + r) * kC + tid]);
  __syncthreads();
  // ---- NIN_3: thread (nq, cg) owns output channels 4nq .. 4nq+3 and input channels 64cg .. 64cg+63
  {
    float acc[kR][4];
#pragma unroll
    for (int r = 0; r < kR; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
    stream_fma<kStreamU, kR>(acc, k.w3 + 4 * cq, kC, oT, grp * (kC / 4), 1, kC / 4, true);
    __syncthreads();                                 // everyone has read its share of red
#pragma unroll
    for (int r = 0; r < kR; ++r)
      *reinterpret_cast<float4*>(red + (static_cast<size_t>(grp) * kR + r) * kC + 4 * cq) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  // ---- residual + 1/sqrt(2) + statistics: thread owns output channel tid
  const float bias = __ldg(k.b3 + tid);
  float qs_s = 0.f, qs_q = 0.f;
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    if (l0 + r < L) {
      const float y = (red[(0 * kR + r) * kC + tid] + red[(1 * kR + r) * kC + tid]) + (red[(2 * kR + r) * kC + tid] + red[(3 * kR + r) * kC + tid]);
      const size_t oidx = (static_cast<size_t>(b) * L + l0 + r) * kC + tid;
      const float val = __fdiv_rn(__ldg(k.x + oidx) + (y + bias), kSqrt2);
      k.out[oidx] = val;
      qs_s += val; qs_q += val * val;
    }
  }
  if (k.qstats) {
    // a channel quad = 4 consecutive lanes
    qs_s += __shfl_xor_sync(0xffffffffu, qs_s, 1); qs_q += __shfl_xor_sync(0xffffffffu, qs_q, 1);
    qs_s += __shfl_xor_sync(0xffffffffu, qs_s, 2); qs_q += __shfl_xor_sync(0xffffffffu, qs_q, 2);
    if ((lane & 3) == 0) {
      double* dst = qstat_slot(k.qstats, b, blockIdx.x, kC / 4) + static_cast<size_t>(tid >> 2) * 2;
      atomicAdd(dst, static_cast<double>(qs_s));
      atomicAdd(dst + 1, static_cast<double>(qs_q));
    }
  }
}

}  // namespace

size_t attention_scratch_floats(int B, int L) { return static_cast<size_t>(3) * B * L * kC; }

int launch_attention(const AttentionArgs& a, cudaStream_t s, std::string* err) {
  if (a.C != kC) { if (err) *err = "attention: the kernels are specialised to 256 channels"; return 1; }
  if (a.L <= 0 || (a.L & 3)) { if (err) *err = "attention: token count must be a positive multiple of 4"; return 1; }
  const int Lpad = ((a.L + kKeysPerPass - 1) / kKeysPerPass) * kKeysPerPass;
  // rows per CTA of the core kernel: the most that still spreads the block over about one CTA per SM (it is bound by FMA
  // issue per CTA; measured at B = 1, L = 512: 8 rows -> 64 CTAs 59 us, 4 rows -> 128 CTAs 35.5 us; L = 32: 45 -> 15 us).
  // Every row's arithmetic is independent of the choice, so results do not depend on the batch size.  (Staggering the
  // CTAs' passes over the shared operands to spread them over the L2 slices gained 6 % of the attention time and cost
  // that property: not kept.)
  int R = kRMax;
  while (R > 2 && static_cast<long long>((a.L + R - 1) / R) * a.B < ATTN_MIN_CTAS) R >>= 1;
  const size_t smem = (static_cast<size_t>(2) * R * Lpad + static_cast<size_t>(kC) * R + static_cast<size_t>(4) * R * kC) * sizeof(float);
  if (smem > 200 * 1024) { if (err) *err = "attention: too many tokens for the shared-memory score rows"; return 1; }
  static PerDevice<bool> attr_done(false);
  bool& attr_set = attr_done.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_core_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(200 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_core_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(200 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_core_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(200 * 1024));
    if (e != cudaSuccess) { if (err) *err = std::string("attention: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_set = true;
  }
  float* q = a.scratch;
  float* kT = q + static_cast<size_t>(a.B) * a.L * kC;
  float* v = kT + static_cast<size_t>(a.B) * a.L * kC;
  // the QKV kernel streams 768 KB of weights per CTA and is bound by that, not by FMA issue: 8 rows per CTA unless the
  // block is tiny (measured at L = 512, B = 1: 8 rows 18.7 us, 4 rows 23.4 us; L = 32: 8 rows 19.4 us, 2 rows 16.3 us)
  const int Rq = (static_cast<long long>((a.L + kRMax - 1) / kRMax) * a.B >= 32) ? kRMax : 2;
  dim3 grid((a.L + R - 1) / R, a.B), gridq((a.L + Rq - 1) / Rq, a.B);
  QkvK qk{a.x, a.qs, a.gamma, a.beta, a.wqkv, a.bqkv, q, kT, v, a.L};
  CoreK ck{a.x, q, kT, v, a.w3, a.b3, a.out, a.qstats, a.L, Lpad, 1.0f / sqrtf(static_cast<float>(kC))};
  if (Rq == kRMax) launch_k(attn_qkv_kernel<8>, gridq, dim3(kQkvThreads), 0, s, qk);
  else launch_k(attn_qkv_kernel<2>, gridq, dim3(kQkvThreads), 0, s, qk);
  switch (R) {
    case 8: launch_k(attn_core_kernel<8>, grid, dim3(kThreads), smem, s, ck); break;
    case 4: launch_k(attn_core_kernel<4>, grid, dim3(kThreads), smem, s, ck); break;
    default: launch_k(attn_core_kernel<2>, grid, dim3(kThreads), smem, s, ck); break;
  }
  return 0;
}

}  // namespace flowse
