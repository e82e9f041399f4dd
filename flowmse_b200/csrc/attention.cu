// Single-head spatial self-attention block of NCSN++ (AttnBlockpp, /root/reference/flowmse/backbones/ncsnpp_utils/
// layerspp.py:62-91; NIN: layers.py:546-555) in TWO kernels instead of six:
//
//   attn_qkv_kernel   h = GroupNorm(x);  q, k, v = NIN_0..2(h)              (layerspp.py:76-79)
//   attn_core_kernel  w = softmax(q k^T C^-1/2);  h = w v;  out = (x + NIN_3(h)) / sqrt(2)   (layerspp.py:81-91)
//                     + quad statistics of `out` for the next GroupNorm
//
// 1.63 GFLOP per network evaluation (0.15 % of the FLOPs): exact fp32 SIMT FMAs, no tensor cores.  What the old path
// (four GEMM launches + a softmax kernel + a prep kernel, scores round-tripping through global memory) paid for was
// launches and latency; here a CTA owns R = 8 tokens (query rows) and keeps everything of those rows on chip.
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
constexpr int kR = 8;            // tokens (query rows) per CTA
constexpr int kThreads = 256;

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

// grid (ceil(L / 8), B), 256 threads: thread n owns output columns n (q), C + n (k), 2C + n (v) for the CTA's 8 tokens.
__global__ void __launch_bounds__(kThreads)
attn_qkv_kernel(const QkvK k) {
  __shared__ __align__(16) float hT[kC][kR];        // normalised rows, transposed
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kR;
  const int L = k.L;
  // the CTA's 8 x 256 input block, requested before the statistics are assembled (two dependent round trips overlap)
  float xr[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r)
    xr[r] = (l0 + r < L) ? __ldg(k.x + (static_cast<size_t>(b) * L + l0 + r) * kC + tid) : 0.f;
  const float ga = __ldg(k.gamma + tid), be = __ldg(k.beta + tid);
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
  {
    const int g = tid / (kC / kGroups);
    const float sc = s_rstd[g] * ga;
    const float sh = fmaf(-s_mean[g], sc, be);
#pragma unroll
    for (int r = 0; r < kR; ++r) hT[tid][r] = fmaf(xr[r], sc, sh);
  }
  __syncthreads();
  float aq[kR], ak[kR], av[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) { aq[r] = 0.f; ak[r] = 0.f; av[r] = 0.f; }
  const float* w = k.wqkv + tid;
  constexpr int U = 4;                                // weight rows in flight per thread
#pragma unroll 1
  for (int c0 = 0; c0 < kC; c0 += U) {
    float wq[U], wk[U], wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* row = w + static_cast<size_t>(c0 + u) * (3 * kC);
      wq[u] = __ldg(row); wk[u] = __ldg(row + kC); wv[u] = __ldg(row + 2 * kC);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 h0 = *reinterpret_cast<const float4*>(&hT[c0 + u][0]);
      const float4 h1 = *reinterpret_cast<const float4*>(&hT[c0 + u][4]);
      const float h[kR] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        aq[r] = fmaf(h[r], wq[u], aq[r]); ak[r] = fmaf(h[r], wk[u], ak[r]); av[r] = fmaf(h[r], wv[u], av[r]);
      }
    }
  }
  const float bq = __ldg(k.bqkv + tid), bk = __ldg(k.bqkv + kC + tid), bv = __ldg(k.bqkv + 2 * kC + tid);
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    if (l0 + r < L) {
      const size_t o = (static_cast<size_t>(b) * L + l0 + r) * kC + tid;
      k.q[o] = aq[r] + bq;
      k.v[o] = av[r] + bv;
    }
  }
  // keys transposed: kT[b][n][l0 .. l0+7] (32 contiguous bytes per thread)
  float* kt = k.kT + (static_cast<size_t>(b) * kC + tid) * L + l0;
  if (l0 + kR <= L && (L & 3) == 0) {
    *reinterpret_cast<float4*>(kt) = make_float4(ak[0] + bk, ak[1] + bk, ak[2] + bk, ak[3] + bk);
    *reinterpret_cast<float4*>(kt + 4) = make_float4(ak[4] + bk, ak[5] + bk, ak[6] + bk, ak[7] + bk);
  } else {
#pragma unroll
    for (int r = 0; r < kR; ++r) if (l0 + r < L) kt[r] = ak[r] + bk;
  }
}

struct CoreK {
  const float* x;            // [B][L][C] block input (residual)
  const float* q; const float* kT; const float* v;
  const float* w3;           // [C][C] NIN_3.W ([in, out])
  const float* b3;           // [C]
  float* out;                // [B][L][C]
  double* qstats;            // optional quad statistics of out (zeroed buffer)
  int L, Lpad;               // Lpad: L rounded up to a multiple of 2 * kThreads (keys per scores pass)
  float scale;               // C^-1/2
};

// grid (ceil(L / 8), B), 256 threads.  Dynamic shared memory: S [8][Lpad] scores, Pt [Lpad][8] probabilities (transposed),
// qT / oT [C][8] query rows, later the context rows.
__global__ void __launch_bounds__(kThreads)
attn_core_kernel(const CoreK k) {
  extern __shared__ __align__(16) float sm[];
  float* S = sm;                                     // [kR][Lpad]
  float* Pt = S + static_cast<size_t>(kR) * k.Lpad;  // [Lpad][kR]
  float* qT = Pt + static_cast<size_t>(k.Lpad) * kR; // [kC][kR]
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
  // ---- scores: thread owns keys (2 tid, 2 tid + 1) of each 512-key pass, all 8 rows
  const float* kTb = k.kT + static_cast<size_t>(b) * kC * L;
  for (int key0 = 0; key0 < L; key0 += 2 * kThreads) {
    const int key = key0 + 2 * tid;
    const bool inb = key < L;                        // L is even (a multiple of 4): key + 1 < L as well
    float a0[kR], a1[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) { a0[r] = 0.f; a1[r] = 0.f; }
    constexpr int U = 8;                             // key rows (channels) in flight per thread
#pragma unroll 1
    for (int c0 = 0; c0 < kC; c0 += U) {
      float2 kv[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        kv[u] = inb ? __ldg(reinterpret_cast<const float2*>(kTb + static_cast<size_t>(c0 + u) * L + key)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 q0 = *reinterpret_cast<const float4*>(qT + (c0 + u) * kR);
        const float4 q1 = *reinterpret_cast<const float4*>(qT + (c0 + u) * kR + 4);
        const float qq[kR] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int r = 0; r < kR; ++r) { a0[r] = fmaf(qq[r], kv[u].x, a0[r]); a1[r] = fmaf(qq[r], kv[u].y, a1[r]); }
      }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r)
      *reinterpret_cast<float2*>(S + static_cast<size_t>(r) * k.Lpad + key) =
          inb ? make_float2(a0[r] * k.scale, a1[r] * k.scale) : make_float2(-INFINITY, -INFINITY);
  }
  __syncthreads();
  // ---- softmax over the keys: warp r owns row r (layerspp.py:84)
  {
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
  // ---- context rows h = P v: thread owns channel tid for all 8 rows; v[l][tid] is coalesced
  float o[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) o[r] = 0.f;
  {
    const float* vb = k.v + static_cast<size_t>(b) * L * kC + tid;
    constexpr int U = 8;
    int l = 0;
#pragma unroll 1
    for (; l + U <= L; l += U) {
      float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) vv[u] = __ldg(vb + static_cast<size_t>(l + u) * kC);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 p0 = *reinterpret_cast<const float4*>(Pt + static_cast<size_t>(l + u) * kR);
        const float4 p1 = *reinterpret_cast<const float4*>(Pt + static_cast<size_t>(l + u) * kR + 4);
        const float pp[kR] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int r = 0; r < kR; ++r) o[r] = fmaf(pp[r], vv[u], o[r]);
      }
    }
    for (; l < L; ++l) {
      const float vv = __ldg(vb + static_cast<size_t>(l) * kC);
#pragma unroll
      for (int r = 0; r < kR; ++r) o[r] = fmaf(Pt[static_cast<size_t>(l) * kR + r], vv, o[r]);
    }
  }
  // park the context rows transposed (the query rows are no longer needed)
  float* oT = qT;
#pragma unroll
  for (int r = 0; r < kR; ++r) oT[tid * kR + r] = o[r];
  __syncthreads();
  // ---- NIN_3 + residual + 1/sqrt(2): thread owns output channel tid
  float y[kR];
#pragma unroll
  for (int r = 0; r < kR; ++r) y[r] = 0.f;
  {
    const float* w = k.w3 + tid;
    constexpr int U = 8;
#pragma unroll 1
    for (int c0 = 0; c0 < kC; c0 += U) {
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) wv[u] = __ldg(w + static_cast<size_t>(c0 + u) * kC);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 h0 = *reinterpret_cast<const float4*>(oT + (c0 + u) * kR);
        const float4 h1 = *reinterpret_cast<const float4*>(oT + (c0 + u) * kR + 4);
        const float h[kR] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int r = 0; r < kR; ++r) y[r] = fmaf(h[r], wv[u], y[r]);
      }
    }
  }
  const float bias = __ldg(k.b3 + tid);
  float qs_s = 0.f, qs_q = 0.f;
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    if (l0 + r < L) {
      const size_t oidx = (static_cast<size_t>(b) * L + l0 + r) * kC + tid;
      const float val = __fdiv_rn(__ldg(k.x + oidx) + (y[r] + bias), kSqrt2);
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
  if (a.L <= 0 || (a.L & 1)) { if (err) *err = "attention: token count must be a positive even number"; return 1; }
  const int Lpad = ((a.L + 2 * kThreads - 1) / (2 * kThreads)) * (2 * kThreads);
  const size_t smem = (static_cast<size_t>(2) * kR * Lpad + static_cast<size_t>(kC) * kR) * sizeof(float);
  if (smem > 200 * 1024) { if (err) *err = "attention: too many tokens for the shared-memory score rows"; return 1; }
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(200 * 1024));
    if (e != cudaSuccess) { if (err) *err = std::string("attention: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    attr_smem = 200 * 1024;
  }
  float* q = a.scratch;
  float* kT = q + static_cast<size_t>(a.B) * a.L * kC;
  float* v = kT + static_cast<size_t>(a.B) * a.L * kC;
  dim3 grid((a.L + kR - 1) / kR, a.B);
  QkvK qk{a.x, a.qs, a.gamma, a.beta, a.wqkv, a.bqkv, q, kT, v, a.L};
  launch_k(attn_qkv_kernel, grid, dim3(kThreads), 0, s, qk);
  CoreK ck{a.x, q, kT, v, a.w3, a.b3, a.out, a.qstats, a.L, Lpad, 1.0f / sqrtf(static_cast<float>(kC))};
  launch_k(attn_core_kernel, grid, dim3(kThreads), smem, s, ck);
  return 0;
}

}  // namespace flowse
