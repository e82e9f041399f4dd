// GroupNorm statistics and the fused operand-preparation kernel.
//
// Reference semantics: nn.GroupNorm(32, C, eps=1e-6) followed by SiLU, optionally followed by the StyleGAN2 FIR
// resampling of BOTH the activated tensor h and the raw input x
// (/root/reference/flowmse/backbones/ncsnpp_utils/layerspp.py:243-258; FIR closed forms: up_or_down_sampling.py:195-257,
// SURVEY.md Appendix B).  The reference runs these as 3-6 separate full-tensor passes; here ONE HBM-bound pass reads x
// once and writes the conv operands directly in the fp16 hi/lo split format the tcgen05 GEMM consumes.
#include "flowse_internal.h"

namespace flowse {

namespace {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

// exact hi/lo split: v ~= hi + lo with hi, lo fp16 (v saturated to the fp16 range first)
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const float a = clamp_h(v.x), b = clamp_h(v.y), c = clamp_h(v.z), d = clamp_h(v.w);
  const __half ha = __float2half_rn(a), hb = __float2half_rn(b), hc = __float2half_rn(c), hd = __float2half_rn(d);
  const __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
  const __half lc = __float2half_rn(c - __half2float(hc)), ld = __float2half_rn(d - __half2float(hd));
  __half2 h01 = __halves2half2(ha, hb), h23 = __halves2half2(hc, hd);
  __half2 l01 = __halves2half2(la, lb), l23 = __halves2half2(lc, ld);
  hi.x = *reinterpret_cast<uint32_t*>(&h01); hi.y = *reinterpret_cast<uint32_t*>(&h23);
  lo.x = *reinterpret_cast<uint32_t*>(&l01); lo.y = *reinterpret_cast<uint32_t*>(&l23);
}

// ---------------------------------------------------------------------------------------------
// statistics: per (batch, group) mean and 1/sqrt(var + eps).
// Deterministic: block partials (fixed-order shared-memory reduce) go to a scratch array; the LAST block to finish
// (ticket counter) sums them in block order in double and writes {mean, rstd}.  No float atomics, no memset.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ s1, int C1, const float* __restrict__ s2, int C2, int npix, int chunk,
                double* __restrict__ stats, double* __restrict__ partials, unsigned* __restrict__ counters) {
  const int C = C1 + C2;
  const int cvec = C >> 2;                 // float4 per pixel
  const int ppi = blockDim.x / cvec;       // pixels per iteration
  const int b = blockIdx.y;
  const int nblk = gridDim.x;
  __shared__ float ts[256], tq[256];
  __shared__ double fs[8][kGroups], fq[8][kGroups];
  __shared__ bool is_last;
  const int v = threadIdx.x % cvec;
  const int pp = threadIdx.x / cvec;
  float s = 0.f, q = 0.f;
  if (pp < ppi) {
    const int c = v << 2;
    const float* src;
    int ld, cc;
    if (c < C1) { src = s1; ld = C1; cc = c; } else { src = s2; ld = C2; cc = c - C1; }
    const int p0 = blockIdx.x * chunk;
    const int p1 = min(npix, p0 + chunk);
    const float* base = src + static_cast<size_t>(b) * npix * ld + cc;
    int p = p0 + pp;
    float s1a = 0.f, q1a = 0.f, s2a = 0.f, q2a = 0.f, s3a = 0.f, q3a = 0.f;
    for (; p + 3 * ppi < p1; p += 4 * ppi) {        // 4 independent 16-byte loads in flight per thread
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * ld));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + ppi) * ld));
      const float4 x2 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 2 * ppi) * ld));
      const float4 x3 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 3 * ppi) * ld));
      s += (x0.x + x0.y) + (x0.z + x0.w);   q += (x0.x * x0.x + x0.y * x0.y) + (x0.z * x0.z + x0.w * x0.w);
      s1a += (x1.x + x1.y) + (x1.z + x1.w); q1a += (x1.x * x1.x + x1.y * x1.y) + (x1.z * x1.z + x1.w * x1.w);
      s2a += (x2.x + x2.y) + (x2.z + x2.w); q2a += (x2.x * x2.x + x2.y * x2.y) + (x2.z * x2.z + x2.w * x2.w);
      s3a += (x3.x + x3.y) + (x3.z + x3.w); q3a += (x3.x * x3.x + x3.y * x3.y) + (x3.z * x3.z + x3.w * x3.w);
    }
    for (; p < p1; p += ppi) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * ld));
      s += (x.x + x.y) + (x.z + x.w);
      q += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    }
    s = (s + s1a) + (s2a + s3a);
    q = (q + q1a) + (q2a + q3a);
  }
  ts[threadIdx.x] = s; tq[threadIdx.x] = q;
  __syncthreads();
  const int vpg = cvec / kGroups;          // float4 lanes per group (1, 2, 3 or 4)
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    double as = 0.0, aq = 0.0;
    for (int r = 0; r < ppi; ++r)
      for (int j = 0; j < vpg; ++j) {
        const int t = r * cvec + g * vpg + j;
        as += static_cast<double>(ts[t]); aq += static_cast<double>(tq[t]);
      }
    double* dst = partials + ((static_cast<size_t>(b) * nblk + blockIdx.x) * kGroups + g) * 2;
    dst[0] = as; dst[1] = aq;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&counters[b], 1u) == static_cast<unsigned>(nblk - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  {
    const int g = threadIdx.x & 31, part = threadIdx.x >> 5;     // 8 parts x 32 groups
    double as = 0.0, aq = 0.0;
    for (int k = part; k < nblk; k += 8) {
      const double* src = partials + ((static_cast<size_t>(b) * nblk + k) * kGroups + g) * 2;
      as += __ldcg(src); aq += __ldcg(src + 1);
    }
    fs[part][g] = as; fq[part][g] = aq;
  }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    double as = 0.0, aq = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { as += fs[k][g]; aq += fq[k][g]; }
    const double n = static_cast<double>(npix) * (C / kGroups);
    const double mean = as / n;
    double var = aq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(static_cast<size_t>(b) * kGroups + g) * 2 + 0] = mean;
    stats[(static_cast<size_t>(b) * kGroups + g) * 2 + 1] = 1.0 / sqrt(var + static_cast<double>(kGnEps));
  }
  if (threadIdx.x == 0) counters[b] = 0u;    // ready for the next GroupNorm on this stream
}

// ---------------------------------------------------------------------------------------------
// prep: y = act(GN(x)) (+ FIR), split to fp16 hi/lo; optional raw-x split for the 1x1 shortcut
// ---------------------------------------------------------------------------------------------
struct PrepK {
  const float* s1; int C1;
  const float* s2; int C2;
  const double* stats;
  const float* gamma; const float* beta;
  int H, W, Ho, Wo, mode, silu, B;
  __half* outA; __half* outX; float* outF; float* outXF;
};

__device__ __forceinline__ float4 load_src(const PrepK& k, int b, int h, int w, int c) {
  const size_t pix = (static_cast<size_t>(b) * k.H + h) * k.W + w;
  if (c < k.C1) return __ldg(reinterpret_cast<const float4*>(k.s1 + pix * k.C1 + c));
  return __ldg(reinterpret_cast<const float4*>(k.s2 + pix * k.C2 + (c - k.C1)));
}

__device__ __forceinline__ float4 norm_act(const float4 x, const float4 sc, const float4 sh, int silu) {
  float4 y;
  y.x = fmaf(x.x, sc.x, sh.x); y.y = fmaf(x.y, sc.y, sh.y);
  y.z = fmaf(x.z, sc.z, sh.z); y.w = fmaf(x.w, sc.w, sh.w);
  if (silu) { y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w); }
  return y;
}

__device__ __forceinline__ void fma4(float4& acc, float wgt, const float4 v) {
  acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
  acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
}

__global__ void __launch_bounds__(256)
gn_prep_kernel(const PrepK k) {
  const int C = k.C1 + k.C2;
  const int cvec = C >> 2;
  const int b = blockIdx.y;
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  if (threadIdx.x < kGroups) {
    s_mean[threadIdx.x] = static_cast<float>(k.stats[(static_cast<size_t>(b) * kGroups + threadIdx.x) * 2 + 0]);
    s_rstd[threadIdx.x] = static_cast<float>(k.stats[(static_cast<size_t>(b) * kGroups + threadIdx.x) * 2 + 1]);
  }
  __syncthreads();
  const size_t total = static_cast<size_t>(k.Ho) * k.Wo * cvec;
  const size_t plane = static_cast<size_t>(k.B) * k.Ho * k.Wo * C;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = idx % cvec;
    const size_t opix = idx / cvec;
    const int wo = opix % k.Wo;
    const int ho = opix / k.Wo;
    const int c = v << 2;
    const int g = c / (C / kGroups);
    const float mean = s_mean[g], rstd = s_rstd[g];
    const float4 ga = __ldg(reinterpret_cast<const float4*>(k.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(k.beta + c));
    float4 sc, sh;   // y = x*sc + sh  with sc = rstd*gamma, sh = beta - mean*sc
    sc.x = rstd * ga.x; sc.y = rstd * ga.y; sc.z = rstd * ga.z; sc.w = rstd * ga.w;
    sh.x = fmaf(-mean, sc.x, be.x); sh.y = fmaf(-mean, sc.y, be.y);
    sh.z = fmaf(-mean, sc.z, be.z); sh.w = fmaf(-mean, sc.w, be.w);

    float4 ya, xa;
    if (k.mode == kPrepPlain) {
      xa = load_src(k, b, ho, wo, c);
      ya = norm_act(xa, sc, sh, k.silu);
    } else if (k.mode == kPrepDown) {
      // out[m] = (x[2m-1] + 3x[2m] + 3x[2m+1] + x[2m+2]) / 8 per axis; 2-D taps outer([1,3,3,1])/64
      ya = make_float4(0.f, 0.f, 0.f, 0.f); xa = ya;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int hi = 2 * ho - 1 + i;
        if (hi < 0 || hi >= k.H) continue;
        const float wi_ = (i == 0 || i == 3) ? 1.f : 3.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int wj = 2 * wo - 1 + j;
          if (wj < 0 || wj >= k.W) continue;
          const float wgt = wi_ * ((j == 0 || j == 3) ? 1.f : 3.f) * (1.f / 64.f);
          const float4 x = load_src(k, b, hi, wj, c);
          fma4(xa, wgt, x);
          fma4(ya, wgt, norm_act(x, sc, sh, k.silu));
        }
      }
    } else {
      // up x2: out[2m] = .25 x[m-1] + .75 x[m];  out[2m+1] = .75 x[m] + .25 x[m+1]; 2-D taps {1,3,3,9}/16
      ya = make_float4(0.f, 0.f, 0.f, 0.f); xa = ya;
      const int mh = ho >> 1, mw = wo >> 1;
      const int h_a = (ho & 1) ? mh : mh - 1, h_b = (ho & 1) ? mh + 1 : mh;      // weights: a,b
      const float wha = (ho & 1) ? 3.f : 1.f, whb = (ho & 1) ? 1.f : 3.f;
      const int w_a = (wo & 1) ? mw : mw - 1, w_b = (wo & 1) ? mw + 1 : mw;
      const float wwa = (wo & 1) ? 3.f : 1.f, wwb = (wo & 1) ? 1.f : 3.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int hi = i ? h_b : h_a;
        if (hi < 0 || hi >= k.H) continue;
        const float wi_ = i ? whb : wha;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int wj = j ? w_b : w_a;
          if (wj < 0 || wj >= k.W) continue;
          const float wgt = wi_ * (j ? wwb : wwa) * (1.f / 16.f);
          const float4 x = load_src(k, b, hi, wj, c);
          fma4(xa, wgt, x);
          fma4(ya, wgt, norm_act(x, sc, sh, k.silu));
        }
      }
    }
    const size_t o = (static_cast<size_t>(b) * k.Ho * k.Wo + opix) * C + c;
    if (k.outA) {
      uint2 hi, lo;
      split4(ya, hi, lo);
      *reinterpret_cast<uint2*>(k.outA + o) = hi;
      *reinterpret_cast<uint2*>(k.outA + plane + o) = lo;
    }
    if (k.outX) {
      uint2 hi, lo;
      split4(xa, hi, lo);
      *reinterpret_cast<uint2*>(k.outX + o) = hi;
      *reinterpret_cast<uint2*>(k.outX + plane + o) = lo;
    }
    if (k.outF) *reinterpret_cast<float4*>(k.outF + o) = ya;
    if (k.outXF) *reinterpret_cast<float4*>(k.outXF + o) = xa;
  }
}

// Plain (no resampling) variant: 8 channels per thread -> two 16-byte loads in flight and 16-byte fp16 stores.
__device__ __forceinline__ uint4 pack8(const uint2 a, const uint2 b) { return make_uint4(a.x, a.y, b.x, b.y); }

__global__ void __launch_bounds__(256)
gn_prep_plain8_kernel(const PrepK k) {
  const int C = k.C1 + k.C2;
  const int c8 = C >> 3;
  const int cpg = C / kGroups;
  const int b = blockIdx.y;
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  if (threadIdx.x < kGroups) {
    s_mean[threadIdx.x] = static_cast<float>(k.stats[(static_cast<size_t>(b) * kGroups + threadIdx.x) * 2 + 0]);
    s_rstd[threadIdx.x] = static_cast<float>(k.stats[(static_cast<size_t>(b) * kGroups + threadIdx.x) * 2 + 1]);
  }
  __syncthreads();
  const size_t npix = static_cast<size_t>(k.H) * k.W;
  const size_t total = npix * c8;
  const size_t plane = static_cast<size_t>(k.B) * npix * C;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = idx % c8;
    const size_t pix = static_cast<size_t>(b) * npix + idx / c8;
    const int c = v << 3;
    const float* src = (c < k.C1) ? k.s1 + pix * k.C1 + c : k.s2 + pix * k.C2 + (c - k.C1);
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(src + 4));
    const float4 ga0 = __ldg(reinterpret_cast<const float4*>(k.gamma + c));
    const float4 ga1 = __ldg(reinterpret_cast<const float4*>(k.gamma + c + 4));
    const float4 be0 = __ldg(reinterpret_cast<const float4*>(k.beta + c));
    const float4 be1 = __ldg(reinterpret_cast<const float4*>(k.beta + c + 4));
    const int g0 = c / cpg, g1 = (c + 4) / cpg;
    const float m0 = s_mean[g0], r0 = s_rstd[g0], m1 = s_mean[g1], r1 = s_rstd[g1];
    float4 sc0, sh0, sc1, sh1;
    sc0.x = r0 * ga0.x; sc0.y = r0 * ga0.y; sc0.z = r0 * ga0.z; sc0.w = r0 * ga0.w;
    sc1.x = r1 * ga1.x; sc1.y = r1 * ga1.y; sc1.z = r1 * ga1.z; sc1.w = r1 * ga1.w;
    sh0.x = fmaf(-m0, sc0.x, be0.x); sh0.y = fmaf(-m0, sc0.y, be0.y); sh0.z = fmaf(-m0, sc0.z, be0.z); sh0.w = fmaf(-m0, sc0.w, be0.w);
    sh1.x = fmaf(-m1, sc1.x, be1.x); sh1.y = fmaf(-m1, sc1.y, be1.y); sh1.z = fmaf(-m1, sc1.z, be1.z); sh1.w = fmaf(-m1, sc1.w, be1.w);
    const float4 y0 = norm_act(x0, sc0, sh0, k.silu);
    const float4 y1 = norm_act(x1, sc1, sh1, k.silu);
    const size_t o = pix * C + c;
    if (k.outA) {
      uint2 h0, l0, h1, l1;
      split4(y0, h0, l0); split4(y1, h1, l1);
      *reinterpret_cast<uint4*>(k.outA + o) = pack8(h0, h1);
      *reinterpret_cast<uint4*>(k.outA + plane + o) = pack8(l0, l1);
    }
    if (k.outX) {
      uint2 h0, l0, h1, l1;
      split4(x0, h0, l0); split4(x1, h1, l1);
      *reinterpret_cast<uint4*>(k.outX + o) = pack8(h0, h1);
      *reinterpret_cast<uint4*>(k.outX + plane + o) = pack8(l0, l1);
    }
    if (k.outF) { *reinterpret_cast<float4*>(k.outF + o) = y0; *reinterpret_cast<float4*>(k.outF + o + 4) = y1; }
    if (k.outXF) { *reinterpret_cast<float4*>(k.outXF + o) = x0; *reinterpret_cast<float4*>(k.outXF + o + 4) = x1; }
  }
}

}  // namespace

int gn_stats_max_blocks() { return 296; }

void launch_gn_stats(const float* src1, int C1, const float* src2, int C2, int B, int npix, double* stats,
                     double* partials, unsigned* counters, cudaStream_t s) {
  const int C = C1 + (src2 ? C2 : 0);
  const int cvec = C / 4;
  const int ppi = 256 / cvec;
  int target_blocks = std::max(1, std::min(gn_stats_max_blocks(), 592 / B));
  int chunk = (npix + target_blocks - 1) / target_blocks;
  chunk = ((chunk + ppi - 1) / ppi) * ppi;
  if (chunk < ppi * 4) chunk = ppi * 4;
  dim3 grid((npix + chunk - 1) / chunk, B);
  gn_stats_kernel<<<grid, 256, 0, s>>>(src1, C1, src2, src2 ? C2 : 0, npix, chunk, stats, partials, counters);
}

void launch_gn_prep(const PrepArgs& a, cudaStream_t s) {
  PrepK k;
  k.s1 = a.src1; k.C1 = a.C1; k.s2 = a.src2; k.C2 = a.src2 ? a.C2 : 0;
  k.stats = a.stats; k.gamma = a.gamma; k.beta = a.beta;
  k.H = a.H; k.W = a.W; k.mode = a.mode; k.silu = a.silu; k.B = a.B;
  k.Ho = a.mode == kPrepDown ? a.H / 2 : (a.mode == kPrepUp ? a.H * 2 : a.H);
  k.Wo = a.mode == kPrepDown ? a.W / 2 : (a.mode == kPrepUp ? a.W * 2 : a.W);
  k.outA = a.outA; k.outX = a.outX; k.outF = a.outF; k.outXF = a.outXF;
  const bool plain8 = (a.mode == kPrepPlain) && (k.C1 % 8 == 0) && (k.C2 % 8 == 0);
  const size_t total = static_cast<size_t>(k.Ho) * k.Wo * ((k.C1 + k.C2) / (plain8 ? 8 : 4));
  size_t blocks = (total + 255) / 256;
  const size_t cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  dim3 grid(static_cast<unsigned>(blocks), a.B);
  if (plain8) gn_prep_plain8_kernel<<<grid, 256, 0, s>>>(k);
  else gn_prep_kernel<<<grid, 256, 0, s>>>(k);
}

}  // namespace flowse
