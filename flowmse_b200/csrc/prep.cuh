// Operand preparation shared by the standalone prep kernels (kernels_gn.cu) and the low-resolution conv kernel, which runs
// the plain variant as its own prologue phase (conv_gemm.cu): y = act(GN(x)) split to fp16 hi/lo, optional raw-x split for
// the 1x1 shortcut (/root/reference/flowmse/backbones/ncsnpp_utils/layerspp.py:243-258).
#pragma once
#include "flowse_internal.h"
#include "operand.cuh"

namespace flowse {
namespace {

struct PrepK {
  const float* s1; int C1;
  const float* s2; int C2;
  const double* qs1; const double* qs2;     // quad statistics of s1 / s2: [B][C/4][2]
  const float* gamma; const float* beta;
  int H, W, Ho, Wo, mode, silu, B;
  __half* outA; __half* outX; float* outF; float* outXF;
  unsigned long long* overflow;
};

__device__ __forceinline__ void fma4(float4& acc, float wgt, const float4 v) {
  acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
  acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
}

// y = x*sc + sh with sc = rstd*gamma, sh = beta - mean*sc for the 4 channels starting at c
__device__ __forceinline__ void scale_shift(const PrepK& k, const float* s_mean, const float* s_rstd, int c, int cpg,
                                            float4& sc, float4& sh) {
  const int g = c / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(k.gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(k.beta + c));
  sc.x = rstd * ga.x; sc.y = rstd * ga.y; sc.z = rstd * ga.z; sc.w = rstd * ga.w;
  sh.x = fmaf(-mean, sc.x, be.x); sh.y = fmaf(-mean, sc.y, be.y);
  sh.z = fmaf(-mean, sc.z, be.z); sh.w = fmaf(-mean, sc.w, be.w);
}

// Group mean / rstd from the quad statistics of the (virtually concatenated) sources; tid = index in the thread group
// that `sync` synchronises (the whole CTA for the standalone kernels).
template <class Sync>
__device__ __forceinline__ void load_stats_t(const PrepK& k, int b, int tid, float* s_mean, float* s_rstd, Sync sync) {
  if (tid < kGroups) {
    const int C = k.C1 + k.C2;
    const int qpg = (C / kGroups) >> 2;            // quads per group
    const int q1 = k.C1 >> 2;
    double su = 0.0, sq = 0.0;
    for (int j = 0; j < qpg; ++j) {
      const int qd = tid * qpg + j;
#pragma unroll
      for (int r = 0; r < kStatReplicas; ++r) {       // fixed order over the replicas
        const double2 v = (qd < q1)
            ? reinterpret_cast<const double2*>(qstat_slot(k.qs1, b, r, q1))[qd]
            : reinterpret_cast<const double2*>(qstat_slot(k.qs2, b, r, k.C2 >> 2))[qd - q1];
        su += v.x; sq += v.y;
      }
    }
    const double n = static_cast<double>(k.H) * k.W * (C / kGroups);
    const double mean = su / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kGnEps)));
  }
  sync();
}
__device__ __forceinline__ void load_stats(const PrepK& k, int b, float* s_mean, float* s_rstd) {
  load_stats_t(k, b, static_cast<int>(threadIdx.x), s_mean, s_rstd, [] { __syncthreads(); });
}

// Plain (no resampling): 8 channels per thread, two pixels per loop trip (4 x 16-byte loads in flight).
// gamma / beta and the first trip's activations are requested BEFORE the statistics are assembled: at the low
// resolutions the pass is one dependent-load chain (statistics -> affine parameters -> activations), and the three
// round trips overlap this way.
// Work of "block" bx of nbx for batch element b, done by a group of 256 threads (tid = 0..255) that `sync` synchronises.
// The block set walks pixel slots 0 .. npix-1; map(slot) gives the image pixel (h * W + w) or -1 for a slot to skip (the
// whole image: identity; the conv kernel's prologue: the halo region of its output tile).
template <int NPRE, class Map, class Sync>
__device__ __forceinline__ void prep_pixels_body(const PrepK& k, int tid, int bx, int nbx, int b, int nslots, Map map,
                                                 float* s_mean, float* s_rstd, Sync sync) {
  const int C = k.C1 + k.C2;
  const int c8 = C >> 3;
  const int ppi = 256 / c8;
  const int v = tid % c8;
  const int pp = tid / c8;
  const bool active = pp < ppi;
  const int c = v << 3;
  const int npix = k.H * k.W;
  const float* src; int ld;
  if (c < k.C1) { src = k.s1 + c; ld = k.C1; } else { src = k.s2 + (c - k.C1); ld = k.C2; }
  src += static_cast<size_t>(b) * npix * ld;
  const int stride = nbx * ppi;
  int p = bx * ppi + pp;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // the first NPRE pixels of this thread are requested before anything else
  int idx[NPRE];
  float4 xa[NPRE], xb[NPRE];
#pragma unroll
  for (int i = 0; i < NPRE; ++i) {
    idx[i] = (active && p + i * stride < nslots) ? map(p + i * stride) : -1;
    xa[i] = zero4; xb[i] = zero4;
  }
  float4 ga0 = zero4, ga1 = zero4, be0 = zero4, be1 = zero4;
  if (active) {
    ga0 = __ldg(reinterpret_cast<const float4*>(k.gamma + c)); ga1 = __ldg(reinterpret_cast<const float4*>(k.gamma + c + 4));
    be0 = __ldg(reinterpret_cast<const float4*>(k.beta + c)); be1 = __ldg(reinterpret_cast<const float4*>(k.beta + c + 4));
  }
#pragma unroll
  for (int i = 0; i < NPRE; ++i) {
    if (idx[i] >= 0) {
      const float* a = src + static_cast<size_t>(idx[i]) * ld;
      xa[i] = __ldg(reinterpret_cast<const float4*>(a)); xb[i] = __ldg(reinterpret_cast<const float4*>(a + 4));
    }
  }
  load_stats_t(k, b, tid, s_mean, s_rstd, sync);
  if (!active || p >= nslots) return;
  float4 sc0, sh0, sc1, sh1;
  {
    const int cpg = C / kGroups;
    const float m0 = s_mean[c / cpg], r0 = s_rstd[c / cpg], m1 = s_mean[(c + 4) / cpg], r1 = s_rstd[(c + 4) / cpg];
    sc0.x = r0 * ga0.x; sc0.y = r0 * ga0.y; sc0.z = r0 * ga0.z; sc0.w = r0 * ga0.w;
    sh0.x = fmaf(-m0, sc0.x, be0.x); sh0.y = fmaf(-m0, sc0.y, be0.y); sh0.z = fmaf(-m0, sc0.z, be0.z); sh0.w = fmaf(-m0, sc0.w, be0.w);
    sc1.x = r1 * ga1.x; sc1.y = r1 * ga1.y; sc1.z = r1 * ga1.z; sc1.w = r1 * ga1.w;
    sh1.x = fmaf(-m1, sc1.x, be1.x); sh1.y = fmaf(-m1, sc1.y, be1.y); sh1.z = fmaf(-m1, sc1.z, be1.z); sh1.w = fmaf(-m1, sc1.w, be1.w);
  }
  const size_t obase = static_cast<size_t>(b) * npix * C + c;
  const size_t plane = static_cast<size_t>(k.B) * npix * C;

  float vmax = 0.f;                       // largest operand magnitude this thread split to fp16
  auto emit = [&](int pi, const float4 a0, const float4 a1) {
    const float4 y0 = norm_act(a0, sc0, sh0, k.silu);
    const float4 y1 = norm_act(a1, sc1, sh1, k.silu);
    const size_t o = obase + static_cast<size_t>(pi) * C;
    if (k.outA) {
      uint2 h0, l0, h1, l1;
      split4(y0, h0, l0); split4(y1, h1, l1);
      vmax = amax4(y0, amax4(y1, vmax));
      *reinterpret_cast<uint4*>(k.outA + o) = pack8(h0, h1);
      *reinterpret_cast<uint4*>(k.outA + plane + o) = pack8(l0, l1);
    }
    if (k.outX) {
      uint2 h0, l0, h1, l1;
      split4(a0, h0, l0); split4(a1, h1, l1);
      vmax = amax4(a0, amax4(a1, vmax));
      *reinterpret_cast<uint4*>(k.outX + o) = pack8(h0, h1);
      *reinterpret_cast<uint4*>(k.outX + plane + o) = pack8(l0, l1);
    }
    if (k.outF) { *reinterpret_cast<float4*>(k.outF + o) = y0; *reinterpret_cast<float4*>(k.outF + o + 4) = y1; }
    if (k.outXF) { *reinterpret_cast<float4*>(k.outXF + o) = a0; *reinterpret_cast<float4*>(k.outXF + o + 4) = a1; }
  };

#pragma unroll
  for (int i = 0; i < NPRE; ++i)
    if (idx[i] >= 0) emit(idx[i], xa[i], xb[i]);
  p += NPRE * stride;
  for (; p + stride < nslots; p += 2 * stride) {           // two pixels per trip: 4 x 16-byte loads in flight
    const int ja = map(p), jb = map(p + stride);
    float4 x0 = zero4, x1 = zero4, z0 = zero4, z1 = zero4;
    if (ja >= 0) {
      const float* a = src + static_cast<size_t>(ja) * ld;
      x0 = __ldg(reinterpret_cast<const float4*>(a)); x1 = __ldg(reinterpret_cast<const float4*>(a + 4));
    }
    if (jb >= 0) {
      const float* bq = src + static_cast<size_t>(jb) * ld;
      z0 = __ldg(reinterpret_cast<const float4*>(bq)); z1 = __ldg(reinterpret_cast<const float4*>(bq + 4));
    }
    if (ja >= 0) emit(ja, x0, x1);
    if (jb >= 0) emit(jb, z0, z1);
  }
  if (p < nslots) {
    const int ja = map(p);
    if (ja >= 0) {
      const float* a = src + static_cast<size_t>(ja) * ld;
      emit(ja, __ldg(reinterpret_cast<const float4*>(a)), __ldg(reinterpret_cast<const float4*>(a + 4)));
    }
  }
  if (vmax > kHalfMax && k.overflow) atomicAdd(k.overflow, 1ull);
}

// the whole image
template <class Sync>
__device__ __forceinline__ void prep_plain_body(const PrepK& k, int tid, int bx, int nbx, int b, float* s_mean,
                                                float* s_rstd, Sync sync) {
  prep_pixels_body<2>(k, tid, bx, nbx, b, k.H * k.W, [](int p) { return p; }, s_mean, s_rstd, sync);
}

}  // namespace
}  // namespace flowse
