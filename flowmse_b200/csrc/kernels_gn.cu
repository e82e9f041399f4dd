// GroupNorm statistics and the fused operand-preparation kernels.
//
// Reference semantics: nn.GroupNorm(32, C, eps=1e-6) followed by SiLU, optionally followed by the StyleGAN2 FIR
// resampling of BOTH the activated tensor h and the raw input x
// (/root/reference/flowmse/backbones/ncsnpp_utils/layerspp.py:243-258; FIR closed forms: up_or_down_sampling.py:195-257,
// SURVEY.md Appendix B).  The reference runs these as 3-6 separate full-tensor passes; here ONE HBM-bound pass reads x
// once and writes the conv operands directly in the fp16 hi/lo split format the tcgen05 GEMM consumes.
//
// Thread mapping (all kernels): a thread owns a fixed channel chunk (4 or 8 channels) for the whole kernel, so the
// per-channel scale/shift (rstd*gamma, beta - mean*rstd*gamma) live in registers and the pixel loop contains no
// integer division; consecutive threads cover consecutive channels of one pixel, i.e. fully coalesced rows.
#include "flowse_internal.h"
#include "operand.cuh"
#include "prep.cuh"

namespace flowse {

namespace {

// ---------------------------------------------------------------------------------------------
// "Quad statistics": per (batch, 4-channel quad) sum and sum of squares in double.  Any GroupNorm grouping whose group
// size is a multiple of 4 channels - including groups that straddle a channel concatenation - is assembled from them
// in the consumer's prologue.  Producers (conv epilogue, split-K reduce, conv_in, combine) accumulate these with
// fp64 atomics while they still hold the values, so the tensor is not re-read.  This standalone kernel serves the
// few tensors without a fusing producer (attention outputs, op-level API): block partials + last-block finalize,
// fixed summation order, overwrites its output (no zeroing needed).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quad_stats_kernel(const float* __restrict__ src, int C, int npix, int chunk, double* __restrict__ qs,
                  double* __restrict__ partials, unsigned* __restrict__ counters) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvec = C >> 2;                 // quads per pixel
  const int ppi = blockDim.x / cvec;       // pixels per iteration
  const int b = blockIdx.y;
  const int nblk = gridDim.x;
  __shared__ float ts[256], tq[256];
  __shared__ bool is_last;
  const int v = threadIdx.x % cvec;
  const int pp = threadIdx.x / cvec;
  float s = 0.f, q = 0.f;
  if (pp < ppi) {
    const int p0 = blockIdx.x * chunk;
    const int p1 = min(npix, p0 + chunk);
    const float* base = src + static_cast<size_t>(b) * npix * C + (v << 2);
    int p = p0 + pp;
    float s1a = 0.f, q1a = 0.f, s2a = 0.f, q2a = 0.f, s3a = 0.f, q3a = 0.f;
    for (; p + 3 * ppi < p1; p += 4 * ppi) {        // 4 independent 16-byte loads in flight per thread
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * C));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + ppi) * C));
      const float4 x2 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 2 * ppi) * C));
      const float4 x3 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 3 * ppi) * C));
      s += (x0.x + x0.y) + (x0.z + x0.w);   q += (x0.x * x0.x + x0.y * x0.y) + (x0.z * x0.z + x0.w * x0.w);
      s1a += (x1.x + x1.y) + (x1.z + x1.w); q1a += (x1.x * x1.x + x1.y * x1.y) + (x1.z * x1.z + x1.w * x1.w);
      s2a += (x2.x + x2.y) + (x2.z + x2.w); q2a += (x2.x * x2.x + x2.y * x2.y) + (x2.z * x2.z + x2.w * x2.w);
      s3a += (x3.x + x3.y) + (x3.z + x3.w); q3a += (x3.x * x3.x + x3.y * x3.y) + (x3.z * x3.z + x3.w * x3.w);
    }
    for (; p < p1; p += ppi) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * C));
      s += (x.x + x.y) + (x.z + x.w);
      q += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    }
    s = (s + s1a) + (s2a + s3a);
    q = (q + q1a) + (q2a + q3a);
  }
  ts[threadIdx.x] = s; tq[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < cvec) {                // one thread per quad sums the ppi pixel rows in fixed order
    double as = 0.0, aq = 0.0;
    for (int r = 0; r < ppi; ++r) {
      as += static_cast<double>(ts[r * cvec + threadIdx.x]); aq += static_cast<double>(tq[r * cvec + threadIdx.x]);
    }
    reinterpret_cast<double2*>(partials)[(static_cast<size_t>(b) * nblk + blockIdx.x) * cvec + threadIdx.x] =
        make_double2(as, aq);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&counters[b], 1u) == static_cast<unsigned>(nblk - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < cvec) {
    const double2* src2 = reinterpret_cast<const double2*>(partials) + static_cast<size_t>(b) * nblk * cvec + threadIdx.x;
    double as = 0.0, aq = 0.0, bs = 0.0, bq = 0.0;
    int k = 0;
    for (; k + 3 < nblk; k += 4) {        // 4 coalesced loads in flight, fixed order
      const double2 v0 = __ldcg(src2 + static_cast<size_t>(k) * cvec);
      const double2 v1 = __ldcg(src2 + static_cast<size_t>(k + 1) * cvec);
      const double2 v2 = __ldcg(src2 + static_cast<size_t>(k + 2) * cvec);
      const double2 v3 = __ldcg(src2 + static_cast<size_t>(k + 3) * cvec);
      as += v0.x; aq += v0.y; bs += v1.x; bq += v1.y; as += v2.x; aq += v2.y; bs += v3.x; bq += v3.y;
    }
    for (; k < nblk; ++k) { const double2 v0 = __ldcg(src2 + static_cast<size_t>(k) * cvec); as += v0.x; aq += v0.y; }
    reinterpret_cast<double2*>(qstat_slot(qs, b, 0, cvec))[threadIdx.x] = make_double2(as + bs, aq + bq);
  }
  if (threadIdx.x == 0) counters[b] = 0u;    // ready for the next call on this stream
}

// Plain (no resampling) preparation as a kernel of its own (prep.cuh holds the body).
__global__ void __launch_bounds__(256)
gn_prep_plain_kernel(const PrepK k) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  prep_plain_body(k, static_cast<int>(threadIdx.x), static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x),
                  static_cast<int>(blockIdx.y), s_mean, s_rstd, [] { __syncthreads(); });
}

// FIR down / up x2 fused with the normalisation (single-source inputs only).
// A CTA owns an output tile x CC channels: the input region it needs (tile + FIR halo) is read ONCE, normalised and
// activated once per input element, and parked in shared memory as fp32 (activated and raw); the 4x4 (down) / 2x2 (up)
// taps are then applied from shared memory.  (The first version recomputed GN+SiLU per tap - 16x / 4x per output -
// and was MUFU-bound at 3x the HBM time.)  The tap order and fma chain are those of the reference's 2-D kernel
// outer([1,3,3,1]) (/root/reference/flowmse/backbones/ncsnpp_utils/up_or_down_sampling.py:181-257).
template <int CC, int TOH, int TOW, bool DOWN>
__global__ void __launch_bounds__(256)
gn_prep_resample_kernel(const PrepK k) {
  constexpr int L = CC / 4;                                   // float4 lanes per pixel
  constexpr int RH = DOWN ? 2 * TOH + 2 : TOH / 2 + 2;        // input region (with halo)
  constexpr int RW = DOWN ? 2 * TOW + 2 : TOW / 2 + 2;
  __shared__ float4 s_act[RH * RW * L];
  __shared__ float4 s_raw[RH * RW * L];
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  pdl_launch_dependents();
  pdl_wait();
  const int C = k.C1;
  const int nchunk = C / CC;
  const int b = blockIdx.z / nchunk;
  const int c0 = (blockIdx.z - b * nchunk) * CC;
  load_stats(k, b, s_mean, s_rstd);
  const int tid = threadIdx.x;
  const int lane = tid % L;
  const int c = c0 + lane * 4;
  float4 sc, sh;
  scale_shift(k, s_mean, s_rstd, c, C / kGroups, sc, sh);
  const int oh0 = blockIdx.y * TOH, ow0 = blockIdx.x * TOW;
  const int ih0 = DOWN ? 2 * oh0 - 1 : oh0 / 2 - 1;           // input coordinates of region element (0, 0)
  const int iw0 = DOWN ? 2 * ow0 - 1 : ow0 / 2 - 1;
  const float* src = k.s1 + static_cast<size_t>(b) * k.H * k.W * C + c;
  // all of this thread's loads are issued before the first one is consumed (the kernel is latency-bound otherwise)
  constexpr int PPI = 256 / L;                                // pixels per pass
  constexpr int NPASS = (RH * RW + PPI - 1) / PPI;
  float4 xr[NPASS];
  bool inb[NPASS];
#pragma unroll
  for (int it = 0; it < NPASS; ++it) {
    const int px = tid / L + it * PPI;
    const int r = px / RW, q = px - r * RW;
    const int h = ih0 + r, w = iw0 + q;
    inb[it] = px < RH * RW && h >= 0 && h < k.H && w >= 0 && w < k.W;
    xr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inb[it]) xr[it] = __ldg(reinterpret_cast<const float4*>(src + (static_cast<size_t>(h) * k.W + w) * C));
  }
#pragma unroll
  for (int it = 0; it < NPASS; ++it) {
    const int px = tid / L + it * PPI;
    if (px < RH * RW) {
      // upfirdn2d pads the ACTIVATED tensor with zeros
      s_act[px * L + lane] = inb[it] ? norm_act(xr[it], sc, sh, k.silu) : make_float4(0.f, 0.f, 0.f, 0.f);
      s_raw[px * L + lane] = xr[it];
    }
  }
  __syncthreads();
  const size_t plane = static_cast<size_t>(k.B) * k.Ho * k.Wo * C;
  float vmax = 0.f;
  for (int op = tid / L; op < TOH * TOW; op += 256 / L) {
    const int oy = op / TOW, ox = op - oy * TOW;
    const int ho = oh0 + oy, wo = ow0 + ox;
    if (ho >= k.Ho || wo >= k.Wo) continue;
    float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), xa = ya;
    if (DOWN) {
      // out[m] = (x[2m-1] + 3x[2m] + 3x[2m+1] + x[2m+2]) / 8 per axis; 2-D taps outer([1,3,3,1]) / 64
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float wi_ = (i == 0 || i == 3) ? 1.f : 3.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wgt = wi_ * ((j == 0 || j == 3) ? 1.f : 3.f) * (1.f / 64.f);
          const int e = ((2 * oy + i) * RW + 2 * ox + j) * L + lane;
          fma4(xa, wgt, s_raw[e]);
          fma4(ya, wgt, s_act[e]);
        }
      }
    } else {
      // up x2: out[2m] = .25 x[m-1] + .75 x[m];  out[2m+1] = .75 x[m] + .25 x[m+1]; 2-D taps {1,3,3,9} / 16
      const int ra = (oy >> 1) + (ho & 1), qa = (ox >> 1) + (wo & 1);   // region row / col of the first tap
      const float wha = (ho & 1) ? 3.f : 1.f, whb = (ho & 1) ? 1.f : 3.f;
      const float wwa = (wo & 1) ? 3.f : 1.f, wwb = (wo & 1) ? 1.f : 3.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float wi_ = i ? whb : wha;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float wgt = wi_ * (j ? wwb : wwa) * (1.f / 16.f);
          const int e = ((ra + i) * RW + qa + j) * L + lane;
          fma4(xa, wgt, s_raw[e]);
          fma4(ya, wgt, s_act[e]);
        }
      }
    }
    const size_t o = (static_cast<size_t>(b) * k.Ho * k.Wo + static_cast<size_t>(ho) * k.Wo + wo) * C + c;
    if (k.outA) {
      uint2 hi, lo;
      split4(ya, hi, lo);
      vmax = amax4(ya, vmax);
      *reinterpret_cast<uint2*>(k.outA + o) = hi;
      *reinterpret_cast<uint2*>(k.outA + plane + o) = lo;
    }
    if (k.outX) {
      uint2 hi, lo;
      split4(xa, hi, lo);
      vmax = amax4(xa, vmax);
      *reinterpret_cast<uint2*>(k.outX + o) = hi;
      *reinterpret_cast<uint2*>(k.outX + plane + o) = lo;
    }
    if (k.outF) *reinterpret_cast<float4*>(k.outF + o) = ya;
    if (k.outXF) *reinterpret_cast<float4*>(k.outXF + o) = xa;
  }
  if (vmax > kHalfMax && k.overflow) atomicAdd(k.overflow, 1ull);
}


// ---------------------------------------------------------------------------------------------
// Pyramid head, fully fused: pyr = FIR-up(prev) + conv3x3(C -> 4)(SiLU(GN(h))) + bias
// (/root/reference/flowmse/backbones/ncsnpp.py:347-366; FIR-up closed form of SURVEY.md Appendix B).
// Cout = 4 is the wrong shape for the tensor cores (a 128 x 16 MMA tile is 3/4 padding and still needs the
// fp16 hi/lo operand pass over h); here ONE kernel reads h (fp32) once, normalises + activates it into a shared-memory
// halo tile, and evaluates the 4 outputs with exact fp32 FMAs.  Output tile (RG*P) x TW pixels; a thread owns P vertically
// adjacent pixels x 4 outputs and, per (dx, channel quad), loads the P + 2 activation rows it needs once for all 3 dy
// taps; the channels of a chunk are split over KS thread groups whose partial sums are folded through shared memory in
// a fixed order.  Two shapes: 16 x 16 tiles with P = 4 for the full-resolution level (FMA-bound, 1.27x halo, 512 CTAs at
// B = 1), 4 x 8 tiles with P = 1 and an 8-way channel split for every smaller level (latency-bound: many short CTAs).
// ---------------------------------------------------------------------------------------------
struct HeadK {
  const float* h; const double* qs; const float* gamma; const float* beta;
  const float* wf;        // [9][C][4] fp32: tap-major, the 4 outputs of one input channel contiguous
  const float* bias;      // [4]
  const float4* prev;     // [B][H/2][W/2] previous (coarser) pyramid level or null
  float4* out;            // [B][H][W]
  int B, H, W, C;
};

template <int TW, int RG, int P, int KS, int CC>
__global__ void __launch_bounds__(TW * RG * KS)
head_conv_kernel(const HeadK k) {
  constexpr int TH = RG * P;
  constexpr int NT = TW * RG * KS;
  constexpr int RW = TW + 2, RH = TH + 2;
  constexpr int PS = CC + 4;                       // floats per staged pixel (+4: conflict-free float4 rows)
  constexpr int CPK = CC / KS;                     // channels of a chunk per thread group
  static_assert(CPK % 4 == 0 && CPK >= 4, "channel split");
  extern __shared__ __align__(16) float sm[];
  constexpr int TILE_F = (RH * RW * PS > KS * TH * TW * 4) ? RH * RW * PS : KS * TH * TW * 4;
  float* s_act = sm;                               // [RH*RW][PS]; reused for the [KS][TH*TW] float4 partial sums
  float* s_w = s_act + TILE_F;                     // [9][CC][4]
  float* s_sc = s_w + 9 * CC * 4;                  // [C] scale, [C] shift
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int C = k.C;
  const int tiles_w = (k.W + TW - 1) / TW;
  const int b = blockIdx.y;
  const int h0 = (blockIdx.x / tiles_w) * TH, w0 = (blockIdx.x % tiles_w) * TW;
  {
    PrepK pk{};
    pk.C1 = C; pk.C2 = 0; pk.qs1 = k.qs; pk.qs2 = nullptr; pk.H = k.H; pk.W = k.W;
    load_stats(pk, b, s_mean, s_rstd);
    pk.gamma = k.gamma; pk.beta = k.beta;
    float* s_sh = s_sc + C;
    for (int c = tid * 4; c < C; c += NT * 4) {
      float4 sc, sh;
      scale_shift(pk, s_mean, s_rstd, c, C / kGroups, sc, sh);
      *reinterpret_cast<float4*>(s_sc + c) = sc;
      *reinterpret_cast<float4*>(s_sh + c) = sh;
    }
  }
  const float* s_sh = s_sc + C;
  const int ks = tid / (TW * RG);
  const int rgp = (tid / TW) % RG;
  const int col = tid % TW;
  float acc[P][4];
#pragma unroll
  for (int i = 0; i < P; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
  const float* src = k.h + static_cast<size_t>(b) * k.H * k.W * C;

  for (int c0 = 0; c0 < C; c0 += CC) {
    __syncthreads();                               // previous chunk fully consumed (and s_sc complete on the first trip)
    // all global loads of this chunk (activations and weights) are issued before the first one is consumed: the small
    // levels are pure latency chains, one round trip per chunk instead of one per element
    constexpr int L = CC / 4;
    constexpr int NPA = (RH * RW * L + NT - 1) / NT;
    constexpr int NPW = (9 * CC + NT - 1) / NT;
    float4 xr[NPA], wr[NPW];
    bool inb[NPA];
#pragma unroll
    for (int it = 0; it < NPA; ++it) {
      const int idx = tid + it * NT;
      const int px = idx / L, c4 = idx - px * L;
      const int r = px / RW, q = px - r * RW;
      const int hh = h0 - 1 + r, ww = w0 - 1 + q;
      inb[it] = idx < RH * RW * L && hh >= 0 && hh < k.H && ww >= 0 && ww < k.W;
      xr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb[it])
        xr[it] = __ldg(reinterpret_cast<const float4*>(src + (static_cast<size_t>(hh) * k.W + ww) * C + c0 + c4 * 4));
    }
#pragma unroll
    for (int it = 0; it < NPW; ++it) {
      const int idx = tid + it * NT;
      wr[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < 9 * CC) {
        const int tap = idx / CC, c = idx - tap * CC;
        wr[it] = __ldg(reinterpret_cast<const float4*>(k.wf + (static_cast<size_t>(tap) * C + c0 + c) * 4));
      }
    }
#pragma unroll
    for (int it = 0; it < NPA; ++it) {
      const int idx = tid + it * NT;
      if (idx < RH * RW * L) {
        const int px = idx / L, c4 = idx - px * L;
        const int c = c0 + c4 * 4;
        // the conv pads the ACTIVATED tensor with zeros
        const float4 y = inb[it] ? norm_act(xr[it], *reinterpret_cast<const float4*>(s_sc + c),
                                            *reinterpret_cast<const float4*>(s_sh + c), 1)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(s_act + px * PS + c4 * 4) = y;
      }
    }
#pragma unroll
    for (int it = 0; it < NPW; ++it) {
      const int idx = tid + it * NT;
      if (idx < 9 * CC) *reinterpret_cast<float4*>(s_w + idx * 4) = wr[it];
    }
    __syncthreads();
#pragma unroll 1
    for (int cq = 0; cq < CPK / 4; ++cq) {
      const int cc = ks * CPK + cq * 4;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float4 a[P + 2];
#pragma unroll
        for (int r = 0; r < P + 2; ++r)
          a[r] = *reinterpret_cast<const float4*>(s_act + ((rgp * P + r) * RW + col + dx) * PS + cc);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float* wp = s_w + ((dy * 3 + dx) * CC + cc) * 4;
          const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
          const float4 wc = *reinterpret_cast<const float4*>(wp + 8), wd = *reinterpret_cast<const float4*>(wp + 12);
#pragma unroll
          for (int i = 0; i < P; ++i) {
            const float4 v = a[i + dy];
            acc[i][0] = fmaf(v.x, wa.x, acc[i][0]); acc[i][1] = fmaf(v.x, wa.y, acc[i][1]);
            acc[i][2] = fmaf(v.x, wa.z, acc[i][2]); acc[i][3] = fmaf(v.x, wa.w, acc[i][3]);
            acc[i][0] = fmaf(v.y, wb.x, acc[i][0]); acc[i][1] = fmaf(v.y, wb.y, acc[i][1]);
            acc[i][2] = fmaf(v.y, wb.z, acc[i][2]); acc[i][3] = fmaf(v.y, wb.w, acc[i][3]);
            acc[i][0] = fmaf(v.z, wc.x, acc[i][0]); acc[i][1] = fmaf(v.z, wc.y, acc[i][1]);
            acc[i][2] = fmaf(v.z, wc.z, acc[i][2]); acc[i][3] = fmaf(v.z, wc.w, acc[i][3]);
            acc[i][0] = fmaf(v.w, wd.x, acc[i][0]); acc[i][1] = fmaf(v.w, wd.y, acc[i][1]);
            acc[i][2] = fmaf(v.w, wd.z, acc[i][2]); acc[i][3] = fmaf(v.w, wd.w, acc[i][3]);
          }
        }
      }
    }
  }
  // fold the KS channel groups in fixed order through shared memory (the activation tile is no longer needed)
  __syncthreads();
  float4* s_part = reinterpret_cast<float4*>(sm);  // [KS][TH*TW]
#pragma unroll
  for (int i = 0; i < P; ++i)
    s_part[ks * (TH * TW) + (rgp * P + i) * TW + col] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  const float4 bv = __ldg(reinterpret_cast<const float4*>(k.bias));
  for (int px = tid; px < TH * TW; px += NT) {
    const int h = h0 + px / TW, w = w0 + px % TW;
    if (h >= k.H || w >= k.W) continue;
    float4 v = s_part[px];
#pragma unroll
    for (int g = 1; g < KS; ++g) {
      const float4 u = s_part[g * (TH * TW) + px];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    if (k.prev) {
      // StyleGAN2 FIR upsample x2 of the coarser pyramid: 2-D taps {1,3,3,9}/16, zero boundary (up_or_down_sampling.py:195-224)
      const int Hp = k.H >> 1, Wp = k.W >> 1;
      const int mh = h >> 1, mw = w >> 1;
      const int h_a = (h & 1) ? mh : mh - 1, h_b = (h & 1) ? mh + 1 : mh;
      const float wha = (h & 1) ? 3.f : 1.f, whb = (h & 1) ? 1.f : 3.f;
      const int w_a = (w & 1) ? mw : mw - 1, w_b = (w & 1) ? mw + 1 : mw;
      const float wwa = (w & 1) ? 3.f : 1.f, wwb = (w & 1) ? 1.f : 3.f;
      float4 up = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        const int hi = a2 ? h_b : h_a;
        if (hi < 0 || hi >= Hp) continue;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int wi = c2 ? w_b : w_a;
          if (wi < 0 || wi >= Wp) continue;
          const float wgt = (a2 ? whb : wha) * (c2 ? wwb : wwa) * (1.f / 16.f);
          fma4(up, wgt, __ldg(k.prev + (static_cast<size_t>(b) * Hp + hi) * Wp + wi));
        }
      }
      v.x = up.x + v.x; v.y = up.y + v.y; v.z = up.z + v.z; v.w = up.w + v.w;
    }
    k.out[(static_cast<size_t>(b) * k.H + h) * k.W + w] = v;
  }
}

template <int TW, int RG, int P, int KS, int CC>
void launch_head_t(const HeadK& k, cudaStream_t s) {
  constexpr int TH = RG * P, RW = TW + 2, RH = TH + 2;
  const size_t tile = std::max(static_cast<size_t>(RH) * RW * (CC + 4), static_cast<size_t>(KS) * TH * TW * 4);
  const size_t smem = (tile + 9 * CC * 4 + 2 * k.C) * sizeof(float);
  static PerDevice<bool> attr_done(false);
  bool& attr_set = attr_done.get();
  if (!attr_set) {
    cudaFuncSetAttribute(head_conv_kernel<TW, RG, P, KS, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  dim3 grid(((k.W + TW - 1) / TW) * ((k.H + TH - 1) / TH), k.B);
  launch_k(head_conv_kernel<TW, RG, P, KS, CC>, grid, dim3(TW * RG * KS), smem, s, k);
}

}  // namespace

int gn_stats_max_blocks() { return 296; }

void launch_quad_stats(const float* src, int C, int B, int npix, double* qs, double* partials, unsigned* counters,
                       cudaStream_t s) {
  const int cvec = C / 4;
  const int ppi = 256 / cvec;
  int target_blocks = std::max(1, std::min(gn_stats_max_blocks(), 592 / B));
  int chunk = (npix + target_blocks - 1) / target_blocks;
  chunk = ((chunk + ppi - 1) / ppi) * ppi;
  if (chunk < ppi * 4) chunk = ppi * 4;
  dim3 grid((npix + chunk - 1) / chunk, B);
  launch_k(quad_stats_kernel, grid, dim3(256), 0, s, src, C, npix, chunk, qs, partials, counters);
}

void launch_gn_prep(const PrepArgs& a, cudaStream_t s) {
  PrepK k;
  k.s1 = a.src1; k.C1 = a.C1; k.s2 = a.src2; k.C2 = a.src2 ? a.C2 : 0;
  k.qs1 = a.qs1; k.qs2 = a.qs2; k.gamma = a.gamma; k.beta = a.beta;
  k.H = a.H; k.W = a.W; k.mode = a.mode; k.silu = a.silu; k.B = a.B;
  k.Ho = a.mode == kPrepDown ? a.H / 2 : (a.mode == kPrepUp ? a.H * 2 : a.H);
  k.Wo = a.mode == kPrepDown ? a.W / 2 : (a.mode == kPrepUp ? a.W * 2 : a.W);
  k.outA = a.outA; k.outX = a.outX; k.outF = a.outF; k.outXF = a.outXF; k.overflow = a.overflow;
  const int C = k.C1 + k.C2;
  const int nout = k.Ho * k.Wo;
  const int per_thread = (a.mode == kPrepPlain) ? 8 : 4;
  const int ppi = 256 / (C / per_thread);
  int blocks = (nout + ppi - 1) / ppi;
  const int cap = std::max(1, (148 * 8) / a.B);
  if (blocks > cap) blocks = cap;
  dim3 grid(blocks, a.B);
  if (a.mode == kPrepPlain) {
    launch_k(gn_prep_plain_kernel, grid, dim3(256), 0, s, k);
  } else if (a.mode == kPrepDown) {       // output tile 4 x 8, 32 channels per CTA: 10 x 18 input region, 45 KB smem;
    // 8 float4 lanes per pixel = 128 B rows: the 8 threads of an LDS.128 phase read one contiguous row (conflict-free)
    dim3 g((k.Wo + 7) / 8, (k.Ho + 3) / 4, a.B * (C / 32));
    launch_k(gn_prep_resample_kernel<32, 4, 8, true>, g, dim3(256), 0, s, k);
  } else {                                // output tile 8 x 32, 32 channels per CTA: 6 x 18 input region, 27.6 KB smem
    dim3 g((k.Wo + 31) / 32, (k.Ho + 7) / 8, a.B * (C / 32));
    launch_k(gn_prep_resample_kernel<32, 8, 32, false>, g, dim3(256), 0, s, k);
  }
}

void launch_head_conv(const float* h, const double* qs, const float* gamma, const float* beta, const float* wf,
                      const float* bias, const float4* prev, float4* out, int B, int H, int W, int C, cudaStream_t s) {
  HeadK k{h, qs, gamma, beta, wf, bias, prev, out, B, H, W, C};
  if (static_cast<long>(B) * H * W >= 65536) launch_head_t<16, 4, 4, 4, 16>(k, s);
  else launch_head_t<8, 4, 1, 8, 128>(k, s);
}

}  // namespace flowse
