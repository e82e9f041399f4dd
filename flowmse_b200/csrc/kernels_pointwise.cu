// HBM-bound element-wise kernels of the sampler plus the small 4-channel "pyramid" kernels of NCSN++.
// Each cites the reference line it replaces (paths relative to /root/reference).
#include "flowse_internal.h"

#include <cstdlib>

namespace flowse {

long long& launch_counter() {
  static long long n = 0;
  return n;
}
int& pdl_mode() {
  // Programmatic dependent launch: 0 = off, 1 = every kernel, 2 = only the low-resolution conv_gemm launches, whose
  // prologue (barrier init, TMEM allocation, tensor-map prefetch, ~0.75 us) then overlaps the tail of the small kernel
  // before them.  Measured on B200 under graph replay (bench.py, B=1, T=512, N=5, two runs each, same box):
  // mode 0 23.39 / 23.14 ms, mode 2 23.00 / 23.03 ms, mode 1 24.26 / 24.05 ms per sampler call - with every kernel
  // opted in, early-scheduled CTAs of the next kernel compete with the running persistent kernels.  Mode 4 = mode 2 + the
  // halo conv launches (round 2, most of them follow another halo launch since the operand fusion): 21.14 vs 21.08 ms,
  // neutral.  Default 2; FLOWSE_PDL=<mode> / option "pdl" select another.
  static int mode = [] { const char* e = getenv("FLOWSE_PDL"); return (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 2; }();
  return mode;
}

namespace {

constexpr float kPiF = 3.14159274101257324219f;   // np.pi rounded to fp32 (scalar * fp32 tensor, layerspp.py:40)

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
// a*b + c with two roundings (no FMA contraction): bit-identical to torch's separate mul and add
__device__ __forceinline__ float madd(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }

// ------------------------------------------------------------------------------------------------
// Sampler element-wise updates.  complex64 is handled as float2 pairs; two complex per thread (float4).
// ------------------------------------------------------------------------------------------------
// x = y + sigma * z                                   (flowmse/odes.py:93-100)
__global__ void __launch_bounds__(256)
prior_kernel(const float4* __restrict__ y, const float4* __restrict__ z, float sigma, float4* __restrict__ x,
             size_t n4, const float2* y2, const float2* z2, float2* x2, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(y + i), b = __ldg(z + i);
    // complex * real-tensor std: (z.re*s - z.im*0, z.re*0 + z.im*s) == (z.re*s, z.im*s) for finite z
    x[i] = make_float4(madd(b.x, sigma, a.x), madd(b.y, sigma, a.y), madd(b.z, sigma, a.z), madd(b.w, sigma, a.w));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) {
    const float2 a = y2[n - 1], b = z2[n - 1];
    x2[n - 1] = make_float2(madd(b.x, sigma, a.x), madd(b.y, sigma, a.y));
  }
}

// out = a + c * b                                     (flowmse/sampling/odesolvers.py:42-47 with c = dt)
__global__ void __launch_bounds__(256)
axpy_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float c, float4* __restrict__ out, size_t n4,
            const float2* a2, const float2* b2, float2* o2, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 u = __ldg(a + i), v = __ldg(b + i);
    out[i] = make_float4(madd(v.x, c, u.x), madd(v.y, c, u.y), madd(v.z, c, u.z), madd(v.w, c, u.w));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) {
    const float2 u = a2[n - 1], v = b2[n - 1];
    o2[n - 1] = make_float2(madd(v.x, c, u.x), madd(v.y, c, u.y));
  }
}

// out = x + c * (v0 + v1)                             (Heun corrector)
__global__ void __launch_bounds__(256)
heun_kernel(const float2* __restrict__ x, const float2* __restrict__ v0, const float2* __restrict__ v1, float c,
            float2* __restrict__ out, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 a = x[i], p = v0[i], q = v1[i];
    out[i] = make_float2(madd(c, __fadd_rn(p.x, q.x), a.x), madd(c, __fadd_rn(p.y, q.y), a.y));
  }
}

// Adaptive Runge-Kutta building block (black-box RK45 solver, sampling/__init__.py:64-114 via scipy's solve_ivp):
//   v = base + sum_s coef[s] * K[s]        base, v complex128; K[s] complex64 stage derivatives (promoted exactly)
// optionally stored as complex128 (the integrator state) and / or rounded to complex64 (the next network input), and
// optionally reduced to sum |v / (atol + rtol * max(|ya|, |yb|))|^2 (scipy's scaled RMS error norm before the sqrt / n).
struct RkArgs {
  const double2* base; const float2* K; long long k_stride; double coef[8]; int S;
  double2* out64; float2* out32;
  const double2* ya; const double2* yb; double rtol, atol; double* sumsq;
};
__global__ void __launch_bounds__(256)
rk_lincomb_kernel(const RkArgs a, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  double local = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double2 v = a.base ? a.base[i] : make_double2(0.0, 0.0);
    for (int s = 0; s < a.S; ++s) {
      const float2 k = a.K[static_cast<size_t>(s) * a.k_stride + i];
      v.x += a.coef[s] * static_cast<double>(k.x);
      v.y += a.coef[s] * static_cast<double>(k.y);
    }
    if (a.out64) a.out64[i] = v;
    if (a.out32) a.out32[i] = make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
    if (a.sumsq) {
      const double2 p = a.ya[i], q = a.yb[i];
      const double scale = a.atol + a.rtol * fmax(hypot(p.x, p.y), hypot(q.x, q.y));
      const double m = hypot(v.x, v.y) / scale;
      local += m * m;
    }
  }
  if (a.sumsq) {
    __shared__ double red[256];
    red[threadIdx.x] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(a.sumsq, red[0]);
  }
}

__global__ void set_scalars_kernel(float* t_dev, int B, float t, float* step_dev, float step) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) t_dev[i] = t;
  if (i == 0) step_dev[0] = step;
}

struct TimeList { float t[64]; };
__global__ void set_times_kernel(float* t_all, int B, int count, const TimeList tl) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * count) t_all[i] = tl.t[i / B];
}

inline unsigned grid_for(size_t n, int per_block = 256, unsigned cap = 148 * 8) {
  size_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b > cap ? cap : b);
}

// ------------------------------------------------------------------------------------------------
// Time embedding (layerspp.py:39-41, ncsnpp.py:259,271-275) and all 49 Dense_0 biases (layerspp.py:262-263)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
temb_mlp_kernel(const TembWeights w, const float* __restrict__ t, float* __restrict__ temb_act) {
  __shared__ __align__(16) float emb[256];
  __shared__ __align__(16) float h1[512];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int j = threadIdx.x;
  if (j < 128) {
    const float lx = logf(t[b]);
    const float xp = ((lx * w.fourier_W[j]) * 2.0f) * kPiF;   // same fp32 op order as the reference
    emb[j] = sinf(xp);
    emb[128 + j] = cosf(xp);
  }
  __syncthreads();
  // one warp per output row, lanes stride over the inputs (coalesced weight reads), 16 warps x 32 rows each.  The kernel
  // is a single latency chain (one CTA per batch element), so a warp requests the weights of RB = 8 rows before it
  // consumes the first (RB2 = 4 rows for the 512-wide second layer: register budget): 4 / 8 instead of 32 dependent
  // global round trips per layer.
  const int warp = j >> 5, lane = j & 31;
  constexpr int RB = 8;
  for (int r0 = warp * RB; r0 < 512; r0 += 16 * RB) {
    float4 wv[RB][2];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const float4* row = reinterpret_cast<const float4*>(w.l1_w + static_cast<size_t>(r0 + k) * 256);
      wv[k][0] = __ldg(row + lane); wv[k][1] = __ldg(row + lane + 32);
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(&emb[4 * (lane + 32 * i)]);
        acc = fmaf(wv[k][i].x, x.x, acc); acc = fmaf(wv[k][i].y, x.y, acc);
        acc = fmaf(wv[k][i].z, x.z, acc); acc = fmaf(wv[k][i].w, x.w, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) h1[r0 + k] = silu_f(acc + w.l1_b[r0 + k]);
    }
  }
  __syncthreads();
  constexpr int RB2 = 4;
  for (int r0 = warp * RB2; r0 < 512; r0 += 16 * RB2) {
    float4 wv[RB2][4];
#pragma unroll
    for (int k = 0; k < RB2; ++k) {
      const float4* row = reinterpret_cast<const float4*>(w.l2_w + static_cast<size_t>(r0 + k) * 512);
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[k][i] = __ldg(row + lane + 32 * i);
    }
#pragma unroll
    for (int k = 0; k < RB2; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(&h1[4 * (lane + 32 * i)]);
        acc = fmaf(wv[k][i].x, x.x, acc); acc = fmaf(wv[k][i].y, x.y, acc);
        acc = fmaf(wv[k][i].z, x.z, acc); acc = fmaf(wv[k][i].w, x.w, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      // act(temb), input of every Dense_0
      if (lane == 0) temb_act[static_cast<size_t>(b) * 512 + r0 + k] = silu_f(acc + w.l2_b[r0 + k]);
    }
  }
}

// one warp per output row r: table[b][r] = dense_b[r] + dense_w[r] . temb_act[b]
__global__ void __launch_bounds__(256)
temb_dense_kernel(const float* __restrict__ dw, const float* __restrict__ db, const float* __restrict__ act, int R,
                  int B, float* __restrict__ table) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float4* row = reinterpret_cast<const float4*>(dw + static_cast<size_t>(warp) * 512);
  float4 wv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wv[i] = __ldg(row + lane + 32 * i);
  for (int b = 0; b < B; ++b) {
    const float4* a = reinterpret_cast<const float4*>(act + static_cast<size_t>(b) * 512);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 x = __ldg(a + lane + 32 * i);
      acc = fmaf(wv[i].x, x.x, acc); acc = fmaf(wv[i].y, x.y, acc);
      acc = fmaf(wv[i].z, x.z, acc); acc = fmaf(wv[i].w, x.w, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) table[static_cast<size_t>(b) * R + warp] = acc + db[warp];
  }
}

// ------------------------------------------------------------------------------------------------
// Input conv 3x3, 4 -> 128 (ncsnpp.py:253-254,285).  SIMT fp32: K = 36 is too thin for tensor cores.
// Block = 8 warps, tile 4 rows x 32 cols; warp g owns 16 pixels, lane owns 4 output channels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_in_kernel(const float2* __restrict__ x, const float2* __restrict__ y, const float* __restrict__ wgt,
               const float* __restrict__ bias, float* __restrict__ out, float4* __restrict__ pyr,
               double* __restrict__ qstats, int H, int W) {
  __shared__ float4 s_in[6][34];          // 4 channels per pixel, halo 1
  __shared__ float4 s_w[9][4][32];        // [tap][ci][lane] -> 4 consecutive output channels
  pdl_launch_dependents();
  const int b = blockIdx.z;
  const int h0 = blockIdx.y * 4, w0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * 4 * 32; i += 256) {      // weights do not depend on the previous kernel
    const int l = i & 31, ci = (i >> 5) & 3, tap = i >> 7;
    float4 v;
    // weight layout [128][4][3][3]
    v.x = wgt[((4 * l + 0) * 4 + ci) * 9 + tap];
    v.y = wgt[((4 * l + 1) * 4 + ci) * 9 + tap];
    v.z = wgt[((4 * l + 2) * 4 + ci) * 9 + tap];
    v.w = wgt[((4 * l + 3) * 4 + ci) * 9 + tap];
    s_w[tap][ci][l] = v;
  }
  pdl_wait();
  for (int i = tid; i < 6 * 34; i += 256) {
    const int r = i / 34, c = i % 34;
    const int h = h0 + r - 1, w = w0 + c - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const size_t p = (static_cast<size_t>(b) * H + h) * W + w;
      const float2 a = x[p], c2 = y[p];
      v = make_float4(a.x, a.y, c2.x, c2.y);
      if (r >= 1 && r <= 4 && c >= 1 && c <= 32) pyr[p] = v;
    }
    s_in[r][c] = v;
  }
  __syncthreads();
  const int g = tid >> 5, lane = tid & 31;
  const int row = g >> 1, col0 = (g & 1) * 16;
  float4 acc[16];
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + lane);
#pragma unroll
  for (int p = 0; p < 16; ++p) acc[p] = bv;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap % 3;
    const float4 w0v = s_w[tap][0][lane], w1v = s_w[tap][1][lane], w2v = s_w[tap][2][lane], w3v = s_w[tap][3][lane];
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const float4 in = s_in[row + dy][col0 + p + dx];
      acc[p].x = fmaf(in.x, w0v.x, acc[p].x); acc[p].y = fmaf(in.x, w0v.y, acc[p].y);
      acc[p].z = fmaf(in.x, w0v.z, acc[p].z); acc[p].w = fmaf(in.x, w0v.w, acc[p].w);
      acc[p].x = fmaf(in.y, w1v.x, acc[p].x); acc[p].y = fmaf(in.y, w1v.y, acc[p].y);
      acc[p].z = fmaf(in.y, w1v.z, acc[p].z); acc[p].w = fmaf(in.y, w1v.w, acc[p].w);
      acc[p].x = fmaf(in.z, w2v.x, acc[p].x); acc[p].y = fmaf(in.z, w2v.y, acc[p].y);
      acc[p].z = fmaf(in.z, w2v.z, acc[p].z); acc[p].w = fmaf(in.z, w2v.w, acc[p].w);
      acc[p].x = fmaf(in.w, w3v.x, acc[p].x); acc[p].y = fmaf(in.w, w3v.y, acc[p].y);
      acc[p].z = fmaf(in.w, w3v.z, acc[p].z); acc[p].w = fmaf(in.w, w3v.w, acc[p].w);
    }
  }
  const int h = h0 + row;
  float qs_s = 0.f, qs_q = 0.f;
  if (h < H) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const int w = w0 + col0 + p;
      if (w < W) {
        const size_t pix = (static_cast<size_t>(b) * H + h) * W + w;
        reinterpret_cast<float4*>(out + pix * 128)[lane] = acc[p];
        qs_s += (acc[p].x + acc[p].y) + (acc[p].z + acc[p].w);
        qs_q += (acc[p].x * acc[p].x + acc[p].y * acc[p].y) + (acc[p].z * acc[p].z + acc[p].w * acc[p].w);
      }
    }
  }
  if (qstats) {            // lane == channel quad; fold the 8 warps through shared memory, one atomic pair per quad
    __shared__ float rs[8][32], rq[8][32];
    rs[g][lane] = qs_s; rq[g][lane] = qs_q;
    __syncthreads();
    if (g == 0) {
      float as = 0.f, aq = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { as += rs[k][lane]; aq += rq[k][lane]; }
      double* dst = qstat_slot(qstats, b, blockIdx.y * gridDim.x + blockIdx.x, 32) + static_cast<size_t>(lane) * 2;
      atomicAdd(dst, static_cast<double>(as));
      atomicAdd(dst + 1, static_cast<double>(aq));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 4-channel pyramids (ncsnpp.py:310, 347-366)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fir_down4_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  // in: [B][2H][2W], out: [B][H][W]
  const size_t total = static_cast<size_t>(B) * H * W;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = i % W;
  const int h = (i / W) % H;
  const int b = i / (static_cast<size_t>(W) * H);
  const int Hi = 2 * H, Wi = 2 * W;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int hi = 2 * h - 1 + a;
    if (hi < 0 || hi >= Hi) continue;
    const float wa = (a == 0 || a == 3) ? 1.f : 3.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int wi = 2 * w - 1 + c;
      if (wi < 0 || wi >= Wi) continue;
      const float wgt = wa * ((c == 0 || c == 3) ? 1.f : 3.f) * (1.f / 64.f);
      const float4 v = __ldg(in + (static_cast<size_t>(b) * Hi + hi) * Wi + wi);
      acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
      acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
    }
  }
  out[i] = acc;
}

// out = h + conv1x1(4->C)(pyr) + b                  (Combine 'sum', layerspp.py:52-57); grid = (blocks, B).
// A thread owns one channel quad for all its pixels (weights in registers) and accumulates the output's quad stats.
__global__ void __launch_bounds__(256)
combine_kernel(const float* __restrict__ h, const float4* __restrict__ pyr, const float* __restrict__ w,
               const float* __restrict__ bias, float* __restrict__ out, double* __restrict__ qstats, int npix, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvec = C >> 2;
  const int ppi = blockDim.x / cvec;
  const int b = blockIdx.y;
  const int v = threadIdx.x % cvec, pp = threadIdx.x / cvec;
  const int c = v << 2;
  __shared__ float ss[256], sq[256];
  float ls = 0.f, lq = 0.f;
  if (pp < ppi) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c));
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (c + 0) * 4));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (c + 1) * 4));
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(w + (c + 2) * 4));
    const float4 w3 = __ldg(reinterpret_cast<const float4*>(w + (c + 3) * 4));
    for (int p = blockIdx.x * ppi + pp; p < npix; p += gridDim.x * ppi) {
      const size_t pix = static_cast<size_t>(b) * npix + p;
      const float4 pv = __ldg(pyr + pix);
      const float4 hv = __ldg(reinterpret_cast<const float4*>(h + pix * C + c));
      float4 o;
      o.x = hv.x + (fmaf(w0.w, pv.w, fmaf(w0.z, pv.z, fmaf(w0.y, pv.y, w0.x * pv.x))) + bv.x);
      o.y = hv.y + (fmaf(w1.w, pv.w, fmaf(w1.z, pv.z, fmaf(w1.y, pv.y, w1.x * pv.x))) + bv.y);
      o.z = hv.z + (fmaf(w2.w, pv.w, fmaf(w2.z, pv.z, fmaf(w2.y, pv.y, w2.x * pv.x))) + bv.z);
      o.w = hv.w + (fmaf(w3.w, pv.w, fmaf(w3.z, pv.z, fmaf(w3.y, pv.y, w3.x * pv.x))) + bv.w);
      *reinterpret_cast<float4*>(out + pix * C + c) = o;
      ls += (o.x + o.y) + (o.z + o.w);
      lq += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
    }
  }
  if (qstats) {
    ss[threadIdx.x] = ls; sq[threadIdx.x] = lq;
    __syncthreads();
    if (threadIdx.x < cvec) {
      float as = 0.f, aq = 0.f;
      for (int r = 0; r < ppi; ++r) { as += ss[r * cvec + threadIdx.x]; aq += sq[r * cvec + threadIdx.x]; }
      double* dst = qstat_slot(qstats, b, blockIdx.x, cvec) + static_cast<size_t>(threadIdx.x) * 2;
      atomicAdd(dst, static_cast<double>(as));
      atomicAdd(dst + 1, static_cast<double>(aq));
    }
  }
}

// d = output_layer(pyr / t); see launch_final for the modes
__global__ void __launch_bounds__(256)
final_kernel(const float4* __restrict__ pyr, const float* __restrict__ t, const float* __restrict__ wo,
             const float* __restrict__ bo, const float2* __restrict__ xin, const float* __restrict__ step_dev,
             float2* __restrict__ out, int mode, int B, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(B) * HW;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float w00 = wo[0], w01 = wo[1], w02 = wo[2], w03 = wo[3];
  const float w10 = wo[4], w11 = wo[5], w12 = wo[6], w13 = wo[7];
  const float b0 = bo[0], b1 = bo[1];
  const float step = (mode == 2) ? step_dev[0] : 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = i / HW;
    const float tb = t[b];
    float4 p = __ldg(pyr + i);
    p.x = __fdiv_rn(p.x, tb); p.y = __fdiv_rn(p.y, tb); p.z = __fdiv_rn(p.z, tb); p.w = __fdiv_rn(p.w, tb);
    const float dre = fmaf(w03, p.w, fmaf(w02, p.z, fmaf(w01, p.y, w00 * p.x))) + b0;
    const float dim = fmaf(w13, p.w, fmaf(w12, p.z, fmaf(w11, p.y, w10 * p.x))) + b1;
    float2 o;
    if (mode == 0) o = make_float2(dre, dim);
    else if (mode == 1) o = make_float2(-dre, -dim);
    else { const float2 xv = xin[i]; o = make_float2(madd(step, dre, xv.x), madd(step, dim, xv.y)); }
    out[i] = o;
  }
}

// row softmax, one warp per row                        (layerspp.py:84)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ s, int rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = s + static_cast<size_t>(row) * cols;
  float m = -INFINITY;
  for (int i = lane; i < cols; i += 32) m = fmaxf(m, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int i = lane; i < cols; i += 32) { const float e = expf(r[i] - m); r[i] = e; sum += e; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int i = lane; i < cols; i += 32) r[i] = __fdiv_rn(r[i], sum);
}

}  // namespace

void launch_set_scalars(float* t_dev, int B, float t, float* step_dev, float step, cudaStream_t s) {
  launch_k(set_scalars_kernel, dim3((B + 127) / 128), dim3(128), 0, s, t_dev, B, t, step_dev, step);
}

void launch_rk_lincomb(const double2* base, const float2* K, long long k_stride, const double* coef, int S, double2* out64,
                       float2* out32, const double2* ya, const double2* yb, double rtol, double atol, double* sumsq, size_t n,
                       cudaStream_t s) {
  RkArgs a{};
  a.base = base; a.K = K; a.k_stride = k_stride; a.S = S;
  for (int i = 0; i < S && i < 8; ++i) a.coef[i] = coef[i];
  a.out64 = out64; a.out32 = out32; a.ya = ya; a.yb = yb; a.rtol = rtol; a.atol = atol; a.sumsq = sumsq;
  launch_k(rk_lincomb_kernel, dim3(grid_for(n)), dim3(256), 0, s, a, n);
}

void launch_set_times(float* t_all, int B, const float* times_host, int count, cudaStream_t s) {
  TimeList tl{};
  for (int i = 0; i < count && i < 64; ++i) tl.t[i] = times_host[i];
  launch_k(set_times_kernel, dim3((B * count + 127) / 128), dim3(128), 0, s, t_all, B, count, tl);
}

void launch_prior(const float2* y, const float2* z, float sigma, float2* x, size_t n, cudaStream_t s) {
  const size_t n4 = n / 2;
  launch_k(prior_kernel, dim3(grid_for(n4)), dim3(256), 0, s, reinterpret_cast<const float4*>(y),
           reinterpret_cast<const float4*>(z), sigma, reinterpret_cast<float4*>(x), n4, y, z, x, n);
}

void launch_axpy_c(const float2* a, const float2* b, float c, float2* out, size_t n, cudaStream_t s) {
  const size_t n4 = n / 2;
  launch_k(axpy_kernel, dim3(grid_for(n4)), dim3(256), 0, s, reinterpret_cast<const float4*>(a),
           reinterpret_cast<const float4*>(b), c, reinterpret_cast<float4*>(out), n4, a, b, out, n);
}

void launch_euler_update(const float2* x, const float2* v, float dt, float2* out, size_t n, cudaStream_t s) {
  launch_axpy_c(x, v, dt, out, n, s);
}

void launch_heun_combine(const float2* x, const float2* v0, const float2* v1, float c, float2* out, size_t n,
                         cudaStream_t s) {
  launch_k(heun_kernel, dim3(grid_for(n)), dim3(256), 0, s, x, v0, v1, c, out, n);
}

void launch_temb(const TembWeights& w, const float* t, int B, float* temb_act, float* bias_table, cudaStream_t s) {
  launch_k(temb_mlp_kernel, dim3(B), dim3(512), 0, s, w, t, temb_act);
  const int warps_per_block = 8;
  launch_k(temb_dense_kernel, dim3((w.R + warps_per_block - 1) / warps_per_block), dim3(256), 0, s, w.dense_w, w.dense_b,
           static_cast<const float*>(temb_act), w.R, B, bias_table);
}

void launch_conv_in(const float2* x, const float2* y, const float* w, const float* bias, float* out, float4* pyr,
                    double* qstats, int B, int H, int W, cudaStream_t s) {
  dim3 grid((W + 31) / 32, (H + 3) / 4, B);
  launch_k(conv_in_kernel, grid, dim3(256), 0, s, x, y, w, bias, out, pyr, qstats, H, W);
}

void launch_fir_down4(const float4* in, float4* out, int B, int H, int W, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W;
  launch_k(fir_down4_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, in, out, B, H, W);
}

void launch_combine(const float* h, const float4* pyr, const float* w, const float* b, float* out, double* qstats,
                    int B, int H, int W, int C, cudaStream_t s) {
  const int npix = H * W;
  const int ppi = 256 / (C / 4);
  int blocks = (npix + ppi - 1) / ppi;
  const int cap = std::max(1, (148 * 8) / B);
  if (blocks > cap) blocks = cap;
  dim3 grid(blocks, B);
  launch_k(combine_kernel, grid, dim3(256), 0, s, h, pyr, w, b, out, qstats, npix, C);
}

// Holds the stream for `ns` nanoseconds: flowse_profile_forward queues a whole evaluation behind it so that the per-op
// events time kernels running back to back, not the host's launch cadence.
__global__ void spin_kernel(unsigned long long ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns) break;
  }
}
void launch_spin(unsigned long long ns, cudaStream_t s) { spin_kernel<<<1, 1, 0, s>>>(ns); }

void launch_final(const float4* pyr, const float* t, const float* wo, const float* bo, const float2* xin,
                  const float* stepsize_dev, float2* out, int mode, int B, int HW, cudaStream_t s) {
  launch_k(final_kernel, dim3(grid_for(static_cast<size_t>(B) * HW)), dim3(256), 0, s, pyr, t, wo, bo, xin, stepsize_dev,
           out, mode, B, HW);
}

void launch_softmax_rows(float* sm, int rows, int cols, cudaStream_t st) {
  launch_k(softmax_rows_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, sm, rows, cols);
}

}  // namespace flowse
