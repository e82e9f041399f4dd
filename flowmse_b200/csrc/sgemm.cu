// Small batched fp32 SIMT GEMM used by the four attention blocks of NCSN++ (0.15 % of the FLOPs per NFE):
// q/k/v projections (NIN, /root/reference/flowmse/backbones/ncsnpp_utils/layers.py:546-555), q.k^T scores,
// softmax(scores).v and the NIN_3 output projection with the residual epilogue (layerspp.py:75-91).
// Exact fp32 FMA arithmetic (these products feed a softmax: no operand splitting, no tensor cores).
//
// TM x 64 tile (TM = 64 or 32, picked so that the few-hundred-token problems still fill the GPU), K step 32,
// 256 threads, (TM/16) x 4 register micro-tile.  Operands are staged k-major in shared memory so that the inner loop
// is two 16-byte shared loads per 4 x 4 FMAs; the global loads of the next TWO K steps are in flight in registers while
// the current step is computed (the problems are latency-bound: K <= 512).
#include "flowse_internal.h"

namespace flowse {

namespace {

constexpr int TN = 64, TK = 32, LDS_PAD = 4;

template <int TM>
__global__ void __launch_bounds__(256)
sgemm_kernel(const SgemmArgs a) {
  constexpr int RM = TM / 16;                         // rows per thread
  constexpr int A_F4 = TM * TK / 4 / 256;             // float4 loads of A per thread per K step (2 or 1)
  constexpr int B_F4 = TN * TK / 4 / 256;             // 2
  __shared__ __align__(16) float sA[TK][TM + LDS_PAD];
  __shared__ __align__(16) float sB[TK][TN + LDS_PAD];
  pdl_launch_dependents();
  pdl_wait();
  const int bz = blockIdx.z;
  const float* __restrict__ A = a.A + bz * a.strideA;
  const float* __restrict__ Bm = a.Bm + bz * a.strideB;
  float* __restrict__ C = a.C + bz * a.strideC;
  const float* __restrict__ R = a.residual ? a.residual + bz * a.strideR : nullptr;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vecA = (a.lda % 4 == 0) && ((reinterpret_cast<size_t>(A) & 15) == 0);
  const bool vecB = (a.ldb % 4 == 0) && ((reinterpret_cast<size_t>(Bm) & 15) == 0);

  float4 ra0[A_F4], rb0[B_F4], ra1[A_F4], rb1[B_F4];     // two K steps of global loads in flight
  // A tile: TM rows x 32 k, row-major in global: float4 f -> row f / 8, k4 = f % 8
  auto load_a = [&](int k0, float4* ra) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      const int f = tid + i * 256;
      const int r = f >> 3, k = k0 + ((f & 7) << 2);
      const int m = m0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M) {
        const float* src = A + static_cast<size_t>(m) * a.lda + k;
        if (vecA && k + 3 < a.K) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (k < a.K) v.x = __ldg(src);
          if (k + 1 < a.K) v.y = __ldg(src + 1);
          if (k + 2 < a.K) v.z = __ldg(src + 2);
          if (k + 3 < a.K) v.w = __ldg(src + 3);
        }
      }
      ra[i] = v;
    }
  };
  auto store_a = [&](const float4* ra) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      const int f = tid + i * 256;
      const int r = f >> 3, k = (f & 7) << 2;
      sA[k][r] = ra[i].x; sA[k + 1][r] = ra[i].y; sA[k + 2][r] = ra[i].z; sA[k + 3][r] = ra[i].w;
    }
  };
  // B tile: transB ? [N][K] (float4 along k, stored transposed) : [K][N] (float4 along n, stored as is)
  auto load_b = [&](int k0, float4* rb) {
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      const int f = tid + i * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.transB) {
        const int c = f >> 3, k = k0 + ((f & 7) << 2);
        const int n = n0 + c;
        if (n < a.N) {
          const float* src = Bm + static_cast<size_t>(n) * a.ldb + k;
          if (vecB && k + 3 < a.K) v = __ldg(reinterpret_cast<const float4*>(src));
          else {
            if (k < a.K) v.x = __ldg(src);
            if (k + 1 < a.K) v.y = __ldg(src + 1);
            if (k + 2 < a.K) v.z = __ldg(src + 2);
            if (k + 3 < a.K) v.w = __ldg(src + 3);
          }
        }
      } else {
        const int kk = f >> 4, n = n0 + ((f & 15) << 2);
        const int k = k0 + kk;
        if (k < a.K) {
          const float* src = Bm + static_cast<size_t>(k) * a.ldb + n;
          if (vecB && n + 3 < a.N) v = __ldg(reinterpret_cast<const float4*>(src));
          else {
            if (n < a.N) v.x = __ldg(src);
            if (n + 1 < a.N) v.y = __ldg(src + 1);
            if (n + 2 < a.N) v.z = __ldg(src + 2);
            if (n + 3 < a.N) v.w = __ldg(src + 3);
          }
        }
      }
      rb[i] = v;
    }
  };
  auto store_b = [&](const float4* rb) {
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      const int f = tid + i * 256;
      if (a.transB) {
        const int c = f >> 3, k = (f & 7) << 2;
        sB[k][c] = rb[i].x; sB[k + 1][c] = rb[i].y; sB[k + 2][c] = rb[i].z; sB[k + 3][c] = rb[i].w;
      } else {
        const int kk = f >> 4, c = (f & 15) << 2;
        *reinterpret_cast<float4*>(&sB[kk][c]) = rb[i];
      }
    }
  };

  float acc[RM][4] = {};
  auto compute = [&]() {
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[RM];
      if constexpr (RM == 4) {
        const float4 v = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
        av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(&sA[kk][ty * 2]);
        av[0] = v.x; av[1] = v.y;
      }
      const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        acc[i][0] = fmaf(av[i], bv.x, acc[i][0]); acc[i][1] = fmaf(av[i], bv.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], bv.z, acc[i][2]); acc[i][3] = fmaf(av[i], bv.w, acc[i][3]);
      }
    }
  };
  // The problems are latency-bound (K <= 512, a handful of CTAs per SM): the loads of K steps s+1 and s+2 are in flight
  // while step s is computed (two register sets, loop unrolled by two).
  load_a(0, ra0); load_b(0, rb0);
  if (TK < a.K) { load_a(TK, ra1); load_b(TK, rb1); }
  for (int k0 = 0; k0 < a.K; k0 += 2 * TK) {
    store_a(ra0); store_b(rb0);
    __syncthreads();
    if (k0 + 2 * TK < a.K) { load_a(k0 + 2 * TK, ra0); load_b(k0 + 2 * TK, rb0); }
    compute();
    __syncthreads();
    if (k0 + TK >= a.K) break;
    store_a(ra1); store_b(rb1);
    __syncthreads();
    if (k0 + 3 * TK < a.K) { load_a(k0 + 3 * TK, ra1); load_b(k0 + 3 * TK, rb1); }
    compute();
    __syncthreads();
  }
  const bool vecC = (a.ldc % 4 == 0) && ((reinterpret_cast<size_t>(C) & 15) == 0) &&
                    (!R || ((a.ldr % 4 == 0) && ((reinterpret_cast<size_t>(R) & 15) == 0)));
  float qs_s = 0.f, qs_q = 0.f;                      // this thread's share of the quad statistics (its 4 columns = one quad)
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= a.M) continue;
    const int n = n0 + tx * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] * a.alpha;
      if (n + j < a.N) {
        if (a.bias) v[j] += a.bias[n + j];
        if (R) v[j] = R[static_cast<size_t>(m) * a.ldr + n + j] + v[j];
        if (a.div_sqrt2) v[j] = __fdiv_rn(v[j], kSqrt2);
      }
    }
    float* dst = C + static_cast<size_t>(m) * a.ldc + n;
    if (vecC && n + 3 < a.N) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else
#pragma unroll
      for (int j = 0; j < 4; ++j) if (n + j < a.N) dst[j] = v[j];
    if (n + 3 < a.N) {
      qs_s += (v[0] + v[1]) + (v[2] + v[3]);
      qs_q += (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
    }
  }
  if (a.qstats) {
    // the tile's rows belong to one batch element (launcher contract): fold the 16 row groups in fixed order, then one
    // fp64 atomic pair per quad into the replica keyed by the row-tile index
    float (*s_q)[16][2] = reinterpret_cast<float (*)[16][2]>(&sA[0][0]);     // the K loop has ended with a barrier
    s_q[ty][tx][0] = qs_s; s_q[ty][tx][1] = qs_q;
    __syncthreads();
    if (ty == 0 && n0 + tx * 4 + 3 < a.N) {
      float as = 0.f, aq = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) { as += s_q[g][tx][0]; aq += s_q[g][tx][1]; }
      const int b = m0 / a.qs_rows_per_batch;
      double* dst = qstat_slot(a.qstats, b, blockIdx.y, a.N >> 2) + static_cast<size_t>((n0 >> 2) + tx) * 2;
      atomicAdd(dst, static_cast<double>(as));
      atomicAdd(dst + 1, static_cast<double>(aq));
    }
  }
}

}  // namespace

int sgemm_tile_rows(const SgemmArgs& a) {
  const long long tiles64 = static_cast<long long>((a.N + TN - 1) / TN) * ((a.M + 63) / 64) * a.batch;
  return (tiles64 >= 120 && a.M > 32) ? 64 : 32;
}

void launch_sgemm(const SgemmArgs& a, cudaStream_t s) {
  if (sgemm_tile_rows(a) == 64) {
    dim3 grid((a.N + TN - 1) / TN, (a.M + 63) / 64, a.batch);
    launch_k(sgemm_kernel<64>, grid, dim3(256), 0, s, a);
  } else {
    dim3 grid((a.N + TN - 1) / TN, (a.M + 31) / 32, a.batch);
    launch_k(sgemm_kernel<32>, grid, dim3(256), 0, s, a);
  }
}

}  // namespace flowse
