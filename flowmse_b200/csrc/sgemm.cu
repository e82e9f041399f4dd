// Small batched fp32 SIMT GEMM used by the four attention blocks of NCSN++ (0.15 % of the FLOPs per NFE):
// q/k/v projections (NIN, /root/reference/flowmse/backbones/ncsnpp_utils/layers.py:546-555), q.k^T scores,
// softmax(scores).v and the NIN_3 output projection with the residual epilogue (layerspp.py:75-91).
// Exact fp32 FMA arithmetic; 64x64 tile, K step 16, 256 threads, 4x4 register micro-tile.
#include "flowse_internal.h"

namespace flowse {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(const SgemmArgs a) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int bz = blockIdx.z;
  const float* A = a.A + bz * a.strideA;
  const float* Bm = a.Bm + bz * a.strideB;
  float* C = a.C + bz * a.strideC;
  const float* R = a.residual ? a.residual + bz * a.strideR : nullptr;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;        // 16 x 16 threads, each 4 rows x 4 cols
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += TK) {
    // A tile: 64 rows x 16 k
    for (int i = tid; i < TM * TK; i += 256) {
      const int r = i / TK, kk = i % TK;
      const int m = m0 + r, k = k0 + kk;
      sA[kk][r] = (m < a.M && k < a.K) ? A[static_cast<size_t>(m) * a.lda + k] : 0.f;
    }
    if (a.transB) {   // B given as [N][K]
      for (int i = tid; i < TN * TK; i += 256) {
        const int c = i / TK, kk = i % TK;
        const int n = n0 + c, k = k0 + kk;
        sB[kk][c] = (n < a.N && k < a.K) ? Bm[static_cast<size_t>(n) * a.ldb + k] : 0.f;
      }
    } else {          // B given as [K][N]
      for (int i = tid; i < TN * TK; i += 256) {
        const int kk = i / TN, c = i % TN;
        const int n = n0 + c, k = k0 + kk;
        sB[kk][c] = (n < a.N && k < a.K) ? Bm[static_cast<size_t>(k) * a.ldb + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] * a.alpha;
      if (a.bias) v += a.bias[n];
      if (R) v = R[static_cast<size_t>(m) * a.ldr + n] + v;
      if (a.div_sqrt2) v = __fdiv_rn(v, kSqrt2);
      C[static_cast<size_t>(m) * a.ldc + n] = v;
    }
  }
}

}  // namespace

void launch_sgemm(const SgemmArgs& a, cudaStream_t s) {
  dim3 grid((a.N + TN - 1) / TN, (a.M + TM - 1) / TM, a.batch);
  sgemm_kernel<<<grid, 256, 0, s>>>(a);
  ++launch_counter();
}

}  // namespace flowse
