// Probe of tcgen05 shared-memory descriptor semantics on sm_100a (measurement tool, not product code).
// Question: can the M rows of a K-major SWIZZLE_128B operand start at an arbitrary 128-byte row of a larger
// swizzled region (row-shifted "views" of one halo tile), with 8-row groups SBO bytes apart, and what does the
// descriptor's base_offset field do?  Also probes the no-swizzle (interleaved) K-major layout.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_probe tools/umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../flowmse_b200/csrc/ptx.cuh"

constexpr int N = 64;          // MMA N
constexpr int ROWS = 416;      // rows of 128 B in the A region
constexpr int KSTEPS = 4;      // 4 x K=16 = one 64-channel chunk

struct Case {
  int layout;        // 2 = SWIZZLE_128B, 0 = no swizzle (interleaved)
  int start_row;     // A start row shift
  int sbo;           // bytes
  int lbo;           // bytes (no-swizzle only)
  int base_offset;   // descriptor bits [49,52)
  int row_pitch16;   // no-swizzle: unused
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int layout, int sbo, int lbo, int base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__host__ __device__ inline float a_val(int r, int k) { return static_cast<float>(((r * 7 + k * 3) % 17) - 8); }
__host__ __device__ inline float b_val(int n, int k) { return static_cast<float>(((n * 5 + k) % 13) - 6); }

__global__ void __launch_bounds__(128, 1) probe_kernel(Case c, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot_var;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  uint8_t* sA = gen;                       // ROWS x 128 B
  uint8_t* sB = gen + ROWS * 128;          // N rows x 128 B (always canonical SW128, 1024-aligned)
  const int tid = threadIdx.x;

  // ---- fill A
  if (c.layout == 2) {
    // what TMA SWIZZLE_128B writes for a box whose smem base is 1024-aligned: row r at r*128, 16-byte chunk j
    // of the row stored at chunk position j ^ (r & 7)
    for (int i = tid; i < ROWS * 64; i += 128) {
      const int r = i / 64, k = i % 64;
      const int chunk = (k >> 3) ^ (r & 7);
      reinterpret_cast<__half*>(sA + r * 128 + chunk * 16)[k & 7] = __float2half(a_val(r, k));
    }
  } else {
    // interleaved K-major: [k8 = k/8][row][8 elems]; row pitch 16 B, chunk pitch = lbo
    for (int i = tid; i < ROWS * 64; i += 128) {
      const int r = i / 64, k = i % 64;
      if (r * 16 + 16 <= c.lbo && (k >> 3) * c.lbo + r * 16 + 16 <= ROWS * 128)
        reinterpret_cast<__half*>(sA + (k >> 3) * c.lbo + r * 16)[k & 7] = __float2half(a_val(r, k));
    }
  }
  for (int i = tid; i < N * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const int chunk = (k >> 3) ^ (n & 7);
    reinterpret_cast<__half*>(sB + n * 128 + chunk * 16)[k & 7] = __float2half(b_val(n, k));
  }
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (tid < 32) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot_var), 64); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_var);

  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(128, N);
    const uint32_t a0 = base + c.start_row * (c.layout == 2 ? 128 : 16);
    const uint32_t b0 = base + ROWS * 128;
    for (int k = 0; k < KSTEPS; ++k) {
      const uint32_t koffA = (c.layout == 2) ? k * 32 : k * 2 * c.lbo;
      const uint64_t dA = make_desc(a0 + koffA, c.layout, c.sbo, c.lbo, c.base_offset);
      const uint64_t dB = make_desc(b0 + k * 32, 2, 1024, 16, 0);
      ptx::mma_f16_ss(tmem, dA, dB, idesc, k > 0 ? 1u : 0u);
    }
    ptx::mma_commit(ptx::smem_u32(&bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    ptx::tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

int main() {
  const int smem = ROWS * 128 + N * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* d_out;
  cudaMalloc(&d_out, 128 * N * sizeof(float));
  std::vector<float> h(128 * N);
  std::vector<Case> cases;
  // SWIZZLE_128B: (start_row, sbo, base_offset)
  for (int sbo : {1024, 1280, 2048})
    for (int s : {0, 1, 2, 5, 8, 11, 21})
      for (int bo : {0, -1}) cases.push_back(Case{2, s, sbo, 16, bo < 0 ? (s & 7) : 0, 0});
  // no swizzle: rows 16 B apart, core matrices sbo apart, k chunks lbo apart
  for (int sbo : {128, 160})
    for (int s : {0, 1, 11})
      for (int lbo : {2880, 4096}) cases.push_back(Case{0, s, sbo, lbo, 0, 0});
  int n_ok = 0;
  for (const Case& c : cases) {
    cudaMemset(d_out, 0, 128 * N * sizeof(float));
    probe_kernel<<<1, 128, smem>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case layout=%d s=%d sbo=%d lbo=%d bo=%d: CUDA error %s\n", c.layout, c.start_row, c.sbo, c.lbo, c.base_offset, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    // expected: row of M index m = start_row + (m / 8) * (sbo / rowbytes) + m % 8
    const int rowbytes = c.layout == 2 ? 128 : 16;
    int bad = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) {
      const int r = c.start_row + (m / 8) * (c.sbo / rowbytes) + (m % 8);
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += a_val(r, k) * b_val(n, k);
        if (h[m * N + n] != ref) { ++bad; if (first_bad < 0) first_bad = m; }
      }
    }
    printf("layout=%d start_row=%2d sbo=%4d lbo=%4d base_offset=%d : %s (bad=%d first_bad_row=%d)\n", c.layout,
           c.start_row, c.sbo, c.lbo, c.base_offset, bad ? "MISMATCH" : "ok", bad, first_bad);
    n_ok += bad == 0;
  }
  printf("%d / %zu cases ok\n", n_ok, cases.size());
  return 0;
}
