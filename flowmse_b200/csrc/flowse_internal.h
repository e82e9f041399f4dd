// Internal declarations shared by the .cu translation units of libflowse.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstddef>
#include <cstdint>
#include <string>

namespace flowse {

// Process-wide count of kernel launches issued by this library (incremented by every launch_* helper).
long long& launch_counter();

// Function attributes (dynamic shared-memory limit, non-portable cluster sizes) and occupancy answers belong to a DEVICE:
// a per-launcher cache has one entry per device, indexed by the current one.
template <class T>
struct PerDevice {
  T v[64];
  explicit PerDevice(T init) { for (auto& e : v) e = init; }
  T& get() {
    int d = 0;
    cudaGetDevice(&d);
    return v[(d >= 0 && d < 64) ? d : 0];
  }
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  One NFE is a chain of ~330 mostly short kernels; every kernel of the library
//   1. calls pdl_launch_dependents() first: the NEXT kernel of the stream / graph may be scheduled as soon as all CTAs
//      of this one are running, so its launch latency and input-independent prologue (barrier init, TMEM allocation,
//      tensor-map prefetch, weight staging in registers) overlap this kernel;
//   2. calls pdl_wait() before its first access to global memory another kernel may have produced or may still read
//      (griddepcontrol.wait returns once all prerequisite grids have COMPLETED and their writes are visible).
// launch_k() sets cudaLaunchAttributeProgrammaticStreamSerialization in mode 1 only; the conv_gemm launches set it in
// modes 1 and 2 (the default: measurements in pdl_mode(), kernels_pointwise.cu).
// ---------------------------------------------------------------------------------------------
int& pdl_mode();      // 0 off, 1 all kernels, 2 low-resolution conv_gemm launches only, 3 = 2 + small grids, 4 = 2 + halo launches
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed =
      (pdl_mode() == 1 || (pdl_mode() == 3 && static_cast<size_t>(grid.x) * grid.y * grid.z <= 148)) ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ++launch_counter();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

constexpr int kGroups = 32;          // GroupNorm groups: min(C/4, 32) == 32 for every C in the net (layerspp.py:219)
constexpr float kGnEps = 1e-6f;
constexpr float kSqrt2 = 1.41421356237309504880f;   // np.sqrt(2.) rounded to fp32 (layerspp.py:274)

// Quad statistics of one tensor: [B][kStatReplicas][C/4]{sum, sumsq} in fp64.  Producers spread their atomics over the
// replicas (keyed by tile / block index) so that one address is not hammered by every CTA; consumers add them up.
constexpr int kStatReplicas = 8;
__host__ __device__ inline double* qstat_slot(double* base, int b, int replica_key, int quads) {
  return base + (static_cast<size_t>(b) * kStatReplicas + (replica_key & (kStatReplicas - 1))) * quads * 2;
}
__host__ __device__ inline const double* qstat_slot(const double* base, int b, int replica_key, int quads) {
  return base + (static_cast<size_t>(b) * kStatReplicas + (replica_key & (kStatReplicas - 1))) * quads * 2;
}

// ---------------------------------------------------------------------------------------------
// Element-wise / small kernels (kernels_pointwise.cu)
// ---------------------------------------------------------------------------------------------
// x = y + sigma * z  (FLOWMATCHING.prior_sampling, odes.py:93-100); n complex elements.
void launch_prior(const float2* y, const float2* z, float sigma, float2* x, size_t n, cudaStream_t s);
// out = x + v * dt (EulerODEsolver.update_fn, odesolvers.py:42-47, dt = -stepsize); n complex elements.
void launch_euler_update(const float2* x, const float2* v, float dt, float2* out, size_t n, cudaStream_t s);
// out = a + c * b   (complex a, b; real scalar c) - Heun / midpoint helper.
void launch_axpy_c(const float2* a, const float2* b, float c, float2* out, size_t n, cudaStream_t s);
// out = x + c * (v0 + v1)
void launch_heun_combine(const float2* x, const float2* v0, const float2* v1, float c, float2* out, size_t n,
                         cudaStream_t s);

// v = base + sum_s coef[s] K[s] (complex128 state, complex64 stages): see rk_lincomb_kernel; sumsq must be zeroed
void launch_rk_lincomb(const double2* base, const float2* K, long long k_stride, const double* coef, int S, double2* out64,
                       float2* out32, const double2* ya, const double2* yb, double rtol, double atol, double* sumsq, size_t n,
                       cudaStream_t s);

// t_dev[0..B) = t; step_dev[0] = step (scalars passed as kernel arguments)
void launch_set_scalars(float* t_dev, int B, float t, float* step_dev, float step, cudaStream_t s);
// t_all[e*B + b] = times[e] for e < count <= 64 (the evaluation times of a sampler call, passed by value)
void launch_set_times(float* t_all, int B, const float* times_host, int count, cudaStream_t s);

struct TembWeights {
  const float* fourier_W;   // [128]
  const float* l1_w;        // [512][256]
  const float* l1_b;        // [512]
  const float* l2_w;        // [512][512]
  const float* l2_b;        // [512]
  const float* dense_w;     // [R][512]  all Dense_0 stacked
  const float* dense_b;     // [R]       Dense_0.bias + Conv_0.bias
  int R;
};
// temb_act[b][512] = SiLU(temb(t[b])); bias_table[b][R] = dense_b + dense_w . temb_act[b]
void launch_temb(const TembWeights& w, const float* t, int B, float* temb_act, float* bias_table, cudaStream_t s);

// conv3x3 4->128 on (x.re, x.im, y.re, y.im) (ncsnpp.py:253-254,285); also writes the 4-plane input pyramid.
void launch_conv_in(const float2* x, const float2* y, const float* w /*[128][4][3][3]*/, const float* bias,
                    float* out /*[B,H,W,128]*/, float4* pyr /*[B,H,W,4]*/, double* qstats /*quad statistics of the output, zeroed*/,
                    int B, int H, int W, cudaStream_t s);
// 4-channel FIR downsample x2 of the input pyramid (ncsnpp.py:310): [B,2H,2W,4] -> [B,H,W,4]
void launch_fir_down4(const float4* in, float4* out, int B, int H, int W, cudaStream_t s);
// Combine(method='sum') (layerspp.py:52-57): out = h + conv1x1(4->C)(pyr) + b
void launch_combine(const float* h, const float4* pyr, const float* w /*[C][4]*/, const float* b, float* out,
                    double* qstats /*quad statistics of the output, zeroed*/, int B, int H, int W, int C, cudaStream_t s);
// d = output_layer(pyr / t) (ncsnpp.py:398-403).
//   mode 0: out = d                         (NCSNpp.forward result)
//   mode 1: out = -d                        (VFModel.forward result, model.py:164-170)
//   mode 2: out = xin + stepsize * d        (fused Euler update: x + (-d) * (-stepsize))
void launch_final(const float4* pyr, const float* t /*[B]*/, const float* wo /*[2][4]*/, const float* bo /*[2]*/,
                  const float2* xin, const float* stepsize_dev, float2* out, int mode, int B, int HW, cudaStream_t s);
// busy-wait `ns` nanoseconds on the stream (measurement helper, not counted as a library launch)
void launch_spin(unsigned long long ns, cudaStream_t s);
// row softmax in place, rows x cols fp32
void launch_softmax_rows(float* s, int rows, int cols, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics + operand preparation (kernels_gn.cu)
// ---------------------------------------------------------------------------------------------
// Quad statistics (layout above) of an NHWC fp32 tensor [B][npix][C]; GroupNorm consumers assemble their (possibly
// concat-straddling) groups from them.  Fused producers accumulate with fp64 atomics into a zeroed buffer; this
// standalone kernel overwrites replica 0 (the others must be zero).  partials: scratch of B * gn_stats_max_blocks() * 256 doubles;
// counters: B unsigned, zero before first use (left zero).
int gn_stats_max_blocks();
void launch_quad_stats(const float* src, int C, int B, int npix, double* qs, double* partials, unsigned* counters,
                       cudaStream_t s);

enum PrepMode { kPrepPlain = 0, kPrepDown = 1, kPrepUp = 2 };
struct PrepArgs {
  const float* src1; int C1;
  const float* src2; int C2;      // virtual concat [src1, src2] on the channel axis
  const double* qs1; const double* qs2;   // quad statistics of src1 / src2
  const float* gamma; const float* beta;
  int B, H, W;                    // INPUT resolution
  int mode;                       // PrepMode: output resolution is H/2 (down), 2H (up)
  int silu;                       // apply SiLU after the affine normalisation
  __half* outA;                   // [2][B][Ho][Wo][C] fp16 hi/lo split of act(GN(x)) (may be null)
  __half* outX;                   // [2][B][Ho][Wo][C] fp16 hi/lo split of (resampled) raw x (may be null)
  float* outF;                    // [B][Ho][Wo][C] fp32 act(GN(x)) (may be null)
  float* outXF;                   // [B][Ho][Wo][C] fp32 resampled raw x (may be null)
  unsigned long long* overflow;   // optional: counts operand values outside the fp16 hi/lo range (|v| > 65504)
};
void launch_gn_prep(const PrepArgs& a, cudaStream_t s);

// Fused pyramid head (ncsnpp.py:347-366): out = FIR-up(prev) + conv3x3(C -> 4)(SiLU(GN(h))) + bias, fp32 SIMT.
// wf: [9][C][4] fp32 (tap-major); prev: coarser pyramid level [B][H/2][W/2] or null; C % 64 == 0.
void launch_head_conv(const float* h, const double* qs, const float* gamma, const float* beta, const float* wf,
                      const float* bias, const float4* prev, float4* out, int B, int H, int W, int C, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM convolution on tcgen05 (conv_gemm.cu)
// ---------------------------------------------------------------------------------------------
// An operand the conv kernel prepares ITSELF from fp32 NHWC activations (halo kernel only): the virtual channel concat
// [s1, s2] is normalised (GroupNorm from the producers' quad statistics, affine gamma / beta), activated (SiLU) and split
// to fp16 hi/lo by transform warps straight into the swizzled shared-memory operand tile - the standalone prep pass
// (one full read + write of the tensor) disappears.  gamma == nullptr: raw split only (the 1x1 shortcut operand).
struct FusedOperand {
  const float* s1 = nullptr; int C1 = 0;
  const float* s2 = nullptr; int C2 = 0;
  const double* qs1 = nullptr; const double* qs2 = nullptr;
  const float* gamma = nullptr; const float* beta = nullptr;
  int silu = 0;
};

struct ConvGemmArgs {
  const __half* A;      // [2][B][H][W][Cin]  hi/lo split activations
  int Cin;              // multiple of 64
  int ntaps;            // 9 (3x3, pad 1) or 1 (1x1)
  const __half* X;      // optional shortcut operand [2][B][H][W][Cin2] (extra 1x1 K-blocks), may be null
  int Cin2;             // multiple of 64 or 0
  const __half* Wp;     // [2][Npad][K] hi/lo split packed weights, K = ntaps*Cin + Cin2, scaled by 2^wexp
  int Npad;             // rows of Wp (multiple of the N tile)
  float wscale_inv;     // 2^-wexp
  const float* bias;    // [bias_bstride ? B : 1][Cout]
  int bias_bstride;     // 0 or stride (floats) between batch elements
  const float* residual;// optional fp32 [B][H][W][ldc]
  int div_sqrt2;        // divide the result by sqrt(2) (skip_rescale)
  float* out;           // fp32 [B][H][W][ldc]
  int Cout;             // valid output channels (<= Npad)
  int ldc;
  int B, H, W;
  double* qstats;               // optional: accumulate quad statistics of the OUTPUT (zeroed buffer)
  float* splitk_scratch;        // optional fp32 scratch enabling split-K for low-resolution layers (may be null)
  size_t splitk_scratch_elems;
  FusedOperand fA, fX;          // halo kernel: prepare the main / shortcut operand in the kernel (s1 != nullptr) instead of A / X
  unsigned long long* overflow; // optional fp16-range counter for the fused operands
};
constexpr size_t kSplitKScratchElems = static_cast<size_t>(148) * 128 * 128;   // enough for any one-wave split
// returns 0 on success; fills err otherwise.
int launch_conv_gemm(const ConvGemmArgs& a, cudaStream_t s, std::string* err);
// Slow SIMT evaluation of exactly the same operands (debug / cross-check only; never on the product path).
int launch_conv_gemm_simt(const ConvGemmArgs& a, cudaStream_t s, std::string* err);
// Halo variant (conv_halo.cu): persistent, one TMA halo load per 64-channel chunk, 3x3 only, H % 16 == 0, W % 8 == 0.
// nmain = number of rotating hi*hi accumulator slots (1: double-buffered TMEM, 3: single buffer).
bool conv_halo_supported(const ConvGemmArgs& a);
int launch_conv_halo(const ConvGemmArgs& a, int nmain, cudaStream_t s, std::string* err);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda), and the weight
// tensor map both conv kernels use: [2][Npad][K] fp16 -> 3-D map with box {64, rows, 1}, 128-byte swizzle.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder(std::string* err);
bool make_weight_map(CUtensorMap* m, const __half* base, int Npad, int K, int rows, std::string* err);
// Host-side packing: fp32 [Cout][Cin][kh][kw] (+ optional 1x1 shortcut [Cout][Cin2]) -> K-major fp16 hi/lo.
// out_hi/out_lo: [Npad][K]; returns the power-of-two exponent used.
int pack_conv_weights_host(const float* w_main, int Cout, int Cin, int ntaps, const float* w_sc, int Cin2,
                           int Npad, __half* out_hi, __half* out_lo);

// ---------------------------------------------------------------------------------------------
// fp32 SIMT GEMM for the attention blocks (sgemm.cu)
// ---------------------------------------------------------------------------------------------
struct SgemmArgs {
  const float* A; int lda; long long strideA;      // [M][K] row-major
  const float* Bm; int ldb; long long strideB;     // transB ? [N][K] : [K][N]
  int transB;
  float* C; int ldc; long long strideC;
  int M, N, K, batch;
  float alpha;                                     // applied to the accumulated product
  const float* bias;                               // [N] or null
  const float* residual; int ldr; long long strideR;   // optional, added after bias
  int div_sqrt2;
  __half* split_out;                               // optional: also not used (reserved)
  // optional: accumulate the quad statistics of C (layout above, zeroed buffer) in the epilogue.  Rows are grouped into
  // batch elements of qs_rows_per_batch rows; needs batch == 1, N % 4 == 0 and qs_rows_per_batch % sgemm_tile_rows() == 0.
  double* qstats; int qs_rows_per_batch;
};
void launch_sgemm(const SgemmArgs& a, cudaStream_t s);
int sgemm_tile_rows(const SgemmArgs& a);           // M tile the launcher will pick (64 or 32)

// ---------------------------------------------------------------------------------------------
// Fused attention block (attention.cu): GroupNorm + QKV projection, then scores -> softmax -> PV -> NIN_3 -> residual
// ---------------------------------------------------------------------------------------------
struct AttentionArgs {
  const float* x;            // [B][L][C] NHWC block input
  const double* qs;          // quad statistics of x
  const float* gamma; const float* beta;       // GroupNorm_0
  const float* wqkv; const float* bqkv;        // [C][3C] (columns q | k | v), [3C]
  const float* w3; const float* b3;            // NIN_3: [C][C] ([in, out]), [C]
  float* out;                // [B][L][C]
  double* qstats;            // optional quad statistics of out (zeroed buffer)
  float* scratch;            // attention_scratch_floats(B, L) floats: q, kT, v
  int B, L, C;
};
size_t attention_scratch_floats(int B, int L);
int launch_attention(const AttentionArgs& a, cudaStream_t s, std::string* err);

// ---------------------------------------------------------------------------------------------
// Batched STFT / iSTFT + amplitude compression (stft.cu); n_fft 510, hop 128, hann, center=True
// ---------------------------------------------------------------------------------------------
size_t stft_basis_floats();                       // forward basis [510][512] + inverse basis [512][512] + window [512]
void launch_stft_basis(float* basis, int sqrt_window, cudaStream_t s);
int stft_frames(int L);                           // 1 + L / 128
// mode: transform_type 0 exponent, 1 log, 2 none (data_module.py:149-175)
// wav [B][wav_stride] -> Y complex [B][256][Tpad]; scratch: xpad [B][xpad_stride] (xpad_stride >= Lmax + 510 + 128,
// multiple of 4), S complex [B][frames(Lmax)][256], peak_bits [B] (the peaks as fp32 bit patterns when normalize)
void launch_stft_spec(const float* basis, const float* wav, long long wav_stride, const int* lengths_dev, int B, int Lmax,
                      bool normalize, float factor, float expo, int mode, float* xpad, long long xpad_stride,
                      float2* S, unsigned* peak_bits, float2* Y, int Tpad, cudaStream_t s);
// X complex [B][256][Tpad] -> wav_out [B][wav_stride] (zeros beyond each length); scratch S as above, frames [B][T][512]
void launch_spec_istft(const float* basis, const float2* X, int Tpad, const int* lengths_dev, int B, int Lmax, float factor,
                       float expo, int mode, const float* peak, float2* S, float* frames, float* wav_out, long long wav_stride,
                       cudaStream_t s);

}  // namespace flowse
