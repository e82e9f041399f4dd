/* libflowse - C ABI of the B200-native FlowSE reverse-ODE sampling hot path.
 *
 * Every entry point replaces a piece of the reference's Python hot path (paths relative to the seongq/flowmse
 * checkout, /root/reference).  The reference has no FFI for this path (it is pure PyTorch plus one JIT-built torch
 * extension), so the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - complex64 tensors are passed as pointers to interleaved (re, im) fp32 pairs, layout [B,1,F=256,T] contiguous
 *     (the reference's NCHW layout with C=1), T % 64 == 0 (flowmse/util/other.py:83-90).
 *   - All data pointers are DEVICE pointers owned by the caller unless a parameter says "host".
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden synchronisation
 *     (except flowse_load_weights and the first call for a new (B,T), which allocate).
 *   - Return value 0 = ok; otherwise an error code, message via flowse_last_error().  There is no CPU fallback.
 *   - One context per (process, device); a context is not re-entrant.
 */
#ifndef FLOWSE_H
#define FLOWSE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flowse_ctx flowse_ctx;

/* One tensor of the backbone state_dict inside a flat fp32 blob (reference names, e.g.
 * "all_modules.4.Conv_0.weight"; flowmse/backbones/ncsnpp.py:99-245). */
typedef struct {
  char name[96];
  long long offset; /* in floats from the start of the blob */
  long long numel;
} flowse_tensor_desc;

enum { FLOWSE_SOLVER_EULER = 0, FLOWSE_SOLVER_HEUN = 1, FLOWSE_SOLVER_MIDPOINT = 2 };

/* Context life cycle.  `device` is a CUDA ordinal. */
int flowse_create(flowse_ctx** out, int device);
void flowse_destroy(flowse_ctx* ctx);
/* Last error message of `ctx` (or of the failed flowse_create when ctx == NULL). */
const char* flowse_last_error(const flowse_ctx* ctx);

/* Load the NCSN++ weights (EMA weights, as VFModel.eval() would select: flowmse/model.py:92-106) from a HOST blob.
 * Replaces NCSNpp.__init__ + load_state_dict (flowmse/backbones/ncsnpp.py:45-245).  Packs the conv weights into the
 * K-major fp16 hi/lo format of the tcgen05 kernels and uploads them; synchronous. */
int flowse_load_weights(flowse_ctx* ctx, const float* host_blob, const flowse_tensor_desc* descs, int n);

/* Packed weights (the on-disk format either side of the path).  After flowse_load_weights the context
 * holds the conv weights in the K-major fp16 hi/lo layout of the tcgen05 kernels (same bytes as fp32: 2 x 2 B) and the
 * small tensors in fp32.  flowse_export_packed copies that state into a relocatable HOST blob of flowse_packed_bytes()
 * bytes (header "FLSEPK01", segment sizes, per-conv power-of-two scales, 256-byte aligned segments);
 * flowse_load_packed rebuilds a fresh context from such a blob without the fp32 checkpoint (torch_ema shadow_params,
 * flowmse/model.py:81-106) and without the host-side packing pass.  Outputs are bit-identical to the fp32-loaded
 * context.  A blob written by another library version (other segment layout) is rejected. */
size_t flowse_packed_bytes(flowse_ctx* ctx);
int flowse_export_packed(flowse_ctx* ctx, void* host_out, size_t bytes);
int flowse_load_packed(flowse_ctx* ctx, const void* host_blob, size_t bytes);

/* Bytes of device workspace the context holds for batch B, T frames (allocates the plan if needed). */
size_t flowse_workspace_bytes(flowse_ctx* ctx, int B, int T);

/* x = y + sigma * z.  Replaces FLOWMATCHING.prior_sampling (flowmse/odes.py:93-100) with z drawn by the caller
 * (torch.randn_like on y's device keeps "same seed" semantics).  n = number of complex elements. */
int flowse_prior_sample(flowse_ctx* ctx, const void* y, const void* z, float sigma, void* x, long long n, void* stream);

/* out = NCSNpp.forward(cat([x, y], 1), t)  (flowmse/backbones/ncsnpp.py:247-404); negate != 0 gives
 * VFModel.forward = -dnn(...) (flowmse/model.py:164-170).
 * x, y: complex [B,1,256,T] with batch strides x_bstride / y_bstride in complex elements (so a [B,2,256,T] dnn
 * input is passed as x = base, y = base + 256*T, strides 2*256*T).  t: fp32 [B] device.  out: complex [B,1,256,T]. */
int flowse_ncsnpp_forward(flowse_ctx* ctx, const void* x, long long x_bstride, const void* y, long long y_bstride,
                          const float* t, void* out, int negate, int B, int T, void* stream);

/* x_out = x + v * (-stepsize).  Replaces EulerODEsolver.update_fn's arithmetic
 * (flowmse/sampling/odesolvers.py:42-47) for a caller-supplied vector field v; n complex elements. */
int flowse_euler_step(flowse_ctx* ctx, const void* x, const void* v, float stepsize, void* x_out, long long n,
                      void* stream);

/* The whole reverse-ODE sampler: prior sample + N solver steps with the NCSN++ vector field, entirely on device.
 * Replaces get_white_box_solver(...)() (flowmse/sampling/__init__.py:27-62).
 *   y, z      complex [B,1,256,T] (noisy spectrogram = conditioning, prior noise)
 *   y_prior   mean of the prior sample x_T = y_prior + sigma z; NULL means y (evaluate.py:118 passes Y twice)
 *   timesteps HOST fp32 [N] = torch.linspace(T_rev, t_eps, N); step sizes are their fp32 differences, last = t[N-1]
 *   solver    FLOWSE_SOLVER_*; Heun/midpoint use an Euler step for the last interval (the reference would evaluate
 *             the network at t = 0, i.e. log 0)
 *   x_out     complex [B,1,256,T] */
int flowse_sample(flowse_ctx* ctx, const void* y, const void* y_prior, const void* z, const float* timesteps_host, int N, int solver,
                  float sigma, void* x_out, int B, int T, void* stream);

/* Building block of the adaptive black-box solver (replaces the NumPy arithmetic scipy's solve_ivp RK45 does on the host
 * for get_black_box_solver, flowmse/sampling/__init__.py:64-114, with a device<->host round trip per network evaluation):
 *   v = base + sum_{s<S} coef[s] * K[s]     base / v complex128 [n] (NULL base = 0), K complex64 [S][k_stride] (exact in fp64)
 * out64 (complex128) / out32 (complex64, round-to-nearest) receive v when not NULL.  When sumsq_host != NULL the call also
 * returns sum_i |v_i / (atol + rtol * max(|ya_i|, |yb_i|))|^2 (scipy's scaled error norm is sqrt(that / n)) and
 * synchronises the stream - the single host round trip of an RK step.  coef_host: HOST doubles, S <= 8. */
int flowse_rk_lincomb(flowse_ctx* ctx, const void* base64, const void* K32, long long k_stride, const double* coef_host, int S,
                      void* out64, void* out32, const void* ya64, const void* yb64, double rtol, double atol,
                      double* sumsq_host, long long n, void* stream);

/* Sticky fp16-range flag.  Conv operands travel as fp16 hi/lo pairs (5 exponent bits): a value with |v| > 65504 - only
 * possible for the un-normalised shortcut operand of a ResBlock (layerspp.py:268-270) with pathological weights - would
 * saturate.  The operand-producing kernels count such values; *count receives the number of detections since the last
 * reset (synchronises the device).  Nothing in the reference corresponds to this: its convolutions are fp32. */
int flowse_fp16_overflow(flowse_ctx* ctx, long long* count, int reset);

/* Options: "conv_impl" 0 = tcgen05 (default: halo kernel on high-resolution layers, per-tap kernel otherwise),
 * 1 = SIMT cross-check, 2 = per-tap kernel everywhere, 3 = like 0 with 3 rotating main accumulators in the halo kernel,
 * 4 = like 0 with the CTA-pair (cta_group::2) halo kernel; "graph" 0/1 = replay each NFE as a CUDA graph; "pdl" = programmatic dependent launch (process-wide):
 * 0 off, 1 every kernel, 2 (default) the low-resolution conv launches only - the fastest under graph replay, 3 = 2 + every
 * launch of at most 148 CTAs, 4 = 2 + the halo conv launches (both measured neutral);
 * "whole_graph" 0/1 (default 1) = flowse_sample replays prior + all evaluations + updates as ONE graph from the second call
 * with a given schedule; "fuse_prep" = who prepares the conv operands (GroupNorm + SiLU + fp16 hi/lo split) of the
 * high-resolution layers: 0 a standalone pass per conv, 1 (default) the halo conv kernel itself;
 * "fork" 0/1 (default 1) = under graph capture the pyramid heads and the input-pyramid FIR chain become parallel graph
 * branches (nothing on the main chain needs them before the final kernel);
 * "spec_transform" = SpecsDataModule.transform_type used by flowse_stft_spec / flowse_spec_istft: 0 "exponent" (default),
 * 1 "log" (log(1+|X|) e^{j angle} * factor and its inverse), 2 "none" (flowmse/data_module.py:149-175);
 * "stft_window" = get_window (flowmse/data_module.py:13-19): 0 "hann" (default), 1 "sqrthann". */
int flowse_set_option(flowse_ctx* ctx, const char* key, int value);

/* Number of this library's kernel launches (graph kernel nodes included) since the context was created. */
long long flowse_kernel_launches(const flowse_ctx* ctx);

/* Measurement hook: run ONE network evaluation of the current plan op by op on a private stream with a CUDA event
 * after every op (the evaluation is queued behind a spin kernel so the kernels run back to back).  Per op: kind (0 misc,
 * 1 gn_stats, 2 gn_prep, 3 conv_gemm = per-tap conv kernel, 4 attention, 5 small, 6 temb, 7 conv_halo = halo conv kernel),
 * elapsed ms, algorithmic FLOPs and info = {H, W, K, Cout} (convs and pyramid heads only). */
int flowse_profile_forward(flowse_ctx* ctx, int max_ops, int* kinds, float* ms, double* flops, int* info, int* n_ops);

/* Test hook: device pointer / shape (NHWC fp32) of the output of all_modules[module_idx] from the last forward of
 * the current (B,T) plan. */
int flowse_debug_tap(flowse_ctx* ctx, int module_idx, const float** ptr, int* C, int* H, int* W);

/* Test hook: synchronous device-to-device copy (pairs with flowse_debug_tap). */
int flowse_debug_copy(flowse_ctx* ctx, const void* src, void* dst, size_t bytes);

/* ---- the step either side of the sampler: batched STFT / iSTFT + amplitude compression ---- */

/* wav -> model-domain spectrogram for a ragged batch, no host synchronisation.  Replaces, per utterance,
 *   y / y.abs().max()                                   (evaluate.py:109-110, when normalize != 0)
 *   VFModel._stft = torch.stft(n_fft 510, hop 128, hann periodic, center=True)   (flowmse/data_module.py:163-170)
 *   VFModel._forward_transform = |X|^e exp(j angle X) * spec_factor              (flowmse/data_module.py:149-162)
 *   pad_spec: zero-pad the frame axis                                            (flowmse/util/other.py:83-90)
 * wav: DEVICE fp32 [B][wav_stride]; lengths_host: HOST int [B] samples per utterance (>= 256);
 * Y: DEVICE complex [B,1,256,Tpad], Tpad >= 1 + max(len)/128 (frames beyond an utterance's own count are zero);
 * peak_out: DEVICE fp32 [B], receives max|wav| per utterance (required when normalize != 0). */
int flowse_stft_spec(flowse_ctx* ctx, const float* wav, long long wav_stride, const int* lengths_host, int B, int normalize,
                     float spec_factor, float abs_exponent, void* Y, int Tpad, float* peak_out, void* stream);

/* Model-domain spectrogram -> wav.  Replaces VFModel.to_audio = istft(_backward_transform(spec), length)
 * (flowmse/model.py:190-203, data_module.py:164-175) and the `* norm_factor` of evaluate.py:134-135 (peak may be NULL).
 * X: DEVICE complex [B,1,256,Tpad]; wav_out: DEVICE fp32 [B][wav_stride], zero beyond each utterance's length. */
int flowse_spec_istft(flowse_ctx* ctx, const void* X, int Tpad, const int* lengths_host, int B, float spec_factor,
                      float abs_exponent, const float* peak, float* wav_out, long long wav_stride, void* stream);

/* ---- op-level entry points (used by the parity tests; same kernels as the path above) ---- */

/* Pack fp32 conv weights [Cout][Cin][kh*kw] (+ optional 1x1 shortcut [Cout][Cin2]) from HOST memory into a DEVICE
 * buffer of 2*Npad*K halves (hi plane then lo plane, K = ntaps*Cin + Cin2); returns the scale exponent in *wexp. */
int flowse_pack_conv_weights(const float* w_main_host, int Cout, int Cin, int ntaps, const float* w_sc_host, int Cin2,
                             int Npad, void* dev_out, int* wexp);

/* GroupNorm(32 groups, eps 1e-6) [+SiLU] [+FIR up/down x2] of the channel-concat of src1/src2 (NHWC fp32), written as
 * fp16 hi/lo operands (outA: activated, outX: raw input) and/or fp32 (outF).  mode 0 plain, 1 down, 2 up. */
int flowse_op_gn_prep(flowse_ctx* ctx, const float* src1, int C1, const float* src2, int C2, const float* gamma,
                      const float* beta, int B, int H, int W, int mode, int silu, void* outA, void* outX, float* outF,
                      float* outXF, void* stream);

/* Implicit-GEMM conv (3x3 pad 1 when ntaps == 9, 1x1 when 1) on hi/lo operands; impl 0 = per-tap tcgen05 kernel,
 * 1 = SIMT cross-check, 2 / 3 = halo tcgen05 kernel (3x3, H % 16 == 0, W % 8 == 0) with 1 / 3 main accumulators,
 * 4 = halo kernel on CTA pairs (cta_group::2). */
int flowse_op_conv_gemm(flowse_ctx* ctx, const void* A, int Cin, int ntaps, const void* X, int Cin2, const void* Wp,
                        int Npad, int wexp, const float* bias, int bias_bstride, const float* residual, int div_sqrt2,
                        float* out, int Cout, int ldc, int B, int H, int W, int impl, void* stream);

/* Pyramid head (ncsnpp.py:347-366): out = FIR-up(prev) + conv3x3(C -> 4)(SiLU(GroupNorm(h))) + bias on NHWC fp32 h
 * [B,H,W,C]; wf: DEVICE fp32 [9][C][4] (tap-major); prev: [B,H/2,W/2,4] or NULL; out: [B,H,W,4]. */
int flowse_op_head_conv(flowse_ctx* ctx, const float* h, const float* gamma, const float* beta, const float* wf,
                        const float* bias, const void* prev, void* out, int B, int H, int W, int C, void* stream);

/* Single-head spatial self-attention block (AttnBlockpp, layerspp.py:62-91) on NHWC fp32 x [B,H,W,256]. */
int flowse_op_attention(flowse_ctx* ctx, int module_idx, const float* x, float* out, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWSE_H */
