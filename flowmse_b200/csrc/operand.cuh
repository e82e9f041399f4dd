// Device helpers shared by the operand-producing kernels (kernels_gn.cu: standalone prep; conv_halo.cu / conv_gemm.cu:
// operand transform inside the conv kernel): SiLU, the exact fp32 -> fp16 hi/lo split and the fp16-range check.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace flowse {
namespace {

// SiLU from the two MUFU approximations directly: v * rcp(1 + ex2(-v * log2 e)), 5 instructions.  (__expf / __fdividef
// wrap the same MUFU.EX2 / MUFU.RCP in denormal-range scaling - FSETP + 2 FMUL per call - which only matters when
// exp(-v) is denormal, i.e. when 1 + exp(-v) == 1 anyway.)  |rel err| ~ 2e-7 near 0, absolute error < 1e-9 in the tails.
// Together with the saturating pack below this removes ~6 of the ~25 instructions per element; end to end it measured
// within noise (22.78-22.86 vs 22.6-22.9 ms per sampler call), kept because it is the simpler code.
__device__ __forceinline__ float silu_f(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

// two fp32 -> packed fp16x2 (e0 in the low half), round-to-nearest-even, saturating to +-65504 (F2FP.SATFINITE: one
// instruction, replaces a 2-instruction clamp per element)
__device__ __forceinline__ uint32_t pack_h2_sat(float e0, float e1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}

// exact hi/lo split: v ~= hi + lo with hi, lo fp16 (saturating; lo = v - hi is exact in fp32)
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  hi.x = pack_h2_sat(v.x, v.y); hi.y = pack_h2_sat(v.z, v.w);
  const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&hi.x));
  const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&hi.y));
  lo.x = pack_h2_sat(v.x - f01.x, v.y - f01.y); lo.y = pack_h2_sat(v.z - f23.x, v.w - f23.y);
}
__device__ __forceinline__ uint4 pack8(const uint2 a, const uint2 b) { return make_uint4(a.x, a.y, b.x, b.y); }
// largest magnitude of an operand vector: anything above 65504 saturates in the fp16 hi/lo split, which the kernels
// report through the context's sticky overflow counter (flowse_fp16_overflow) instead of clipping silently
__device__ __forceinline__ float amax4(const float4 v, float m) {
  return fmaxf(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))), m);
}
constexpr float kHalfMax = 65504.0f;

// GroupNorm affine (x * sc + sh with sc = rstd*gamma, sh = beta - mean*sc) followed by an optional SiLU
__device__ __forceinline__ float4 norm_act(const float4 x, const float4 sc, const float4 sh, int silu) {
  float4 y;
  y.x = fmaf(x.x, sc.x, sh.x); y.y = fmaf(x.y, sc.y, sh.y);
  y.z = fmaf(x.z, sc.z, sh.z); y.w = fmaf(x.w, sc.w, sh.w);
  if (silu) { y.x = silu_f(y.x); y.y = silu_f(y.y); y.z = silu_f(y.z); y.w = silu_f(y.w); }
  return y;
}


}  // namespace
}  // namespace flowse
