// The per-element arithmetic of GroupNorm32 (+FiLM) (+SiLU) on an fp16 activation, shared by
// gn_apply_kernel (unet_ops.cu) and the transform warps of conv_halo_kernel (conv_tc.cu) so both
// produce the same bits.  Reference rounding points: GroupNorm32 computes in fp32 and casts back
// (nn.py:17-19), h * (1 + scale) + shift and SiLU are fp16 tensor ops (unet.py:248-252).
#pragma once
#include <cuda_fp16.h>

namespace pdr {

__device__ __forceinline__ float gnm_round_h(float x) { return __half2float(__float2half_rn(x)); }
// SiLU whose result is immediately rounded to fp16: ex2.approx / rcp.approx (abs error ~1e-6)
// (.ftz forms: 5 instructions - FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL - without the denormal-range
// fix-ups nvcc adds around the non-ftz ex2/rcp; 1 + 2^t never is denormal and x is an fp16 value)
__device__ __forceinline__ float gnm_silu_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// x: fp16 value as float; ga/gb: GroupNorm affine (rstd*gamma, beta - mean*rstd*gamma);
// fs/fsh: FiLM (1+scale) and shift, both fp16 values as float
__device__ __forceinline__ float gn_apply_one(float x, float ga, float gb, float fs, float fsh,
                                              bool film, bool silu) {
  float t = gnm_round_h(x * ga + gb);
  if (film) t = gnm_round_h(gnm_round_h(t * fs) + fsh);
  if (silu) t = gnm_round_h(gnm_silu_fast(t));
  return t;
}

// Two elements at once with packed conversions (same results as gn_apply_one: the fp16 product
// t*fs is exact in fp32, so __hmul2's single rounding equals round_h(fp32 product); the add stays
// an fp32 add + rounding like the reference's fp16 tensor add).
template <bool FILM, bool SILU = true>
__device__ __forceinline__ __half2 gn_apply_two(__half2 x, float2 ga, float2 gb, __half2 fs,
                                                float2 fsh) {
  const float2 xf = __half22float2(x);
  __half2 t = __floats2half2_rn(xf.x * ga.x + gb.x, xf.y * ga.y + gb.y);
  if (FILM) {
    const float2 m = __half22float2(__hmul2(t, fs));
    t = __floats2half2_rn(m.x + fsh.x, m.y + fsh.y);
  }
  if (!SILU) return t;
  const float2 tf = __half22float2(t);
  return __floats2half2_rn(gnm_silu_fast(tf.x), gnm_silu_fast(tf.y));
}

}  // namespace pdr
