// Shared device helpers of the geometry kernels (project / splat / fill / unproject).
//
// These translation units are compiled with -fmad=false: every fp32 expression below is a
// sequence of single IEEE-754 operations in the order documented in oracle/camera.py and
// oracle/project.py, so integer results (pixels, masks, ids) match the oracle bit for bit.
#pragma once
#include "common.cuh"

namespace pdr {

static constexpr int CAM_PARAM_FLOATS = 16;  // r00..r22, t0..t2, f, za, zb, pad
static constexpr int MAX_VIEWS = 32;

// world -> NDC.  Restates kaolin Camera.transform (call sites ours_utils.py:99,
// unproject.py:241) with the canonical op order of oracle/camera.py:transform.
__device__ __forceinline__ void cam_transform(const float* __restrict__ p, float x, float y,
                                              float z, float& nx, float& ny, float& nz) {
  const float cx = ((p[0] * x + p[1] * y) + p[2] * z) + p[9];
  const float cy = ((p[3] * x + p[4] * y) + p[5] * z) + p[10];
  const float cz = ((p[6] * x + p[7] * y) + p[8] * z) + p[11];
  const float d = -cz;
  nx = (cx * p[12]) / d;
  ny = (cy * p[12]) / d;
  nz = p[13] - p[14] / d;
}

// order-preserving float <-> int key (for atomicMin/Max on floats)
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF);
}
__device__ __forceinline__ unsigned int float_to_ordered_u32(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_u32_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// torch `.clip(lo, hi)` on floats: min(max(x, lo), hi) (NaN propagates like torch)
__device__ __forceinline__ float clipf(float x, float lo, float hi) {
  return fminf(fmaxf(x, lo), hi);
}
__device__ __forceinline__ long long clipll(long long x, long long lo, long long hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

// ---- rasteriser conventions shared by rasterize / interpolate (oracle/project.py:rasterize) ----
static constexpr int SUBPIX = 256;  // vertex xy snapped to a 1/256-pixel grid

__device__ __forceinline__ long long snap_coord(float ndc, int res) {
  const float s = ((ndc + 1.0f) * 0.5f) * (float)res;
  return (long long)floorf(s * (float)SUBPIX + 0.5f);
}
__device__ __forceinline__ bool edge_inclusive(long long dx, long long dy) {
  return (dy > 0) || (dy == 0 && dx < 0);
}
__device__ __forceinline__ long long floordiv(long long a, long long b) {
  long long q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

}  // namespace pdr
