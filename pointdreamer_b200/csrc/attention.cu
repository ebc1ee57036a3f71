// Self-attention of the ADM U-Net's AttentionBlock (QKVAttentionLegacy).
//
// Reference: models/DDNM/guided_diffusion/unet.py:299-305, 337-354.  qkv comes from the 1x1
// conv as [B, T, 3C] (NHWC) with the legacy head-major channel order
// c = head*(3*64) + {q:0..63, k:64..127, v:128..191}.  Rounding points mirror the reference's
// fp16 tensors: q*scale and k*scale (scale = 64^-1/4) are fp16, the logits are an fp16 tensor,
// softmax runs in fp32 and is cast to fp16, the weighted sum is an fp16 tensor.
// Attention is 0.5 % of the U-Net's FLOPs: this kernel favours exactness over
// speed — two passes over the keys (row max / sum, then normalised probabilities x V) with
// mma.sync m16n8k16 tiles; no T x T matrix is ever written to HBM.
#include "common.cuh"
#include "unet_ops.h"

namespace pdr {

static constexpr int DH = 64;       // head dim
static constexpr int QT = 64;       // queries per CTA
static constexpr int KT = 64;       // keys per tile
static constexpr int LDS = 72;      // padded smem row (halfs)

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4],
                                          const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *(uint32_t*)&h;
}

// load a [64 x 64] tile (rows = tokens) of q, k or v into smem, optionally scaled (fp16 result)
__device__ __forceinline__ void load_tile(const __half* __restrict__ src, int row_stride,
                                          float scale, bool do_scale, __half* dst /*[64][LDS]*/) {
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c8 = (i & 7) * 8;
    uint4 v = __ldg((const uint4*)(src + (size_t)r * row_stride + c8));
    if (do_scale) {
      __half* h = (__half*)&v;
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __float2half_rn(__half2float(h[j]) * scale);
    }
    *(uint4*)(dst + r * LDS + c8) = v;
  }
}
__device__ __forceinline__ void load_tile_transposed(const __half* __restrict__ src,
                                                     int row_stride, __half* dst /*[64 d][LDS]*/) {
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c8 = (i & 7) * 8;  // r = key, c8 = first of 8 d's
    const uint4 v = __ldg((const uint4*)(src + (size_t)r * row_stride + c8));
    const __half* h = (const __half*)&v;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[(c8 + j) * LDS + r] = h[j];
  }
}

// S = Q K^T for this warp's 16 query rows against the 64 keys in sK; result rounded to fp16
__device__ __forceinline__ void compute_scores(const __half* sQ, const __half* sK, int warp,
                                               int lane, float (&s)[8][4]) {
  const int r0 = warp * 16 + (lane >> 2);
  const int cq = (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = *(const uint32_t*)(sQ + r0 * LDS + ks * 16 + cq);
    a[1] = *(const uint32_t*)(sQ + (r0 + 8) * LDS + ks * 16 + cq);
    a[2] = *(const uint32_t*)(sQ + r0 * LDS + ks * 16 + 8 + cq);
    a[3] = *(const uint32_t*)(sQ + (r0 + 8) * LDS + ks * 16 + 8 + cq);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t b[2];
      const int kr = nt * 8 + (lane >> 2);
      b[0] = *(const uint32_t*)(sK + kr * LDS + ks * 16 + cq);
      b[1] = *(const uint32_t*)(sK + kr * LDS + ks * 16 + 8 + cq);
      mma_16816(s[nt], a, b);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[nt][j] = round_h(s[nt][j]);  // logits are an fp16 tensor
}

__global__ void __launch_bounds__(128)
attention_kernel(const __half* __restrict__ qkv, int T, int heads, __half* __restrict__ out) {
  __shared__ __align__(16) __half sQ[QT * LDS];
  __shared__ __align__(16) __half sK[KT * LDS];
  __shared__ __align__(16) __half sVt[DH * LDS];
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int C3 = heads * 3 * DH, C = heads * DH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = 0.35355339059327373f;  // 1/sqrt(sqrt(64))
  const __half* base = qkv + (size_t)b * T * C3 + head * 3 * DH;
  load_tile(base + (size_t)qt * QT * C3, C3, scale, true, sQ);

  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float s[8][4];
  // ---- pass 1: row max and sum of exp ----
  for (int kt = 0; kt < T / KT; ++kt) {
    __syncthreads();
    load_tile(base + (size_t)kt * KT * C3 + DH, C3, scale, true, sK);
    __syncthreads();
    compute_scores(sQ, sK, warp, lane, s);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float tm = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) tm = fmaxf(tm, fmaxf(s[nt][2 * h], s[nt][2 * h + 1]));
      tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
      tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 2));
      const float mn = fmaxf(m[h], tm);
      float ts = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) ts += expf(s[nt][2 * h] - mn) + expf(s[nt][2 * h + 1] - mn);
      ts += __shfl_xor_sync(0xffffffffu, ts, 1);
      ts += __shfl_xor_sync(0xffffffffu, ts, 2);
      l[h] = l[h] * expf(m[h] - mn) + ts;
      m[h] = mn;
    }
  }
  // ---- pass 2: P = softmax (fp16), O = P V ----
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[nt][j] = 0.f;
  for (int kt = 0; kt < T / KT; ++kt) {
    __syncthreads();
    load_tile(base + (size_t)kt * KT * C3 + DH, C3, scale, true, sK);
    load_tile_transposed(base + (size_t)kt * KT * C3 + 2 * DH, C3, sVt);
    __syncthreads();
    compute_scores(sQ, sK, warp, lane, s);
    uint32_t p[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      p[nt][0] = pack_h2(expf(s[nt][0] - m[0]) / l[0], expf(s[nt][1] - m[0]) / l[0]);
      p[nt][1] = pack_h2(expf(s[nt][2] - m[1]) / l[1], expf(s[nt][3] - m[1]) / l[1]);
    }
    const int cq = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4] = {p[2 * ks][0], p[2 * ks][1], p[2 * ks + 1][0], p[2 * ks + 1][1]};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bb[2];
        const int dr = nt * 8 + (lane >> 2);
        bb[0] = *(const uint32_t*)(sVt + dr * LDS + ks * 16 + cq);
        bb[1] = *(const uint32_t*)(sVt + dr * LDS + ks * 16 + 8 + cq);
        mma_16816(o[nt], a, bb);
      }
    }
  }
  // ---- write O (fp16) ----
  const int r0 = qt * QT + warp * 16 + (lane >> 2);
  __half* ob = out + (size_t)b * T * C + head * DH;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int d = nt * 8 + (lane & 3) * 2;
    *(uint32_t*)(ob + (size_t)r0 * C + d) = pack_h2(o[nt][0], o[nt][1]);
    *(uint32_t*)(ob + (size_t)(r0 + 8) * C + d) = pack_h2(o[nt][2], o[nt][3]);
  }
}

int attention_launch(const __half* qkv, int B, int T, int heads, __half* out,
                     cudaStream_t stream) {
  PDR_CHECK_ARG(T % 64 == 0 && T >= 64, "attention: sequence length %d must be a multiple of 64", T);
  PDR_CHECK_ARG(heads >= 1 && B >= 1, "attention: bad shape");
  attention_kernel<<<dim3(T / QT, heads, B), 128, 0, stream>>>(qkv, T, heads, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
