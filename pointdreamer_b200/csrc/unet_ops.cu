// Memory-bound glue kernels of the ADM U-Net (everything that is not a tensor-core GEMM):
// time embedding MLP + per-ResBlock emb_layers (fp32), stem conv 3->C, GroupNorm32 statistics
// and apply (+SiLU, +FiLM scale/shift, +AvgPool2/nearest-up2 resampling, two-source concat),
// plain resampling, and the fp32 output head (GN + SiLU + conv C->6).
//
// Reference: models/DDNM/guided_diffusion/unet.py (ResBlock._forward 236-256,
// UNetModel.forward 635-664, Upsample 92-110, Downsample 113-140), nn.py (GroupNorm32 17-19,
// timestep_embedding 103-121).  Activations are NHWC fp16; rounding to fp16 happens at exactly
// the points where the reference materialises an fp16 tensor (see oracle/unet.py).
#include "common.cuh"
#include "unet_ops.h"
#include "gn_math.cuh"

namespace pdr {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// SiLU whose result is immediately rounded to fp16: ex2.approx / rcp.approx (abs error ~1e-6) is far
// below the fp16 quantum, and the memory-bound GroupNorm pass stays memory-bound.
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

// ------------------------------------------------------------ linear (fp32) ----
// out[b][n] = bias[n] + sum_k f(in[b][k]) * W[n][k];  f = identity | SiLU | timestep embedding.
// One warp per output feature, all (<= 8) batch rows at once; in[] staged in shared memory.
// mode_in: 0 identity, 1 SiLU(in), 2 in = timestep_embedding(t[b], K)   (nn.py:103-121)
// mode_out: 0 fp32, 1 fp32 SiLU(out)... only 0 used; out16 != null additionally stores fp16.
__global__ void linear_kernel(const float* __restrict__ in, const float* __restrict__ W,
                              const float* __restrict__ bias, int B, int K, int N, int mode_in,
                              float* __restrict__ out, __half* __restrict__ out16,
                              const int* __restrict__ skip) {
  extern __shared__ float s_in[];  // [bchunk<=8][K]
  if (skip && *skip >= 0) return;  // embedding cache hit: the result is already known
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = min(8, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * K; i += blockDim.x) {
      const int b = i / K, k = i - b * K;
      float v;
      if (mode_in == 2) {
        const int half = K / 2;
        const int j = k < half ? k : k - half;
        const float freq = expf((-9.210340371976184f * (float)j) / (float)half);
        const float a = in[b0 + b] * freq;
        v = k < half ? cosf(a) : sinf(a);
      } else {
        v = in[(size_t)(b0 + b) * K + k];
        if (mode_in == 1) v = silu_f(v);
      }
      s_in[i] = v;
    }
    __syncthreads();
    // a warp owns LF consecutive output features at a time: LF independent 16-byte weight
    // loads in flight per step and each staged input vector is read once for LF features
    constexpr int LF = 4;
    for (int n0 = (blockIdx.x * warps + warp) * LF; n0 < N; n0 += gridDim.x * warps * LF) {
      float acc[LF][8];
#pragma unroll
      for (int f = 0; f < LF; ++f)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[f][b] = 0.f;
      const float* wr[LF];
#pragma unroll
      for (int f = 0; f < LF; ++f) wr[f] = W + (size_t)min(n0 + f, N - 1) * K;
      for (int k = lane * 4; k < K; k += 128) {
        float4 w4[LF];
#pragma unroll
        for (int f = 0; f < LF; ++f) w4[f] = __ldg((const float4*)(wr[f] + k));
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (b < nb) {
            const float4 s = *(const float4*)(s_in + b * K + k);
#pragma unroll
            for (int f = 0; f < LF; ++f)
              acc[f][b] += w4[f].x * s.x + w4[f].y * s.y + w4[f].z * s.z + w4[f].w * s.w;
          }
        }
      }
#pragma unroll
      for (int f = 0; f < LF; ++f)
#pragma unroll
        for (int b = 0; b < 8; ++b)
          for (int o = 16; o > 0; o >>= 1)
            acc[f][b] += __shfl_xor_sync(0xffffffffu, acc[f][b], o);
      if (lane == 0) {
#pragma unroll
        for (int f = 0; f < LF; ++f) {
          const int n = n0 + f;
          if (n >= N) break;
          const float bv = bias ? bias[n] : 0.f;
          for (int b = 0; b < nb; ++b) {
            const float r = acc[f][b] + bv;
            if (out) out[(size_t)(b0 + b) * N + n] = r;
            if (out16) out16[(size_t)(b0 + b) * N + n] = __float2half_rn(r);
          }
        }
      }
    }
  }
}

int linear_launch(const float* in, const float* W, const float* bias, int B, int K, int N,
                  int mode_in, float* out, __half* out16, cudaStream_t stream, const int* skip) {
  PDR_CHECK_ARG(K % 4 == 0 && K <= 4096, "linear: K=%d must be a multiple of 4 and <= 4096", K);
  const int threads = 256;
  const int nbk = B < 8 ? B : 8;
  const size_t smem = (size_t)nbk * K * sizeof(float);
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  8 * 4096 * 4));
    configured = true;
  }
  int grid = cdiv(N, (threads / 32) * 4);
  if (grid > 148 * 8) grid = 148 * 8;
  linear_kernel<<<grid, threads, smem, stream>>>(in, W, bias, B, K, N, mode_in, out, out16, skip);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------- timestep-embedding cache ----
__global__ void __launch_bounds__(32)
emb_cache_lookup_kernel(const float* __restrict__ t, int B, EmbCacheMeta* __restrict__ m, int enabled) {
  const int lane = threadIdx.x;
  const unsigned t0 = __float_as_uint(t[0]);
  bool same = true;
  for (int b = lane; b < B; b += 32) same = same && __float_as_uint(t[b]) == t0;
  same = __all_sync(0xffffffffu, same) && enabled != 0;
  int hit = -1;
  const int n = m->n;
  if (same)
    for (int k = lane; k < n; k += 32)
      if (__float_as_uint(m->t[k]) == t0) hit = k;
  for (int o = 16; o > 0; o >>= 1) hit = max(hit, __shfl_xor_sync(0xffffffffu, hit, o));
  if (lane == 0) {
    m->hit = hit;
    m->store = (hit < 0 && same && n < EMB_CACHE_SLOTS) ? n : -1;
  }
}

__global__ void emb_cache_finish_kernel(__half* __restrict__ emb16, int B, int etot,
                                        __half* __restrict__ cache, EmbCacheMeta* __restrict__ m,
                                        const float* __restrict__ t) {
  const int hit = m->hit, store = m->store;  // written by the lookup launch only
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte piece of a row
  const int pieces = etot / 8;
  if (hit >= 0) {
    if (i < pieces) {
      const uint4 v = reinterpret_cast<const uint4*>(cache + (size_t)hit * etot)[i];
      for (int b = 0; b < B; ++b) reinterpret_cast<uint4*>(emb16 + (size_t)b * etot)[i] = v;
    }
  } else if (store >= 0) {
    if (i < pieces)
      reinterpret_cast<uint4*>(cache + (size_t)store * etot)[i] = reinterpret_cast<const uint4*>(emb16)[i];
    if (i == 0) {  // visible to the next lookup launch (stream order)
      m->t[store] = t[0];
      m->n = store + 1;
    }
  }
}

int emb_cache_lookup_launch(const float* t, int B, EmbCacheMeta* meta, int enabled,
                            cudaStream_t stream) {
  emb_cache_lookup_kernel<<<1, 32, 0, stream>>>(t, B, meta, enabled);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int emb_cache_finish_launch(__half* emb16, int B, int etot, __half* cache, EmbCacheMeta* meta,
                            const float* t, cudaStream_t stream) {
  PDR_CHECK_ARG(etot % 8 == 0, "embedding cache: row length %d must be a multiple of 8", etot);
  emb_cache_finish_kernel<<<cdiv(etot / 8, 256), 256, 0, stream>>>(emb16, B, etot, cache, meta, t);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- stem conv ----
// x [B,3,H,W] fp32 NCHW -> fp16 (unet.py:655) -> conv3x3(3->C) -> NHWC fp16.
// A warp produces one pixel's C-vector at a time (lane = 8-channel chunk): coalesced 16-B stores.
__global__ void stem_conv_kernel(const float* __restrict__ x, const __half* __restrict__ w,
                                 const float* __restrict__ bias, int B, int H, int W, int C,
                                 __half* __restrict__ out) {
  extern __shared__ float s_w[];  // [27][C]
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    s_w[i] = h2f(w[(size_t)c * 27 + k]);  // w is [C][tap*3 + cin]
  }
  __syncthreads();
  const int chunks = C / 8;
  const int lane_chunk = threadIdx.x % chunks;
  const int pix_in_block = threadIdx.x / chunks;
  const int pix_per_block = blockDim.x / chunks;
  const size_t npix = (size_t)B * H * W;
  for (size_t p = (size_t)blockIdx.x * pix_per_block + pix_in_block; p < npix;
       p += (size_t)gridDim.x * pix_per_block) {
    const int xw = p % W, yh = (p / W) % H, b = p / ((size_t)W * H);
    float in[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = yh + ky - 1, xx = xw + kx - 1;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          in[(ky * 3 + kx) * 3 + c] =
              ok ? round_h(__ldg(x + (((size_t)b * 3 + c) * H + yy) * W + xx)) : 0.f;
      }
    float acc[8];
    const int c0 = lane_chunk * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4 w0 = *(const float4*)(s_w + k * C + c0);
      const float4 w1 = *(const float4*)(s_w + k * C + c0 + 4);
      acc[0] += in[k] * w0.x, acc[1] += in[k] * w0.y, acc[2] += in[k] * w0.z, acc[3] += in[k] * w0.w;
      acc[4] += in[k] * w1.x, acc[5] += in[k] * w1.y, acc[6] += in[k] * w1.z, acc[7] += in[k] * w1.w;
    }
    __align__(16) __half o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(acc[j] + bias[c0 + j]);
    *(uint4*)(out + p * C + c0) = *(const uint4*)o;
  }
}

int stem_conv_launch(const float* x, const __half* w, const float* bias, int B, int H, int W,
                     int C, __half* out, cudaStream_t stream) {
  PDR_CHECK_ARG(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "stem: unsupported C=%d", C);
  const size_t smem = (size_t)27 * C * sizeof(float);
  PDR_CHECK_ARG(smem <= 200 * 1024, "stem: C too large");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PDR_CUDA(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = smem;
  }
  const int ppb = 256 / (C / 8);
  long long blocks = cdiv((long long)B * H * W, ppb);
  if (blocks > 148 * 16) blocks = 148 * 16;
  stem_conv_kernel<<<(int)blocks, 256, smem, stream>>>(x, w, bias, B, H, W, C, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// stem as a tensor-core GEMM: x [B,3,H,W] fp32 -> fp16 patches [B,H,W,64] (k = (ky*3+kx)*3 + c for
// k < 27, zero above) consumed by conv_tc as a 1x1 conv with a [C][64] zero-padded weight
__global__ void stem_im2col_kernel(const float* __restrict__ x, int B, int H, int W,
                                   __half* __restrict__ out) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)B * H * W) return;
  const int xw = p % W, yh = (p / W) % H, b = p / ((size_t)W * H);
  __align__(16) __half v[64];
#pragma unroll
  for (int k = 27; k < 64; ++k) v[k] = __float2half_rn(0.f);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = yh + ky - 1, xx = xw + kx - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[(ky * 3 + kx) * 3 + c] =
            __float2half_rn(ok ? __ldg(x + (((size_t)b * 3 + c) * H + yy) * W + xx) : 0.f);
    }
  uint4* o = (uint4*)(out + p * 64);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = ((const uint4*)v)[j];
}

int stem_im2col_launch(const float* x, int B, int H, int W, __half* out, cudaStream_t stream) {
  stem_im2col_kernel<<<cdiv((long long)B * H * W, 128), 128, 0, stream>>>(x, B, H, W, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------ GroupNorm32 ----
// statistics: per-(batch, slab, channel) partial sum / sum of squares in fp32, finalised in
// double per (batch, group).  x = concat(x1[C1], x2[C2]) along channels, NHWC fp16.
__device__ __forceinline__ const __half* src_ptr(const __half* x1, const __half* x2, int C1,
                                                 int C2, size_t pixel, int c) {
  return c < C1 ? x1 + pixel * C1 + c : x2 + pixel * C2 + (c - C1);
}

__global__ void gn_partial_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2,
                                  int C1, int C2, int HW, int slabs,
                                  float* __restrict__ partial) {
  // grid (slabs, B); thread -> (pixel lane, 8-channel chunk)
  const int C = C1 + C2;
  const int chunks = C / 8;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int chunk = threadIdx.x % chunks;
  const int plane = threadIdx.x / chunks;
  const int planes = blockDim.x / chunks;
  const int per_slab = (HW + slabs - 1) / slabs;
  const int p0 = slab * per_slab, p1 = min(HW, p0 + per_slab);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const int c0 = chunk * 8;
  {
    const __half* src = c0 < C1 ? x1 + c0 : x2 + (c0 - C1);
    const size_t stride = c0 < C1 ? C1 : C2;
    src += (size_t)b * HW * stride;
    int p = p0 + plane;
    for (; p + 3 * planes < p1; p += 4 * planes) {  // 4 independent 16-B loads in flight
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg((const uint4*)(src + (size_t)(p + u * planes) * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half* h = (const __half*)&v[u];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float f = h2f(h[j]);
          s[j] += f;
          q[j] += f * f;
        }
      }
    }
    for (; p < p1; p += planes) {
      const uint4 v = __ldg((const uint4*)(src + (size_t)p * stride));
      const __half* h = (const __half*)&v;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = h2f(h[j]);
        s[j] += f;
        q[j] += f * f;
      }
    }
  }
  // deterministic block reduction: per-plane partials, summed in plane order
  extern __shared__ float sm[];  // [planes][2][C]
  if (plane < planes) {
    float* mine = sm + (size_t)plane * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[c0 + j] = s[j];
      mine[C + c0 + j] = q[j];
    }
  }
  __syncthreads();
  float* dst = partial + ((size_t)b * slabs + slab) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < planes; ++p) acc += sm[(size_t)p * 2 * C + i];
    dst[i] = acc;
  }
}

__global__ void gn_finalize_kernel(const float* __restrict__ partial, int C, int HW, int slabs,
                                   float eps, float* __restrict__ stats) {
  // one warp per (b, group)
  const int b = blockIdx.y, g = blockIdx.x;
  const int cpg = C / 32;
  double s = 0.0, q = 0.0;
  const int n = slabs * cpg;
  for (int i = threadIdx.x; i < n; i += 32) {
    const int slab = i / cpg, c = g * cpg + i % cpg;
    const float* src = partial + ((size_t)b * slabs + slab) * 2 * C;
    s += (double)src[c];
    q += (double)src[C + c];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (threadIdx.x == 0) {
    const double cnt = (double)HW * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((size_t)b * 32 + g) * 2 + 0] = (float)mean;
    stats[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

int gn_stats_slabs(int B, int HW) {
  // Depends on the image size only: the summation order (and so every bit of the result) of a
  // chain is the same whatever batch it runs in.
  (void)B;
  int slabs = (HW + 255) / 256;
  if (slabs > 74) slabs = 74;
  if (slabs < 1) slabs = 1;
  return slabs;
}

// Small feature maps (HW <= 256, 8-channel aligned groups): one block per (group, image) reads its
// HW x C/32 slab once and writes (mean, rstd) directly - one launch, 32*B blocks, instead of the
// two-stage slab reduction whose grid would be B blocks.  Fixed summation order.
__global__ void __launch_bounds__(256)
gn_stats_small_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, int C1, int C2,
                      int HW, float eps, float* __restrict__ stats) {
  const int g = blockIdx.x, b = blockIdx.y;
  const int C = C1 + C2, cpg = C / 32, chunks = cpg / 8;
  const int planes = 256 / chunks;
  const int chunk = threadIdx.x % chunks, plane = threadIdx.x / chunks;
  const int c0 = g * cpg + chunk * 8;
  const __half* src = c0 < C1 ? x1 + c0 : x2 + (c0 - C1);
  const size_t stride = c0 < C1 ? C1 : C2;
  src += (size_t)b * HW * stride;
  float s = 0.f, q = 0.f;
  if (plane < planes) {
    for (int p = plane; p < HW; p += planes) {
      const uint4 v = __ldg((const uint4*)(src + (size_t)p * stride));
      const __half* h = (const __half*)&v;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = h2f(h[j]);
        s += f;
        q += f * f;
      }
    }
  }
  double ds = (double)s, dq = (double)q;
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
  __shared__ double sm[8][2];
  if ((threadIdx.x & 31) == 0) {
    sm[threadIdx.x >> 5][0] = ds;
    sm[threadIdx.x >> 5][1] = dq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      ts += sm[w][0];
      tq += sm[w][1];
    }
    const double cnt = (double)HW * cpg;
    const double mean = ts / cnt;
    double var = tq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((size_t)b * 32 + g) * 2 + 0] = (float)mean;
    stats[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

int gn_stats_launch(const __half* x1, const __half* x2, int B, int HW, int C1, int C2,
                    float* ws_partial, float* stats, cudaStream_t stream) {
  const int C = C1 + C2;
  PDR_CHECK_ARG(C % 32 == 0 && C1 % 8 == 0 && C2 % 8 == 0 && C <= 4096,
                "GroupNorm32: unsupported channel count %d+%d", C1, C2);
  if (HW <= 256 && C % 256 == 0 && C / 256 <= 256) {
    gn_stats_small_kernel<<<dim3(32, B), 256, 0, stream>>>(x1, x2 ? x2 : x1, C1, C2, HW, 1e-5f,
                                                           stats);
    PDR_COUNT_LAUNCH();
    PDR_LAUNCH_CHECK();
    return 0;
  }
  const int chunks = C / 8;
  const int threads = chunks >= 256 ? chunks : 256 / chunks * chunks;
  PDR_CHECK_ARG(threads <= 1024, "GroupNorm32: too many channels");
  const int slabs = gn_stats_slabs(B, HW);
  gn_partial_kernel<<<dim3(slabs, B), threads, (size_t)(threads / chunks) * 2 * C * sizeof(float),
                      stream>>>(
      x1, x2 ? x2 : x1, C1, C2, HW, slabs, ws_partial);
  PDR_COUNT_LAUNCH();
  gn_finalize_kernel<<<dim3(32, B), 32, 0, stream>>>(ws_partial, C, HW, slabs, 1e-5f, stats);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}


// ---- GroupNorm statistics fused into the producing conv (conv_tc epilogue) ----
// reduce the epilogue's partial rows [B*R][C8][2] (fp32) to per-image per-8-channel sums
// sums8[B][C8][2] (fp64), fixed summation order.
template <int PARTS>
__global__ void __launch_bounds__(32 * PARTS)
sums8_reduce_kernel(const float* __restrict__ partial, int R, int C8,
                    double* __restrict__ sums8) {
  // grid (ceil(C8*2 / 32), B), block (32, PARTS): thread (col, part) sums a contiguous row range
  __shared__ double sm[PARTS][32];
  const int b = blockIdx.y;
  const int col = blockIdx.x * 32 + threadIdx.x;  // index into [C8][2]
  const int part = threadIdx.y;
  const int per = (R + PARTS - 1) / PARTS;
  const int r0 = part * per, r1 = min(R, r0 + per);
  double acc = 0.0;
  if (col < C8 * 2) {
    const float* p = partial + ((size_t)b * R) * (C8 * 2) + col;
    int r = r0;
    for (; r + 4 <= r1; r += 4) {  // 4 independent loads in flight, summed in row order
      const float v0 = p[(size_t)r * (C8 * 2)], v1 = p[(size_t)(r + 1) * (C8 * 2)];
      const float v2 = p[(size_t)(r + 2) * (C8 * 2)], v3 = p[(size_t)(r + 3) * (C8 * 2)];
      acc += (double)v0;
      acc += (double)v1;
      acc += (double)v2;
      acc += (double)v3;
    }
    for (; r < r1; ++r) acc += (double)p[(size_t)r * (C8 * 2)];
  }
  sm[part][threadIdx.x] = acc;
  __syncthreads();
  if (part == 0 && col < C8 * 2) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < PARTS; ++k) t += sm[k][threadIdx.x];
    sums8[(size_t)b * C8 * 2 + col] = t;
  }
}

int sums8_reduce_launch(const float* partial, int B, int R, int C, double* sums8,
                        cudaStream_t stream) {
  const int C8 = C / 8;
  // the split depends on R (the image size) only, never on the batch: a chain's bits are the
  // same whatever batch it runs in
  if (R >= 256)
    sums8_reduce_kernel<32><<<dim3(cdiv(C8 * 2, 32), B), dim3(32, 32), 0, stream>>>(partial, R, C8, sums8);
  else
    sums8_reduce_kernel<8><<<dim3(cdiv(C8 * 2, 32), B), dim3(32, 8), 0, stream>>>(partial, R, C8, sums8);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// mean / rstd of GroupNorm32 over cat(x1, x2) from the per-8-channel sums of its sources
__global__ void gn_finalize_sums_kernel(const double* __restrict__ s1, const double* __restrict__ s2,
                                        int C1, int C2, int HW, float eps,
                                        float* __restrict__ stats) {
  const int b = blockIdx.x, g = threadIdx.x;  // 32 threads = 32 groups
  const int C = C1 + C2, cpg = C / 32;
  double s = 0.0, q = 0.0;
  for (int c = g * cpg; c < (g + 1) * cpg; c += 8) {
    const double* src = c < C1 ? s1 + ((size_t)b * (C1 / 8) + c / 8) * 2
                               : s2 + ((size_t)b * (C2 / 8) + (c - C1) / 8) * 2;
    s += src[0];
    q += src[1];
  }
  const double cnt = (double)HW * cpg;
  const double mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[((size_t)b * 32 + g) * 2 + 0] = (float)mean;
  stats[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

int gn_finalize_sums_launch(const double* s1, const double* s2, int B, int HW, int C1, int C2,
                            float* stats, cudaStream_t stream) {
  PDR_CHECK_ARG((C1 + C2) % 256 == 0 && C1 % 8 == 0 && C2 % 8 == 0,
                "fused GroupNorm statistics need 8-channel aligned groups");
  gn_finalize_sums_kernel<<<B, 32, 0, stream>>>(s1, s2 ? s2 : s1, C1, C2, HW, 1e-5f, stats);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// apply: y = GN(x)*gamma+beta  [-> fp16] [ *(1+scale)+shift -> fp16 ] [ SiLU -> fp16 ]
//        [ AvgPool2 / nearest-up2 ]  -> out NHWC fp16 [B,Ho,Wo,C]
// resample: 0 none, 1 down (avg 2x2), 2 up (nearest 2x)
struct GnApplyArgs {
  const __half* x1;
  const __half* x2;
  int C1, C2, B, H, W;
  const float* stats;    // [B,32,2] (mean, rstd) or null when sums1/sums2 are given
  const double* sums1;   // per-8-channel (sum, sumsq) of x1: double[B][C1/8][2]
  const double* sums2;   // same for x2
  const float* gamma;    // [C]
  const float* beta;     // [C]
  const __half* film;    // [B, film_stride] fp16: scale at [coff + c], shift at [coff + C + c]
  int film_stride, film_off;
  int silu, resample;
  __half* out;
  __half* raw_out;       // resample != 0: also emit the resampled RAW input (ResBlock's x_upd,
                         // unet.py:241) from the same read; same arithmetic as resample_kernel
};

// thread = fixed 8-channel chunk (affine constants live in registers), loops over pixels of its
// block's slab with several 16-B loads in flight; grid = (pixel slabs, B).  The per-element math is
// gn_apply_two (gn_math.cuh): packed half2 conversions, ~9 (plain) / ~12 (FiLM) instructions per
// element instead of ~15 - the kernel is issue/MUFU-bound, not HBM-bound (2 MUFU per SiLU).
struct GnConst {
  float2 ga[4], gb[4], fsh[4];
  __half2 fs[4];
};
template <bool FILM, bool SILU>
__device__ __forceinline__ uint4 gn_transform_8(uint4 v, const GnConst& k) {
  __half2* h = (__half2*)&v;
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = gn_apply_two<FILM, SILU>(h[j], k.ga[j], k.gb[j], k.fs[j], k.fsh[j]);
  return v;
}

template <int RESAMPLE, bool FILM, bool SILU>
__global__ void __launch_bounds__(512, 2)  // <= 64 registers: 4 resident 256-thread CTAs per SM
gn_apply_kernel(const GnApplyArgs a, int pix_per_block) {
  const int C = a.C1 + a.C2;
  const int chunks = C / 8;
  const int chunk = threadIdx.x % chunks, plane = threadIdx.x / chunks;
  const int planes = blockDim.x / chunks;
  const int b = blockIdx.y;
  const int c0 = chunk * 8;
  const int cpg = C / 32;
  float mean8 = 0.f, rstd8 = 0.f;
  if (a.sums1) {
    // statistics straight from the conv epilogue's sums (cpg is a multiple of 8 here, so the
    // thread's 8 channels share one group); same arithmetic as gn_finalize_sums_kernel
    const int g = c0 / cpg;
    double s = 0.0, q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; c += 8) {
      const double* src = c < a.C1 ? a.sums1 + ((size_t)b * (a.C1 / 8) + c / 8) * 2
                                   : a.sums2 + ((size_t)b * (a.C2 / 8) + (c - a.C1) / 8) * 2;
      s += src[0];
      q += src[1];
    }
    const double cnt = (double)a.H * a.W * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    mean8 = (float)mean;
    rstd8 = (float)(1.0 / sqrt(var + 1e-5));
  }
  GnConst k;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const int g = c / cpg;
    const float mean = a.sums1 ? mean8 : a.stats[((size_t)b * 32 + g) * 2];
    const float rstd = a.sums1 ? rstd8 : a.stats[((size_t)b * 32 + g) * 2 + 1];
    const float ga = rstd * a.gamma[c];
    const float gb = a.beta[c] - mean * ga;
    float fs = 1.f, fsh = 0.f;
    if (FILM) {
      const __half* f = a.film + (size_t)b * a.film_stride + a.film_off;
      fs = round_h(1.0f + h2f(f[c]));  // (1 + scale) in fp16
      fsh = h2f(f[C + c]);
    }
    if (j & 1) {
      k.ga[j / 2].y = ga, k.gb[j / 2].y = gb, k.fsh[j / 2].y = fsh;
      k.fs[j / 2] = __halves2half2(__low2half(k.fs[j / 2]), __float2half_rn(fs));
    } else {
      k.ga[j / 2].x = ga, k.gb[j / 2].x = gb, k.fsh[j / 2].x = fsh;
      k.fs[j / 2] = __halves2half2(__float2half_rn(fs), __float2half_rn(fs));
    }
  }
  const __half* src = c0 < a.C1 ? a.x1 + c0 : a.x2 + (c0 - a.C1);
  const size_t sstride = c0 < a.C1 ? a.C1 : a.C2;
  const int H = a.H, W = a.W;
  src += (size_t)b * H * W * sstride;
  if (RESAMPLE == 0) {
    __half* dst = a.out + (size_t)b * H * W * C + c0;
    const int npix = H * W;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
    int p = p0 + plane;
    for (; p + 3 * planes < p1; p += 4 * planes) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg((const uint4*)(src + (size_t)(p + u * planes) * sstride));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *(uint4*)(dst + (size_t)(p + u * planes) * C) = gn_transform_8<FILM, SILU>(v[u], k);
    }
    for (; p < p1; p += planes) {
      const uint4 v = __ldg((const uint4*)(src + (size_t)p * sstride));
      *(uint4*)(dst + (size_t)p * C) = gn_transform_8<FILM, SILU>(v, k);
    }
  } else if (RESAMPLE == 1) {  // AvgPool2d(2) of the activated tensor: loop over OUTPUT pixels
    const int Ho = H / 2, Wo = W / 2;
    __half* dst = a.out + (size_t)b * Ho * Wo * C + c0;
    const int npix = Ho * Wo;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
    for (int p = p0 + plane; p < p1; p += planes) {
      const int yo = p / Wo, xo = p - yo * Wo;
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg((const uint4*)(src + ((size_t)(2 * yo + (u >> 1)) * W + 2 * xo + (u & 1)) * sstride));
      if (a.raw_out) {  // AvgPool2d(2) of the raw input, summed in the order of resample_kernel
        float racc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) racc[j] = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const __half* hh = (const __half*)&v[u];
#pragma unroll
          for (int j = 0; j < 8; ++j) racc[j] += h2f(hh[j]);
        }
        __align__(16) __half ro[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ro[j] = __float2half_rn(racc[j] * 0.25f);
        *(uint4*)(a.raw_out + ((size_t)b * Ho * Wo + p) * C + c0) = *(const uint4*)ro;
      }
      float2 acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 t = gn_transform_8<FILM, SILU>(v[u], k);
        const __half2* th = (const __half2*)&t;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(th[j]);
          acc[j].x += f.x, acc[j].y += f.y;
        }
      }
      uint4 o;
      __half2* oh = (__half2*)&o;
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[j].x * 0.25f, acc[j].y * 0.25f);
      *(uint4*)(dst + (size_t)p * C) = o;
    }
  } else {  // nearest x2: loop over INPUT pixels, write the 2x2 block
    const int Wo = W * 2;
    __half* dst = a.out + (size_t)b * H * 2 * Wo * C + c0;
    const int npix = H * W;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
    for (int p = p0 + plane; p < p1; p += planes) {
      const int yi = p / W, xi = p - yi * W;
      const uint4 v = __ldg((const uint4*)(src + (size_t)p * sstride));
      const uint4 o = gn_transform_8<FILM, SILU>(v, k);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *(uint4*)(dst + ((size_t)(2 * yi + (u >> 1)) * Wo + 2 * xi + (u & 1)) * C) = o;
      if (a.raw_out) {
        __half* rdst = a.raw_out + (size_t)b * H * 2 * Wo * C + c0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *(uint4*)(rdst + ((size_t)(2 * yi + (u >> 1)) * Wo + 2 * xi + (u & 1)) * C) = v;
      }
    }
  }
}

// Per-(image, channel) constants of a GroupNorm32 (+FiLM) whose application is fused into the
// consuming conv (conv_halo_kernel's transform warps): coeff[b][c] = (ga, gb, fs, fsh), the same
// values gn_apply_kernel keeps in registers.  thread = (b, c).
__global__ void gn_coeff_kernel(const GnApplyArgs a, float4* __restrict__ coeff) {
  const int C = a.C1 + a.C2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * C) return;
  const int b = i / C, c = i - b * C;
  const int cpg = C / 32, g = c / cpg;
  float mean, rstd;
  if (a.sums1) {
    double s = 0.0, q = 0.0;
    for (int cc = g * cpg; cc < (g + 1) * cpg; cc += 8) {
      const double* src = cc < a.C1 ? a.sums1 + ((size_t)b * (a.C1 / 8) + cc / 8) * 2
                                    : a.sums2 + ((size_t)b * (a.C2 / 8) + (cc - a.C1) / 8) * 2;
      s += src[0];
      q += src[1];
    }
    const double cnt = (double)a.H * a.W * cpg;
    const double m = s / cnt;
    double var = q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + 1e-5));
  } else {
    mean = a.stats[((size_t)b * 32 + g) * 2];
    rstd = a.stats[((size_t)b * 32 + g) * 2 + 1];
  }
  const float ga = rstd * a.gamma[c];
  const float gb = a.beta[c] - mean * ga;
  float fs = 1.f, fsh = 0.f;
  if (a.film) {
    const __half* f = a.film + (size_t)b * a.film_stride + a.film_off;
    fs = round_h(1.0f + h2f(f[c]));
    fsh = h2f(f[C + c]);
  }
  coeff[i] = make_float4(ga, gb, fs, fsh);
}

int gn_coeff_launch(int B, int H, int W, int C1, int C2, const float* stats, const double* sums1,
                    const double* sums2, const float* gamma, const float* beta, const __half* film,
                    int film_stride, int film_off, float4* coeff, cudaStream_t stream) {
  const int C = C1 + C2;
  PDR_CHECK_ARG(C % 32 == 0 && (stats || sums1), "gn_coeff: bad arguments");
  PDR_CHECK_ARG(!sums1 || C % 256 == 0, "gn_coeff from sums needs 8-aligned groups");
  GnApplyArgs a;
  a.x1 = a.x2 = nullptr;
  a.C1 = C1, a.C2 = C2, a.B = B, a.H = H, a.W = W;
  a.stats = stats, a.gamma = gamma, a.beta = beta;
  a.sums1 = sums1, a.sums2 = sums2 ? sums2 : sums1;
  a.film = film, a.film_stride = film_stride, a.film_off = film_off;
  a.silu = 1, a.resample = 0, a.out = nullptr;
  a.raw_out = nullptr;
  gn_coeff_kernel<<<cdiv((long long)B * C, 256), 256, 0, stream>>>(a, coeff);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int gn_apply_launch(const __half* x1, const __half* x2, int B, int H, int W, int C1, int C2,
                    const float* stats, const double* sums1, const double* sums2,
                    const float* gamma, const float* beta,
                    const __half* film, int film_stride, int film_off, int silu, int resample,
                    __half* out, cudaStream_t stream, __half* raw_out) {
  const int C = C1 + C2;
  PDR_CHECK_ARG(C % 32 == 0 && C1 % 8 == 0 && C2 % 8 == 0, "GroupNorm32 apply: bad channels");
  PDR_CHECK_ARG(raw_out == nullptr || (resample != 0 && C2 == 0),
                "GroupNorm32 apply: the raw resampled copy needs a single-source resampling call");
  PDR_CHECK_ARG(resample != 1 || (H % 2 == 0 && W % 2 == 0), "avg-pool needs even size");
  GnApplyArgs a;
  a.x1 = x1;
  a.x2 = x2 ? x2 : x1;
  a.C1 = C1, a.C2 = C2, a.B = B, a.H = H, a.W = W;
  a.stats = stats, a.gamma = gamma, a.beta = beta;
  a.sums1 = sums1, a.sums2 = sums2 ? sums2 : sums1;
  PDR_CHECK_ARG(stats || sums1, "GroupNorm32 apply: no statistics given");
  PDR_CHECK_ARG(!sums1 || (C % 256 == 0), "GroupNorm32 apply from sums needs 8-aligned groups");
  a.film = film, a.film_stride = film_stride, a.film_off = film_off;
  a.silu = silu, a.resample = resample, a.out = out;
  a.raw_out = raw_out;
  const int chunks = C / 8;
  const int threads = chunks >= 256 ? chunks : 256 / chunks * chunks;
  PDR_CHECK_ARG(threads <= 1024, "GroupNorm32 apply: too many channels");
  const int planes = threads / chunks;
  // pixels the kernel loops over: outputs for none/down, inputs for up
  const int npix = resample == 1 ? (H / 2) * (W / 2) : H * W;
  int ppb = planes * 32;  // ~32 pixels per thread
  while (ppb > planes && (long long)cdiv(npix, ppb) * B < 148 * 4) ppb /= 2;
  dim3 grid(cdiv(npix, ppb), B);
  const bool fl = film != nullptr, si = silu != 0;
#define PDR_GN_CASE(R)                                                                      \
  do {                                                                                      \
    if (fl && si) gn_apply_kernel<R, true, true><<<grid, threads, 0, stream>>>(a, ppb);     \
    else if (fl) gn_apply_kernel<R, true, false><<<grid, threads, 0, stream>>>(a, ppb);     \
    else if (si) gn_apply_kernel<R, false, true><<<grid, threads, 0, stream>>>(a, ppb);     \
    else gn_apply_kernel<R, false, false><<<grid, threads, 0, stream>>>(a, ppb);            \
  } while (0)
  if (resample == 0)
    PDR_GN_CASE(0);
  else if (resample == 1)
    PDR_GN_CASE(1);
  else
    PDR_GN_CASE(2);
#undef PDR_GN_CASE
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// plain resampling of x (x_upd in ResBlock._forward, unet.py:241): mode 1 avg-pool 2, 2 nearest x2
__global__ void resample_kernel(const __half* __restrict__ x, int B, int H, int W, int C, int mode,
                                __half* __restrict__ out) {
  const int chunks = C / 8;
  const int Ho = mode == 1 ? H / 2 : H * 2, Wo = mode == 1 ? W / 2 : W * 2;
  const size_t total = (size_t)B * Ho * Wo * chunks;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int chunk = i % chunks;
    const size_t po = i / chunks;
    const int xo = po % Wo, yo = (po / Wo) % Ho, b = po / ((size_t)Wo * Ho);
    const int c0 = chunk * 8;
    __align__(16) __half o[8];
    if (mode == 1) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          const size_t pin = ((size_t)b * H + 2 * yo + dy) * W + 2 * xo + dx;
          const uint4 v = __ldg((const uint4*)(x + pin * C + c0));
          const __half* h = (const __half*)&v;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += h2f(h[j]);
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(acc[j] * 0.25f);
      *(uint4*)(out + po * C + c0) = *(const uint4*)o;
    } else {
      const size_t pin = ((size_t)b * H + (yo >> 1)) * W + (xo >> 1);
      *(uint4*)(out + po * C + c0) = __ldg((const uint4*)(x + pin * C + c0));
    }
  }
}

int resample_launch(const __half* x, int B, int H, int W, int C, int mode, __half* out,
                    cudaStream_t stream) {
  PDR_CHECK_ARG(C % 8 == 0 && (mode == 1 || mode == 2), "resample: bad arguments");
  const int Ho = mode == 1 ? H / 2 : H * 2, Wo = mode == 1 ? W / 2 : W * 2;
  long long blocks = cdiv((long long)B * Ho * Wo * (C / 8), 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  resample_kernel<<<(int)blocks, 256, 0, stream>>>(x, B, H, W, C, mode, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// --------------------------------------------------------------- fp32 head ----
// unet.py:663-664: h.float() -> GroupNorm -> SiLU -> conv3x3(C -> n_out) in fp32, NCHW fp32 out.
// Block = 16 rows x 32 cols of output pixels, thread = 4 consecutive pixels of a row;
// channels are streamed through shared memory in chunks of HEAD_CC.  The staging pass (GroupNorm
// affine + SiLU of the halo tile) used to cost more instructions than the convolution itself:
// SiLU uses ex2.approx / rcp.approx (relative error ~1e-6, three orders of magnitude below the
// fp16 noise of the tensor it is applied to).
static constexpr int HEAD_CC = 16;
static constexpr int HEAD_TW = 32, HEAD_TH = 16;
static constexpr int HEAD_THREADS = 128;  // 16 rows x 8 groups of 4 pixels

// NO = number of output channels computed: 3 (the sampler keeps only eps, diffusion.py:529-530)
// or 6
template <int NO>
__global__ void __launch_bounds__(HEAD_THREADS)
head_kernel(const __half* __restrict__ h, const float* __restrict__ stats,
            const float* __restrict__ gamma, const float* __restrict__ beta,
            const float* __restrict__ w, const float* __restrict__ bias, int B, int H, int W,
            int C, int n_out, float* __restrict__ out, int out_channels_total) {
  // w: [n_out_total(6)][C][3][3] fp32 (PyTorch layout); only the first n_out rows are used
  constexpr int PW = HEAD_TW + 4, PH = HEAD_TH + 2;  // rows padded to 36 floats: aligned float4
  constexpr int WP = NO <= 4 ? 4 : 8;                // padded outputs per (tap, channel)
  __shared__ __align__(16) float s_act[HEAD_CC * PH * PW];
  __shared__ __align__(16) float s_w[9 * HEAD_CC * WP];  // [tap][cc][WP padded outputs]
  __shared__ float s_ab[2 * HEAD_CC];     // GroupNorm affine of the chunk's channels
  const int tiles_x = W / HEAD_TW, tiles_y = H / HEAD_TH;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int x0 = tx * HEAD_TW, y0 = ty * HEAD_TH;
  const int tid = threadIdx.x;
  const int row = tid / 8, col4 = (tid % 8) * 4;
  const int cpg = C / 32;
  float acc[4][NO];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[p][o] = 0.f;

  for (int cb = 0; cb < C; cb += HEAD_CC) {
    __syncthreads();
    if (tid < HEAD_CC) {
      const int c = cb + tid;
      const int g = c / cpg;
      const float mean = stats[((size_t)b * 32 + g) * 2], rstd = stats[((size_t)b * 32 + g) * 2 + 1];
      const float a = rstd * gamma[c];
      s_ab[tid] = a;
      s_ab[HEAD_CC + tid] = beta[c] - mean * a;
    }
    for (int i = tid; i < 9 * HEAD_CC * WP; i += HEAD_THREADS) {
      const int o = i % WP, cc = (i / WP) % HEAD_CC, tap = i / (WP * HEAD_CC);
      s_w[i] = o < n_out ? w[((size_t)o * C + cb + cc) * 9 + tap] : 0.f;
    }
    __syncthreads();
    // stage normalised + SiLU activations (zero outside the image = conv zero padding)
    for (int i = tid; i < PH * (HEAD_TW + 2) * (HEAD_CC / 8); i += HEAD_THREADS) {
      const int c8 = i % (HEAD_CC / 8);
      const int pq = i / (HEAD_CC / 8);
      const int px = pq % (HEAD_TW + 2), py = pq / (HEAD_TW + 2);
      const int pp = py * PW + px;
      const int yy = y0 + py - 1, xx = x0 + px - 1;
      float v[8];
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const uint4 raw =
            __ldg((const uint4*)(h + (((size_t)b * H + yy) * W + xx) * C + cb + c8 * 8));
        const __half* hh = (const __half*)&raw;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[j] = silu_fast(h2f(hh[j]) * s_ab[c8 * 8 + j] + s_ab[HEAD_CC + c8 * 8 + j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_act[(c8 * 8 + j) * PH * PW + pp] = v[j];
    }
    __syncthreads();
    for (int cc = 0; cc < HEAD_CC; ++cc) {
      const float* sa = s_act + cc * PH * PW;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float4 a03 = *(const float4*)(sa + (row + ky) * PW + col4);
        const float2 a45 = *(const float2*)(sa + (row + ky) * PW + col4 + 4);
        const float a6[6] = {a03.x, a03.y, a03.z, a03.w, a45.x, a45.y};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 w0 = *(const float4*)(s_w + ((ky * 3 + kx) * HEAD_CC + cc) * WP);
          float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (NO > 4) w1 = *(const float4*)(s_w + ((ky * 3 + kx) * HEAD_CC + cc) * WP + WP - 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float av = a6[p + kx];
#pragma unroll
            for (int o = 0; o < NO; ++o) acc[p][o] += av * wv[o];
          }
        }
      }
    }
  }
  const int y = y0 + row;
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    if (o >= n_out) break;
    float4 r = make_float4(acc[0][o] + bias[o], acc[1][o] + bias[o], acc[2][o] + bias[o],
                           acc[3][o] + bias[o]);
    *(float4*)(out + (((size_t)b * out_channels_total + o) * H + y) * W + x0 + col4) = r;
  }
}

int head_launch(const __half* h, const float* stats, const float* gamma, const float* beta,
                const float* w, const float* bias, int B, int H, int W, int C, int n_out,
                float* out, int out_channels_total, cudaStream_t stream) {
  PDR_CHECK_ARG(W % HEAD_TW == 0 && H % HEAD_TH == 0, "head: image %dx%d not tileable by %dx%d", H,
                W, HEAD_TH, HEAD_TW);
  PDR_CHECK_ARG(C % HEAD_CC == 0 && C % 32 == 0 && n_out >= 1 && n_out <= 6, "head: bad channels");
  const int tiles = B * (H / HEAD_TH) * (W / HEAD_TW);
  if (n_out <= 3)
    head_kernel<3><<<tiles, HEAD_THREADS, 0, stream>>>(h, stats, gamma, beta, w, bias, B, H, W, C,
                                                       n_out, out, out_channels_total);
  else
    head_kernel<6><<<tiles, HEAD_THREADS, 0, stream>>>(h, stats, gamma, beta, w, bias, B, H, W, C,
                                                       n_out, out, out_channels_total);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
