// Texture optimiser: optimize_color ("next" row N1 of SURVEY.md §8f).
//
// Reference: pointdreamer/ours_utils.py:1583-1785 - 100 Adam iterations (lr 5e-2, StepLR 15/0.5)
// on the atlas so that its bilinear renders (kaolin texture_mapping == F.grid_sample
// align_corners=False, padding 'border', y reversed, computed in float64) match the inpainted
// views under an L1 loss.  The texture coordinates per view pixel never change, so instead of
// autograd's scatter (atomicAdd in grid_sample's backward, nondeterministic) the backward pass is
// a GATHER over a texel-major list of (pixel, corner) contributions built once:
//   prepare : per view pixel -> active flag (foreground & shrinked visibility), resized target
//             colour, and up to four sort keys (texel << 32 | pixel*4 + corner)
//   [host: sort keys, count]   build : per sorted contribution -> pixel id, bilinear weight (fp64),
//             segment-head flag       [host: nonzero(head) -> segment starts]
//   forward : per active pixel -> sign(clamp(render) - target) per channel (0 where the clamp
//             blocks the gradient)
//   step    : per touched texel -> fp64 sum of weight*sign/N per view, cast to fp32 and summed
//             over views (autograd's repeat / .double() backward), then the Adam update.
// Every sum has a fixed order: the result is deterministic (the reference's is not).
#include "common.cuh"
#include "geom.h"

namespace pdr {

struct Bilerp {
  int x0, y0;        // north-west texel
  double w[4];       // nw, ne, sw, se
  bool ok[4];        // corner inside the atlas
};

// kaolin texture_mapping (uv*2-1, y negated) + grid_sample unnormalise / border clip / weights,
// all in float64 in the order of ATen's grid_sampler_2d CUDA kernel
__device__ __forceinline__ Bilerp bilerp_setup(float u, float v, int R) {
  const double gx = (double)u * 2.0 - 1.0;
  const double gy = -((double)v * 2.0 - 1.0);
  double ix = ((gx + 1.0) * (double)R - 1.0) / 2.0;
  double iy = ((gy + 1.0) * (double)R - 1.0) / 2.0;
  ix = fmin((double)(R - 1), fmax(ix, 0.0));
  iy = fmin((double)(R - 1), fmax(iy, 0.0));
  const double fx = floor(ix), fy = floor(iy);
  Bilerp b;
  b.x0 = (int)fx;
  b.y0 = (int)fy;
  const double ix_nw = fx, iy_nw = fy, ix_ne = fx + 1.0, iy_ne = fy, ix_sw = fx, iy_sw = fy + 1.0,
               ix_se = fx + 1.0, iy_se = fy + 1.0;
  b.w[0] = (ix_se - ix) * (iy_se - iy);
  b.w[1] = (ix - ix_sw) * (iy_sw - iy);
  b.w[2] = (ix_ne - ix) * (iy - iy_ne);
  b.w[3] = (ix - ix_nw) * (iy - iy_nw);
  const bool x1 = b.x0 + 1 < R, y1 = b.y0 + 1 < R;
  b.ok[0] = true;
  b.ok[1] = x1;
  b.ok[2] = y1;
  b.ok[3] = x1 && y1;
  return b;
}
__device__ __forceinline__ int corner_texel(const Bilerp& b, int k, int R) {
  return (b.y0 + (k >> 1)) * R + b.x0 + (k & 1);
}

// torch upsample_bilinear2d, align_corners=False (transforms.Resize at ours_utils.py:1750)
__device__ __forceinline__ void resize_src(int dst, float scale, int in_size, int& i0, int& step,
                                           float& l0, float& l1) {
  float r = scale * ((float)dst + 0.5f) - 0.5f;
  r = r < 0.f ? 0.f : r;
  i0 = (int)r;
  if (i0 > in_size - 1) i0 = in_size - 1;
  step = i0 < in_size - 1 ? 1 : 0;
  l1 = r - (float)i0;
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  l0 = 1.f - l1;
}

__global__ void texopt_prepare_kernel(const float* __restrict__ uv_map,
                                      const uint8_t* __restrict__ mask,
                                      const uint8_t* __restrict__ vis,
                                      const float* __restrict__ inpainted, int r0, int V, int res,
                                      int R, uint8_t* __restrict__ active,
                                      float* __restrict__ target, long long* __restrict__ keys) {
  const size_t n = (size_t)V * res * res;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = i % res, y = (i / res) % res, v = i / ((size_t)res * res);
  const float2 uv = reinterpret_cast<const float2*>(uv_map)[i];
  bool act = mask[i] != 0;
  if (act && vis) {
    // shrinked visibility looked up at long(uv * R) clipped (ours_utils.py:1713-1716, 1728-1733)
    long long tx = (long long)(uv.x * (float)R), ty = (long long)(uv.y * (float)R);
    tx = tx < 0 ? 0 : (tx > R - 1 ? R - 1 : tx);
    ty = ty < 0 ? 0 : (ty > R - 1 ? R - 1 : ty);
    act = vis[((size_t)v * R + ty) * R + tx] != 0;
  }
  active[i] = act ? 1 : 0;
  float t[3] = {0.f, 0.f, 0.f};
  if (act) {
    const float scale = (float)r0 / (float)res;
    int h1, hp, w1, wp;
    float h0l, h1l, w0l, w1l;
    resize_src(y, scale, r0, h1, hp, h0l, h1l);
    resize_src(x, scale, r0, w1, wp, w0l, w1l);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* img = inpainted + ((size_t)v * 3 + c) * r0 * r0;
      const float a = img[(size_t)h1 * r0 + w1], b = img[(size_t)h1 * r0 + w1 + wp];
      const float cc = img[(size_t)(h1 + hp) * r0 + w1], d = img[(size_t)(h1 + hp) * r0 + w1 + wp];
      t[c] = h0l * (w0l * a + w1l * b) + h1l * (w0l * cc + w1l * d);
    }
  }
  target[i * 3 + 0] = t[0];
  target[i * 3 + 1] = t[1];
  target[i * 3 + 2] = t[2];
  long long k4[4] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MAX};
  if (act) {
    const Bilerp b = bilerp_setup(uv.x, uv.y, R);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (b.ok[k]) k4[k] = ((long long)corner_texel(b, k, R) << 32) | (long long)(i * 4 + k);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) keys[i * 4 + k] = k4[k];
}

__global__ void texopt_build_kernel(const long long* __restrict__ keys, long long n_valid,
                                    const float* __restrict__ uv_map, int R,
                                    unsigned int* __restrict__ entry_pix,
                                    double* __restrict__ entry_w, uint8_t* __restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  const long long key = keys[i];
  const unsigned int e = (unsigned int)(key & 0xFFFFFFFFll);
  const unsigned int pix = e >> 2;
  const int k = e & 3;
  const float2 uv = reinterpret_cast<const float2*>(uv_map)[pix];
  const Bilerp b = bilerp_setup(uv.x, uv.y, R);
  entry_pix[i] = pix;
  entry_w[i] = b.w[k];
  head[i] = (i == 0 || (keys[i - 1] >> 32) != (key >> 32)) ? 1 : 0;
}

// signs: char4 per view pixel (x,y,z = channel signs; w unused)
__global__ void texopt_forward_kernel(const float* __restrict__ atlas,
                                      const float* __restrict__ uv_map,
                                      const uint8_t* __restrict__ active,
                                      const float* __restrict__ target, int V, int res, int R,
                                      char4* __restrict__ signs, double* __restrict__ images) {
  const size_t n = (size_t)V * res * res;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool act = active[i] != 0;
  double img[3] = {0.0, 0.0, 0.0};
  if (act) {
    const float2 uv = reinterpret_cast<const float2*>(uv_map)[i];
    const Bilerp b = bilerp_setup(uv.x, uv.y, R);
    const size_t RR = (size_t)R * R;
    signed char s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* ch = atlas + c * RR;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (b.ok[k]) acc += (double)ch[corner_texel(b, k, R)] * b.w[k];
      const bool pass = acc >= 0.0 && acc <= 1.0;  // clamp(0,1) lets the gradient through
      const double cl = fmin(fmax(acc, 0.0), 1.0);
      const double d = cl - (double)target[i * 3 + c];
      s[c] = pass ? (d > 0.0 ? 1 : (d < 0.0 ? -1 : 0)) : 0;
      img[c] = cl;
    }
    signs[i] = make_char4(s[0], s[1], s[2], 0);
  }
  if (images) {
    const size_t px = i % ((size_t)res * res), v = i / ((size_t)res * res);
#pragma unroll
    for (int c = 0; c < 3; ++c) images[(v * 3 + c) * (size_t)res * res + px] = img[c];
  }
}

struct AdamArgs {
  float lerp_w;         // 1 - beta1
  float beta2;          // beta2
  float one_m_beta2;    // 1 - beta2
  float bc2_sqrt;       // sqrt(1 - beta2^t)
  float eps;
  float neg_step_size;  // -(lr / (1 - beta1^t))
};

__global__ void texopt_step_kernel(float* __restrict__ atlas, float* __restrict__ m_buf,
                                   float* __restrict__ v_buf, const long long* __restrict__ keys,
                                   const long long* __restrict__ seg_start, long long n_seg,
                                   const unsigned int* __restrict__ entry_pix,
                                   const double* __restrict__ entry_w,
                                   const char4* __restrict__ signs, unsigned int pix_per_view,
                                   int R, double inv_n, AdamArgs a) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const long long i0 = seg_start[s], i1 = seg_start[s + 1];
  const long long t = keys[i0] >> 32;
  double acc[3] = {0.0, 0.0, 0.0};
  float g[3] = {0.f, 0.f, 0.f};
  unsigned int cur_view = entry_pix[i0] / pix_per_view;
  for (long long i = i0; i < i1; ++i) {
    const unsigned int pix = entry_pix[i];
    const unsigned int view = pix / pix_per_view;
    if (view != cur_view) {  // .double() backward casts each view's gradient to fp32,
      g[0] += (float)acc[0];  // repeat() backward sums the views in fp32
      g[1] += (float)acc[1];
      g[2] += (float)acc[2];
      acc[0] = acc[1] = acc[2] = 0.0;
      cur_view = view;
    }
    const double w = entry_w[i];
    const char4 sg = signs[pix];
    acc[0] += ((double)sg.x * inv_n) * w;
    acc[1] += ((double)sg.y * inv_n) * w;
    acc[2] += ((double)sg.z * inv_n) * w;
  }
  g[0] += (float)acc[0];
  g[1] += (float)acc[1];
  g[2] += (float)acc[2];
  const size_t RR = (size_t)R * R;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t o = c * RR + (size_t)t;
    float m = m_buf[o], v = v_buf[o];
    m = m + a.lerp_w * (g[c] - m);
    v = v * a.beta2;
    v = v + (a.one_m_beta2 * g[c]) * g[c];
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    atlas[o] = atlas[o] + a.neg_step_size * (m / denom);
    m_buf[o] = m;
    v_buf[o] = v;
  }
}

int texopt_prepare_launch(const float* uv_map, const uint8_t* mask, const uint8_t* vis,
                          const float* inpainted, int r0, int V, int res, int R, uint8_t* active,
                          float* target, long long* keys, cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && res > 0 && R > 0 && r0 > 0, "texopt_prepare: bad sizes");
  PDR_CHECK_ARG((size_t)V * res * res * 4 <= 0xFFFFFFFFull, "texopt: V*res*res*4 must fit in 32 bits");
  const size_t n = (size_t)V * res * res;
  texopt_prepare_kernel<<<cdiv(n, 256), 256, 0, stream>>>(uv_map, mask, vis, inpainted, r0, V, res,
                                                         R, active, target, keys);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int texopt_build_launch(const long long* sorted_keys, long long n_valid, const float* uv_map, int R,
                        unsigned int* entry_pix, double* entry_w, uint8_t* head,
                        cudaStream_t stream) {
  if (n_valid <= 0) return 0;
  texopt_build_kernel<<<cdiv(n_valid, 256), 256, 0, stream>>>(sorted_keys, n_valid, uv_map, R,
                                                             entry_pix, entry_w, head);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int texopt_forward_launch(const float* atlas, const float* uv_map, const uint8_t* active,
                          const float* target, int V, int res, int R, signed char* signs,
                          double* images, cudaStream_t stream) {
  const size_t n = (size_t)V * res * res;
  texopt_forward_kernel<<<cdiv(n, 256), 256, 0, stream>>>(atlas, uv_map, active, target, V, res, R,
                                                         (char4*)signs, images);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int texopt_step_launch(float* atlas, float* m, float* v, const long long* sorted_keys,
                       const long long* seg_start, long long n_seg, const unsigned int* entry_pix,
                       const double* entry_w, const signed char* signs, int V, int res, int R,
                       float lerp_w, float beta2, float one_m_beta2, float bc2_sqrt, float eps,
                       float neg_step_size, cudaStream_t stream) {
  if (n_seg <= 0) return 0;
  AdamArgs a{lerp_w, beta2, one_m_beta2, bc2_sqrt, eps, neg_step_size};
  const double inv_n = 1.0 / ((double)V * 3.0 * (double)res * (double)res);
  texopt_step_kernel<<<cdiv(n_seg, 128), 128, 0, stream>>>(
      atlas, m, v, sorted_keys, seg_start, n_seg, entry_pix, entry_w, (const char4*)signs,
      (unsigned int)res * (unsigned int)res, R, inv_n, a);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
