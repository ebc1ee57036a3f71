// UNPROJECT stage kernels (K11-K13): texel -> view projection + depth visibility, Non-Border-First
// (NBF) border removal on bit-packed view masks, view selection, colour gather, compaction.
//
// Reference: pointdreamer/unproject.py:201-425 (unproject), :429-475
// (get_shrinked_per_view_per_pixel_visibility_torch), utils/utils_2d.py:799-827 (Scharr),
// :833-845 (dilate_torch_batch).  The Scharr thresholds reduce to "gx != 0 or gy != 0" in exact
// integer arithmetic and reflect-pad + max-pool equals a clamped-window OR (oracle/unproject.py).
// All V views of a texel live in the bits of one 32-bit word, so Scharr, dilation and the
// candidate logic handle every view at once.
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

struct UnprojConst {
  float cams[MAX_VIEWS * CAM_PARAM_FLOATS];
  float base_dirs[MAX_VIEWS * 3];
};

// uv of a texel in view v: (uv - c)/s [* scale_factor] * pad_mul + 0.5   (unproject.py:247-262)
__device__ __forceinline__ void texel_uv(float nx, float ny, float cx, float cy, float sc,
                                         float pad_mul, int rescale, float sf, float& u_ns,
                                         float& v_ns, float& u_s, float& v_s) {
  if (rescale) {
    const float a = (nx - cx) / sc, b = (ny - cy) / sc;
    u_ns = a * pad_mul + 0.5f;
    v_ns = b * pad_mul + 0.5f;
    u_s = (a * sf) * pad_mul + 0.5f;
    v_s = (b * sf) * pad_mul + 0.5f;
  } else {
    u_ns = u_s = nx * 0.5f + 0.5f;
    v_ns = v_s = ny * 0.5f + 0.5f;
  }
}

// u1: per-texel visibility bits
__global__ void unproj_visibility_kernel(const float* __restrict__ cams,
                                         const float* __restrict__ gb_pos,
                                         const uint8_t* __restrict__ mask, int R, int V,
                                         const float* __restrict__ centers,
                                         const float* __restrict__ scales, float pad_mul,
                                         int rescale, const float* __restrict__ mesh_depths,
                                         int cam_res, float offset,
                                         unsigned int* __restrict__ vis_bits) {
  __shared__ float sc[MAX_VIEWS * CAM_PARAM_FLOATS];
  for (int i = threadIdx.x; i < V * CAM_PARAM_FLOATS; i += blockDim.x) sc[i] = cams[i];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * R) return;
  unsigned int bits = 0;
  if (mask[i]) {
    const float x = gb_pos[3 * i], y = gb_pos[3 * i + 1], z = gb_pos[3 * i + 2];
    for (int v = 0; v < V; ++v) {
      float nx, ny, nz, u_ns, v_ns, u_s, v_s;
      cam_transform(sc + v * CAM_PARAM_FLOATS, x, y, z, nx, ny, nz);
      // rescale == 0 (unproject.py:262-264): the crop arrays may be NULL and are not read
      texel_uv(nx, ny, rescale ? centers[2 * v] : 0.f, rescale ? centers[2 * v + 1] : 0.f,
               rescale ? scales[v] : 2.f, pad_mul, rescale, 1.0f, u_ns, v_ns, u_s, v_s);
      const float fc = (float)cam_res;
      const long long col = (long long)clipf(u_ns * fc, 0.f, (float)(cam_res - 1));
      const long long row = (long long)clipf(v_ns * fc, 0.f, (float)(cam_res - 1));
      const float ref = mesh_depths[((size_t)v * cam_res + row) * cam_res + col];
      if (nz - ref <= offset) bits |= 1u << v;
    }
  }
  vis_bits[i] = bits;
}

// u2: edges.  bit v of the result: Scharr(vis_v) != 0 and Scharr(chart mask) == 0
__global__ void unproj_edges_kernel(const unsigned int* __restrict__ vis_bits,
                                    const uint8_t* __restrict__ mask, int R, int V,
                                    unsigned int* __restrict__ edge_bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * R) return;
  const int y = i / R, x = i % R;
  unsigned int nb[3][3];
  int mk[3][3];
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      const bool in = yy >= 0 && yy < R && xx >= 0 && xx < R;
      nb[dy + 1][dx + 1] = in ? vis_bits[(size_t)yy * R + xx] : 0u;  // zero padding
      mk[dy + 1][dx + 1] = in ? (mask[(size_t)yy * R + xx] ? 1 : 0) : 0;
    }
  const int cgx = 3 * (mk[0][2] - mk[0][0]) + 10 * (mk[1][2] - mk[1][0]) + 3 * (mk[2][2] - mk[2][0]);
  const int cgy = 3 * (mk[2][0] - mk[0][0]) + 10 * (mk[2][1] - mk[0][1]) + 3 * (mk[2][2] - mk[0][2]);
  unsigned int out = 0;
  if (cgx == 0 && cgy == 0) {
    for (int v = 0; v < V; ++v) {
#define BIT(a, b) ((int)((nb[a][b] >> v) & 1u))
      const int gx = 3 * (BIT(0, 2) - BIT(0, 0)) + 10 * (BIT(1, 2) - BIT(1, 0)) +
                     3 * (BIT(2, 2) - BIT(2, 0));
      const int gy = 3 * (BIT(2, 0) - BIT(0, 0)) + 10 * (BIT(2, 1) - BIT(0, 1)) +
                     3 * (BIT(2, 2) - BIT(0, 2));
#undef BIT
      if (gx != 0 || gy != 0) out |= 1u << v;
    }
  }
  edge_bits[i] = out;
}

// u3/u4: separable clamped-window OR (all views at once)
__global__ void unproj_dilate_rows_kernel(const unsigned int* __restrict__ in, int R, int p,
                                          unsigned int* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * R) return;
  const int y = i / R, x = i % R;
  unsigned int acc = 0;
  for (int xx = max(x - p, 0); xx <= min(x + p, R - 1); ++xx) acc |= in[(size_t)y * R + xx];
  out[i] = acc;
}
__global__ void unproj_dilate_cols_kernel(const unsigned int* __restrict__ in,
                                          const unsigned int* __restrict__ vis_bits, int R, int p,
                                          unsigned int* __restrict__ shrinked) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * R) return;
  const int y = i / R, x = i % R;
  unsigned int acc = 0;
  for (int yy = max(y - p, 0); yy <= min(y + p, R - 1); ++yy) acc |= in[(size_t)yy * R + x];
  shrinked[i] = vis_bits[i] & ~acc;
}

// u5: candidate views, softmax-weighted argmax, colour gather (unproject.py:298-400)
__global__ void unproj_select_kernel(const float* __restrict__ cams,
                                     const float* __restrict__ base_dirs,
                                     const float* __restrict__ gb_pos,
                                     const uint8_t* __restrict__ mask,
                                     const long long* __restrict__ face_id,
                                     const float* __restrict__ f_normals, int F, int R, int V,
                                     const float* __restrict__ centers,
                                     const float* __restrict__ scales, float pad_mul, int rescale,
                                     const float* __restrict__ scale_factors,
                                     const unsigned int* __restrict__ vis_bits,
                                     const unsigned int* __restrict__ shrinked, int K,
                                     int n_levels, int complete_unseen,
                                     const float* __restrict__ images, int res,
                                     float* __restrict__ atlas, int* __restrict__ view_id_dense,
                                     uint8_t* __restrict__ painted) {
  __shared__ float sc[MAX_VIEWS * CAM_PARAM_FLOATS];
  __shared__ float sb[MAX_VIEWS * 3];
  for (int i = threadIdx.x; i < V * CAM_PARAM_FLOATS; i += blockDim.x) sc[i] = cams[i];
  for (int i = threadIdx.x; i < V * 3; i += blockDim.x) sb[i] = base_dirs[i];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t RR = (size_t)R * R;
  if (i >= RR) return;
  float r = 0.f, g = 0.f, b = 0.f;
  int vid = -1000;  // outside the chart mask
  uint8_t pt = 0;
  if (mask[i]) {
    // candidates: level 0, then looser levels only while no view is left (unproject.py:324-356)
    unsigned int cand = shrinked[i];
    for (int l = 1; l < n_levels; ++l)
      if (cand == 0) cand |= shrinked[(size_t)l * RR + i];
    if (complete_unseen && cand == 0) cand |= vis_bits[i];
    (void)K;
    long long f = face_id[i];
    if (f < 0) f += F;  // python negative indexing (unproject.py:298)
    const float nx_ = f_normals[3 * f], ny_ = f_normals[3 * f + 1], nz_ = f_normals[3 * f + 2];
    float sim[MAX_VIEWS];
    float mx = -INFINITY;
    for (int v = 0; v < V; ++v) {
      sim[v] = (nx_ * sb[3 * v] + ny_ * sb[3 * v + 1]) + nz_ * sb[3 * v + 2];
      mx = fmaxf(mx, sim[v]);
    }
    float sum = 0.f;
    for (int v = 0; v < V; ++v) {
      // correctly rounded fp32 exp (float64 evaluation, one rounding): the canonical rule of
      // oracle/unproject.py:softmax_rows, independent of the library's expf
      sim[v] = (float)exp((double)(sim[v] - mx));
      sum += sim[v];
    }
    float best = -INFINITY;
    int arg = 0;
    for (int v = 0; v < V; ++v) {
      const float w = ((cand >> v) & 1u) ? sim[v] / sum : -100.0f;
      if (w > best) best = w, arg = v;  // first maximum
    }
    vid = arg;
    if (!complete_unseen && cand == 0) vid = -100;
    if (vid >= 0) {
      const float x = gb_pos[3 * i], y = gb_pos[3 * i + 1], z = gb_pos[3 * i + 2];
      float nx, ny, nz, u_ns, v_ns, u_s, v_s;
      cam_transform(sc + vid * CAM_PARAM_FLOATS, x, y, z, nx, ny, nz);
      texel_uv(nx, ny, rescale ? centers[2 * vid] : 0.f, rescale ? centers[2 * vid + 1] : 0.f,
               rescale ? scales[vid] : 2.f, pad_mul, rescale, rescale ? scale_factors[vid] : 1.f,
               u_ns, v_ns, u_s, v_s);
      const float fr = (float)res;
      const long long col = (long long)clipf(u_s * fr, 0.f, (float)(res - 1));
      const long long row = (long long)clipf(v_s * fr, 0.f, (float)(res - 1));
      const size_t o = ((size_t)vid * 3 * res + (res - 1 - row)) * res + col;  // flipped rows
      const size_t cs = (size_t)res * res;
      r = images[o];
      g = images[o + cs];
      b = images[o + 2 * cs];
      pt = 1;
    }
  }
  atlas[3 * i] = r;
  atlas[3 * i + 1] = g;
  atlas[3 * i + 2] = b;
  view_id_dense[i] = vid;
  painted[i] = pt;
}

// expand the bits of the returned shrinked visibility to [V,R,R] uint8
__global__ void unproj_expand_kernel(const unsigned int* __restrict__ bits, int R, int V,
                                     uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t RR = (size_t)R * R;
  if (i >= RR) return;
  const unsigned int b = bits[i];
  for (int v = 0; v < V; ++v) out[(size_t)v * RR + i] = (b >> v) & 1u;
}

// ---- compaction of the masked texels (row-major order == torch boolean indexing order) ----
static constexpr int SCAN_BLOCK = 1024;

__global__ void scan_block_kernel(const uint8_t* __restrict__ mask, size_t n,
                                  int* __restrict__ local, int* __restrict__ block_sums) {
  __shared__ int s[SCAN_BLOCK];
  const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int val = (i < n && mask[i]) ? 1 : 0;
  s[threadIdx.x] = val;
  __syncthreads();
  for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
    int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) local[i] = s[threadIdx.x] - val;  // exclusive
  if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void scan_sums_kernel(int* block_sums, int nb, int* total) {
  // single block: sequential chunks of SCAN_BLOCK
  __shared__ int s[SCAN_BLOCK];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const int val = i < nb ? block_sums[i] : 0;
    s[threadIdx.x] = val;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
      int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_sums[i] = carry + s[threadIdx.x] - val;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void compact_texels_kernel(const uint8_t* __restrict__ mask,
                                      const int* __restrict__ local,
                                      const int* __restrict__ block_sums,
                                      const float* __restrict__ gb_pos,
                                      const int* __restrict__ view_id_dense, int R,
                                      long long* __restrict__ view_ids,
                                      long long* __restrict__ coords, float* __restrict__ points) {
  const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (i >= (size_t)R * R || !mask[i]) return;
  const size_t k = (size_t)block_sums[blockIdx.x] + local[i];
  if (view_ids) view_ids[k] = view_id_dense[i];
  if (coords) {
    coords[2 * k] = i / R;
    coords[2 * k + 1] = i % R;
  }
  if (points) {
    points[3 * k] = gb_pos[3 * i];
    points[3 * k + 1] = gb_pos[3 * i + 1];
    points[3 * k + 2] = gb_pos[3 * i + 2];
  }
}

__global__ void mask_count_kernel(const uint8_t* __restrict__ mask, size_t n, int* out) {
  int c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    c += mask[i] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int mask_count_sync(const uint8_t* mask, size_t n, int* ws_counter, int* out_host,
                    cudaStream_t stream) {
  PDR_CUDA(cudaMemsetAsync(ws_counter, 0, sizeof(int), stream));
  mask_count_kernel<<<148 * 4, 256, 0, stream>>>(mask, n, ws_counter);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  PDR_CUDA(cudaMemcpyAsync(out_host, ws_counter, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PDR_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

size_t unproject_workspace_bytes(int R, int n_levels) {
  const size_t RR = (size_t)R * R;
  const size_t nb = (RR + SCAN_BLOCK - 1) / SCAN_BLOCK;
  // vis, edges, tmp, shrinked[n_levels], view_id_dense, scan local, block sums
  return RR * 4 * (size_t)(3 + n_levels + 2) + nb * 4 + 1024;
}

int unproject_launch(const float* images, int res, const float* cams, int V, int cam_res,
                     const float* base_dirs, const float* gb_pos, const uint8_t* mask,
                     const long long* face_id, int R, const float* f_normals, int F,
                     const float* uv_centers, const float* uv_scales, double padding, int rescale,
                     const float* scale_factors, const float* mesh_depths, const int* kernels_host,
                     int n_levels, int n_kernels_total, int complete_unseen, void* workspace,
                     float* atlas, uint8_t* shrinked_vis, long long* point_view_ids,
                     long long* point_coords, float* points, uint8_t* painted,
                     cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && V <= MAX_VIEWS, "view count %d out of range", V);
  PDR_CHECK_ARG(n_levels >= 1 && n_levels <= 16, "edge_dilate_kernels must have 1..16 entries");
  PDR_CHECK_ARG(R > 0 && res > 0 && F > 0, "empty input");
  (void)n_kernels_total;
  const size_t RR = (size_t)R * R;
  const bool nbf_off = kernels_host[0] == 0;  // unproject.py:436-437
  PDR_CHECK_ARG(!nbf_off || n_levels == 1,
                "edge_dilate_kernels=[0, ...] with more than one level indexes past the single "
                "unshrunk level in the reference (IndexError)");
  for (int l = 0; l < n_levels && !nbf_off; ++l)
    PDR_CHECK_ARG(kernels_host[l] % 2 == 1, "dilation kernel %d must be odd", kernels_host[l]);
  uint8_t* w = (uint8_t*)workspace;
  unsigned int* vis = (unsigned int*)w;
  w += RR * 4;
  unsigned int* edges = (unsigned int*)w;
  w += RR * 4;
  unsigned int* tmp = (unsigned int*)w;
  w += RR * 4;
  unsigned int* shr = (unsigned int*)w;
  w += RR * 4 * n_levels;
  int* vid_dense = (int*)w;
  w += RR * 4;
  int* scan_local = (int*)w;
  w += RR * 4;
  int* block_sums = (int*)w;
  const int nb = cdiv(RR, SCAN_BLOCK);
  const float pad_mul = (float)(1.0 - 2.0 * padding);
  const int grid = cdiv(RR, 256);

  unproj_visibility_kernel<<<grid, 256, 0, stream>>>(cams, gb_pos, mask, R, V, uv_centers,
                                                     uv_scales, pad_mul, rescale, mesh_depths,
                                                     cam_res, 0.0001f, vis);
  PDR_COUNT_LAUNCH();
  if (nbf_off) {
    PDR_CUDA(cudaMemcpyAsync(shr, vis, RR * 4, cudaMemcpyDeviceToDevice, stream));
  } else {
    unproj_edges_kernel<<<grid, 256, 0, stream>>>(vis, mask, R, V, edges);
    PDR_COUNT_LAUNCH();
    for (int l = 0; l < n_levels; ++l) {
      const int p = (kernels_host[l] - 1) / 2;
      unproj_dilate_rows_kernel<<<grid, 256, 0, stream>>>(edges, R, p, tmp);
      PDR_COUNT_LAUNCH();
      unproj_dilate_cols_kernel<<<grid, 256, 0, stream>>>(tmp, vis, R, p, shr + (size_t)l * RR);
      PDR_COUNT_LAUNCH();
    }
  }
  unproj_select_kernel<<<grid, 256, 0, stream>>>(
      cams, base_dirs, gb_pos, mask, face_id, f_normals, F, R, V, uv_centers, uv_scales, pad_mul,
      rescale, scale_factors, vis, shr, n_levels, n_levels, complete_unseen, images, res, atlas,
      vid_dense, painted);
  PDR_COUNT_LAUNCH();
  if (shrinked_vis) {
    // the reference returns the LAST level it looked at (unproject.py:324,338,425)
    unproj_expand_kernel<<<grid, 256, 0, stream>>>(shr + (size_t)(n_levels - 1) * RR, R, V,
                                                   shrinked_vis);
    PDR_COUNT_LAUNCH();
  }
  if (point_view_ids || point_coords || points) {
    scan_block_kernel<<<nb, SCAN_BLOCK, 0, stream>>>(mask, RR, scan_local, block_sums);
    PDR_COUNT_LAUNCH();
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, stream>>>(block_sums, nb, nullptr);
    PDR_COUNT_LAUNCH();
    compact_texels_kernel<<<nb, SCAN_BLOCK, 0, stream>>>(mask, scan_local, block_sums, gb_pos,
                                                         vid_dense, R, point_view_ids,
                                                         point_coords, points);
    PDR_COUNT_LAUNCH();
  }
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
