// Sparse-image construction (K6/K7): deterministic point splat, inner-edge mask, nearest valid
// point for edge pixels, hard_mask0 / hard_mask2, vertical flip.
//
// Reference: pointdreamer/ours_utils.py:456-495 (paint_pixels), :497-532
// (get_forground_inner_edge_mask 'dilate'), :954-1044 (get_one_sparse_img), :848-882
// (get_sparse_images).  Canonical rules: duplicate splat winner = highest
// point index; kaolin sided_distance = exact squared pixel distance, lowest index on ties.
//
// All views are processed by every launch; there is no host round trip (the reference syncs on
// `.item()` at ours_utils.py:987).
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

struct ViewParams {  // one per view, lives in the workspace
  int n_fg;
  int n_valid;
  int rescaled;   // 1 when mask_ratio > thresh
  int after_res;
  int pad;
  float scale;    // scale_factor (1 when not rescaled)
  int status;     // bit 0: view without valid points (the reference raises there)
  int _pad;
};

__global__ void splat_init_kernel(ViewParams* vp, int V, int* ctr_max, int* ctr_min, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)V) {
    ViewParams z = {0, 0, 0, 0, 0, 1.0f, 0, 0};
    vp[i] = z;
  }
  if (i < n) {
    ctr_max[i] = -1;
    ctr_min[i] = INT_MAX;
  }
}

// counts: foreground pixels of the (res x res) mask and valid points, per view
__global__ void splat_count_kernel(const uint8_t* __restrict__ hard_masks,
                                   const uint8_t* __restrict__ valid, int V, int N, int res,
                                   ViewParams* vp) {
  const int v = blockIdx.y;
  const int npx = res * res;
  int fg = 0, nv = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x)
    fg += hard_masks[(size_t)v * npx + i] ? 1 : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    nv += valid[(size_t)v * N + i] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) {
    fg += __shfl_xor_sync(0xffffffffu, fg, o);
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (fg) atomicAdd(&vp[v].n_fg, fg);
    if (nv) atomicAdd(&vp[v].n_valid, nv);
  }
}

// ours_utils.py:967-987: mask ratio test and shrink parameters (fp32 scalar arithmetic)
__global__ void splat_decide_kernel(ViewParams* vp, int V, int res, float thresh,
                                    float one_minus_thresh, float* __restrict__ scale_out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  ViewParams p = vp[v];
  const float fg = (float)p.n_fg;
  const float mask_ratio = 1.0f - (float)p.n_valid / fg;
  if (p.n_valid == 0) p.status |= 1;
  if (mask_ratio > thresh) {
    const float wanted = (float)p.n_valid / one_minus_thresh;
    const float scale = wanted / fg;
    int after = (int)floorf((float)res * scale);
    if ((res - after) % 2 == 1) after += 1;
    p.rescaled = 1;
    p.after_res = after;
    p.pad = (res - after) / 2;
    p.scale = scale;
  } else {
    p.rescaled = 0;
    p.after_res = res;
    p.pad = 0;
    p.scale = 1.0f;
  }
  vp[v] = p;
  scale_out[v] = p.scale;
}

// effective foreground mask: copy, or bilinear(any-nonzero) shrink + zero pad
// (ours_utils.py:989-995; torchvision Resize on a bool tensor, antialias off)
__global__ void splat_mask_kernel(const uint8_t* __restrict__ hard_masks, const ViewParams* vp,
                                  int V, int res, uint8_t* __restrict__ fg_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int npx = res * res;
  if (i >= (size_t)V * npx) return;
  const int v = i / npx, y = (i % npx) / res, x = i % res;
  const ViewParams p = vp[v];
  const uint8_t* m = hard_masks + (size_t)v * npx;
  if (!p.rescaled) {
    fg_out[i] = m[y * res + x] ? 1 : 0;
    return;
  }
  const int oy = y - p.pad, ox = x - p.pad;
  if (oy < 0 || ox < 0 || oy >= p.after_res || ox >= p.after_res) {
    fg_out[i] = 0;
    return;
  }
  const float scale = (float)res / (float)p.after_res;  // area_pixel_compute_scale<float>
  float sy = scale * ((float)oy + 0.5f) - 0.5f;
  float sx = scale * ((float)ox + 0.5f) - 0.5f;
  sy = fmaxf(sy, 0.f);
  sx = fmaxf(sx, 0.f);
  const int y0 = min((int)sy, res - 1), x0 = min((int)sx, res - 1);
  const int y1 = min(y0 + 1, res - 1), x1 = min(x0 + 1, res - 1);
  const bool wy1 = clipf(sy - (float)y0, 0.f, 1.f) > 0.f;
  const bool wx1 = clipf(sx - (float)x0, 0.f, 1.f) > 0.f;
  bool any = m[y0 * res + x0] != 0;
  any |= wx1 && m[y0 * res + x1] != 0;
  any |= wy1 && m[y1 * res + x0] != 0;
  any |= wy1 && wx1 && m[y1 * res + x1] != 0;
  fg_out[i] = any ? 1 : 0;
}

// per valid point: (optionally rescaled) centre pixel; record max / min point index per pixel
__global__ void splat_scatter_kernel(const long long* __restrict__ point_pixels,
                                     const uint8_t* __restrict__ valid, const ViewParams* vp,
                                     int V, int N, int res, int* __restrict__ ctr_max,
                                     int* __restrict__ ctr_min) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  if (!valid[i]) return;
  const int v = i / N, n = i % N;
  long long r = point_pixels[2 * i], c = point_pixels[2 * i + 1];
  const ViewParams p = vp[v];
  if (p.rescaled) {
    // ours_utils.py:975-981
    const float fr = (float)res;
    float ur = (float)r / fr, uc = (float)c / fr;
    ur = ur * 2.0f - 1.0f;
    uc = uc * 2.0f - 1.0f;
    ur = ur * p.scale;
    uc = uc * p.scale;
    ur = (ur + 1.0f) * 0.5f;
    uc = (uc + 1.0f) * 0.5f;
    r = (long long)clipf(ur * fr, 0.f, (float)(res - 1));
    c = (long long)clipf(uc * fr, 0.f, (float)(res - 1));
  }
  const size_t q = ((size_t)v * res + r) * res + c;
  atomicMax(&ctr_max[q], n);
  atomicMin(&ctr_min[q], n);
}

// inner edge pixels (ours_utils.py:519-522) and, for each, the nearest valid point
// (expanding square rings over the per-pixel min-index map; exact, lowest index on ties)
__global__ void splat_edge_kernel(const uint8_t* __restrict__ fg, const int* __restrict__ ctr_min,
                                  const ViewParams* vp, int V, int res,
                                  int* __restrict__ edge_src) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int npx = res * res;
  if (i >= (size_t)V * npx) return;
  const int v = i / npx, y = (i % npx) / res, x = i % res;
  const uint8_t* m = fg + (size_t)v * npx;
  int src = -1;  // -1: not an edge pixel
  if (m[y * res + x]) {
    bool bg_near = false;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < res && xx >= 0 && xx < res && !m[yy * res + xx]) bg_near = true;
      }
    if (bg_near) {
      src = -2;  // edge pixel without any valid point in the view
      if (vp[v].n_valid > 0) {
        const int* cm = ctr_min + (size_t)v * npx;
        long long best_d = LLONG_MAX;
        int best_i = INT_MAX;
        for (int r = 0; r < res; ++r) {
          if (best_d < (long long)r * r) break;  // all remaining rings are strictly farther
          const int y0 = y - r, y1 = y + r, x0 = x - r, x1 = x + r;
          // top and bottom rows of the ring
          for (int xx = max(x0, 0); xx <= min(x1, res - 1); ++xx) {
            if (y0 >= 0) {
              const int c = cm[y0 * res + xx];
              if (c != INT_MAX) {
                const long long d = (long long)r * r + (long long)(xx - x) * (xx - x);
                if (d < best_d || (d == best_d && c < best_i)) best_d = d, best_i = c;
              }
            }
            if (r > 0 && y1 < res) {
              const int c = cm[y1 * res + xx];
              if (c != INT_MAX) {
                const long long d = (long long)r * r + (long long)(xx - x) * (xx - x);
                if (d < best_d || (d == best_d && c < best_i)) best_d = d, best_i = c;
              }
            }
          }
          // left and right columns (without the corners)
          for (int yy = max(y0 + 1, 0); yy <= min(y1 - 1, res - 1); ++yy) {
            if (x0 >= 0) {
              const int c = cm[yy * res + x0];
              if (c != INT_MAX) {
                const long long d = (long long)r * r + (long long)(yy - y) * (yy - y);
                if (d < best_d || (d == best_d && c < best_i)) best_d = d, best_i = c;
              }
            }
            if (x1 < res) {
              const int c = cm[yy * res + x1];
              if (c != INT_MAX) {
                const long long d = (long long)r * r + (long long)(yy - y) * (yy - y);
                if (d < best_d || (d == best_d && c < best_i)) best_d = d, best_i = c;
              }
            }
          }
        }
        src = best_i;
      }
    }
  }
  edge_src[i] = src;
}

// compose sparse image, hard_mask0, hard_mask2 and write them vertically flipped
// (ours_utils.py:1002-1043 and :866).  Windows of size (2s-1)^2 are resolved by gathering:
// the winner at a pixel is the highest flattened write index among the windows covering it.
__global__ void splat_compose_kernel(const float* __restrict__ colors,
                                     const uint8_t* __restrict__ fg,
                                     const int* __restrict__ ctr_max,
                                     const int* __restrict__ edge_src, int V, int res,
                                     int point_size, int edge_point_size,
                                     float* __restrict__ sparse, float* __restrict__ m0,
                                     float* __restrict__ m2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int npx = res * res;
  if (i >= (size_t)V * npx) return;
  const int v = i / npx, y = (i % npx) / res, x = i % res;
  const size_t base = (size_t)v * npx;
  const bool is_fg = fg[i] != 0;
  int pidx = -1;  // winning point
  {
    const int s = point_size - 1;
    for (int dy = -s; dy <= s; ++dy)
      for (int dx = -s; dx <= s; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < res && xx >= 0 && xx < res)
          pidx = max(pidx, ctr_max[base + yy * res + xx]);
      }
  }
  int eidx = -1;  // point whose colour the winning edge pixel carries
  bool edge_hit = false;
  {
    const int s = edge_point_size - 1;
    int best_lin = -1;
    for (int dy = -s; dy <= s; ++dy)
      for (int dx = -s; dx <= s; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < res && xx >= 0 && xx < res) {
          const int e = edge_src[base + yy * res + xx];
          if (e != -1 && yy * res + xx > best_lin) {
            best_lin = yy * res + xx;
            eidx = e;
            edge_hit = true;
          }
        }
      }
  }
  float r = 0.f, g = 0.f, b = 0.f;
  if (edge_hit && eidx >= 0) {
    r = colors[3 * eidx], g = colors[3 * eidx + 1], b = colors[3 * eidx + 2];
  } else if (!edge_hit && pidx >= 0) {
    r = colors[3 * pidx], g = colors[3 * pidx + 1], b = colors[3 * pidx + 2];
  }
  const float f0 = is_fg ? 1.f : 0.f;
  const float f2 = (pidx >= 0 || edge_hit) ? 1.f : (1.f - f0);
  const size_t o = ((size_t)v * 3 * res + (res - 1 - y)) * res + x;  // channel 0, flipped row
  const size_t cs = (size_t)res * res;
  sparse[o] = r * f0;
  sparse[o + cs] = g * f0;
  sparse[o + 2 * cs] = b * f0;
  m0[o] = f0;
  m0[o + cs] = f0;
  m0[o + 2 * cs] = f0;
  m2[o] = f2;
  m2[o + cs] = f2;
  m2[o + 2 * cs] = f2;
}

size_t sparse_images_workspace_bytes(int V, int res) {
  const size_t npx = (size_t)V * res * res;
  return (size_t)MAX_VIEWS * sizeof(ViewParams) + npx * (3 * sizeof(int) + 1) + 256;
}

int sparse_images_launch(const long long* point_pixels, const float* colors, const uint8_t* valid,
                         const uint8_t* hard_masks, int V, int N, int res, int point_size,
                         int edge_point_size, double mask_ratio_thresh, void* workspace,
                         float* sparse, float* m0, float* m2, float* scale_factors,
                         cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && V <= MAX_VIEWS, "view count %d out of range", V);
  PDR_CHECK_ARG(point_size >= 1 && edge_point_size >= 1, "point sizes must be >= 1");
  PDR_CHECK_ARG(N > 0 && res > 0, "empty input");
  const size_t npx = (size_t)V * res * res;
  uint8_t* w = (uint8_t*)workspace;
  ViewParams* vp = (ViewParams*)w;
  w += (size_t)MAX_VIEWS * sizeof(ViewParams);
  int* ctr_max = (int*)w;
  w += npx * sizeof(int);
  int* ctr_min = (int*)w;
  w += npx * sizeof(int);
  int* edge_src = (int*)w;
  w += npx * sizeof(int);
  uint8_t* fg = w;

  const float thresh = (float)mask_ratio_thresh;
  const float one_minus = (float)(1.0 - mask_ratio_thresh);
  const size_t nmax = npx > (size_t)V ? npx : (size_t)V;
  splat_init_kernel<<<cdiv(nmax, 256), 256, 0, stream>>>(vp, V, ctr_max, ctr_min, npx);
  PDR_COUNT_LAUNCH();
  splat_count_kernel<<<dim3(32, V), 256, 0, stream>>>(hard_masks, valid, V, N, res, vp);
  PDR_COUNT_LAUNCH();
  splat_decide_kernel<<<1, 32, 0, stream>>>(vp, V, res, thresh, one_minus, scale_factors);
  PDR_COUNT_LAUNCH();
  splat_mask_kernel<<<cdiv(npx, 256), 256, 0, stream>>>(hard_masks, vp, V, res, fg);
  PDR_COUNT_LAUNCH();
  splat_scatter_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(point_pixels, valid, vp, V, N,
                                                                    res, ctr_max, ctr_min);
  PDR_COUNT_LAUNCH();
  splat_edge_kernel<<<cdiv(npx, 128), 128, 0, stream>>>(fg, ctr_min, vp, V, res, edge_src);
  PDR_COUNT_LAUNCH();
  splat_compose_kernel<<<cdiv(npx, 256), 256, 0, stream>>>(colors, fg, ctr_max, edge_src, V, res,
                                                          point_size, edge_point_size, sparse, m0,
                                                          m2);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
