// PROJECT stage kernels: camera transform + crop/rescale (K1), mesh z-buffer rasteriser (K2),
// mask down-resolution (K3), depth visibility (K4), point pixel quantisation.
//
// Reference: pointdreamer/ours_utils.py:93-150 (get_rendered_hard_mask_and_face_idx_batch),
// :153-202 (get_point_validation_by_depth), demo.py:103-104, 121-125.  kaolin's
// Camera.transform and nvdiffrast's rasterize are third-party and unvendored: the canonical
// rules implemented here are those of oracle/camera.py and oracle/project.py:rasterize.
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

// ------------------------------------------------------------------ K1 ----
__global__ void minmax_init_kernel(int* mm, int V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V * 4) mm[i] = (i & 2) ? INT_MIN : INT_MAX;  // [min_x, min_y, max_x, max_y]
}

// transform mesh vertices for every view, store raw NDC into pos, reduce per-view uv min/max
__global__ void vertex_transform_kernel(const float* __restrict__ cams,
                                        const float* __restrict__ verts, int Vm, int V,
                                        float* __restrict__ pos, int* __restrict__ mm) {
  __shared__ float sp[CAM_PARAM_FLOATS];
  const int v = blockIdx.y;
  if (threadIdx.x < CAM_PARAM_FLOATS) sp[threadIdx.x] = cams[v * CAM_PARAM_FLOATS + threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int kminx = INT_MAX, kminy = INT_MAX, kmaxx = INT_MIN, kmaxy = INT_MIN;
  if (i < Vm) {
    float nx, ny, nz;
    cam_transform(sp, verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], nx, ny, nz);
    float4 o = make_float4(nx, ny, nz, 1.0f);
    reinterpret_cast<float4*>(pos)[(size_t)v * Vm + i] = o;
    kminx = kmaxx = float_to_ordered(nx);
    kminy = kmaxy = float_to_ordered(ny);
  }
  // warp reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    kminx = min(kminx, __shfl_xor_sync(0xffffffffu, kminx, o));
    kminy = min(kminy, __shfl_xor_sync(0xffffffffu, kminy, o));
    kmaxx = max(kmaxx, __shfl_xor_sync(0xffffffffu, kmaxx, o));
    kmaxy = max(kmaxy, __shfl_xor_sync(0xffffffffu, kmaxy, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[v * 4 + 0], kminx);
    atomicMin(&mm[v * 4 + 1], kminy);
    atomicMax(&mm[v * 4 + 2], kmaxx);
    atomicMax(&mm[v * 4 + 3], kmaxy);
  }
}

// per view: centres / scale from min/max (ours_utils.py:112-118)
__global__ void crop_params_kernel(const int* __restrict__ mm, int V, int rescale,
                                   float* __restrict__ centers, float* __restrict__ scales) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  if (rescale) {
    const float mnx = ordered_to_float(mm[v * 4 + 0]), mny = ordered_to_float(mm[v * 4 + 1]);
    const float mxx = ordered_to_float(mm[v * 4 + 2]), mxy = ordered_to_float(mm[v * 4 + 3]);
    centers[v * 2 + 0] = (mnx + mxx) / 2.0f;
    centers[v * 2 + 1] = (mny + mxy) / 2.0f;
    scales[v] = fmaxf(mxx - mnx, mxy - mny);
  } else {
    centers[v * 2 + 0] = 0.f;
    centers[v * 2 + 1] = 0.f;
    scales[v] = 2.f;
  }
}

// rescale vertices in place (ours_utils.py:119-123 / 132-133) and transform+rescale points
// (ours_utils.py:125-130 / 135-141); one pass over the cloud for all views.
__global__ void rescale_kernel(const float* __restrict__ cams, const float* __restrict__ points,
                               int N, int Vm, int V, int rescale, float pad_mul,
                               const float* __restrict__ centers, const float* __restrict__ scales,
                               float* __restrict__ pos, float* __restrict__ vuv,
                               float* __restrict__ puv, float* __restrict__ pdepth) {
  __shared__ float sp[CAM_PARAM_FLOATS];
  const int v = blockIdx.y;
  if (threadIdx.x < CAM_PARAM_FLOATS) sp[threadIdx.x] = cams[v * CAM_PARAM_FLOATS + threadIdx.x];
  __syncthreads();
  const float cx = centers[v * 2], cy = centers[v * 2 + 1], sc = scales[v];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Vm) {
    float4 p = reinterpret_cast<float4*>(pos)[(size_t)v * Vm + i];
    float u, w;
    if (rescale) {
      u = clipf(((p.x - cx) / sc) * pad_mul + 0.5f, 0.f, 1.f);
      w = clipf(((p.y - cy) / sc) * pad_mul + 0.5f, 0.f, 1.f);
      p.x = u * 2.0f - 1.0f;
      p.y = w * 2.0f - 1.0f;
      reinterpret_cast<float4*>(pos)[(size_t)v * Vm + i] = p;
    } else {
      u = clipf((p.x + 1.0f) * 0.5f, 0.f, 1.f);
      w = clipf((p.y + 1.0f) * 0.5f, 0.f, 1.f);
    }
    reinterpret_cast<float2*>(vuv)[(size_t)v * Vm + i] = make_float2(u, w);
  }
  if (i < N) {
    float nx, ny, nz;
    cam_transform(sp, points[3 * i], points[3 * i + 1], points[3 * i + 2], nx, ny, nz);
    float u, w;
    if (rescale) {
      u = ((nx - cx) / sc) * pad_mul + 0.5f;
      w = ((ny - cy) / sc) * pad_mul + 0.5f;
    } else {
      u = (nx + 1.0f) * 0.5f;
      w = (ny + 1.0f) * 0.5f;
    }
    reinterpret_cast<float2*>(puv)[(size_t)v * N + i] = make_float2(u, w);
    pdepth[(size_t)v * N + i] = nz;
  }
}

int project_launch(const float* cams, const float* vertices, int Vm, const float* points, int N,
                   int V, int rescale, double padding, int* ws_minmax, float* pos,
                   float* vertice_uvs, float* uv_centers, float* uv_scales, float* point_uvs,
                   float* point_depths, cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && V <= MAX_VIEWS, "view count %d out of range (1..%d)", V, MAX_VIEWS);
  PDR_CHECK_ARG(Vm > 0 && N >= 0, "empty mesh");
  const float pad_mul = (float)(1.0 - 2.0 * padding);
  minmax_init_kernel<<<1, 128, 0, stream>>>(ws_minmax, V);
  PDR_COUNT_LAUNCH();
  dim3 gv(cdiv(Vm, 256), V);
  vertex_transform_kernel<<<gv, 256, 0, stream>>>(cams, vertices, Vm, V, pos, ws_minmax);
  PDR_COUNT_LAUNCH();
  crop_params_kernel<<<1, 32, 0, stream>>>(ws_minmax, V, rescale, uv_centers, uv_scales);
  PDR_COUNT_LAUNCH();
  const int M = Vm > N ? Vm : N;
  dim3 gp(cdiv(M, 256), V);
  rescale_kernel<<<gp, 256, 0, stream>>>(cams, points, N, Vm, V, rescale, pad_mul, uv_centers,
                                         uv_scales, pos, vertice_uvs, point_uvs, point_depths);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ K2 ----
// (SUBPIX, snap_coord, edge_inclusive, floordiv: geom_common.cuh)

__global__ void zkey_init_kernel(unsigned long long* keys, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ~0ull;
}

// one warp per (view, triangle); lanes stride over the bounding box
__global__ void raster_kernel(const float* __restrict__ pos, const int* __restrict__ faces, int V,
                              int Vm, int F, int res, unsigned long long* __restrict__ keys) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= V * F) return;
  const int v = warp / F, f = warp - v * F;
  const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
  const float4* P = reinterpret_cast<const float4*>(pos) + (size_t)v * Vm;
  const float4 A = P[ia], B = P[ib], C = P[ic];
  const long long ax = snap_coord(A.x, res), ay = snap_coord(A.y, res);
  const long long bx = snap_coord(B.x, res), by = snap_coord(B.y, res);
  const long long cx = snap_coord(C.x, res), cy = snap_coord(C.y, res);
  const long long area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
  if (area == 0) return;
  const long long sgn = area > 0 ? 1 : -1;
  const long long H = SUBPIX / 2;
  long long xmin = floordiv(min(ax, min(bx, cx)) - H + SUBPIX - 1, SUBPIX);
  long long xmax = floordiv(max(ax, max(bx, cx)) - H, SUBPIX);
  long long ymin = floordiv(min(ay, min(by, cy)) - H + SUBPIX - 1, SUBPIX);
  long long ymax = floordiv(max(ay, max(by, cy)) - H, SUBPIX);
  xmin = max(xmin, 0ll);
  ymin = max(ymin, 0ll);
  xmax = min(xmax, (long long)res - 1);
  ymax = min(ymax, (long long)res - 1);
  if (xmin > xmax || ymin > ymax) return;
  const bool incA = edge_inclusive(sgn * (cx - bx), sgn * (cy - by));
  const bool incB = edge_inclusive(sgn * (ax - cx), sgn * (ay - cy));
  const bool incC = edge_inclusive(sgn * (bx - ax), sgn * (by - ay));
  const int bw = (int)(xmax - xmin + 1);
  const long long npx = (long long)bw * (ymax - ymin + 1);
  unsigned long long* kv = keys + (size_t)v * res * res;
  for (long long t = lane; t < npx; t += 32) {
    const long long yy = ymin + t / bw, xx = xmin + t % bw;
    const long long px = xx * SUBPIX + H, py = yy * SUBPIX + H;
    const long long eA = sgn * ((cx - bx) * (py - by) - (cy - by) * (px - bx));
    const long long eB = sgn * ((ax - cx) * (py - cy) - (ay - cy) * (px - cx));
    const long long eC = sgn * ((bx - ax) * (py - ay) - (by - ay) * (px - ax));
    const bool in = (eA > 0 || (eA == 0 && incA)) && (eB > 0 || (eB == 0 && incB)) &&
                    (eC > 0 || (eC == 0 && incC));
    if (!in) continue;
    const float wa = __ll2float_rn(eA), wb = __ll2float_rn(eB), wc = __ll2float_rn(eC);
    const float tot = __ll2float_rn(eA + eB + eC);
    const float z = ((wa * A.z + wb * B.z) + wc * C.z) / tot;
    if (!(z >= -1.0f && z <= 1.0f)) continue;
    const unsigned long long key =
        ((unsigned long long)float_to_ordered_u32(z) << 32) | (unsigned int)f;
    atomicMin(&kv[yy * res + xx], key);
  }
}

// keys -> depth / face id / mask (+ the res-sized "any" mask, demo.py:103-104)
__global__ void raster_resolve_kernel(const unsigned long long* __restrict__ keys, int V, int res,
                                      int out_res, float* __restrict__ depth,
                                      long long* __restrict__ face_idx,
                                      uint8_t* __restrict__ mask_cam,
                                      uint8_t* __restrict__ mask_out) {
  // thread per OUTPUT pixel of the (possibly half-resolution) mask; ratio = res / out_res (1 or 2)
  const int ratio = res / out_res;
  const size_t n = (size_t)V * out_res * out_res;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ox = i % out_res, oy = (i / out_res) % out_res, v = i / ((size_t)out_res * out_res);
  bool any = false;
  for (int dy = 0; dy < ratio; ++dy)
    for (int dx = 0; dx < ratio; ++dx) {
      const size_t p = ((size_t)v * res + (oy * ratio + dy)) * res + (ox * ratio + dx);
      const unsigned long long k = keys[p];
      const bool hit = k != ~0ull;
      depth[p] = hit ? ordered_u32_to_float((unsigned int)(k >> 32)) : 0.0f;
      face_idx[p] = hit ? (long long)(unsigned int)(k & 0xFFFFFFFFu) : -1ll;
      mask_cam[p] = hit ? 1 : 0;
      any |= hit;
    }
  mask_out[i] = any ? 1 : 0;
}

int rasterize_launch(const float* pos, const int* faces, int V, int Vm, int F, int res,
                     int out_res, unsigned long long* ws_keys, float* depth, long long* face_idx,
                     uint8_t* mask_cam, uint8_t* mask_out, cudaStream_t stream) {
  PDR_CHECK_ARG(out_res == res || out_res * 2 == res,
                "mask resize %d -> %d unsupported (cam_res must equal res or 2*res)", res, out_res);
  PDR_CHECK_ARG(F > 0 && V > 0, "empty mesh");
  const size_t n = (size_t)V * res * res;
  zkey_init_kernel<<<cdiv(n, 256), 256, 0, stream>>>(ws_keys, n);
  PDR_COUNT_LAUNCH();
  const long long warps = (long long)V * F;
  raster_kernel<<<cdiv(warps * 32, 256), 256, 0, stream>>>(pos, faces, V, Vm, F, res, ws_keys);
  PDR_COUNT_LAUNCH();
  const size_t no = (size_t)V * out_res * out_res;
  raster_resolve_kernel<<<cdiv(no, 256), 256, 0, stream>>>(ws_keys, V, res, out_res, depth,
                                                          face_idx, mask_cam, mask_out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}


// ------------------------------------------------------------------ K3 ----
// demo.py:103-104: Resize((res,res)) of the float mask (bilinear, antialias off) then .bool();
// for the exact 2x reduction every output pixel is the OR of its 2x2 block.
__global__ void mask_half_any_kernel(const uint8_t* __restrict__ in, int V, int res_in,
                                     uint8_t* __restrict__ out) {
  const int ro = res_in / 2;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * ro * ro) return;
  const int x = i % ro, y = (i / ro) % ro, v = i / ((size_t)ro * ro);
  const uint8_t* m = in + (size_t)v * res_in * res_in;
  const int y0 = 2 * y, x0 = 2 * x;
  out[i] = (m[(size_t)y0 * res_in + x0] | m[(size_t)y0 * res_in + x0 + 1] |
            m[(size_t)(y0 + 1) * res_in + x0] | m[(size_t)(y0 + 1) * res_in + x0 + 1])
               ? 1
               : 0;
}

int mask_half_any_launch(const uint8_t* in, int V, int res_in, uint8_t* out,
                         cudaStream_t stream) {
  PDR_CHECK_ARG(res_in % 2 == 0 && V > 0, "mask resolution must be even");
  const size_t n = (size_t)V * (res_in / 2) * (res_in / 2);
  mask_half_any_kernel<<<cdiv(n, 256), 256, 0, stream>>>(in, V, res_in, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ K4 ----
// ours_utils.py:153-202: pixel = long(clip(uv*cam_res, 0, cam_res-1)), (row, col) = (y, x);
// visible iff point_depth - mesh_depth[row, col] <= offset.  Optionally also emits
// demo.py:121-125's pixel at `res` (long() BEFORE clip).
__global__ void point_visibility_kernel(const float* __restrict__ puv,
                                        const float* __restrict__ pdepth,
                                        const float* __restrict__ mesh_depths, int V, int N,
                                        int cam_res, float offset, int res,
                                        uint8_t* __restrict__ vis, long long* __restrict__ pix_cam,
                                        long long* __restrict__ pix_res) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N;
  const float2 uv = reinterpret_cast<const float2*>(puv)[i];
  const float fc = (float)cam_res;
  const long long col = (long long)clipf(uv.x * fc, 0.f, (float)(cam_res - 1));
  const long long row = (long long)clipf(uv.y * fc, 0.f, (float)(cam_res - 1));
  if (vis) {
    const float ref = mesh_depths[((size_t)v * cam_res + row) * cam_res + col];
    vis[i] = (pdepth[i] - ref <= offset) ? 1 : 0;
  }
  if (pix_cam) {
    pix_cam[2 * i] = row;
    pix_cam[2 * i + 1] = col;
  }
  if (pix_res) {
    const float fr = (float)res;
    const long long c2 = clipll((long long)(uv.x * fr), 0, res - 1);
    const long long r2 = clipll((long long)(uv.y * fr), 0, res - 1);
    pix_res[2 * i] = r2;
    pix_res[2 * i + 1] = c2;
  }
}

int point_visibility_launch(const float* puv, const float* pdepth, const float* mesh_depths,
                            int V, int N, int cam_res, float offset, int res, uint8_t* vis,
                            long long* pix_cam, long long* pix_res, cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && N > 0, "empty input");
  PDR_CHECK_ARG(!vis || (pdepth && mesh_depths), "visibility needs depths");
  point_visibility_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(
      puv, pdepth, mesh_depths, V, N, cam_res, offset, res, vis, pix_cam, pix_res);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
