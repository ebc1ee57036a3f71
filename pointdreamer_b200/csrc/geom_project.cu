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
// Tile-binned z-buffer rasteriser (SUBPIX, snap_coord, edge_inclusive, floordiv: geom_common.cuh).
//
//   raster_setup_kernel   thread per (view, triangle): snap the vertices to the 1/256-px grid, store a
//                         64-byte record, count the 32 x 8 pixel tiles its bounding box overlaps
//                         (triangles spanning more than RT_MAX_SPAN tiles go to a per-view list);
//   raster_scan_kernel    exclusive scan of the per-tile counts (one block);
//   raster_bin_kernel     thread per (view, triangle): write the face id into the bins of its tiles;
//   raster_tile_kernel    one CTA per tile with the tile's z-buffer in SHARED memory: the tile's
//                         triangle records are staged through shared memory in chunks, each warp
//                         takes whole triangles and its lanes walk the triangle's bounding box
//                         inside the tile (three exact fp64 edge functions per pixel), hits go to
//                         the shared z-buffer with a shared-memory atomicMin on the packed
//                         (z, face id) key - no global atomics, no z-buffer in global memory -
//                         then thread = pixel writes depth / face id / masks (and the res/2 "any"
//                         mask of demo.py:103-104) once, coalesced.
// The result is the minimum over a SET of (z, id) keys, so the (non-deterministic) order of the
// faces inside a bin does not matter: output is bit-identical to the one-warp-per-triangle +
// atomicMin version it replaces, and to oracle/project.py:rasterize.
static constexpr int RT_W = 32, RT_H = 8;   // pixel tile of one CTA
static constexpr int RT_MAX_SPAN = 16;      // more tiles than this: the per-view "large" list
static constexpr int RT_CHUNK = 64;         // triangle records staged per round

struct __align__(16) RasterTri {
  long long ax, ay, bx, by, cx, cy;  // snapped vertices, 1/256 px
  float az, bz, cz;
  int f;
};
static_assert(sizeof(RasterTri) == 64, "RasterTri must be 64 bytes");

// Fast-path record of a triangle whose snapped coordinates are below 2^24: the three (sign-
// normalised) edge functions e = A*px + B*py + C are evaluated in fp64, where every product and
// sum is an integer below 2^53 and therefore EXACT - the same values as the int64 evaluation of
// raster_shade, at 2 DFMA per edge instead of ~16 integer instructions.
struct __align__(16) RasterEdge {
  double A[3], B[3], C[3];  // edges opposite to vertex a, b, c
  float thr[3];             // 0 when the edge owns its boundary pixels, 1 otherwise (e >= thr)
  float az, bz, cz;
  short x0, y0, x1, y1;     // pixel bounding box (clipped to the image)
  int f;
  int pad;
};
static_assert(sizeof(RasterEdge) == 112, "RasterEdge must be 112 bytes");
static constexpr long long RT_EXACT_LIMIT = 1ll << 24;

struct RasterWs {
  RasterTri* tris;     // [V*F]
  RasterEdge* edges;   // [V*F]
  int* counts;         // [V*T + 1]  per-tile counts, then write cursors
  int* offs;           // [V*T + 1]  exclusive scan
  int* bins;           // [V*F*RT_MAX_SPAN]
  int* large;          // [V*F]
  int* large_count;    // [V]
};

__host__ __device__ inline size_t rt_align(size_t x) { return (x + 255) & ~(size_t)255; }

static RasterWs raster_ws_carve(void* ws, int V, int F, int tiles) {
  uint8_t* w = (uint8_t*)ws;
  RasterWs r;
  r.tris = (RasterTri*)w;
  w += rt_align((size_t)V * F * sizeof(RasterTri));
  r.edges = (RasterEdge*)w;
  w += rt_align((size_t)V * F * sizeof(RasterEdge));
  r.counts = (int*)w;
  w += rt_align(((size_t)V * tiles + 1) * 4);
  r.large_count = (int*)w;  // directly after counts: one memset clears both
  w += rt_align((size_t)V * 4);
  r.offs = (int*)w;
  w += rt_align(((size_t)V * tiles + 1) * 4);
  r.bins = (int*)w;
  w += rt_align((size_t)V * F * RT_MAX_SPAN * 4);
  r.large = (int*)w;
  return r;
}

size_t rasterize_workspace_bytes(int V, int F, int res) {
  const size_t tiles = (size_t)cdiv(res, RT_W) * cdiv(res, RT_H);
  return rt_align((size_t)V * F * sizeof(RasterTri)) + rt_align((size_t)V * F * sizeof(RasterEdge)) +
         2 * rt_align(((size_t)V * tiles + 1) * 4) +
         rt_align((size_t)V * 4) + rt_align((size_t)V * F * RT_MAX_SPAN * 4) +
         rt_align((size_t)V * F * 4) + 256;
}

// pixel bounding box of a snapped triangle clipped to the image; false when empty / degenerate
__device__ __forceinline__ bool raster_bbox(const RasterTri& t, int res, int& x0, int& y0, int& x1,
                                            int& y1) {
  const long long area = (t.bx - t.ax) * (t.cy - t.ay) - (t.by - t.ay) * (t.cx - t.ax);
  if (area == 0) return false;
  const long long H = SUBPIX / 2;
  long long xmin = floordiv(min(t.ax, min(t.bx, t.cx)) - H + SUBPIX - 1, SUBPIX);
  long long xmax = floordiv(max(t.ax, max(t.bx, t.cx)) - H, SUBPIX);
  long long ymin = floordiv(min(t.ay, min(t.by, t.cy)) - H + SUBPIX - 1, SUBPIX);
  long long ymax = floordiv(max(t.ay, max(t.by, t.cy)) - H, SUBPIX);
  xmin = max(xmin, 0ll);
  ymin = max(ymin, 0ll);
  xmax = min(xmax, (long long)res - 1);
  ymax = min(ymax, (long long)res - 1);
  if (xmin > xmax || ymin > ymax) return false;
  x0 = (int)xmin, x1 = (int)xmax, y0 = (int)ymin, y1 = (int)ymax;
  return true;
}

// every snapped coordinate below 2^24: the fp64 edge functions of RasterEdge are exact
__device__ __forceinline__ bool raster_exact_range(const RasterTri& t) {
  const long long m = max(max(max(llabs(t.ax), llabs(t.ay)), max(llabs(t.bx), llabs(t.by))),
                          max(llabs(t.cx), llabs(t.cy)));
  return m < RT_EXACT_LIMIT;
}

__global__ void raster_setup_kernel(const float* __restrict__ pos, const int* __restrict__ faces,
                                    int V, int Vm, int F, int res, int tiles_x, int tiles,
                                    RasterWs ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * F) return;
  const int v = i / F, f = i - v * F;
  const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
  const float4* P = reinterpret_cast<const float4*>(pos) + (size_t)v * Vm;
  const float4 A = P[ia], B = P[ib], C = P[ic];
  RasterTri t;
  t.ax = snap_coord(A.x, res), t.ay = snap_coord(A.y, res);
  t.bx = snap_coord(B.x, res), t.by = snap_coord(B.y, res);
  t.cx = snap_coord(C.x, res), t.cy = snap_coord(C.y, res);
  t.az = A.z, t.bz = B.z, t.cz = C.z;
  t.f = f;
  ws.tris[i] = t;
  int x0, y0, x1, y1;
  if (!raster_bbox(t, res, x0, y0, x1, y1)) return;
  const int tx0 = x0 / RT_W, tx1 = x1 / RT_W, ty0 = y0 / RT_H, ty1 = y1 / RT_H;
  if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > RT_MAX_SPAN || !raster_exact_range(t)) {
    ws.large[(size_t)v * F + atomicAdd(&ws.large_count[v], 1)] = f;
    return;
  }
  {
    const long long area = (t.bx - t.ax) * (t.cy - t.ay) - (t.by - t.ay) * (t.cx - t.ax);
    const long long sgn = area > 0 ? 1 : -1;
    RasterEdge e;
    // eA = sgn*((cx-bx)*(py-by) - (cy-by)*(px-bx)) = A*px + B*py + C with
    const long long dxs[3] = {sgn * (t.cx - t.bx), sgn * (t.ax - t.cx), sgn * (t.bx - t.ax)};
    const long long dys[3] = {sgn * (t.cy - t.by), sgn * (t.ay - t.cy), sgn * (t.by - t.ay)};
    const long long ox[3] = {t.bx, t.cx, t.ax}, oy[3] = {t.by, t.cy, t.ay};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      e.A[k] = (double)(-dys[k]);
      e.B[k] = (double)dxs[k];
      e.C[k] = (double)(dys[k] * ox[k] - dxs[k] * oy[k]);
      e.thr[k] = edge_inclusive(dxs[k], dys[k]) ? 0.0f : 1.0f;
    }
    e.az = t.az, e.bz = t.bz, e.cz = t.cz;
    e.x0 = (short)x0, e.y0 = (short)y0, e.x1 = (short)x1, e.y1 = (short)y1;
    e.f = f;
    e.pad = 0;
    ws.edges[i] = e;
  }
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&ws.counts[(size_t)v * tiles + ty * tiles_x + tx], 1);
}

// exclusive scan of counts[0..n) into offs[0..n], single block; counts are zeroed (they become
// the write cursors of raster_bin_kernel)
__global__ void raster_scan_kernel(int* __restrict__ counts, int* __restrict__ offs, int n) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int val = i < n ? counts[i] : 0;
    s[threadIdx.x] = val;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) {
      offs[i] = carry + s[threadIdx.x] - val;
      counts[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) offs[n] = carry;
}

__global__ void raster_bin_kernel(int V, int F, int res, int tiles_x, int tiles, RasterWs ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * F) return;
  const int v = i / F, f = i - v * F;
  const RasterTri t = ws.tris[i];
  int x0, y0, x1, y1;
  if (!raster_bbox(t, res, x0, y0, x1, y1)) return;
  const int tx0 = x0 / RT_W, tx1 = x1 / RT_W, ty0 = y0 / RT_H, ty1 = y1 / RT_H;
  if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > RT_MAX_SPAN || !raster_exact_range(t)) return;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) {
      const size_t tile = (size_t)v * tiles + ty * tiles_x + tx;
      ws.bins[ws.offs[tile] + atomicAdd(&ws.counts[tile], 1)] = f;
    }
}

// coverage + depth of pixel centre (px, py) (sub-pixel units) for one triangle record (int64 path);
// returns the (z, face id) key or ~0
__device__ __forceinline__ unsigned long long raster_shade(const RasterTri& t, long long px,
                                                           long long py) {
  const long long area = (t.bx - t.ax) * (t.cy - t.ay) - (t.by - t.ay) * (t.cx - t.ax);
  const long long sgn = area > 0 ? 1 : -1;
  const long long eA = sgn * ((t.cx - t.bx) * (py - t.by) - (t.cy - t.by) * (px - t.bx));
  const long long eB = sgn * ((t.ax - t.cx) * (py - t.cy) - (t.ay - t.cy) * (px - t.cx));
  const long long eC = sgn * ((t.bx - t.ax) * (py - t.ay) - (t.by - t.ay) * (px - t.ax));
  if (eA < 0 || eB < 0 || eC < 0) return ~0ull;
  if ((eA == 0 && !edge_inclusive(sgn * (t.cx - t.bx), sgn * (t.cy - t.by))) ||
      (eB == 0 && !edge_inclusive(sgn * (t.ax - t.cx), sgn * (t.ay - t.cy))) ||
      (eC == 0 && !edge_inclusive(sgn * (t.bx - t.ax), sgn * (t.by - t.ay))))
    return ~0ull;
  const float wa = __ll2float_rn(eA), wb = __ll2float_rn(eB), wc = __ll2float_rn(eC);
  const float tot = __ll2float_rn(eA + eB + eC);
  const float z = ((wa * t.az + wb * t.bz) + wc * t.cz) / tot;
  if (!(z >= -1.0f && z <= 1.0f)) return ~0ull;
  return ((unsigned long long)float_to_ordered_u32(z) << 32) | (unsigned int)t.f;
}

// the same for a fast-path record: three exact fp64 edge functions
__device__ __forceinline__ unsigned long long raster_shade_fast(const RasterEdge& t, double px,
                                                                double py) {
  const double eA = fma(t.A[0], px, fma(t.B[0], py, t.C[0]));
  const double eB = fma(t.A[1], px, fma(t.B[1], py, t.C[1]));
  const double eC = fma(t.A[2], px, fma(t.B[2], py, t.C[2]));
  if (eA < (double)t.thr[0] || eB < (double)t.thr[1] || eC < (double)t.thr[2]) return ~0ull;
  // integers below 2^53: double -> float rounds exactly like __ll2float_rn of the int64 value
  const float wa = __double2float_rn(eA), wb = __double2float_rn(eB), wc = __double2float_rn(eC);
  const float tot = __double2float_rn(eA + eB + eC);
  const float z = ((wa * t.az + wb * t.bz) + wc * t.cz) / tot;
  if (!(z >= -1.0f && z <= 1.0f)) return ~0ull;
  return ((unsigned long long)float_to_ordered_u32(z) << 32) | (unsigned int)t.f;
}

// One CTA per 32 x 8 pixel tile, z-buffer of the tile in SHARED memory.  The tile's triangle records
// are staged in shared memory RT_CHUNK at a time; each of the 8 warps takes whole triangles and its
// lanes walk the part of the triangle's bounding box that lies inside the tile (16 x 2 pixels per
// step), so the work is proportional to the covered area, not to triangles x tile pixels.  Hits go
// to the shared z-buffer with a 64-bit shared-memory atomicMin on the packed (z, face id) key (the
// minimum of a set: order independent); then thread = pixel writes all outputs once, coalesced.
__global__ void __launch_bounds__(RT_W * RT_H, 4)  // <= 64 registers: 4+ tiles resident per SM
raster_tile_kernel(int V, int F, int res, int out_res, int tiles_x, int tiles, RasterWs ws,
                   float* __restrict__ depth, long long* __restrict__ face_idx,
                   uint8_t* __restrict__ mask_cam, uint8_t* __restrict__ mask_out) {
  __shared__ RasterEdge se[RT_CHUNK];
  __shared__ RasterTri st[RT_CHUNK];
  __shared__ unsigned long long zbuf[RT_H * RT_W];
  __shared__ int s_n;
  __shared__ uint8_t s_hit[RT_H][RT_W];
  const int tile = blockIdx.x % tiles, v = blockIdx.x / tiles;
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sx = lane & 15, sy = lane >> 4;  // a warp step covers 16 x 2 pixels
  const int X0 = tx * RT_W, X1 = min(X0 + RT_W, res) - 1, Y0 = ty * RT_H, Y1 = min(Y0 + RT_H, res) - 1;
  const RasterTri* tris = ws.tris + (size_t)v * F;
  const RasterEdge* edges = ws.edges + (size_t)v * F;
  zbuf[threadIdx.x] = ~0ull;
  // ---- the tile's bin (fast-path records) ----
  const int b0 = ws.offs[(size_t)v * tiles + tile], b1 = ws.offs[(size_t)v * tiles + tile + 1];
  for (int base = b0; base < b1; base += RT_CHUNK) {
    const int n = min(RT_CHUNK, b1 - base);
    __syncthreads();
    // 7 threads copy one 112-byte record (16 bytes each)
    for (int k = threadIdx.x; k < n * 7; k += RT_W * RT_H) {
      const int r = k / 7, part = k - r * 7;
      reinterpret_cast<uint4*>(&se[r])[part] =
          reinterpret_cast<const uint4*>(&edges[ws.bins[base + r]])[part];
    }
    __syncthreads();
    for (int r = warp; r < n; r += RT_H) {
      const RasterEdge& t = se[r];
      const int x0 = max((int)t.x0, X0), x1 = min((int)t.x1, X1);
      const int y0 = max((int)t.y0, Y0), y1 = min((int)t.y1, Y1);
      for (int yy = y0 + sy; yy <= y1; yy += 2)
        for (int xx = x0 + sx; xx <= x1; xx += 16) {
          const unsigned long long key =
              raster_shade_fast(t, (double)(xx * SUBPIX + SUBPIX / 2), (double)(yy * SUBPIX + SUBPIX / 2));
          if (key != ~0ull) atomicMin(&zbuf[(yy - Y0) * RT_W + (xx - X0)], key);
        }
    }
  }
  // ---- the view's large triangles (bounding box > RT_MAX_SPAN tiles, or coordinates beyond the
  //      exact fp64 range), filtered per tile, int64 edge functions ----
  const int nl = ws.large_count[v];
  const int* large = ws.large + (size_t)v * F;
  for (int base = 0; base < nl; base += RT_CHUNK) {
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (threadIdx.x < RT_CHUNK && base + threadIdx.x < nl) {
      const RasterTri t = tris[large[base + threadIdx.x]];
      int x0, y0, x1, y1;
      if (raster_bbox(t, res, x0, y0, x1, y1) && x0 <= X1 && x1 >= X0 && y0 <= Y1 && y1 >= Y0)
        st[atomicAdd(&s_n, 1)] = t;
    }
    __syncthreads();
    const int n = s_n;
    for (int r = warp; r < n; r += RT_H) {
      const RasterTri& t = st[r];
      int bx0, by0, bx1, by1;
      raster_bbox(t, res, bx0, by0, bx1, by1);
      const int x0 = max(bx0, X0), x1 = min(bx1, X1), y0 = max(by0, Y0), y1 = min(by1, Y1);
      for (int yy = y0 + sy; yy <= y1; yy += 2)
        for (int xx = x0 + sx; xx <= x1; xx += 16) {
          const unsigned long long key = raster_shade(t, (long long)xx * SUBPIX + SUBPIX / 2,
                                                      (long long)yy * SUBPIX + SUBPIX / 2);
          if (key != ~0ull) atomicMin(&zbuf[(yy - Y0) * RT_W + (xx - X0)], key);
        }
    }
  }
  __syncthreads();
  // ---- thread = pixel: write the outputs ----
  const int lx = threadIdx.x % RT_W, ly = threadIdx.x / RT_W;
  const int x = X0 + lx, y = Y0 + ly;
  const bool inside = x < res && y < res;
  const unsigned long long best = zbuf[threadIdx.x];
  const bool hit = inside && best != ~0ull;
  if (inside) {
    const size_t p = ((size_t)v * res + y) * res + x;
    depth[p] = hit ? ordered_u32_to_float((unsigned int)(best >> 32)) : 0.0f;
    face_idx[p] = hit ? (long long)(unsigned int)(best & 0xFFFFFFFFu) : -1ll;
    mask_cam[p] = hit ? 1 : 0;
  }
  if (out_res == res) {
    if (inside) mask_out[((size_t)v * res + y) * res + x] = hit ? 1 : 0;
  } else {  // res == 2 * out_res: OR of each 2 x 2 block (RT_W, RT_H are even, so blocks stay in the tile)
    s_hit[ly][lx] = hit ? 1 : 0;
    __syncthreads();
    if (ly < RT_H / 2 && lx < RT_W / 2) {
      const int ox = tx * (RT_W / 2) + lx, oy = ty * (RT_H / 2) + ly;
      if (ox < out_res && oy < out_res)
        mask_out[((size_t)v * out_res + oy) * out_res + ox] =
            s_hit[2 * ly][2 * lx] | s_hit[2 * ly][2 * lx + 1] | s_hit[2 * ly + 1][2 * lx] |
            s_hit[2 * ly + 1][2 * lx + 1];
    }
  }
}

int rasterize_launch(const float* pos, const int* faces, int V, int Vm, int F, int res,
                     int out_res, void* workspace, float* depth, long long* face_idx,
                     uint8_t* mask_cam, uint8_t* mask_out, cudaStream_t stream) {
  PDR_CHECK_ARG(out_res == res || out_res * 2 == res,
                "mask resize %d -> %d unsupported (cam_res must equal res or 2*res)", res, out_res);
  PDR_CHECK_ARG(F > 0 && V > 0, "empty mesh");
  PDR_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "rasterize workspace must be 16-byte aligned");
  const int tiles_x = cdiv(res, RT_W), tiles_y = cdiv(res, RT_H), tiles = tiles_x * tiles_y;
  PDR_CHECK_ARG((long long)V * tiles < (1ll << 30), "raster grid too large");
  RasterWs ws = raster_ws_carve(workspace, V, F, tiles);
  // counts and large_count are adjacent: one memset
  PDR_CUDA(cudaMemsetAsync(ws.counts, 0, (uint8_t*)ws.offs - (uint8_t*)ws.counts, stream));
  raster_setup_kernel<<<cdiv((long long)V * F, 256), 256, 0, stream>>>(pos, faces, V, Vm, F, res,
                                                                      tiles_x, tiles, ws);
  PDR_COUNT_LAUNCH();
  raster_scan_kernel<<<1, 1024, 0, stream>>>(ws.counts, ws.offs, V * tiles);
  PDR_COUNT_LAUNCH();
  raster_bin_kernel<<<cdiv((long long)V * F, 256), 256, 0, stream>>>(V, F, res, tiles_x, tiles, ws);
  PDR_COUNT_LAUNCH();
  raster_tile_kernel<<<V * tiles, RT_W * RT_H, 0, stream>>>(V, F, res, out_res, tiles_x, tiles, ws,
                                                           depth, face_idx, mask_cam, mask_out);
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
