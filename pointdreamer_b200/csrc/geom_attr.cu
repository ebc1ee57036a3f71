// Attribute interpolation over a rasterised mesh, face normals, and the fixed-crop vertex
// projection of the texture optimiser ("next" rows N1 / N4 of SURVEY.md §8f).
//
// Reference: nvdiffrast.torch.interpolate at models/get3d/extract_texture_map.py:60 (world
// position per atlas texel) and pointdreamer/ours_utils.py:1697 (texture uv per view pixel);
// kaolin.ops.mesh.face_normals at demo.py:422; the vertex rescaling of optimize_color,
// ours_utils.py:1676-1693.  nvdiffrast / kaolin are third-party and unvendored: the canonical
// rule (oracle/project.py:interpolate) follows nvdiffrast's published scheme - the rasteriser
// hands fp32 barycentrics (u, v) of vertices 0 and 1 to the interpolator, which forms
// attr = (u*a0 + v*a1) + ((1 - u) - v)*a2 - with u = fp32(eA)/fp32(eA+eB+eC),
// v = fp32(eB)/fp32(eA+eB+eC) from the rasteriser's exact int64 edge functions.
#include "geom_common.cuh"
#include "geom.h"

namespace pdr {

// thread per output pixel; face_idx < 0 -> zeros (nvdiffrast returns 0 for empty pixels).
// flip_y != 0 writes row (res-1-y) of the raster frame to output row y (the reference's
// torch.flip(..., [1]) at ours_utils.py:1702-1705) and also flips mask_out.
template <int C>
__global__ void interpolate_kernel(const float* __restrict__ pos, const int* __restrict__ faces,
                                   const long long* __restrict__ face_idx,
                                   const float* __restrict__ attr,
                                   const int* __restrict__ attr_faces, int V, int Vm, int res,
                                   int flip_y, float* __restrict__ out,
                                   uint8_t* __restrict__ mask_out) {
  const size_t n = (size_t)V * res * res;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = i % res, y = (i / res) % res, v = i / ((size_t)res * res);
  const int yr = flip_y ? res - 1 - y : y;  // raster-frame row
  const long long f = face_idx[((size_t)v * res + yr) * res + x];
  float o[C];
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = 0.f;
  if (f >= 0) {
    const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
    const float4* P = reinterpret_cast<const float4*>(pos) + (size_t)v * Vm;
    const float4 A = P[ia], B = P[ib], Cc = P[ic];
    const long long ax = snap_coord(A.x, res), ay = snap_coord(A.y, res);
    const long long bx = snap_coord(B.x, res), by = snap_coord(B.y, res);
    const long long cx = snap_coord(Cc.x, res), cy = snap_coord(Cc.y, res);
    const long long area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    const long long sgn = area > 0 ? 1 : -1;
    const long long px = (long long)x * SUBPIX + SUBPIX / 2, py = (long long)yr * SUBPIX + SUBPIX / 2;
    const long long eA = sgn * ((cx - bx) * (py - by) - (cy - by) * (px - bx));
    const long long eB = sgn * ((ax - cx) * (py - cy) - (ay - cy) * (px - cx));
    const long long eC = sgn * ((bx - ax) * (py - ay) - (by - ay) * (px - ax));
    const float wa = __ll2float_rn(eA), wb = __ll2float_rn(eB);
    const float tot = __ll2float_rn(eA + eB + eC);
    const int ja = attr_faces[3 * f], jb = attr_faces[3 * f + 1], jc = attr_faces[3 * f + 2];
    const float u = wa / tot, w = wb / tot;
    const float b2 = (1.0f - u) - w;
#pragma unroll
    for (int c = 0; c < C; ++c)
      o[c] = (u * attr[(size_t)ja * C + c] + w * attr[(size_t)jb * C + c]) +
             b2 * attr[(size_t)jc * C + c];
  }
#pragma unroll
  for (int c = 0; c < C; ++c) out[i * C + c] = o[c];
  if (mask_out) mask_out[i] = f >= 0 ? 1 : 0;
}

int interpolate_launch(const float* pos, const int* faces, const long long* face_idx,
                       const float* attr, const int* attr_faces, int V, int Vm, int res, int C,
                       int flip_y, float* out, uint8_t* mask_out, cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && res > 0, "interpolate: empty raster");
  PDR_CHECK_ARG(C == 2 || C == 3, "interpolate: %d attribute channels unsupported (2 or 3)", C);
  const size_t n = (size_t)V * res * res;
  if (C == 2)
    interpolate_kernel<2><<<cdiv(n, 256), 256, 0, stream>>>(pos, faces, face_idx, attr, attr_faces,
                                                           V, Vm, res, flip_y, out, mask_out);
  else
    interpolate_kernel<3><<<cdiv(n, 256), 256, 0, stream>>>(pos, faces, face_idx, attr, attr_faces,
                                                           V, Vm, res, flip_y, out, mask_out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// unit face normals, kal.ops.mesh.face_normals(face_vertices, unit=True) (demo.py:422):
// n = (v1 - v0) x (v2 - v0), n / max(||n||, eps)  [canonical: eps = 1e-12 like F.normalize]
__global__ void face_normals_kernel(const float* __restrict__ verts, const int* __restrict__ faces,
                                    int F, float* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
  const float ax = verts[3 * ia], ay = verts[3 * ia + 1], az = verts[3 * ia + 2];
  const float e1x = verts[3 * ib] - ax, e1y = verts[3 * ib + 1] - ay, e1z = verts[3 * ib + 2] - az;
  const float e2x = verts[3 * ic] - ax, e2y = verts[3 * ic + 1] - ay, e2z = verts[3 * ic + 2] - az;
  const float nx = e1y * e2z - e1z * e2y;
  const float ny = e1z * e2x - e1x * e2z;
  const float nz = e1x * e2y - e1y * e2x;
  const float len = sqrtf((nx * nx + ny * ny) + nz * nz);
  const float d = fmaxf(len, 1e-12f);
  out[3 * f] = nx / d;
  out[3 * f + 1] = ny / d;
  out[3 * f + 2] = nz / d;
}

int face_normals_launch(const float* verts, const int* faces, int F, float* out,
                        cudaStream_t stream) {
  PDR_CHECK_ARG(F > 0, "face_normals: empty mesh");
  face_normals_kernel<<<cdiv(F, 256), 256, 0, stream>>>(verts, faces, F, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// optimize_color's per-view clip-space vertices (ours_utils.py:1676-1693): transform, then
// ((uv - c) / s) * (1 - 2*padding) * inpaint_scale + 0.5, clip to [0,1], * 2 - 1; z kept, w = 1.
__global__ void project_fixed_kernel(const float* __restrict__ cams,
                                     const float* __restrict__ verts, int Vm, float pad_mul,
                                     const float* __restrict__ centers,
                                     const float* __restrict__ scales,
                                     const float* __restrict__ inpaint_scales,
                                     float* __restrict__ pos) {
  __shared__ float sp[CAM_PARAM_FLOATS];
  const int v = blockIdx.y;
  if (threadIdx.x < CAM_PARAM_FLOATS) sp[threadIdx.x] = cams[v * CAM_PARAM_FLOATS + threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Vm) return;
  const float cx = centers[v * 2], cy = centers[v * 2 + 1], sc = scales[v], is = inpaint_scales[v];
  float nx, ny, nz;
  cam_transform(sp, verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], nx, ny, nz);
  const float u = clipf((((nx - cx) / sc) * pad_mul) * is + 0.5f, 0.f, 1.f);
  const float w = clipf((((ny - cy) / sc) * pad_mul) * is + 0.5f, 0.f, 1.f);
  reinterpret_cast<float4*>(pos)[(size_t)v * Vm + i] =
      make_float4(u * 2.0f - 1.0f, w * 2.0f - 1.0f, nz, 1.0f);
}

int project_fixed_launch(const float* cams, const float* vertices, int Vm, int V, double padding,
                         const float* centers, const float* scales, const float* inpaint_scales,
                         float* pos, cudaStream_t stream) {
  PDR_CHECK_ARG(V > 0 && V <= MAX_VIEWS && Vm > 0, "project_fixed: bad sizes");
  const float pad_mul = (float)(1.0 - 2.0 * padding);
  project_fixed_kernel<<<dim3(cdiv(Vm, 256), V), 256, 0, stream>>>(cams, vertices, Vm, pad_mul,
                                                                    centers, scales,
                                                                    inpaint_scales, pos);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// 8-bit atlas as demo.py:283-301 writes it: (img - 0) * 255 in fp32, clip [0,255], truncate to
// uint8, rows flipped; optional RGBA copy whose alpha is mask*255.  Quantising on the device cuts
// the device->host copy of the result 4x (12.6 MB -> 3.1 MB at R = 1024).
__global__ void atlas_to_u8_kernel(const float* __restrict__ atlas, const uint8_t* __restrict__ mask,
                                   int R, uint8_t* __restrict__ rgb, uint8_t* __restrict__ rgba) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * R) return;
  const int x = i % R, y = i / R;
  const size_t src = ((size_t)(R - 1 - y) * R + x);
  uint8_t q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = clipf(atlas[src * 3 + c] * 255.0f, 0.f, 255.f);
    q[c] = (uint8_t)v;  // astype(np.uint8) of a value in [0,255]: truncation (NaN -> 0)
  }
  rgb[i * 3] = q[0], rgb[i * 3 + 1] = q[1], rgb[i * 3 + 2] = q[2];
  if (rgba) {
    rgba[i * 4] = q[0], rgba[i * 4 + 1] = q[1], rgba[i * 4 + 2] = q[2];
    rgba[i * 4 + 3] = (mask && mask[src]) ? 255 : 0;
  }
}

int atlas_to_u8_launch(const float* atlas, const uint8_t* mask, int R, uint8_t* rgb, uint8_t* rgba,
                       cudaStream_t stream) {
  PDR_CHECK_ARG(R > 0, "atlas_to_u8: bad size");
  atlas_to_u8_kernel<<<cdiv((long long)R * R, 256), 256, 0, stream>>>(atlas, mask, R, rgb, rgba);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
