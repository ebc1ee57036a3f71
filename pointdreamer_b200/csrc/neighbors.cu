// Completion of never-seen texels from mesh neighbours ("next" row N2 of SURVEY.md §8f).
//
// Reference: pointdreamer/unproject.py:93-196 paint_invisible_areas_by_neighbors (use_atlas=True),
// :17-38 compute_vertex_only_uv_mask.  The reference builds a DENSE V x V uniform Laplacian
// (kaolin) and multiplies it by the colour matrix every round; here the same rows are a CSR
// adjacency walked in ascending neighbour order (oracle/neighbors.py states the canonical fp32
// summation order), one thread per vertex, Jacobi double buffering.  Kernels:
//   vertex_uv        : per face corner, atomicMax of the uv index per vertex (the pair that sorts
//                      last wins the reference's duplicate index_put)
//   vertex_colors    : uv -> atlas pixel (clip(uv*R, 0, R-1).long(), swapped to row/col), colour
//                      and "has colour" gathers
//   laplacian_round  : one round of unproject.py:160-163 for every never-coloured vertex, plus
//                      the number of coloured vertices (integer atomicAdd)
//   scatter          : colours back to the atlas, highest vertex index wins a shared texel
#include "geom_common.cuh"
#include "geom.h"

namespace pdr {

__global__ void fill_int_kernel(int* p, size_t n, int v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void vertex_uv_kernel(const int* __restrict__ faces, const int* __restrict__ face_uv,
                                 int n_corners, int* __restrict__ uv_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_corners) return;
  atomicMax(&uv_idx[faces[i]], face_uv[i]);
}

__global__ void vertex_colors_kernel(const int* __restrict__ uv_idx, const float* __restrict__ uvs,
                                     int Vn, const float* __restrict__ atlas,
                                     const uint8_t* __restrict__ mask, int R,
                                     long long* __restrict__ pix, float* __restrict__ colors,
                                     float* __restrict__ count, uint8_t* __restrict__ has_color) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= Vn) return;
  const int u = uv_idx[v];
  float ux = 0.f, uy = 0.f;  // vertices no face references keep uv (0,0) (unproject.py:120)
  if (u >= 0) {
    ux = uvs[2 * u];
    uy = uvs[2 * u + 1];
  }
  const long long px = (long long)clipf(ux * (float)R, 0.f, (float)(R - 1));
  const long long py = (long long)clipf(uy * (float)R, 0.f, (float)(R - 1));
  const long long p = py * R + px;  // (row = y, col = x)
  pix[v] = p;
  colors[3 * v] = atlas[3 * p];
  colors[3 * v + 1] = atlas[3 * p + 1];
  colors[3 * v + 2] = atlas[3 * p + 2];
  const uint8_t h = mask[p] ? 1 : 0;
  has_color[v] = h;
  count[v] = h ? 1.f : 0.f;
}

__global__ void laplacian_round_kernel(const int* __restrict__ rowptr,
                                       const int* __restrict__ colidx, int Vn,
                                       const uint8_t* __restrict__ fixed,
                                       const float* __restrict__ c_in,
                                       const float* __restrict__ n_in, float* __restrict__ c_out,
                                       float* __restrict__ n_out, int* __restrict__ total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int colored = 0;
  if (i < Vn) {
    if (fixed[i]) {
      colored = 1;  // c_out / n_out of fixed vertices were initialised once and never change
    } else {
      const int j0 = rowptr[i], j1 = rowptr[i + 1];
      const float w = j1 > j0 ? 1.0f / (float)(j1 - j0) : 0.f;
      float nc0 = 0.f, nc1 = 0.f, nc2 = 0.f, nn = 0.f;
      for (int k = j0; k < j1; ++k) {
        const int j = colidx[k];
        const float cnt = n_in[j];
        nc0 = nc0 + w * (c_in[3 * j] * cnt);
        nc1 = nc1 + w * (c_in[3 * j + 1] * cnt);
        nc2 = nc2 + w * (c_in[3 * j + 2] * cnt);
        nn = nn + w * cnt;
      }
      if (nn > 0.f) {
        c_out[3 * i] = nc0 / nn;
        c_out[3 * i + 1] = nc1 / nn;
        c_out[3 * i + 2] = nc2 / nn;
        n_out[i] = 1.f;
        colored = 1;
      } else {
        c_out[3 * i] = c_in[3 * i];
        c_out[3 * i + 1] = c_in[3 * i + 1];
        c_out[3 * i + 2] = c_in[3 * i + 2];
        n_out[i] = 0.f;
      }
    }
  }
  const unsigned int ballot = __ballot_sync(0xffffffffu, colored);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(total, __popc(ballot));
}

__global__ void scatter_winner_kernel(const long long* __restrict__ pix, int Vn,
                                      int* __restrict__ winner) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < Vn) atomicMax(&winner[pix[v]], v);
}
__global__ void scatter_write_kernel(const long long* __restrict__ pix,
                                     const float* __restrict__ colors, int Vn,
                                     const int* __restrict__ winner, float* __restrict__ atlas,
                                     uint8_t* __restrict__ mask) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= Vn) return;
  const long long p = pix[v];
  if (winner[p] != v) return;
  atlas[3 * p] = colors[3 * v];
  atlas[3 * p + 1] = colors[3 * v + 1];
  atlas[3 * p + 2] = colors[3 * v + 2];
  mask[p] = 1;
}

int vertex_colors_launch(const int* faces, const int* face_uv_idx, int F, const float* uvs, int Vn,
                         const float* atlas, const uint8_t* mask, int R, int* ws_uv_idx,
                         long long* pix, float* colors, float* count, uint8_t* has_color,
                         cudaStream_t stream) {
  PDR_CHECK_ARG(F > 0 && Vn > 0 && R > 0, "vertex_colors: bad sizes");
  fill_int_kernel<<<cdiv(Vn, 256), 256, 0, stream>>>(ws_uv_idx, (size_t)Vn, -1);
  PDR_COUNT_LAUNCH();
  vertex_uv_kernel<<<cdiv(3 * (long long)F, 256), 256, 0, stream>>>(faces, face_uv_idx, 3 * F,
                                                                   ws_uv_idx);
  PDR_COUNT_LAUNCH();
  vertex_colors_kernel<<<cdiv(Vn, 256), 256, 0, stream>>>(ws_uv_idx, uvs, Vn, atlas, mask, R, pix,
                                                         colors, count, has_color);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int laplacian_round_launch(const int* rowptr, const int* colidx, int Vn, const uint8_t* fixed,
                           const float* colors_in, const float* count_in, float* colors_out,
                           float* count_out, int* colored_total, cudaStream_t stream) {
  PDR_CHECK_ARG(Vn > 0, "laplacian_round: empty mesh");
  PDR_CUDA(cudaMemsetAsync(colored_total, 0, sizeof(int), stream));
  laplacian_round_kernel<<<cdiv(Vn, 256), 256, 0, stream>>>(rowptr, colidx, Vn, fixed, colors_in,
                                                           count_in, colors_out, count_out,
                                                           colored_total);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

int scatter_vertex_colors_launch(const long long* pix, const float* colors, int Vn, int R,
                                 int* ws_winner, float* atlas, uint8_t* mask,
                                 cudaStream_t stream) {
  PDR_CHECK_ARG(Vn > 0 && R > 0, "scatter_vertex_colors: bad sizes");
  const size_t RR = (size_t)R * R;
  fill_int_kernel<<<cdiv(RR, 256), 256, 0, stream>>>(ws_winner, RR, -1);
  PDR_COUNT_LAUNCH();
  scatter_winner_kernel<<<cdiv(Vn, 256), 256, 0, stream>>>(pix, Vn, ws_winner);
  PDR_COUNT_LAUNCH();
  scatter_write_kernel<<<cdiv(Vn, 256), 256, 0, stream>>>(pix, colors, Vn, ws_winner, atlas, mask);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
