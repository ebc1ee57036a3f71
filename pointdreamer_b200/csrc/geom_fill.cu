// Exact nearest-valid-pixel fill (K8).
//
// Reference: pointdreamer/ours_utils.py:610-643 naive_inpainting(method='nearest') (scipy
// griddata -> cKDTree 1-NN on the CPU) used by texture_gen_method 'nearest'
// (ours_utils.py:930-941) and by pointdreamer/unproject.py:480-504 dilate_atlas.
// Canonical tie rule (oracle/fill.py): minimum squared distance, then lowest linear index of
// the source pixel.
//
// Two passes: (1) per column, distance to the nearest valid pixel above/below (sequential scan
// down a column, coalesced across columns); (2) per pixel, walk columns outward (dx = 0, -1, +1,
// ...) combining dx^2 with the column distances, stopping once dx^2 exceeds the best distance —
// O(distance) probes per pixel instead of a kd-tree.
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

static constexpr int FILL_BIG = 1 << 20;

__global__ void fill_column_scan_kernel(const uint8_t* __restrict__ known, int B, int H, int W,
                                        int* __restrict__ up, int* __restrict__ dn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  const int b = i / W, x = i % W;
  const uint8_t* k = known + (size_t)b * H * W;
  int* u = up + (size_t)b * H * W;
  int* d = dn + (size_t)b * H * W;
  int last = -FILL_BIG;
  for (int y = 0; y < H; ++y) {
    if (k[(size_t)y * W + x]) last = y;
    u[(size_t)y * W + x] = min(y - last, FILL_BIG);
  }
  int nxt = 3 * FILL_BIG;
  for (int y = H - 1; y >= 0; --y) {
    if (k[(size_t)y * W + x]) nxt = y;
    d[(size_t)y * W + x] = min(nxt - y, FILL_BIG);
  }
}

__global__ void fill_gather_kernel(const float* __restrict__ img, const int* __restrict__ up,
                                   const int* __restrict__ dn, int B, int C, int H, int W,
                                   long long sb, long long sc, long long sy, long long sx,
                                   float* __restrict__ out, int* __restrict__ src_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W) return;
  const int b = i / ((size_t)H * W), y = (i / W) % H, x = i % W;
  const int* u = up + (size_t)b * H * W;
  const int* d = dn + (size_t)b * H * W;
  long long best_d = LLONG_MAX;
  int best_r = -1, best_c = -1;
  for (int a = 0; a < W; ++a) {
    if ((long long)a * a > best_d) break;
    for (int s = 0; s < 2; ++s) {
      if (a == 0 && s == 1) break;
      const int c = s == 0 ? x - a : x + a;
      if (c < 0 || c >= W) continue;
      const int uu = u[(size_t)y * W + c];
      if (uu < FILL_BIG) {
        const long long dd = (long long)a * a + (long long)uu * uu;
        const int r = y - uu;
        if (dd < best_d || (dd == best_d && (r < best_r || (r == best_r && c < best_c))))
          best_d = dd, best_r = r, best_c = c;
      }
      const int dw = d[(size_t)y * W + c];
      if (dw < FILL_BIG && dw > 0) {
        const long long dd = (long long)a * a + (long long)dw * dw;
        const int r = y + dw;
        if (dd < best_d || (dd == best_d && (r < best_r || (r == best_r && c < best_c))))
          best_d = dd, best_r = r, best_c = c;
      }
    }
  }
  if (src_out) src_out[i] = best_r < 0 ? -1 : best_r * W + best_c;
  const long long ob = (long long)b * sb + (long long)y * sy + (long long)x * sx;
  for (int ch = 0; ch < C; ++ch) {
    float val = 0.f;
    if (best_r >= 0) val = img[(long long)b * sb + ch * sc + (long long)best_r * sy + best_c * sx];
    out[ob + ch * sc] = val;
  }
}

size_t nearest_fill_workspace_bytes(int B, int H, int W) {
  return (size_t)B * H * W * 2 * sizeof(int) + 256;
}

int nearest_fill_launch(const float* img, const uint8_t* known, int B, int C, int H, int W,
                        int channels_last, void* workspace, float* out, int* src_index,
                        cudaStream_t stream) {
  PDR_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "empty image");
  PDR_CHECK_ARG(img != out, "nearest fill cannot run in place");
  int* up = (int*)workspace;
  int* dn = up + (size_t)B * H * W;
  long long sb = (long long)C * H * W, sc, sy, sx;
  if (channels_last) {
    sc = 1, sx = C, sy = (long long)W * C;
  } else {
    sc = (long long)H * W, sy = W, sx = 1;
  }
  fill_column_scan_kernel<<<cdiv((long long)B * W, 128), 128, 0, stream>>>(known, B, H, W, up, dn);
  PDR_COUNT_LAUNCH();
  fill_gather_kernel<<<cdiv((long long)B * H * W, 256), 256, 0, stream>>>(
      img, up, dn, B, C, H, W, sb, sc, sy, sx, out, src_index);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
