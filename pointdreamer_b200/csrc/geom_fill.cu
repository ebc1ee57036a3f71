// Exact nearest-valid-pixel fill (K8).
//
// Reference: pointdreamer/ours_utils.py:610-643 naive_inpainting(method='nearest') (scipy
// griddata -> cKDTree 1-NN on the CPU) used by texture_gen_method 'nearest'
// (ours_utils.py:930-941) and by pointdreamer/unproject.py:480-504 dilate_atlas.
// Canonical tie rule (oracle/fill.py): minimum squared distance, then lowest linear index of
// the source pixel.
//
// Two passes (an exact Euclidean distance transform with source tracking):
//  (1) fill_column_kernel: per column, the SIGNED offset to the nearest valid pixel of that column
//      (negative = above; on a tie the pixel above wins because it has the lower linear index), as
//      one int16 per pixel.  A thread owns 32 consecutive rows of one column as a 32-bit mask and
//      the per-column carries of the other segments go through shared memory, so the pass is fully
//      parallel (no serial walk down the column) and its loads/stores are coalesced across columns.
//  (2) fill_row_kernel: per pixel, min over columns c of (x-c)^2 + off(c)^2.  The row of offsets is
//      staged in SHARED memory once per CTA; each thread walks outwards (dx = 0, -1, +1, ...) and
//      stops once dx^2 exceeds the best distance - O(distance) shared-memory probes per pixel.
// One candidate per column is enough: if the nearest pixels above and below are equally far the
// upper one wins every tie-break, otherwise the nearer one strictly beats the other.
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

static constexpr int FILL_NONE = 0x7FFF;   // int16 offset sentinel: the column has no valid pixel
static constexpr int FILL_COLS = 32;       // columns per CTA in pass 1
static constexpr int FILL_MAX_H = 8192;    // offsets must fit int16

// pass 1.  grid = (ceil(W / 32), B), block = (32 columns, 32 segment slots); segment = 32 rows.
__global__ void __launch_bounds__(1024)
fill_column_kernel(const uint8_t* __restrict__ known, int H, int W, short* __restrict__ off) {
  extern __shared__ int s_carry[];  // [2][nseg][32]: last valid row of the segment, first valid row
  const int nseg = (H + 31) / 32;
  int* s_last = s_carry;
  int* s_first = s_carry + nseg * FILL_COLS;
  const int b = blockIdx.y;
  const int x = blockIdx.x * FILL_COLS + threadIdx.x;
  const uint8_t* k = known + (size_t)b * H * W;
  short* o = off + (size_t)b * H * W;
  const bool col_ok = x < W;
  // (a) build the 32-row masks
  for (int seg = threadIdx.y; seg < nseg; seg += blockDim.y) {
    unsigned m = 0;
    if (col_ok) {
      const int y0 = seg * 32;
#pragma unroll 8
      for (int r = 0; r < 32; ++r)
        if (y0 + r < H && k[(size_t)(y0 + r) * W + x]) m |= 1u << r;
    }
    s_last[seg * FILL_COLS + threadIdx.x] = m ? seg * 32 + 31 - __clz(m) : -1;
    s_first[seg * FILL_COLS + threadIdx.x] = m ? seg * 32 + __ffs(m) - 1 : -1;
  }
  __syncthreads();
  // (b) per column: nearest valid row strictly above / below each segment (serial over <= 256
  //     segments, one thread per column; replaces the carries in place)
  if (threadIdx.y == 0) {
    int last = -1;
    for (int seg = 0; seg < nseg; ++seg) {
      const int mine = s_last[seg * FILL_COLS + threadIdx.x];
      s_last[seg * FILL_COLS + threadIdx.x] = last;  // last valid row ABOVE this segment
      if (mine >= 0) last = mine;
    }
    int next = -1;
    for (int seg = nseg - 1; seg >= 0; --seg) {
      const int mine = s_first[seg * FILL_COLS + threadIdx.x];
      s_first[seg * FILL_COLS + threadIdx.x] = next;  // first valid row BELOW this segment
      if (mine >= 0) next = mine;
    }
  }
  __syncthreads();
  // (c) offsets of the 32 rows of each segment (mask re-read: it is an L1/L2 hit)
  for (int seg = threadIdx.y; seg < nseg; seg += blockDim.y) {
    if (!col_ok) continue;
    const int y0 = seg * 32;
    unsigned m = 0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r)
      if (y0 + r < H && k[(size_t)(y0 + r) * W + x]) m |= 1u << r;
    const int above = s_last[seg * FILL_COLS + threadIdx.x];
    const int below = s_first[seg * FILL_COLS + threadIdx.x];
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const int y = y0 + r;
      if (y >= H) break;
      const unsigned lo = m & (0xFFFFFFFFu >> (31 - r));  // rows <= r of the segment
      const unsigned hi = m >> r;                          // rows >= r
      const int up = lo ? y - (y0 + 31 - __clz(lo)) : (above >= 0 ? y - above : INT_MAX);
      const int dn = hi ? __ffs(hi) - 1 : (below >= 0 ? below - y : INT_MAX);
      int v = FILL_NONE;
      if (up != INT_MAX || dn != INT_MAX) v = up <= dn ? -up : dn;
      o[(size_t)y * W + x] = (short)v;
    }
  }
}

// pass 2.  grid = (ceil(W / 256), H, B): a CTA owns 256 consecutive pixels of one row and stages the
// whole row of column offsets in shared memory.
__global__ void __launch_bounds__(256)
fill_row_kernel(const float* __restrict__ img, const short* __restrict__ off, int C, int H, int W,
                long long sb, long long sc, long long sy, long long sx, float* __restrict__ out,
                int* __restrict__ src_out) {
  extern __shared__ short s_off[];  // [W]
  const int b = blockIdx.z, y = blockIdx.y;
  const short* orow = off + ((size_t)b * H + y) * W;
  for (int c = threadIdx.x; c < W; c += blockDim.x) s_off[c] = orow[c];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  long long best_d = LLONG_MAX;
  int best_r = -1, best_c = -1;
  for (int a = 0; a < W; ++a) {
    if ((long long)a * a > best_d) break;
    for (int s = 0; s < 2; ++s) {
      if (a == 0 && s == 1) break;
      const int c = s == 0 ? x - a : x + a;
      if (c < 0 || c >= W) continue;
      const int o = s_off[c];
      if (o == FILL_NONE) continue;
      const long long dd = (long long)a * a + (long long)o * o;
      const int r = y + o;
      if (dd < best_d || (dd == best_d && (r < best_r || (r == best_r && c < best_c))))
        best_d = dd, best_r = r, best_c = c;
    }
  }
  const size_t i = ((size_t)b * H + y) * W + x;
  if (src_out) src_out[i] = best_r < 0 ? -1 : best_r * W + best_c;
  const long long ob = (long long)b * sb + (long long)y * sy + (long long)x * sx;
  for (int ch = 0; ch < C; ++ch) {
    float val = 0.f;
    if (best_r >= 0) val = img[(long long)b * sb + ch * sc + (long long)best_r * sy + best_c * sx];
    out[ob + ch * sc] = val;
  }
}

size_t nearest_fill_workspace_bytes(int B, int H, int W) {
  return (size_t)B * H * W * sizeof(short) + 256;
}

int nearest_fill_launch(const float* img, const uint8_t* known, int B, int C, int H, int W,
                        int channels_last, void* workspace, float* out, int* src_index,
                        cudaStream_t stream) {
  PDR_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "empty image");
  PDR_CHECK_ARG(H <= FILL_MAX_H && W <= 16384, "nearest fill: image larger than %d x 16384",
                FILL_MAX_H);
  PDR_CHECK_ARG(img != out, "nearest fill cannot run in place");
  short* off = (short*)workspace;
  long long sb = (long long)C * H * W, sc, sy, sx;
  if (channels_last) {
    sc = 1, sx = C, sy = (long long)W * C;
  } else {
    sc = (long long)H * W, sy = W, sx = 1;
  }
  const int nseg = (H + 31) / 32;
  fill_column_kernel<<<dim3(cdiv(W, FILL_COLS), B), dim3(FILL_COLS, 32),
                       2 * nseg * FILL_COLS * sizeof(int), stream>>>(known, H, W, off);
  PDR_COUNT_LAUNCH();
  fill_row_kernel<<<dim3(cdiv(W, 256), H, B), 256, W * sizeof(short), stream>>>(
      img, off, C, H, W, sb, sc, sy, sx, out, src_index);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
