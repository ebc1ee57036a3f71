// Hidden point removal on the GPU (K5) — replaces open3d's PointCloud.hidden_point_removal as
// called by pointdreamer/ours_utils.py:204-225 (Katz et al.: spherical flip + convex hull, visible
// points = hull vertices; the reference runs float64 Qhull on the CPU once per view).
//
// Instead of building a hull, every point is tested for being a hull VERTEX directly:
//   * the flipped points q_i = s_i p_i' (s_i = 2R/|p_i'| - 1, p_i' = p_i - eye) all lie in the half
//     space in front of the eye; the projective map T(q) = (q.ex/q.ez, q.ey/q.ez, -1/q.ez) sends the
//     eye (the extra hull point) to infinity, so the hull of {q_i} U {eye} becomes the UPPER hull of
//     the points (u_i, v_i, w_i) = T(q_i);
//   * point i is a vertex of that upper hull iff a plane through it keeps every other point on or
//     below:  exists (a,b):  a (u_j-u_i) + b (v_j-v_i) >= w_j - w_i  for all j  — a 2-variable LP
//     (checked against Qhull: identical vertex sets, tests/test_hpr_*.py).
// Each lane owns one point and streams through all constraints (Seidel's incremental LP, constraints
// visited in a pseudo-random order); when a lane's optimum is cut off, the whole warp re-solves its
// 1-D LP on the new constraint's line cooperatively.  Everything is fp64 like the reference.
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

static constexpr double HPR_WSCALE = 1048576.0;        // 2^20: exact rescale of w
static constexpr double HPR_BOX = 1073741824.0;        // |a|,|b| <= 2^30 (slope cap)
static constexpr int HPR_TILE = 128;

// frames: [V][12] doubles = eye(3), ex(3), ey(3), ez(3).
// Output Q[v][k] = (u, v, w, 0) of point perm(k): the constraints are stored in the pseudo-random
// visiting order, so every scan below is a sequential, fully coalesced 32-byte-per-lane stream.
__device__ __forceinline__ int hpr_perm(int k, int N, int stride, int offset) {
  return (int)(((long long)k * stride + offset) % N);
}

__global__ void hpr_prepare_kernel(const float* __restrict__ points, int N, int V,
                                   const double* __restrict__ frames, double radius, int stride,
                                   int offset, double4* __restrict__ Q) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N, k = i % N;
  const int n = hpr_perm(k, N, stride, offset);
  const double* f = frames + v * 12;
  const double px = (double)points[3 * n] - f[0], py = (double)points[3 * n + 1] - f[1],
               pz = (double)points[3 * n + 2] - f[2];
  const double nrm = sqrt(px * px + py * py + pz * pz);
  const double s = 2.0 * radius / nrm - 1.0;
  const double x = px * f[3] + py * f[4] + pz * f[5];
  const double y = px * f[6] + py * f[7] + pz * f[8];
  const double z = px * f[9] + py * f[10] + pz * f[11];
  Q[i] = make_double4(x / z, y / z, -HPR_WSCALE / (s * z), 0.0);
}

// tighten [lo, hi] on the line p0 + t d with one earlier constraint; divisions only when the bound
// actually moves (rare), comparisons by cross-multiplication otherwise
__device__ __forceinline__ void hpr_clip(const double4 c, double uL, double vL, double wL,
                                         double p0x, double p0y, double dx, double dy, double& lo,
                                         double& hi) {
  const double ax = c.x - uL, ay = c.y - vL, ah = c.z - wL;
  const double den = ax * dx + ay * dy;
  const double rhs = ah - (ax * p0x + ay * p0y);
  if (den > 0.0) {
    if (rhs > lo * den) lo = rhs / den;
  } else if (den < 0.0) {
    if (rhs > hi * den) hi = rhs / den;  // rhs/den < hi  <=>  rhs > hi*den  (den < 0)
  } else if (rhs > 0.0) {
    lo = INFINITY;
  }
}

// one warp per block: warps re-solve at very different times, so nothing may couple them
__global__ void __launch_bounds__(32)
hpr_lp_kernel(const double4* __restrict__ Q, int N, int stride, int offset,
              uint8_t* __restrict__ vis) {
  __shared__ double4 sq[HPR_TILE];
  const int v = blockIdx.y;
  const double4* q = Q + (size_t)v * N;
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x * 32 + lane;  // this lane's point, as a POSITION in visiting order
  const bool active0 = pi < N;
  const double4 me = active0 ? q[pi] : make_double4(0.0, 0.0, 0.0, 0.0);
  const double ui = me.x, vi = me.y, wi = me.z;
  // maximise c.x with c = (1, 0.5) inside the box: start at the (+,+) corner
  double a = HPR_BOX, b = HPR_BOX;
  bool feasible = active0;
  const double c0 = 1.0, c1 = 0.5;

  for (int base = 0; base < N; base += HPR_TILE) {
    if (!__any_sync(0xffffffffu, feasible)) break;  // every point of this warp is decided
    __syncwarp();
    for (int t = lane; t < HPR_TILE && base + t < N; t += 32) sq[t] = q[base + t];
    __syncwarp();
    const int cnt = min(HPR_TILE, N - base);
    for (int t = 0; t < cnt; ++t) {
      const double4 cj = sq[t];
      const double du = cj.x - ui, dv = cj.y - vi, dw = cj.z - wi;
      const bool viol = feasible && (base + t) != pi && (du * a + dv * b < dw);
      unsigned m = __ballot_sync(0xffffffffu, viol);
      while (m) {
        const int L = __ffs(m) - 1;
        m &= m - 1;
        // everything about lane L's sub-problem, broadcast to the warp
        const double uL = __shfl_sync(0xffffffffu, ui, L), vL = __shfl_sync(0xffffffffu, vi, L),
                     wL = __shfl_sync(0xffffffffu, wi, L);
        const int pL = __shfl_sync(0xffffffffu, pi, L);
        const double nx = cj.x - uL, ny = cj.y - vL, h = cj.z - wL;
        const double nn = nx * nx + ny * ny;
        double lo = -INFINITY, hi = INFINITY;
        double p0x = 0.0, p0y = 0.0, dx = 0.0, dy = 0.0;
        bool ok = nn > 0.0;  // a point exactly above in the same direction: infeasible
        if (ok) {
          const double sc = h / nn;
          p0x = nx * sc, p0y = ny * sc;
          dx = -ny, dy = nx;
          // box |p0 + t d| <= BOX
          if (dx != 0.0) {
            const double t1 = (-HPR_BOX - p0x) / dx, t2 = (HPR_BOX - p0x) / dx;
            lo = fmax(lo, fmin(t1, t2));
            hi = fmin(hi, fmax(t1, t2));
          } else if (fabs(p0x) > HPR_BOX) {
            ok = false;
          }
          if (dy != 0.0) {
            const double t1 = (-HPR_BOX - p0y) / dy, t2 = (HPR_BOX - p0y) / dy;
            lo = fmax(lo, fmin(t1, t2));
            hi = fmin(hi, fmax(t1, t2));
          } else if (fabs(p0y) > HPR_BOX) {
            ok = false;
          }
          // all earlier constraints (positions < base + t), split over the lanes, 4 loads in flight
          const int pos = base + t;
          int k = lane;
          for (; k + 96 < pos; k += 128) {
            double4 c4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) c4[r] = q[k + 32 * r];
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (k + 32 * r != pL) hpr_clip(c4[r], uL, vL, wL, p0x, p0y, dx, dy, lo, hi);
          }
          for (; k < pos; k += 32)
            if (k != pL) hpr_clip(q[k], uL, vL, wL, p0x, p0y, dx, dy, lo, hi);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmin(hi, __shfl_xor_sync(0xffffffffu, hi, o));
          }
          ok = ok && lo <= hi;
        }
        if (lane == L) {
          if (!ok) {
            feasible = false;
          } else {
            const double tt = (c0 * dx + c1 * dy > 0.0) ? hi : lo;
            a = p0x + tt * dx;
            b = p0y + tt * dy;
          }
        }
      }
    }
  }
  if (active0) vis[(size_t)v * N + hpr_perm(pi, N, stride, offset)] = feasible ? 1 : 0;
}

size_t hpr_workspace_bytes(int V, int N) { return (size_t)V * N * sizeof(double4) + 256; }

int hpr_launch(const float* points, int N, int V, const double* frames_dev, double radius,
               void* workspace, uint8_t* vis, cudaStream_t stream) {
  PDR_CHECK_ARG(N > 0 && V > 0, "hidden point removal: empty input");
  PDR_CHECK_ARG(((uintptr_t)workspace & 31) == 0, "hidden point removal: workspace must be 32-byte aligned");
  double4* Q = (double4*)workspace;
  // visiting order of the constraints: k -> (k*stride + offset) mod N, stride coprime with N
  auto gcd = [](long long x, long long y) {
    while (y) {
      long long t = x % y;
      x = y;
      y = t;
    }
    return x;
  };
  int stride = 1;
  for (int p : {7919, 104729, 1299709, 15485863, 32452843}) {
    if (gcd(p % N, N) == 1 && p % N > 1) {
      stride = p % N;
      break;
    }
  }
  const int offset = N / 3;
  hpr_prepare_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(points, N, V, frames_dev, radius,
                                                                  stride, offset, Q);
  PDR_COUNT_LAUNCH();
  hpr_lp_kernel<<<dim3(cdiv(N, 32), V), 32, 0, stream>>>(Q, N, stride, offset, vis);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
