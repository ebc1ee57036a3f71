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
static constexpr int HPR_TILE = 256;

// frames: [V][12] doubles = eye(3), ex(3), ey(3), ez(3)
__global__ void hpr_prepare_kernel(const float* __restrict__ points, int N, int V,
                                   const double* __restrict__ frames, double radius,
                                   double* __restrict__ U, double* __restrict__ Vv,
                                   double* __restrict__ Wt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N, n = i % N;
  const double* f = frames + v * 12;
  const double px = (double)points[3 * n] - f[0], py = (double)points[3 * n + 1] - f[1],
               pz = (double)points[3 * n + 2] - f[2];
  const double nrm = sqrt(px * px + py * py + pz * pz);
  const double s = 2.0 * radius / nrm - 1.0;
  const double x = px * f[3] + py * f[4] + pz * f[5];
  const double y = px * f[6] + py * f[7] + pz * f[8];
  const double z = px * f[9] + py * f[10] + pz * f[11];
  U[i] = x / z;
  Vv[i] = y / z;
  Wt[i] = -HPR_WSCALE / (s * z);
}

__device__ __forceinline__ int hpr_perm(int k, int N, int stride, int offset) {
  return (int)(((long long)k * stride + offset) % N);
}

__global__ void __launch_bounds__(128)
hpr_lp_kernel(const double* __restrict__ U, const double* __restrict__ Vv,
              const double* __restrict__ Wt, int N, int stride, int offset,
              uint8_t* __restrict__ vis) {
  __shared__ double su[HPR_TILE], sv[HPR_TILE], sw[HPR_TILE];
  __shared__ int sj[HPR_TILE];
  const int v = blockIdx.y;
  const double* u = U + (size_t)v * N;
  const double* vv = Vv + (size_t)v * N;
  const double* w = Wt + (size_t)v * N;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active0 = i < N;
  const double ui = active0 ? u[i] : 0.0, vi = active0 ? vv[i] : 0.0, wi = active0 ? w[i] : 0.0;
  // maximise c.x with c = (1, 0.5) inside the box: start at the (+,+) corner
  double a = HPR_BOX, b = HPR_BOX;
  bool feasible = active0;
  const double c0 = 1.0, c1 = 0.5;

  for (int base = 0; base < N; base += HPR_TILE) {
    __syncthreads();
    for (int t = threadIdx.x; t < HPR_TILE && base + t < N; t += blockDim.x) {
      const int j = hpr_perm(base + t, N, stride, offset);
      sj[t] = j;
      su[t] = u[j];
      sv[t] = vv[j];
      sw[t] = w[j];
    }
    __syncthreads();
    if (!__any_sync(0xffffffffu, feasible)) continue;  // warp finished; keep hitting the barriers
    const int cnt = min(HPR_TILE, N - base);
    for (int t = 0; t < cnt; ++t) {
      const int j = sj[t];
      const double du = su[t] - ui, dv = sv[t] - vi, dw = sw[t] - wi;
      const bool viol = feasible && j != i && (du * a + dv * b < dw);
      unsigned m = __ballot_sync(0xffffffffu, viol);
      while (m) {
        const int L = __ffs(m) - 1;
        m &= m - 1;
        // everything about lane L's sub-problem, broadcast to the warp
        const double uL = __shfl_sync(0xffffffffu, ui, L), vL = __shfl_sync(0xffffffffu, vi, L),
                     wL = __shfl_sync(0xffffffffu, wi, L);
        const int iL = __shfl_sync(0xffffffffu, i, L);
        const double nx = su[t] - uL, ny = sv[t] - vL, h = sw[t] - wL;
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
          // all earlier constraints (positions < base + t), split over the lanes
          const int pos = base + t;
          for (int k = lane; k < pos; k += 32) {
            const int jk = hpr_perm(k, N, stride, offset);
            if (jk == iL) continue;
            const double kx = u[jk] - uL, ky = vv[jk] - vL, kh = w[jk] - wL;
            const double den = kx * dx + ky * dy;
            const double rhs = kh - (kx * p0x + ky * p0y);
            if (den > 0.0)
              lo = fmax(lo, rhs / den);
            else if (den < 0.0)
              hi = fmin(hi, rhs / den);
            else if (rhs > 0.0)
              lo = INFINITY;
          }
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
  if (active0) vis[(size_t)v * N + i] = feasible ? 1 : 0;
}

size_t hpr_workspace_bytes(int V, int N) { return (size_t)V * N * 3 * sizeof(double) + 256; }

int hpr_launch(const float* points, int N, int V, const double* frames_dev, double radius,
               void* workspace, uint8_t* vis, cudaStream_t stream) {
  PDR_CHECK_ARG(N > 0 && V > 0, "hidden point removal: empty input");
  double* U = (double*)workspace;
  double* Vv = U + (size_t)V * N;
  double* Wt = Vv + (size_t)V * N;
  hpr_prepare_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(points, N, V, frames_dev, radius,
                                                                  U, Vv, Wt);
  PDR_COUNT_LAUNCH();
  // visiting order of the constraints: k -> (k*stride + offset) mod N, stride coprime with N
  static const int primes[] = {7919, 104729, 1299709, 15485863, 32452843};
  int stride = 1;
  for (int p : primes)
    if (N % p != 0 && p % N != 0) {
      stride = p % N;
      break;
    }
  // make sure gcd(stride, N) == 1
  auto gcd = [](int x, int y) {
    while (y) {
      int t = x % y;
      x = y;
      y = t;
    }
    return x;
  };
  while (stride < 1 || gcd(stride, N) != 1) stride = (stride + 1) % N == 0 ? 1 : stride + 1;
  hpr_lp_kernel<<<dim3(cdiv(N, 128), V), 128, 0, stream>>>(U, Vv, Wt, N, stride, N / 3, vis);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
