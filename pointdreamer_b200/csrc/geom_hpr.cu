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
//     (checked against Qhull: identical vertex sets, tests/test_hpr_*.py), solved with Seidel's
//     incremental algorithm: one lane per point streams through the constraints; when a lane's
//     optimum is cut off the whole warp re-solves its 1-D LP on the new constraint's line.
//
// Constraints from points that are not hull vertices are redundant (they are convex combinations of
// vertices), and constraints from ANY subset of the cloud are necessary conditions.  So:
//   1. FILTER: a G x G grid over (u, v); the highest point of every cell is an "extreme" E (<= G^2
//      of them).  Every point runs its LP against E only, with E resident in shared memory:
//      infeasible => certainly hidden.  (28 % of a 30k cloud survive per view.)
//   2. EXACT: the survivors C (E first, then the others in pseudo-random order) run the LP against
//      C only, CONTINUING from the optimum the filter left (Seidel's invariant holds: E is the
//      prefix of the constraint sequence), i.e. ~2 ln(|C|/|E|) instead of ~2 ln N re-solves.
// Work drops from V N^2 to V (N |E| + |C|^2) constraint checks, all in fp64 like the reference.
#include "geom_common.cuh"
#include <limits.h>
#include "geom.h"

namespace pdr {

static constexpr double HPR_WSCALE = 1048576.0;        // 2^20: exact rescale of w
static constexpr double HPR_BOX = 1073741824.0;        // |a|,|b| <= 2^30 (slope cap)
static constexpr int HPR_TILE = 128;
static constexpr int HPR_G = 40;                       // filter grid (G^2 = 1600 cells, 50 KB of E)
static constexpr int HPR_G2 = HPR_G * HPR_G;
static constexpr int HPR_CELL_STRIDE = 1543;           // prime, coprime with G^2: visiting order of cells
static constexpr int HPR_FILTER_WARPS = 8;

struct HprWs {
  double4* Q;        // [V][N]  (u, v, w, index)
  double4* E;        // [V][G2] extremes, visiting order
  double4* C;        // [V][N]  survivors: E first, then the rest
  double2* ab;       // [V][N]  filter optimum per point
  double2* Cab;      // [V][N]  the same, in C order
  long long* bbox;   // [V][4]  ordered keys: min u, min v, max u, max v
  long long* cellmax;  // [V][G2]
  int* cellidx;      // [V][G2]
  int* nE;           // [V]
  int* nC;           // [V]
  uint8_t* surv;     // [V][N]
  uint8_t* isE;      // [V][N]
};

static inline size_t hpr_al(size_t x) { return (x + 255) & ~(size_t)255; }

size_t hpr_workspace_bytes(int V, int N) {
  const size_t vn = (size_t)V * N, vg = (size_t)V * HPR_G2;
  return hpr_al(vn * 32) * 2 + hpr_al(vg * 32) + hpr_al(vn * 16) * 2 + hpr_al((size_t)V * 32) +
         hpr_al(vg * 8) + hpr_al(vg * 4) + 2 * hpr_al((size_t)V * 4) + 2 * hpr_al(vn) + 256;
}

static HprWs hpr_carve(void* ws, int V, int N) {
  const size_t vn = (size_t)V * N, vg = (size_t)V * HPR_G2;
  uint8_t* w = (uint8_t*)ws;
  HprWs r;
  r.Q = (double4*)w, w += hpr_al(vn * 32);
  r.C = (double4*)w, w += hpr_al(vn * 32);
  r.E = (double4*)w, w += hpr_al(vg * 32);
  r.ab = (double2*)w, w += hpr_al(vn * 16);
  r.Cab = (double2*)w, w += hpr_al(vn * 16);
  r.bbox = (long long*)w, w += hpr_al((size_t)V * 32);
  r.cellmax = (long long*)w, w += hpr_al(vg * 8);
  r.cellidx = (int*)w, w += hpr_al(vg * 4);
  r.nE = (int*)w, w += hpr_al((size_t)V * 4);
  r.nC = (int*)w, w += hpr_al((size_t)V * 4);
  r.surv = w, w += hpr_al(vn);
  r.isE = w;
  return r;
}

__device__ __forceinline__ long long hpr_ordered(double d) {
  const long long b = __double_as_longlong(d);
  return b >= 0 ? b : b ^ 0x7FFFFFFFFFFFFFFFll;
}
__device__ __forceinline__ double hpr_unordered(long long k) {
  return __longlong_as_double(k >= 0 ? k : k ^ 0x7FFFFFFFFFFFFFFFll);
}

__global__ void hpr_init_kernel(HprWs ws, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V * HPR_G2) {
    ws.cellmax[i] = LLONG_MIN;
    ws.cellidx[i] = INT_MAX;
  }
  if (i < V * 4) ws.bbox[i] = (i & 2) ? LLONG_MIN : LLONG_MAX;
}

// frames: [V][12] doubles = eye(3), ex(3), ey(3), ez(3).  Q[v][n] = (u, v, w, n)
__global__ void hpr_prepare_kernel(const float* __restrict__ points, int N, int V,
                                   const double* __restrict__ frames, double radius, HprWs ws) {
  // grid = (blocks per view, V): a block never straddles two views
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  const size_t i = (size_t)v * N + n;
  long long k0 = LLONG_MAX, k1 = LLONG_MAX, k2 = LLONG_MIN, k3 = LLONG_MIN;
  if (n < N) {
    const double* f = frames + v * 12;
    const double px = (double)points[3 * n] - f[0], py = (double)points[3 * n + 1] - f[1],
                 pz = (double)points[3 * n + 2] - f[2];
    const double nrm = sqrt(px * px + py * py + pz * pz);
    const double s = 2.0 * radius / nrm - 1.0;
    const double x = px * f[3] + py * f[4] + pz * f[5];
    const double y = px * f[6] + py * f[7] + pz * f[8];
    const double z = px * f[9] + py * f[10] + pz * f[11];
    const double4 q = make_double4(x / z, y / z, -HPR_WSCALE / (s * z), (double)n);
    ws.Q[i] = q;
    k0 = k2 = hpr_ordered(q.x);
    k1 = k3 = hpr_ordered(q.y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    k0 = min(k0, __shfl_xor_sync(0xffffffffu, k0, o));
    k1 = min(k1, __shfl_xor_sync(0xffffffffu, k1, o));
    k2 = max(k2, __shfl_xor_sync(0xffffffffu, k2, o));
    k3 = max(k3, __shfl_xor_sync(0xffffffffu, k3, o));
  }
  if ((threadIdx.x & 31) == 0 && k0 != LLONG_MAX) {
    atomicMin(&ws.bbox[v * 4 + 0], k0);
    atomicMin(&ws.bbox[v * 4 + 1], k1);
    atomicMax(&ws.bbox[v * 4 + 2], k2);
    atomicMax(&ws.bbox[v * 4 + 3], k3);
  }
}

__device__ __forceinline__ int hpr_cell(const double4 q, const long long* bbox) {
  const double u0 = hpr_unordered(bbox[0]), v0 = hpr_unordered(bbox[1]);
  const double du = hpr_unordered(bbox[2]) - u0, dv = hpr_unordered(bbox[3]) - v0;
  int cu = du > 0.0 ? (int)((q.x - u0) / du * HPR_G) : 0;
  int cv = dv > 0.0 ? (int)((q.y - v0) / dv * HPR_G) : 0;
  cu = min(max(cu, 0), HPR_G - 1);
  cv = min(max(cv, 0), HPR_G - 1);
  return cv * HPR_G + cu;
}

// pass 0: highest w per cell; pass 1: lowest index among the points that reach it
__global__ void hpr_cell_kernel(int N, int V, HprWs ws, int pass) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N;
  const double4 q = ws.Q[i];
  const int c = v * HPR_G2 + hpr_cell(q, ws.bbox + v * 4);
  const long long key = hpr_ordered(q.z);
  if (pass == 0)
    atomicMax(&ws.cellmax[c], key);
  else if (key == ws.cellmax[c])
    atomicMin(&ws.cellidx[c], (int)(i % N));
}

// block-wide exclusive scan of one flag per thread (1024 threads); returns the block total
__device__ __forceinline__ int hpr_block_scan(int flag, int& total, int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int within = __popc(bal & ((1u << lane) - 1));
  __syncthreads();
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int x = s_warp[lane];
    int incl = x;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    s_warp[lane] = incl - x;
    if (lane == 31) s_warp[32] = incl;
  }
  __syncthreads();
  total = s_warp[32];
  return s_warp[warp] + within;
}

// E[v] = the cell maxima in the (pseudo-random) visiting order of the cells; one block per view
__global__ void __launch_bounds__(1024) hpr_compact_e_kernel(int N, HprWs ws) {
  __shared__ int s_warp[33];
  const int v = blockIdx.x;
  int base = 0;
  for (int p0 = 0; p0 < HPR_G2; p0 += 1024) {
    const int p = p0 + threadIdx.x;
    int idx = INT_MAX;
    if (p < HPR_G2) idx = ws.cellidx[v * HPR_G2 + (int)(((long long)p * HPR_CELL_STRIDE + 7) % HPR_G2)];
    const int flag = idx != INT_MAX;
    int total;
    const int pos = hpr_block_scan(flag, total, s_warp);
    if (flag) {
      ws.E[(size_t)v * HPR_G2 + base + pos] = ws.Q[(size_t)v * N + idx];
      ws.isE[(size_t)v * N + idx] = 1;
    }
    base += total;
  }
  if (threadIdx.x == 0) ws.nE[v] = base;
}

// C[v] = E[v] followed by the other survivors of the filter in the visiting order
// k -> (k*stride + offset) mod N; points the filter rejected are hidden: vis = 0
__global__ void __launch_bounds__(1024)
hpr_compact_c_kernel(int N, int stride, int offset, HprWs ws, uint8_t* __restrict__ vis) {
  __shared__ int s_warp[33];
  const int v = blockIdx.x;
  const int nE = ws.nE[v];
  for (int k = threadIdx.x; k < nE; k += 1024) {
    const double4 e = ws.E[(size_t)v * HPR_G2 + k];
    ws.C[(size_t)v * N + k] = e;
    ws.Cab[(size_t)v * N + k] = ws.ab[(size_t)v * N + (int)e.w];
  }
  int base = nE;
  for (int p0 = 0; p0 < N; p0 += 1024) {
    const int p = p0 + threadIdx.x;
    int n = -1, flag = 0;
    if (p < N) {
      n = (int)(((long long)p * stride + offset) % N);
      const size_t i = (size_t)v * N + n;
      const bool s = ws.surv[i] != 0;
      if (!s) vis[i] = 0;
      flag = s && !ws.isE[i];
    }
    int total;
    const int pos = hpr_block_scan(flag, total, s_warp);
    if (flag) {
      ws.C[(size_t)v * N + base + pos] = ws.Q[(size_t)v * N + n];
      ws.Cab[(size_t)v * N + base + pos] = ws.ab[(size_t)v * N + n];
    }
    base += total;
  }
  if (threadIdx.x == 0) ws.nC[v] = base;
}

// tighten [lo, hi] on the line p0 + t d with one earlier constraint (skipped when `use` is false).
// Branch-light: the only divergent path is the division, taken when a bound actually moves (rare
// once the first few constraints of a re-solve are in); comparisons by cross-multiplication.
__device__ __forceinline__ void hpr_clip(const double4 c, bool use, double uL, double vL, double wL,
                                         double p0x, double p0y, double dx, double dy, double& lo,
                                         double& hi) {
  const double ax = c.x - uL, ay = c.y - vL, ah = c.z - wL;
  const double den = ax * dx + ay * dy;
  const double rhs = ah - (ax * p0x + ay * p0y);
  const bool pos = den > 0.0, neg = den < 0.0;
  // den > 0: t >= rhs/den, moves lo when rhs > lo*den;  den < 0: t <= rhs/den, moves hi when
  // rhs/den < hi  <=>  rhs > hi*den
  const double bound = pos ? lo : hi;
  const bool moves = use && (pos || neg) && rhs > bound * den;
  if (moves) {
    const double t = rhs / den;
    if (pos) lo = t; else hi = t;
  }
  if (use && !pos && !neg && rhs > 0.0) lo = INFINITY;
}

// Lane `L`'s optimum was cut off by constraint cj at position `pos` of the sequence q: the warp
// solves the 1-D LP on cj's line over the constraints before `pos` (lane L's own point excluded by
// index) and lane L takes the new optimum or becomes infeasible.  UNROLL = loads in flight per lane
// (q in shared memory: 4; q in global memory / L2: 8).
template <int UNROLL>
__device__ __forceinline__ void hpr_resolve(const double4* q, int pos, const double4 cj, int L,
                                            int lane, const double4 me, double& a, double& b,
                                            bool& feasible) {
  const double c0 = 1.0, c1 = 0.5;  // objective: maximise c.x inside the box
  const double uL = __shfl_sync(0xffffffffu, me.x, L), vL = __shfl_sync(0xffffffffu, me.y, L),
               wL = __shfl_sync(0xffffffffu, me.z, L), iL = __shfl_sync(0xffffffffu, me.w, L);
  const double nx = cj.x - uL, ny = cj.y - vL, h = cj.z - wL;
  const double nn = nx * nx + ny * ny;
  double lo = -INFINITY, hi = INFINITY;
  double p0x = 0.0, p0y = 0.0, dx = 0.0, dy = 0.0;
  bool ok = nn > 0.0;  // a point exactly above in the same direction: infeasible
  if (ok) {
    const double sc = h / nn;
    p0x = nx * sc, p0y = ny * sc;
    dx = -ny, dy = nx;
    // all earlier constraints, split over the lanes
    int k = lane;
    for (; k + (UNROLL - 1) * 32 < pos; k += UNROLL * 32) {
      double4 cu[UNROLL];
#pragma unroll
      for (int r = 0; r < UNROLL; ++r) cu[r] = q[k + 32 * r];
#pragma unroll
      for (int r = 0; r < UNROLL; ++r)
        hpr_clip(cu[r], cu[r].w != iL, uL, vL, wL, p0x, p0y, dx, dy, lo, hi);
    }
    for (; k < pos; k += 32) {
      const double4 c = q[k];
      hpr_clip(c, c.w != iL, uL, vL, wL, p0x, p0y, dx, dy, lo, hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmin(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // the box |p0 + t d| <= BOX only matters on the side the objective pushes to while that side
    // is still unbounded (the first few constraints of a sequence): its four divisions are skipped
    // otherwise.  (A finite bound beyond the box is clamped by it all the same.)
    const bool up = c0 * dx + c1 * dy > 0.0;
    const double tsel = up ? hi : lo;
    if (!(fabs(p0x + tsel * dx) <= HPR_BOX && fabs(p0y + tsel * dy) <= HPR_BOX)) {
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
    }
    ok = ok && lo <= hi;
    if (lane == L) {
      if (!ok) {
        feasible = false;
      } else {
        const double tt = up ? hi : lo;
        a = p0x + tt * dx;
        b = p0y + tt * dy;
      }
    }
  } else if (lane == L) {
    feasible = false;
  }
}

// the streaming part: constraints sq[0..cnt) are positions base.. of the sequence q
template <int UNROLL>
__device__ __forceinline__ void hpr_scan_tile(const double4* q, const double4* sq, int base, int cnt,
                                              int lane, const double4 me, double& a, double& b,
                                              bool& feasible) {
  for (int t0 = 0; t0 < cnt; t0 += 4) {
    // fast path: none of the next four constraints cuts off any lane's optimum
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (t0 + j < cnt) {
        const double4 cj = sq[t0 + j];
        any |= feasible && cj.w != me.w && ((cj.x - me.x) * a + (cj.y - me.y) * b < cj.z - me.z);
      }
    }
    if (!__any_sync(0xffffffffu, any)) continue;
    for (int t = t0; t < min(t0 + 4, cnt); ++t) {
      const double4 cj = sq[t];
      const bool viol =
          feasible && cj.w != me.w && ((cj.x - me.x) * a + (cj.y - me.y) * b < cj.z - me.z);
      unsigned m = __ballot_sync(0xffffffffu, viol);
      while (m) {
        const int L = __ffs(m) - 1;
        m &= m - 1;
        hpr_resolve<UNROLL>(q, base + t, cj, L, lane, me, a, b, feasible);
      }
    }
  }
}

// FILTER: every point of the cloud against the extremes E (resident in shared memory).
__global__ void __launch_bounds__(32 * HPR_FILTER_WARPS, 3)  // <= 85 registers: 3 blocks (24 warps) per SM
hpr_filter_kernel(int N, HprWs ws) {
  extern __shared__ double4 se[];
  const int v = blockIdx.y;
  const int nE = ws.nE[v];
  for (int k = threadIdx.x; k < nE; k += blockDim.x) se[k] = ws.E[(size_t)v * HPR_G2 + k];
  __syncthreads();  // the only block-wide synchronisation: warps run independently afterwards
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pi < N;
  const double4 me = active ? ws.Q[(size_t)v * N + pi] : make_double4(0.0, 0.0, 0.0, -1.0);
  double a = HPR_BOX, b = HPR_BOX;  // the (+,+) corner maximises (1, 0.5).x
  bool feasible = active;
  for (int base = 0; base < nE; base += HPR_TILE) {
    if (!__any_sync(0xffffffffu, feasible)) break;
    hpr_scan_tile<4>(se, se + base, base, min(HPR_TILE, nE - base), lane, me, a, b, feasible);
  }
  if (active) {
    ws.surv[(size_t)v * N + pi] = feasible ? 1 : 0;
    ws.ab[(size_t)v * N + pi] = make_double2(a, b);
  }
}

// EXACT: the survivors against each other, continuing after the E prefix.  One warp per block
// (warps re-solve at very different times, so nothing may couple them) and only HPR_EXACT_POINTS
// points per warp: the warp-wide re-solves of its points are serialised, and they - not the
// streaming scan - are what the pass spends its time on, so fewer points per warp (more warps in
// flight) shortens every warp's critical path.
static constexpr int HPR_EXACT_POINTS = 16;
__global__ void __launch_bounds__(32)
hpr_exact_kernel(int N, HprWs ws, uint8_t* __restrict__ vis) {
  __shared__ double4 sq[HPR_TILE];
  const int v = blockIdx.y;
  const int nC = ws.nC[v], nE = ws.nE[v];
  const int lane = threadIdx.x;
  if (blockIdx.x * HPR_EXACT_POINTS >= nC) return;
  const int pi = blockIdx.x * HPR_EXACT_POINTS + lane;
  const double4* q = ws.C + (size_t)v * N;
  const bool active = lane < HPR_EXACT_POINTS && pi < nC;
  const double4 me = active ? q[pi] : make_double4(0.0, 0.0, 0.0, -1.0);
  const double2 ab0 = active ? ws.Cab[(size_t)v * N + pi] : make_double2(0.0, 0.0);
  double a = ab0.x, b = ab0.y;
  // an extreme that the filter itself rejected stays in the E prefix as a (valid) constraint only
  bool feasible = active && ws.surv[(size_t)v * N + (int)me.w] != 0;
  for (int base = nE; base < nC; base += HPR_TILE) {
    if (!__any_sync(0xffffffffu, feasible)) break;
    __syncwarp();
    const int cnt = min(HPR_TILE, nC - base);
    for (int t = lane; t < cnt; t += 32) sq[t] = q[base + t];
    __syncwarp();
    hpr_scan_tile<8>(q, sq, base, cnt, lane, me, a, b, feasible);
  }
  if (active) vis[(size_t)v * N + (int)me.w] = feasible ? 1 : 0;
}

int hpr_launch(const float* points, int N, int V, const double* frames_dev, double radius,
               void* workspace, uint8_t* vis, cudaStream_t stream) {
  PDR_CHECK_ARG(N > 0 && V > 0, "hidden point removal: empty input");
  PDR_CHECK_ARG(((uintptr_t)workspace & 31) == 0, "hidden point removal: workspace must be 32-byte aligned");
  HprWs ws = hpr_carve(workspace, V, N);
  // visiting order of the survivors: k -> (k*stride + offset) mod N, stride coprime with N
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
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(hpr_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HPR_G2 * (int)sizeof(double4)));
    configured = true;
  }
  PDR_CUDA(cudaMemsetAsync(ws.isE, 0, (size_t)V * N, stream));
  hpr_init_kernel<<<cdiv(V * HPR_G2, 256), 256, 0, stream>>>(ws, V);
  PDR_COUNT_LAUNCH();
  hpr_prepare_kernel<<<dim3(cdiv(N, 256), V), 256, 0, stream>>>(points, N, V, frames_dev, radius, ws);
  PDR_COUNT_LAUNCH();
  hpr_cell_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(N, V, ws, 0);
  PDR_COUNT_LAUNCH();
  hpr_cell_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(N, V, ws, 1);
  PDR_COUNT_LAUNCH();
  hpr_compact_e_kernel<<<V, 1024, 0, stream>>>(N, ws);
  PDR_COUNT_LAUNCH();
  hpr_filter_kernel<<<dim3(cdiv(N, 32 * HPR_FILTER_WARPS), V), 32 * HPR_FILTER_WARPS,
                      HPR_G2 * sizeof(double4), stream>>>(N, ws);
  PDR_COUNT_LAUNCH();
  hpr_compact_c_kernel<<<V, 1024, 0, stream>>>(N, stride, offset, ws, vis);
  PDR_COUNT_LAUNCH();
  hpr_exact_kernel<<<dim3(cdiv(N, HPR_EXACT_POINTS), V), 32, 0, stream>>>(N, ws, vis);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
