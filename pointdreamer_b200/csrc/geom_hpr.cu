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
//   2. EXACT: the survivors C = E + S run the LP against C only, CONTINUING from the optimum the
//      filter left (Seidel's invariant holds: E is the prefix of every point's constraint sequence).
//      The other survivors S are sorted by the Morton code of their cell on a 64 x 64 grid (ties by
//      index: a deterministic order) and cut into tiles of 64 with a bounding box each; a warp owns
//      16 consecutive points and visits the tiles outwards from its own ("neighbour first": the
//      constraints that bind are met early).  The boxes prune both halves of the work, exactly:
//        * scan: a tile none of whose members can violate any lane's optimum is skipped;
//        * re-solve: the 1-D LP on the new constraint's line is a max/min over the earlier
//          constraints, so their order is free and a tile none of whose members can move lo or hi
//          is skipped; E is kept a second time in spatial blocks of 8 x 4 cells for this.
//      A box test is the constraint's own inequality evaluated at the box corner that maximises it
//      plus a rounding margin, so a skipped tile provably holds no constraint the unpruned pass
//      would have acted on: decisions equal the unpruned pass over the same sequence.  A re-solve
//      whose interval is already empty stops early.
// Work drops from V N^2 to V (N |E| + |C| (|C| / prune)) constraint checks (measured prune: 3x of
// the scan, 6x of the re-solve clips on the demo clouds), all in fp64 like the reference.
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
static constexpr int HPR_M = 64;                       // Morton grid that orders the survivors
static constexpr int HPR_M2 = HPR_M * HPR_M;
static constexpr int HPR_ST = 64;                      // survivors per tile (one bounding box each)
static constexpr int HPR_EBW = 8, HPR_EBH = 4;         // E blocks: 8 x 4 cells = one warp of extremes
static constexpr int HPR_NET = (HPR_G / HPR_EBW) * (HPR_G / HPR_EBH);  // 50 blocks
static constexpr double HPR_EPS = 64.0 * 2.220446049250313e-16;  // rounding margin of a box test
static_assert(HPR_G % HPR_EBW == 0 && HPR_G % HPR_EBH == 0 && HPR_EBW * HPR_EBH == 32, "E blocks");

struct HprWs {
  double4* Q;        // [V][N]  (u, v, w, index)
  double4* E;        // [V][G2] extremes, visiting order (the filter's constraint sequence)
  double4* E2;       // [V][NET][32] the same extremes in spatial blocks of 8 x 4 cells
  double4* S;        // [V][N]  survivors that are not extremes, Morton order
  double2* ab;       // [V][N]  filter optimum per point
  double2* Sab;      // [V][N]  the same, in S order
  double* sbox;      // [V][NT][6] bounding box of every S tile: u0, u1, v0, v1, w0, w1
  double* ebox;      // [V][NET][6] bounding box of every E block
  long long* bbox;   // [V][4]  ordered keys: min u, min v, max u, max v
  long long* cellmax;  // [V][G2]
  int* cellidx;      // [V][G2]
  int* tmpidx;       // [V][N]  survivors bucketed by Morton cell, unordered inside a cell
  int* celloffs;     // [V][M2+1] first S position of every Morton cell
  int* cellcnt;      // [V][M2]  (zeroed per call, contiguous with cellcur)
  int* cellcur;      // [V][M2]
  int* ecnt;         // [V][NET]
  int* nE;           // [V]
  int* nS;           // [V]
  uint16_t* mcell;   // [V][N]  Morton cell of every point
  uint8_t* surv;     // [V][N]
  uint8_t* isE;      // [V][N]
};

static inline size_t hpr_al(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int hpr_tiles(int N) { return (N + HPR_ST - 1) / HPR_ST; }
__device__ __forceinline__ int hpr_tiles_dev(int N) { return (N + HPR_ST - 1) / HPR_ST; }

// one table drives both the size query and the carving, so they cannot drift apart
template <typename F>
static void hpr_layout(int V, int N, F&& field) {
  const size_t vn = (size_t)V * N, vg = (size_t)V * HPR_G2, vm = (size_t)V * HPR_M2,
               ve = (size_t)V * HPR_NET;
  field(0, vn * 32);                                // Q
  field(1, vg * 32);                                // E
  field(2, ve * 32 * 32);                           // E2
  field(3, vn * 32);                                // S
  field(4, vn * 16);                                // ab
  field(5, vn * 16);                                // Sab
  field(6, (size_t)V * hpr_tiles(N) * 6 * 8);       // sbox
  field(7, ve * 6 * 8);                             // ebox
  field(8, (size_t)V * 32);                         // bbox
  field(9, vg * 8);                                 // cellmax
  field(10, vg * 4);                                // cellidx
  field(11, vn * 4);                                // tmpidx
  field(12, (size_t)V * (HPR_M2 + 1) * 4);          // celloffs
  field(13, 2 * vm * 4);                            // cellcnt + cellcur
  field(14, ve * 4);                                // ecnt
  field(15, (size_t)V * 4);                         // nE
  field(16, (size_t)V * 4);                         // nS
  field(17, vn * 2);                                // mcell
  field(18, vn);                                    // surv
  field(19, vn);                                    // isE
}

size_t hpr_workspace_bytes(int V, int N) {
  size_t total = 256;
  hpr_layout(V, N, [&](int, size_t bytes) { total += hpr_al(bytes); });
  return total;
}

static HprWs hpr_carve(void* ws, int V, int N) {
  uint8_t* w = (uint8_t*)ws;
  void* f[20];
  hpr_layout(V, N, [&](int i, size_t bytes) {
    f[i] = w;
    w += hpr_al(bytes);
  });
  HprWs r;
  r.Q = (double4*)f[0], r.E = (double4*)f[1], r.E2 = (double4*)f[2], r.S = (double4*)f[3];
  r.ab = (double2*)f[4], r.Sab = (double2*)f[5], r.sbox = (double*)f[6], r.ebox = (double*)f[7];
  r.bbox = (long long*)f[8], r.cellmax = (long long*)f[9], r.cellidx = (int*)f[10];
  r.tmpidx = (int*)f[11], r.celloffs = (int*)f[12], r.cellcnt = (int*)f[13];
  r.cellcur = r.cellcnt + (size_t)V * HPR_M2;
  r.ecnt = (int*)f[14], r.nE = (int*)f[15], r.nS = (int*)f[16], r.mcell = (uint16_t*)f[17];
  r.surv = (uint8_t*)f[18], r.isE = (uint8_t*)f[19];
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

// ---- ordering of the survivors ------------------------------------------------------------------
__device__ __forceinline__ unsigned hpr_spread(unsigned x) {  // 0b abcdef -> 0b 0a0b0c0d0e0f
  x = (x | (x << 4)) & 0x0F0Fu;
  x = (x | (x << 2)) & 0x3333u;
  x = (x | (x << 1)) & 0x5555u;
  return x;
}
__device__ __forceinline__ int hpr_mcell(const double4 q, const long long* bbox) {
  const double u0 = hpr_unordered(bbox[0]), v0 = hpr_unordered(bbox[1]);
  const double du = hpr_unordered(bbox[2]) - u0, dv = hpr_unordered(bbox[3]) - v0;
  int cu = du > 0.0 ? (int)((q.x - u0) / du * HPR_M) : 0;
  int cv = dv > 0.0 ? (int)((q.y - v0) / dv * HPR_M) : 0;
  cu = min(max(cu, 0), HPR_M - 1);
  cv = min(max(cv, 0), HPR_M - 1);
  return (int)(hpr_spread((unsigned)cu) | (hpr_spread((unsigned)cv) << 1));
}

__device__ __forceinline__ void hpr_box_reduce(double& x0, double& x1, double& y0, double& y1,
                                               double& z0, double& z1) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o));
    x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    z0 = fmin(z0, __shfl_xor_sync(0xffffffffu, z0, o));
    z1 = fmax(z1, __shfl_xor_sync(0xffffffffu, z1, o));
  }
}

// E2: the extremes once more, grouped by blocks of 8 x 4 cells (one warp per block) with a box each
__global__ void __launch_bounds__(32) hpr_eblock_kernel(int N, HprWs ws) {
  const int t = blockIdx.x, v = blockIdx.y, lane = threadIdx.x;
  const int cu = (t % (HPR_G / HPR_EBW)) * HPR_EBW + (lane & (HPR_EBW - 1));
  const int cv = (t / (HPR_G / HPR_EBW)) * HPR_EBH + lane / HPR_EBW;
  const int idx = ws.cellidx[v * HPR_G2 + cv * HPR_G + cu];
  const bool valid = idx != INT_MAX;
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY, z0 = INFINITY, z1 = -INFINITY;
  if (valid) {
    const double4 q = ws.Q[(size_t)v * N + idx];
    ws.E2[((size_t)v * HPR_NET + t) * 32 + __popc(bal & ((1u << lane) - 1))] = q;
    x0 = x1 = q.x, y0 = y1 = q.y, z0 = z1 = q.z;
  }
  hpr_box_reduce(x0, x1, y0, y1, z0, z1);
  if (lane == 0) {
    double* bx = ws.ebox + ((size_t)v * HPR_NET + t) * 6;
    bx[0] = x0, bx[1] = x1, bx[2] = y0, bx[3] = y1, bx[4] = z0, bx[5] = z1;
    ws.ecnt[v * HPR_NET + t] = __popc(bal);
  }
}

// points the filter rejected are hidden (vis = 0); the other non-extremes are counted per Morton cell
__global__ void hpr_count_kernel(int N, int V, HprWs ws, uint8_t* __restrict__ vis) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N;
  const int mc = hpr_mcell(ws.Q[i], ws.bbox + v * 4);
  ws.mcell[i] = (uint16_t)mc;
  if (!ws.surv[i]) {
    vis[i] = 0;
    return;
  }
  if (!ws.isE[i]) atomicAdd(&ws.cellcnt[v * HPR_M2 + mc], 1);
}

// exclusive scan of the 4096 cell counts of a view (one block per view, 4 cells per thread)
__global__ void __launch_bounds__(1024) hpr_offsets_kernel(HprWs ws) {
  __shared__ int s_warp[33];
  static_assert(HPR_M2 == 4 * 1024, "4 cells per thread");
  const int v = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int c[4], sum = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) sum += c[j] = ws.cellcnt[v * HPR_M2 + 4 * t + j];
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int x = s_warp[lane];
    int w = x;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    s_warp[lane] = w - x;
  }
  __syncthreads();
  int base = s_warp[warp] + incl - sum;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ws.celloffs[v * (HPR_M2 + 1) + 4 * t + j] = base;
    base += c[j];
  }
  if (t == 1023) {
    ws.celloffs[v * (HPR_M2 + 1) + HPR_M2] = base;
    ws.nS[v] = base;
  }
}

// bucket the survivors by cell (the order inside a cell is whatever the atomics give) ...
__global__ void hpr_scatter_kernel(int N, int V, HprWs ws) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N;
  if (!ws.surv[i] || ws.isE[i]) return;
  const int mc = ws.mcell[i];
  const int pos = ws.celloffs[v * (HPR_M2 + 1) + mc] + atomicAdd(&ws.cellcur[v * HPR_M2 + mc], 1);
  ws.tmpidx[(size_t)v * N + pos] = (int)(i % N);
}

// ... then every survivor takes the rank of its index inside its cell: S is ordered by (cell, index)
__global__ void hpr_place_kernel(int N, int V, HprWs ws) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * N) return;
  const int v = i / N, n = (int)(i % N);
  if (!ws.surv[i] || ws.isE[i]) return;
  const int mc = ws.mcell[i];
  const int o0 = ws.celloffs[v * (HPR_M2 + 1) + mc], o1 = ws.celloffs[v * (HPR_M2 + 1) + mc + 1];
  int rank = 0;
  for (int k = o0; k < o1; ++k) rank += ws.tmpidx[(size_t)v * N + k] < n;
  ws.S[(size_t)v * N + o0 + rank] = ws.Q[i];
  ws.Sab[(size_t)v * N + o0 + rank] = ws.ab[i];
}

// bounding box of every tile of 64 survivors (one warp per tile)
__global__ void __launch_bounds__(256) hpr_tilebox_kernel(int N, HprWs ws) {
  const int v = blockIdx.y, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int nS = ws.nS[v];
  if (tile * HPR_ST >= nS) return;
  double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY, z0 = INFINITY, z1 = -INFINITY;
  for (int k = tile * HPR_ST + lane; k < min(nS, (tile + 1) * HPR_ST); k += 32) {
    const double4 q = ws.S[(size_t)v * N + k];
    x0 = fmin(x0, q.x), x1 = fmax(x1, q.x), y0 = fmin(y0, q.y), y1 = fmax(y1, q.y);
    z0 = fmin(z0, q.z), z1 = fmax(z1, q.z);
  }
  hpr_box_reduce(x0, x1, y0, y1, z0, z1);
  if (lane == 0) {
    double* bx = ws.sbox + ((size_t)v * hpr_tiles_dev(N) + tile) * 6;
    bx[0] = x0, bx[1] = x1, bx[2] = y0, bx[3] = y1, bx[4] = z0, bx[5] = z1;
  }
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

// ---- EXACT pass ---------------------------------------------------------------------------------
// can a member of the box violate  (x - u) a + (y - v) b >= z - w ?  The inequality is evaluated at
// the corner of the box that maximises (z - w) - (x - u) a - (y - v) b; `ma`, `mb` bound |a|, |b|
// INCLUDING the terms they were summed from, so that HPR_EPS * mag covers every rounding of the
// member-by-member evaluation this test stands in for.  NaN / infinite inputs answer "yes".
__device__ __forceinline__ bool hpr_box_reach(const double* bx, double u, double v, double w, double a,
                                              double b, double ma, double mb) {
  const double x0 = bx[0] - u, x1 = bx[1] - u, y0 = bx[2] - v, y1 = bx[3] - v, z0 = bx[4] - w,
               z1 = bx[5] - w;
  const double g = z1 - (a > 0.0 ? x0 : x1) * a - (b > 0.0 ? y0 : y1) * b;
  const double mag = fmax(fabs(z0), fabs(z1)) + fmax(fabs(x0), fabs(x1)) * ma +
                     fmax(fabs(y0), fabs(y1)) * mb;
  return !(g + HPR_EPS * mag <= 0.0);
}

// tile visited at step s: the warp's own tile, then alternately right / left of it, then whatever
// side is left
__device__ __forceinline__ int hpr_tile_at(int s, int home, int nT) {
  if (s == 0) return home;
  const int r = nT - 1 - home, l = home, m = min(r, l);
  if (s <= 2 * m) return (s & 1) ? home + (s + 1) / 2 : home - s / 2;
  return r > l ? home + (s - m) : home - (s - m);
}

struct HprSeq {       // the constraint sequence of one warp
  const double4* S;   // survivors of the view, Morton order
  const double* sbox;
  const double4* E2;  // extremes of the view in spatial blocks
  const double* ebox;
  const int* ecnt;
  int nS, nT, home;
};

struct HprLine {  // the 1-D LP of a re-solve: optimum = p0 + t d, t in [lo, hi]
  double uL, vL, wL, iL, p0x, p0y, dx, dy, lo, hi;
};

__device__ __forceinline__ void hpr_line_clip(HprLine& ln, const double4 c) {
  hpr_clip(c, c.w != ln.iL, ln.uL, ln.vL, ln.wL, ln.p0x, ln.p0y, ln.dx, ln.dy, ln.lo, ln.hi);
}
// every lane leaves with the warp-wide interval; true when it is empty (the re-solve can stop:
// further constraints only shrink it)
__device__ __forceinline__ bool hpr_line_reduce(HprLine& ln) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ln.lo = fmax(ln.lo, __shfl_xor_sync(0xffffffffu, ln.lo, o));
    ln.hi = fmin(ln.hi, __shfl_xor_sync(0xffffffffu, ln.hi, o));
  }
  return ln.lo > ln.hi;
}
// can a member of the box move lo or hi?  It moves lo (hi) exactly when it is violated by the plane
// at t = lo (t = hi), so two box tests; an interval that is still open on a side is not prunable
__device__ __forceinline__ bool hpr_line_box(const HprLine& ln, const double* bx) {
  if (!(fabs(ln.lo) < INFINITY && fabs(ln.hi) < INFINITY)) return true;
  const double tl = ln.lo * ln.dx, sl = ln.lo * ln.dy, th = ln.hi * ln.dx, sh = ln.hi * ln.dy;
  return hpr_box_reach(bx, ln.uL, ln.vL, ln.wL, ln.p0x + tl, ln.p0y + sl, fabs(ln.p0x) + fabs(tl),
                       fabs(ln.p0y) + fabs(sl)) ||
         hpr_box_reach(bx, ln.uL, ln.vL, ln.wL, ln.p0x + th, ln.p0y + sh, fabs(ln.p0x) + fabs(th),
                       fabs(ln.p0y) + fabs(sh));
}

// Lane L's optimum was cut off by constraint cj = sq[t] of the tile visited at step `step`: the warp
// solves the 1-D LP on cj's line over everything before it - the tile's prefix, the tiles of the
// earlier steps and E - pruned by the boxes (the LP is a max / min, the order is free).
__device__ __forceinline__ void hpr_resolve_tiles(const HprSeq& sq_, const double4* sq, int t, int step,
                                               const double4 cj, int L, int lane, const double4 me,
                                               double& a, double& b, bool& feasible) {
  const double c0 = 1.0, c1 = 0.5;  // the filter's objective
  HprLine ln;
  ln.uL = __shfl_sync(0xffffffffu, me.x, L), ln.vL = __shfl_sync(0xffffffffu, me.y, L);
  ln.wL = __shfl_sync(0xffffffffu, me.z, L), ln.iL = __shfl_sync(0xffffffffu, me.w, L);
  const double nx = cj.x - ln.uL, ny = cj.y - ln.vL, h = cj.z - ln.wL;
  const double nn = nx * nx + ny * ny;
  bool ok = nn > 0.0;  // a point exactly above in the same direction: infeasible
  if (ok) {
    const double sc = h / nn;
    ln.p0x = nx * sc, ln.p0y = ny * sc, ln.dx = -ny, ln.dy = nx;
    ln.lo = -INFINITY, ln.hi = INFINITY;
    bool empty = false;
    // the prefix of the current tile (shared memory) and the warp's own tile: the nearest constraints
    for (int k = lane; k < t; k += 32) hpr_line_clip(ln, sq[k]);
    if (step > 0) {
      const int k0 = sq_.home * HPR_ST, k1 = min(sq_.nS, k0 + HPR_ST);
      const double4 c0_ = k0 + lane < k1 ? sq_.S[k0 + lane] : make_double4(0, 0, 0, ln.iL);
      const double4 c1_ = k0 + lane + 32 < k1 ? sq_.S[k0 + lane + 32] : make_double4(0, 0, 0, ln.iL);
      hpr_line_clip(ln, c0_);
      hpr_line_clip(ln, c1_);
    }
    empty = hpr_line_reduce(ln);
    // the tiles of steps 1 .. step-1, 32 box tests at a time
    for (int s0 = 1; s0 < step && !empty; s0 += 32) {
      const int s = s0 + lane;
      int tl = -1;
      bool need = false;
      if (s < step) {
        tl = hpr_tile_at(s, sq_.home, sq_.nT);
        need = hpr_line_box(ln, sq_.sbox + (size_t)tl * 6);
      }
      unsigned m = __ballot_sync(0xffffffffu, need);
      if (!m) continue;
      while (m) {  // two tiles per round: four loads in flight per lane
        const int j0 = __ffs(m) - 1;
        m &= m - 1;
        const int j1 = m ? __ffs(m) - 1 : j0;
        m &= m - 1;  // (0 & anything = 0)
        const int k0 = __shfl_sync(0xffffffffu, tl, j0) * HPR_ST, k1 = min(sq_.nS, k0 + HPR_ST);
        const int k2 = __shfl_sync(0xffffffffu, tl, j1) * HPR_ST, k3 = j1 != j0 ? min(sq_.nS, k2 + HPR_ST) : 0;
        const double4 dummy = make_double4(0.0, 0.0, 0.0, ln.iL);
        const double4 c0_ = k0 + lane < k1 ? sq_.S[k0 + lane] : dummy;
        const double4 c1_ = k0 + lane + 32 < k1 ? sq_.S[k0 + lane + 32] : dummy;
        const double4 c2_ = k2 + lane < k3 ? sq_.S[k2 + lane] : dummy;
        const double4 c3_ = k2 + lane + 32 < k3 ? sq_.S[k2 + lane + 32] : dummy;
        hpr_line_clip(ln, c0_);
        hpr_line_clip(ln, c1_);
        hpr_line_clip(ln, c2_);
        hpr_line_clip(ln, c3_);
      }
      empty = hpr_line_reduce(ln);
    }
    // the extremes, block by block
    for (int e0 = 0; e0 < HPR_NET && !empty; e0 += 32) {
      const int e = e0 + lane;
      const bool need = e < HPR_NET && sq_.ecnt[e] > 0 && hpr_line_box(ln, sq_.ebox + (size_t)e * 6);
      unsigned m = __ballot_sync(0xffffffffu, need);
      if (!m) continue;
      while (m) {  // two blocks per round
        const int ej0 = e0 + __ffs(m) - 1;
        m &= m - 1;
        const int ej1 = m ? e0 + __ffs(m) - 1 : -1;
        m &= m - 1;
        const double4 dummy = make_double4(0.0, 0.0, 0.0, ln.iL);
        const double4 c0_ = lane < sq_.ecnt[ej0] ? sq_.E2[ej0 * 32 + lane] : dummy;
        const double4 c1_ = ej1 >= 0 && lane < sq_.ecnt[ej1] ? sq_.E2[ej1 * 32 + lane] : dummy;
        hpr_line_clip(ln, c0_);
        hpr_line_clip(ln, c1_);
      }
      empty = hpr_line_reduce(ln);
    }
    double lo = ln.lo, hi = ln.hi;
    const double p0x = ln.p0x, p0y = ln.p0y, dx = ln.dx, dy = ln.dy;
    // the box |p0 + t d| <= BOX only matters on the side the objective pushes to while that side is
    // still unbounded (see hpr_resolve)
    const bool up = c0 * dx + c1 * dy > 0.0;
    const double tsel = up ? hi : lo;
    if (!empty && !(fabs(p0x + tsel * dx) <= HPR_BOX && fabs(p0y + tsel * dy) <= HPR_BOX)) {
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

// EXACT: HPR_EXACT_WARPS independent warps per block (nothing couples them after the start: they
// re-solve at very different times) sharing the view's extremes E2 + boxes in shared memory - a
// re-solve touches a dozen E blocks one after the other, so their latency is its critical path.
// HPR_EXACT_POINTS points per warp (the other lanes only help in the re-solves, which are warp wide
// and serialised).  Warps [0, 2 NET) of a view own the surviving extremes (two half blocks of E2
// each), the others 16 consecutive survivors of S.  ncu (profiles/r02q_*): fixed-latency fp64
// dependency waits dominate (3 of 7 stall cycles per issue), loads do not.
static constexpr int HPR_EXACT_POINTS = 16;  // measured: 8 -> -2 % / -6 %, 4 -> -12 % / -22 % (2 / 8 views)
static constexpr int HPR_EXACT_WARPS = 8;    // 126 registers, no spills; 10 warps (96 registers, spills) was slower
static constexpr int HPR_EXACT_SMEM_E = HPR_NET * 32 * 32 + HPR_NET * 6 * 8 + 256;  // E2, ebox, ecnt
static constexpr int HPR_EXACT_SMEM_W = HPR_ST * 32 + 32 * 6 * 8;                   // per warp: tile, boxes
static constexpr int HPR_EXACT_SMEM = HPR_EXACT_SMEM_E + HPR_EXACT_WARPS * HPR_EXACT_SMEM_W;
static_assert(HPR_NET * 4 <= 256, "ecnt region");
__global__ void __launch_bounds__(32 * HPR_EXACT_WARPS, 2)
hpr_exact_kernel(int N, HprWs ws, uint8_t* __restrict__ vis) {
  extern __shared__ __align__(32) uint8_t hpr_smem[];
  double4* sE2 = (double4*)hpr_smem;
  double* sebox = (double*)(hpr_smem + HPR_NET * 32 * 32);
  int* secnt = (int*)(hpr_smem + HPR_NET * 32 * 32 + HPR_NET * 6 * 8);
  const int v = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double4* sq = (double4*)(hpr_smem + HPR_EXACT_SMEM_E + warp * HPR_EXACT_SMEM_W);
  double* sbx = (double*)(sq + HPR_ST);
  HprSeq seq;
  seq.nS = ws.nS[v];
  seq.nT = hpr_tiles_dev(seq.nS);
  // warps past the last point of the view: the whole block leaves before the barrier
  constexpr int EW = 32 / HPR_EXACT_POINTS;  // warps per E block
  if (blockIdx.x * HPR_EXACT_WARPS >= EW * HPR_NET + (seq.nS + HPR_EXACT_POINTS - 1) / HPR_EXACT_POINTS)
    return;
  for (int k = threadIdx.x; k < HPR_NET * 32; k += blockDim.x)
    sE2[k] = ws.E2[(size_t)v * HPR_NET * 32 + k];
  for (int k = threadIdx.x; k < HPR_NET * 6; k += blockDim.x)
    sebox[k] = ws.ebox[(size_t)v * HPR_NET * 6 + k];
  for (int k = threadIdx.x; k < HPR_NET; k += blockDim.x) secnt[k] = ws.ecnt[v * HPR_NET + k];
  __syncthreads();  // the only block-wide synchronisation
  const int vb = blockIdx.x * HPR_EXACT_WARPS + warp;
  seq.S = ws.S + (size_t)v * N;
  seq.sbox = ws.sbox + (size_t)v * hpr_tiles_dev(N) * 6;
  seq.E2 = sE2;
  seq.ebox = sebox;
  seq.ecnt = secnt;
  double4 me = make_double4(0.0, 0.0, 0.0, -1.0);
  double2 ab0 = make_double2(0.0, 0.0);
  bool active = false;
  if (vb < EW * HPR_NET) {
    const int t = vb / EW, r0 = (vb % EW) * HPR_EXACT_POINTS;
    const int cnt = seq.ecnt[t];
    if (r0 >= cnt) return;
    if (lane < HPR_EXACT_POINTS && r0 + lane < cnt) {
      me = seq.E2[t * 32 + r0 + lane];
      // an extreme that the filter itself rejected stays a (valid) constraint only
      active = ws.surv[(size_t)v * N + (int)me.w] != 0;
      ab0 = ws.ab[(size_t)v * N + (int)me.w];
    }
    const int first = (int)seq.E2[t * 32 + r0].w;
    seq.home = min(seq.nT - 1, ws.celloffs[v * (HPR_M2 + 1) + ws.mcell[(size_t)v * N + first]] / HPR_ST);
  } else {
    const int p0 = (vb - EW * HPR_NET) * HPR_EXACT_POINTS;
    if (p0 >= seq.nS) return;
    if (lane < HPR_EXACT_POINTS && p0 + lane < seq.nS) {
      me = seq.S[p0 + lane];
      ab0 = ws.Sab[(size_t)v * N + p0 + lane];
      active = true;
    }
    seq.home = p0 / HPR_ST;
  }
  double a = ab0.x, b = ab0.y;
  bool feasible = active;
  for (int s0 = 0; s0 < seq.nT; s0 += 32) {
    if (!__any_sync(0xffffffffu, feasible)) break;
    // the boxes of the next 32 steps
    __syncwarp();
    if (s0 + lane < seq.nT) {
      const double* src = seq.sbox + (size_t)hpr_tile_at(s0 + lane, seq.home, seq.nT) * 6;
#pragma unroll
      for (int j = 0; j < 6; ++j) sbx[lane * 6 + j] = src[j];
    }
    __syncwarp();
    for (int s = s0; s < min(seq.nT, s0 + 32); ++s) {
      const bool need =
          feasible && hpr_box_reach(sbx + (s - s0) * 6, me.x, me.y, me.z, a, b, fabs(a), fabs(b));
      if (!__any_sync(0xffffffffu, need)) continue;
      const int k0 = hpr_tile_at(s, seq.home, seq.nT) * HPR_ST;
      const int cnt = min(HPR_ST, seq.nS - k0);
      __syncwarp();
      for (int t = lane; t < cnt; t += 32) sq[t] = seq.S[k0 + t];
      __syncwarp();
      // Pre-test of the scan: constraint j cuts the optimum off iff  z_j - x_j a - y_j b  exceeds the
      // same expression of the lane's own point.  Evaluated in that form (two FMAs per constraint
      // instead of the difference form below) against a threshold lowered by a rounding margin
      // taken from the tile's box, it can only err towards "look again": the decision itself is
      // always taken by the difference form.
      const double* bx = sbx + (s - s0) * 6;
      const double zm = fmax(fabs(bx[4]), fabs(bx[5])) + fabs(me.z);
      const double xm = fmax(fabs(bx[0]), fabs(bx[1])) + fabs(me.x);
      const double ym = fmax(fabs(bx[2]), fabs(bx[3])) + fabs(me.y);
      auto threshold = [&]() {
        return feasible ? fma(-me.y, b, fma(-me.x, a, me.z)) - HPR_EPS * (zm + xm * fabs(a) + ym * fabs(b))
                        : INFINITY;
      };
      double thr = threshold();
      for (int t0 = 0; t0 < cnt; t0 += 4) {
        // fast path: none of the next four constraints can cut off any lane's optimum
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (t0 + j < cnt) {
            const double2 xy = *reinterpret_cast<const double2*>(&sq[t0 + j]);
            any |= fma(-xy.y, b, fma(-xy.x, a, sq[t0 + j].z)) > thr;
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
            hpr_resolve_tiles(seq, sq, t, s, cj, L, lane, me, a, b, feasible);
          }
        }
        thr = threshold();  // (a, b) or feasible may have changed
      }
    }
  }
  if (active) vis[(size_t)v * N + (int)me.w] = feasible ? 1 : 0;
}

int hpr_launch(const float* points, int N, int V, const double* frames_dev, double radius,
               void* workspace, uint8_t* vis, cudaStream_t stream) {
  PDR_CHECK_ARG(N > 0 && V > 0, "hidden point removal: empty input");
  PDR_CHECK_ARG(((uintptr_t)workspace & 31) == 0, "hidden point removal: workspace must be 32-byte aligned");
  HprWs ws = hpr_carve(workspace, V, N);
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(hpr_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HPR_G2 * (int)sizeof(double4)));
    PDR_CUDA(cudaFuncSetAttribute(hpr_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HPR_EXACT_SMEM));
    configured = true;
  }
  PDR_CUDA(cudaMemsetAsync(ws.isE, 0, (size_t)V * N, stream));
  PDR_CUDA(cudaMemsetAsync(ws.cellcnt, 0, (size_t)V * HPR_M2 * 2 * sizeof(int), stream));
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
  hpr_eblock_kernel<<<dim3(HPR_NET, V), 32, 0, stream>>>(N, ws);
  PDR_COUNT_LAUNCH();
  hpr_filter_kernel<<<dim3(cdiv(N, 32 * HPR_FILTER_WARPS), V), 32 * HPR_FILTER_WARPS,
                      HPR_G2 * sizeof(double4), stream>>>(N, ws);
  PDR_COUNT_LAUNCH();
  hpr_count_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(N, V, ws, vis);
  PDR_COUNT_LAUNCH();
  hpr_offsets_kernel<<<V, 1024, 0, stream>>>(ws);
  PDR_COUNT_LAUNCH();
  hpr_scatter_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(N, V, ws);
  PDR_COUNT_LAUNCH();
  hpr_place_kernel<<<cdiv((size_t)V * N, 256), 256, 0, stream>>>(N, V, ws);
  PDR_COUNT_LAUNCH();
  hpr_tilebox_kernel<<<dim3(cdiv(hpr_tiles(N), 8), V), 256, 0, stream>>>(N, ws);
  PDR_COUNT_LAUNCH();
  const int exact_warps = (32 / HPR_EXACT_POINTS) * HPR_NET + cdiv(N, HPR_EXACT_POINTS);
  hpr_exact_kernel<<<dim3(cdiv(exact_warps, HPR_EXACT_WARPS), V), 32 * HPR_EXACT_WARPS, HPR_EXACT_SMEM,
                     stream>>>(N, ws, vis);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
