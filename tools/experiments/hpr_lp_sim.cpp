// CPU simulation of the hidden-point-removal LP passes of csrc/geom_hpr.cu (same arithmetic, serial):
// counts constraint checks / re-solves / clip evaluations of the exact pass for different constraint
// orders, and compares the per-point decisions.  Findings:
//   * the exact pass costs ~2(|C|-|E|) clips per surviving point in a random order (Seidel's expectation);
//   * Morton-sorted survivors visited neighbour-first (128-constraint tiles, outwards from the warp's
//     own tile) need 1.6-2.1x fewer clips with identical decisions;
//   * bounding boxes per tile of 64 survivors / per block of 8 x 4 cells of E prune the exact pass 3x
//     (scan checks) and 6x (re-solve clips) with identical decisions: built (csrc/geom_hpr.cu);
//   * the same scheme for the FILTER (cloud sorted by E block, ring order around the warp's block) does
//     4x fewer checks and 5x fewer clips here, but on the GPU it executed only 20 % fewer instructions
//     (box tests, interval reductions and the sort eat the saving) and its spatially coherent warps
//     are unbalanced (all-visible warps run 3x longer than the average): 2.29 ms vs 2.18 ms at 8 views
//     with 8 points per warp, 6.3 ms with 32 - not built (profiles/r02u_filter_experiment.md);
//   * a sequence that repeats constraints (E members also listed among the survivors) flips ~4 % of the
//     decisions: a duplicate of the BINDING constraint can test as violated by rounding and its 1-D
//     re-solve is degenerate.  Constraint sequences must be duplicate free.
// Input: a binary file of N x 3 float64 (u, v, w) of one view, written e.g. by
//   python - <<'PY'
//   import numpy as np; from pointdreamer_b200.io_utils import read_ply_xyzrgb, normalize_cloud
//   from pointdreamer_b200.hpr import view_frames; from oracle import camera as ocam
//   xyz = normalize_cloud(read_ply_xyzrgb('tests/golden/clock.ply')[0]); f = view_frames(ocam.create_cameras(8, 1.6, 512)[2])[0]
//   p = xyz.astype(np.float64) - f[0:3]; s = 2 * 100.0 / np.linalg.norm(p, axis=1) - 1; z = p @ f[9:12]
//   np.stack([p @ f[3:6] / z, p @ f[6:9] / z, -1048576.0 / (s * z)], 1).tofile('/tmp/clock_v0.bin')
//   PY
// Build / run:  g++ -O2 -o hpr_lp_sim tools/experiments/hpr_lp_sim.cpp && ./hpr_lp_sim /tmp/clock_v0.bin 40
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>
#include <cstdint>
struct P4 { double x, y, z, w; };
static const double BOX = 1073741824.0;
struct Stats { long long clips = 0, resolves = 0, checks = 0; };
// returns feasible; processes constraints q[start..n) given state (a,b) valid for q[0..start)
static bool lp(const std::vector<P4>& q, int start, int n, const P4& me, double& a, double& b, Stats& st) {
  for (int t = start; t < n; ++t) {
    const P4& c = q[t];
    st.checks++;
    if (c.w == me.w) continue;
    if ((c.x - me.x) * a + (c.y - me.y) * b < c.z - me.z) {
      st.resolves++;
      double nx = c.x - me.x, ny = c.y - me.y, h = c.z - me.z, nn = nx * nx + ny * ny;
      if (!(nn > 0)) return false;
      double sc = h / nn, p0x = nx * sc, p0y = ny * sc, dx = -ny, dy = nx;
      double lo = -INFINITY, hi = INFINITY;
      if (dx != 0) { double t1 = (-BOX - p0x) / dx, t2 = (BOX - p0x) / dx; lo = fmax(lo, fmin(t1, t2)); hi = fmin(hi, fmax(t1, t2)); }
      else if (fabs(p0x) > BOX) return false;
      if (dy != 0) { double t1 = (-BOX - p0y) / dy, t2 = (BOX - p0y) / dy; lo = fmax(lo, fmin(t1, t2)); hi = fmin(hi, fmax(t1, t2)); }
      else if (fabs(p0y) > BOX) return false;
      for (int k = 0; k < t; ++k) {
        const P4& e = q[k];
        if (e.w == me.w) continue;
        st.clips++;
        double ax = e.x - me.x, ay = e.y - me.y, ah = e.z - me.z;
        double den = ax * dx + ay * dy, rhs = ah - (ax * p0x + ay * p0y);
        if (den > 0) { if (rhs > lo * den) lo = rhs / den; }
        else if (den < 0) { if (rhs > hi * den) hi = rhs / den; }
        else if (rhs > 0) lo = INFINITY;
      }
      if (!(lo <= hi)) return false;
      double tt = (1.0 * dx + 0.5 * dy > 0) ? hi : lo;
      a = p0x + tt * dx; b = p0y + tt * dy;
    }
  }
  return true;
}
static uint32_t morton(uint32_t x, uint32_t y) {
  auto sp = [](uint32_t v) { v &= 0xFFFF; v = (v | (v << 8)) & 0x00FF00FF; v = (v | (v << 4)) & 0x0F0F0F0F; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v; };
  return sp(x) | (sp(y) << 1);
}
int main(int argc, char** argv) {
  // input: binary file of N x 3 doubles (u, v, w) for one view
  FILE* f = fopen(argv[1], "rb");
  int G = argc > 2 ? atoi(argv[2]) : 40;
  std::vector<P4> Q;
  double buf[3];
  while (fread(buf, 8, 3, f) == 3) Q.push_back({buf[0], buf[1], buf[2], (double)Q.size()});
  fclose(f);
  const int N = Q.size();
  double u0 = 1e300, v0 = 1e300, u1 = -1e300, v1 = -1e300;
  for (auto& p : Q) { u0 = fmin(u0, p.x); u1 = fmax(u1, p.x); v0 = fmin(v0, p.y); v1 = fmax(v1, p.y); }
  auto cell = [&](const P4& p, int g) { int cu = std::min(g - 1, std::max(0, (int)((p.x - u0) / (u1 - u0) * g))); int cv = std::min(g - 1, std::max(0, (int)((p.y - v0) / (v1 - v0) * g))); return cv * g + cu; };
  std::vector<int> cm(G * G, -1);
  for (int i = 0; i < N; ++i) { int c = cell(Q[i], G); if (cm[c] < 0 || Q[i].z > Q[cm[c]].z) cm[c] = i; }
  std::vector<P4> E; std::vector<char> isE(N, 0);
  for (int p = 0; p < G * G; ++p) { int c = (int)(((long long)p * 1543 + 7) % (G * G)); if (cm[c] >= 0) { E.push_back(Q[cm[c]]); isE[cm[c]] = 1; } }
  const int nE = E.size();
  // filter
  Stats sf; std::vector<char> surv(N); std::vector<double> A(N), B(N);
  for (int i = 0; i < N; ++i) { double a = BOX, b = BOX; surv[i] = lp(E, 0, nE, Q[i], a, b, sf); A[i] = a; B[i] = b; }
  int ns = 0; for (int i = 0; i < N; ++i) ns += surv[i];
  printf("N %d  nE %d  survivors %d | filter: checks %lld resolves %lld clips %lld\n", N, nE, ns, sf.checks, sf.resolves, sf.clips);
  // filter, variant (b): E in spatial blocks of 8 x 4 cells with boxes; a point visits the blocks in ring
  // order around its own block; scan and re-solve pruned by the boxes, empty interval stops a re-solve
  {
    const int BW = 8, BH = 4, NBX = G / BW, NBY = G / BH, NB = NBX * NBY;
    struct Bx { double x0, x1, y0, y1, z0, z1; };
    std::vector<std::vector<P4>> blk(NB); std::vector<Bx> bx(NB);
    for (int c = 0; c < G * G; ++c) if (cm[c] >= 0) { int cu = c % G, cv = c / G; blk[(cv / BH) * NBX + cu / BW].push_back(Q[cm[c]]); }
    for (int t = 0; t < NB; ++t) { Bx b{1e300, -1e300, 1e300, -1e300, 1e300, -1e300}; for (auto& p : blk[t]) { b.x0 = fmin(b.x0, p.x); b.x1 = fmax(b.x1, p.x); b.y0 = fmin(b.y0, p.y); b.y1 = fmax(b.y1, p.y); b.z0 = fmin(b.z0, p.z); b.z1 = fmax(b.z1, p.z); } bx[t] = b; }
    const double EPS = 64 * 2.220446049250313e-16;
    auto gmax = [&](const Bx& b, const P4& me, double a, double bb, double ma, double mb) {
      double xs = (a > 0 ? b.x0 : b.x1) - me.x, ys = (bb > 0 ? b.y0 : b.y1) - me.y;
      double g = (b.z1 - me.z) - xs * a - ys * bb;
      double mx = fmax(fabs(b.x0 - me.x), fabs(b.x1 - me.x)), my = fmax(fabs(b.y0 - me.y), fabs(b.y1 - me.y));
      return g + EPS * (fmax(fabs(b.z1 - me.z), fabs(b.z0 - me.z)) + mx * ma + my * mb);
    };
    long long checks = 0, clips = 0, resolves = 0, boxt = 0; int ns2 = 0, mism = 0; double maxd = 0;
    for (int mode = 0; mode < 2; ++mode) {
      checks = clips = resolves = boxt = 0; ns2 = 0; mism = 0; maxd = 0;
      for (int i = 0; i < N; ++i) {
        const P4& me = Q[i];
        int c = cell(me, G), hb = ((c / G) / BH) * NBX + (c % G) / BW, hx = hb % NBX, hy = hb / NBX;
        std::vector<int> order(NB); std::iota(order.begin(), order.end(), 0);
        if (mode == 0) std::sort(order.begin(), order.end(), [&](int x, int y) { int dx = std::max(abs(x % NBX - hx) * 2, abs(x / NBX - hy)), dy = std::max(abs(y % NBX - hx) * 2, abs(y / NBX - hy)); return dx != dy ? dx < dy : x < y; });
        else { for (int k = 0; k < NB; ++k) order[k] = (k * 17 + 7 + hb) % NB; }
        double a = BOX, b = BOX; bool feas = true;
        for (int s = 0; s < NB && feas; ++s) {
          const int t = order[s]; if (blk[t].empty()) continue;
          boxt++;
          if (!(gmax(bx[t], me, a, b, fabs(a), fabs(b)) > 0)) continue;
          for (size_t k = 0; k < blk[t].size() && feas; ++k) {
            const P4& cj = blk[t][k]; checks++;
            if (cj.w == me.w) continue;
            if (!((cj.x - me.x) * a + (cj.y - me.y) * b < cj.z - me.z)) continue;
            resolves++;
            double nx = cj.x - me.x, ny = cj.y - me.y, h = cj.z - me.z, nn = nx * nx + ny * ny;
            if (!(nn > 0)) { feas = false; break; }
            double sc = h / nn, p0x = nx * sc, p0y = ny * sc, dx = -ny, dy = nx, lo = -INFINITY, hi = INFINITY;
            auto clip = [&](const P4& e) { if (e.w == me.w) return; clips++; double ax = e.x - me.x, ay = e.y - me.y, ah = e.z - me.z; double den = ax * dx + ay * dy, rhs = ah - (ax * p0x + ay * p0y); if (den > 0) { if (rhs > lo * den) lo = rhs / den; } else if (den < 0) { if (rhs > hi * den) hi = rhs / den; } else if (rhs > 0) lo = INFINITY; };
            for (size_t kk = 0; kk < k; ++kk) clip(blk[t][kk]);
            for (int s2 = 0; s2 < s && lo <= hi; ++s2) {
              const int t2 = order[s2]; if (blk[t2].empty()) continue;
              if (std::isfinite(lo) && std::isfinite(hi)) {
                boxt++;
                double g1 = gmax(bx[t2], me, p0x + lo * dx, p0y + lo * dy, fabs(p0x) + fabs(lo * dx), fabs(p0y) + fabs(lo * dy));
                double g2 = gmax(bx[t2], me, p0x + hi * dx, p0y + hi * dy, fabs(p0x) + fabs(hi * dx), fabs(p0y) + fabs(hi * dy));
                if (g1 <= 0 && g2 <= 0) continue;
              }
              for (auto& e : blk[t2]) clip(e);
            }
            if (dx != 0) { double t1 = (-BOX - p0x) / dx, t2 = (BOX - p0x) / dx; lo = fmax(lo, fmin(t1, t2)); hi = fmin(hi, fmax(t1, t2)); } else if (fabs(p0x) > BOX) { feas = false; break; }
            if (dy != 0) { double t1 = (-BOX - p0y) / dy, t2 = (BOX - p0y) / dy; lo = fmax(lo, fmin(t1, t2)); hi = fmin(hi, fmax(t1, t2)); } else if (fabs(p0y) > BOX) { feas = false; break; }
            if (!(lo <= hi)) { feas = false; break; }
            double tt = (1.0 * dx + 0.5 * dy > 0) ? hi : lo; a = p0x + tt * dx; b = p0y + tt * dy;
          }
        }
        ns2 += feas; if (feas != (bool)surv[i]) mism++;
        if (feas && surv[i]) maxd = fmax(maxd, fmax(fabs(a - A[i]) / fmax(1.0, fabs(A[i])), fabs(b - B[i]) / fmax(1.0, fabs(B[i]))));
      }
      printf("filter (E blocks, %s order): survivors %d mismatches %d max rel d(a,b) %.2e | checks %lld resolves %lld clips %lld box tests %lld\n", mode == 0 ? "ring" : "pseudo-random", ns2, mism, maxd, checks, resolves, clips, boxt);
    }
  }
  // exact, order (a): stride permutation
  int stride = 7919 % N, offset = N / 3;
  std::vector<P4> C = E; std::vector<int> rest;
  for (int p = 0; p < N; ++p) { int n = (int)(((long long)p * stride + offset) % N); if (surv[n] && !isE[n]) { C.push_back(Q[n]); rest.push_back(n); } }
  const int nC = C.size();
  Stats sa; int vis = 0; std::vector<char> resA(N,0);
  for (int k = 0; k < nC; ++k) { int i = (int)C[k].w; if (!surv[i]) continue; double a = A[i], b = B[i]; bool r = lp(C, nE, nC, C[k], a, b, sa); resA[i]=r; vis += r; }
  printf("exact (random order): nC %d visible %d | checks %lld resolves %lld clips %lld  (clips/pt %.0f)\n", nC, vis, sa.checks, sa.resolves, sa.clips, (double)sa.clips / nC);
  // exact, order (b): Morton-sorted rest; each point's sequence = E, then blocks expanding outward from its own position
  std::vector<int> ord = rest;
  std::sort(ord.begin(), ord.end(), [&](int x, int y) { uint32_t mx = morton((uint32_t)((Q[x].x - u0) / (u1 - u0) * 1023), (uint32_t)((Q[x].y - v0) / (v1 - v0) * 1023)); uint32_t my = morton((uint32_t)((Q[y].x - u0) / (u1 - u0) * 1023), (uint32_t)((Q[y].y - v0) / (v1 - v0) * 1023)); return mx != my ? mx < my : x < y; });
  const int BL = 32; const int nb = (ord.size() + BL - 1) / BL;
  Stats sb; int vis2 = 0;
  std::vector<P4> seq;
  for (int wb = 0; wb < nb; ++wb) {
    // sequence for the points of block wb: E, then blocks wb, wb-1, wb+1, wb-2, ...
    seq.assign(E.begin(), E.end());
    for (int d = 0; d < nb; ++d) {
      int cand[2] = {wb - d, wb + d};
      for (int s = 0; s < (d == 0 ? 1 : 2); ++s) { int bb = cand[s]; if (bb < 0 || bb >= nb) continue; for (int k = bb * BL; k < std::min<int>((bb + 1) * BL, ord.size()); ++k) seq.push_back(Q[ord[k]]); }
    }
    for (int k = wb * BL; k < std::min<int>((wb + 1) * BL, ord.size()); ++k) { int i = ord[k]; double a = A[i], b = B[i]; vis2 += lp(seq, nE, seq.size(), Q[i], a, b, sb); }
  }
  // E members themselves
  for (int k = 0; k < nE; ++k) { int i = (int)E[k].w; if (!surv[i]) continue; double a = A[i], b = B[i]; vis2 += lp(C, nE, nC, E[k], a, b, sb); }
  printf("exact (neighbour-first): visible %d | checks %lld resolves %lld clips %lld  (clips/pt %.0f)\n", vis2, sb.checks, sb.resolves, sb.clips, (double)sb.clips / nC);
  // exact, order (c): E prefix + ALL survivors sorted by 64x64 Morton cell then index; warps of 16 points,
  // tiles of 128 constraints visited home tile first then alternating right/left
  {
    std::vector<int> S; for (int i = 0; i < N; ++i) if (surv[i] && !isE[i]) S.push_back(i);
    auto key = [&](int x) { int cu = std::min(63, std::max(0, (int)((Q[x].x - u0) / (u1 - u0) * 64))); int cv = std::min(63, std::max(0, (int)((Q[x].y - v0) / (v1 - v0) * 64))); return morton(cu, cv); };
    std::sort(S.begin(), S.end(), [&](int x, int y) { uint32_t a = key(x), b = key(y); return a != b ? a < b : x < y; });
    const int nS = S.size(), TL = 128, nT = (nS + TL - 1) / TL, PW = 16;
    Stats sc; int vis3 = 0; std::vector<P4> seq;
    for (int s0 = 0; s0 < nS; s0 += PW) {
      int home = s0 / TL, ta = home, tb = home;
      seq.assign(E.begin(), E.end());
      for (int step = 0; step < nT; ++step) {
        int tau; bool right;
        if (step == 0) { tau = home; right = true; }
        else { right = (tb < nT) && (ta == 0 || (step & 1)); tau = right ? tb : ta - 1; }
        for (int k = tau * TL; k < std::min(nS, (tau + 1) * TL); ++k) seq.push_back(Q[S[k]]);
        if (right) tb = tau + 1; else ta = tau;
      }
      for (int k = s0; k < std::min(nS, s0 + PW); ++k) { int i = S[k]; double a = A[i], b = B[i]; bool r = lp(seq, nE, seq.size(), Q[i], a, b, sc); if (r != (bool)resA[i]) { static int shown=0; if (shown++<5) printf("  mismatch point %d isE %d resA %d resC %d\n", i, (int)isE[i], (int)resA[i], (int)r);} vis3 += r; }
    }
    // E survivors: warps of 16 in E order, home = tile of the first point's Morton cell
    std::vector<int> Es; for (int k = 0; k < nE; ++k) if (surv[(int)E[k].w]) Es.push_back((int)E[k].w);
    for (size_t e0 = 0; e0 < Es.size(); e0 += PW) {
      uint32_t kk = key(Es[e0]);
      int pos = std::lower_bound(S.begin(), S.end(), kk, [&](int x, uint32_t v) { return key(x) < v; }) - S.begin();
      int home = std::min(nT - 1, pos / TL), ta = home, tb = home;
      seq.assign(E.begin(), E.end());
      for (int step = 0; step < nT; ++step) {
        int tau; bool right;
        if (step == 0) { tau = home; right = true; }
        else { right = (tb < nT) && (ta == 0 || (step & 1)); tau = right ? tb : ta - 1; }
        for (int k = tau * TL; k < std::min(nS, (tau + 1) * TL); ++k) seq.push_back(Q[S[k]]);
        if (right) tb = tau + 1; else ta = tau;
      }
      for (size_t k = e0; k < std::min(Es.size(), e0 + PW); ++k) { int i = Es[k]; double a = A[i], b = B[i]; bool r = lp(seq, nE, seq.size(), Q[i], a, b, sc); if (r != (bool)resA[i]) printf("  E mismatch %d\n", i); vis3 += r; }
    }
    printf("exact (tile-128 neighbour-first, all survivors): nS %d visible %d | checks %lld resolves %lld clips %lld (clips/pt %.0f)\n", nS, vis3, sc.checks, sc.resolves, sc.clips, (double)sc.clips / nS);
  }
  // exact, order (d): as (c) plus bounding boxes per 128-constraint tile: a tile is skipped in the scan
  // when no member can violate any lane's optimum, and in a re-solve when no member can move lo or hi
  {
    const int TLs[3] = {128, 64, 32};
    for (int TL : TLs) {
    std::vector<int> S; for (int i = 0; i < N; ++i) if (surv[i] && !isE[i]) S.push_back(i);
    auto key = [&](int x) { int cu = std::min(63, std::max(0, (int)((Q[x].x - u0) / (u1 - u0) * 64))); int cv = std::min(63, std::max(0, (int)((Q[x].y - v0) / (v1 - v0) * 64))); return morton(cu, cv); };
    std::sort(S.begin(), S.end(), [&](int x, int y) { uint32_t a = key(x), b = key(y); return a != b ? a < b : x < y; });
    const int nS = S.size(), nT = (nS + TL - 1) / TL, PW = 16;
    struct Bx { double x0, x1, y0, y1, z0, z1; };
    std::vector<Bx> bx(nT);
    for (int t = 0; t < nT; ++t) { Bx b{1e300, -1e300, 1e300, -1e300, 1e300, -1e300}; for (int k = t * TL; k < std::min(nS, (t + 1) * TL); ++k) { const P4& p = Q[S[k]]; b.x0 = fmin(b.x0, p.x); b.x1 = fmax(b.x1, p.x); b.y0 = fmin(b.y0, p.y); b.y1 = fmax(b.y1, p.y); b.z0 = fmin(b.z0, p.z); b.z1 = fmax(b.z1, p.z); } bx[t] = b; }
    const double EPS = 64 * 2.220446049250313e-16;
    // max over the box of (z - w) - (x - u) a - (y - v) b, plus a rounding margin
    auto gmax = [&](const Bx& b, const P4& me, double a, double bb, double ma, double mb) {
      double xs = (a > 0 ? b.x0 : b.x1) - me.x, ys = (bb > 0 ? b.y0 : b.y1) - me.y;
      double g = (b.z1 - me.z) - xs * a - ys * bb;
      double mx = fmax(fabs(b.x0 - me.x), fabs(b.x1 - me.x)), my = fmax(fabs(b.y0 - me.y), fabs(b.y1 - me.y));
      double mag = fmax(fabs(b.z1 - me.z), fabs(b.z0 - me.z)) + mx * ma + my * mb;
      return g + EPS * mag;
    };
    // E sorted by the Morton code of its 64x64 cell, tiles of ET with boxes (used by the re-solves only)
    const int ET = 32;
    std::vector<P4> E2 = E; std::sort(E2.begin(), E2.end(), [&](const P4& x, const P4& y) { uint32_t a = key((int)x.w), b = key((int)y.w); return a != b ? a < b : x.w < y.w; });
    const int nET = (nE + ET - 1) / ET; std::vector<Bx> ebx(nET);
    for (int t = 0; t < nET; ++t) { Bx b{1e300, -1e300, 1e300, -1e300, 1e300, -1e300}; for (int k = t * ET; k < std::min(nE, (t + 1) * ET); ++k) { const P4& p = E2[k]; b.x0 = fmin(b.x0, p.x); b.x1 = fmax(b.x1, p.x); b.y0 = fmin(b.y0, p.y); b.y1 = fmax(b.y1, p.y); b.z0 = fmin(b.z0, p.z); b.z1 = fmax(b.z1, p.z); } ebx[t] = b; }
    long long clips = 0, boxt = 0, resolves = 0, checks = 0, tiles_scanned = 0, tiles_skipped = 0, rtiles_done = 0, rtiles_skip = 0, eclips = 0;
    int vis4 = 0, mism = 0;
    auto run_warp = [&](const std::vector<int>& pts, int home) {
      std::vector<int> order; int ta = home, tb = home;
      for (int step = 0; step < nT; ++step) { int tau; bool right; if (step == 0) { tau = home; right = true; } else { right = (tb < nT) && (ta == 0 || (step & 1)); tau = right ? tb : ta - 1; } order.push_back(tau); if (right) tb = tau + 1; else ta = tau; }
      const int np = pts.size();
      std::vector<double> a(np), b(np); std::vector<char> feas(np, 1);
      for (int l = 0; l < np; ++l) { a[l] = A[pts[l]]; b[l] = B[pts[l]]; }
      for (int step = 0; step < nT; ++step) {
        const int tau = order[step]; const Bx& bb = bx[tau];
        bool need = false;
        for (int l = 0; l < np; ++l) if (feas[l]) { boxt++; if (gmax(bb, Q[pts[l]], a[l], b[l], fabs(a[l]), fabs(b[l])) > 0) need = true; }
        if (!need) { tiles_skipped++; continue; }
        tiles_scanned++;
        for (int k = tau * TL; k < std::min(nS, (tau + 1) * TL); ++k) {
          const P4& c = Q[S[k]];
          for (int l = 0; l < np; ++l) {
            if (!feas[l]) continue;
            const P4& me = Q[pts[l]];
            checks++;
            if (c.w == me.w) continue;
            if ((c.x - me.x) * a[l] + (c.y - me.y) * b[l] < c.z - me.z) {
              resolves++;
              double nx = c.x - me.x, ny = c.y - me.y, h = c.z - me.z, nn = nx * nx + ny * ny;
              if (!(nn > 0)) { feas[l] = 0; continue; }
              double sc = h / nn, p0x = nx * sc, p0y = ny * sc, dx = -ny, dy = nx;
              double lo = -INFINITY, hi = INFINITY;
              if (dx != 0) { double t1 = (-BOX - p0x) / dx, t2 = (BOX - p0x) / dx; lo = fmax(lo, fmin(t1, t2)); hi = fmin(hi, fmax(t1, t2)); } else if (fabs(p0x) > BOX) { feas[l] = 0; continue; }
              if (dy != 0) { double t1 = (-BOX - p0y) / dy, t2 = (BOX - p0y) / dy; lo = fmax(lo, fmin(t1, t2)); hi = fmin(hi, fmax(t1, t2)); } else if (fabs(p0y) > BOX) { feas[l] = 0; continue; }
              auto clip = [&](const P4& e) {
                if (e.w == me.w) return;
                clips++;
                double ax = e.x - me.x, ay = e.y - me.y, ah = e.z - me.z;
                double den = ax * dx + ay * dy, rhs = ah - (ax * p0x + ay * p0y);
                if (den > 0) { if (rhs > lo * den) lo = rhs / den; } else if (den < 0) { if (rhs > hi * den) hi = rhs / den; } else if (rhs > 0) lo = INFINITY;
              };
              auto cantouch = [&](const Bx& b2) {
                if (!(std::isfinite(lo) && std::isfinite(hi))) return true;
                boxt++;
                double alo = p0x + lo * dx, blo = p0y + lo * dy, ahi = p0x + hi * dx, bhi = p0y + hi * dy;
                double g1 = gmax(b2, me, alo, blo, fabs(p0x) + fabs(lo * dx), fabs(p0y) + fabs(lo * dy));
                double g2 = gmax(b2, me, ahi, bhi, fabs(p0x) + fabs(hi * dx), fabs(p0y) + fabs(hi * dy));
                return !(g1 <= 0 && g2 <= 0);
              };
              // current tile prefix first, then the earlier tiles (nearest first), then the E tiles
              for (int kk = tau * TL; kk < k; ++kk) clip(Q[S[kk]]);
              for (int s2 = 0; s2 < step; ++s2) {
                const int t2 = order[s2];
                if (!cantouch(bx[t2])) { rtiles_skip++; continue; }
                rtiles_done++;
                for (int kk = t2 * TL; kk < std::min(nS, (t2 + 1) * TL); ++kk) clip(Q[S[kk]]);
              }
              for (int t2 = 0; t2 < nET; ++t2) {
                if (!cantouch(ebx[t2])) continue;
                for (int kk = t2 * ET; kk < std::min(nE, (t2 + 1) * ET); ++kk) { long long c0 = clips; clip(E2[kk]); eclips += clips - c0; }
              }
              for (int s2 = 0; s2 < 0; ++s2) {
                const int t2 = order[s2]; const Bx& b2 = bx[t2];
                const int kend = s2 == step ? k : std::min(nS, (t2 + 1) * TL);
                bool skip = false;
                if (std::isfinite(lo) && std::isfinite(hi)) {
                  boxt++;
                  double alo = p0x + lo * dx, blo = p0y + lo * dy, ahi = p0x + hi * dx, bhi = p0y + hi * dy;
                  double g1 = gmax(b2, me, alo, blo, fabs(p0x) + fabs(lo * dx), fabs(p0y) + fabs(lo * dy));
                  double g2 = gmax(b2, me, ahi, bhi, fabs(p0x) + fabs(hi * dx), fabs(p0y) + fabs(hi * dy));
                  skip = g1 <= 0 && g2 <= 0;
                }
                if (skip) { rtiles_skip++; continue; }
                rtiles_done++;
                for (int kk = t2 * TL; kk < kend; ++kk) clip(Q[S[kk]]);
              }
              if (!(lo <= hi)) { feas[l] = 0; continue; }
              double tt = (1.0 * dx + 0.5 * dy > 0) ? hi : lo;
              a[l] = p0x + tt * dx; b[l] = p0y + tt * dy;
            }
          }
        }
      }
      for (int l = 0; l < np; ++l) { vis4 += feas[l]; if ((bool)feas[l] != (bool)resA[pts[l]]) mism++; }
    };
    for (int s0 = 0; s0 < nS; s0 += PW) { std::vector<int> pts; for (int k = s0; k < std::min(nS, s0 + PW); ++k) pts.push_back(S[k]); run_warp(pts, s0 / TL); }
    std::vector<int> Es; for (int k = 0; k < nE; ++k) if (surv[(int)E[k].w]) Es.push_back((int)E[k].w);
    for (size_t e0 = 0; e0 < Es.size(); e0 += PW) {
      uint32_t kk = key(Es[e0]);
      int pos = std::lower_bound(S.begin(), S.end(), kk, [&](int x, uint32_t v) { return key(x) < v; }) - S.begin();
      std::vector<int> pts; for (size_t k = e0; k < std::min(Es.size(), e0 + PW); ++k) pts.push_back(Es[k]);
      run_warp(pts, std::min(nT - 1, pos / TL));
    }
    printf("exact (boxes, tile %d): visible %d mismatches %d | checks %lld resolves %lld clips %lld (clips/pt %.0f) (E part %lld) box tests %lld | scan tiles %lld scanned %lld skipped | resolve tiles %lld done %lld skipped\n",
           TL, vis4, mism, checks, resolves, clips, (double)clips / nC, eclips, boxt, tiles_scanned, tiles_skipped, rtiles_done, rtiles_skip);
    }
  }
  return 0;
}
