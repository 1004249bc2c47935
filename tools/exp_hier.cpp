// exp_hier.cpp -- CPU design study for the hierarchical Bader assignment (NOT product code, NOT the oracle).
// Reads a cubic-cell density (n^3 doubles, x fastest), computes every point's own full near-grid trajectory
// (same step rule as bader@proc.f90:455-494) and then measures, for candidate GPU schemes:
//   E2  walk lengths of boundary points when a walk may stop at a point whose (2m+1)^3 neighbourhood is
//       uniformly labelled (margin m), and whether the adopted label equals the own-trajectory label;
//   E3  walker counts per level of a corner-agreement hierarchy started at stride L0, and the label errors
//       left after the edge-fix fixed point.
// build: g++ -O2 -fopenmp -ffp-contract=off -o /tmp/exp/exp_hier tools/exp_hier.cpp
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <omp.h>

static int N;
static const double* F;
static double C2L;       // car2lat diagonal (cubic)
static double LID[27];

static inline int wrap(int p) { return p < 0 ? p + N : (p >= N ? p - N : p); }
static inline int lin(int x, int y, int z) { return x + N * (y + N * z); }
static inline int nint_(double x) { return (int)std::lround(x); }

static bool is_max(int x, int y, int z) {
  const double r0 = F[lin(x, y, z)];
  for (int a = -1; a <= 1; a++) for (int b = -1; b <= 1; b++) for (int c = -1; c <= 1; c++)
    if (F[lin(wrap(x + a), wrap(y + b), wrap(z + c))] > r0) return false;
  return true;
}
static int step_ongrid(int x, int y, int z) {
  const double r0 = F[lin(x, y, z)];
  double rm = r0; int best = lin(x, y, z);
  for (int a = -1; a <= 1; a++) for (int b = -1; b <= 1; b++) for (int c = -1; c <= 1; c++) {
    const int q = lin(wrap(x + a), wrap(y + b), wrap(z + c));
    const double rt = r0 + (F[q] - r0) * LID[(a + 1) * 9 + (b + 1) * 3 + (c + 1)];
    if (rt > rm) { rm = rt; best = q; }
  }
  return best;
}

// generic walk: stop(q) >= 0 -> adopt that label (checked on every point entered, not on the start)
template <class Stop>
static int walk(int start, Stop stop, long* nsteps, std::vector<int>& path) {
  int x = start % N, y = (start / N) % N, z = start / (N * N);
  double dr[3] = {0, 0, 0};
  double rhomax = -1e300;
  path.clear();
  int id = start;
  for (;;) {
    const double r0 = F[id];
    const double rxp = F[lin(wrap(x + 1), y, z)], rxm = F[lin(wrap(x - 1), y, z)];
    const double ryp = F[lin(x, wrap(y + 1), z)], rym = F[lin(x, wrap(y - 1), z)];
    const double rzp = F[lin(x, y, wrap(z + 1))], rzm = F[lin(x, y, wrap(z - 1))];
    double gl[3] = {(rxp - rxm) * 0.5, (ryp - rym) * 0.5, (rzp - rzm) * 0.5};
    if (rxp < r0 && rxm < r0) gl[0] = 0;
    if (ryp < r0 && rym < r0) gl[1] = 0;
    if (rzp < r0 && rzm < r0) gl[2] = 0;
    double g[3];
    for (int i = 0; i < 3; i++) g[i] = C2L * (gl[i] * C2L);
    const double gmax = std::max(std::fabs(g[0]), std::max(std::fabs(g[1]), std::fabs(g[2])));
    int nx, ny, nz, nid;
    if (gmax < 1e-30) {
      dr[0] = dr[1] = dr[2] = 0;
      if (is_max(x, y, z)) break;
      nid = step_ongrid(x, y, z);
    } else {
      const double c = 1.0 / gmax;
      int pm[3] = {x, y, z};
      for (int i = 0; i < 3; i++) {
        g[i] = c * g[i];
        const int ng = nint_(g[i]);
        pm[i] += ng;
        dr[i] = dr[i] + g[i] - (double)ng;
        const int nd = nint_(dr[i]);
        pm[i] += nd;
        dr[i] -= (double)nd;
      }
      nid = lin(wrap(wrap(pm[0])), wrap(wrap(pm[1])), wrap(wrap(pm[2])));
    }
    path.push_back(id);
    rhomax = std::max(rhomax, r0);
    if (F[nid] <= rhomax && std::find(path.begin(), path.end(), nid) != path.end()) {
      nid = step_ongrid(x, y, z);
      dr[0] = dr[1] = dr[2] = 0;
    }
    if (nid == id) break;
    id = nid;
    x = id % N; y = (id / N) % N; z = id / (N * N);
    const int s = stop(id);
    if (s >= 0) { *nsteps += (long)path.size(); return s; }
  }
  *nsteps += (long)path.size();
  return id;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: exp_hier rho.bin n celllen [L0]\n"); return 1; }
  N = atoi(argv[2]);
  const double a = atof(argv[3]);
  const long nn = (long)N * N * N;
  std::vector<double> f(nn);
  FILE* fp = fopen(argv[1], "rb");
  if (!fp || fread(f.data(), 8, nn, fp) != (size_t)nn) { fprintf(stderr, "read error\n"); return 1; }
  fclose(fp);
  F = f.data();
  const double l2c = a / N;
  C2L = 1.0 / l2c;
  for (int i = -1; i <= 1; i++) for (int j = -1; j <= 1; j++) for (int k = -1; k <= 1; k++) {
    const double d = l2c * std::sqrt((double)(i * i + j * j + k * k));
    LID[(i + 1) * 9 + (j + 1) * 3 + (k + 1)] = (i || j || k) ? 1.0 / d : 0.0;
  }
  // ---- E1: canonical terminals ----
  std::vector<int> term(nn);
  long tot = 0, longest = 0;
  double t0 = omp_get_wtime();
  {
    char cache[256]; snprintf(cache, sizeof cache, "%s.term", argv[1]);
    FILE* fc = fopen(cache, "rb");
    if (fc && fread(term.data(), 4, nn, fc) == (size_t)nn) { fclose(fc); printf("E1: terminals read from cache\n"); }
    else {
#pragma omp parallel reduction(+ : tot) reduction(max : longest)
      {
        std::vector<int> path;
#pragma omp for schedule(dynamic, 4096)
        for (long s = 0; s < nn; s++) {
          long st = 0;
          term[s] = walk((int)s, [](int) { return -1; }, &st, path);
          tot += st; longest = std::max(longest, st);
        }
      }
      printf("E1: canonical walks: %.2f steps/pt, longest %ld, %.1f s\n", (double)tot / nn, longest, omp_get_wtime() - t0);
      fc = fopen(cache, "wb"); fwrite(term.data(), 4, nn, fc); fclose(fc);
    }
  }
  {
    std::vector<int> u(term); std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end());
    printf("   %zu basins\n", u.size());
  }
  // uniformity radius of every point: largest m in 0..MMAX such that the (2m+1)^3 neighbourhood has one label
  const int MMAX = 4;
  std::vector<unsigned char> urad(nn);
#pragma omp parallel for schedule(static)
  for (long s = 0; s < nn; s++) {
    const int x = s % N, y = (s / N) % N, z = s / ((long)N * N);
    const int t = term[s];
    int m = 0;
    for (; m < MMAX; m++) {
      const int r = m + 1; bool ok = true;
      for (int c = -r; c <= r && ok; c++) for (int b = -r; b <= r && ok; b++) for (int a2 = -r; a2 <= r; a2++) {
        if (std::max(std::abs(a2), std::max(std::abs(b), std::abs(c))) != r) continue;
        if (term[lin(wrap(x + a2), wrap(y + b), wrap(z + c))] != t) { ok = false; break; }
      }
      if (!ok) break;
    }
    urad[s] = (unsigned char)m;
  }
  long cntm[MMAX + 1] = {0};
  for (long s = 0; s < nn; s++) cntm[urad[s]]++;
  printf("   points by uniformity radius (0 = has a different 26-neighbour):");
  for (int m = 0; m <= MMAX; m++) printf(" r%d: %.3f%%", m, 100.0 * cntm[m] / nn);
  printf("\n");
  // ---- E2: walk lengths of edge points (urad == 0) with stop at a point of uniformity radius >= m ----
  for (int m = 1; m <= 3; m++) {
    long nw = 0, st = 0, mism = 0, mx = 0;
    std::vector<long> hist(12, 0);
#pragma omp parallel reduction(+ : nw, st, mism) reduction(max : mx)
    {
      std::vector<int> path; std::vector<long> h(12, 0);
#pragma omp for schedule(dynamic, 4096)
      for (long s = 0; s < nn; s++) {
        if (urad[s] >= m) continue;  // certified itself
        long k = 0;
        const int lab = walk((int)s, [&](int q) { return urad[q] >= m ? term[q] : -1; }, &k, path);
        nw++; st += k; mx = std::max(mx, k);
        if (lab != term[s]) mism++;
        int b = 0; long kk = k; while (kk > 1 && b < 11) { kk >>= 1; b++; } h[b]++;
      }
#pragma omp critical
      for (int i = 0; i < 12; i++) hist[i] += h[i];
    }
    printf("E2 margin %d: walkers %.3f%% of grid, mean steps %.2f, max %ld, label mismatches %ld; log2 hist:", m, 100.0 * nw / nn,
           (double)st / nw, mx, mism);
    for (int i = 0; i < 12; i++) printf(" %ld", hist[i]);
    printf("\n");
  }
  // ---- E4: the GPU's cube certificate: stride-2 cube uniform (8 corners of the TRUE labelling agree, no maximum inside)
  //          and so are its 26 neighbour cubes; walkers = the 7 non-lattice points of every non-uniform cube ----
  {
    const int C = N / 2;
    std::vector<unsigned char> hm(nn, 0);
    { std::vector<int> u(term); std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end()); for (int t : u) hm[t] = 1; }
    std::vector<int> uni((size_t)C * C * C), safe((size_t)C * C * C);
    auto cl = [&](int a, int b, int c) { return ((a + C) % C) + C * (((b + C) % C) + C * ((c + C) % C)); };
    for (int cz = 0; cz < C; cz++) for (int cy = 0; cy < C; cy++) for (int cx = 0; cx < C; cx++) {
      const int x0 = 2 * cx, y0 = 2 * cy, z0 = 2 * cz, x1 = (x0 + 2) % N, y1 = (y0 + 2) % N, z1 = (z0 + 2) % N;
      const int l0 = term[lin(x0, y0, z0)];
      bool u = term[lin(x1, y0, z0)] == l0 && term[lin(x0, y1, z0)] == l0 && term[lin(x1, y1, z0)] == l0 && term[lin(x0, y0, z1)] == l0 &&
               term[lin(x1, y0, z1)] == l0 && term[lin(x0, y1, z1)] == l0 && term[lin(x1, y1, z1)] == l0;
      for (int c = 0; c < 2 && u; c++) for (int b = 0; b < 2 && u; b++) for (int a2 = 0; a2 < 2; a2++) if (hm[lin(x0 + a2, y0 + b, z0 + c)]) { u = false; break; }
      uni[cl(cx, cy, cz)] = u ? l0 : -1;
    }
    for (int cz = 0; cz < C; cz++) for (int cy = 0; cy < C; cy++) for (int cx = 0; cx < C; cx++) {
      const int v = uni[cl(cx, cy, cz)]; int out = v;
      for (int c = -1; c <= 1 && out >= 0; c++) for (int b = -1; b <= 1 && out >= 0; b++) for (int a2 = -1; a2 <= 1; a2++) if (uni[cl(cx + a2, cy + b, cz + c)] != v) { out = -1; break; }
      safe[cl(cx, cy, cz)] = out;
    }
    long nsafe = 0, nnon = 0; for (size_t i = 0; i < safe.size(); i++) { nsafe += safe[i] >= 0; nnon += uni[i] < 0; }
    printf("E4: stride-2 cubes: non-uniform %.2f%%, safe %.2f%%\n", 100.0 * nnon / safe.size(), 100.0 * nsafe / safe.size());
    long nw = 0, st = 0, mism = 0, mx = 0;
#pragma omp parallel reduction(+ : nw, st, mism) reduction(max : mx)
    {
      std::vector<int> path;
#pragma omp for schedule(dynamic, 4096)
      for (long s = 0; s < nn; s++) {
        const int x = s % N, y = (s / N) % N, z = s / ((long)N * N);
        if (!((x | y | z) & 1)) continue;
        if (uni[cl(x >> 1, y >> 1, z >> 1)] >= 0) continue;
        long k = 0;
        const int lab = walk((int)s, [&](int q) { const int qx = q % N, qy = (q / N) % N, qz = q / (N * N); return safe[cl(qx >> 1, qy >> 1, qz >> 1)]; }, &k, path);
        nw++; st += k; mx = std::max(mx, k);
        if (lab != term[s]) mism++;
      }
    }
    printf("E4: walkers %.3f%% of grid, mean steps %.2f, max %ld, label mismatches %ld\n", 100.0 * nw / nn, (double)st / nw, mx, mism);
  }
  // ---- E3: corner-agreement hierarchy from stride L0 ----
  const int L0 = argc > 4 ? atoi(argv[4]) : 32;
  std::vector<unsigned char> hasmax(nn, 0);
  {
    std::vector<int> u(term); std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end());
    for (int t : u) hasmax[t] = 1;
  }
  std::vector<int> lab(nn, -1);
  std::vector<unsigned char> filled(nn, 0);
  long walked = 0;
  for (int z = 0; z < N; z += L0) for (int y = 0; y < N; y += L0) for (int x = 0; x < N; x += L0) { lab[lin(x, y, z)] = term[lin(x, y, z)]; walked++; }
  printf("E3: L0=%d lattice walkers %ld\n", L0, walked);
  for (int s = L0; s >= 2; s >>= 1) {
    const int h = s / 2;
    long nwalk = 0, nfill = 0, ncube = 0, nuni = 0;
    for (int z0 = 0; z0 < N; z0 += s) for (int y0 = 0; y0 < N; y0 += s) for (int x0 = 0; x0 < N; x0 += s) {
      ncube++;
      const int x1 = (x0 + s) % N, y1 = (y0 + s) % N, z1 = (z0 + s) % N;
      const int l0 = lab[lin(x0, y0, z0)];
      bool uni = lab[lin(x1, y0, z0)] == l0 && lab[lin(x0, y1, z0)] == l0 && lab[lin(x1, y1, z0)] == l0 && lab[lin(x0, y0, z1)] == l0 &&
                 lab[lin(x1, y0, z1)] == l0 && lab[lin(x0, y1, z1)] == l0 && lab[lin(x1, y1, z1)] == l0;
      if (uni) {  // no maximum strictly inside the cube [x0,x0+s)^3
        for (int c = 0; c < s && uni; c++) for (int b = 0; b < s && uni; b++) for (int a2 = 0; a2 < s; a2++)
          if (hasmax[lin(x0 + a2, y0 + b, z0 + c)]) { uni = false; break; }
      }
      nuni += uni;
      for (int o = 1; o < 8; o++) {
        const int q = lin(x0 + ((o & 1) ? h : 0), y0 + ((o & 2) ? h : 0), z0 + ((o & 4) ? h : 0));
        if (uni) { lab[q] = l0; filled[q] = 1; nfill++; }
        else { lab[q] = term[q]; filled[q] = 0; nwalk++; }
      }
    }
    printf("   level %2d -> %2d: cubes %ld uniform %.2f%%, walkers %ld (%.3f%% of grid), filled %ld\n", s, h, ncube, 100.0 * nuni / ncube, nwalk,
           100.0 * nwalk / nn, nfill);
    walked += nwalk;
  }
  long wrong = 0;
  for (long s = 0; s < nn; s++) wrong += lab[s] != term[s];
  printf("   before edge fix: walked %.3f%%, wrong labels %ld\n", 100.0 * walked / nn, wrong);
  for (int pass = 1; pass < 100; pass++) {
    std::vector<int> fix;
    for (long s = 0; s < nn; s++) {
      if (!filled[s]) continue;
      const int x = s % N, y = (s / N) % N, z = s / ((long)N * N);
      bool edge = false;
      for (int c = -1; c <= 1 && !edge; c++) for (int b = -1; b <= 1 && !edge; b++) for (int a2 = -1; a2 <= 1; a2++)
        if (lab[lin(wrap(x + a2), wrap(y + b), wrap(z + c))] != lab[s]) { edge = true; break; }
      if (edge) fix.push_back((int)s);
    }
    if (fix.empty()) { printf("   edge fix converged after %d passes\n", pass - 1); break; }
    long changed = 0;
    for (int s : fix) { filled[s] = 0; if (lab[s] != term[s]) { lab[s] = term[s]; changed++; } }
    walked += (long)fix.size();
    printf("   edge pass %d: %zu walkers (%.3f%%), %ld changed\n", pass, fix.size(), 100.0 * fix.size() / nn, changed);
  }
  wrong = 0;
  for (long s = 0; s < nn; s++) wrong += lab[s] != term[s];
  printf("   final: walked %.3f%% of grid, wrong labels %ld\n", 100.0 * walked / nn, wrong);
  return 0;
}
