// Experiment (round 2): early-stop rules of the FAST walkers against the own-trajectory ("canonical") labelling, on the CPU.
// For every edge point of the canonical labelling (a point with a 26-neighbour in another basin) the trajectory is walked
// with a stop at the first point that satisfies the rule; the label it would take there is compared with its own terminal.
//   rule 0: uniform 3x3x3 neighbourhood (round 1)        rule 1: uniform 5x5x5 (the reference's refine_edge margin)
//   rule 2: uniform 7x7x7                                 rule 3: 3x3x3 and |dr| < 0.25
//   rule 4: 5x5x5, or 3x3x3 and |dr| < 0.25               rule 5: no early stop (this model has no revisit rule: its
//   rule 6: stride-2 cube + its 26 neighbour cubes uniform        "wrong" count is the model's own error floor)
// build: g++ -O3 -std=c++17 -fopenmp -ffp-contract=off -o exp_stoprule tools/exp_stoprule.cpp   (from tools/)
// input: tools/exp_stoprule_dump.py <case of tests/sized_cases.py> writes f / terminals / metrics as raw files.
// Results of round 2: profiles/r02_stoprule_model.txt
#include "../oracle/oracle.cpp"
#include <cstdio>
int main(int argc, char** argv) {
  // argv: f.raw term.raw n1 n2 n3 meta.raw(car2lat 9 + lid 27)
  int n[3] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  long nn = (long)n[0] * n[1] * n[2];
  std::vector<double> f(nn); std::vector<int> term(nn);
  FILE* fp = fopen(argv[1], "rb"); fread(f.data(), 8, nn, fp); fclose(fp);
  fp = fopen(argv[2], "rb"); fread(term.data(), 4, nn, fp); fclose(fp);
  double meta[36]; fp = fopen(argv[6], "rb"); fread(meta, 8, 36, fp); fclose(fp);
  // uniformity radius: largest R<=3 with (2R+1)^3 uniform
  std::vector<unsigned char> ur(nn);
  auto lin = [&](int x, int y, int z) { return (long)wrap0(x, n[0]) + (long)n[0] * (wrap0(y, n[1]) + (long)n[1] * wrap0(z, n[2])); };
#pragma omp parallel for schedule(dynamic, 64)
  for (int z = 0; z < n[2]; z++) for (int y = 0; y < n[1]; y++) for (int x = 0; x < n[0]; x++) {
    const int l = term[lin(x, y, z)]; int R = 0;
    for (int r = 1; r <= 3; r++) {
      bool ok = true;
      for (int c = -r; c <= r && ok; c++) for (int b = -r; b <= r && ok; b++) for (int a = -r; a <= r; a++) {
        if (std::max(std::abs(a), std::max(std::abs(b), std::abs(c))) != r) continue;
        if (term[lin(x + a, y + b, z + c)] != l) { ok = false; break; }
      }
      if (!ok) break; R = r;
    }
    ur[lin(x, y, z)] = R;
  }
  // cube certificates from the canonical labels: cu[c] = label if all points of [2c,2c+2]^3 agree else -1; cc[c] = cu if 27 cubes agree
  int c1 = (n[0] + 1) / 2, c2 = (n[1] + 1) / 2, c3 = (n[2] + 1) / 2;
  std::vector<int> cu((long)c1 * c2 * c3), cc((long)c1 * c2 * c3);
#pragma omp parallel for
  for (int cz = 0; cz < c3; cz++) for (int cy = 0; cy < c2; cy++) for (int cx = 0; cx < c1; cx++) {
    int l = term[lin(2 * cx, 2 * cy, 2 * cz)]; bool ok = true;
    for (int c = 0; c <= 2 && ok; c++) for (int b = 0; b <= 2 && ok; b++) for (int a = 0; a <= 2; a++) if (term[lin(2 * cx + a, 2 * cy + b, 2 * cz + c)] != l) { ok = false; break; }
    cu[cx + (long)c1 * (cy + (long)c2 * cz)] = ok ? l : -1;
  }
#pragma omp parallel for
  for (int cz = 0; cz < c3; cz++) for (int cy = 0; cy < c2; cy++) for (int cx = 0; cx < c1; cx++) {
    int l = cu[cx + (long)c1 * (cy + (long)c2 * cz)]; bool ok = l >= 0;
    for (int c = -1; c <= 1 && ok; c++) for (int b = -1; b <= 1 && ok; b++) for (int a = -1; a <= 1; a++) if (cu[wrap0(cx + a, c1) + (long)c1 * (wrap0(cy + b, c2) + (long)c2 * wrap0(cz + c, c3))] != l) { ok = false; break; }
    cc[cx + (long)c1 * (cy + (long)c2 * cz)] = ok ? l : -1;
  }
  long nedge = 0; for (long i = 0; i < nn; i++) nedge += ur[i] == 0;
  printf("points %ld edge %ld (%.2f%%)\n", nn, nedge, 100.0 * nedge / nn);
  // rules: 0: uni1, 1: uni2, 2: uni3, 3: uni1 & |dr|<0.25 , 4: uni1 && (uni2 || |dr|<.25), 5: none
  for (int rule = 0; rule < 7; rule++) {
    long steps = 0, bad = 0, walks = 0;
#pragma omp parallel reduction(+ : steps, bad, walks)
    {
      Bader b; for (int i = 0; i < 3; i++) b.n[i] = n[i]; b.f = f.data();
      std::memcpy(b.car2lat, meta, 72); std::memcpy(b.lat_i_dist, meta + 9, 27 * 8);
#pragma omp for schedule(dynamic, 4096)
      for (long s = 0; s < nn; s++) {
        if (ur[s] != 0) continue;
        walks++;
        int p[3]; b.unlin((int)s, p);
        double dr[3] = {0, 0, 0}; int res = -1;
        for (int it = 0; it < 100000; it++) {
          double g[3]; b.rho_grad_dir(p, g); int pm[3];
          const double gmax = std::max(std::fabs(g[0]), std::max(std::fabs(g[1]), std::fabs(g[2])));
          if (gmax < 1e-30) { dr[0] = dr[1] = dr[2] = 0; if (b.is_max(p)) { res = b.lin(p); break; } pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2]; b.step_ongrid(pm); }
          else { const double coeff = 1.0 / gmax; for (int i = 0; i < 3; i++) { g[i] = coeff * g[i]; const int ng = nint_(g[i]); pm[i] = p[i] + ng; dr[i] = dr[i] + g[i] - ng; const int nd = nint_(dr[i]); pm[i] += nd; dr[i] -= nd; } }
          b.pbc(pm); steps++;
          if (b.lin(pm) == b.lin(p)) { res = b.lin(p); break; }
          p[0] = pm[0]; p[1] = pm[1]; p[2] = pm[2];
          const int q = b.lin(p); const int R = ur[q];
          const double drm = std::max(std::fabs(dr[0]), std::max(std::fabs(dr[1]), std::fabs(dr[2])));
          bool stop = false;
          if (rule == 0) stop = R >= 1; else if (rule == 1) stop = R >= 2; else if (rule == 2) stop = R >= 3;
          else if (rule == 3) stop = R >= 1 && drm < 0.25; else if (rule == 4) stop = R >= 2 || (R >= 1 && drm < 0.25);
          if (rule == 6) stop = cc[(p[0] >> 1) + (long)c1 * ((p[1] >> 1) + (long)c2 * (p[2] >> 1))] >= 0;
          if (stop) { res = term[q]; break; }
        }
        if (res != term[s]) bad++;
      }
    }
    printf("rule %d: walks %ld steps %ld (%.2f per walk) wrong %ld\n", rule, walks, steps, (double)steps / walks, bad);
  }
}
