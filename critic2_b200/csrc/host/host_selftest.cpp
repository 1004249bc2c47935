// host_selftest.cpp -- drives the C++ host mirror the way critic2's intgrid_driver does
// (BADER, then INTEGRABLE rho; then YT is left to the Python tests).  Reads a raw field written by the
// test (n1 n2 n3, x2c, atoms, f), prints nattr, a label checksum and the basin populations as JSON.
// usage: host_selftest <input.bin>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "critic2_host.hpp"

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: host_selftest input.bin\n"); return 2; }
  FILE* fp = std::fopen(argv[1], "rb");
  if (!fp) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  int hdr[4];
  double x2c[9];
  if (std::fread(hdr, sizeof(int), 4, fp) != 4 || std::fread(x2c, sizeof(double), 9, fp) != 9) return 2;
  c2h::system s;
  s.set_cell(x2c);
  s.xat.resize(3 * (size_t)hdr[3]);
  if (std::fread(s.xat.data(), sizeof(double), s.xat.size(), fp) != s.xat.size()) return 2;
  c2h::basindat bas;
  for (int i = 0; i < 3; i++) bas.n[i] = s.grid.n[i] = hdr[i];
  bas.f.resize((size_t)hdr[0] * hdr[1] * hdr[2]);
  if (std::fread(bas.f.data(), sizeof(double), bas.f.size(), fp) != bas.f.size()) return 2;
  std::fclose(fp);
  try {
    c2h::gpu_init(0);
    c2h::bader_integrate(s, bas);
    std::vector<c2h::int_result> res;
    std::vector<double> vol;
    c2h::intgrid_fields(s, bas, {bas.f.data()}, res, vol);
    uint64_t hsh = 1469598103934665603ull;  // FNV-1a over the labels
    for (int v : bas.idg) { hsh ^= (uint32_t)v; hsh *= 1099511628211ull; }
    std::printf("{\"nattr\": %d, \"labels_fnv\": \"%llx\", \"pop\": [", bas.nattr, (unsigned long long)hsh);
    for (int i = 0; i < bas.nattr; i++) std::printf("%s%.17g", i ? ", " : "", res[0].psum[i]);
    std::printf("], \"vol\": [");
    for (int i = 0; i < bas.nattr; i++) std::printf("%s%.17g", i ? ", " : "", vol[i]);
    // INTEGRABLE 1 MULTIPOLES 2
    std::vector<double> mpole;
    c2h::intgrid_multipoles(s, bas, bas.f.data(), 2, {}, mpole);
    std::printf("], \"mpole_lmax2\": [");
    for (size_t i = 0; i < mpole.size(); i++) std::printf("%s%.17g", i ? ", " : "", mpole[i]);
    std::printf("]}\n");
    // error behaviour: a bad call must raise like ferror(...,faterr)
    bool raised = false;
    try { c2h::basindat bad = bas; bad.f.resize(5); c2h::bader_integrate(s, bad); } catch (const c2h::fatal_error&) { raised = true; }
    if (!raised) { std::fprintf(stderr, "expected a fatal error\n"); return 3; }
    c2h::gpu_end();
  } catch (const c2h::fatal_error& e) {
    return 1;
  }
  return 0;
}
