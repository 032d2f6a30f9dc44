// Compile-and-run check of the C++ facade (include/SLAM.h) against libsrukf_b200.so.
// usage: test_facade <scenario.bin>   (written by tests/test_facade.py: L, steps, x0, S0 dense, u, z)
// Prints the final robot pose and trace(P); exit code 0 on success, 3 when no CUDA device is present
// (the facade must throw, never fall back).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "SLAM.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* fp = fopen(argv[1], "rb");
  if (!fp) return 2;
  int hdr[2];
  if (fread(hdr, sizeof(int), 2, fp) != 2) return 2;
  const int L = hdr[0], steps = hdr[1], n = 6 * L + 4;
  std::vector<double> x0(n), S0((size_t)n * n), u((size_t)steps * 3), z((size_t)steps * L * 2);
  if (fread(x0.data(), 8, x0.size(), fp) != x0.size() || fread(S0.data(), 8, S0.size(), fp) != S0.size() ||
      fread(u.data(), 8, u.size(), fp) != u.size() || fread(z.data(), 8, z.size(), fp) != z.size())
    return 2;
  fclose(fp);
  try {
    srukf::CSLAM slam(L);
    slam.m_X_k.d = x0;
    slam.m_S_k.d = S0;
    slam.uploadState();
    for (int s = 0; s < steps; ++s) {
      for (int k = 0; k < 3; ++k) slam.Ut(k) = u[(size_t)s * 3 + k];
      slam.SLAM([&](srukf::CSLAM& me) {
        for (int i = 0; i < L; ++i) {
          me.map[i].isMatching = me.map[i].isVisible;
          me.map[i].matchLocation.x = z[((size_t)s * L + i) * 2];
          me.map[i].matchLocation.y = z[((size_t)s * L + i) * 2 + 1];
        }
      });
    }
    slam.updateCovariance();
    double tr = 0;
    for (int i = 0; i < n; ++i) tr += slam.m_P_k(i, i);
    printf("%.17g %.17g %.17g %.17g %.17g %d %d\n", slam.m_X_k(n - 4), slam.m_X_k(n - 3), slam.m_X_k(n - 2),
           slam.m_X_k(n - 1), tr, slam.m_nMatches, slam.m_nPredicts);
  } catch (const std::exception& e) {
    fprintf(stderr, "facade error: %s\n", e.what());
    return 3;
  }
  return 0;
}
