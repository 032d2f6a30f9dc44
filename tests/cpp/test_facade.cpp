// Compile-and-run check of the C++ facade (include/SLAM.h) against libsrukf_b200.so.
// usage: test_facade <scenario.bin> [<out.bin>]
//   scenario.bin (tests/test_facade.py): L, steps, x0, S0 dense, u, z, then G (n x n, symmetric) and A (2n x n)
//   out.bin: [x (n) | S (n*n) | trace P | mchol(G) (n*n) | R of A (n*n) | sigma (Na*(2Na+1), Na = n) |
//             S after GSLCholeskyUpdate(u2, DOWNDATING, NEEDNOT_REORDER) (n*n)]  -- compared with the oracle in Python
// Prints the final robot pose and trace(P); exit code 0 on success, 3 when no CUDA device is present
// (the facade must throw, never fall back).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "SLAM.h"

static void put(FILE* f, const std::vector<double>& v) { fwrite(v.data(), 8, v.size(), f); }

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
  srukf::Mat64 G(n, n), A(2 * n, n);
  const bool helpers = fread(G.d.data(), 8, G.d.size(), fp) == G.d.size() && fread(A.d.data(), 8, A.d.size(), fp) == A.d.size();
  fclose(fp);
  try {
    srukf::CSLAM slam(L);
    slam.m_X_k.d = x0;
    slam.m_S_k.d = S0;
    slam.uploadState();
    for (int s = 0; s < steps; ++s) {
      for (int k = 0; k < 3; ++k) slam.Ut(k) = u[(size_t)s * 3 + k];
      for (int i = 0; i < L; ++i) {   // candidates for the gate of dataAssociation
        slam.map[i].matchLocation.x = z[((size_t)s * L + i) * 2];
        slam.map[i].matchLocation.y = z[((size_t)s * L + i) * 2 + 1];
      }
      if (s % 2 == 0) {
        slam.SLAM();   // the reference's no-argument frame: gate decides isMatching
      } else {
        slam.SLAM([&](srukf::CSLAM& me) {
          for (int i = 0; i < L; ++i) me.map[i].isMatching = me.visibleNow(i);
        });
      }
    }
    slam.updateCovariance();
    double tr = 0;
    for (int i = 0; i < n; ++i) tr += slam.m_P_k(i, i);
    printf("%.17g %.17g %.17g %.17g %.17g %d %d\n", slam.m_X_k(n - 4), slam.m_X_k(n - 3), slam.m_X_k(n - 2),
           slam.m_X_k(n - 1), tr, slam.m_nMatches, slam.m_nPredicts);
    if (argc > 2 && helpers) {
      FILE* fo = fopen(argv[2], "wb");
      if (!fo) return 2;
      put(fo, slam.m_X_k.d);
      put(fo, slam.m_S_k.d);
      fwrite(&tr, 8, 1, fo);
      srukf::Mat64 sr, R, sigma;
      slam.modifiedCholeskyDecomposition(sr, G);
      put(fo, sr.d);
      slam.GSLQrDecomposition(R, A);
      put(fo, R.d);
      srukf::Mat64 mu(n, 1);
      mu.d = slam.m_X_k.d;
      slam.calculateSampleParameter(n);
      slam.generateSigmaPoints(sigma, mu, slam.m_S_k);
      put(fo, sigma.d);
      srukf::Mat64 u2(n, 2);   // two small downdate columns built from rows of S
      for (int r = 0; r < n; ++r) { u2(r, 0) = 0.25 * slam.m_S_k(0, r); u2(r, 1) = 0.25 * slam.m_S_k(n - 4, r); }
      slam.GSLCholeskyUpdate(u2, slam.FLAG_4_DOWNDATING, slam.FLAG_4_NEEDNOT_REORDER);
      put(fo, slam.m_S_k.d);
      fclose(fo);
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "facade error: %s\n", e.what());
    return 3;
  }
  return 0;
}
