/*
 * harness.c -- headless Linux timing harness around the CPU oracle (TEST/BASELINE INFRASTRUCTURE).
 * Times the reference algorithm (literal mode: materialised sigma matrices, 100-iteration Newton,
 * one dense S^T S + GMW per U column; SLAM.cpp:1430-1775, 2020-2327) on host cores.
 * Input: a binary blob written by bench.py / tests (see oracle/oracle.py: write_harness_input).
 * usage: srukf_harness <input.bin> <nsteps> <nthreads> <downdate_mode>
 * Prints one JSON line: {"filters":B,"L":L,"steps":K,"threads":T,"seconds":s,"filter_steps_per_s":v}
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "srukf_oracle.h"

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s input.bin nsteps nthreads downdate_mode\n", argv[0]);
    return 2;
  }
  FILE *fp = fopen(argv[1], "rb");
  if (!fp) { perror("open"); return 2; }
  int hdr[4]; /* B, L, nsteps_available, reserved */
  if (fread(hdr, sizeof(int), 4, fp) != 4) return 2;
  int B = hdr[0], L = hdr[1], K_av = hdr[2];
  int n = 6 * L + 4;
  int K = atoi(argv[2]);
  if (K > K_av) K = K_av;
  int T = atoi(argv[3]);
  OracleParams p;
  oracle_default_params(&p);
  p.downdate_mode = atoi(argv[4]);
  size_t nx = (size_t)B * n, nS = (size_t)B * n * n, nu = (size_t)K_av * B * 3, nz = (size_t)K_av * B * L * 2,
         nm = (size_t)K_av * B * L;
  double *x = malloc(nx * 8), *S = malloc(nS * 8), *u = malloc(nu * 8), *z = malloc(nz * 8);
  unsigned char *m = malloc(nm);
  if (fread(x, 8, nx, fp) != nx || fread(S, 8, nS, fp) != nS || fread(u, 8, nu, fp) != nu ||
      fread(z, 8, nz, fp) != nz || fread(m, 1, nm, fp) != nm) {
    fprintf(stderr, "short read\n");
    return 2;
  }
  fclose(fp);
  double t0 = now_s();
  oracle_batch_step(B, L, &p, x, S, u, z, m, 0, K, T, NULL);
  double t1 = now_s();
  double cs = 0;
  for (size_t i = 0; i < nx; i++) cs += x[i];
  printf("{\"filters\": %d, \"L\": %d, \"steps\": %d, \"threads\": %d, \"seconds\": %.6f, "
         "\"filter_steps_per_s\": %.6f, \"downdate_mode\": %d, \"checksum_x\": %.17g}\n",
         B, L, K, T, t1 - t0, (double)B * K / (t1 - t0), p.downdate_mode, cs);
  free(x); free(S); free(u); free(z); free(m);
  return 0;
}
