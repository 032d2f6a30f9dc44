// DMMA throughput vs resident warps per SM and independent accumulators per warp (B200, sm_100a).
// Answers: can 16 warps/SM (4 per SMSP, what a 128-register kernel allows) saturate the FP64 tensor pipe?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
double run(int warps_per_sm, int sms, double* out) {
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int threads = 32 * warps_per_sm;   // one CTA per SM
  if (threads > 1024) return 0;
  k<ILP><<<sms, threads>>>(out, 1000, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<ILP><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return 2.0 * 256.0 * sms * warps_per_sm * (double)iters * ILP / ms * 1e-9;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 64);
  int sms = p.multiProcessorCount;
  printf("warps/SM  ILP1   ILP2   ILP4   ILP8   ILP20  (TFLOP/s)\n");
  for (int w : {4, 8, 16, 32}) {
    printf("%7d  %6.2f %6.2f %6.2f %6.2f %6.2f\n", w, run<1>(w, sms, out), run<2>(w, sms, out), run<4>(w, sms, out),
           run<8>(w, sms, out), run<20>(w, sms, out));
  }
  return 0;
}
