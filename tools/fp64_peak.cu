// FP64 pipe micro-benchmarks for B200 (sm_100a): DFMA vector peak, DMMA (mma.sync.m8n8k4.f64)
// peak, both interleaved, and shared-memory LDS.64/LDS.128 throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Output: one JSON object on stdout (flops are 2 per FMA).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
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

// interleave NF DFMA per 1 DMMA
template <int ILP, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], acc[ILP * NF];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < ILP * NF; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      dmma(c0[i], c1[i], a, b);
#pragma unroll
      for (int j = 0; j < NF; ++j) acc[i * NF + j] = fma(acc[i * NF + j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < ILP * NF; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// shared-memory read throughput: each thread reads VEC doubles per access, conflict-free
template <int VEC>
__global__ void __launch_bounds__(256) k_lds(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int base = threadIdx.x * VEC;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int idx = (base + u * 256 * VEC + it) & 4095 & ~(VEC - 1);
      if (VEC == 1) { s0 += sm[idx]; }
      else { double2 v = *reinterpret_cast<double2*>(&sm[idx]); s0 += v.x; s1 += v.y; }
    }
  }
  double s = s0 + s1 + s2 + s3;
  if (s == 12345.678) out[0] = s;
}

// DFMA fed from shared memory with R x C register tile (rank-1 update pattern): measures how close
// a smem-fed outer-product loop gets to the DFMA peak.
template <int R, int C>
__global__ void __launch_bounds__(256) k_outer(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  double acc[R][C];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[r][c] = 0;
  int tr = (threadIdx.x >> 4) * R;   // 16 row-groups
  int tc = (threadIdx.x & 15) * C;   // 16 col-groups
  for (int it = 0; it < iters; ++it) {
    const double* pa = sm + ((it & 31) * 128);          // A panel row (k): 16*R <= 128 values
    const double* pb = sm + 4096 + ((it & 31) * 128);   // B panel row (k)
    double a[R], b[C];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = pa[tr + r];
#pragma unroll
    for (int c = 0; c < C; ++c) b[c] = pb[tc + c];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[r][c];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  const int iters = 20000;
  const int blocks = sms * 8;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);

  { // DFMA
    float ms = time_ms([&] { k_dfma<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * blocks * 256.0 * iters * 8;
    printf(", \"dfma_tflops\": %.3f", fl / ms * 1e-9);
  }
  { // DFMA sustained ~2 s
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 400; ++n) k_dfma<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * blocks * 256.0 * iters * 8 * n;
    printf(", \"dfma_tflops_sustained\": %.3f, \"dfma_sustained_ms\": %.1f", fl / ms * 1e-9, ms);
  }
  { // DMMA: m8n8k4 = 256 FMA per warp instr
    float ms = time_ms([&] { k_dmma<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * 256.0 * blocks * 8.0 /*warps*/ * iters * 8;
    printf(", \"dmma_tflops\": %.3f", fl / ms * 1e-9);
  }
  { // mix 1 DMMA : 8 DFMA (equal flops)
    float ms = time_ms([&] { k_mix<4, 8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = (2.0 * 256.0 * blocks * 8.0 * iters * 4) + (2.0 * blocks * 256.0 * iters * 32);
    printf(", \"mix_1dmma_8dfma_tflops\": %.3f", fl / ms * 1e-9);
  }
  { // mix 1 DMMA : 2 DFMA
    float ms = time_ms([&] { k_mix<4, 2><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = (2.0 * 256.0 * blocks * 8.0 * iters * 4) + (2.0 * blocks * 256.0 * iters * 8);
    printf(", \"mix_1dmma_2dfma_tflops\": %.3f", fl / ms * 1e-9);
  }
  { // LDS
    float ms1 = time_ms([&] { k_lds<1><<<blocks, 256, 32768>>>(out, 4000); }, 5);
    float ms2 = time_ms([&] { k_lds<2><<<blocks, 256, 32768>>>(out, 4000); }, 5);
    double b1 = 8.0 * blocks * 256.0 * 4000 * 8, b2 = 16.0 * blocks * 256.0 * 4000 * 8;
    printf(", \"lds64_TBps\": %.3f, \"lds128_TBps\": %.3f", b1 / ms1 * 1e-9, b2 / ms2 * 1e-9);
  }
  { // smem-fed outer product
    CK(cudaFuncSetAttribute(k_outer<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_outer<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_outer<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_outer<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    float m22 = time_ms([&] { k_outer<2, 2><<<sms * 3, 256, 65536>>>(out, 40000); }, 5);
    float m44 = time_ms([&] { k_outer<4, 4><<<sms * 3, 256, 65536>>>(out, 40000); }, 5);
    float m84 = time_ms([&] { k_outer<8, 4><<<sms * 3, 256, 65536>>>(out, 40000); }, 5);
    float m88 = time_ms([&] { k_outer<8, 8><<<sms * 3, 256, 65536>>>(out, 40000); }, 5);
    double base = 2.0 * sms * 3 * 256.0 * 40000;
    printf(", \"outer2x2_tflops\": %.3f, \"outer4x4_tflops\": %.3f, \"outer8x4_tflops\": %.3f, \"outer8x8_tflops\": %.3f",
           base * 4 / m22 * 1e-9, base * 16 / m44 * 1e-9, base * 32 / m84 * 1e-9, base * 64 / m88 * 1e-9);
  }
  printf("}\n");
  return 0;
}
