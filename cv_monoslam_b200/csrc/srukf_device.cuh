// srukf_device.cuh -- device-side building blocks of the batched SRUKF (sm_100a, FP64).
//
// Everything here follows the arithmetic of MonoSLAM/SLAM.cpp (cited per function) but never
// materialises the reference's Na x (2Na+1) sigma matrix: sigma point i is x +- gamma * row_i of
// blockdiag(S, Mt, Qt) (SLAM.cpp:1148-1162,1461-1463) and is generated on the fly from packed S.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/srukf.h"

namespace srukf {

// Per-handle constants, passed by value to every kernel.
struct DevParams {
  int B, L, n, nf, Na, P, ntri;
  int np;    // n rounded up to a multiple of 8 (internal padded state dimension)
  int Lc;    // 2L rounded up to a multiple of 8 (padded measurement dimension)
  int nbp;   // doubles per filter in the internal layout of S (np * np)
  // camera (SLAM.cpp:329-337)
  double cam_dx, cam_dy, cam_cx, cam_cy, cam_k1, cam_k2, f1, f2;
  double inv_dx, inv_dy;  // 1/cam_dx, 1/cam_dy
  double img_w, img_h;
  // noise (SLAM.cpp:195-198, 238)
  double a1, a2, a3, a4, sigma_measure;
  // sample parameters for Na (SLAM.cpp:1050-1103)
  double gamma, wm0, wc0, wi, wi_sr;
  double Wsum;   // wm0 + 2 Na wi  (== 1 analytically)
  double cpair;  // sqrt(2) * wi_sr * gamma (== 1 analytically for all three weight types)
  double epsilon;
  int newton_iters;
  // distortion fast paths, decided on the host from the camera parameters (fill_dev_params):
  int dist_series;   // k2 == 0 and |k1| ru^2 <= 2.5e-4 everywhere in the image: closed-form series root (no iteration)
  int dist_inward;   // k1 >= 0 and k2 >= 0: distortion moves a pixel towards the principal point, so a pixel that passed
                     // the [10, W-10] x [10, H-10] test (or was zeroed by it) cannot leave the image: second test skipped
  int pred_free;     // k_predict: per-warp partial sums of all features fit in shared memory (no block barriers)
  int force_fb_ppm;  // measurement only (SRUKF_FORCE_FALLBACK_PPM): parts per million of the filters forced through the fallback
  int dbg_skip_mma;  // diagnostics only (SRUKF_DBG_SKIP_MMA): bit 0 stream the K chunks but skip the DMMAs, bit 1 k_gain without loads
};

// packed upper-triangular row-major: row i holds columns i..n-1
__host__ __device__ __forceinline__ int tri_off(int i, int n) { return i * n - (i * (i - 1)) / 2; }

// Internal layout of S in HBM: one np x np row-major square per filter (np = n rounded up to a multiple of 8),
// upper triangular, everything below the diagonal stays zero, rows/columns n..np-1 are padding (identity).
// The square costs 2x the packed size in capacity but no extra traffic (kernels only stream the row suffixes
// they need), and it makes every K-chunk of the DMMA kernels a regular strided box that ONE TMA tensor copy
// can fetch (the packed layouts needed one copy per row and were copy-count bound).
__host__ __device__ __forceinline__ int bp_idx(int k, int c, int np) { return k * np + c; }

__device__ __forceinline__ double S_at(const double* __restrict__ S, int np, int i, int c) {
  return (c >= i) ? S[bp_idx(i, c, np)] : 0.0;
}

// FP64 tensor-core MMA (DMMA): D(8x8) += A(8x4, row) * B(4x8, col).
//   a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], c0/c1 = C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) -------------------------------------------
// K-chunks are streamed HBM/L2 -> shared memory by one elected thread with 1-D bulk copies that complete on
// an mbarrier ("full"); consumer warps release a stage through a second mbarrier ("empty").
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      #ifdef SRUKF_TEST_WAIT
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 3-D tiled TMA load (cp.async.bulk.tensor, SASS UTMALDG): box of the tensor map at (c0, c1, c2) -> smem
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of a box (no shared-memory destination): hides the DRAM latency of chunks further ahead than the ring
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// orders this thread's earlier generic-proxy accesses (shared AND global) before later async-proxy (TMA) accesses
// same with an L2 eviction-priority hint (createpolicy): evict_last for operands that are re-streamed several times,
// evict_first for data read once
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, "
      "%5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// 16-byte asynchronous global->shared copy (LDGSTS, L2 only) and its completion hook onto an mbarrier:
// the executing thread arrives on `bar` once all of its earlier cp.async operations have landed.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 1/d for |d| well inside the normal range: MUFU.RCP64H seed + two Newton steps (5 dependent FP64 ops instead of
// the ~15 of an IEEE division), accurate to ~1 ulp.  Used on the pivot chain of the panel factorisation.
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

// 1/d = x0 (1 + e2) with x0 = MUFU.RCP64H(d) (20-bit input, ~2^-20 relative error), e = 1 - d x0, e2 = e + e^2: relative
// error e^3 < 2^-57.  Used on the pivot chains: 1/d is never formed there -- a product p = u x0 is started as soon as x0
// is known and corrected with ONE fused multiply-add, fma(-p, e2, r - p), so that only three FP64 operations (e, e2, the
// correction) depend on each other per pivot instead of seven.
__device__ __forceinline__ void rcp_parts(double d, double& x0, double& e2) {
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
  const double e = fma(-d, x0, 1.0);
  e2 = fma(e, e, e);
}

// ---------------------------------------------------------------------------------------------
// distortOnePointRW, SLAM.cpp:3177-3213.
// The reference solves rd + k1 rd^3 + k2 rd^5 = ru by 100 Newton steps and returns c + (xu/d)/dx with
// d = 1 + k1 rd^2 + k2 rd^4.  The same root is found here for t = rd/ru = 1/d directly:
//     t (1 + a t^2 + b t^4) = 1,   a = k1 ru^2,  b = k2 ru^4
// which needs no square root and no division (Newton with the derivative's reciprocal replaced by its
// first-order expansion 2 - f'; it converges to the same fixed point, to rounding).  The loop stops when the
// step is below one ulp (the reference's fixed 100 iterations sit on that fixed point).  |result - reference|
// is a few ulp of the pixel coordinate.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void distort_point(const DevParams& p, double ux, double uy, double& ox, double& oy,
                                              uint32_t& flags) {
  const double xu = (ux - p.cam_cx) * p.cam_dx;
  const double yu = (uy - p.cam_cy) * p.cam_dy;
  const double ru2 = xu * xu + yu * yu;
  if (p.dist_series) {
    // t + a t^3 = 1 with 0 <= |a| <= 2.5e-4:  t = sum_k binom(3k,k)/(2k+1) (-a)^k = 1 - a + 3a^2 - 12a^3 + 55a^4 - 273a^5 + ..
    // (the first omitted term, 1428 a^6 < 4e-19, is below half an ulp of t).  The reference's 100 Newton steps sit on
    // the same root to rounding.
    const double a = p.cam_k1 * ru2;
    double t = fma(a, -273.0, 55.0);
    t = fma(a, t, -12.0);
    t = fma(a, t, 3.0);
    t = fma(a, t, -1.0);
    t = fma(a, t, 1.0);
    ox = p.cam_cx + (xu * t) * p.inv_dx;
    oy = p.cam_cy + (yu * t) * p.inv_dy;
    return;   // dist_series implies dist_inward or a negligible outward shift checked on the host
  }
  const double a = p.cam_k1 * ru2, bq = p.cam_k2 * (ru2 * ru2);
  // second-order estimate of the root of t + a t^3 + bq t^5 = 1 (t = 1 - d, d (1 + 3a + 5bq) = a + bq + O(d^2))
  const double s1 = a + bq, s3 = 3.0 * a + 5.0 * bq;
  double t = fma(s1, s3, 1.0 - s1);
  // The step uses 2 - f' for 1/f', so the error contracts by (f' - 1)^2 ~ s3^2 per step plus the Newton term
  // (3|a| + 10|bq|) e: once |step| * (s3^2 + K |step|) is below half an ulp of t the NEW iterate is converged and the
  // confirming step can be skipped (the reference runs 100 steps to the same fixed point).
  const double c1 = s3 * s3, K = 3.0 * fabs(a) + 10.0 * fabs(bq);
  for (int it = 0; it < p.newton_iters; ++it) {
    const double t2 = t * t;
    const double g = a * t2 + bq * (t2 * t2);
    const double f = fma(t, g, t - 1.0);
    const double ff = 1.0 + 3.0 * a * t2 + 5.0 * bq * (t2 * t2);
    const double nt = fma(-f, 2.0 - ff, t);
    const double dl = fabs(nt - t);
    t = nt;
    if (dl * fma(K, dl, c1) <= 2.0e-17 || dl <= 1.2e-16) break;
  }
  ox = p.cam_cx + (xu * t) * p.inv_dx;
  oy = p.cam_cy + (yu * t) * p.inv_dy;
  if (p.dist_inward) return;
  bool vis = (ox >= 0) && (ox <= p.img_w) && (oy >= 0) && (oy <= p.img_h);
  if (!vis) {
    ox = 0;
    oy = 0;
    flags |= SRUKF_FLAG_OUT_OF_VIEW;
  }
}

// ---------------------------------------------------------------------------------------------
// World-frame ray (hx,hy,hz) = feature point - robot position  ->  distorted pixel:
//   World2Camera :3289-3292 with Rcw = Rwc.inv() (:1642-1643, OpenCV closed-form 3x3 inverse = adj/det; the
//   three distinct entries cd = c/det, sd = s/det, zd = det/det are precomputed per sigma point),
//   Camera2Image :3324-3347 (x/y swap kept), distortion.  (e0, e1) = pixel-noise components of the sigma point.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pixel_from_ray(const DevParams& p, double hx, double hy, double hz, double cd,
                                               double sd, double zd, double e0, double e1, double& ox, double& oy,
                                               uint32_t& flags) {
  const double rxx = cd * hx + sd * hy;
  const double ryy = -sd * hx + cd * hy;
  const double rzz = zd * hz;
  double ux, uy;
  if (rzz == 0) {
    ux = 0;
    uy = 0;
  } else {
    const double r = fast_rcp(rzz);   // rzz ~ ceiling height: well inside the normal range
    uy = p.cam_cx + (p.f1 * rxx) * r + e0;
    ux = p.cam_cy + (p.f2 * ryy) * r + e1;
    if (ux < 10 || ux > p.img_w - 10 || uy < 10 || uy > p.img_h - 10) {
      ux = 0;
      uy = 0;
      flags |= SRUKF_FLAG_OUT_OF_VIEW;
    }
  }
  distort_point(p, ux, uy, ox, oy, flags);
}

// sin/cos of (a0 + d) and (a0 - d) from sin/cos(a0): angle addition with a degree-11/12 Taylor series of the
// small perturbation d = gamma * S(k, col) (the +- sigma points share it); full sincos for |d| >= 1/8.
__device__ __forceinline__ void sincos_pm(double a0, double s0, double c0, double d, double& sp, double& cp, double& sm,
                                          double& cm) {
  if (fabs(d) < 0.125) {
    const double d2 = d * d;
    double sd = fma(d2, -1.0 / 110.0, 1.0);
    sd = fma(d2 * (-1.0 / 72.0), sd, 1.0);
    sd = fma(d2 * (-1.0 / 42.0), sd, 1.0);
    sd = fma(d2 * (-1.0 / 20.0), sd, 1.0);
    sd = fma(d2 * (-1.0 / 6.0), sd, 1.0) * d;
    double cd = fma(d2, -1.0 / 132.0, 1.0);
    cd = fma(d2 * (-1.0 / 90.0), cd, 1.0);
    cd = fma(d2 * (-1.0 / 56.0), cd, 1.0);
    cd = fma(d2 * (-1.0 / 30.0), cd, 1.0);
    cd = fma(d2 * (-1.0 / 12.0), cd, 1.0);
    cd = fma(d2 * (-0.5), cd, 1.0);
    sp = fma(s0, cd, c0 * sd);
    cp = fma(c0, cd, -s0 * sd);
    sm = fma(s0, cd, -c0 * sd);
    cm = fma(c0, cd, s0 * sd);
  } else {
    sincos(a0 + d, &sp, &cp);
    sincos(a0 - d, &sm, &cm);
  }
}

// deterministic block reductions (fixed tree order => run-to-run and 1-vs-N-GPU bit identical)
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
  int tid = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (tid < 32) {
    r = (tid < NT / 32) ? red[tid] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    if (tid == 0) red[32] = r;
  }
  __syncthreads();
  return red[32];
}

// N sums at once behind ONE pair of barriers (block_sum costs three per value): warp shuffle trees, then every
// thread adds the NT/32 warp partials in fixed order.  `scratch` holds N * (NT/32) doubles.
template <int NT, int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double* scratch) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int c = 0; c < N; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_down_sync(0xffffffffu, v[c], o);
  }
  __syncthreads();
  if ((tid & 31) == 0) {
#pragma unroll
    for (int c = 0; c < N; ++c) scratch[c * (NT / 32) + (tid >> 5)] = v[c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < N; ++c) {
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) r += scratch[c * (NT / 32) + w];
    v[c] = r;
  }
}

template <int NT>
__device__ __forceinline__ double block_max(double v, double* red) {
  int tid = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    double r = (tid < NT / 32) ? red[tid] : -1.0e300;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(0xffffffffu, r, o));
    if (tid == 0) red[32] = r;
  }
  __syncthreads();
  return red[32];
}

}  // namespace srukf
