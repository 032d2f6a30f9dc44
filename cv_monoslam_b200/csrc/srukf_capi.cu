// srukf_capi.cu -- C ABI (include/srukf.h) over the CUDA kernels.  No CPU fallback: every entry
// point needs a CUDA device and fails with SRUKF_ENODEV / SRUKF_ECUDA otherwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "srukf_device.cuh"

namespace srukf {
struct StepPtrs {
  double* x; double* S; const double* u; const double* z; const uint8_t* matched;
  double* hbar; double* si; uint8_t* visible; double* cshift; double* pxyr; double* rsig;
  double* dZ; double* U; double* G; uint32_t* flags; int chunk0;
  double* S2; int* worklist; int rel0; unsigned long long* dbg;
  const CUtensorMap* tmaps; int sbuf; int tm_dz; int tm_ut; int dz_filter0;
  double* Pd; double* Pd2; int carry_p;
  double* G2; int n_new;
  double* Gp;
  double* Ed;
  double* Useq;
  int* nact;
};
// tensor-map table layout and box geometry (must match srukf_kernels.cu)
constexpr int TM_S0 = 0, TM_S1 = 8, TM_UT = 16, TM_DZ = 24, TM_DZ_ALL = 25, TM_USEQ = 26, TM_UT2 = 34, TM_COUNT = 42, TM_ROWSETS = 8;
#ifndef SRUKF_TW
#define SRUKF_TW 64
#endif
#ifndef SRUKF_PAD
#define SRUKF_PAD 4
#endif
constexpr int TP = SRUKF_TW + SRUKF_PAD;
int gain_dz_box(const DevParams& p);
int gain_variant(const DevParams& p);
int tile_warps(const DevParams& p);
int gain_row_ctas(const DevParams& p);
bool predict_free(const DevParams& p);
cudaError_t configure_kernels(const DevParams& p);
size_t predict_smem_bytes(const DevParams& p);
size_t gain_smem_bytes(const DevParams& p);
size_t update_smem_bytes(const DevParams& p);
void launch_predict(const DevParams& p, const StepPtrs& q, int nblocks, bool motion, bool meas, bool save_rsig,
                    cudaStream_t st);
void launch_gain(const DevParams& p, const StepPtrs& q, int nblocks, cudaStream_t st);
void launch_update(const DevParams& p, const StepPtrs& q, int nblocks, cudaStream_t st);
void launch_downdate(const DevParams& p, const StepPtrs& q, int nblocks, int mode, int use_worklist, cudaStream_t st);
void launch_update_seq(const DevParams& p, const StepPtrs& q, int nblocks, cudaStream_t st);
bool update_seq_available(const DevParams& p);
void launch_import(const DevParams& p, int nb, int fmt, const double* ext, double* bp, cudaStream_t st);
void launch_export(const DevParams& p, int nb, int fmt, const double* bp, double* ext, cudaStream_t st);
void launch_form_P(const DevParams& p, int b0, int nb, double* S, double* Pd, cudaStream_t st);
void launch_cov_block(const DevParams& p, const double* S, int r0, int nr, double* out, cudaStream_t st);
void launch_init_features(const DevParams& p, int nblocks, const double* x4, const double* S4, const double* kp,
                          double rho0, double sigma_rho, double gamma, double wi, double* x, double* S, double* Pd,
                          double* G, uint32_t* flags, cudaStream_t st);
void launch_add_features(const DevParams& p, int nblocks, const double* kp, double rho0, double sigma_rho, double gamma,
                         double wi, const double* xs, const double* Ss, int ns, int nps, int M, double* x, double* S,
                         double* Pd, double* G, double* Ascr, const uint32_t* flags_src, uint32_t* flags, cudaStream_t st);
void launch_delete_feature(const DevParams& p, int nblocks, const double* xs, const double* Ss, int ns, int nps,
                           const int* ids, double* x, double* S, double* Pd, double* G, double* V,
                           const uint32_t* flags_src, uint32_t* flags, cudaStream_t st);
void launch_gate(const DevParams& p, const double* z, const double* hbar, const double* si, const uint8_t* visible,
                 double threshold, uint8_t* accept, double* d2, cudaStream_t st);
void launch_stats(const DevParams& p, const double* x, const double* S, const double* truth, double* perf,
                  const uint32_t* flags, double* out, cudaStream_t st);
size_t downdate_smem_bytes(const DevParams& p);
void launch_mchol_batch(int nb, int n, double eps, const double* Gd, double* Gp, double* Sd, uint32_t* flags, int grid,
                        cudaStream_t st);
void launch_qr_batch(int nb, int m, int n, const double* A, double* W, double* R, int grid, cudaStream_t st);
void launch_sigma_points(int nb, int Na, double gamma, const double* mu, const double* sr, double* sigma, cudaStream_t st);
void launch_chol_update(const DevParams& p, int grid, double* S, double* Pd, const double* Ut, int k, double sign, int M,
                        double* G, double* G2, uint32_t* flags, cudaStream_t st);
cudaError_t measure_fp64_peak(int sms, int iters, int reps, double* tflops, cudaStream_t st);
}  // namespace srukf

using namespace srukf;

static thread_local std::string g_err;
static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  g_err = what;
  if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
  return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? SRUKF_ENOMEM : SRUKF_ECUDA, #x, e_); } while (0)

struct srukf_handle {
  int device = 0;
  DevParams p{};
  SrukfParams prm{};
  cudaStream_t stream = nullptr;
  // state: S lives in the internal square layout; k_update ping-pongs between the two buffers
  // state: ONE np x np square per filter in the internal layout (factor in the upper triangle, carried covariance in
  // the lower one); k_update works in place (a panel's P_old is read before its positions are rewritten)
  double *x = nullptr, *S = nullptr;
  double* Pd = nullptr;                               // diagonal of the carried covariance (fused mode)
  double* Ed = nullptr;                               // E_j = d_j - c_jj of the last update: lets the fallback rebuild P_old
  CUtensorMap* tmaps = nullptr;                       // device table of TMA tensor maps
  // per-step inputs (device copies for the host-pointer API)
  double *u = nullptr, *z = nullptr; uint8_t* matched = nullptr;
  // srukf_step (host pointers): the inputs of step s+1 are copied on copy_stream into the other input set while step s
  // computes; srukf_get_x_async reads m_X_k back through a staging copy on d2h_stream while the next step computes
  double *u_in[2] = {nullptr, nullptr}, *z_in[2] = {nullptr, nullptr}; uint8_t* m_in[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_xs = nullptr, ev_xd = nullptr;
  int in_slot = 0;
  bool inputs_dirty = true;   // an entry point other than srukf_step has queued work that reads input set 0
  double* x_stage = nullptr;
  // prediction outputs
  double *hbar = nullptr, *si = nullptr, *cshift = nullptr, *pxyr = nullptr; uint8_t* visible = nullptr;
  uint32_t* flags = nullptr;
  // scratch
  int chunk = 0;           // filters per pipeline pass of srukf_step
  double *dZ = nullptr, *U = nullptr, *G = nullptr;
  double* G2 = nullptr;    // scratch of the NEED_REORDER update (allocated on first use)
  double* Gp = nullptr;    // carried covariance of the reference-order fallback, one packed matrix per fallback CTA
  double* Useq = nullptr;  // per fallback CTA a square [np][np]: U rows of a bisection pass / work area of the literal step
  // The guard's fallback of chunk i runs on fb_stream while the main stream goes on with chunk i+1; for that the chunk
  // scratch the fallback reads (U, nact, work list) exists twice and the chunks alternate between the two sets.
  double* U_set[2] = {nullptr, nullptr}; int* nact_set[2] = {nullptr, nullptr}; int* wl_set[2] = {nullptr, nullptr};
  cudaStream_t fb_stream = nullptr;
  cudaEvent_t ev_upd[2] = {nullptr, nullptr}, ev_fb[2] = {nullptr, nullptr};
  struct { int b0, nb; bool pending; } fbp[2] = {{0, 0, false}, {0, 0, false}};
  int uset = 0;
  int gslots = 0;          // CTAs (and G scratch slots) of the reference-order fallback kernel
  int* worklist = nullptr;
  int* nact = nullptr;     // [chunk] per-filter count of features used by k_gain
  unsigned long long* dbg = nullptr;  // phase-cycle counters (SRUKF_PHASE_TIMING=1)
  // split-API persistent intermediates (allocated on first use)
  double *rsig = nullptr, *dZ_all = nullptr, *U_all = nullptr, *G_all = nullptr;
  int phase = 0;           // 0 idle, 1 motion done, 2 measurement done
  double *perf = nullptr, *stats_out = nullptr, *truth = nullptr;
  uint64_t launches = 0;
  // optional per-kernel event timing
  bool profiling = false;
  std::vector<cudaEvent_t> ev;   // triples of (start, stop, kind) are stored as pairs + kinds
  std::vector<int> ev_kind;
  size_t ev_used = 0;
  double prof_ms[3] = {0, 0, 0};
  uint64_t prof_n[3] = {0, 0, 0};
};

static void prof_begin(srukf_handle* h, int kind) {
  if (!h->profiling) return;
  if (h->ev_used + 2 > h->ev.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    h->ev.push_back(a); h->ev.push_back(b);
  }
  h->ev_kind.push_back(kind);
  cudaEventRecord(h->ev[h->ev_used], h->stream);
}
static void prof_end(srukf_handle* h) {
  if (!h->profiling) return;
  cudaEventRecord(h->ev[h->ev_used + 1], h->stream);
  h->ev_used += 2;
}
static void prof_collect(srukf_handle* h) {
  for (size_t i = 0; i + 1 < h->ev_used + 1 && i / 2 < h->ev_kind.size(); i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) {
      h->prof_ms[h->ev_kind[i / 2]] += ms;
      h->prof_n[h->ev_kind[i / 2]]++;
    }
  }
  h->ev_used = 0;
  h->ev_kind.clear();
}

// Outstanding side-stream fallbacks: the main stream waits for the ones that touch filters [b0, b0+nb) (b0 < 0: all)
static void fb_join(srukf_handle* h, int b0 = -1, int nb = 0) {
  if (!h->fb_stream) return;
  for (int s = 0; s < 2; ++s) {
    if (!h->fbp[s].pending) continue;
    const bool overlap = b0 < 0 || (h->fbp[s].b0 < b0 + nb && b0 < h->fbp[s].b0 + h->fbp[s].nb);
    if (overlap) { cudaStreamWaitEvent(h->stream, h->ev_fb[s], 0); h->fbp[s].pending = false; }
  }
}

extern "C" {

void srukf_default_params(SrukfParams* p) {
  // CSLAM::initializeParameters, SLAM.cpp:158-343
  p->cam_dx = 0.0028; p->cam_dy = 0.0028; p->cam_cx = 310.1129; p->cam_cy = 236.7526;
  p->cam_k1 = 0.0001; p->cam_k2 = 0.0; p->cam_f = 2.1735;
  p->image_width = 640; p->image_height = 480;
  p->a1 = 8; p->a2 = 8; p->a3 = 8; p->a4 = 8;
  p->sigma_measure = 3.0;
  p->weight_type = 0; p->alpha = 1e-3; p->beta = 2;
  p->epsilon = 1e-13;
  p->newton_iters = 100;
  p->downdate_mode = 0;
}

const char* srukf_last_error(void) { return g_err.c_str(); }
const char* srukf_version(void) { return "srukf-b200 0.1 (sm_100a, fp64)"; }

// calculateSampleParameter, SLAM.cpp:1050-1103 (operation order kept)
static void sample_weights(const SrukfParams& s, int Na, double& wm0, double& wc0, double& wi, double& wi_sr,
                           double& gamma) {
  if (s.weight_type == 0) {
    wm0 = 1.0 - Na / 3.0; wc0 = 1.0 - Na / 3.0;
    wi = (1.0 - wc0) / (2 * Na); wi_sr = std::sqrt(wi);
    gamma = std::sqrt(Na / (1.0 - wm0));
  } else if (s.weight_type == 1) {
    double kappa = 0;
    double lambda = std::pow(s.alpha, 2) * (Na + kappa) - Na;
    gamma = std::sqrt(Na + lambda);
    wm0 = lambda / (Na + lambda);
    wc0 = wm0 + (1 - std::pow(s.alpha, 2) + s.beta);
    wi = 1.0 / (2 * (Na + lambda)); wi_sr = std::sqrt(std::fabs(wi));
  } else {
    gamma = std::sqrt(3.0 * Na / 2.0);
    wm0 = 1.0 / 3.0; wc0 = 1.0 / 3.0;
    wi = 1.0 / (3.0 * Na); wi_sr = std::sqrt(wi);
  }
}

static void fill_dev_params(DevParams& d, const SrukfParams& s, int B, int L) {
  d.B = B; d.L = L; d.n = 6 * L + 4; d.nf = 6 * L; d.Na = d.n + 5; d.P = 2 * d.Na + 1;
  d.ntri = d.n * (d.n + 1) / 2;
  d.np = (d.n + 7) & ~7;
  d.Lc = (2 * L + 7) & ~7;
  d.nbp = d.np * d.np;
  d.cam_dx = s.cam_dx; d.cam_dy = s.cam_dy; d.cam_cx = s.cam_cx; d.cam_cy = s.cam_cy;
  d.cam_k1 = s.cam_k1; d.cam_k2 = s.cam_k2;
  d.f1 = s.cam_f / s.cam_dx; d.f2 = s.cam_f / s.cam_dy;  // SLAM.cpp:336-337
  d.inv_dx = 1.0 / s.cam_dx; d.inv_dy = 1.0 / s.cam_dy;
  d.img_w = s.image_width; d.img_h = s.image_height;
  d.a1 = s.a1; d.a2 = s.a2; d.a3 = s.a3; d.a4 = s.a4; d.sigma_measure = s.sigma_measure;
  d.epsilon = s.epsilon; d.newton_iters = s.newton_iters;
  {
    // distortion fast paths (srukf_device.cuh distort_point): largest |k1| ru^2 over the image (the corners; a pixel
    // zeroed by the view test is the corner (0, 0))
    double r2max = 0.0;
    for (int cx = 0; cx < 2; ++cx)
      for (int cy = 0; cy < 2; ++cy) {
        const double xu = (cx * (double)s.image_width - s.cam_cx) * s.cam_dx, yu = (cy * (double)s.image_height - s.cam_cy) * s.cam_dy;
        r2max = std::fmax(r2max, xu * xu + yu * yu);
      }
    d.dist_inward = (s.cam_k1 >= 0.0 && s.cam_k2 >= 0.0) ? 1 : 0;
    d.dist_series = (s.cam_k2 == 0.0 && d.dist_inward && s.cam_k1 * r2max <= 2.5e-4 && s.newton_iters >= 6) ? 1 : 0;
    if (getenv("SRUKF_NO_DIST_FASTPATH")) { d.dist_series = 0; d.dist_inward = 0; }
  }
  d.pred_free = (predict_free(d) && !getenv("SRUKF_PREDICT_BARRIERS")) ? 1 : 0;
  { const char* e_ = getenv("SRUKF_FORCE_FALLBACK_PPM"); d.force_fb_ppm = e_ ? atoi(e_) : 0; }
  { const char* e_ = getenv("SRUKF_DBG_SKIP_MMA"); d.dbg_skip_mma = e_ ? atoi(e_) : 0; }  // bit 0: skip DMMAs, bit 1: skip k_gain's loads
  double wm0, wc0, wi, wi_sr, gamma;
  sample_weights(s, d.Na, wm0, wc0, wi, wi_sr, gamma);
  const int Na = d.Na;
  d.gamma = gamma; d.wm0 = wm0; d.wc0 = wc0; d.wi = wi; d.wi_sr = wi_sr;
  d.Wsum = wm0 + 2.0 * Na * wi;
  d.cpair = std::sqrt(2.0) * wi_sr * gamma;
}

// 3-D tiled tensor map over doubles: dims {d0, d1, d2} (d0 contiguous), box {b0, b1, 1}
static int encode_map(CUtensorMap* m, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return fail(SRUKF_ECUDA, "cuTensorMapEncodeTiled is not available from the driver", e);
    encode = (PFN_cuTensorMapEncodeTiled)fn;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * sizeof(double), d0 * d1 * sizeof(double)};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return SRUKF_ECUDA;
  }
  return SRUKF_OK;
}

// (re)build the handle's tensor-map table: S buffers (by allocation order), Ut, dZ scratch, dZ_all
static int build_tensor_maps(srukf_handle* h, double* sbuf0, double* sbuf1) {
  const DevParams& p = h->p;
  CUtensorMap hm[TM_COUNT];
  memset(hm, 0, sizeof(hm));
  int rc;
  for (int r = 0; r < TM_ROWSETS; ++r) {
    if ((rc = encode_map(&hm[TM_S0 + r], sbuf0, p.np, p.np, p.B, TP, 8 * (r + 1)))) return rc;
    hm[TM_S1 + r] = hm[TM_S0 + r];   // in-place update: "old" and "new" factor are the same buffer
    (void)sbuf1;
    if ((rc = encode_map(&hm[TM_UT + r], h->U, p.np, p.Lc, h->chunk, TP, 8 * (r + 1)))) return rc;
    if (h->Useq) { if ((rc = encode_map(&hm[TM_USEQ + r], h->Useq, p.np, p.np, h->gslots, TP, 8 * (r + 1)))) return rc; }
    else hm[TM_USEQ + r] = hm[TM_UT + r];
    if (h->U_set[1]) { if ((rc = encode_map(&hm[TM_UT2 + r], h->U_set[1], p.np, p.Lc, h->chunk, TP, 8 * (r + 1)))) return rc; }
    else hm[TM_UT2 + r] = hm[TM_UT + r];
  }
  const uint32_t bpb = (uint32_t)gain_dz_box(p);
  if ((rc = encode_map(&hm[TM_DZ], h->dZ, p.Lc, p.np, h->chunk, bpb, 8))) return rc;
  if ((rc = encode_map(&hm[TM_DZ_ALL], h->dZ_all ? h->dZ_all : h->dZ, p.Lc, p.np, h->dZ_all ? p.B : h->chunk, bpb, 8)))
    return rc;
  if (!h->tmaps) CU(cudaMalloc(&h->tmaps, sizeof(hm)));
  CU(cudaMemcpyAsync(h->tmaps, hm, sizeof(hm), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SRUKF_OK;
}

int srukf_create(int device, int B, int L, const SrukfParams* params, srukf_t** out) {
  if (!out || B <= 0 || L <= 0) return fail(SRUKF_EINVAL, "srukf_create: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SRUKF_ENODEV, "srukf_create: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(SRUKF_EINVAL, "srukf_create: bad device index");
  SrukfParams prm;
  if (params) prm = *params; else srukf_default_params(&prm);
  if (prm.weight_type < 0 || prm.weight_type > 2) return fail(SRUKF_EINVAL, "srukf_create: weight_type must be 0..2");
  srukf_handle* h = new (std::nothrow) srukf_handle();
  if (!h) return fail(SRUKF_ENOMEM, "srukf_create: host allocation failed");
  h->device = device; h->prm = prm;
  fill_dev_params(h->p, prm, B, L);
  const DevParams& p = h->p;
  if (std::fabs(p.cpair - 1.0) > 1e-12) { delete h; return fail(SRUKF_EINVAL, "srukf_create: sqrt(2)*wi_sr*gamma != 1"); }
  if (predict_smem_bytes(p) > 227 * 1024 || tile_warps(p) == 0 || update_smem_bytes(p) > 227 * 1024 ||
      gain_smem_bytes(p) > 227 * 1024) {
    delete h;
    return fail(SRUKF_EINVAL, "srukf_create: L too large for one CTA per filter (supported: L <= 212, n = 1276)");
  }
  if (gain_row_ctas(p) > 1 && p.wc0 != p.wm0) {
    delete h;
    return fail(SRUKF_EINVAL, "srukf_create: weight type 1 (FLAG_4_WEIGHT2) is limited to L <= 106: its feature-sequential "
                              "state shift (SLAM.cpp:2030) needs all rows of U in one CTA");
  }
  if (prm.downdate_mode < 0 || prm.downdate_mode > 2) { delete h; return fail(SRUKF_EINVAL, "srukf_create: downdate_mode must be 0..2"); }
  cudaError_t e;
#define CUH(x) do { e = (x); if (e != cudaSuccess) { int c_ = fail(e == cudaErrorMemoryAllocation ? SRUKF_ENOMEM : SRUKF_ECUDA, #x, e); srukf_destroy(h); return c_; } } while (0)
  CUH(cudaSetDevice(device));
  CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUH(configure_kernels(p));
  const size_t n = p.n, L2 = 2 * (size_t)L;
  CUH(cudaMalloc(&h->x, sizeof(double) * B * n));
  CUH(cudaMalloc(&h->S, sizeof(double) * (size_t)B * p.nbp));
  if (prm.downdate_mode == 0) {
    CUH(cudaMalloc(&h->Pd, sizeof(double) * 2 * (size_t)B * p.np));   // [Pd | Ed]: k_update reaches Ed through its Pd pointer
    h->Ed = h->Pd + (size_t)B * p.np;
    CUH(cudaMemsetAsync(h->Pd, 0, sizeof(double) * 2 * (size_t)B * p.np, h->stream));
  }
  CUH(cudaMalloc(&h->u, sizeof(double) * B * 3));
  CUH(cudaMalloc(&h->z, sizeof(double) * B * L2));
  CUH(cudaMalloc(&h->matched, (size_t)B * L));
  CUH(cudaMalloc(&h->hbar, sizeof(double) * B * L2));
  CUH(cudaMalloc(&h->si, sizeof(double) * B * 4 * L));
  CUH(cudaMalloc(&h->cshift, sizeof(double) * B * L2));
  CUH(cudaMalloc(&h->pxyr, sizeof(double) * B * 4 * L2));
  CUH(cudaMalloc(&h->visible, (size_t)B * L));
  CUH(cudaMalloc(&h->flags, sizeof(uint32_t) * B));
  CUH(cudaMalloc(&h->perf, sizeof(double) * B * 4));
  CUH(cudaMalloc(&h->stats_out, sizeof(double) * 8));
  CUH(cudaMalloc(&h->truth, sizeof(double) * B * 3));
  CUH(cudaMemsetAsync(h->flags, 0, sizeof(uint32_t) * B, h->stream));
  CUH(cudaMemsetAsync(h->x, 0, sizeof(double) * B * n, h->stream));
  CUH(cudaMemsetAsync(h->S, 0, sizeof(double) * (size_t)B * p.nbp, h->stream));
  CUH(cudaMemsetAsync(h->visible, 0, (size_t)B * L, h->stream));
  // scratch: chunk sized so the pipeline scratch (dZ and Ut per filter) stays <= 12 GiB (SRUKF_SCRATCH_GIB): few, long launches keep
  // the idle tail of each kernel (the last wave of CTAs) small against its run time
  size_t per = sizeof(double) * 2 * (size_t)p.np * p.Lc;
  size_t budget = (size_t)12 << 30;   // 65,536 x L = 50: 3 chunks of 21,904 (6 GiB / 6 chunks: +0.65 % per step)
  if (const char* e_ = getenv("SRUKF_SCRATCH_GIB")) { const long g_ = atol(e_); if (g_ > 0) budget = (size_t)g_ << 30; }
  long chunk = (long)(budget / per);
  if (const char* e_ = getenv("SRUKF_CHUNK")) { const long c_ = atol(e_); if (c_ > 0 && c_ < chunk) chunk = c_; }   // tests: several chunks on a small batch
  if (chunk < 1) chunk = 1;
  if (chunk > B) chunk = B;
  if (chunk >= 592) chunk -= chunk % 296;  // whole waves of two CTAs per SM
  if (chunk >= 592 && chunk < B) {         // equal chunks (whole waves) instead of full ones and a short last one
    const long nch = (B + chunk - 1) / chunk;
    long even = (B + nch - 1) / nch;
    even += (296 - even % 296) % 296;
    if (even < chunk) chunk = even;
  }
  h->chunk = (int)chunk;
  h->gslots = (int)(chunk < 148 ? chunk : 148);
  if (prm.downdate_mode == 0 && update_seq_available(p) && tile_warps(p) == 8 && chunk > 148) h->gslots = (int)(chunk < 296 ? chunk : 296);
  CUH(cudaMalloc(&h->dZ, sizeof(double) * (size_t)chunk * p.np * p.Lc));
  CUH(cudaMalloc(&h->U, sizeof(double) * (size_t)chunk * p.Lc * p.np));
  CUH(cudaMalloc(&h->G, sizeof(double) * (size_t)h->gslots * p.ntri));
  if (prm.downdate_mode == 0) {
    CUH(cudaMalloc(&h->Gp, sizeof(double) * (size_t)h->gslots * p.ntri));
    if (update_seq_available(p)) {
      CUH(cudaMalloc(&h->Useq, sizeof(double) * (size_t)h->gslots * p.nbp));   // [np][np] per fallback CTA
      CUH(cudaMemsetAsync(h->Useq, 0, sizeof(double) * (size_t)h->gslots * p.nbp, h->stream));
    }
  }
  CUH(cudaMalloc(&h->worklist, sizeof(int) * ((size_t)chunk + 1)));
  CUH(cudaMalloc(&h->nact, sizeof(int) * (size_t)chunk));
  CUH(cudaMemsetAsync(h->nact, 0, sizeof(int) * (size_t)chunk, h->stream));
  CUH(cudaMemsetAsync(h->dZ, 0, sizeof(double) * (size_t)chunk * p.np * p.Lc, h->stream));
  CUH(cudaMemsetAsync(h->U, 0, sizeof(double) * (size_t)chunk * p.Lc * p.np, h->stream));
  CUH(cudaMemsetAsync(h->worklist, 0, sizeof(int) * ((size_t)chunk + 1), h->stream));
  h->U_set[0] = h->U; h->nact_set[0] = h->nact; h->wl_set[0] = h->worklist;
  if (h->Useq && chunk < B && !getenv("SRUKF_FALLBACK_INLINE")) {   // several chunks per step: overlap the fallback with the next chunk
    CUH(cudaMalloc(&h->U_set[1], sizeof(double) * (size_t)chunk * p.Lc * p.np));
    CUH(cudaMalloc(&h->wl_set[1], sizeof(int) * ((size_t)chunk + 1)));
    CUH(cudaMalloc(&h->nact_set[1], sizeof(int) * (size_t)chunk));
    CUH(cudaMemsetAsync(h->U_set[1], 0, sizeof(double) * (size_t)chunk * p.Lc * p.np, h->stream));
    CUH(cudaMemsetAsync(h->wl_set[1], 0, sizeof(int) * ((size_t)chunk + 1), h->stream));
    CUH(cudaMemsetAsync(h->nact_set[1], 0, sizeof(int) * (size_t)chunk, h->stream));
    int lo = 0, hi = 0;
    CUH(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority: the fallback fills what the main kernels leave
    CUH(cudaStreamCreateWithPriority(&h->fb_stream, cudaStreamNonBlocking, lo));
    for (int i = 0; i < 2; ++i) {
      CUH(cudaEventCreateWithFlags(&h->ev_upd[i], cudaEventDisableTiming));
      CUH(cudaEventCreateWithFlags(&h->ev_fb[i], cudaEventDisableTiming));
    }
  }
  if (const char* e_ = getenv("SRUKF_PHASE_TIMING")) {
    if (e_[0] == '1') {
      CUH(cudaMalloc(&h->dbg, sizeof(unsigned long long) * 16));
      CUH(cudaMemsetAsync(h->dbg, 0, sizeof(unsigned long long) * 16, h->stream));
    }
  }
  CUH(cudaStreamSynchronize(h->stream));
#undef CUH
  {
    int rc_ = build_tensor_maps(h, h->S, nullptr);
    if (rc_) { srukf_destroy(h); return rc_; }
  }
  *out = h;
  return SRUKF_OK;
}

int srukf_destroy(srukf_t* h) {
  if (!h) return SRUKF_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->fb_stream) { cudaStreamSynchronize(h->fb_stream); cudaStreamDestroy(h->fb_stream); }
  for (cudaEvent_t e : {h->ev_upd[0], h->ev_upd[1], h->ev_fb[0], h->ev_fb[1]}) if (e) cudaEventDestroy(e);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->d2h_stream) { cudaStreamSynchronize(h->d2h_stream); cudaStreamDestroy(h->d2h_stream); }
  for (cudaEvent_t e : {h->ev_in[0], h->ev_in[1], h->ev_done[0], h->ev_done[1], h->ev_xs, h->ev_xd}) if (e) cudaEventDestroy(e);
  void* ptrs[] = {h->U_set[1], h->nact_set[1], h->wl_set[1], h->Useq, h->Gp, h->u_in[1], h->z_in[1], h->m_in[1], h->x_stage, h->nact, h->G2, h->Pd, h->tmaps, h->dbg, h->worklist, h->x, h->S, h->u, h->z, h->matched, h->hbar, h->si, h->cshift, h->pxyr, h->visible, h->flags,
                  h->dZ, h->U, h->G, h->rsig, h->dZ_all, h->U_all, h->G_all, h->perf, h->stats_out, h->truth};
  for (void* q : ptrs) if (q) cudaFree(q);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return SRUKF_OK;
}

static StepPtrs base_ptrs(srukf_t* h) {
  StepPtrs q{};
  q.x = h->x; q.S = h->S; q.u = h->u; q.z = h->z; q.matched = h->matched;
  q.hbar = h->hbar; q.si = h->si; q.visible = h->visible; q.cshift = h->cshift; q.pxyr = h->pxyr;
  q.rsig = h->rsig; q.dZ = h->dZ; q.U = h->U; q.G = h->G; q.flags = h->flags; q.chunk0 = 0;
  q.S2 = h->S; q.worklist = h->worklist; q.rel0 = 0; q.dbg = h->dbg;
  q.tmaps = h->tmaps; q.sbuf = 0; q.tm_dz = TM_DZ; q.tm_ut = TM_UT; q.dz_filter0 = 0;
  q.Pd = h->Pd; q.Pd2 = h->Pd; q.Ed = h->Ed; q.carry_p = (h->prm.downdate_mode == 0) ? 1 : 0;
  q.G2 = h->G2; q.n_new = 0; q.nact = h->nact; q.Gp = h->Gp; q.Useq = h->Useq;
  return q;
}

// external S (fmt 0 dense [B][n][n], fmt 1 upper-packed [B][ntri]) <-> internal layout, staged in slabs <= 256 MiB
static int transfer_S(srukf_t* h, int fmt, const double* src_host, double* dst_host) {
  const DevParams& p = h->p;
  fb_join(h);
  const size_t per = sizeof(double) * (fmt ? (size_t)p.ntri : (size_t)p.n * p.n);
  int slab = (int)(((size_t)256 << 20) / per);
  if (slab < 1) slab = 1;
  if (slab > p.B) slab = p.B;
  double* tmp = nullptr;
  CU(cudaMalloc(&tmp, per * slab));
  for (int b0 = 0; b0 < p.B; b0 += slab) {
    const int nb = p.B - b0 < slab ? p.B - b0 : slab;
    cudaError_t e = cudaSuccess;
    if (src_host) {
      e = cudaMemcpyAsync(tmp, (const char*)src_host + per * b0, per * nb, cudaMemcpyHostToDevice, h->stream);
      if (e == cudaSuccess) {
        launch_import(p, nb, fmt, tmp, h->S + (size_t)b0 * p.nbp, h->stream);
        if (h->Pd) { launch_form_P(p, b0, nb, h->S, h->Pd, h->stream); h->launches++; }
      }
    } else {
      launch_export(p, nb, fmt, h->S + (size_t)b0 * p.nbp, tmp, h->stream);
      e = cudaMemcpyAsync((char*)dst_host + per * b0, tmp, per * nb, cudaMemcpyDeviceToHost, h->stream);
    }
    h->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(tmp); return fail(SRUKF_ECUDA, "srukf state transfer", e); }
  }
  cudaFree(tmp);
  return SRUKF_OK;
}

static int set_state_any(srukf_t* h, const double* x, const double* S, int fmt, const char* who) {
  if (!h || !x || !S) return fail(SRUKF_EINVAL, who);
  CU(cudaSetDevice(h->device));
  int rc = transfer_S(h, fmt, S, nullptr);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->x, x, sizeof(double) * (size_t)h->p.B * h->p.n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->phase = 0;
  return SRUKF_OK;
}

static int get_state_any(srukf_t* h, double* x, double* S, int fmt, const char* who) {
  if (!h) return fail(SRUKF_EINVAL, who);
  CU(cudaSetDevice(h->device));
  if (S) {
    int rc = transfer_S(h, fmt, nullptr, S);
    if (rc) return rc;
  }
  if (x) CU(cudaMemcpyAsync(x, h->x, sizeof(double) * (size_t)h->p.B * h->p.n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SRUKF_OK;
}

int srukf_init_features(srukf_t* h, const double* x4, const double* S4, const double* keypoints, double rho0,
                        double sigma_rho) {
  if (!h || !x4 || !S4 || !keypoints) return fail(SRUKF_EINVAL, "srukf_init_features: null argument");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  const DevParams& p = h->p;
  const size_t B = (size_t)p.B;
  double wm0, wc0, wi, wi_sr, gamma;
  sample_weights(h->prm, 4 + 3 * p.L, wm0, wc0, wi, wi_sr, gamma);   // Na of the initialisation, SLAM.cpp:827,867
  double* d_in = nullptr;
  CU(cudaMalloc(&d_in, sizeof(double) * B * (20 + 2 * (size_t)p.L)));
  double *d_x4 = d_in, *d_S4 = d_in + 4 * B, *d_kp = d_in + 20 * B;
  cudaError_t e = cudaMemcpyAsync(d_x4, x4, sizeof(double) * 4 * B, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_S4, S4, sizeof(double) * 16 * B, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_kp, keypoints, sizeof(double) * 2 * p.L * B, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    double* Pd = h->Pd;   // null in downdate modes 1 / 2 (no carried covariance)
    const int nblocks = (int)(B < (size_t)h->gslots ? B : (size_t)h->gslots);
    launch_init_features(p, nblocks, d_x4, d_S4, d_kp, rho0, sigma_rho, gamma, wi, h->x, h->S, Pd, h->G, h->flags,
                         h->stream);
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d_in);
  if (e != cudaSuccess) return fail(SRUKF_ECUDA, "srukf_init_features", e);
  h->phase = 0;
  return SRUKF_OK;
}

int srukf_add_features(srukf_t* src, srukf_t* dst, const double* keypoints, double rho0, double sigma_rho) {
  if (!src || !dst || !keypoints) return fail(SRUKF_EINVAL, "srukf_add_features: null argument");
  const int M = dst->p.L - src->p.L;
  if (src == dst || dst->p.B != src->p.B || M < 1 || dst->device != src->device)
    return fail(SRUKF_EINVAL, "srukf_add_features: dst must be another handle on the same device with the same B and more features");
  CU(cudaSetDevice(dst->device));
  CU(cudaStreamSynchronize(src->stream));
  fb_join(src); fb_join(dst);
  if (src->fb_stream) CU(cudaStreamSynchronize(src->fb_stream));
  const DevParams& p = dst->p;
  const int ns = src->p.n;
  double wm0, wc0, wi, wi_sr, gamma;
  sample_weights(dst->prm, ns + 3 * M, wm0, wc0, wi, wi_sr, gamma);   // Na = dim + 3 m_nFilters, SLAM.cpp:827,867
  const int nblocks = p.B < dst->gslots ? p.B : dst->gslots;
  double *d_kp = nullptr, *d_A = nullptr;
  CU(cudaMalloc(&d_kp, sizeof(double) * (size_t)p.B * 2 * M));
  cudaError_t e = cudaMalloc(&d_A, sizeof(double) * (size_t)nblocks * ((size_t)M * 4 * ns + 12 * M));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_kp, keypoints, sizeof(double) * (size_t)p.B * 2 * M, cudaMemcpyHostToDevice, dst->stream);
  if (e == cudaSuccess) {
    launch_add_features(p, nblocks, d_kp, rho0, sigma_rho, gamma, wi, src->x, src->S, ns, src->p.np, M, dst->x, dst->S,
                        dst->Pd, dst->G, d_A, src->flags, dst->flags, dst->stream);
    dst->launches++;
    e = cudaStreamSynchronize(dst->stream);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d_kp);
  if (d_A) cudaFree(d_A);
  if (e != cudaSuccess) return fail(SRUKF_ECUDA, "srukf_add_features", e);
  dst->phase = 0;
  return SRUKF_OK;
}

int srukf_delete_feature(srukf_t* src, srukf_t* dst, const int32_t* ids) {
  if (!src || !dst || !ids) return fail(SRUKF_EINVAL, "srukf_delete_feature: null argument");
  if (src == dst || dst->p.B != src->p.B || dst->p.L != src->p.L - 1 || dst->device != src->device)
    return fail(SRUKF_EINVAL, "srukf_delete_feature: dst must be another handle on the same device with the same B and L-1 features");
  for (int b = 0; b < src->p.B; ++b)
    if (ids[b] < 0 || ids[b] >= src->p.L) return fail(SRUKF_EINVAL, "srukf_delete_feature: feature id out of range");
  CU(cudaSetDevice(dst->device));
  CU(cudaStreamSynchronize(src->stream));
  fb_join(src); fb_join(dst);
  if (src->fb_stream) CU(cudaStreamSynchronize(src->fb_stream));
  const DevParams& p = dst->p;
  const int nblocks = p.B < dst->gslots ? p.B : dst->gslots;
  int* d_ids = nullptr;
  double* d_V = nullptr;
  CU(cudaMalloc(&d_ids, sizeof(int) * (size_t)p.B));
  cudaError_t e = cudaMalloc(&d_V, sizeof(double) * (size_t)nblocks * 6 * p.np);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_ids, ids, sizeof(int) * (size_t)p.B, cudaMemcpyHostToDevice, dst->stream);
  if (e == cudaSuccess) {
    launch_delete_feature(p, nblocks, src->x, src->S, src->p.n, src->p.np, d_ids, dst->x, dst->S, dst->Pd, dst->G, d_V,
                          src->flags, dst->flags, dst->stream);
    dst->launches++;
    e = cudaStreamSynchronize(dst->stream);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d_ids);
  if (d_V) cudaFree(d_V);
  if (e != cudaSuccess) return fail(SRUKF_ECUDA, "srukf_delete_feature", e);
  dst->phase = 0;
  return SRUKF_OK;
}

int srukf_set_state(srukf_t* h, const double* x, const double* S_packed) {
  return set_state_any(h, x, S_packed, 1, "srukf_set_state: null argument");
}
int srukf_get_state(srukf_t* h, double* x, double* S_packed) {
  return get_state_any(h, x, S_packed, 1, "srukf_get_state: null handle");
}
int srukf_set_state_dense(srukf_t* h, const double* x, const double* S_dense) {
  return set_state_any(h, x, S_dense, 0, "srukf_set_state_dense: null argument");
}
int srukf_get_state_dense(srukf_t* h, double* x, double* S_dense) {
  return get_state_any(h, x, S_dense, 0, "srukf_get_state_dense: null handle");
}

static int ensure_split_buffers(srukf_t* h) {
  const DevParams& p = h->p;
  if (!h->rsig) CU(cudaMalloc(&h->rsig, sizeof(double) * (size_t)p.B * p.P * 4));
  if (h->chunk >= p.B) return SRUKF_OK;  // the step scratch already covers the whole batch
  if (!h->dZ_all) {
    CU(cudaMalloc(&h->dZ_all, sizeof(double) * (size_t)p.B * p.np * p.Lc));
    CU(cudaMemsetAsync(h->dZ_all, 0, sizeof(double) * (size_t)p.B * p.np * p.Lc, h->stream));
    // buffers by allocation order: the current one is index h->sbuf
    int rc_ = build_tensor_maps(h, h->S, nullptr);
    if (rc_) return rc_;
  }
  return SRUKF_OK;
}

int srukf_predict_motion(srukf_t* h, const double* u) {
  if (!h || !u) return fail(SRUKF_EINVAL, "srukf_predict_motion: null argument");
  h->inputs_dirty = true;
  CU(cudaSetDevice(h->device));
  fb_join(h);
  int rc = ensure_split_buffers(h);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->u, u, sizeof(double) * (size_t)h->p.B * 3, cudaMemcpyHostToDevice, h->stream));
  StepPtrs q = base_ptrs(h);
  launch_predict(h->p, q, h->p.B, true, false, true, h->stream);
  h->launches++;
  CU(cudaGetLastError());
  h->phase = 1;
  return SRUKF_OK;
}

int srukf_predict_measurement(srukf_t* h) {
  if (!h) return fail(SRUKF_EINVAL, "srukf_predict_measurement: null handle");
  if (h->phase != 1) return fail(SRUKF_ESTATE, "srukf_predict_measurement: call srukf_predict_motion first");
  CU(cudaSetDevice(h->device));
  StepPtrs q = base_ptrs(h);
  if (h->dZ_all) q.dZ = h->dZ_all;
  launch_predict(h->p, q, h->p.B, false, true, false, h->stream);
  h->launches++;
  CU(cudaGetLastError());
  h->phase = 2;
  return SRUKF_OK;
}

int srukf_get_prediction(srukf_t* h, double* hbar, double* si, uint8_t* visible) {
  if (!h) return fail(SRUKF_EINVAL, "srukf_get_prediction: null handle");
  if (h->phase < 2) return fail(SRUKF_ESTATE, "srukf_get_prediction: no prediction available");
  CU(cudaSetDevice(h->device));
  const size_t B = h->p.B, L = h->p.L;
  if (hbar) CU(cudaMemcpyAsync(hbar, h->hbar, sizeof(double) * B * 2 * L, cudaMemcpyDeviceToHost, h->stream));
  if (si) CU(cudaMemcpyAsync(si, h->si, sizeof(double) * B * 4 * L, cudaMemcpyDeviceToHost, h->stream));
  if (visible) CU(cudaMemcpyAsync(visible, h->visible, B * L, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SRUKF_OK;
}

int srukf_chi2_gate(srukf_t* h, const double* z, double threshold, uint8_t* accept, double* d2) {
  if (!h || !z || !accept) return fail(SRUKF_EINVAL, "srukf_chi2_gate: null argument");
  if (h->phase != 2) return fail(SRUKF_ESTATE, "srukf_chi2_gate: call srukf_predict_measurement first");
  h->inputs_dirty = true;
  CU(cudaSetDevice(h->device));
  const DevParams& p = h->p;
  const size_t BL = (size_t)p.B * p.L;
  CU(cudaMemcpyAsync(h->z, z, sizeof(double) * BL * 2, cudaMemcpyHostToDevice, h->stream));
  double* d_d2 = nullptr;
  if (d2) CU(cudaMalloc(&d_d2, sizeof(double) * BL));
  launch_gate(p, h->z, h->hbar, h->si, h->visible, threshold, h->matched, d_d2, h->stream);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(accept, h->matched, BL, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && d2) e = cudaMemcpyAsync(d2, d_d2, sizeof(double) * BL, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (d_d2) cudaFree(d_d2);
  if (e != cudaSuccess) return fail(SRUKF_ECUDA, "srukf_chi2_gate", e);
  return SRUKF_OK;
}

// gain + covariance update over [b0, b0+nb) with scratch indexed from 0
static void run_update(srukf_t* h, StepPtrs q, int b0, int nb) {
  q.chunk0 = b0;
  const bool side = h->fb_stream != nullptr && h->prm.downdate_mode == 0;
  int set = 0;
  if (side) {
    set = h->uset;
    h->uset ^= 1;
    if (h->fbp[set].pending) {   // the fallback that still reads this scratch set
      cudaStreamWaitEvent(h->stream, h->ev_fb[set], 0);
      h->fbp[set].pending = false;
    }
    q.U = h->U_set[set]; q.nact = h->nact_set[set]; q.worklist = h->wl_set[set];
    q.tm_ut = set ? TM_UT2 : TM_UT;
  }
  prof_begin(h, 1);
  launch_gain(h->p, q, nb, h->stream);
  prof_end(h);
  h->launches++;
  prof_begin(h, 2);
  if (h->prm.downdate_mode == 0) {
    cudaMemsetAsync(q.worklist, 0, sizeof(int), h->stream);
    launch_update(h->p, q, nb, h->stream);
    // redo of the filters the guard queued (usually none): bisection over column groups on the tensor pipe (on the side
    // stream when the step has several chunks), or the literal per-column sequence where that kernel is not built
    if (side) {
      cudaEventRecord(h->ev_upd[set], h->stream);
      cudaStreamWaitEvent(h->fb_stream, h->ev_upd[set], 0);
      launch_update_seq(h->p, q, h->gslots, h->fb_stream);
      cudaEventRecord(h->ev_fb[set], h->fb_stream);
      h->fbp[set].b0 = b0; h->fbp[set].nb = nb; h->fbp[set].pending = true;
    } else if (h->Useq) {
      launch_update_seq(h->p, q, h->gslots, h->stream);
    } else {
      launch_downdate(h->p, q, h->gslots < 148 ? h->gslots : 148, 1, 1, h->stream);
    }
    h->launches += 2;
  } else {
    for (int r0 = 0; r0 < nb; r0 += h->gslots) {
      q.rel0 = r0;
      int m = nb - r0 < h->gslots ? nb - r0 : h->gslots;
      launch_downdate(h->p, q, m, h->prm.downdate_mode, 0, h->stream);
      h->launches++;
    }
  }
  prof_end(h);
}

// (the fused update works in place: there is no second S buffer to swap)
static void flip_buffers(srukf_t*) {}

int srukf_kalman_update(srukf_t* h, const double* z, const uint8_t* matched) {
  if (!h || !z || !matched) return fail(SRUKF_EINVAL, "srukf_kalman_update: null argument");
  if (h->phase != 2) return fail(SRUKF_ESTATE, "srukf_kalman_update: call srukf_predict_measurement first");
  h->inputs_dirty = true;
  CU(cudaSetDevice(h->device));
  fb_join(h);
  const DevParams& p = h->p;
  CU(cudaMemcpyAsync(h->z, z, sizeof(double) * (size_t)p.B * 2 * p.L, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->matched, matched, (size_t)p.B * p.L, cudaMemcpyHostToDevice, h->stream));
  StepPtrs q = base_ptrs(h);
  for (int b0 = 0; b0 < p.B; b0 += h->chunk) {
    int nb = p.B - b0 < h->chunk ? p.B - b0 : h->chunk;
    StepPtrs qq = q;
    if (h->dZ_all) { qq.dZ = h->dZ_all + (size_t)b0 * p.np * p.Lc; qq.tm_dz = TM_DZ_ALL; qq.dz_filter0 = b0; }
    run_update(h, qq, b0, nb);
  }
  flip_buffers(h);
  CU(cudaGetLastError());
  h->phase = 0;
  return SRUKF_OK;
}

int srukf_kalman_update_reorder(srukf_t* h, const double* z, const uint8_t* matched, int n_new) {
  if (!h || !z || !matched) return fail(SRUKF_EINVAL, "srukf_kalman_update_reorder: null argument");
  if (n_new < 0 || n_new > h->p.L) return fail(SRUKF_EINVAL, "srukf_kalman_update_reorder: n_new must be in 0..L");
  if (n_new == 0) return srukf_kalman_update(h, z, matched);   // m_nAddings == 0: NEEDNOT_REORDER (:2087-2090)
  if (h->phase != 2) return fail(SRUKF_ESTATE, "srukf_kalman_update_reorder: call srukf_predict_measurement first");
  h->inputs_dirty = true;
  CU(cudaSetDevice(h->device));
  fb_join(h);
  const DevParams& p = h->p;
  if (!h->G2) CU(cudaMalloc(&h->G2, sizeof(double) * (size_t)h->gslots * ((size_t)p.ntri + 2 * (size_t)p.nbp)));
  CU(cudaMemcpyAsync(h->z, z, sizeof(double) * (size_t)p.B * 2 * p.L, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->matched, matched, (size_t)p.B * p.L, cudaMemcpyHostToDevice, h->stream));
  StepPtrs q = base_ptrs(h);
  q.n_new = n_new;
  for (int b0 = 0; b0 < p.B; b0 += h->chunk) {
    const int nb = p.B - b0 < h->chunk ? p.B - b0 : h->chunk;
    StepPtrs qq = q;
    if (h->dZ_all) { qq.dZ = h->dZ_all + (size_t)b0 * p.np * p.Lc; qq.tm_dz = TM_DZ_ALL; qq.dz_filter0 = b0; }
    qq.chunk0 = b0;
    launch_gain(p, qq, nb, h->stream);
    h->launches++;
    for (int r0 = 0; r0 < nb; r0 += h->gslots) {   // reference order, in place on the current factor
      qq.rel0 = r0;
      const int m = nb - r0 < h->gslots ? nb - r0 : h->gslots;
      launch_downdate(p, qq, m, 3, 0, h->stream);
      h->launches++;
    }
    if (h->Pd) { launch_form_P(p, b0, nb, h->S, h->Pd, h->stream); h->launches++; }   // carried covariance rebuilt
  }
  CU(cudaGetLastError());
  h->phase = 0;
  return SRUKF_OK;
}

int srukf_step_dev(srukf_t* h, const double* d_u, const double* d_z, const uint8_t* d_matched) {
  if (!h || !d_u || !d_z || !d_matched) return fail(SRUKF_EINVAL, "srukf_step_dev: null argument");
  CU(cudaSetDevice(h->device));
  const DevParams& p = h->p;
  StepPtrs q = base_ptrs(h);
  q.u = d_u; q.z = d_z; q.matched = d_matched;
  for (int b0 = 0; b0 < p.B; b0 += h->chunk) {
    int nb = p.B - b0 < h->chunk ? p.B - b0 : h->chunk;
    q.chunk0 = b0;
    fb_join(h, b0, nb);   // a fallback of the previous frame that is still rewriting these filters
    prof_begin(h, 0);
    launch_predict(p, q, nb, true, true, false, h->stream);
    prof_end(h);
    h->launches++;
    run_update(h, q, b0, nb);
  }
  flip_buffers(h);
  CU(cudaGetLastError());
  h->phase = 0;
  return SRUKF_OK;
}

// lazily created: second input set, copy / read-back streams and their events
static int ensure_async(srukf_t* h) {
  if (h->copy_stream) return SRUKF_OK;
  const DevParams& p = h->p;
  CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  h->u_in[0] = h->u; h->z_in[0] = h->z; h->m_in[0] = h->matched;
  CU(cudaMalloc(&h->u_in[1], sizeof(double) * (size_t)p.B * 3));
  CU(cudaMalloc(&h->z_in[1], sizeof(double) * (size_t)p.B * 2 * p.L));
  CU(cudaMalloc(&h->m_in[1], (size_t)p.B * p.L));
  for (int i = 0; i < 2; ++i) {
    CU(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&h->ev_xs, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->ev_xd, cudaEventDisableTiming));
  return SRUKF_OK;
}

int srukf_step(srukf_t* h, const double* u, const double* z, const uint8_t* matched) {
  if (!h || !u || !z || !matched) return fail(SRUKF_EINVAL, "srukf_step: null argument");
  CU(cudaSetDevice(h->device));
  int rc = ensure_async(h);
  if (rc) return rc;
  const DevParams& p = h->p;
  const int s = h->in_slot;
  h->in_slot ^= 1;
  // this input set was last read by the srukf_step before the previous one (ev_done[s]); after any other entry point
  // that uses h->u / h->z / h->matched (= set 0) the copy waits for everything queued on the handle's stream
  if (h->inputs_dirty) {
    CU(cudaEventRecord(h->ev_done[s], h->stream));
    h->inputs_dirty = false;
  }
  CU(cudaStreamWaitEvent(h->copy_stream, h->ev_done[s], 0));
  CU(cudaMemcpyAsync(h->u_in[s], u, sizeof(double) * (size_t)p.B * 3, cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaMemcpyAsync(h->z_in[s], z, sizeof(double) * (size_t)p.B * 2 * p.L, cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaMemcpyAsync(h->m_in[s], matched, (size_t)p.B * p.L, cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaEventRecord(h->ev_in[s], h->copy_stream));
  CU(cudaStreamWaitEvent(h->stream, h->ev_in[s], 0));
  rc = srukf_step_dev(h, h->u_in[s], h->z_in[s], h->m_in[s]);
  if (rc) return rc;
  CU(cudaEventRecord(h->ev_done[s], h->stream));
  return SRUKF_OK;
}

int srukf_get_x_async(srukf_t* h, double* x_host) {
  if (!h || !x_host) return fail(SRUKF_EINVAL, "srukf_get_x_async: null argument");
  CU(cudaSetDevice(h->device));
  int rc = ensure_async(h);
  if (rc) return rc;
  const size_t bytes = sizeof(double) * (size_t)h->p.B * h->p.n;
  if (!h->x_stage) {
    CU(cudaMalloc(&h->x_stage, bytes));
    CU(cudaEventRecord(h->ev_xd, h->d2h_stream));
  }
  CU(cudaStreamWaitEvent(h->stream, h->ev_xd, 0));    // the previous read-back has left the staging copy
  CU(cudaMemcpyAsync(h->x_stage, h->x, bytes, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaEventRecord(h->ev_xs, h->stream));
  CU(cudaStreamWaitEvent(h->d2h_stream, h->ev_xs, 0));
  CU(cudaMemcpyAsync(x_host, h->x_stage, bytes, cudaMemcpyDeviceToHost, h->d2h_stream));
  CU(cudaEventRecord(h->ev_xd, h->d2h_stream));
  return SRUKF_OK;
}

int srukf_set_state_dev(srukf_t* h, int b0, int nb, const double* d_x, const double* d_S_packed) {
  if (!h || b0 < 0 || nb <= 0 || b0 + nb > h->p.B) return fail(SRUKF_EINVAL, "srukf_set_state_dev: bad arguments");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  if (d_x)
    CU(cudaMemcpyAsync(h->x + (size_t)b0 * h->p.n, d_x, sizeof(double) * (size_t)nb * h->p.n, cudaMemcpyDeviceToDevice,
                       h->stream));
  if (d_S_packed) {
    launch_import(h->p, nb, 1, d_S_packed, h->S + (size_t)b0 * h->p.nbp, h->stream);
    h->launches++;
    if (h->Pd) { launch_form_P(h->p, b0, nb, h->S, h->Pd, h->stream); h->launches++; }
  }
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  h->phase = 0;
  return SRUKF_OK;
}

int srukf_get_cov_block(srukf_t* h, int r0, int nr, double* out) {
  if (!h || !out || r0 < 0 || nr <= 0 || r0 + nr > h->p.n) return fail(SRUKF_EINVAL, "srukf_get_cov_block: bad arguments");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  double* tmp = nullptr;
  size_t bytes = sizeof(double) * (size_t)h->p.B * nr * nr;
  CU(cudaMalloc(&tmp, bytes));
  launch_cov_block(h->p, h->S, r0, nr, tmp, h->stream);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(SRUKF_ECUDA, "srukf_get_cov_block", e);
  return SRUKF_OK;
}

int srukf_get_flags(srukf_t* h, uint32_t* flags) {
  if (!h || !flags) return fail(SRUKF_EINVAL, "srukf_get_flags: null argument");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  CU(cudaMemcpyAsync(flags, h->flags, sizeof(uint32_t) * h->p.B, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SRUKF_OK;
}

int srukf_clear_flags(srukf_t* h) {
  if (!h) return fail(SRUKF_EINVAL, "srukf_clear_flags: null handle");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  CU(cudaMemsetAsync(h->flags, 0, sizeof(uint32_t) * h->p.B, h->stream));
  return SRUKF_OK;
}

int srukf_stats(srukf_t* h, const double* truth, double* out8) {
  if (!h || !truth || !out8) return fail(SRUKF_EINVAL, "srukf_stats: null argument");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  CU(cudaMemcpyAsync(h->truth, truth, sizeof(double) * (size_t)h->p.B * 3, cudaMemcpyHostToDevice, h->stream));
  launch_stats(h->p, h->x, h->S, h->truth, h->perf, h->flags, h->stats_out, h->stream);
  h->launches += 2;
  CU(cudaMemcpyAsync(out8, h->stats_out, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SRUKF_OK;
}

// ---- CSLAM helper methods (SLAM.h:322,341,347-348,355) as stand-alone entry points ---------------------------------
static int helper_device(int device, const char* who) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SRUKF_ENODEV, who);
  if (device < 0 || device >= ndev) return fail(SRUKF_EINVAL, who);
  CU(cudaSetDevice(device));
  DevParams dummy{};
  CU(configure_kernels(dummy));
  return SRUKF_OK;
}

int srukf_mchol(int device, int nb, int n, double epsilon, const double* G, double* S, uint32_t* flags) {
  if (nb <= 0 || n <= 0 || !G || !S) return fail(SRUKF_EINVAL, "srukf_mchol: bad arguments");
  int rc = helper_device(device, "srukf_mchol: no CUDA device (there is no CPU fallback)");
  if (rc) return rc;
  const size_t mat = sizeof(double) * (size_t)n * n, ntri = (size_t)n * (n + 1) / 2;
  const int grid = nb < 296 ? nb : 296;
  double *dG = nullptr, *dS = nullptr, *dP = nullptr;
  uint32_t* dF = nullptr;
  cudaError_t e = cudaMalloc(&dG, mat * nb);
  if (!e) e = cudaMalloc(&dS, mat * nb);
  if (!e) e = cudaMalloc(&dP, sizeof(double) * ntri * grid);
  if (!e) e = cudaMalloc(&dF, sizeof(uint32_t) * nb);
  if (!e) e = cudaMemcpy(dG, G, mat * nb, cudaMemcpyHostToDevice);
  if (!e) { launch_mchol_batch(nb, n, epsilon, dG, dP, dS, dF, grid, 0); e = cudaGetLastError(); }
  if (!e) e = cudaMemcpy(S, dS, mat * nb, cudaMemcpyDeviceToHost);
  if (!e && flags) e = cudaMemcpy(flags, dF, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost);
  cudaFree(dG); cudaFree(dS); cudaFree(dP); cudaFree(dF);
  if (e) return fail(e == cudaErrorMemoryAllocation ? SRUKF_ENOMEM : SRUKF_ECUDA, "srukf_mchol", e);
  return SRUKF_OK;
}

int srukf_qr_R(int device, int nb, int m, int n, const double* A, double* R) {
  if (nb <= 0 || m <= 0 || n <= 0 || m < n || !A || !R) return fail(SRUKF_EINVAL, "srukf_qr_R: bad arguments (need m >= n)");
  int rc = helper_device(device, "srukf_qr_R: no CUDA device (there is no CPU fallback)");
  if (rc) return rc;
  const size_t amat = sizeof(double) * (size_t)m * n, rmat = sizeof(double) * (size_t)n * n;
  const int grid = nb < 296 ? nb : 296;
  double *dA = nullptr, *dW = nullptr, *dR = nullptr;
  cudaError_t e = cudaMalloc(&dA, amat * nb);
  if (!e) e = cudaMalloc(&dW, amat * grid);
  if (!e) e = cudaMalloc(&dR, rmat * nb);
  if (!e) e = cudaMemcpy(dA, A, amat * nb, cudaMemcpyHostToDevice);
  if (!e) { launch_qr_batch(nb, m, n, dA, dW, dR, grid, 0); e = cudaGetLastError(); }
  if (!e) e = cudaMemcpy(R, dR, rmat * nb, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dW); cudaFree(dR);
  if (e) return fail(e == cudaErrorMemoryAllocation ? SRUKF_ENOMEM : SRUKF_ECUDA, "srukf_qr_R", e);
  return SRUKF_OK;
}

int srukf_generate_sigma_points(int device, int nb, int Na, double gamma, const double* mu, const double* sr,
                                double* sigma) {
  if (nb <= 0 || Na <= 0 || !mu || !sr || !sigma) return fail(SRUKF_EINVAL, "srukf_generate_sigma_points: bad arguments");
  int rc = helper_device(device, "srukf_generate_sigma_points: no CUDA device (there is no CPU fallback)");
  if (rc) return rc;
  const size_t P = 2 * (size_t)Na + 1;
  double *dm = nullptr, *ds = nullptr, *dg = nullptr;
  cudaError_t e = cudaMalloc(&dm, sizeof(double) * Na * nb);
  if (!e) e = cudaMalloc(&ds, sizeof(double) * (size_t)Na * Na * nb);
  if (!e) e = cudaMalloc(&dg, sizeof(double) * Na * P * nb);
  if (!e) e = cudaMemcpy(dm, mu, sizeof(double) * Na * nb, cudaMemcpyHostToDevice);
  if (!e) e = cudaMemcpy(ds, sr, sizeof(double) * (size_t)Na * Na * nb, cudaMemcpyHostToDevice);
  if (!e) { launch_sigma_points(nb, Na, gamma, dm, ds, dg, 0); e = cudaGetLastError(); }
  if (!e) e = cudaMemcpy(sigma, dg, sizeof(double) * Na * P * nb, cudaMemcpyDeviceToHost);
  cudaFree(dm); cudaFree(ds); cudaFree(dg);
  if (e) return fail(e == cudaErrorMemoryAllocation ? SRUKF_ENOMEM : SRUKF_ECUDA, "srukf_generate_sigma_points", e);
  return SRUKF_OK;
}

int srukf_cholesky_update(srukf_t* h, const double* U, int k, int up_or_down, int order, int n_new) {
  if (!h || !U || k <= 0) return fail(SRUKF_EINVAL, "srukf_cholesky_update: bad arguments");
  if ((up_or_down != SRUKF_UPDATING && up_or_down != SRUKF_DOWNDATING) ||
      (order != SRUKF_NEED_REORDER && order != SRUKF_NEEDNOT_REORDER))
    return fail(SRUKF_EINVAL, "srukf_cholesky_update: bad flag (use the SRUKF_* values of srukf.h)");
  const DevParams& p = h->p;
  const int M = (order == SRUKF_NEED_REORDER) ? n_new : 0;
  if (order == SRUKF_NEED_REORDER && (n_new < 1 || n_new > p.L))
    return fail(SRUKF_EINVAL, "srukf_cholesky_update: NEED_REORDER needs 1 <= n_new <= L");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  if (M && !h->G2) CU(cudaMalloc(&h->G2, sizeof(double) * (size_t)h->gslots * ((size_t)p.ntri + 2 * (size_t)p.nbp)));
  // u is dim x k per filter (a cv::Mat in the reference); the kernels want its columns contiguous and padded to np
  std::vector<double> ut((size_t)p.B * k * p.np, 0.0);
  for (int b = 0; b < p.B; ++b)
    for (int r = 0; r < p.n; ++r)
      for (int c = 0; c < k; ++c) ut[((size_t)b * k + c) * p.np + r] = U[((size_t)b * p.n + r) * k + c];
  double* dU = nullptr;
  CU(cudaMalloc(&dU, sizeof(double) * ut.size()));
  cudaError_t e = cudaMemcpyAsync(dU, ut.data(), sizeof(double) * ut.size(), cudaMemcpyHostToDevice, h->stream);
  if (!e) {
    const int grid = p.B < h->gslots ? p.B : h->gslots;
    launch_chol_update(p, grid, h->S, h->Pd, dU, k, up_or_down == SRUKF_UPDATING ? +1.0 : -1.0, M, h->G, h->G2, h->flags,
                       h->stream);
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
  }
  if (!e) e = cudaGetLastError();
  cudaFree(dU);
  if (e) return fail(SRUKF_ECUDA, "srukf_cholesky_update", e);
  h->phase = 0;
  return SRUKF_OK;
}

int srukf_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(SRUKF_EINVAL, "srukf_fp64_peak: null argument");
  int rc = helper_device(device, "srukf_fp64_peak: no CUDA device");
  if (rc) return rc;
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CU(measure_fp64_peak(sms, 20000, 5, tflops, 0));
  return SRUKF_OK;
}

int srukf_sync(srukf_t* h) {
  if (!h) return fail(SRUKF_EINVAL, "srukf_sync: null handle");
  CU(cudaSetDevice(h->device));
  fb_join(h);
  CU(cudaStreamSynchronize(h->stream));
  if (h->fb_stream) CU(cudaStreamSynchronize(h->fb_stream));
  if (h->copy_stream) CU(cudaStreamSynchronize(h->copy_stream));
  if (h->d2h_stream) CU(cudaStreamSynchronize(h->d2h_stream));
  CU(cudaGetLastError());
  return SRUKF_OK;
}

int srukf_set_profiling(srukf_t* h, int on) {
  if (!h) return fail(SRUKF_EINVAL, "srukf_set_profiling: null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  prof_collect(h);
  h->profiling = on != 0;
  for (int i = 0; i < 3; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  return SRUKF_OK;
}

int srukf_get_kernel_times(srukf_t* h, double* ms3, uint64_t* launches3) {
  if (!h || !ms3) return fail(SRUKF_EINVAL, "srukf_get_kernel_times: null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  prof_collect(h);
  for (int i = 0; i < 3; ++i) { ms3[i] = h->prof_ms[i]; if (launches3) launches3[i] = h->prof_n[i]; }
  return SRUKF_OK;
}

int srukf_get_phase_cycles(srukf_t* h, uint64_t* out8) {
  if (!h || !out8) return fail(SRUKF_EINVAL, "srukf_get_phase_cycles: null argument");
  if (!h->dbg) return fail(SRUKF_ESTATE, "srukf_get_phase_cycles: create the handle with SRUKF_PHASE_TIMING=1");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(out8, h->dbg, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SRUKF_OK;
}

int srukf_stream(srukf_t* h, uint64_t* stream) {
  if (!h || !stream) return fail(SRUKF_EINVAL, "srukf_stream: null argument");
  *stream = (uint64_t)(uintptr_t)h->stream;
  return SRUKF_OK;
}

int srukf_launch_count(srukf_t* h, uint64_t* count) {
  if (!h || !count) return fail(SRUKF_EINVAL, "srukf_launch_count: null argument");
  *count = h->launches;
  return SRUKF_OK;
}

}  // extern "C"
