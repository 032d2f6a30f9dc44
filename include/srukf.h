/*
 * srukf.h -- C ABI of the B200-native batched SRUKF predict/update (libsrukf_b200.so).
 *
 * The reference (junliu111/CV-MonoSLAM) has no plugin/FFI layer: its boundary for this path is the
 * C++ class CSLAM (MonoSLAM/SLAM.h:118-398) whose methods communicate through public members.
 * Each entry point below names the CSLAM method / member it replaces.  A reference maintainer binds
 * them as shown in INTEGRATION.md (include/SLAM.h is the ready-made C++ facade).
 *
 * Conventions
 *   - plain pointers and sizes only; all host arrays are caller-owned, C-contiguous doubles
 *   - B independent filters, L landmarks each, n = 6L+4, state order
 *       x = [f_0(6) .. f_{L-1}(6) | rx ry rz rtheta]            (SLAM.cpp:1659,2427-2432,1492-1523)
 *   - S is the upper-triangular factor with P = S^T S (SLAM.cpp:2118), exchanged PACKED row-major:
 *       row i holds columns i..n-1 at offset i*n - i*(i-1)/2 ; ntri = n(n+1)/2 doubles per filter
 *   - every function returns 0 on success or a negative SRUKF_E* code; nothing aborts or prints
 *   - a handle is bound to one device and one stream; calls on a handle are ordered
 *   - *_dev variants take DEVICE pointers (same layouts) and enqueue without host copies
 */
#ifndef SRUKF_B200_H
#define SRUKF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRUKF_OK 0
#define SRUKF_EINVAL (-1)   /* bad argument */
#define SRUKF_ECUDA (-2)    /* CUDA runtime error (srukf_last_error gives the string) */
#define SRUKF_ENOMEM (-3)   /* device allocation failed */
#define SRUKF_ESTATE (-4)   /* call order violated (e.g. update before predict_measurement) */
#define SRUKF_ENODEV (-5)   /* no CUDA device: there is no CPU fallback */

/* per-filter flag bits (srukf_get_flags) */
#define SRUKF_FLAG_NAN 1u            /* non-finite value seen in x or S */
#define SRUKF_FLAG_GMW_FLOOR 2u      /* GMW pivot raised to the EPSILON floor (expected, rank-deficient prior) */
#define SRUKF_FLAG_GMW_MODIFIED 4u   /* GMW pivot modified beyond the floor: one-shot downdate not proven equal */
#define SRUKF_FLAG_OUT_OF_VIEW 8u    /* some sigma-point projection left the image and was zeroed (SLAM.cpp:3341-3345) */
#define SRUKF_FLAG_INVISIBLE 16u     /* some feature had a zero predicted pixel (SLAM.cpp:1727) */
#define SRUKF_FLAG_FALLBACK 32u      /* the fused update was redone in reference order (sequential GMW) for this filter */

/* Replaces the scalar members CSLAM::cam_*, a1..a4, m_sigmaMeasure, m_weightType, m_sample.Alpha/Beta,
 * EPSILON, imageWidth/imageHeight (SLAM.h:148,203-204,206,240,261,293-301; defaults SLAM.cpp:164-343). */
typedef struct SrukfParams {
  double cam_dx, cam_dy, cam_cx, cam_cy, cam_k1, cam_k2, cam_f;
  int32_t image_width, image_height;
  double a1, a2, a3, a4;
  double sigma_measure;  /* Qt = I2 * sigma_measure (SLAM.cpp:238) */
  int32_t weight_type;   /* 0,1,2 = FLAG_4_WEIGHT1..3 (SLAM.cpp:1062-1102) */
  double alpha, beta;    /* m_sample.Alpha / Beta (SLAM.cpp:263-264), weight type 1 only */
  double epsilon;        /* EPSILON (SLAM.cpp:52) */
  int32_t newton_iters;  /* cap of the distortion Newton loop (SLAM.cpp:3186: 100); exits early once converged */
  int32_t downdate_mode; /* 0 = fused blocked GMW of S^T S - U U^T on the FP64 tensor pipe, with a device-side guard
                                that re-runs flagged filters in reference order (default);
                            1 = sequential per-column GMW exactly as SLAM.cpp:2116-2153 for every filter;
                            2 = one unblocked GMW of S^T S - U U^T for every filter */
} SrukfParams;

typedef struct srukf_handle srukf_t;

/* defaults of CSLAM::initializeParameters (SLAM.cpp:158-343) */
void srukf_default_params(SrukfParams *p);

/* CSLAM::CSLAM / ~CSLAM (SLAM.cpp:21-78) for B filters on CUDA device `device`. */
int srukf_create(int device, int B, int L, const SrukfParams *params, srukf_t **out);
int srukf_destroy(srukf_t *h);

/* m_X_k, m_S_k (SLAM.h:271-272).  x: [B][n]; S_packed: [B][n(n+1)/2]. */
int srukf_set_state(srukf_t *h, const double *x, const double *S_packed);
int srukf_get_state(srukf_t *h, double *x, double *S_packed);
/* dense helpers: S as [B][n][n] row-major (lower part ignored on input, zero on output) */
int srukf_set_state_dense(srukf_t *h, const double *x, const double *S_dense);
int srukf_get_state_dense(srukf_t *h, double *x, double *S_dense);

/* CSLAM::predictMotion, motion part (SLAM.cpp:1430-1465): u is [B][3] = Ut (rot1, trans, rot2). */
int srukf_predict_motion(srukf_t *h, const double *u);
/* CSLAM::predictMeasurement (SLAM.cpp:1604-1608). */
int srukf_predict_measurement(srukf_t *h);
/* m_allPredictSet / map_p->predictLocation, map_p->Si, map_p->isVisible (SLAM.cpp:1724-1738):
 * hbar [B][L][2], si [B][L][4] (2x2 row-major upper triangular), visible [B][L]; any may be NULL.
 * Sign convention: si is the Cholesky factor of the 2x2 innovation Gram matrix (positive diagonal).  The reference's
 * GSL Householder QR (:1771-1775) returns the same matrix up to the sign of each row (R(0,0) = -sign(alpha) |x|,
 * typically negative); Si^T Si, the gain, U and the chi-square gate are independent of it.  A caller that derives a
 * search window from 2*Si(0,0) (dataAssociation, :1952-1955) should use |Si(k,k)|. */
int srukf_get_prediction(srukf_t *h, double *hbar, double *si, uint8_t *visible);
/* Feature initialisation at frame 1: CSLAM::addFeatures with an empty map (SLAM.cpp:818-871), i.e.
 * passSigmaThroughMapingFunction (:1177-1250), QrAndCholeskyForInitilization (:1260-1300) and getPermutationMatrix
 * (:1303-1334).  For every filter: robot prior x4 [B][4] = (x, y, z, theta) and its factor S4 [B][4][4] (dense rows,
 * :851-857), L key-points in distorted pixels keypoints [B][L][2] = (pt.x, pt.y) (:859-861), inverse-depth prior rho0
 * +- sigma_rho (:172-173); the pixel sigma is SrukfParams.sigma_measure.  Replaces the whole state (m_X_k, m_S_k) of
 * the handle, in canonical order [f_0(6) .. f_{L-1}(6) | robot(4)]; flags are reset (SRUKF_FLAG_GMW_FLOOR is expected:
 * the anchors make the covariance rank 4 + 3L). */
int srukf_init_features(srukf_t *h, const double *x4, const double *S4, const double *keypoints, double rho0,
                        double sigma_rho);
/* CSLAM::integrateFeaturesInformation on a NON-empty map (SLAM.cpp:818-871 with dim > 4): M = dst.L - src.L key-points
 * (keypoints [B][M][2], distorted pixels) are appended to every filter of src; the augmented state goes to dst (same B,
 * same device) in canonical order [old features | new features | robot].  Follow it with
 * srukf_kalman_update_reorder(dst, ..., M) on the next frame, as the reference does.  src is left untouched. */
int srukf_add_features(srukf_t *src, srukf_t *dst, const double *keypoints, double rho0, double sigma_rho);
/* CSLAM::KalmanUpdate on the frame after features were added (m_nAddings != 0, SLAM.cpp:2083-2086): the covariance
 * downdate takes GSLCholeskyUpdate's NEED_REORDER branch (:2122-2138) with CholeskyDecompositionWithPivoting
 * (:2158-2179) for every U column.  n_new = m_nFilters, the number of features (the last ones of the state) added on the
 * previous frame -- L on the frame after srukf_init_features; 0 behaves as srukf_kalman_update.  Reference order, one
 * CTA per filter (not the fused kernel): meant for the one frame that follows an addition.  The covariance is carried
 * across the U columns and factorised once at the end (the reference's QR between columns adds nothing to it). */
int srukf_kalman_update_reorder(srukf_t *h, const double *z, const uint8_t *matched, int n_new);
/* CSLAM::deleteOneFeature (SLAM.cpp:2637-2663), state and factor part: filter b of src drops feature ids[b] (0-based
 * position in the state); the reduced state is written to dst, a handle created for the same B, L-1 features and the
 * same device (the state dimension of a handle is fixed).  The factor is rebuilt as the reference does it: the dropped
 * rows V of S re-enter through GSLCholeskyUpdate(V^T, UPDATING, NEEDNOT_REORDER) (:2661-2662, :2139-2153), six
 * modified-Cholesky re-factorisations of S^T S + v v^T.  Flags of src carry over.  src is left untouched. */
int srukf_delete_feature(srukf_t *src, srukf_t *dst, const int32_t *ids);
/* Chi-square gate of CSLAM::dataAssociation (SLAM.cpp:1946-1977, CHI2INV_TABLE(0,2) = 5.99146454710798 at :54):
 * candidate pixels z [B][L][2] are accepted when (z - predictLocation) (Si^T Si)^-1 (z - predictLocation)^T < threshold
 * and the feature is visible.  accept [B][L] (the isMatching mask for srukf_kalman_update), d2 [B][L] or NULL.
 * Call between srukf_predict_measurement and srukf_kalman_update.  (The image-patch correlation that proposes the
 * candidates in the reference is outside this path.) */
int srukf_chi2_gate(srukf_t *h, const double *z, double threshold, uint8_t *accept, double *d2);
/* CSLAM::KalmanUpdate (SLAM.cpp:2048-2096): z [B][L][2] = matchLocation (x,y); matched [B][L] = isMatching. */
int srukf_kalman_update(srukf_t *h, const double *z, const uint8_t *matched);
/* predictMotion + predictMeasurement + KalmanUpdate of one CSLAM::SLAM() frame (SLAM.cpp:91,93,99). */
int srukf_step(srukf_t *h, const double *u, const double *z, const uint8_t *matched);

/* Read-back of m_X_k that overlaps the next frame: x_host [B][n] (pinned memory for a truly asynchronous copy) is
 * filled from a device-side snapshot taken in stream order after everything queued so far; it is complete after
 * srukf_sync().  srukf_step itself copies its inputs on a second stream into alternating device buffers, so the inputs
 * of frame s+1 travel while frame s computes; u / z / matched must stay unmodified until the next srukf_sync() or the
 * second-next srukf_step() of the handle. */
int srukf_get_x_async(srukf_t *h, double *x_host);

/* ---- CSLAM helper methods as stand-alone entry points (the facade include/SLAM.h binds them) ---- */
#define SRUKF_UPDATING 0          /* FLAG_4_UPDATING,        SLAM.cpp:31 */
#define SRUKF_DOWNDATING 1        /* FLAG_4_DOWNDATING,      SLAM.cpp:32 */
#define SRUKF_NEED_REORDER 0      /* FLAG_4_NEED_REORDER,    SLAM.cpp:36 */
#define SRUKF_NEEDNOT_REORDER 1   /* FLAG_4_NEEDNOT_REORDER, SLAM.cpp:37 */
/* CSLAM::modifiedCholeskyDecomposition(Mat &sr, const Mat &Cov) (SLAM.h:355, SLAM.cpp:2197-2327) for nb matrices:
 * G [nb][n][n] dense row-major (lower triangle factorised, maxima over the whole matrix) -> S [nb][n][n] upper
 * triangular, S = sqrt(D) L^T; flags [nb] (SRUKF_FLAG_GMW_* / NAN) or NULL. */
int srukf_mchol(int device, int nb, int n, double epsilon, const double *G, double *S, uint32_t *flags);
/* CSLAM::GSLQrDecomposition(Mat &R, const Mat &A) const (SLAM.h:348, SLAM.cpp:2330-2353): triu(R) of the Householder
 * QR in GSL's sign convention; A [nb][m][n] (m >= n), R [nb][n][n]. */
int srukf_qr_R(int device, int nb, int m, int n, const double *A, double *R);
/* CSLAM::generateSigmaPoints(Mat &sigma, const Mat &mu, const Mat &sr) (SLAM.h:341, SLAM.cpp:1148-1162):
 * mu [nb][Na], sr [nb][Na][Na] (rows are the directions) -> sigma [nb][Na][2Na+1].  The step kernels never materialise
 * this matrix; the entry exists for callers of the reference helper. */
int srukf_generate_sigma_points(int device, int nb, int Na, double gamma, const double *mu, const double *sr,
                                double *sigma);
/* CSLAM::GSLCholeskyUpdate(const Mat &u, const int &flag4UpOrDown, const int &flag4Order) (SLAM.h:347,
 * SLAM.cpp:2106-2155) on the handle's m_S_k: U [B][n][k] (u is dim x k), applied column by column in reference order:
 * S <- modifiedCholesky(S^T S +- u u^T); with SRUKF_NEED_REORDER the last n_new features are the ones added on the
 * previous frame (:2122-2138; n_new is ignored otherwise).  Synchronous. */
int srukf_cholesky_update(srukf_t *h, const double *U, int k, int up_or_down, int order, int n_new);
/* Diagnostics: FP64 tensor-pipe throughput (back-to-back DMMA m8n8k4) of `device` in TFLOP/s, measured now (bench.py's
 * roofline denominator, taken in the same process and clock state as the timed run). */
int srukf_fp64_peak(int device, double *tflops);

/* device-pointer variants (inputs already resident in HBM; asynchronous on the handle's stream).  The handle's stream is
 * non-blocking: it is NOT ordered after the caller's streams, so the buffers must be complete (synchronise or wait on an
 * event of the stream that produced them) before the call, and must stay valid until srukf_sync(). */
int srukf_step_dev(srukf_t *h, const double *d_u, const double *d_z, const uint8_t *d_matched);
/* device-to-device load of filters [b0, b0+nb): d_x [nb][n], d_S_packed [nb][n(n+1)/2] (either may be NULL).
 * (The library keeps S in its own blocked layout in HBM; all exchanged formats are the ones of this header.)
 * Returns after the copy has completed; the same readiness rule for d_x / d_S_packed applies on entry. */
int srukf_set_state_dev(srukf_t *h, int b0, int nb, const double *d_x, const double *d_S_packed);

/* m_P_k block (SLAM.cpp:2404): P[r0:r0+nr, r0:r0+nr] of S^T S per filter, out [B][nr][nr]. */
int srukf_get_cov_block(srukf_t *h, int r0, int nr, double *out);
int srukf_get_flags(srukf_t *h, uint32_t *flags /* [B] */);
int srukf_clear_flags(srukf_t *h);

/* Monte-Carlo statistics (no reference equivalent): truth [B][3] = true (rx, ry, rtheta).
 * out[8] = { sum ex^2, sum ey^2, sum etheta^2, sum NEES(x,y,theta), count, #flag NAN, #flag GMW_MODIFIED, 0 }
 * These partial sums are what the multi-GPU driver all-reduces. */
int srukf_stats(srukf_t *h, const double *truth, double *out8);

int srukf_sync(srukf_t *h);
/* cudaStream_t of the handle as an integer (for event timing on the launching stream) */
int srukf_stream(srukf_t *h, uint64_t *stream);
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int srukf_launch_count(srukf_t *h, uint64_t *count);
/* Optional per-kernel timing with CUDA events on the handle's stream (bench.py's roofline leg).
 * ms[3] = cumulative milliseconds of {k_predict, k_gain, k_update (+ fallback)} launches, launches[3] their counts,
 * both since profiling was last switched on.  srukf_get_kernel_times synchronises the stream. */
int srukf_set_profiling(srukf_t *h, int on);
int srukf_get_kernel_times(srukf_t *h, double *ms3, uint64_t *launches3);
/* Diagnostics: cumulative SM cycles spent by k_update's CTAs in {K loop, post-K barrier, panel store, diagonal
 * block, solve, end barrier, -, #CTAs} and, in out[8..13], inside the K loop {stage acquire, copy issue, data wait,
 * DMMA, release, #chunks}; only when the handle was created with SRUKF_PHASE_TIMING=1 in the env.  out has 16 slots. */
int srukf_get_phase_cycles(srukf_t *h, uint64_t *out16);
const char *srukf_last_error(void);
const char *srukf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SRUKF_B200_H */
