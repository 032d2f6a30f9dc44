/*
 * srukf_oracle.h -- CPU restatement of CV-MonoSLAM's SRUKF predict/update path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * PARITY PINNED TO THE REFERENCE'S TEXT (except at the GSL boundary).  The reference (MonoSLAM/SLAM.cpp, Win32/MFC +
 * OpenCV 2.4.3 + GSL 1.8) ships no tests, golden vectors or recorded data and its own build cannot run here, but the
 * bodies of its functions on this path compile unchanged against a small cv::Mat / MFC stand-in: oracle/ref_shim/
 * extracts them verbatim from /root/reference at build time into oracle/_ref/ (git-ignored) and
 * tests/test_ref_pin.py runs them beside this file on the same inputs.  Every function below reproduces the
 * reference's result BIT FOR BIT in literal mode (downdate_mode 0 / 2): sample parameters, the modified Cholesky,
 * the camera chain, whole predict / measure / update frames up to L = 50, the NEED_REORDER update, feature
 * initialisation, augmentation and deletion.  tests/golden/*.npz are outputs of that reference build.
 * What stays a restatement: gsl_linalg_QR_decomp (GSL is not vendored by the reference; the reference build calls
 * oracle_qr_decomp below, which is cross-checked against LAPACK dgeqrf and mpmath in tests/test_oracle.py) and the
 * OpenCV primitives inside the stand-in (checked against cv2 in tests/test_ref_shim.py).
 * This file restates the arithmetic line by line (citations are MonoSLAM/SLAM.cpp:line unless noted).
 *
 * Third-party arithmetic restated (not vendored by the reference):
 *   - GSL 1.8 gsl_linalg_QR_decomp (unblocked Householder, linalg/qr.c + householder.c), called at
 *     SLAM.cpp:2339.
 *   - OpenCV 2.4.3 addWeighted / Mat::inv (<=3x3 closed form) / divide / minMaxLoc / Mat products.
 *
 * Layout: all matrices dense row-major doubles.  State x = [f_0(6) .. f_{L-1}(6) | rx ry rz rtheta],
 * n = 6L+4 (SLAM.cpp:1659,2427-2432,1492-1523).  S is n x n upper triangular with P = S^T S
 * (SLAM.cpp:2118).
 */
#ifndef SRUKF_ORACLE_H
#define SRUKF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleParams {
  /* camera, SLAM.cpp:329-337 */
  double cam_dx, cam_dy, cam_cx, cam_cy, cam_k1, cam_k2, cam_f;
  int image_width, image_height; /* SLAM.cpp:312-313 (read from the first frame) */
  /* odometry noise, SLAM.cpp:195-198 */
  double a1, a2, a3, a4;
  /* Qt = I2 * sigma_measure, SLAM.cpp:189,238 */
  double sigma_measure;
  /* weights, SLAM.cpp:241,263-264 */
  int weight_type; /* 0,1,2 = FLAG_4_WEIGHT1..3 */
  double alpha, beta;
  /* EPSILON, SLAM.cpp:52 */
  double epsilon;
  /* distortion Newton iterations, SLAM.cpp:3186 */
  int newton_iters;
  /* covariance downdate mode:
   *   0 literal : P = S^T S re-formed per U column, then GMW (SLAM.cpp:2116-2153)
   *   1 carry-P : same sequence, but P carried as G+E between columns (algebraically S^T S)
   *   2 literal with the dense (non-triangular-aware) n^3 product, for CPU-baseline timing only */
  int downdate_mode;
} OracleParams;

typedef struct OracleWeights {
  double gamma, wm0, wm0_sr, wc0, wc0_sr, wi, wi_sr;
} OracleWeights;

void oracle_default_params(OracleParams *p);

/* SLAM.cpp:1050-1103 */
void oracle_sample_parameters(int Na, const OracleParams *p, OracleWeights *w);

/* GSL 1.8 gsl_linalg_QR_decomp semantics; A is m x n row-major and is overwritten (R in the upper
 * triangle, Householder vectors below).  tau has min(m,n) entries.  SLAM.cpp:2330-2353 keeps
 * triu(R) only: oracle_qr_R writes the n x n upper-triangular R (requires m >= n). */
void oracle_qr_decomp(double *A, int m, int n, double *tau);
void oracle_qr_R(const double *A, int m, int n, double *R);

/* SLAM.cpp:2197-2327.  G n x n symmetric (full storage).  S receives sqrt(D) L^T (upper triangular,
 * dense n x n).  E (may be NULL) receives the n diagonal modifications D_j - C_jj.  Returns the number
 * of pivots whose D_j != C_jj. */
int oracle_mchol(const double *G, int n, double epsilon, double *S, double *E);

/* camera chain, SLAM.cpp:3177-3213, 3224-3236, 3250-3276, 3289-3292, 3324-3347, 3358-3420 */
void oracle_distort(const OracleParams *p, double uvu_x, double uvu_y, double *uvd_x, double *uvd_y);
void oracle_undistort(const OracleParams *p, double uvd_x, double uvd_y, double *uvu_x, double *uvu_y);
/* one feature (6-vector), robot position (3), robot heading, pixel noise (2) -> distorted pixel */
void oracle_project(const OracleParams *p, const double *feat6, const double *pos3, double theta,
                    const double *err2, double *uvd_x, double *uvd_y);

/* SLAM.cpp:1446-1450: odometry poses (x,y,theta) at k-1 and k -> Ut = (rot1, trans, rot2) */
void oracle_odometry_to_control(const double *odo_prev3, const double *odo_now3, double *u3);

/* Workspace for one filter (materialised sigma matrices as in the reference). */
typedef struct OracleFilter {
  int L, n, Na, P;
  OracleParams prm;
  OracleWeights w;
  double *x;        /* n            m_X_k   */
  double *S;        /* n x n        m_S_k   */
  double *sigma;    /* Na x P       m_sigma */
  double *pix;      /* 2L x P       m_sigma_allPixel */
  double *hbar;     /* 2L           m_allPredictSet  */
  double *si;       /* L x 4        map_p->Si        */
  unsigned char *visible; /* L      map_p->isVisible (this frame) */
  double Mt[3], Ut[3];
  int n_new;        /* m_nAddings: the last n_new features were added on the previous frame */
  /* diagnostics */
  long n_mchol_calls, n_mchol_modified;
  double max_E; /* largest diagonal modification seen in the last update */
  /* scratch */
  double *work_qr, *work_P, *work_U;
} OracleFilter;

OracleFilter *oracle_filter_create(int L, const OracleParams *p);
void oracle_filter_destroy(OracleFilter *f);
void oracle_filter_set_state(OracleFilter *f, const double *x, const double *S);
void oracle_filter_get_state(const OracleFilter *f, double *x, double *S);

/* SLAM.cpp:1430-1465 (+1476-1532, 1539-1556) */
void oracle_predict_motion(OracleFilter *f, const double *u3);
/* SLAM.cpp:1604-1608 (+1615-1682, 1700-1775) */
void oracle_predict_measurement(OracleFilter *f);
/* SLAM.cpp:2048-2096 (+2020-2038, 2106-2153); z is L x 2 (matchLocation.x, .y), matched L flags */
void oracle_kalman_update(OracleFilter *f, const double *z, const unsigned char *matched);
void oracle_filter_set_new_features(OracleFilter *f, int n_new);
void oracle_filter_get_prediction(const OracleFilter *f, double *hbar, double *si, unsigned char *visible);
/* SLAM.cpp:1946-1977 chi-square gate of candidate pixels z (L x 2) against the current prediction */
void oracle_chi2_gate(const OracleFilter *f, const double *z, double threshold, unsigned char *accept, double *d2);
/* one full frame of the hot path */
void oracle_step(OracleFilter *f, const double *u3, const double *z, const unsigned char *matched);

/* Feature initialisation at frame 1 (SLAM.cpp:818-871, 1177-1334): robot prior x4 / S4 (4x4 upper),
 * M key-points (distorted pixels, kp[2*i]=pt.x, kp[2*i+1]=pt.y), rho0, sigma_rho.
 * Outputs x (6M+4) and S ((6M+4)^2) in canonical order. */
void oracle_init_features(const OracleParams *p, const double *x4, const double *S4, int M,
                          const double *kp, double rho0, double sigma_rho, double *x_out, double *S_out);

/* deleteOneFeature (SLAM.cpp:2637-2663) + GSLCholeskyUpdate(UPDATING, NEEDNOT_REORDER) (:2139-2153) */
void oracle_delete_feature(const OracleParams *p, int L, const double *x, const double *S, int id, double *x_out,
                           double *S_out);

/* integrateFeaturesInformation for a non-empty map (SLAM.cpp:818-871 with dim = 6 Lold + 4 > 4): M key-points are
 * appended to the Lold-feature state; outputs in canonical order [old features | new features | robot]. */
void oracle_add_features(const OracleParams *p, int Lold, const double *x, const double *S, int M, const double *kp,
                         double rho0, double sigma_rho, double *x_out, double *S_out);

/* batch helpers used by tests / bench (OpenMP over filters when available) */
void oracle_batch_step(int B, int L, const OracleParams *p, double *x /*B x n*/, double *S /*B x n x n*/,
                       const double *u /*B x 3*/, const double *z /*B x L x 2*/,
                       const unsigned char *matched /*B x L*/, int steps_stride_u, int nsteps,
                       int nthreads, double *max_E_out /*B or NULL*/);

#ifdef __cplusplus
}
#endif
#endif
