/*
 * srukf_oracle.c -- CPU restatement of CV-MonoSLAM's SRUKF predict/update (see srukf_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference tests / golden vectors exist, the
 * reference cannot be built here).  Every function cites the MonoSLAM/SLAM.cpp lines it follows.
 * The arithmetic keeps the reference's operation order where the order is visible in the source
 * (addWeighted accumulations, sequential per-feature update, GMW column sweeps).
 */
#include "srukf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------ */
/* parameters                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* SLAM.cpp:164-214 (debug-model defaults), :221-224, :238-242, :263-264, :329-337, :52 */
void oracle_default_params(OracleParams *p) {
  p->cam_dx = 0.0028;
  p->cam_dy = 0.0028;
  p->cam_cx = 310.1129;
  p->cam_cy = 236.7526;
  p->cam_k1 = 0.0001;
  p->cam_k2 = 0.0000;
  p->cam_f = 2.1735;
  p->image_width = 640; /* the reference reads these from the first frame (:312-313) */
  p->image_height = 480;
  p->a1 = 8;
  p->a2 = 8;
  p->a3 = 8;
  p->a4 = 8;
  p->sigma_measure = 3.0;
  p->weight_type = 0;
  p->alpha = 1e-3;
  p->beta = 2;
  p->epsilon = 1e-13;
  p->newton_iters = 100;
  p->downdate_mode = 0;
}

/* SLAM.cpp:1050-1103 */
void oracle_sample_parameters(int Na, const OracleParams *p, OracleWeights *w) {
  /* :1052-1057 (m_sample.*; Kappa = 0) */
  double kappa = 0;
  double lambda = pow(p->alpha, 2) * (Na + kappa) - Na;
  double s_gamma = sqrt(Na + lambda);
  double s_wm0 = lambda / (Na + lambda);
  double s_wc0 = s_wm0 + (1 - pow(p->alpha, 2) + p->beta);
  double s_wi = 1.0 / (2 * (Na + lambda));
  switch (p->weight_type) {
    case 0: /* :1064-1075 */
      w->wm0 = 1.0 - Na / 3.0;
      w->wm0_sr = sqrt(fabs(w->wm0));
      w->wc0 = 1.0 - Na / 3.0;
      w->wc0_sr = sqrt(fabs(w->wm0));
      w->wi = (1.0 - w->wc0) / (2 * Na);
      w->wi_sr = sqrt(w->wi);
      w->gamma = sqrt(Na / (1.0 - w->wm0));
      break;
    case 1: /* :1077-1088 */
      w->gamma = s_gamma;
      w->wm0 = s_wm0;
      w->wm0_sr = sqrt(fabs(s_wm0));
      w->wc0 = s_wc0;
      w->wc0_sr = sqrt(fabs(s_wc0));
      w->wi = s_wi;
      w->wi_sr = sqrt(fabs(s_wi));
      break;
    default: /* :1090-1101 */
      w->gamma = sqrt(3.0 * Na / 2.0);
      w->wm0 = 1.0 / 3.0;
      w->wm0_sr = sqrt(w->wm0);
      w->wc0 = 1.0 / 3.0;
      w->wc0_sr = sqrt(w->wc0);
      w->wi = 1.0 / (3.0 * Na);
      w->wi_sr = sqrt(w->wi);
      break;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* GSL 1.8 Householder QR (called at SLAM.cpp:2339)                                            */
/* ------------------------------------------------------------------------------------------ */

/* reference-BLAS dnrm2 (scaled sum of squares), as GSL's cblas does */
static double nrm2_strided(const double *x, int n, int inc) {
  double scale = 0.0, ssq = 1.0;
  if (n <= 0) return 0.0;
  if (n == 1) return fabs(x[0]);
  for (int i = 0; i < n; i++) {
    double v = x[(size_t)i * inc];
    if (v != 0.0) {
      double ax = fabs(v);
      if (scale < ax) {
        ssq = 1.0 + ssq * (scale / ax) * (scale / ax);
        scale = ax;
      } else {
        ssq += (ax / scale) * (ax / scale);
      }
    }
  }
  return scale * sqrt(ssq);
}

/* gsl_linalg_householder_transform on the strided vector v[0..len) (GSL 1.8 linalg/householder.c) */
static double householder_transform(double *v, int len, int inc) {
  if (len == 1) return 0.0;
  double xnorm = nrm2_strided(v + inc, len - 1, inc);
  if (xnorm == 0) return 0.0;
  double alpha = v[0];
  double beta = -(alpha >= 0.0 ? +1.0 : -1.0) * hypot(alpha, xnorm);
  double tau = (beta - alpha) / beta;
  double sc = 1.0 / (alpha - beta);
  for (int i = 1; i < len; i++) v[(size_t)i * inc] *= sc;
  v[0] = beta;
  return tau;
}

/* gsl_linalg_QR_decomp: for each column, transform then gsl_linalg_householder_hm on the trailing
 * block, column by column (w_j = A_0j + sum_i A_ij v_i; A_ij -= tau v_i w_j). */
void oracle_qr_decomp(double *A, int m, int n, double *tau) {
  int kmax = m < n ? m : n;
  for (int i = 0; i < kmax; i++) {
    double *c = A + (size_t)i * n + i; /* column i from row i, stride n */
    double t = householder_transform(c, m - i, n);
    tau[i] = t;
    if (i + 1 < n && t != 0.0) {
      int rows = m - i;
      for (int j = i + 1; j < n; j++) {
        double *a = A + (size_t)i * n + j;
        double wj = a[0];
        for (int r = 1; r < rows; r++) wj += a[(size_t)r * n] * c[(size_t)r * n];
        a[0] = a[0] - t * wj;
        for (int r = 1; r < rows; r++) {
          double vi = c[(size_t)r * n];
          a[(size_t)r * n] = a[(size_t)r * n] - t * vi * wj;
        }
      }
    }
  }
}

/* SLAM.cpp:2330-2353: copy in, decompose, keep triu */
void oracle_qr_R(const double *A, int m, int n, double *R) {
  double *W = (double *)malloc(sizeof(double) * (size_t)m * n);
  double *tau = (double *)malloc(sizeof(double) * (size_t)n);
  memcpy(W, A, sizeof(double) * (size_t)m * n);
  oracle_qr_decomp(W, m, n, tau);
  memset(R, 0, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n && i < m; i++)
    for (int j = i; j < n; j++) R[(size_t)i * n + j] = W[(size_t)i * n + j];
  free(tau);
  free(W);
}

/* ------------------------------------------------------------------------------------------ */
/* Gill-Murray-Wright modified Cholesky, SLAM.cpp:2197-2327                                    */
/* ------------------------------------------------------------------------------------------ */
int oracle_mchol(const double *G, int n, double epsilon, double *S, double *E) {
  /* :2204 gamma = max diag; :2205 zi = max (signed) entry of G with its diagonal zeroed */
  double gamma = G[0], zi = 0.0; /* the zeroed diagonal contributes 0 to the max when n >= 1 */
  for (int i = 0; i < n; i++) {
    if (G[(size_t)i * n + i] > gamma) gamma = G[(size_t)i * n + i];
    for (int j = 0; j < n; j++)
      if (i != j && G[(size_t)i * n + j] > zi) zi = G[(size_t)i * n + j];
  }
  /* :2207-2211 */
  double nu = sqrt((double)n * n - 1.0);
  if (nu < 1.0) nu = 1.0;
  double beta2 = gamma;
  if (zi / nu > beta2) beta2 = zi / nu;
  if (1e-15 > beta2) beta2 = 1e-15;

  /* C holds the working columns (lower triangle + diagonal), D the pivots. */
  double *C = (double *)calloc((size_t)n * n, sizeof(double));
  double *Lm = (double *)calloc((size_t)n * n, sizeof(double));
  double *D = (double *)calloc((size_t)n, sizeof(double));
  for (int i = 0; i < n; i++) C[(size_t)i * n + i] = G[(size_t)i * n + i]; /* :2217 */
  int modified = 0;

  for (int j = 0; j < n; j++) {
    /* :2224-2234  L(j,0:j) = C(j,0:j) ./ D(0:j)   (cv::divide gives 0 on a zero divisor) */
    for (int k = 0; k < j; k++) Lm[(size_t)j * n + k] = (D[k] != 0.0) ? C[(size_t)j * n + k] / D[k] : 0.0;
    /* :2237-2261  C(j+1:n, j) = G(j+1:n, j) - L(j,0:j) * C(j+1:n,0:j)^T */
    for (int i = j + 1; i < n; i++) {
      double acc = 0.0;
      for (int k = 0; k < j; k++) acc += Lm[(size_t)j * n + k] * C[(size_t)i * n + k];
      C[(size_t)i * n + j] = G[(size_t)i * n + j] - acc;
    }
    /* :2264-2276 theta_j = max_{i>j} |C(i,j)| */
    double theta = 0.0;
    for (int i = j + 1; i < n; i++) {
      double a = fabs(C[(size_t)i * n + j]);
      if (a > theta) theta = a;
    }
    /* :2279-2285 D_j = max(EPSILON, |C_jj|, theta^2/beta2) */
    double cjj = C[(size_t)j * n + j];
    double d = epsilon;
    if (fabs(cjj) > d) d = fabs(cjj);
    if (theta * theta / beta2 > d) d = theta * theta / beta2;
    D[j] = d;
    /* :2288 */
    if (E) E[j] = d - cjj;
    if (d != cjj) modified++;
    /* :2291-2295 */
    for (int i = j + 1; i < n; i++)
      C[(size_t)i * n + i] = C[(size_t)i * n + i] - C[(size_t)i * n + j] * C[(size_t)i * n + j] / d;
  }
  /* :2299-2302 unit diagonal; :2319-2321 S = sqrt(D) * L^T.  (:2304-2317 pneg is computed and
   * discarded by the reference; it has no effect on the state.) */
  memset(S, 0, sizeof(double) * (size_t)n * n);
  for (int j = 0; j < n; j++) {
    double sd = sqrt(D[j]);
    S[(size_t)j * n + j] = sd * 1.0;
    for (int i = j + 1; i < n; i++) S[(size_t)j * n + i] = sd * Lm[(size_t)i * n + j];
  }
  free(D);
  free(Lm);
  free(C);
  return modified;
}

/* ------------------------------------------------------------------------------------------ */
/* camera chain                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* SLAM.cpp:3177-3213 */
void oracle_distort(const OracleParams *p, double uvu_x, double uvu_y, double *uvd_x, double *uvd_y) {
  double f, ff;
  double xu = (uvu_x - p->cam_cx) * p->cam_dx;
  double yu = (uvu_y - p->cam_cy) * p->cam_dy;
  double ru = sqrt(xu * xu + yu * yu);
  double rd = ru / (1 + p->cam_k1 * ru * ru + p->cam_k2 * pow(ru, 4));
  for (int i = 0; i < p->newton_iters; i++) {
    f = rd + p->cam_k1 * pow(rd, 3) + p->cam_k2 * pow(rd, 5) - ru;
    ff = 1.0 + 3.0 * p->cam_k1 * rd * rd + 5.0 * p->cam_k2 * pow(rd, 4);
    rd = rd - f / ff;
  }
  double d = 1 + p->cam_k1 * rd * rd + p->cam_k2 * pow(rd, 4);
  if (d == 0) d = p->epsilon;
  double xd = xu / d;
  double yd = yu / d;
  double ox = p->cam_cx + xd / p->cam_dx;
  double oy = p->cam_cy + yd / p->cam_dy;
  int vis = (ox >= 0) && (ox <= p->image_width) && (oy >= 0) && (oy <= p->image_height);
  if (!vis) {
    ox = 0;
    oy = 0;
  }
  *uvd_x = ox;
  *uvd_y = oy;
}

/* SLAM.cpp:3224-3236 */
void oracle_undistort(const OracleParams *p, double uvd_x, double uvd_y, double *uvu_x, double *uvu_y) {
  double xd = (uvd_x - p->cam_cx) * p->cam_dx;
  double yd = (uvd_y - p->cam_cy) * p->cam_dy;
  double rd = sqrt(xd * xd + yd * yd);
  double d = 1 + p->cam_k1 * pow(rd, 2) + p->cam_k2 * pow(rd, 4);
  double xu = xd * d;
  double yu = yd * d;
  *uvu_x = p->cam_cx + xu / p->cam_dx;
  *uvu_y = p->cam_cy + yu / p->cam_dy;
}

/* OpenCV 2.4.3 Mat::inv() on a 3x3 (DECOMP_LU, n<=3 closed form, modules/core/src/lapack.cpp):
 * d = det; if d != 0: d = 1/d; dst = adj * d.  Applied to Rwc of SLAM.cpp:1031-1037 at :1643. */
static void inv3x3(const double *s, double *t) {
  double d = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) +
             s[2] * (s[3] * s[7] - s[4] * s[6]);
  if (d != 0.) {
    d = 1. / d;
    t[0] = (s[4] * s[8] - s[5] * s[7]) * d;
    t[1] = (s[2] * s[7] - s[1] * s[8]) * d;
    t[2] = (s[1] * s[5] - s[2] * s[4]) * d;
    t[3] = (s[5] * s[6] - s[3] * s[8]) * d;
    t[4] = (s[0] * s[8] - s[2] * s[6]) * d;
    t[5] = (s[2] * s[3] - s[0] * s[5]) * d;
    t[6] = (s[3] * s[7] - s[4] * s[6]) * d;
    t[7] = (s[1] * s[6] - s[0] * s[7]) * d;
    t[8] = (s[0] * s[4] - s[1] * s[3]) * d;
  } else {
    memset(t, 0, 9 * sizeof(double));
  }
}

/* SLAM.cpp:1031-1037 */
static void transfer_matrix(double theta, double *R) {
  R[0] = cos(theta);
  R[1] = -sin(theta);
  R[2] = 0;
  R[3] = sin(theta);
  R[4] = cos(theta);
  R[5] = 0;
  R[6] = 0;
  R[7] = 0;
  R[8] = 1;
}

/* one projection given Rcw: State2World :3250-3276, World2Camera :3289-3292, Camera2Image
 * :3324-3347, distortOnePointRW :3177-3213 */
static void project_rcw(const OracleParams *p, const double *st, const double *pos, const double *Rcw,
                        const double *err, double *ox, double *oy) {
  double xi = st[0], yi = st[1], zi = st[2], theta = st[3], phi = st[4], rho = st[5];
  double Hlw[3], Hlr[3];
  Hlw[0] = xi + 1 / rho * cos(phi) * sin(theta) - pos[0];
  Hlw[1] = yi - 1 / rho * sin(phi) - pos[1];
  Hlw[2] = zi + 1 / rho * cos(phi) * cos(theta) - pos[2];
  for (int r = 0; r < 3; r++) Hlr[r] = Rcw[3 * r + 0] * Hlw[0] + Rcw[3 * r + 1] * Hlw[1] + Rcw[3 * r + 2] * Hlw[2];
  double ux, uy;
  double f1 = p->cam_f / p->cam_dx, f2 = p->cam_f / p->cam_dy; /* :336-337 */
  if (Hlr[2] == 0) {
    ux = 0;
    uy = 0;
  } else {
    uy = p->cam_cx + f1 * Hlr[0] / Hlr[2] + err[0]; /* :3338  (x/y swap is the reference's) */
    ux = p->cam_cy + f2 * Hlr[1] / Hlr[2] + err[1]; /* :3339 */
    if (ux < 10 || ux > p->image_width - 10 || uy < 10 || uy > p->image_height - 10) {
      ux = 0;
      uy = 0;
    }
  }
  oracle_distort(p, ux, uy, ox, oy);
}

void oracle_project(const OracleParams *p, const double *feat6, const double *pos3, double theta,
                    const double *err2, double *uvd_x, double *uvd_y) {
  double Rwc[9], Rcw[9];
  transfer_matrix(theta, Rwc);
  inv3x3(Rwc, Rcw);
  project_rcw(p, feat6, pos3, Rcw, err2, uvd_x, uvd_y);
}

/* SLAM.cpp:1446-1450 */
void oracle_odometry_to_control(const double *o0, const double *o1, double *u3) {
  double dx = o1[0] - o0[0];
  double dy = o1[1] - o0[1];
  double rot1 = atan2(dy, dx) - o0[2];
  double trans = sqrt(dy * dy + dx * dx);
  double rot2 = o1[2] - o0[2] - rot1;
  u3[0] = rot1;
  u3[1] = trans;
  u3[2] = rot2;
}

/* ------------------------------------------------------------------------------------------ */
/* filter object                                                                               */
/* ------------------------------------------------------------------------------------------ */

OracleFilter *oracle_filter_create(int L, const OracleParams *p) {
  OracleFilter *f = (OracleFilter *)calloc(1, sizeof(OracleFilter));
  f->L = L;
  f->n = 6 * L + 4;
  f->Na = f->n + 5; /* :1432 */
  f->P = 2 * f->Na + 1;
  f->prm = *p;
  size_t n = f->n, Na = f->Na, P = f->P;
  f->x = (double *)calloc(n, sizeof(double));
  f->S = (double *)calloc(n * n, sizeof(double));
  f->sigma = (double *)calloc(Na * P, sizeof(double));
  f->pix = (double *)calloc((size_t)2 * L * P + 1, sizeof(double));
  f->hbar = (double *)calloc((size_t)2 * L + 1, sizeof(double));
  f->si = (double *)calloc((size_t)4 * L + 1, sizeof(double));
  f->visible = (unsigned char *)calloc((size_t)L + 1, 1);
  f->work_qr = (double *)calloc((size_t)2 * Na * n, sizeof(double));
  f->work_P = (double *)calloc(n * n, sizeof(double));
  f->work_U = (double *)calloc(n * 8, sizeof(double));
  return f;
}

void oracle_filter_destroy(OracleFilter *f) {
  if (!f) return;
  free(f->x);
  free(f->S);
  free(f->sigma);
  free(f->pix);
  free(f->hbar);
  free(f->si);
  free(f->visible);
  free(f->work_qr);
  free(f->work_P);
  free(f->work_U);
  free(f);
}

void oracle_filter_set_state(OracleFilter *f, const double *x, const double *S) {
  memcpy(f->x, x, sizeof(double) * (size_t)f->n);
  memcpy(f->S, S, sizeof(double) * (size_t)f->n * f->n);
}

void oracle_filter_get_state(const OracleFilter *f, double *x, double *S) {
  memcpy(x, f->x, sizeof(double) * (size_t)f->n);
  memcpy(S, f->S, sizeof(double) * (size_t)f->n * f->n);
}

/* SLAM.cpp:1148-1162 with mu = [x; 0_5], sr = blockdiag(S, Mt, Qt) (:1461-1462, :1123-1135).
 * addWeighted(mu, 1, element, +-gamma, 0, dst) = mu*1 + element*(+-gamma) + 0. */
static void generate_sigma_points(OracleFilter *f) {
  int n = f->n, Na = f->Na, P = f->P;
  double g = f->w.gamma;
  double *sg = f->sigma;
  for (int r = 0; r < Na; r++) {
    double mu = r < n ? f->x[r] : 0.0;
    sg[(size_t)r * P + 0] = mu;
    for (int i = 0; i < Na; i++) {
      double e; /* sr(i, r) */
      if (i < n)
        e = r < n ? f->S[(size_t)i * n + r] : 0.0;
      else if (i < n + 3)
        e = (r == i) ? f->Mt[i - n] : 0.0;
      else
        e = (r == i) ? f->prm.sigma_measure : 0.0; /* Qt = I2 * m_sigmaMeasure, :238 */
      sg[(size_t)r * P + i + 1] = mu * 1 + e * g + 0;
      sg[(size_t)r * P + Na + i + 1] = mu * 1 + e * ((-1) * g) + 0;
    }
  }
}

/* SLAM.cpp:1430-1465 */
void oracle_predict_motion(OracleFilter *f, const double *u3) {
  int n = f->n, Na = f->Na, P = f->P;
  const OracleParams *p = &f->prm;
  double rot1 = u3[0], trans = u3[1], rot2 = u3[2];
  f->Ut[0] = rot1;
  f->Ut[1] = trans;
  f->Ut[2] = rot2;
  /* :1456-1458 */
  f->Mt[0] = p->a1 * rot1 * rot1 + p->a2 * trans * trans;
  f->Mt[1] = p->a3 * trans * trans + p->a4 * rot1 * rot1 + p->a4 * rot2 * rot2;
  f->Mt[2] = p->a1 * rot2 * rot2 + p->a2 * trans * trans;
  oracle_sample_parameters(Na, p, &f->w); /* :1460 */
  generate_sigma_points(f);               /* :1461-1463 */

  /* passSigmaThroughMotionFunction, :1476-1532 (noise type 0, :1490-1494) */
  double *sg = f->sigma;
  double mu[4] = {0, 0, 0, 0}; /* uninitialised in the reference, multiplied by 0 at :1527 */
  for (int i = 0; i < P; i++) {
    double r1 = f->Ut[0] - sg[(size_t)(n + 0) * P + i];
    double tr = f->Ut[1] - sg[(size_t)(n + 1) * P + i];
    double r2 = f->Ut[2] - sg[(size_t)(n + 2) * P + i];
    double th = sg[(size_t)(n - 1) * P + i];
    double upd[4];
    upd[0] = tr * cos(th + r1);
    upd[1] = tr * sin(th + r1);
    upd[2] = 0;
    upd[3] = r1 + r2;
    for (int k = 0; k < 4; k++) {
      sg[(size_t)(n - 4 + k) * P + i] += upd[k];
      double e = sg[(size_t)(n - 4 + k) * P + i];
      if (!i)
        mu[k] = e * f->w.wm0 + mu[k] * 0 + 0;
      else
        mu[k] = e * f->w.wi + mu[k] * 1 + 0;
    }
  }
  for (int k = 0; k < 4; k++) f->x[n - 4 + k] = mu[k]; /* :1531 */

  /* QrAndCholeskyForMotion, :1539-1556 */
  int m = 2 * Na;
  double *A = f->work_qr;
  for (int i = 0; i < m; i++)
    for (int c = 0; c < n; c++)
      A[(size_t)i * n + c] = f->w.wi_sr * (sg[(size_t)c * P + i + 1] - sg[(size_t)c * P + 0]);
  oracle_qr_R(A, m, n, f->S);
}

/* SLAM.cpp:1604-1608 */
void oracle_predict_measurement(OracleFilter *f) {
  int n = f->n, P = f->P, L = f->L, Na = f->Na;
  const OracleParams *p = &f->prm;
  double *sg = f->sigma;
  /* passSigmaThroughMesaurementFunction, :1615-1682 */
  for (int i = 0; i < P; i++) {
    double err[2] = {sg[(size_t)(n + 3) * P + i], sg[(size_t)(n + 4) * P + i]};
    double pos[3] = {sg[(size_t)(n - 4) * P + i], sg[(size_t)(n - 3) * P + i], sg[(size_t)(n - 2) * P + i]};
    double Rwc[9], Rcw[9];
    transfer_matrix(sg[(size_t)(n - 1) * P + i], Rwc);
    inv3x3(Rwc, Rcw);
    for (int id = 0; id < L; id++) {
      double st[6];
      for (int k = 0; k < 6; k++) st[k] = sg[(size_t)(6 * id + k) * P + i];
      double ox, oy;
      project_rcw(p, st, pos, Rcw, err, &ox, &oy);
      f->pix[(size_t)(2 * id + 0) * P + i] = ox;
      f->pix[(size_t)(2 * id + 1) * P + i] = oy;
    }
    for (int r = 0; r < 2 * L; r++) {
      double e = f->pix[(size_t)r * P + i];
      if (!i)
        f->hbar[r] = e * f->w.wm0 + 0 + 0;
      else
        f->hbar[r] = e * f->w.wi + f->hbar[r] * 1 + 0;
    }
  }
  /* QrAndCholeskyForMeasurement, :1700-1748, calculateOneFeatureCovariance :1759-1775 */
  int m = 2 * Na;
  double *A = f->work_qr;
  for (int id = 0; id < L; id++) {
    double px = f->hbar[2 * id + 0], py = f->hbar[2 * id + 1];
    f->visible[id] = 0;
    if (px != 0 && py != 0) {
      f->visible[id] = 1;
      for (int i = 0; i < m; i++) {
        A[2 * i + 0] = f->w.wi_sr * (f->pix[(size_t)(2 * id + 0) * P + i + 1] - f->pix[(size_t)(2 * id + 0) * P + 0]);
        A[2 * i + 1] = f->w.wi_sr * (f->pix[(size_t)(2 * id + 1) * P + i + 1] - f->pix[(size_t)(2 * id + 1) * P + 0]);
      }
      oracle_qr_R(A, m, 2, f->si + 4 * id);
    }
  }
}

/* GSLCholeskyUpdate, NEEDNOT_REORDER branch, SLAM.cpp:2106-2121,2139-2153.
 * u is n x nc row-major; sign = -1 downdating (:2149), +1 updating (:2144). */
/* getPermutationMatrix, SLAM.cpp:1303-1334, for a map whose last M = m_nFilters features are new:
 * disordered order = [old features 6(L-M) | robot 4 | (theta,phi,rho) of the new 3M | anchors of the new 3M].
 * canon[a] = canonical (row) index of disordered (column) index a, i.e. m_permutation(canon[a], a) = 1. */
static void reorder_map(int L, int M, int *canon) {
  int n = 6 * L + 4, dimOld = n - 6 * M;
  for (int a = 0; a < dimOld - 4; a++) canon[a] = a;
  for (int k = 0; k < 4; k++) canon[dimOld - 4 + k] = n - 4 + k;
  for (int id = 0; id < M; id++)
    for (int k = 0; k < 3; k++) {
      canon[dimOld + 3 * M + 3 * id + k] = dimOld - 4 + 6 * id + k;
      canon[dimOld + 3 * id + k] = dimOld - 4 + 6 * id + 3 + k;
    }
}

/* cv::Mat::inv() (DECOMP_LU) on an upper-triangular r x r matrix, OpenCV 2.4.3 modules/core/src/lapack.cpp LUImpl
 * applied to [A | I]: partial pivoting finds the diagonal (everything below it is exactly zero), the elimination
 * is a no-op, a pivot below DBL_EPSILON makes inv() return the zero matrix; then back substitution
 * b(i,:) = (b(i,:) - sum_{k>i} A(i,k) b(k,:)) * (1/A(i,i)). */
static int inv_upper(const double *A, int r, double *X) {
  memset(X, 0, sizeof(double) * (size_t)r * r);
  for (int i = 0; i < r; i++) {
    if (fabs(A[(size_t)i * r + i]) < 2.220446049250313e-16) {
      memset(X, 0, sizeof(double) * (size_t)r * r);
      return 0;
    }
    X[(size_t)i * r + i] = 1.0;
  }
  for (int i = r - 1; i >= 0; i--) {
    double dinv = -(-1.0 / A[(size_t)i * r + i]); /* LUImpl keeps d = -1/A(i,i) and stores A(i,i) = -d */
    for (int j = 0; j < r; j++) {
      double s = X[(size_t)i * r + j];
      for (int k = i + 1; k < r; k++) s -= A[(size_t)i * r + k] * X[(size_t)k * r + j];
      X[(size_t)i * r + j] = s * dinv;
    }
  }
  return 1;
}

/* GSLCholeskyUpdate, NEED_REORDER branch (SLAM.cpp:2122-2138) with CholeskyDecompositionWithPivoting
 * (:2158-2179): per column, dst = Perm^T (S^T S -+ u u^T) Perm; R11 = modified Cholesky of the leading
 * covRank x covRank block, R12 = R11^-T dst12, S_dis = [R11 R12; 0 0]; m_S_k = R of QR(Perm S_dis Perm^T).
 * M = m_nFilters (features added on the previous frame), covRank = dim - 3M (:2131). */
static void cholesky_update_reorder(OracleFilter *f, const double *U, int nc, double sign, int M) {
  int n = f->n, r = n - 3 * M, m2 = n - r;
  double *Pm = f->work_P;
  int *canon = (int *)malloc(sizeof(int) * (size_t)n);
  double *dst = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *Sdis = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *B = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *C11 = (double *)malloc(sizeof(double) * (size_t)r * r);
  double *R11 = (double *)malloc(sizeof(double) * (size_t)r * r);
  double *Xi = (double *)malloc(sizeof(double) * (size_t)r * r);
  reorder_map(f->L, M, canon);
  for (int c = 0; c < nc; c++) {
    memset(Pm, 0, sizeof(double) * (size_t)n * n); /* src1 = S^T S, :2118 */
    for (int k = 0; k < n; k++) {
      const double *row = f->S + (size_t)k * n;
      for (int i = k; i < n; i++) {
        double a = row[i];
        double *pr = Pm + (size_t)i * n;
        for (int j = k; j < n; j++) pr[j] += a * row[j];
      }
    }
    for (int i = 0; i < n; i++) {
      double ui = U[(size_t)i * nc + c];
      for (int j = 0; j < n; j++) Pm[(size_t)i * n + j] = Pm[(size_t)i * n + j] + sign * (ui * U[(size_t)j * nc + c]);
    }
    for (int a = 0; a < n; a++) /* :2127 / :2132 */
      for (int b = 0; b < n; b++) dst[(size_t)a * n + b] = Pm[(size_t)canon[a] * n + canon[b]] + 0;
    memset(Sdis, 0, sizeof(double) * (size_t)n * n); /* :2161 */
    if (n == r) {
      oracle_mchol(dst, n, f->prm.epsilon, Sdis, NULL);
    } else {
      for (int a = 0; a < r; a++)
        for (int b = 0; b < r; b++) C11[(size_t)a * r + b] = dst[(size_t)a * n + b];
      int mod = oracle_mchol(C11, r, f->prm.epsilon, R11, NULL); /* :2173 */
      f->n_mchol_calls++;
      if (mod) f->n_mchol_modified++;
      inv_upper(R11, r, Xi); /* :2175 R12 = R11.inv().t() * Cov12 */
      for (int a = 0; a < r; a++) {
        for (int b = 0; b < r; b++) Sdis[(size_t)a * n + b] = R11[(size_t)a * r + b];
        for (int t = 0; t < m2; t++) {
          double s = 0.0;
          for (int k = 0; k < r; k++) s += Xi[(size_t)k * r + a] * dst[(size_t)k * n + r + t];
          Sdis[(size_t)a * n + r + t] = s;
        }
      }
    }
    /* :2137 m_S_k = R of QR(Perm S_dis Perm^T): (Perm X Perm^T)(i,j) = X(dis(i), dis(j)) */
    for (int a = 0; a < n; a++)
      for (int b = 0; b < n; b++) B[(size_t)canon[a] * n + canon[b]] = Sdis[(size_t)a * n + b];
    oracle_qr_R(B, n, n, f->S);
  }
  free(Xi); free(R11); free(C11); free(B); free(Sdis); free(dst); free(canon);
}

static void cholesky_update(OracleFilter *f, const double *U, int nc, double sign) {
  int n = f->n;
  if (sign < 0 && f->n_new > 0) { /* KalmanUpdate :2083-2090: NEED_REORDER while m_nAddings != 0 */
    cholesky_update_reorder(f, U, nc, sign, f->n_new);
    return;
  }
  double *Pm = f->work_P;
  double *E = (double *)malloc(sizeof(double) * (size_t)n);
  int mode = f->prm.downdate_mode;
  for (int c = 0; c < nc; c++) {
    if (mode == 1 && c > 0) {
      /* carry-P: S^T S == G + E exactly in real arithmetic (:2288, :2321) */
      for (int i = 0; i < n; i++) Pm[(size_t)i * n + i] += E[i];
    } else {
      /* src1 = S^T S, :2118, accumulated as sum_k S(k,:)^T S(k,:) with k ascending (the order of a
       * row-by-column product).  mode 0 skips the exact zeros below the diagonal of S; mode 2 runs the
       * full dense n^3 loop as cv::Mat operator* does (CPU-baseline timing). */
      memset(Pm, 0, sizeof(double) * (size_t)n * n);
      for (int k = 0; k < n; k++) {
        const double *row = f->S + (size_t)k * n;
        int lo = (mode == 2) ? 0 : k;
        for (int i = lo; i < n; i++) {
          double a = row[i];
          double *pr = Pm + (size_t)i * n;
          for (int j = lo; j < n; j++) pr[j] += a * row[j];
        }
      }
    }
    /* dst = src1 -+ u u^T, :2119-2120, :2144/:2149 */
    for (int i = 0; i < n; i++) {
      double ui = U[(size_t)i * nc + c];
      for (int j = 0; j < n; j++) Pm[(size_t)i * n + j] = Pm[(size_t)i * n + j] + sign * (ui * U[(size_t)j * nc + c]);
    }
    int mod = oracle_mchol(Pm, n, f->prm.epsilon, f->S, E); /* :2152 */
    f->n_mchol_calls++;
    if (mod) f->n_mchol_modified++;
    for (int i = 0; i < n; i++)
      if (fabs(E[i]) > f->max_E) f->max_E = fabs(E[i]);
  }
  free(E);
}

/* deleteOneFeature, SLAM.cpp:2637-2663 (state and factor part; the list surgery at :2665-2706 has no arithmetic):
 * drop the feature's six entries of x and its six rows and columns of S, keep the dropped rows (restricted to the
 * surviving columns) as V, then GSLCholeskyUpdate(V^T, UPDATING, NEEDNOT_REORDER) (:2661-2662, :2139-2153): six
 * rank-one updates, each re-forming S^T S + v v^T and re-factorising with the modified Cholesky.
 * x [n], S [n x n] of an L-feature filter -> x_out [n-6], S_out [(n-6) x (n-6)]. */
void oracle_delete_feature(const OracleParams *p, int L, const double *x, const double *S, int id, double *x_out,
                           double *S_out) {
  int n = 6 * L + 4, m = n - 6;
  OracleFilter *f = oracle_filter_create(L - 1, p);
  double *U = (double *)calloc((size_t)m * 6, sizeof(double)); /* VT: m x 6 */
  for (int r = 0; r < m; r++) {
    int rs = (r < 6 * id) ? r : r + 6;
    f->x[r] = x[rs];
    for (int c = 0; c < m; c++) {
      int cs = (c < 6 * id) ? c : c + 6;
      f->S[(size_t)r * m + c] = S[(size_t)rs * n + cs];
    }
    for (int k = 0; k < 6; k++) U[(size_t)r * 6 + k] = S[(size_t)(6 * id + k) * n + rs];
  }
  cholesky_update(f, U, 6, +1.0);
  memcpy(x_out, f->x, sizeof(double) * (size_t)m);
  memcpy(S_out, f->S, sizeof(double) * (size_t)m * m);
  free(U);
  oracle_filter_destroy(f);
}

/* SLAM.cpp:2048-2096 (non-RANSAC branch) */
void oracle_kalman_update(OracleFilter *f, const double *z, const unsigned char *matched) {
  int n = f->n, P = f->P, L = f->L;
  double *sg = f->sigma;
  double *Pxy = f->work_U;         /* n x 2 */
  double *Ki = f->work_U + 2 * n;  /* n x 2 */
  double *U = f->work_U + 4 * n;   /* n x 2 */
  double *T = f->work_U + 6 * n;   /* n x 2 */
  f->max_E = 0.0;
  for (int id = 0; id < L; id++) {
    /* isMatching can only be set for visible features (dataAssociation, :1946-2001) */
    if (!(matched[id] && f->visible[id])) continue;
    double zi[2] = {z[2 * id], z[2 * id + 1]};
    double hi[2] = {f->hbar[2 * id], f->hbar[2 * id + 1]};
    const double *si = f->si + 4 * id;
    /* calculateOneFeatureCrossCovariance, :2020-2038 (uses the current, already-updated x) */
    for (int i = 0; i < P; i++) {
      double s0 = f->pix[(size_t)(2 * id + 0) * P + i] - hi[0];
      double s1 = f->pix[(size_t)(2 * id + 1) * P + i] - hi[1];
      for (int r = 0; r < n; r++) {
        double d = sg[(size_t)r * P + i] - f->x[r];
        if (!i) {
          Pxy[2 * r + 0] = f->w.wc0 * d * s0 + 0;
          Pxy[2 * r + 1] = f->w.wc0 * d * s1 + 0;
        } else {
          Pxy[2 * r + 0] += f->w.wi * d * s0;
          Pxy[2 * r + 1] += f->w.wi * d * s1;
        }
      }
    }
    /* sii = si.inv() : OpenCV 2x2 closed form, :2077 */
    double det = si[0] * si[3] - si[1] * si[2];
    double sii[4] = {0, 0, 0, 0};
    if (det != 0.) {
      double d = 1. / det;
      sii[0] = si[3] * d;
      sii[1] = -si[1] * d;
      sii[2] = -si[2] * d;
      sii[3] = si[0] * d;
    }
    /* Ki = Pxy*sii*sii^T, :2078 ; x += Ki*(zi-hi), :2079 ; U = Ki*si^T, :2080 */
    double inn[2] = {zi[0] - hi[0], zi[1] - hi[1]};
    for (int r = 0; r < n; r++) {
      T[2 * r + 0] = Pxy[2 * r + 0] * sii[0] + Pxy[2 * r + 1] * sii[2];
      T[2 * r + 1] = Pxy[2 * r + 0] * sii[1] + Pxy[2 * r + 1] * sii[3];
      Ki[2 * r + 0] = T[2 * r + 0] * sii[0] + T[2 * r + 1] * sii[1];
      Ki[2 * r + 1] = T[2 * r + 0] * sii[2] + T[2 * r + 1] * sii[3];
    }
    for (int r = 0; r < n; r++) f->x[r] += Ki[2 * r + 0] * inn[0] + Ki[2 * r + 1] * inn[1];
    for (int r = 0; r < n; r++) {
      U[2 * r + 0] = Ki[2 * r + 0] * si[0] + Ki[2 * r + 1] * si[1];
      U[2 * r + 1] = Ki[2 * r + 0] * si[2] + Ki[2 * r + 1] * si[3];
    }
    cholesky_update(f, U, 2, -1.0); /* :2089 */
  }
}

/* m_nAddings / m_nFilters of the previous frame's addFeatures (SLAM.cpp:758-766): while non-zero, KalmanUpdate takes
 * the NEED_REORDER branch (:2083-2086).  The reference resets it in addFeatures at the end of every frame (:554). */
void oracle_filter_set_new_features(OracleFilter *f, int n_new) { f->n_new = n_new; }

/* m_allPredictSet / map_p->Si / isVisible as left by predictMeasurement (SLAM.cpp:1724-1738) */
void oracle_filter_get_prediction(const OracleFilter *f, double *hbar, double *si, unsigned char *visible) {
  memcpy(hbar, f->hbar, sizeof(double) * 2 * (size_t)f->L);
  memcpy(si, f->si, sizeof(double) * 4 * (size_t)f->L);
  memcpy(visible, f->visible, (size_t)f->L);
}

/* Chi-square gate of dataAssociation, SLAM.cpp:1946-1977: pi = Si^T Si, pii = err * pi^-1 * err^T with
 * err = candidate - predictLocation, accepted when pii < CHI2INV_TABLE(0,2) (:54).  cv::Mat::inv on a 2x2 is the
 * determinant closed form.  d2 (may be NULL) receives pii; unvisible features are rejected with d2 = -1. */
void oracle_chi2_gate(const OracleFilter *f, const double *z, double threshold, unsigned char *accept, double *d2) {
  for (int id = 0; id < f->L; id++) {
    accept[id] = 0;
    if (d2) d2[id] = -1.0;
    if (!f->visible[id]) continue;
    const double *si = f->si + 4 * id;
    double p00 = si[0] * si[0] + si[2] * si[2], p01 = si[0] * si[1] + si[2] * si[3];
    double p10 = p01, p11 = si[1] * si[1] + si[3] * si[3];
    double det = p00 * p11 - p01 * p10;
    double i00 = 0, i01 = 0, i10 = 0, i11 = 0;
    if (det != 0.) {
      double d = 1. / det;
      i00 = p11 * d; i01 = -p01 * d; i10 = -p10 * d; i11 = p00 * d;
    }
    double e0 = z[2 * id] - f->hbar[2 * id], e1 = z[2 * id + 1] - f->hbar[2 * id + 1];
    double pii = (e0 * i00 + e1 * i10) * e0 + (e0 * i01 + e1 * i11) * e1;
    if (d2) d2[id] = pii;
    accept[id] = pii < threshold ? 1 : 0;
  }
}

void oracle_step(OracleFilter *f, const double *u3, const double *z, const unsigned char *matched) {
  oracle_predict_motion(f, u3);
  oracle_predict_measurement(f);
  oracle_kalman_update(f, z, matched);
}

/* ------------------------------------------------------------------------------------------ */
/* feature initialisation at frame 1, SLAM.cpp:818-871, 1177-1334                              */
/* ------------------------------------------------------------------------------------------ */
/* integrateFeaturesInformation for a map of Lold features (dim = 6 Lold + 4; Lold = 0 is frame 1):
 * x (dim), S (dim x dim) -> x_out (dim + 6M), S_out in canonical order [old features | new features | robot]. */
void oracle_add_features(const OracleParams *p, int Lold, const double *xin, const double *Sin, int M,
                         const double *kp, double rho0, double sigma_rho, double *x_out, double *S_out) {
  int dim = 6 * Lold + 4;
  int Na = dim + 3 * M;     /* :827 */
  int nCols = 2 * Na + 1;
  int dimNew = dim + 6 * M;
  OracleWeights w;
  oracle_sample_parameters(Na, p, &w); /* :867 */

  /* mu, sr : :847-868 (expandMatrix :1123-1135) */
  double *mu = (double *)calloc((size_t)Na, sizeof(double));
  double *sr = (double *)calloc((size_t)Na * Na, sizeof(double));
  for (int i = 0; i < dim; i++) {
    mu[i] = xin[i];
    for (int j = 0; j < dim; j++) sr[(size_t)i * Na + j] = Sin[(size_t)i * dim + j];
  }
  for (int i = 0; i < M; i++) {
    int idx = dim + 3 * i;
    mu[idx + 0] = kp[2 * i + 0];
    mu[idx + 1] = kp[2 * i + 1];
    mu[idx + 2] = rho0;
    sr[(size_t)(idx + 0) * Na + idx + 0] = p->sigma_measure;
    sr[(size_t)(idx + 1) * Na + idx + 1] = p->sigma_measure;
    sr[(size_t)(idx + 2) * Na + idx + 2] = sigma_rho;
  }
  /* generateSigmaPoints, :1148-1162 */
  double *sin_ = (double *)calloc((size_t)Na * nCols, sizeof(double));
  for (int r = 0; r < Na; r++) {
    sin_[(size_t)r * nCols] = mu[r];
    for (int i = 0; i < Na; i++) {
      double e = sr[(size_t)i * Na + r];
      sin_[(size_t)r * nCols + i + 1] = mu[r] * 1 + e * w.gamma + 0;
      sin_[(size_t)r * nCols + Na + i + 1] = mu[r] * 1 + e * ((-1) * w.gamma) + 0;
    }
  }
  /* passSigmaThroughMapingFunction, :1177-1250 */
  double *sout = (double *)calloc((size_t)dimNew * nCols, sizeof(double));
  double *mu_angle = (double *)calloc((size_t)3 * M, sizeof(double));
  double f1 = p->cam_f / p->cam_dx, f2 = p->cam_f / p->cam_dy;
  for (int r = 0; r < dim; r++)
    for (int i = 0; i < nCols; i++) sout[(size_t)r * nCols + i] = sin_[(size_t)r * nCols + i]; /* :1185 */
  for (int i = 0; i < nCols; i++) {
    double Rwc[9];
    transfer_matrix(sin_[(size_t)(dim - 1) * nCols + i], Rwc); /* :1204 */
    for (int id = 0; id < M; id++) {
      int index_in = dim + 3 * id, index_out1 = index_in, index_out2 = index_in + 3 * M;
      double uvd_x = sin_[(size_t)(index_in + 0) * nCols + i];
      double uvd_y = sin_[(size_t)(index_in + 1) * nCols + i];
      double rho = sin_[(size_t)(index_in + 2) * nCols + i];
      double uvu_x, uvu_y;
      oracle_undistort(p, uvd_x, uvd_y, &uvu_x, &uvu_y);           /* :1217 */
      double Hlr[3] = {(uvu_y - p->cam_cx) / f1, (uvu_x - p->cam_cy) / f2, 1}; /* :3360-3363 */
      double Hlw[3];
      for (int r = 0; r < 3; r++) Hlw[r] = Rwc[3 * r] * Hlr[0] + Rwc[3 * r + 1] * Hlr[1] + Rwc[3 * r + 2] * Hlr[2];
      double st[3];
      st[0] = atan2(Hlw[0], Hlw[2]);                                     /* :3411 */
      st[1] = atan2(-Hlw[1], sqrt(Hlw[0] * Hlw[0] + Hlw[2] * Hlw[2]));  /* :3412 */
      st[2] = rho;
      for (int k = 0; k < 3; k++) {
        sout[(size_t)(index_out1 + k) * nCols + i] = st[k];                                  /* :1222 */
        sout[(size_t)(index_out2 + k) * nCols + i] = sin_[(size_t)(dim - 4 + k) * nCols + i]; /* :1223 */
        if (!i)
          mu_angle[3 * id + k] = st[k] * w.wm0 + 0 + 0; /* :1235 */
        else
          mu_angle[3 * id + k] = st[k] * w.wi + mu_angle[3 * id + k] * 1 + 0; /* :1240 */
      }
    }
  }
  /* x_new, :1245-1249 (disordered: [x(dim) | angles(3M) | positions(3M)]) */
  double *xdis = (double *)calloc((size_t)dimNew, sizeof(double));
  for (int r = 0; r < dim; r++) xdis[r] = xin[r];
  for (int r = 0; r < 3 * M; r++) xdis[dim + r] = mu_angle[r];
  for (int id = 0; id < M; id++)
    for (int k = 0; k < 3; k++) xdis[Na + 3 * id + k] = xin[dim - 4 + k];
  /* QrAndCholeskyForInitilization, :1260-1300 */
  int m = 2 * Na;
  double *A = (double *)calloc((size_t)m * dimNew, sizeof(double));
  for (int i = 0; i < m; i++)
    for (int c = 0; c < dimNew; c++)
      A[(size_t)i * dimNew + c] = w.wi_sr * (sout[(size_t)c * nCols + i + 1] - sout[(size_t)c * nCols + 0]);
  double *Sdis = (double *)calloc((size_t)dimNew * dimNew, sizeof(double));
  oracle_qr_R(A, m, dimNew, Sdis);
  /* getPermutationMatrix, :1303-1334: src[canonical row] = disordered index */
  int *src = (int *)calloc((size_t)dimNew, sizeof(int));
  int dimOld = dim;
  for (int r = 0; r < dimOld - 4; r++) src[r] = r;
  for (int k = 0; k < 4; k++) src[dimNew - 4 + k] = dimOld - 4 + k;
  for (int id = 0; id < M; id++) {
    for (int k = 0; k < 3; k++) {
      src[dimOld - 4 + 6 * id + k] = dimOld + 3 * M + 3 * id + k;
      src[dimOld - 4 + 6 * id + 3 + k] = dimOld + 3 * id + k;
    }
  }
  /* :1293-1294  x = Perm x ; S = R of QR(Perm S Perm^T) */
  double *B = (double *)calloc((size_t)dimNew * dimNew, sizeof(double));
  for (int r = 0; r < dimNew; r++) {
    x_out[r] = xdis[src[r]];
    for (int c = 0; c < dimNew; c++) B[(size_t)r * dimNew + c] = Sdis[(size_t)src[r] * dimNew + src[c]];
  }
  oracle_qr_R(B, dimNew, dimNew, S_out);
  free(B);
  free(src);
  free(Sdis);
  free(A);
  free(xdis);
  free(mu_angle);
  free(sout);
  free(sin_);
  free(sr);
  free(mu);
}

void oracle_init_features(const OracleParams *p, const double *x4, const double *S4, int M,
                          const double *kp, double rho0, double sigma_rho, double *x_out, double *S_out) {
  oracle_add_features(p, 0, x4, S4, M, kp, rho0, sigma_rho, x_out, S_out);
}

/* ------------------------------------------------------------------------------------------ */
/* batch driver (tests / CPU baseline): filters are independent, one filter per pthread worker  */
/* ------------------------------------------------------------------------------------------ */
typedef struct BatchJob {
  int B, L, nsteps;
  const OracleParams *p;
  double *x, *S;
  const double *u, *z;
  const unsigned char *matched;
  double *max_E_out;
  int next; /* next filter to claim */
  pthread_mutex_t mu;
} BatchJob;

static void *batch_worker(void *arg) {
  BatchJob *j = (BatchJob *)arg;
  int n = 6 * j->L + 4, L = j->L, B = j->B;
  OracleFilter *f = oracle_filter_create(L, j->p);
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int b = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (b >= B) break;
    oracle_filter_set_state(f, j->x + (size_t)b * n, j->S + (size_t)b * n * n);
    double me = 0.0;
    for (int s = 0; s < j->nsteps; s++) {
      /* u: [nsteps][B][3], z: [nsteps][B][L][2], matched: [nsteps][B][L] */
      oracle_step(f, j->u + ((size_t)s * B + b) * 3, j->z + ((size_t)s * B + b) * L * 2,
                  j->matched + ((size_t)s * B + b) * L);
      if (f->max_E > me) me = f->max_E;
    }
    oracle_filter_get_state(f, j->x + (size_t)b * n, j->S + (size_t)b * n * n);
    if (j->max_E_out) j->max_E_out[b] = me;
  }
  oracle_filter_destroy(f);
  return NULL;
}

void oracle_batch_step(int B, int L, const OracleParams *p, double *x, double *S, const double *u,
                       const double *z, const unsigned char *matched, int step_stride, int nsteps,
                       int nthreads, double *max_E_out) {
  (void)step_stride;
  BatchJob j;
  j.B = B; j.L = L; j.nsteps = nsteps; j.p = p; j.x = x; j.S = S; j.u = u; j.z = z;
  j.matched = matched; j.max_E_out = max_E_out; j.next = 0;
  pthread_mutex_init(&j.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B;
  if (nthreads == 1) {
    batch_worker(&j);
  } else {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  pthread_mutex_destroy(&j.mu);
}
