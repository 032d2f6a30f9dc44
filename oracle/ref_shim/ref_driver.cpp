/*
 * ref_driver.cpp -- C ABI over the reference's OWN SRUKF code (TEST INFRASTRUCTURE ONLY).
 *
 * oracle/_ref/slam_extract.cpp holds the bodies of the CSLAM member functions on the hot path, copied VERBATIM from
 * /root/reference/MonoSLAM/SLAM.cpp at build time by extract_ref.py (never committed); they compile against
 * ref_shim.h.  This file is the glue a test needs around them: it fills the public members the reference methods read
 * (m_X_k, m_S_k, the `map` list, m_odoXY / m_odoTheta, the counters) exactly as the reference's own callers leave them,
 * calls the reference method, and copies the members it wrote back out.  No arithmetic of the path lives here.
 *
 * Members the reference sets in functions that are NOT on the path (and so are not extracted) are set here with the
 * citation of the statement that sets them.
 */
#include "ref_shim.h"
#include "SLAM.h"

/* ---- members of CSLAM that the extracted bodies reference but that are off the path: inert definitions ---- */
CSLAM::~CSLAM(void) {}
void CSLAM::loadOdometryData(void) {}   /* SLAM.cpp:462-496 reads E:\SLAM\...; the driver writes m_odoXY / m_odoTheta itself */
void CSLAM::addFeatures(void) {}        /* image feature detection; the driver builds `map` itself */
double CSLAM::Gauss(const double&) { std::fprintf(stderr, "ref_driver: noise types 2-4 are not pinned\n"); std::abort(); return 0; }
void CSLAM::get3DdisplayInformation(Quaternion&, Point3d&, const Mat&) const { std::abort(); }
void CSLAM::getFeatureCartesianInformation(Point3d&, Mat&, Mat&, const int&) const { std::abort(); }

namespace {

void free_map(CSLAM* s) {
  while (s->map) { PointsMap* nx = s->map->next; delete s->map; s->map = nx; }
}

/* L nodes in list order = state order (integrateFeaturesInformation appends at the tail, SLAM.cpp:913-942) */
void build_map(CSLAM* s, int L) {
  free_map(s);
  PointsMap* tail = NULL;
  for (int i = 0; i < L; i++) {
    PointsMap* p = new PointsMap();
    p->ID = s->ID++;
    p->isVisible = false; p->isMatching = false; p->isLoop = false;
    p->nPredictTimes = 0; p->nMatchTimes = 0;
    p->next = NULL;
    if (!tail) s->map = p; else tail->next = p;
    tail = p;
  }
  s->m_nMapFeatures = L;
}

Mat from_array(const double* a, int r, int c) {
  Mat m(r, c, CV_64F);
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.ptr<double>(i)[j] = a[(size_t)i * c + j];
  return m;
}
void to_array(const Mat& m, double* a) {
  for (int i = 0; i < m.rows; i++) for (int j = 0; j < m.cols; j++) a[(size_t)i * m.cols + j] = m.ptr<double>(i)[j];
}

}  // namespace

extern "C" {

void* ref_create(void) { return new CSLAM(); }
void ref_destroy(void* h) { CSLAM* s = (CSLAM*)h; free_map(s); delete s; }

/* the parameters initializeParameters() left in the object (SLAM.cpp:158-343), for the test to compare with
 * oracle_default_params: cam_dx dy cx cy k1 k2 f f1 f2, a1..a4, sigma_measure, rho, sigma_rho, sigmaX Y Z Theta,
 * EPSILON, image width, height, weight type, noise type, alpha, beta */
void ref_get_params(void* h, double* out27) {
  CSLAM* s = (CSLAM*)h;
  double v[] = {s->cam_dx, s->cam_dy, s->cam_cx, s->cam_cy, s->cam_k1, s->cam_k2, s->cam_f, s->cam_f1, s->cam_f2,
                s->a1, s->a2, s->a3, s->a4, s->m_sigmaMeasure, s->m_rho, s->m_sigmaRHO,
                s->m_sigmaX, s->m_sigmaY, s->m_sigmaZ, s->m_sigmaTheta, s->EPSILON,
                (double)s->imageWidth, (double)s->imageHeight, (double)s->m_weightType, (double)s->m_noiseType,
                s->m_sample.Alpha, s->m_sample.Beta};
  for (int i = 0; i < 27; i++) out27[i] = v[i];
}
void ref_set_weight_type(void* h, int t) { ((CSLAM*)h)->m_weightType = t; }
void ref_set_camera(void* h, double k1, double k2) { ((CSLAM*)h)->cam_k1 = k1; ((CSLAM*)h)->cam_k2 = k2; }

/* calculateSampleParameter, SLAM.cpp:1050-1103 -> gamma wm0 wm0_sr wc0 wc0_sr wi wi_sr */
void ref_sample_parameters(void* h, int Na, double* out7) {
  CSLAM* s = (CSLAM*)h;
  s->calculateSampleParameter(Na);
  out7[0] = s->gamma; out7[1] = s->wm0; out7[2] = s->wm0_sr; out7[3] = s->wc0; out7[4] = s->wc0_sr; out7[5] = s->wi; out7[6] = s->wi_sr;
}

/* modifiedCholeskyDecomposition, SLAM.cpp:2197-2327 */
void ref_mchol(void* h, const double* G, int n, double* S_out) {
  CSLAM* s = (CSLAM*)h;
  Mat C = from_array(G, n, n), S;
  s->modifiedCholeskyDecomposition(S, C);
  to_array(S, S_out);
}

/* GSLQrDecomposition, SLAM.cpp:2330-2353 (the QR itself is the restated GSL routine) */
void ref_qr_R(void* h, const double* A, int m, int n, double* R_out) {
  CSLAM* s = (CSLAM*)h;
  Mat R;
  s->GSLQrDecomposition(R, from_array(A, m, n));
  to_array(R, R_out);
}

/* camera chain, SLAM.cpp:3177-3236 */
void ref_distort(void* h, double ux, double uy, double* dx, double* dy) {
  Point2d uvd, uvu(ux, uy);
  ((CSLAM*)h)->distortOnePointRW(uvd, uvu);
  *dx = uvd.x; *dy = uvd.y;
}
void ref_undistort(void* h, double dx, double dy, double* ux, double* uy) {
  Point2d uvu, uvd(dx, dy);
  ((CSLAM*)h)->undistortOnePointRW(uvu, uvd);
  *ux = uvu.x; *uy = uvu.y;
}
/* one projection exactly as passSigmaThroughMesaurementFunction chains it, SLAM.cpp:1640-1662 */
void ref_project(void* h, const double* feat6, const double* pos3, double theta, const double* err2, double* out2) {
  CSLAM* s = (CSLAM*)h;
  Mat state = from_array(feat6, 6, 1), position = from_array(pos3, 3, 1), error = from_array(err2, 2, 1);
  Mat Rwc, Rcw, Hlw, Hlr;
  Point2d uvu, uvd;
  s->getTransferMatrix(Rwc, theta);
  Rcw = Rwc.inv();
  s->coordinatesState2World(Hlw, state, position);
  s->coordinatesWorld2Camera(Hlr, Hlw, Rcw);
  s->coordinatesCamera2Image(uvu, Hlr, error);
  s->distortOnePointRW(uvd, uvu);
  out2[0] = uvd.x; out2[1] = uvd.y;
}

/* state of an L-feature map: x (n), S (n x n dense upper); rebuilds the `map` list */
void ref_set_state(void* h, int L, const double* x, const double* S) {
  CSLAM* s = (CSLAM*)h;
  const int n = 6 * L + 4;
  s->m_X_k = from_array(x, n, 1);
  s->m_S_k = from_array(S, n, n);
  build_map(s, L);
  s->m_nFilters = 0; s->m_nAddings = 0; s->m_nMatches = 0; s->m_nPredicts = 0;
}
int ref_state_dim(void* h) { return ((CSLAM*)h)->m_X_k.rows; }
void ref_get_state(void* h, double* x, double* S) {
  CSLAM* s = (CSLAM*)h;
  if (x) to_array(s->m_X_k, x);
  if (S) to_array(s->m_S_k, S);
}

/* predictMotion, SLAM.cpp:1343-1466, from two odometry poses (x, y, theta) at k-1 and k; the redirection flag
 * (m_odoTheta row 2) is 0.  Returns the control the reference derived (Ut) and Mt's diagonal. */
void ref_predict_motion(void* h, const double* odo_prev3, const double* odo_now3, double* Ut3, double* Mt3) {
  CSLAM* s = (CSLAM*)h;
  s->m_frame.counter = 2;   /* any frame >= 1 whose predecessor exists; `1 == counter` only refreshes display counters */
  const int c = s->m_frame.counter;
  s->m_odoXY[2 * c - 2] = odo_prev3[0]; s->m_odoXY[2 * c - 1] = odo_prev3[1];
  s->m_odoXY[2 * c + 0] = odo_now3[0];  s->m_odoXY[2 * c + 1] = odo_now3[1];
  s->m_odoTheta.ptr<double>(0)[c] = c;  s->m_odoTheta.ptr<double>(2)[c] = 0;
  s->m_odoTheta.ptr<double>(1)[c - 1] = odo_prev3[2];
  s->m_odoTheta.ptr<double>(1)[c] = odo_now3[2];
  s->predictMotion();
  for (int i = 0; i < 3; i++) { if (Ut3) Ut3[i] = s->Ut.ptr<double>(i)[0]; if (Mt3) Mt3[i] = s->Mt.ptr<double>(i)[i]; }
}
/* m_sigma after the motion step (Na x P), for the sigma-point comparison */
void ref_get_sigma(void* h, double* sigma, int* Na, int* P) {
  CSLAM* s = (CSLAM*)h;
  if (Na) *Na = s->m_sigma.rows;
  if (P) *P = s->m_sigma.cols;
  if (sigma) to_array(s->m_sigma, sigma);
}

/* predictMeasurement, SLAM.cpp:1604-1608 */
void ref_predict_measurement(void* h) { ((CSLAM*)h)->predictMeasurement(); }
/* per feature: predictLocation (x, y), Si (2 x 2), isVisible as this frame's QrAndCholeskyForMeasurement left them.
 * isVisible is sticky in the reference (never reset to false on the path), so "visible this frame" is re-derived from
 * m_allPredictSet exactly as SLAM.cpp:1727 tests it. */
void ref_get_prediction(void* h, double* hbar, double* si, unsigned char* visible, double* pix) {
  CSLAM* s = (CSLAM*)h;
  int id = 0;
  for (PointsMap* p = s->map; p; p = p->next, id++) {
    const double px = s->m_allPredictSet.ptr<double>(2 * id + 0)[0], py = s->m_allPredictSet.ptr<double>(2 * id + 1)[0];
    const bool vis = (px != 0 && py != 0);
    if (hbar) { hbar[2 * id] = px; hbar[2 * id + 1] = py; }
    if (visible) visible[id] = vis;
    if (si) for (int k = 0; k < 4; k++) si[4 * id + k] = (vis && !p->Si.empty()) ? p->Si.ptr<double>(k / 2)[k % 2] : 0.0;
  }
  if (pix) to_array(s->m_sigma_allPixel, pix);
}

/* KalmanUpdate, SLAM.cpp:2048-2104.  z[2i] = matchLocation.x, z[2i+1] = .y; n_new = m_nAddings (selects NEED_REORDER,
 * :2082) with m_nFilters = n_new (m_covRank, :2129/2148) and m_permutation from getPermutationMatrix (:1303-1334), as
 * integrateFeaturesInformation leaves them on a frame that added features. */
void ref_kalman_update(void* h, const double* z, const unsigned char* matched, int n_new) {
  CSLAM* s = (CSLAM*)h;
  int id = 0, nm = 0;
  for (PointsMap* p = s->map; p; p = p->next, id++) {
    p->isMatching = matched[id] != 0;
    if (p->isMatching) { p->matchLocation.x = z[2 * id]; p->matchLocation.y = z[2 * id + 1]; nm++; }
  }
  s->m_nMatches = nm;
  s->m_nAddings = n_new;
  s->m_nFilters = n_new;
  if (n_new) s->getPermutationMatrix();
  s->KalmanUpdate();
}

/* GSLCholeskyUpdate on a caller-supplied factor: S (n x n), U (n x k); mode 0 = UPDATING, 1 = DOWNDATING;
 * order 0 = NEED_REORDER, 1 = NEEDNOT_REORDER (the reference's flag values, SLAM.cpp:31-36); n_new sets m_nFilters /
 * m_nAddings (rank of the leading block, :2129) and nmap m_nMapFeatures (:2124/2143). */
void ref_cholesky_update(void* h, int n, const double* S, const double* U, int k, int mode, int order, int n_new,
                         int nmap, double* S_out) {
  CSLAM* s = (CSLAM*)h;
  s->m_X_k = Mat::zeros(n, 1, CV_64F);
  s->m_S_k = from_array(S, n, n);
  s->m_nFilters = n_new; s->m_nAddings = n_new; s->m_nMapFeatures = nmap;
  if (order == 0) s->getPermutationMatrix();
  s->GSLCholeskyUpdate(from_array(U, n, k), mode, order);
  to_array(s->m_S_k, S_out);
}
void ref_get_permutation(void* h, double* Pm) { to_array(((CSLAM*)h)->m_permutation, Pm); }

/* deleteOneFeature, SLAM.cpp:2637-2706, on the current state; id = position in the state, the node's own ID is looked up */
void ref_delete_feature(void* h, int id) {
  CSLAM* s = (CSLAM*)h;
  PointsMap* p = s->map;
  for (int i = 0; i < id && p; i++) p = p->next;
  if (!p) return;
  s->map = s->deleteOneFeature(id, p->ID, s->map);
}

/* Feature initialisation: the unscented transform of integrateFeaturesInformation, SLAM.cpp:818-871 -- the statements
 * up to QrAndCholeskyForInitilization are re-issued here in the reference's order (the function itself also cuts image
 * patches, :925-927, so it is not extractable); every callee is reference text.  kp are the key-points' (pt.x, pt.y);
 * the reference stores them as float (KeyPoint::pt is Point2f), reproduced.  M key-points are appended to the current
 * state (4 rows at frame 1, 6 Lold + 4 later). */
void ref_init_features(void* h, int M, const double* kp, double rho, double sigma_rho) {
  CSLAM* s = (CSLAM*)h;
  s->m_rho = rho; s->m_sigmaRHO = sigma_rho;
  s->m_nFilters = M; s->m_nAddings = M;
  s->m_keyPoints.clear();
  for (int i = 0; i < M; i++) { KeyPoint k; k.pt.x = (float)kp[2 * i]; k.pt.y = (float)kp[2 * i + 1]; s->m_keyPoints.push_back(k); }
  int dim = s->m_X_k.rows;
  s->m_sample.num = dim + 3 * s->m_nFilters;                                    /* :827 */
  int Na = s->m_sample.num;
  Mat mu2(3 * s->m_nFilters, 1, CV_64F);                                        /* :832-833 */
  Mat sr2 = Mat::zeros(3 * s->m_nFilters, 3 * s->m_nFilters, CV_64F);
  for (int i = 0; i < M; i++) {                                                 /* :847-858 */
    int index = 3 * i;
    mu2.ptr<double>(index + 0)[0] = s->m_keyPoints[i].pt.x;
    mu2.ptr<double>(index + 1)[0] = s->m_keyPoints[i].pt.y;
    mu2.ptr<double>(index + 2)[0] = s->m_rho;
    sr2.ptr<double>(index + 0)[index + 0] = s->m_sigmaMeasure;
    sr2.ptr<double>(index + 1)[index + 1] = s->m_sigmaMeasure;
    sr2.ptr<double>(index + 2)[index + 2] = s->m_sigmaRHO;
  }
  Mat mu(Na, 1, CV_64F);                                                        /* :860-865 */
  Mat sr = Mat::zeros(Na, Na, CV_64F);
  Mat sigma_in(dim + 3 * s->m_nFilters, 2 * Na + 1, CV_64F);
  Mat sigma_out(dim + 6 * s->m_nFilters, 2 * Na + 1, CV_64F);
  Mat mu_Hlw(3 * s->m_nFilters, 1, CV_64F);
  Mat mu_angle(3 * s->m_nFilters, 1, CV_64F);
  s->calculateSampleParameter(Na);                                              /* :867-871 */
  s->expandMatrix(mu, sr, s->m_X_k, s->m_S_k, mu2, sr2);
  s->generateSigmaPoints(sigma_in, mu, sr);
  s->passSigmaThroughMapingFunction(sigma_out, mu_Hlw, mu_angle, sigma_in);
  s->QrAndCholeskyForInitilization(sigma_out);
  /* the list grows by M nodes at the tail (:913-942) */
  PointsMap* tail = s->map;
  while (tail && tail->next) tail = tail->next;
  for (int i = 0; i < M; i++) {
    PointsMap* p = new PointsMap();
    p->ID = s->ID++; p->isVisible = false; p->isMatching = false; p->isLoop = false; p->nPredictTimes = 0; p->nMatchTimes = 0; p->next = NULL;
    if (!tail) s->map = p; else tail->next = p;
    tail = p;
  }
  s->m_nMapFeatures += M;
}

}  // extern "C"

/* ---- the stand-in's own cv:: primitives, exposed so that tests/test_ref_shim.py can check them against cv2 ---- */
extern "C" {
void shim_gemm(const double* A, int m, int k, const double* B, int n, double* out) { to_array(Mat(from_array(A, m, k) * from_array(B, k, n)), out); }
void shim_inv(const double* A, int n, double* out) { to_array(Mat(from_array(A, n, n).inv()), out); }
void shim_add_weighted(const double* A, double alpha, const double* B, double beta, double gamma, int m, int n, double* out) {
  Mat d;
  addWeighted(from_array(A, m, n), alpha, from_array(B, m, n), beta, gamma, d);
  to_array(d, out);
}
void shim_divide(const double* A, const double* B, int m, int n, double* out) {
  Mat d;
  divide(from_array(A, m, n), from_array(B, m, n), d);
  to_array(d, out);
}
void shim_min_max_loc(const double* A, int m, int n, double* mn, double* mx, int* loc4) {
  Point a, b;
  minMaxLoc(from_array(A, m, n), mn, mx, &a, &b);
  loc4[0] = a.x; loc4[1] = a.y; loc4[2] = b.x; loc4[3] = b.y;
}
/* ROI semantics: m(rows r0..r1, cols c0..c1) = expr writes through; copyTo into an overlapping ROI of the same buffer */
void shim_roi_assign(double* A, int m, int n, int r0, int r1, int c0, int c1, double s) {
  Mat M = from_array(A, m, n);
  M(Range(r0, r1), Range(c0, c1)) = M(Range(r0, r1), Range(c0, c1)) * s;
  to_array(M, A);
}
void shim_roi_shift_left(double* A, int m, int n, int k) {
  Mat M = from_array(A, m, n);
  M(Range(0, m), Range(k, n)).copyTo(M(Range(0, m), Range(0, n - k)));
  to_array(M, A);
}
}
