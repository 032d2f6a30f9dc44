/*
 * SLAM.h -- header-only C++ facade with the reference's CSLAM interface for the SRUKF path, on top of the
 * C ABI of libsrukf_b200.so (include/srukf.h).  No MFC, OpenCV, OpenGL or GSL types.
 *
 * Mirrors MonoSLAM/SLAM.h:118-398 for the members and methods that the hot path touches:
 *   members  m_X_k, m_S_k, m_P_k, Ut, Mt, Qt (SLAM.h:271-289), gamma, wm0, wc0, wi, wi_sr (SLAM.h:251-257),
 *            m_sample (SLAM.h:72-83,161), cam_* (SLAM.h:293-301), a1..a4, m_sigmaMeasure, m_weightType,
 *            EPSILON, imageWidth/imageHeight, m_nMatches, m_nPredicts, map (PointsMap, SLAM.h:47-70)
 *   methods  SLAM(), predictMotion(), predictMeasurement(), KalmanUpdate(), dataAssociation() (chi-square gate),
 *            calculateSampleParameter(), generateSigmaPoints(), passSigmaThroughMotionFunction(),
 *            QrAndCholeskyForMotion(), passSigmaThroughMesaurementFunction(), QrAndCholeskyForMeasurement(),
 *            GSLCholeskyUpdate(), GSLQrDecomposition(), modifiedCholeskyDecomposition()
 *            (SLAM.h:322,341,347-348,355-363,370,372)
 * As in the reference, data flows through the public members: set Ut, call predictMotion(); set
 * map[i].matchLocation / isMatching, call KalmanUpdate(); read m_X_k / m_S_k.
 *
 * CSLAM is ONE filter (B = 1, for drop-in use and tests); CSLAMBatch is B independent filters with
 * structure-of-arrays members (what the GPU is for).  Errors from the C ABI are thrown as std::runtime_error;
 * there is no CPU fallback.
 */
#ifndef SRUKF_SLAM_FACADE_H
#define SRUKF_SLAM_FACADE_H

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "srukf.h"

namespace srukf {

/* minimal stand-in for the cv::Mat members (CV_64F, row-major) */
struct Mat64 {
  int rows = 0, cols = 0;
  std::vector<double> d;
  Mat64() = default;
  Mat64(int r, int c) : rows(r), cols(c), d((size_t)r * c, 0.0) {}
  static Mat64 zeros(int r, int c) { return Mat64(r, c); }
  double& operator()(int r, int c = 0) { return d[(size_t)r * cols + c]; }
  double operator()(int r, int c = 0) const { return d[(size_t)r * cols + c]; }
  double* ptr(int r = 0) { return d.data() + (size_t)r * cols; }
  const double* ptr(int r = 0) const { return d.data() + (size_t)r * cols; }
};

struct Point2d { double x = 0, y = 0; };

/* SLAM.h:47-70, fields used by the path (the reference keeps a heap linked list; here a vector) */
struct PointsMap {
  int ID = 0;
  bool isVisible = false;
  bool isMatching = false;
  Point2d predictLocation;
  Point2d matchLocation;
  Mat64 Si = Mat64(2, 2);
};

/* SLAM.h:72-83 */
struct SampleParameter {
  int num = 0;
  double Alpha = 1e-3, Beta = 2, Gamma = 0, Kappa = 0, Lammda = 0, wm0 = 0, wc0 = 0, wi = 0;
};

inline void check(int rc, const char* what) {
  if (rc != SRUKF_OK) throw std::runtime_error(std::string(what) + ": " + srukf_last_error());
}

class CSLAM {
 public:
  /* constants of SLAM.h:129-148 that the path uses (values of SLAM.cpp:21-56) */
  const int FLAG_4_WEIGHT1 = 0, FLAG_4_WEIGHT2 = 1, FLAG_4_WEIGHT3 = 2;
  const int FLAG_4_UPDATING = SRUKF_UPDATING, FLAG_4_DOWNDATING = SRUKF_DOWNDATING;
  const int FLAG_4_NEED_REORDER = SRUKF_NEED_REORDER, FLAG_4_NEEDNOT_REORDER = SRUKF_NEEDNOT_REORDER;
  const double EPSILON = 1e-13;
  const double CHI2INV_95_2 = 5.99146454710798;   /* CHI2INV_TABLE(0,2), SLAM.cpp:54 */

  /* ---- members (same names as the reference) ---- */
  Mat64 m_X_k, m_S_k, m_P_k;     /* state, upper-triangular factor (P = S^T S), covariance */
  Mat64 Ut, Mt, Qt;              /* control (rot1, trans, rot2), control noise, measurement noise */
  Mat64 m_allPredictSet;         /* predicted pixels, 2L x 1 */
  SampleParameter m_sample;
  std::vector<PointsMap> map;
  int m_weightType = 0, imageWidth = 640, imageHeight = 480;
  int m_nMapFeatures = 0, m_nPredicts = 0, m_nMatches = 0;
  int m_nAddings = 0, m_nFilters = 0;   /* features added on the previous frame: KalmanUpdate takes NEED_REORDER (:2082) */
  double m_sigmaMeasure = 3.0, a1 = 8, a2 = 8, a3 = 8, a4 = 8;
  double gamma = 0, wm0 = 0, wm0_sr = 0, wc0 = 0, wc0_sr = 0, wi = 0, wi_sr = 0;
  double cam_dx = 0.0028, cam_dy = 0.0028, cam_cx = 310.1129, cam_cy = 236.7526, cam_k1 = 0.0001, cam_k2 = 0.0,
         cam_f = 2.1735, cam_f1 = 0, cam_f2 = 0;

  /* L landmarks already in the map.  The scalar members above start at the reference's defaults
   * (initializeParameters, SLAM.cpp:158-343); after editing any of them call applyParameters(). */
  explicit CSLAM(int L, int device = 0) : m_nMapFeatures(L), device_(device) {
    const int n = 6 * L + 4;
    m_X_k = Mat64(n, 1);
    m_S_k = Mat64(n, n);
    m_P_k = Mat64(n, n);
    Ut = Mat64(3, 1);
    Mt = Mat64(3, 3);
    Qt = Mat64(2, 2);
    m_allPredictSet = Mat64(2 * L, 1);
    map.resize(L);
    for (int i = 0; i < L; ++i) map[i].ID = i + 1;
    visible_now_.assign(L, 0);
    applyParameters();
  }
  ~CSLAM() { srukf_destroy(h_); }
  CSLAM(const CSLAM&) = delete;
  CSLAM& operator=(const CSLAM&) = delete;

  /* The device works from a copy of the scalar members (camera, noise, weight type, image size, EPSILON) taken here:
   * the handle is re-created from the CURRENT member values and m_X_k / m_S_k are uploaded again, so host members
   * and device arithmetic cannot drift apart silently. */
  void applyParameters() {
    SrukfParams p;
    srukf_default_params(&p);
    p.cam_dx = cam_dx; p.cam_dy = cam_dy; p.cam_cx = cam_cx; p.cam_cy = cam_cy; p.cam_k1 = cam_k1; p.cam_k2 = cam_k2;
    p.cam_f = cam_f; p.image_width = imageWidth; p.image_height = imageHeight;
    p.a1 = a1; p.a2 = a2; p.a3 = a3; p.a4 = a4; p.sigma_measure = m_sigmaMeasure; p.weight_type = m_weightType;
    p.alpha = m_sample.Alpha; p.beta = m_sample.Beta; p.epsilon = EPSILON;
    srukf_t* nh = nullptr;
    check(srukf_create(device_, 1, (int)map.size(), &p, &nh), "srukf_create");
    if (h_) srukf_destroy(h_);
    h_ = nh;
    prm_ = p;
    cam_f1 = cam_f / cam_dx;   /* SLAM.cpp:336-337 */
    cam_f2 = cam_f / cam_dy;
    Qt(0, 0) = Qt(1, 1) = m_sigmaMeasure;  /* SLAM.cpp:238 */
    Qt(0, 1) = Qt(1, 0) = 0;
    calculateSampleParameter(m_X_k.rows + 5);
    uploadState();
  }

  /* push m_X_k / m_S_k to the device (call after editing them on the host) */
  void uploadState() { check(srukf_set_state_dense(h_, m_X_k.ptr(), m_S_k.ptr()), "srukf_set_state_dense"); }

  /* SLAM.cpp:1050-1103, all three weight types, operation order kept */
  void calculateSampleParameter(const int& Na) {
    m_sample.num = Na;
    if (m_weightType == FLAG_4_WEIGHT1) {
      wm0 = 1.0 - Na / 3.0;
      wm0_sr = std::sqrt(std::fabs(wm0));
      wc0 = 1.0 - Na / 3.0;
      wc0_sr = std::sqrt(std::fabs(wm0));
      wi = (1.0 - wc0) / (2 * Na);
      wi_sr = std::sqrt(wi);
      gamma = std::sqrt(Na / (1.0 - wm0));
    } else if (m_weightType == FLAG_4_WEIGHT2) {
      m_sample.Kappa = 0;
      m_sample.Lammda = std::pow(m_sample.Alpha, 2) * (Na + m_sample.Kappa) - Na;
      gamma = std::sqrt(Na + m_sample.Lammda);
      wm0 = m_sample.Lammda / (Na + m_sample.Lammda);
      wm0_sr = std::sqrt(std::fabs(wm0));
      wc0 = wm0 + (1 - std::pow(m_sample.Alpha, 2) + m_sample.Beta);
      wc0_sr = std::sqrt(std::fabs(wc0));
      wi = 1.0 / (2 * (Na + m_sample.Lammda));
      wi_sr = std::sqrt(std::fabs(wi));
    } else {
      gamma = std::sqrt(3.0 * Na / 2.0);
      wm0 = 1.0 / 3.0;
      wm0_sr = std::sqrt(wm0);
      wc0 = 1.0 / 3.0;
      wc0_sr = std::sqrt(wc0);
      wi = 1.0 / (3.0 * Na);
      wi_sr = std::sqrt(wi);
    }
    m_sample.wm0 = wm0; m_sample.wc0 = wc0; m_sample.wi = wi; m_sample.Gamma = gamma;
  }

  /* SLAM.cpp:1430-1465 (motion part): uses Ut.  = passSigmaThroughMotionFunction(Ut) + QrAndCholeskyForMotion() */
  void predictMotion() {
    passSigmaThroughMotionFunction(Ut);
    QrAndCholeskyForMotion();
  }
  /* SLAM.h:359 / SLAM.cpp:1476-1532.  On the device, sigma-point generation, the motion model and the square-root
   * update are ONE kernel; this half records the control and Mt (:1456-1458), QrAndCholeskyForMotion launches. */
  void passSigmaThroughMotionFunction(const Mat64& u) {
    if (&u != &Ut) Ut = u;
    const double rot1 = Ut(0), trans = Ut(1), rot2 = Ut(2);
    Mt(0, 0) = a1 * rot1 * rot1 + a2 * trans * trans;
    Mt(1, 1) = a3 * trans * trans + a4 * rot1 * rot1 + a4 * rot2 * rot2;
    Mt(2, 2) = a1 * rot2 * rot2 + a2 * trans * trans;
    motion_pending_ = true;
  }
  /* SLAM.h:361 / SLAM.cpp:1539-1556: m_X_k (robot mean) and m_S_k after the motion step */
  void QrAndCholeskyForMotion() {
    if (!motion_pending_) throw std::runtime_error("QrAndCholeskyForMotion: call passSigmaThroughMotionFunction first");
    check(srukf_predict_motion(h_, Ut.ptr()), "srukf_predict_motion");
    motion_pending_ = false;
    download();
  }

  /* SLAM.cpp:1604-1608: fills m_allPredictSet and map[i].predictLocation / Si / isVisible */
  void predictMeasurement() {
    passSigmaThroughMesaurementFunction();
    QrAndCholeskyForMeasurement();
  }
  /* SLAM.h:360 / SLAM.cpp:1615-1682 (spelling as in the reference): projections of every sigma point, predicted means */
  void passSigmaThroughMesaurementFunction() { check(srukf_predict_measurement(h_), "srukf_predict_measurement"); }
  /* SLAM.h:362 / SLAM.cpp:1700-1748: per-feature Si, visibility, predictLocation */
  void QrAndCholeskyForMeasurement() {
    const int L = (int)map.size();
    std::vector<double> si(4 * (size_t)L);
    std::vector<uint8_t> vis(L);
    check(srukf_get_prediction(h_, m_allPredictSet.ptr(), si.data(), vis.data()), "srukf_get_prediction");
    m_nPredicts = 0;
    for (int i = 0; i < L; ++i) {
      if (vis[i]) {                                                     /* :1727-1738 (isVisible is sticky there too) */
        m_nPredicts++;
        map[i].isVisible = true;
        map[i].isMatching = false;
        map[i].predictLocation.x = m_allPredictSet(2 * i);
        map[i].predictLocation.y = m_allPredictSet(2 * i + 1);
        for (int k = 0; k < 4; ++k) map[i].Si.d[k] = si[4 * (size_t)i + k];
      }
      visible_now_[i] = vis[i];
    }
  }

  /* SLAM.cpp:2048-2096: uses map[i].matchLocation / isMatching; NEED_REORDER while m_nAddings != 0 (:2082-2089) */
  void KalmanUpdate() {
    const int L = (int)map.size();
    std::vector<double> z(2 * (size_t)L);
    std::vector<uint8_t> m(L);
    m_nMatches = 0;
    for (int i = 0; i < L; ++i) {
      z[2 * (size_t)i] = map[i].matchLocation.x;
      z[2 * (size_t)i + 1] = map[i].matchLocation.y;
      m[i] = map[i].isMatching ? 1 : 0;
      m_nMatches += m[i];
    }
    if (m_nAddings != 0) check(srukf_kalman_update_reorder(h_, z.data(), m.data(), m_nFilters), "srukf_kalman_update_reorder");
    else check(srukf_kalman_update(h_, z.data(), m.data()), "srukf_kalman_update");
    download();
  }

  /* Batched part of CSLAM::dataAssociation (SLAM.cpp:1946-1977): the candidates in map[i].matchLocation pass when
   * their Mahalanobis distance to the prediction is below CHI2INV_TABLE(0,2); sets map[i].isMatching.  (The patch
   * correlation that proposes candidates in the reference works on images and is outside this path.) */
  void dataAssociation() {
    const int L = (int)map.size();
    std::vector<double> z(2 * (size_t)L);
    std::vector<uint8_t> acc(L);
    for (int i = 0; i < L; ++i) { z[2 * (size_t)i] = map[i].matchLocation.x; z[2 * (size_t)i + 1] = map[i].matchLocation.y; }
    check(srukf_chi2_gate(h_, z.data(), CHI2INV_95_2, acc.data(), nullptr), "srukf_chi2_gate");
    m_nMatches = 0;
    for (int i = 0; i < L; ++i) { map[i].isMatching = acc[i] != 0; m_nMatches += acc[i]; }
  }

  /* CSLAM::SLAM(), SLAM.cpp:88-110, the stages of this path in the reference's order: predictMotion (:91),
   * predictMeasurement (:93), dataAssociation (:97), KalmanUpdate (:99).  Set Ut and the candidate pixels
   * map[i].matchLocation before the call (the reference's loadPictures / patch matching produce them from images). */
  void SLAM() {
    predictMotion();
    predictMeasurement();
    dataAssociation();
    KalmanUpdate();
  }
  /* same with a caller-supplied association step between prediction and update */
  template <typename Associate>
  void SLAM(Associate&& associate) {
    predictMotion();
    predictMeasurement();
    associate(*this);
    KalmanUpdate();
  }

  /* ---- helper methods of the reference (SLAM.h:341,347-348,355), each one CUDA entry point ---- */
  /* SLAM.cpp:1148-1162: sigma is Na x (2Na+1) */
  void generateSigmaPoints(Mat64& sigma, const Mat64& mu, const Mat64& sr) {
    const int Na = mu.rows;
    if (sigma.rows != Na || sigma.cols != 2 * Na + 1) sigma = Mat64(Na, 2 * Na + 1);
    check(srukf_generate_sigma_points(device_, 1, Na, gamma, mu.ptr(), sr.ptr(), sigma.ptr()), "srukf_generate_sigma_points");
  }
  /* SLAM.cpp:2106-2155 on m_S_k (uploads the host copy first, as the reference works on the member) */
  void GSLCholeskyUpdate(const Mat64& u, const int& flag4UpOrDown, const int& flag4Order) {
    uploadState();
    check(srukf_cholesky_update(h_, u.ptr(), u.cols, flag4UpOrDown, flag4Order, m_nFilters), "srukf_cholesky_update");
    download();
  }
  /* SLAM.cpp:2330-2353: R = triu of the Householder QR of A (GSL sign convention) */
  void GSLQrDecomposition(Mat64& R, const Mat64& A) const {
    if (R.rows != A.cols || R.cols != A.cols) R = Mat64(A.cols, A.cols);
    check(srukf_qr_R(device_, 1, A.rows, A.cols, A.ptr(), R.ptr()), "srukf_qr_R");
  }
  /* SLAM.cpp:2197-2327: sr = sqrt(D) L^T of the Gill-Murray-Wright factorisation of Cov */
  void modifiedCholeskyDecomposition(Mat64& sr, const Mat64& Cov) {
    if (sr.rows != Cov.rows || sr.cols != Cov.cols) sr = Mat64(Cov.rows, Cov.cols);
    check(srukf_mchol(device_, 1, Cov.rows, EPSILON, Cov.ptr(), sr.ptr(), nullptr), "srukf_mchol");
  }

  /* m_P_k = S^T S (SLAM.cpp:2404), computed on the device */
  void updateCovariance() {
    check(srukf_get_cov_block(h_, 0, m_X_k.rows, m_P_k.ptr()), "srukf_get_cov_block");
  }

  uint32_t flags() {
    uint32_t f = 0;
    check(srukf_get_flags(h_, &f), "srukf_get_flags");
    return f;
  }
  /* visible on THIS frame (map[i].isVisible is sticky in the reference) */
  bool visibleNow(int i) const { return visible_now_[i] != 0; }
  srukf_t* handle() { return h_; }

 private:
  void download() { check(srukf_get_state_dense(h_, m_X_k.ptr(), m_S_k.ptr()), "srukf_get_state_dense"); }
  srukf_t* h_ = nullptr;
  SrukfParams prm_{};
  int device_ = 0;
  bool motion_pending_ = false;
  std::vector<uint8_t> visible_now_;
};

/* B independent CSLAM filters; members are structure-of-arrays, state stays resident on the device. */
class CSLAMBatch {
 public:
  const int B, L, n;
  std::vector<double> Ut;             /* [B][3]    */
  std::vector<double> matchLocation;  /* [B][L][2] */
  std::vector<uint8_t> isMatching;    /* [B][L]    */

  CSLAMBatch(int B_, int L_, const SrukfParams* prm = nullptr, int device = 0)
      : B(B_), L(L_), n(6 * L_ + 4), Ut((size_t)B_ * 3), matchLocation((size_t)B_ * L_ * 2),
        isMatching((size_t)B_ * L_, 1) {
    check(srukf_create(device, B, L, prm, &h_), "srukf_create");
  }
  ~CSLAMBatch() { srukf_destroy(h_); }
  CSLAMBatch(const CSLAMBatch&) = delete;
  CSLAMBatch& operator=(const CSLAMBatch&) = delete;

  void setState(const double* x, const double* S_packed) { check(srukf_set_state(h_, x, S_packed), "srukf_set_state"); }
  void getState(double* x, double* S_packed) { check(srukf_get_state(h_, x, S_packed), "srukf_get_state"); }
  /* addFeatures with an empty map (SLAM.cpp:818-871): x4 [B][4], S4 [B][4][4], keypoints [B][L][2] */
  void initFeatures(const double* x4, const double* S4, const double* keypoints, double rho0 = 1.0 / 3.0,
                    double sigma_rho = 1.0 / 6.0) {
    check(srukf_init_features(h_, x4, S4, keypoints, rho0, sigma_rho), "srukf_init_features");
  }
  /* integrateFeaturesInformation on a non-empty map (SLAM.cpp:818-871): dst holds L + M features */
  void addFeatures(CSLAMBatch& dst, const double* keypoints, double rho0 = 1.0 / 3.0, double sigma_rho = 1.0 / 6.0) {
    check(srukf_add_features(h_, dst.h_, keypoints, rho0, sigma_rho), "srukf_add_features");
  }
  /* deleteOneFeature (SLAM.cpp:2637-2663): filter b drops feature ids[b]; dst must hold L-1 features */
  void deleteFeature(CSLAMBatch& dst, const int32_t* ids) {
    check(srukf_delete_feature(h_, dst.h_, ids), "srukf_delete_feature");
  }
  void predictMotion() { check(srukf_predict_motion(h_, Ut.data()), "srukf_predict_motion"); }
  void predictMeasurement() { check(srukf_predict_measurement(h_), "srukf_predict_measurement"); }
  /* chi-square gate of dataAssociation (SLAM.cpp:1946-1977) on the candidates in matchLocation -> isMatching */
  void chi2Gate(double threshold = 5.99146454710798, double* d2 = nullptr) {
    check(srukf_chi2_gate(h_, matchLocation.data(), threshold, isMatching.data(), d2), "srukf_chi2_gate");
  }
  void KalmanUpdate() { check(srukf_kalman_update(h_, matchLocation.data(), isMatching.data()), "srukf_kalman_update"); }
  /* KalmanUpdate while m_nAddings != 0 (NEED_REORDER, SLAM.cpp:2083-2086): nNew = m_nFilters */
  void KalmanUpdateReorder(int nNew) {
    check(srukf_kalman_update_reorder(h_, matchLocation.data(), isMatching.data(), nNew), "srukf_kalman_update_reorder");
  }
  void SLAM() { check(srukf_step(h_, Ut.data(), matchLocation.data(), isMatching.data()), "srukf_step"); }
  void sync() { check(srukf_sync(h_), "srukf_sync"); }
  srukf_t* handle() { return h_; }

 private:
  srukf_t* h_ = nullptr;
};

}  // namespace srukf
#endif /* SRUKF_SLAM_FACADE_H */
