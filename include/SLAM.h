/*
 * SLAM.h -- header-only C++ facade with the reference's CSLAM interface for the SRUKF path, on top of the
 * C ABI of libsrukf_b200.so (include/srukf.h).  No MFC, OpenCV, OpenGL or GSL types.
 *
 * Mirrors MonoSLAM/SLAM.h:118-398 for the members and methods that the hot path touches:
 *   members  m_X_k, m_S_k, m_P_k, Ut, Mt, Qt (SLAM.h:271-289), gamma, wm0, wc0, wi, wi_sr (SLAM.h:251-257),
 *            m_sample (SLAM.h:72-83,161), cam_* (SLAM.h:293-301), a1..a4, m_sigmaMeasure, m_weightType,
 *            EPSILON, imageWidth/imageHeight, m_nMatches, m_nPredicts, map (PointsMap, SLAM.h:47-70)
 *   methods  predictMotion(), predictMeasurement(), KalmanUpdate(), SLAM(), calculateSampleParameter()
 *            (SLAM.h:322,359-360,370,372)
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
  /* constants of SLAM.h:129-148 that the path uses */
  const int FLAG_4_WEIGHT1 = 0, FLAG_4_WEIGHT2 = 1, FLAG_4_WEIGHT3 = 2;
  const double EPSILON = 1e-13;

  /* ---- members (same names as the reference) ---- */
  Mat64 m_X_k, m_S_k, m_P_k;     /* state, upper-triangular factor (P = S^T S), covariance */
  Mat64 Ut, Mt, Qt;              /* control (rot1, trans, rot2), control noise, measurement noise */
  Mat64 m_allPredictSet;         /* predicted pixels, 2L x 1 */
  SampleParameter m_sample;
  std::vector<PointsMap> map;
  int m_weightType = 0, imageWidth = 640, imageHeight = 480;
  int m_nMapFeatures = 0, m_nPredicts = 0, m_nMatches = 0;
  double m_sigmaMeasure = 3.0, a1 = 8, a2 = 8, a3 = 8, a4 = 8;
  double gamma = 0, wm0 = 0, wm0_sr = 0, wc0 = 0, wc0_sr = 0, wi = 0, wi_sr = 0;
  double cam_dx = 0.0028, cam_dy = 0.0028, cam_cx = 310.1129, cam_cy = 236.7526, cam_k1 = 0.0001, cam_k2 = 0.0,
         cam_f = 2.1735, cam_f1 = 0, cam_f2 = 0;

  /* L landmarks already in the map (feature initialisation is outside this path) */
  explicit CSLAM(int L, int device = 0) : m_nMapFeatures(L) {
    const int n = 6 * L + 4;
    m_X_k = Mat64(n, 1);
    m_S_k = Mat64(n, n);
    m_P_k = Mat64(n, n);
    Ut = Mat64(3, 1);
    Mt = Mat64(3, 3);
    Qt = Mat64(2, 2);
    Qt(0, 0) = Qt(1, 1) = m_sigmaMeasure;  /* SLAM.cpp:238 */
    m_allPredictSet = Mat64(2 * L, 1);
    map.resize(L);
    for (int i = 0; i < L; ++i) map[i].ID = i + 1;
    cam_f1 = cam_f / cam_dx;
    cam_f2 = cam_f / cam_dy;
    SrukfParams p;
    srukf_default_params(&p);
    check(srukf_create(device, 1, L, &p, &h_), "srukf_create");
    calculateSampleParameter(n + 5);
  }
  ~CSLAM() { srukf_destroy(h_); }
  CSLAM(const CSLAM&) = delete;
  CSLAM& operator=(const CSLAM&) = delete;

  /* push m_X_k / m_S_k to the device (call after editing them on the host) */
  void uploadState() { check(srukf_set_state_dense(h_, m_X_k.ptr(), m_S_k.ptr()), "srukf_set_state_dense"); }

  /* SLAM.cpp:1050-1103 (weight type 0; the device uses the same formulas for all three types) */
  void calculateSampleParameter(const int& Na) {
    m_sample.num = Na;
    wm0 = 1.0 - Na / 3.0;
    wm0_sr = std::sqrt(std::fabs(wm0));
    wc0 = 1.0 - Na / 3.0;
    wc0_sr = std::sqrt(std::fabs(wm0));
    wi = (1.0 - wc0) / (2 * Na);
    wi_sr = std::sqrt(wi);
    gamma = std::sqrt(Na / (1.0 - wm0));
    m_sample.wm0 = wm0; m_sample.wc0 = wc0; m_sample.wi = wi; m_sample.Gamma = gamma;
  }

  /* SLAM.cpp:1430-1465 (motion part): uses Ut */
  void predictMotion() {
    const double rot1 = Ut(0), trans = Ut(1), rot2 = Ut(2);
    Mt(0, 0) = a1 * rot1 * rot1 + a2 * trans * trans;                      /* :1456-1458 */
    Mt(1, 1) = a3 * trans * trans + a4 * rot1 * rot1 + a4 * rot2 * rot2;
    Mt(2, 2) = a1 * rot2 * rot2 + a2 * trans * trans;
    check(srukf_predict_motion(h_, Ut.ptr()), "srukf_predict_motion");
    download();
  }

  /* SLAM.cpp:1604-1608: fills m_allPredictSet and map[i].predictLocation / Si / isVisible */
  void predictMeasurement() {
    check(srukf_predict_measurement(h_), "srukf_predict_measurement");
    const int L = (int)map.size();
    std::vector<double> si(4 * (size_t)L);
    std::vector<uint8_t> vis(L);
    check(srukf_get_prediction(h_, m_allPredictSet.ptr(), si.data(), vis.data()), "srukf_get_prediction");
    m_nPredicts = 0;
    for (int i = 0; i < L; ++i) {
      map[i].isVisible = vis[i] != 0;                                   /* :1727-1738 */
      if (map[i].isVisible) {
        m_nPredicts++;
        map[i].isMatching = false;
        map[i].predictLocation.x = m_allPredictSet(2 * i);
        map[i].predictLocation.y = m_allPredictSet(2 * i + 1);
        for (int k = 0; k < 4; ++k) map[i].Si.d[k] = si[4 * (size_t)i + k];
      }
    }
  }

  /* SLAM.cpp:2048-2096: uses map[i].matchLocation / isMatching */
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
    check(srukf_kalman_update(h_, z.data(), m.data()), "srukf_kalman_update");
    download();
  }

  /* the three path stages of CSLAM::SLAM(), SLAM.cpp:91,93,99, with the matches supplied in between by the
   * caller-provided data association (the image stages of the reference are outside this path) */
  template <typename Associate>
  void SLAM(Associate&& dataAssociation) {
    predictMotion();
    predictMeasurement();
    dataAssociation(*this);
    KalmanUpdate();
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
  srukf_t* handle() { return h_; }

 private:
  void download() { check(srukf_get_state_dense(h_, m_X_k.ptr(), m_S_k.ptr()), "srukf_get_state_dense"); }
  srukf_t* h_ = nullptr;
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
