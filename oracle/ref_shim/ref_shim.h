/*
 * ref_shim.h -- the slice of OpenCV 2.4 / GSL 1.8 / MFC that the reference's SRUKF functions touch, so that their
 * bodies -- extracted VERBATIM from /root/reference/MonoSLAM/SLAM.cpp at build time (extract_ref.py) -- and the
 * reference's own SLAM.h compile with plain g++ on Linux.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): it pins oracle/srukf_oracle.c to the reference's text.  Nothing here is product.
 *
 * What is reference text and what is restated here:
 *   reference text  every statement of the extracted CSLAM member functions and the CSLAM class definition
 *   restated        the third-party primitives those statements call, neither of which is vendored by the reference:
 *     - OpenCV 2.4.3 core: cv::Mat (double, 2-D, header/ROI semantics), MatExpr assignment semantics (`roi = expr`
 *       writes through when the sizes match), Mat::t / inv / diag / mul-free products, addWeighted, minMaxLoc, divide,
 *       sqrt, abs, repeat.  gemm accumulates sum_k a(i,k) b(k,j) in ascending k in one double accumulator (what 2.4's
 *       GEMMSingleMul / GEMMBlockMul do for CV_64F); Mat::inv() for n <= 3 is the determinant closed form of
 *       cv::invert, for larger n Gaussian elimination with partial pivoting (LUImpl).  tests/test_ref_shim.py checks
 *       these primitives against cv2 4.13's Python bindings (same arithmetic for these functions).
 *     - GSL 1.8 gsl_linalg_QR_decomp: forwarded to oracle_qr_decomp (oracle/srukf_oracle.c), the restatement of
 *       linalg/qr.c + householder.c.  THIS boundary stays a restatement.
 *     - MFC / highgui types that CSLAM only holds as members (CEdit, CListCtrl, CDC, CRect, CString, IplImage,
 *       CvCapture, TickMeter, KeyPoint, RNG): empty or minimal stand-ins; image loading returns a 640 x 480 header.
 *   Memory that the reference leaves uninitialised and then multiplies by 0 (Mat mu(4,1) at SLAM.cpp:1486,
 *   m_allPredictSet.create at :1635) is zero-filled here, as SURVEY 8(c) prescribes.
 */
#ifndef SRUKF_REF_SHIM_H
#define SRUKF_REF_SHIM_H

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

extern "C" void oracle_qr_decomp(double* A, int m, int n, double* tau);   /* oracle/srukf_oracle.c */

/* ------------------------------------------------------------------ MFC / Win32 / CRT stand-ins */
typedef int BOOL;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
struct CEdit {};
struct CListCtrl {};
struct CDC {};
struct CRect {};
struct CString {
  std::string s;
  CString() {}
  CString(const char* c) : s(c) {}
  CString& operator=(const char* c) { s = c; return *this; }
  operator const char*() const { return s.c_str(); }
};
inline int strcpy_s(char* dst, size_t n, const char* src) { std::strncpy(dst, src, n - 1); dst[n - 1] = 0; return 0; }
template <size_t N>
inline int sprintf_s(char (&buf)[N], const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int r = std::vsnprintf(buf, N, fmt, ap);
  va_end(ap);
  return r;
}
inline int fopen_s(FILE** f, const char*, const char*) { *f = NULL; return 0; }   /* the odometry file is not on the path */

/* ------------------------------------------------------------------ OpenCV C structures held by CSLAM */
struct IplROI { int width, height; };
struct IplImage { int width, height, depth, nChannels; IplROI* roi; };
struct CvCapture {};
struct CvRect { int x, y, width, height; };
struct CvSize { int width, height; };
struct CvPoint { int x, y; };
struct CvScalar { double val[4]; };
inline CvSize cvSize(int w, int h) { CvSize s = {w, h}; return s; }
inline CvRect cvRect(int x, int y, int w, int h) { CvRect r = {x, y, w, h}; return r; }
#define IPL_DEPTH_8U 8
#define CV_RGB2GRAY 7
#define CV_CAP_PROP_FPS 5
#define CV_CAP_PROP_POS_FRAMES 1
inline IplImage* ref_shim_image() { IplImage* im = new IplImage(); im->width = 640; im->height = 480; im->depth = 8; im->nChannels = 3; im->roi = NULL; return im; }
inline IplImage* cvLoadImage(const char*) { return ref_shim_image(); }   /* SLAM.cpp:312-313 reads only the size */
inline IplImage* cvCreateImage(CvSize s, int depth, int ch) { IplImage* im = ref_shim_image(); im->width = s.width; im->height = s.height; im->depth = depth; im->nChannels = ch; return im; }
inline void cvCvtColor(const IplImage*, IplImage*, int) {}
inline void cvReleaseImage(IplImage**) {}
inline CvCapture* cvCreateFileCapture(const char*) { return NULL; }
inline double cvGetCaptureProperty(CvCapture*, int) { return 0; }
inline int cvSetCaptureProperty(CvCapture*, int, double) { return 0; }
inline IplImage* cvQueryFrame(CvCapture*) { return ref_shim_image(); }

/* ------------------------------------------------------------------ cv:: subset */
#define CV_64F 6
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

struct Range {
  int start, end;
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
};
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(double v0 = 0) { val[0] = v0; val[1] = val[2] = val[3] = 0; } };
struct KeyPoint { Point2f pt; };
struct TickMeter { void reset() {} void start() {} void stop() {} double getTimeSec() const { return 0; } double getTimeMilli() const { return 0; } };
struct RNG { double gaussian(double) { std::abort(); return 0; } };   /* noise types 2/3 are not on the pinned path */
struct Matx33d {
  double val[9];
  Matx33d() { for (int i = 0; i < 9; i++) val[i] = 0; }
  Matx33d(double v0) { for (int i = 0; i < 9; i++) val[i] = 0; val[0] = v0; }   /* Matx(_Tp v0): used as `cam_K = 0` */
  Matx33d(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
    double v[9] = {a, b, c, d, e, f, g, h, i};
    for (int k = 0; k < 9; k++) val[k] = v[k];
  }
  double& operator()(int i, int j) { return val[3 * i + j]; }
  const double& operator()(int i, int j) const { return val[3 * i + j]; }
};

class MatExpr;
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_;

/* double, 2-D, reference-counted buffer + (rows, cols, step) header, as cv::Mat for CV_64F */
class Mat {
 public:
  int rows, cols;
  size_t step;                       /* doubles between consecutive rows */
  double* data;
  std::shared_ptr<std::vector<double> > buf;

  Mat() : rows(0), cols(0), step(0), data(NULL) {}
  Mat(int r, int c, int /*type*/) : rows(0), cols(0), step(0), data(NULL) { create(r, c, CV_64F); }
  Mat(const IplImage*) : rows(0), cols(0), step(0), data(NULL) {}   /* Mat image(m_gryImage): never dereferenced on the path */

  void create(int r, int c, int /*type*/) {
    if (data && r == rows && c == cols) return;
    buf.reset(new std::vector<double>((size_t)r * c, 0.0));   /* zero-filled, see the header comment */
    rows = r; cols = c; step = (size_t)c; data = buf->empty() ? NULL : &(*buf)[0];
  }
  bool empty() const { return data == NULL || rows == 0 || cols == 0; }
  template <typename T> T* ptr(int i = 0) { return reinterpret_cast<T*>(data + (size_t)i * step); }
  template <typename T> const T* ptr(int i = 0) const { return reinterpret_cast<const T*>(data + (size_t)i * step); }
  template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(data + (size_t)i * step + j); }
  template <typename T> const T& at(int i, int j) const { return *reinterpret_cast<const T*>(data + (size_t)i * step + j); }
  double& e(int i, int j) { return data[(size_t)i * step + j]; }
  double e(int i, int j) const { return data[(size_t)i * step + j]; }

  /* headers onto the same buffer */
  Mat view(int r0, int r1, int c0, int c1) const {
    Mat m;
    m.buf = buf; m.rows = r1 - r0; m.cols = c1 - c0; m.step = step; m.data = data + (size_t)r0 * step + c0;
    return m;
  }
  Mat row(int i) const { return view(i, i + 1, 0, cols); }
  Mat col(int j) const { return view(0, rows, j, j + 1); }
  Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
  Mat rowRange(const Range& r) const { return view(r.start, r.end, 0, cols); }
  Mat colRange(int a, int b) const { return view(0, rows, a, b); }
  Mat colRange(const Range& r) const { return view(0, rows, r.start, r.end); }
  Mat operator()(const Range& rr, const Range& cr) const { return view(rr.start, rr.end, cr.start, cr.end); }
  Mat diag() const {   /* Mat::diag(0): an n x 1 header with step + 1 */
    Mat m;
    m.buf = buf; m.rows = std::min(rows, cols); m.cols = 1; m.step = step + 1; m.data = data;
    return m;
  }
  static Mat diag(const Mat& d);     /* square matrix with the given diagonal (a copy) */

  Mat clone() const {
    Mat m(rows, cols, CV_64F);
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) m.e(i, j) = e(i, j);
    return m;
  }
  /* copyTo(OutputArray): dst.create(size) -- a no-op for a header of the right size, so ROIs are written through */
  void copyTo(const Mat& dst_) const {
    Mat& dst = const_cast<Mat&>(dst_);
    if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step == step) return;
    if (dst.buf.get() == buf.get() && buf) {   /* overlapping source and destination (deleteOneFeature): via a copy */
      Mat tmp = clone();
      dst.create(rows, cols, CV_64F);
      for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) dst.e(i, j) = tmp.e(i, j);
      return;
    }
    dst.create(rows, cols, CV_64F);
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) dst.e(i, j) = e(i, j);
  }

  /* m = expr: Mat::operator=(const MatExpr&) evaluates INTO m (create() + write), i.e. through a ROI header */
  Mat& operator=(const MatExpr& ex);
  Mat& operator=(const Scalar& s) { for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) e(i, j) = s.val[0]; return *this; }
  /* `pneg = NULL;` (SLAM.cpp:2318): MSVC's NULL is the int 0, i.e. Mat::operator=(const Scalar&) -- fill with 0 */
  Mat& operator=(long v) { return *this = Scalar((double)v); }
  Mat& operator=(int v) { return *this = Scalar((double)v); }
  Mat& operator=(double v) { return *this = Scalar(v); }
  Mat& operator+=(const Mat& b) { for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) e(i, j) += b.e(i, j); return *this; }
  Mat& operator-=(const Mat& b) { for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) e(i, j) -= b.e(i, j); return *this; }

  MatExpr t() const;
  MatExpr inv() const;
  static MatExpr zeros(int r, int c, int type);
  static MatExpr eye(int r, int c, int type);
};

/* the value of an expression: owns a fresh buffer.  Deriving from Mat lets it bind to `const Mat&` parameters. */
class MatExpr : public Mat {
 public:
  MatExpr() {}
  explicit MatExpr(const Mat& m) : Mat(m) {}
};
inline Mat& Mat::operator=(const MatExpr& ex) {
  if (data && rows == ex.rows && cols == ex.cols) {
    for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) e(i, j) = ex.e(i, j);
  } else {
    rows = ex.rows; cols = ex.cols; step = ex.step; data = ex.data; buf = ex.buf;
  }
  return *this;
}
inline Mat Mat::diag(const Mat& d) {
  const int n = d.rows * d.cols;
  Mat m(n, n, CV_64F);
  for (int i = 0; i < n; i++) m.e(i, i) = (d.cols == 1) ? d.e(i, 0) : d.e(0, i);
  return m;
}
inline MatExpr Mat::zeros(int r, int c, int) { return MatExpr(Mat(r, c, CV_64F)); }
inline MatExpr Mat::eye(int r, int c, int) { Mat m(r, c, CV_64F); for (int i = 0; i < std::min(r, c); i++) m.e(i, i) = 1.0; return MatExpr(m); }
inline MatExpr Mat::t() const {
  Mat m(cols, rows, CV_64F);
  for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) m.e(j, i) = e(i, j);
  return MatExpr(m);
}
/* cv::invert(DECOMP_LU): closed forms for n <= 3 (zeros when det == 0), LUImpl (partial pivoting) otherwise */
inline MatExpr Mat::inv() const {
  const int n = rows;
  Mat d(n, n, CV_64F);
  if (n == 2) {
    double det = e(0, 0) * e(1, 1) - e(0, 1) * e(1, 0);
    if (det != 0.) {
      det = 1. / det;
      double t0 = e(0, 0) * det, t1 = e(1, 1) * det;
      d.e(1, 1) = t0; d.e(0, 0) = t1;
      t0 = -e(0, 1) * det; t1 = -e(1, 0) * det;
      d.e(0, 1) = t0; d.e(1, 0) = t1;
    }
  } else if (n == 3) {
    double det = e(0, 0) * (e(1, 1) * e(2, 2) - e(1, 2) * e(2, 1)) - e(0, 1) * (e(1, 0) * e(2, 2) - e(1, 2) * e(2, 0)) +
                 e(0, 2) * (e(1, 0) * e(2, 1) - e(1, 1) * e(2, 0));
    if (det != 0.) {
      det = 1. / det;
      double t[9];
      t[0] = (e(1, 1) * e(2, 2) - e(1, 2) * e(2, 1)) * det;
      t[1] = (e(0, 2) * e(2, 1) - e(0, 1) * e(2, 2)) * det;
      t[2] = (e(0, 1) * e(1, 2) - e(0, 2) * e(1, 1)) * det;
      t[3] = (e(1, 2) * e(2, 0) - e(1, 0) * e(2, 2)) * det;
      t[4] = (e(0, 0) * e(2, 2) - e(0, 2) * e(2, 0)) * det;
      t[5] = (e(0, 2) * e(1, 0) - e(0, 0) * e(1, 2)) * det;
      t[6] = (e(1, 0) * e(2, 1) - e(1, 1) * e(2, 0)) * det;
      t[7] = (e(0, 1) * e(2, 0) - e(0, 0) * e(2, 1)) * det;
      t[8] = (e(0, 0) * e(1, 1) - e(0, 1) * e(1, 0)) * det;
      for (int i = 0; i < 9; i++) d.e(i / 3, i % 3) = t[i];
    }
  } else if (n == 1) {
    if (e(0, 0) != 0.) d.e(0, 0) = 1. / e(0, 0);
  } else {
    Mat A = clone();
    for (int i = 0; i < n; i++) d.e(i, i) = 1.0;
    bool ok = true;
    for (int i = 0; i < n && ok; i++) {
      int k = i;
      for (int j = i + 1; j < n; j++) if (std::abs(A.e(j, i)) > std::abs(A.e(k, i))) k = j;
      if (std::abs(A.e(k, i)) < 2.220446049250313e-16) { ok = false; break; }
      if (k != i) {
        for (int j = i; j < n; j++) std::swap(A.e(i, j), A.e(k, j));
        for (int j = 0; j < n; j++) std::swap(d.e(i, j), d.e(k, j));
      }
      const double dd = -1 / A.e(i, i);
      for (int j = i + 1; j < n; j++) {
        const double alpha = A.e(j, i) * dd;
        for (k = i + 1; k < n; k++) A.e(j, k) += alpha * A.e(i, k);
        for (k = 0; k < n; k++) d.e(j, k) += alpha * d.e(i, k);
      }
      A.e(i, i) = -dd;
    }
    if (!ok) { d = Scalar(0); }
    else {
      for (int i = n - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
          double s = d.e(i, j);
          for (int k = i + 1; k < n; k++) s -= A.e(i, k) * d.e(k, j);
          d.e(i, j) = s * A.e(i, i);
        }
    }
  }
  return MatExpr(d);
}

/* ---- expressions (evaluated eagerly, element order as the reference's operands) ---- */
inline MatExpr operator+(const Mat& a, const Mat& b) { Mat m(a.rows, a.cols, CV_64F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.e(i, j) = a.e(i, j) + b.e(i, j); return MatExpr(m); }
inline MatExpr operator-(const Mat& a, const Mat& b) { Mat m(a.rows, a.cols, CV_64F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.e(i, j) = a.e(i, j) - b.e(i, j); return MatExpr(m); }
inline MatExpr operator+(const Mat& a, double s) { Mat m(a.rows, a.cols, CV_64F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.e(i, j) = a.e(i, j) + s; return MatExpr(m); }
inline MatExpr operator*(const Mat& a, double s) { Mat m(a.rows, a.cols, CV_64F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.e(i, j) = a.e(i, j) * s; return MatExpr(m); }
inline MatExpr operator*(double s, const Mat& a) { return a * s; }
/* gemm: one double accumulator per output element, k ascending.  The loops run i-k-j over the output row so that the
 * compiler can vectorise across j; every element still receives its products in ascending k into a zero-initialised
 * accumulator, i.e. the same bits as the textbook i-j-k loop (2.4's GEMMSingleMul order for CV_64F). */
inline MatExpr operator*(const Mat& a, const Mat& b) {
  Mat m(a.rows, b.cols, CV_64F);
  const int K = a.cols, N = b.cols;
  for (int i = 0; i < a.rows; i++) {
    double* __restrict__ mi = m.data + (size_t)i * m.step;
    const double* ai = a.data + (size_t)i * a.step;
    for (int k = 0; k < K; k++) {
      const double aik = ai[k];
      const double* __restrict__ bk = b.data + (size_t)k * b.step;
      for (int j = 0; j < N; j++) mi[j] += aik * bk[j];
    }
  }
  return MatExpr(m);
}
inline MatExpr abs(const Mat& a) { Mat m(a.rows, a.cols, CV_64F); for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.e(i, j) = std::fabs(a.e(i, j)); return MatExpr(m); }

/* dst = src1*alpha + src2*beta + gamma (double arithmetic, in this order) */
inline void addWeighted(const Mat& s1, double alpha, const Mat& s2, double beta, double gamma, const Mat& dst_) {
  Mat& dst = const_cast<Mat&>(dst_);
  dst.create(s1.rows, s1.cols, CV_64F);
  for (int i = 0; i < s1.rows; i++) for (int j = 0; j < s1.cols; j++) dst.e(i, j) = s1.e(i, j) * alpha + s2.e(i, j) * beta + gamma;
}
/* first occurrence of the extreme value, scanning rows then columns; Point(x = column, y = row) */
inline void minMaxLoc(const Mat& a, double* minVal, double* maxVal, Point* minLoc = NULL, Point* maxLoc = NULL) {
  double mn = a.e(0, 0), mx = a.e(0, 0);
  Point pn(0, 0), px(0, 0);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) {
      const double v = a.e(i, j);
      if (v < mn) { mn = v; pn = Point(j, i); }
      if (v > mx) { mx = v; px = Point(j, i); }
    }
  if (minVal) *minVal = mn;
  if (maxVal) *maxVal = mx;
  if (minLoc) *minLoc = pn;
  if (maxLoc) *maxLoc = px;
}
/* 2.4's div_: 0 where the divisor is 0 */
inline void divide(const Mat& a, const Mat& b, const Mat& dst_) {
  Mat& dst = const_cast<Mat&>(dst_);
  dst.create(a.rows, a.cols, CV_64F);
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) dst.e(i, j) = (b.e(i, j) != 0) ? a.e(i, j) / b.e(i, j) : 0.0;
}
inline void sqrt(const Mat& a, Mat& dst) {
  dst.create(a.rows, a.cols, CV_64F);
  for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) dst.e(i, j) = std::sqrt(a.e(i, j));
}
inline void repeat(const Mat& src, int ny, int nx, const Mat& dst_) {
  Mat& dst = const_cast<Mat&>(dst_);
  dst.create(src.rows * ny, src.cols * nx, CV_64F);
  for (int i = 0; i < dst.rows; i++) for (int j = 0; j < dst.cols; j++) dst.e(i, j) = src.e(i % src.rows, j % src.cols);
}

/* Mat_<double>(r, c) << v0, v1, ... */
template <typename T>
class MatCommaInitializer_ {
 public:
  Mat m;
  int idx;
  MatCommaInitializer_(const Mat& mm) : m(mm), idx(0) {}
  template <typename T2> MatCommaInitializer_<T>& operator,(T2 v) { m.e(idx / m.cols, idx % m.cols) = (T)v; idx++; return *this; }
  operator Mat() const { return m; }
};
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_(int r, int c) : Mat(r, c, CV_64F) {}
};
template <typename T, typename T2>
inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v) {
  MatCommaInitializer_<T> ci(m);
  return (ci, v);
}

}  // namespace cv

/* ------------------------------------------------------------------ GSL subset (row-major, tda == size2) */
struct gsl_matrix { size_t size1, size2, tda; double* data; };
struct gsl_vector { size_t size; double* data; };
struct gsl_permutation { size_t size; size_t* data; };
inline gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2) { gsl_matrix* m = new gsl_matrix(); m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = new double[n1 * n2](); return m; }
inline void gsl_matrix_free(gsl_matrix* m) { if (m) { delete[] m->data; delete m; } }
inline gsl_vector* gsl_vector_alloc(size_t n) { gsl_vector* v = new gsl_vector(); v->size = n; v->data = new double[n](); return v; }
inline void gsl_vector_free(gsl_vector* v) { if (v) { delete[] v->data; delete v; } }
inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
inline int gsl_linalg_QR_decomp(gsl_matrix* A, gsl_vector* tau) { oracle_qr_decomp(A->data, (int)A->size1, (int)A->size2, tau->data); return 0; }

#endif
