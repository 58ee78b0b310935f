// TEST INFRASTRUCTURE — OpenCV-API shim for compiling the reference's src/ORBmatcher.cc VERBATIM
// (oracle/Makefile: _ref/libmatcher_ref.so).  Written from scratch; provides only what that file uses:
// KeyPoint / Point2f, and a small cv::Mat (CV_8U descriptor rows, CV_32F algebra).
//
// The float algebra reproduces how OpenCV 4.x evaluates these tiny expressions (pinned against cv2.gemm
// in tests/test_cvprim_vs_cv2.py and by the probes recorded in DESIGN.md §2):
//   * A*B (+C) with no transposed operand: cv::gemm's small-matrix path, float32 products summed left to
//     right, then alpha, then the addend — identical to evaluating eagerly in float;
//   * a transposed operand (A.t()*B, A*B.t()) leaves that path: products and sums in double, alpha (and a
//     fused addend, `C - A.t()*B`) applied in double, one rounding to float.  cv::MatExpr is lazy, so the
//     transpose and the product are proxy types here (MatT, GemmT) that fold the same way;
//   * scalar*Mat, Mat/scalar: elementwise float multiply by the scalar (or its double reciprocal) rounded
//     to float; 3x3 inv(): closed form in double; dot / norm: double accumulation.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;
struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};

class Mat;
struct MatT;    // alpha * A^T, lazy
struct GemmT;   // alpha * op(A) * op(B) with a transposed operand, lazy

class Mat {
 public:
  int rows = 0, cols = 0;
  int type_ = CV_32F;
  size_t step = 0;  // bytes per row
  std::shared_ptr<std::vector<unsigned char>> store;
  unsigned char* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(const GemmT& g);  // materialise
  Mat(const MatT& t);
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * esz();
    store = std::make_shared<std::vector<unsigned char>>((size_t)r * step + 16, 0);
    data = store->data();
  }
  size_t esz() const { return type_ == CV_32F ? 4 : 1; }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || cols == 0; }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < std::min(r, c); ++i) m.at<float>(i, i) = 1.f;
    return m;
  }
  template <typename T> T& at(int i, int j) { return *reinterpret_cast<T*>(data + (size_t)i * step + (size_t)j * sizeof(T)); }
  template <typename T> const T& at(int i, int j) const {
    return *reinterpret_cast<const T*>(data + (size_t)i * step + (size_t)j * sizeof(T));
  }
  template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> T* ptr(int i = 0) { return reinterpret_cast<T*>(data + (size_t)i * step); }
  template <typename T> const T* ptr(int i = 0) const { return reinterpret_cast<const T*>(data + (size_t)i * step); }
  unsigned char* ptr(int i = 0) { return data + (size_t)i * step; }
  const unsigned char* ptr(int i = 0) const { return data + (size_t)i * step; }
  Mat view(int r0, int r1, int c0, int c1) const {
    Mat m;
    m.rows = r1 - r0; m.cols = c1 - c0; m.type_ = type_; m.step = step; m.store = store;
    m.data = data + (size_t)r0 * step + (size_t)c0 * esz();
    return m;
  }
  Mat row(int i) const { return view(i, i + 1, 0, cols); }
  Mat col(int j) const { return view(0, rows, j, j + 1); }
  Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
  Mat colRange(int a, int b) const { return view(0, rows, a, b); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; ++i) std::memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * esz());
    return m;
  }
  void copyTo(Mat& dst) const {
    if (dst.rows != rows || dst.cols != cols || dst.type_ != type_) dst.create(rows, cols, type_);
    for (int i = 0; i < rows; ++i) std::memcpy(dst.data + (size_t)i * dst.step, data + (size_t)i * step, (size_t)cols * esz());
  }
  void release() { rows = cols = 0; store.reset(); data = nullptr; }
  // u8 -> f32 / same-type copy (Frame::ComputeStereoMatches converts 11x11 patches in place, src/Frame.cc:879-880)
  void convertTo(Mat& dst, int type) const {
    Mat out(rows, cols, type);
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) {
        const float v = type_ == CV_32F ? at<float>(i, j) : (float)at<unsigned char>(i, j);
        if (type == CV_32F) out.at<float>(i, j) = v; else out.at<unsigned char>(i, j) = (unsigned char)v;
      }
    dst = out;
  }
  static Mat ones(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.f;
    return m;
  }
  // header over foreign memory (no ownership), e.g. a pyramid level handed in by a test
  static Mat wrap(int r, int c, int type, void* p, size_t step_bytes) {
    Mat m;
    m.rows = r; m.cols = c; m.type_ = type; m.step = step_bytes; m.data = static_cast<unsigned char*>(p);
    return m;
  }
  Mat reshape(int) const { return *this; }  // N x 2 one-channel <-> N x 1 two-channel: same memory, channels are not modelled
  MatT t() const;
  Mat inv() const;  // 3x3 CV_32F: closed form in double
  double dot(const Mat& b) const {
    double s = 0;
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) s += (double)at<float>(i, j) * (double)b.at<float>(i, j);
    return s;
  }
};

struct MatT {
  Mat a;
  double alpha;
  Mat inv() const { return Mat(*this).inv(); }
};
struct GemmT {
  Mat a, b;
  bool ta, tb;
  double alpha;
};

inline MatT Mat::t() const { return MatT{*this, 1.0}; }
inline Mat::Mat(const MatT& t) {
  create(t.a.cols, t.a.rows, CV_32F);
  const float al = (float)t.alpha;
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) at<float>(i, j) = t.alpha == 1.0 ? t.a.at<float>(j, i) : t.a.at<float>(j, i) * al;
}
// alpha * op(A) op(B) + beta_c * C with a transposed operand: double accumulation, one rounding
inline Mat gemm_t(const GemmT& g, const Mat* c, double beta) {
  const int m = g.ta ? g.a.cols : g.a.rows, k = g.ta ? g.a.rows : g.a.cols, n = g.tb ? g.b.rows : g.b.cols;
  Mat d(m, n, CV_32F);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int l = 0; l < k; ++l)
        s += (double)(g.ta ? g.a.at<float>(l, i) : g.a.at<float>(i, l)) * (double)(g.tb ? g.b.at<float>(j, l) : g.b.at<float>(l, j));
      s *= g.alpha;
      if (c) s += beta * (double)c->at<float>(i, j);
      d.at<float>(i, j) = (float)s;
    }
  return d;
}
inline Mat::Mat(const GemmT& g) { *this = gemm_t(g, nullptr, 0.0); }
inline Mat Mat::inv() const {
  double S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i * 3 + j] = at<float>(i, j);
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  Mat r(3, 3, CV_32F);
  if (d == 0) return r;
  d = 1. / d;
  const double t[9] = {(S[4] * S[8] - S[5] * S[7]) * d, (S[2] * S[7] - S[1] * S[8]) * d, (S[1] * S[5] - S[2] * S[4]) * d,
                       (S[5] * S[6] - S[3] * S[8]) * d, (S[0] * S[8] - S[2] * S[6]) * d, (S[2] * S[3] - S[0] * S[5]) * d,
                       (S[3] * S[7] - S[4] * S[6]) * d, (S[1] * S[6] - S[0] * S[7]) * d, (S[0] * S[4] - S[1] * S[3]) * d};
  for (int i = 0; i < 9; ++i) r.at<float>(i / 3, i % 3) = (float)t[i];
  return r;
}

// ---- eager float operators (the small-matrix float path of cv::gemm and elementwise ops) ------------
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat d(a.rows, b.cols, CV_32F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      float s = 0.f;
      for (int l = 0; l < a.cols; ++l) s += a.at<float>(i, l) * b.at<float>(l, j);
      d.at<float>(i, j) = s;
    }
  return d;
}
inline Mat ew(const Mat& a, const Mat& b, float sb) {
  Mat d(a.rows, a.cols, CV_32F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) d.at<float>(i, j) = a.at<float>(i, j) + sb * b.at<float>(i, j);
  return d;
}
inline Mat operator+(const Mat& a, const Mat& b) { return ew(a, b, 1.f); }
inline Mat operator-(const Mat& a, const Mat& b) { return ew(a, b, -1.f); }
inline Mat scale(const Mat& a, float s) {
  Mat d(a.rows, a.cols, CV_32F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) d.at<float>(i, j) = a.at<float>(i, j) * s;
  return d;
}
inline Mat operator-(const Mat& a) { return scale(a, -1.f); }
inline Mat operator*(double s, const Mat& a) { return scale(a, (float)s); }
inline Mat operator*(const Mat& a, double s) { return scale(a, (float)s); }
inline Mat operator/(const Mat& a, double s) { return scale(a, (float)(1.0 / s)); }
// ---- lazy transposes ------------------------------------------------------------------------------
inline MatT operator-(const MatT& t) { return MatT{t.a, -t.alpha}; }
inline MatT operator*(double s, const MatT& t) { return MatT{t.a, t.alpha * s}; }
inline GemmT operator*(const MatT& t, const Mat& b) { return GemmT{t.a, b, true, false, t.alpha}; }
inline GemmT operator*(const Mat& a, const MatT& t) { return GemmT{a, t.a, false, true, t.alpha}; }
inline Mat operator*(const GemmT& g, const Mat& b) { return Mat(g) * b; }
inline Mat operator*(const Mat& a, const GemmT& g) { return a * Mat(g); }
inline GemmT operator-(const GemmT& g) { return GemmT{g.a, g.b, g.ta, g.tb, -g.alpha}; }
inline Mat operator+(const GemmT& g, const Mat& c) { return gemm_t(g, &c, 1.0); }   // fused addend
inline Mat operator+(const Mat& c, const GemmT& g) { return gemm_t(g, &c, 1.0); }
inline Mat operator-(const Mat& c, const GemmT& g) { return gemm_t(GemmT{g.a, g.b, g.ta, g.tb, -g.alpha}, &c, 1.0); }
inline Mat operator-(const GemmT& g, const Mat& c) { return gemm_t(g, &c, -1.0); }

inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }
enum { NORM_L1 = 2 };
// cv::norm(a, b, NORM_L1) of two CV_32F matrices: sum of |a - b| accumulated in double (OpenCV normDiffL1_32f)
inline double norm(const Mat& a, const Mat& b, int) {
  double s = 0;
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) s += std::fabs((double)a.at<float>(i, j) - (double)b.at<float>(i, j));
  return s;
}

// cv::FileStorage / cv::FileNode: only so that DBoW2's YAML save/load members (virtual, hence instantiated with the
// class) compile; they are never called (the vocabulary is read with loadFromTextFile).
class FileNode {
 public:
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator double() const { return 0; }
  operator float() const { return 0; }
  operator std::string() const { return std::string(); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_32F) {}
  Mat_(const Mat& m) : Mat(m) {}
  struct Init {
    Mat_* m;
    int i;
    Init& operator,(T v) { m->template at<T>(i / m->cols, i % m->cols) = v; ++i; return *this; }
    operator Mat() const { return *m; }
    operator Mat_<T>() const { return *m; }
  };
  Init operator<<(T v) {
    this->template at<T>(0, 0) = v;
    return Init{this, 1};
  }
};
}  // namespace cv
