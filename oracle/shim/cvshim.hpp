// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// From-scratch OpenCV-API shim: just enough of cv:: (CV_8UC1 Mat with ROI views, KeyPoint,
// resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 / cvRound) for the reference's
// /root/reference/src/ORBextractor.cc to compile VERBATIM (oracle/Makefile, target
// _ref/liborb_ref.so).  The pixel arithmetic lives in oracle/cvprim.cc, pinned bit-exact
// against cv2 4.13.0.  OpenCV itself is not installed as a C++ library in this image.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../cvprim.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

inline int cvRound(double v) { return cvp::round_d(v); }
inline int cvRound(float v) { return cvp::round_f(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
           int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct MatStep {
  size_t v;
  MatStep(size_t s = 0) : v(s) {}
  operator size_t() const { return v; }
};

class Mat;
// Mat::zeros(...) expression: assigning it to a Mat of matching geometry zero-fills that Mat's
// existing storage in place (OpenCV MatOp_Initializer::assign -> create() is a no-op), which
// the reference relies on at ORBextractor.cc:1038,1090-1091.
struct MatZerosExpr { int rows, cols, type; };

class Mat {
 public:
  int rows, cols;
  uchar* data;
  MatStep step;

  Mat() : rows(0), cols(0), data(nullptr), step(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  // wrap external memory (no ownership)
  Mat(int r, int c, int type, void* ext, size_t stp) : rows(r), cols(c), data((uchar*)ext), step(stp) {
    (void)type;
  }

  void create(int r, int c, int type) {
    assert(type == CV_8UC1);
    (void)type;
    if (data && r == rows && c == cols) return;
    rows = r;
    cols = c;
    step = MatStep((size_t)c);
    store_.reset(new std::vector<uchar>((size_t)r * c));
    data = store_->data();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { store_.reset(); data = nullptr; rows = cols = 0; step = MatStep(0); }

  Mat operator()(const Rect& r) const {
    Mat m;
    m.rows = r.height;
    m.cols = r.width;
    m.step = step;
    m.data = data + (size_t)r.y * step + r.x;
    m.store_ = store_;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
    return m;
  }
  Mat& operator=(const MatZerosExpr& e) {
    create(e.rows, e.cols, e.type);
    for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, cols);
    return *this;
  }
  static MatZerosExpr zeros(int r, int c, int type) { return MatZerosExpr{r, c, type}; }

  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const {
    return *(const T*)(data + (size_t)y * step + x * sizeof(T));
  }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  size_t step1() const { return step; }
  int type() const { return CV_8UC1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }

 private:
  std::shared_ptr<std::vector<uchar>> store_;
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_->empty(); }
  Mat getMat() const { return *m_; }
 private:
  const Mat* m_;
};
typedef const _InputArray& InputArray;

class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void release() const { m_->release(); }
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void create(Size s, int type) const { m_->create(s, type); }
  Mat getMat() const { return *m_; }
  Mat& ref() const { return *m_; }
 private:
  Mat* m_;
};
typedef const _OutputArray& OutputArray;

inline float fastAtan2(float y, float x) { return cvp::fast_atan2(y, x); }

inline void resize(InputArray _src, OutputArray _dst, Size dsize, double, double, int interp) {
  assert(interp == INTER_LINEAR);
  (void)interp;
  Mat src = _src.getMat();
  _dst.create(dsize, CV_8UC1);
  Mat dst = _dst.getMat();
  cvp::resize_linear_u8(src.data, src.step, src.cols, src.rows, dst.data, dst.step, dst.cols, dst.rows);
}

inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                           int borderType) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  assert(top == bottom && top == left && top == right);
  (void)borderType; (void)bottom; (void)left; (void)right;
  Mat src = _src.getMat();
  _dst.create(src.rows + 2 * top, src.cols + 2 * top, CV_8UC1);
  Mat dst = _dst.getMat();
  uchar* interior = dst.data + (size_t)top * dst.step + top;
  if (interior != src.data)
    for (int y = 0; y < src.rows; ++y)
      std::memmove(interior + (size_t)y * dst.step, src.data + (size_t)y * src.step, src.cols);
  cvp::border_reflect101_inplace(dst.data, dst.step, src.cols, src.rows, top);
}

inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy, int borderType) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
  (void)ksize; (void)sx; (void)sy; (void)borderType;
  Mat src = _src.getMat();
  _dst.create(src.rows, src.cols, CV_8UC1);
  Mat dst = _dst.getMat();
  cvp::gaussblur7_sigma2_u8(src.data, src.step, dst.data, dst.step, src.cols, src.rows);
}

inline void FAST(InputArray _img, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
  Mat img = _img.getMat();
  std::vector<cvp::FastKP> tmp;  // no statics: the _ref build resets its bump arena per call
  cvp::fast9_16(img.data, img.step, img.cols, img.rows, threshold, nms, tmp);
  kps.clear();
  for (const cvp::FastKP& k : tmp) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
}

struct KeyPointsFilter {
  // only referenced by the reference's dead ComputeKeyPointsOld (ORBextractor.cc:856-1033)
  static void retainBest(std::vector<KeyPoint>&, int) { assert(!"retainBest: dead code path"); }
};

}  // namespace cv
