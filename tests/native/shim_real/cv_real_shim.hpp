// TEST INFRASTRUCTURE — the few OpenCV names the reference's REAL headers (include/Frame.h, KeyFrame.h, MapPoint.h,
// Map.h, KeyFrameDatabase.h, ORBextractor.h, ORBmatcher.h) mention beyond what oracle/shim_matcher/cvm.hpp already
// provides, so that the product's drop-in translation units can be COMPILED against those headers unmodified
// (tests/test_dropin_build.py).  Declarations only need to parse: nothing here is executed.
#pragma once
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>
#include <vector>

#include "cvm.hpp"

namespace cv {
typedef Point_<int> Point2i;
struct Vec3f { float v[3]; float& operator[](int i) { return v[i]; } const float& operator[](int i) const { return v[i]; } };
// cv::InputArray / cv::OutputArray as OpenCV has them: proxy classes (the reference's ORBextractor.h names them in
// operator()'s signature; Frame::ExtractORB passes cv::Mat objects, src/Frame.cc:397-403)
class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  Mat getMat() const { return *m_; }
  bool empty() const { return m_->empty(); }
 private:
  const Mat* m_;
};
typedef const _InputArray& InputArray;
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  Mat getMat() const { return *m_; }
  void release() const { m_->release(); }
 private:
  Mat* m_;
};
typedef const _OutputArray& OutputArray;
}  // namespace cv
