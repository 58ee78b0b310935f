// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// cvprim: from-scratch scalar restatement of the OpenCV 4.13 primitives that the
// reference's ORB front end calls (OpenCV itself is NOT vendored in /root/reference;
// the reference pins only "OpenCV >= 2.4.3", CMakeLists.txt:43-46).  Call sites in the
// reference: src/ORBextractor.cc:81,103 (cvRound, fastAtan2), :810,815 (FAST),
// :1087 (GaussianBlur), :1122 (resize), :1124-1130 (copyMakeBorder).
//
// Every primitive here is pinned bit-exact against cv2 4.13.0 by tests/test_cvprim_vs_cv2.py
// (run where cv2 is importable) and by the committed fixtures in tests/golden/.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may link or load this code.  The product path (multi_orb_slam_b200/csrc) never does.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace cvp {

// cvRound: round-half-to-even (SSE cvtss2si / lrint under the default rounding mode).
int round_f(float v);
int round_d(double v);

// cv::fastAtan2(y, x) in degrees [0,360): 7th-order odd polynomial, float32, no FMA.
float fast_atan2(float y, float x);

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 (11-bit fixed point).
void resize_linear_u8(const uint8_t* src, size_t sstep, int sw, int sh,
                      uint8_t* dst, size_t dstep, int dw, int dh);

// cv::copyMakeBorder(..., BORDER_REFLECT_101): `buf` points at the top-left of a
// (w+2*b) x (h+2*b) buffer whose interior (offset b,b) already holds the w x h image.
void border_reflect101_inplace(uint8_t* buf, size_t step, int w, int h, int b);

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1, OpenCV 4.x
// fixed-point path (taps 18,34,48,56,48,34,18 / 256 per pass).  src may alias dst.
void gaussblur7_sigma2_u8(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep,
                          int w, int h);

struct FastKP { int x, y, score; };
// cv::FAST(img, kps, threshold, nonmaxSuppression, TYPE_9_16).  Output in row-major order.
void fast9_16(const uint8_t* img, size_t step, int w, int h, int threshold, bool nms,
              std::vector<FastKP>& out);

// FAST arc measure m(x,y) = max over 16 cyclic 9-arcs of min(+-diff) (pixel must be >= 3 px
// from every edge).  Corner iff m > threshold, score = m-1.
int fast9_arc_measure(const uint8_t* p, size_t step);

}  // namespace cvp

namespace cvp {
// cv::undistortPoints(src, dst, K, dist, noArray(), K) for CV_32FC2 points with the default
// termination criteria (COUNT, 5 iterations), OpenCV 4.x cvUndistortPointsInternal: everything in
// double, K = [fx 0 cx; 0 fy cy; 0 0 1], dist = (k1, k2, p1, p2, k3); result rounded to float.
void undistort_points(const float* src_xy, int n, float fx, float fy, float cx, float cy, const float* dist5,
                      float* dst_xy);
}  // namespace cvp
