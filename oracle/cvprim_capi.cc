// TEST INFRASTRUCTURE — C entry points so tests can drive the cvprim primitives directly
// (tests/test_cvprim_vs_cv2.py pins each one bit-exact against cv2 4.13.0).
#include <cmath>
#include <vector>
#include "cvprim.h"
extern "C" {
void cvp_resize(const uint8_t* s, size_t ss, int sw, int sh, uint8_t* d, size_t ds, int dw, int dh) {
  cvp::resize_linear_u8(s, ss, sw, sh, d, ds, dw, dh);
}
void cvp_border(uint8_t* buf, size_t step, int w, int h, int b) { cvp::border_reflect101_inplace(buf, step, w, h, b); }
void cvp_blur(const uint8_t* s, size_t ss, uint8_t* d, size_t ds, int w, int h) {
  cvp::gaussblur7_sigma2_u8(s, ss, d, ds, w, h);
}
int cvp_fast(const uint8_t* img, size_t step, int w, int h, int th, int nms, int* xys, int cap) {
  std::vector<cvp::FastKP> v;
  cvp::fast9_16(img, step, w, h, th, nms != 0, v);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
  return (int)v.size();
}
void cvp_atan2(const float* y, const float* x, float* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = cvp::fast_atan2(y[i], x[i]);
}
void cvp_sincos_libm(const float* a, float* c, float* s, int n) {
  for (int i = 0; i < n; ++i) { c[i] = cosf(a[i]); s[i] = sinf(a[i]); }
}
int cvp_round(float v) { return cvp::round_f(v); }
}

extern "C" void cvp_undistort_points(const float* src_xy, int n, float fx, float fy, float cx, float cy, const float* dist5,
                                     float* dst_xy) {
  cvp::undistort_points(src_xy, n, fx, fy, cx, cy, dist5, dst_xy);
}
