// TEST INFRASTRUCTURE — the oo_* C harness of oracle/ref_capi.cc, here over the PRODUCT's drop-in translation unit
// multi_orb_slam_b200/dropin/ORBextractor_b200.cc (the reference's own class ORB_SLAM2::ORBextractor, declared by the
// reference's include/ORBextractor.h, defined by the drop-in): tests/test_gpu_dropin.py drives the GPU extractor
// through the C++ class exactly as Frame::ExtractORB does (src/Frame.cc:397-403) and compares with the oracle.
#include <cstdlib>
#include <cstring>

#include "ORBextractor.h"  // from /root/reference/include
#include "orb_b200_dropin.h"
#include "orb_oracle.h"

struct oo_extractor {
  int nlevels;
  ORB_SLAM2::ORBextractor* ex;
};

extern "C" {

oo_extractor* oo_create(int nf, float sf, int nl, int ini, int mn) {
  oo_extractor* e = new oo_extractor();
  e->nlevels = nl;
  e->ex = new ORB_SLAM2::ORBextractor(nf, sf, nl, ini, mn);
  ORB_SLAM2::b200::SetPyramidMirror(e->ex, true);
  return e;
}
void oo_destroy(oo_extractor* e) {
  ORB_SLAM2::b200::Release(e->ex);
  delete e->ex;
  delete e;
}

int oo_extract(oo_extractor* e, const uint8_t* img, int rows, int cols, size_t stride, oo_keypoint* kps, uint8_t* desc, int cap,
               int* level_counts) {
  cv::Mat image(rows, cols, CV_8UC1, (void*)img, stride);
  std::vector<cv::KeyPoint> keys;
  cv::Mat descriptors;
  (*e->ex)(image, cv::Mat(), keys, descriptors);
  const int n = (int)keys.size();
  if (level_counts) for (int l = 0; l < e->nlevels; ++l) level_counts[l] = 0;
  if (n > cap) return -1;
  for (int i = 0; i < n; ++i) {
    kps[i].x = keys[i].pt.x; kps[i].y = keys[i].pt.y; kps[i].size = keys[i].size;
    kps[i].angle = keys[i].angle; kps[i].response = keys[i].response; kps[i].octave = keys[i].octave;
    if (level_counts) level_counts[keys[i].octave]++;
    std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
  }
  return n;
}

// mvImagePyramid[level]: the ROI's data pointer / step, like the reference's layout (parent with a 19-px border)
int oo_pyramid_level(oo_extractor* e, int level, const uint8_t** data, int* w, int* h, size_t* step) {
  if (level < 0 || level >= e->nlevels) return -1;
  const cv::Mat& m = e->ex->mvImagePyramid[level];
  if (m.empty()) return -1;
  *data = m.data; *w = m.cols; *h = m.rows; *step = m.step;
  return 0;
}

void oo_scale_tables(oo_extractor* e, float* s, float* is, float* s2, float* is2) {
  std::vector<float> a = e->ex->GetScaleFactors(), b = e->ex->GetInverseScaleFactors(), c = e->ex->GetScaleSigmaSquares(),
                     d = e->ex->GetInverseScaleSigmaSquares();
  for (int i = 0; i < e->nlevels; ++i) {
    if (s) s[i] = a[i];
    if (is) is[i] = b[i];
    if (s2) s2[i] = c[i];
    if (is2) is2[i] = d[i];
  }
}

}  // extern "C"
