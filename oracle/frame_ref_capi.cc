// TEST INFRASTRUCTURE — C API over the reference's own src/Frame.cc (with its real include/Frame.h) and
// src/ORBmatcher.cc, both compiled VERBATIM into oracle/_ref/libframe_ref.so (oracle/Makefile; stubs in
// oracle/shim_frame/frame_stubs.hpp).  Frames are built by the reference's two-camera RGB-D constructor
// (src/Frame.cc:148-346) from keypoints / descriptors queued in a stand-in extractor, so UndistortKeyPoints,
// ComputeImageBounds, ComputeStereoFromRGBD, the multi-camera index maps and AssignFeaturesToGrid all run as
// the reference wrote them; the matcher entry points then search through the reference's
// Frame::GetFeaturesInArea.  Used by tests/test_frame_ref.py to pin the restatements.
#include <cstring>
#include <memory>

#include "Frame.h"
#include "ORBmatcher.h"
#include "orb_oracle.h"

using namespace ORB_SLAM2;

namespace ORB_SLAM2 {
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {  // src/MapPoint.cc:602-617
  const float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(std::log(ratio) / pF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
  return nScale;
}
}  // namespace ORB_SLAM2

namespace {
std::vector<cv::KeyPoint> keys_of(const oo_keypoint* k, int n) {
  std::vector<cv::KeyPoint> v(n);
  for (int i = 0; i < n; ++i) {
    v[i].pt.x = k[i].x; v[i].pt.y = k[i].y; v[i].size = k[i].size; v[i].angle = k[i].angle;
    v[i].response = k[i].response; v[i].octave = k[i].octave;
  }
  return v;
}
cv::Mat desc_rows(const uint8_t* d, int n) {
  cv::Mat m(n, 32, CV_8U);
  if (n) std::memcpy(m.data, d, (size_t)n * 32);
  return m;
}
cv::Mat fmat(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) m.at<float>(i, j) = p[i * c + j];
  return m;
}
struct Rig {
  ORBextractor ex0, ex1;
  ORBVocabulary voc;
  cv::Mat K, dist, calib, gray, depth0, depth1;
  Rig(int nlevels, float sf) : ex0(1000, sf, nlevels, 20, 7), ex1(500, sf, nlevels, 20, 7) {}
};
// the reference's constructor on flat inputs; cols x rows only matter through ComputeImageBounds
std::unique_ptr<Frame> make_frame(Rig& rig, const oo_keypoint* k0, const uint8_t* d0, int n0, const oo_keypoint* k1,
                                  const uint8_t* d1, int n1, const float* depth0, const float* depth1, int cols, int rows,
                                  float fx, float fy, float cx, float cy, const float* dist, int n_dist, float bf, float th_depth,
                                  const float* calib) {
  rig.ex0.next_keys = keys_of(k0, n0); rig.ex0.next_desc = desc_rows(d0, n0);
  rig.ex1.next_keys = keys_of(k1, n1); rig.ex1.next_desc = desc_rows(d1, n1);
  const float Kf[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  rig.K = fmat(Kf, 3, 3);
  rig.dist = fmat(dist, n_dist, 1);
  static const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  rig.calib = fmat(calib ? calib : ident, 4, 3);
  rig.gray = cv::Mat(rows, cols, CV_8U);
  rig.depth0 = cv::Mat(rows, cols, CV_32F);
  rig.depth1 = cv::Mat(rows, cols, CV_32F);
  if (depth0) std::memcpy(rig.depth0.data, depth0, sizeof(float) * (size_t)rows * cols);
  if (depth1) std::memcpy(rig.depth1.data, depth1, sizeof(float) * (size_t)rows * cols);
  Frame::mbInitialComputations = true;  // image bounds and grid cell sizes are statics computed by the first frame
  return std::unique_ptr<Frame>(new Frame(rig.gray, rig.depth0, rig.gray, rig.depth1, 0.0, &rig.ex0, &rig.ex1, &rig.voc, rig.K,
                                          rig.dist, bf, th_depth, rig.calib));
}
}  // namespace

extern "C" {

// Frame::ComputeStereoMatches — the reference's own text (commented out in the fork, src/Frame.cc:782-956; un-commented
// at build time into oracle/_ref/ by oracle/gen_stereo_ref.py, never copied into the repository) run on a default-
// constructed real Frame whose members are set from flat arrays.  pyr_l / pyr_r: mvImagePyramid of the left / right
// extractor (interior pointers into bordered level buffers, like the reference's ROIs).
void ofr_compute_stereo_matches(const oo_keypoint* kl, const uint8_t* dl, int nl, const oo_keypoint* kr, const uint8_t* dr, int nr,
                                const om_image* pyr_l, const om_image* pyr_r, int nlevels, const float* scale_factors,
                                const float* inv_scale_factors, float mbf, float mb, float* uright, float* depth) {
  ORBextractor exl(1000, 1.2f, nlevels, 20, 7), exr(1000, 1.2f, nlevels, 20, 7);
  for (int l = 0; l < nlevels; ++l) {
    exl.mvImagePyramid.push_back(cv::Mat::wrap(pyr_l[l].h, pyr_l[l].w, CV_8U, (void*)pyr_l[l].data, pyr_l[l].step));
    exr.mvImagePyramid.push_back(cv::Mat::wrap(pyr_r[l].h, pyr_r[l].w, CV_8U, (void*)pyr_r[l].data, pyr_r[l].step));
  }
  Frame F;
  F.N = nl;
  F.mvKeys = keys_of(kl, nl);
  F.mvKeysRight = keys_of(kr, nr);
  F.mDescriptors = desc_rows(dl, nl);
  F.mDescriptorsRight = desc_rows(dr, nr);
  F.mpORBextractorLeft = &exl;
  F.mpORBextractorRight = &exr;
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  F.mvInvScaleFactors.assign(inv_scale_factors, inv_scale_factors + nlevels);
  F.mb = mb;
  F.mbf = mbf;
  F.ComputeStereoMatches();
  for (int i = 0; i < nl; ++i) { uright[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

// Frame::Frame (two cameras, RGB-D) -> everything the glue computes.  Outputs sized n0 + n1: k_un (mvKeysUn_total),
// uright / depth (mvuRight_total / mvDepth_total), cam_of (keypoint_to_cam), local_idx (cont_idx_to_local_cam_idx);
// bounds = mnMinX..; grid_start (2 x (64*48+1)) / grid_items (2 x (n0+n1)): mGrids[cam][ix][iy] as CSR over ix*48+iy.
int ofr_frame_glue(const oo_keypoint* k0, const uint8_t* d0, int n0, const oo_keypoint* k1, const uint8_t* d1, int n1,
                   const float* depth0, const float* depth1, int cols, int rows, float fx, float fy, float cx, float cy,
                   const float* dist, int n_dist, float bf, int nlevels, float scale_factor, oo_keypoint* k_un, float* uright,
                   float* depth_out, int32_t* cam_of, int32_t* local_idx, om_bounds* bounds, int32_t* grid_start,
                   int32_t* grid_items) {
  if (dist[0] == 0.f) return -1;
  Rig rig(nlevels, scale_factor);
  std::unique_ptr<Frame> F = make_frame(rig, k0, d0, n0, k1, d1, n1, depth0, depth1, cols, rows, fx, fy, cx, cy, dist, n_dist, bf,
                                        40.f, nullptr);
  const int n = n0 + n1;
  if (F->N_total != n) return -1;
  for (int i = 0; i < n; ++i) {
    const cv::KeyPoint& kp = F->mvKeysUn_total[i];
    k_un[i] = oo_keypoint{kp.pt.x, kp.pt.y, kp.size, kp.angle, kp.response, kp.octave};
    uright[i] = F->mvuRight_total[i];
    depth_out[i] = F->mvDepth_total[i];
    cam_of[i] = F->keypoint_to_cam[i];
    local_idx[i] = F->cont_idx_to_local_cam_idx[i];
  }
  bounds->min_x = Frame::mnMinX; bounds->max_x = Frame::mnMaxX; bounds->min_y = Frame::mnMinY; bounds->max_y = Frame::mnMaxY;
  for (int c = 0; c < 2; ++c) {
    int run = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
      for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
        grid_start[c * (FRAME_GRID_COLS * FRAME_GRID_ROWS + 1) + ix * FRAME_GRID_ROWS + iy] = run;
        for (size_t idx : F->mGrids[c][ix][iy]) grid_items[c * n + run++] = (int32_t)idx;
      }
    grid_start[c * (FRAME_GRID_COLS * FRAME_GRID_ROWS + 1) + FRAME_GRID_COLS * FRAME_GRID_ROWS] = run;
  }
  return n;
}

// NOTE: with k1 == 0 the reference's UndistortKeyPoints returns before it fills mvKeysUn_total (src/Frame.cc:676-680 vs
// :697-704), so AssignFeaturesToGrid then indexes an empty vector — the two-camera constructor is undefined behaviour for
// undistorted cameras.  Every entry point here therefore takes a distortion vector with k1 != 0.

// Frame::GetFeaturesInArea(cam, x, y, r, minLevel, maxLevel) (src/Frame.cc:574-630) of a two-camera frame.
int ofr_features_in_area(const oo_keypoint* k0, int n0, const oo_keypoint* k1, int n1, int cols, int rows, float fx, float fy,
                         float cx, float cy, const float* dist, int n_dist, int cam, float x, float y, float r, int min_level,
                         int max_level, int32_t* out, int cap) {
  if (dist[0] == 0.f) return -1;
  Rig rig(8, 1.2f);
  std::vector<uint8_t> z((size_t)std::max(std::max(n0, n1), 1) * 32, 0);
  std::unique_ptr<Frame> F = make_frame(rig, k0, z.data(), n0, k1, z.data(), n1, nullptr, nullptr, cols, rows, fx, fy, cx, cy, dist,
                                        n_dist, 40.f, 40.f, nullptr);
  const std::vector<size_t> v = F->GetFeaturesInArea(cam, x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:747-878) on two reference-built frames; it reads mvKeysUn and
// searches through Frame::GetFeaturesInArea with the bounds the constructor computed.
int ofr_search_for_initialization(const oo_keypoint* k1, const uint8_t* d1, int n1, const oo_keypoint* k2, const uint8_t* d2,
                                  int n2, int cols, int rows, float fx, float fy, float cx, float cy, const float* dist,
                                  int n_dist, float* prev_xy, int window, float nnratio, int check_ori, int* matches12) {
  if (dist[0] == 0.f) return -1;
  Rig r1(8, 1.2f), r2(8, 1.2f);
  std::unique_ptr<Frame> F1 = make_frame(r1, k1, d1, n1, nullptr, nullptr, 0, nullptr, nullptr, cols, rows, fx, fy, cx, cy, dist,
                                         n_dist, 40.f, 40.f, nullptr);
  std::unique_ptr<Frame> F2 = make_frame(r2, k2, d2, n2, nullptr, nullptr, 0, nullptr, nullptr, cols, rows, fx, fy, cx, cy, dist,
                                         n_dist, 40.f, 40.f, nullptr);
  std::vector<cv::Point2f> prev(n1);
  for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchForInitialization(*F1, *F2, prev, m12, window);
  for (int i = 0; i < n1; ++i) {
    matches12[i] = m12[i];
    prev_xy[2 * i] = prev[i].x;
    prev_xy[2 * i + 1] = prev[i].y;
  }
  return n;
}

}  // extern "C"
