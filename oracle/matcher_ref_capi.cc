// TEST INFRASTRUCTURE — C API over the reference's own src/ORBmatcher.cc, compiled VERBATIM from
// /root/reference against oracle/shim_matcher (OpenCV-API shim + Frame/KeyFrame/MapPoint stand-ins);
// built into oracle/_ref/libmatcher_ref.so by oracle/Makefile.  Every omr_* function takes the same flat
// arguments as the om_* restatement of the same reference function (oracle/orb_oracle.h), builds the
// object graph the reference expects, calls the reference and flattens the result, so that
// tests/test_matcher_ref.py can pin the restatement (and through it the GPU kernels) to the reference code.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

#include "ORBmatcher.h"
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#include "orb_oracle.h"

using namespace ORB_SLAM2;

float Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;

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
  cv::Mat m(std::max(n, 1), 32, CV_8U);
  if (n) std::memcpy(m.data, d, (size_t)n * 32);
  return m;
}
cv::Mat fmat(const float* p, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) m.at<float>(i, j) = p[i * c + j];
  return m;
}
void set_bounds(om_bounds b) {
  Frame::mnMinX = b.min_x; Frame::mnMaxX = b.max_x; Frame::mnMinY = b.min_y; Frame::mnMaxY = b.max_y;
}
#ifndef OMR_HOT_ONLY
DBoW2::FeatureVector featvec(const int32_t* node, const int32_t* start, const int32_t* items, int nn) {
  DBoW2::FeatureVector fv;
  for (int a = 0; a < nn; ++a)
    for (int p = start[a]; p < start[a + 1]; ++p) fv.addFeature((DBoW2::NodeId)node[a], (unsigned)items[p]);
  return fv;
}
#endif
// concatenated multi-camera features -> the members a Frame / KeyFrame holds for them
template <class T>
void fill_rig(T& f, const oo_keypoint* k, const uint8_t* d, const int32_t* cam, const float* uright, int n, om_bounds b) {
  f.mvKeysUn_total = f.mvKeys_total = keys_of(k, n);
  f.N_total = n;
  f.mvuRight_total.assign(n, -1.f);
  if (uright) f.mvuRight_total.assign(uright, uright + n);
  int cnt[2] = {0, 0};
  for (int i = 0; i < n; ++i) {
    const int c = cam ? cam[i] : 0;
    f.keypoint_to_cam[i] = c;
    f.cont_idx_to_local_cam_idx[i] = cnt[c]++;
  }
  f.mDescriptors_total.assign(2, cv::Mat());
  for (int c = 0; c < 2; ++c) f.mDescriptors_total[c] = cv::Mat(std::max(cnt[c], 1), 32, CV_8U);
  for (int i = 0; i < n; ++i)
    std::memcpy(f.mDescriptors_total[f.keypoint_to_cam[i]].ptr(f.cont_idx_to_local_cam_idx[i]), d + (size_t)i * 32, 32);
  f.N = cnt[0];  // camera-1 features come first in the reference's numbering
  f.mvKeysUn.assign(f.mvKeysUn_total.begin(), f.mvKeysUn_total.begin() + std::min(n, f.N));
  f.mvKeys = f.mvKeysUn;
  f.mvuRight.assign(f.mvuRight_total.begin(), f.mvuRight_total.begin() + std::min(n, f.N));
  f.mDescriptors = f.mDescriptors_total[0];
  f.grids.build(f.mvKeysUn_total, f.keypoint_to_cam, 2, b.min_x, b.max_x, b.min_y, b.max_y);
  f.grid1.build(f.mvKeysUn_total, std::unordered_map<size_t, int>(), 1, b.min_x, b.max_x, b.min_y, b.max_y);
}
void set_levels(Frame& f, const float* sf, int nlevels, float log_sf) {
  f.mvScaleFactors.assign(sf, sf + nlevels);
  f.mnScaleLevels = nlevels;
  f.mfLogScaleFactor = log_sf;
}
void set_levels(KeyFrame& f, const float* sf, int nlevels, float log_sf) {
  f.mvScaleFactors.assign(sf, sf + nlevels);
  f.mnScaleLevels = nlevels;
  f.mfLogScaleFactor = log_sf;
}
void set_camera(Frame& f, om_camera c) { f.fx = c.fx; f.fy = c.fy; f.cx = c.cx; f.cy = c.cy; f.mb = c.mb; f.mbf = c.mbf; }
void set_camera(KeyFrame& f, om_camera c, om_bounds b) {
  f.fx = c.fx; f.fy = c.fy; f.cx = c.cx; f.cy = c.cy; f.mb = c.mb; f.mbf = c.mbf;
  f.mnMinX = (int)b.min_x; f.mnMaxX = (int)b.max_x; f.mnMinY = (int)b.min_y; f.mnMaxY = (int)b.max_y;
}
MapPoint make_point(const float* xyz, const float* normal, float max_inv, float min_inv, float max_d, const uint8_t* desc,
                    int n_obs, bool bad) {
  MapPoint p;
  if (xyz) p.worldPos = fmat(xyz, 3, 1);
  if (normal) p.normal = fmat(normal, 3, 1);
  p.maxInvariance = max_inv; p.minInvariance = min_inv; p.mfMaxDistance = max_d;
  if (desc) p.descriptor = desc_rows(desc, 1);
  p.nObs = n_obs; p.bad = bad;
  return p;
}
int index_in(const std::vector<MapPoint>& pool, const MapPoint* p) {
  return (p >= pool.data() && p < pool.data() + pool.size()) ? (int)(p - pool.data()) : -1;
}
}  // namespace

int g_cam1 = 0;  // != 0: the wrappers call the reference's camera-1-only twins (_cam1) instead

extern "C" {

// Select the `_cam1` twins (SearchByProjection_cam1(KeyFrame*, Scw, ...) :753-867, SearchByBoW_cam1 :390-565 and
// :1180-1363, Fuse_cam1 :2518-2813, SearchBySim3_cam1 :3137-3433) for the following omr_* calls.
void omr_set_cam1(int on) { g_cam1 = on; }

int omr_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_rows(a, 1), desc_rows(b, 1)); }

int omr_search_for_initialization(const oo_keypoint* k1, const uint8_t* d1, int n1, const oo_keypoint* k2, const uint8_t* d2,
                                  int n2, om_bounds b2, float* prev_xy, int window, float nnratio, int check_ori,
                                  int* matches12) {
  set_bounds(b2);
  Frame F1, F2;
  fill_rig(F1, k1, d1, nullptr, nullptr, n1, b2);
  fill_rig(F2, k2, d2, nullptr, nullptr, n2, b2);
  std::vector<cv::Point2f> prev(n1);
  for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window);
  for (int i = 0; i < n1; ++i) {
    matches12[i] = m12[i];
    prev_xy[2 * i] = prev[i].x;
    prev_xy[2 * i + 1] = prev[i].y;
  }
  return n;
}

int omr_search_by_projection_points(const oo_keypoint* k, const uint8_t* d, const float* u_right, int n, om_bounds b,
                                    const float* scale_factors, int nlevels, const om_mappoint* mp, const uint8_t* mp_desc,
                                    const int* mp_obs, int nmp, float th, float nnratio, int* frame_mp,
                                    const int* frame_mp_obs) {
  set_bounds(b);
  Frame F;
  fill_rig(F, k, d, nullptr, u_right, n, b);
  set_levels(F, scale_factors, nlevels, 0.f);
  std::vector<MapPoint> held(n), pts(nmp);
  F.mvpMapPoints.assign(n, nullptr);
  for (int i = 0; i < n; ++i)
    if (frame_mp[i] >= 0) {
      held[i].nObs = frame_mp_obs ? (frame_mp_obs[i] > 0) : 0;
      F.mvpMapPoints[i] = &held[i];
    }
  std::vector<MapPoint*> vp(nmp);
  for (int i = 0; i < nmp; ++i) {
    pts[i] = make_point(nullptr, nullptr, 0, 0, 0, mp_desc + (size_t)i * 32, mp_obs ? (mp_obs[i] > 0) : 0, mp[i].bad != 0);
    pts[i].mbTrackInView = mp[i].track_in_view != 0;
    pts[i].mnTrackScaleLevel = mp[i].level;
    pts[i].mTrackViewCos = mp[i].view_cos;
    pts[i].mTrackProjX = mp[i].proj_x; pts[i].mTrackProjY = mp[i].proj_y; pts[i].mTrackProjXR = mp[i].proj_xr;
    vp[i] = &pts[i];
  }
  ORBmatcher matcher(nnratio, true);
  const int nm = matcher.SearchByProjection(F, vp, th);
  for (int i = 0; i < n; ++i) {
    const int j = index_in(pts, F.mvpMapPoints[i]);
    if (j >= 0) frame_mp[i] = j;
    else if (!F.mvpMapPoints[i]) frame_mp[i] = -1;
  }
  return nm;
}

int omr_search_by_projection_frame(const oo_keypoint* cur_k, const uint8_t* cur_desc, const float* cur_uright,
                                   const int32_t* cur_cam, int n_cur, om_bounds b, const float* scale_factors, int nlevels,
                                   om_camera cam, const float* Tcw_cur, const float* Tcw_last, const oo_keypoint* last_k,
                                   const int32_t* last_cam, const int32_t* last_valid, const float* last_xyz,
                                   const uint8_t* last_desc, const int32_t* last_obs, int n_last, const float* calib, float th,
                                   int mono, int check_ori, int32_t* cur_mp, const int32_t* cur_mp_obs) {
  set_bounds(b);
  Frame Cur, Last;
  fill_rig(Cur, cur_k, cur_desc, cur_cam, cur_uright, n_cur, b);
  set_levels(Cur, scale_factors, nlevels, 0.f);
  set_camera(Cur, cam);
  Cur.mTcw = fmat(Tcw_cur, 4, 4);
  std::vector<uint8_t> zeros((size_t)std::max(n_last, 1) * 32, 0);
  fill_rig(Last, last_k, zeros.data(), last_cam, nullptr, n_last, b);
  Last.mTcw = fmat(Tcw_last, 4, 4);
  std::vector<MapPoint> held(n_cur), pts(n_last);
  Cur.mvpMapPoints.assign(n_cur, nullptr);
  for (int i = 0; i < n_cur; ++i)
    if (cur_mp[i] >= 0) {
      held[i].nObs = cur_mp_obs ? (cur_mp_obs[i] > 0) : 0;
      Cur.mvpMapPoints[i] = &held[i];
    }
  Last.mvpMapPoints.assign(n_last, nullptr);
  Last.mvbOutlier.assign(n_last, false);
  for (int i = 0; i < n_last; ++i)
    if (last_valid[i]) {
      pts[i] = make_point(last_xyz + 3 * i, nullptr, 0, 0, 0, last_desc + (size_t)i * 32, last_obs ? (last_obs[i] > 0) : 0, false);
      Last.mvpMapPoints[i] = &pts[i];
    }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int nm = matcher.SearchByProjection(Cur, Last, th, mono != 0, fmat(calib, 4, 3));
  for (int i = 0; i < n_cur; ++i) {
    const int j = index_in(pts, Cur.mvpMapPoints[i]);
    if (j >= 0) cur_mp[i] = j;
    else if (!Cur.mvpMapPoints[i]) cur_mp[i] = -1;
  }
  return nm;
}

int omr_search_by_projection_keyframe(const oo_keypoint* cur_k, const uint8_t* cur_desc, int n_cur, om_bounds b,
                                      const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam,
                                      const float* Tcw_cur, const int32_t* kf_valid, const float* kf_xyz,
                                      const float* kf_max_dist, const float* kf_min_dist, const float* kf_max_d,
                                      const float* kf_angle, const uint8_t* kf_desc, int n_kf, float th, int orb_dist,
                                      int check_ori, int32_t* cur_mp) {
  set_bounds(b);
  Frame Cur;
  fill_rig(Cur, cur_k, cur_desc, nullptr, nullptr, n_cur, b);
  set_levels(Cur, scale_factors, nlevels, log_scale_factor);
  set_camera(Cur, cam);
  Cur.mTcw = fmat(Tcw_cur, 4, 4);
  std::vector<MapPoint> held(n_cur), pts(n_kf);
  Cur.mvpMapPoints.assign(n_cur, nullptr);
  for (int i = 0; i < n_cur; ++i)
    if (cur_mp[i] >= 0) Cur.mvpMapPoints[i] = &held[i];
  KeyFrame KF;
  KF.N = KF.N_total = n_kf;
  KF.mvKeysUn.resize(n_kf);
  KF.mvpMapPoints.assign(n_kf, nullptr);
  for (int i = 0; i < n_kf; ++i) {
    KF.mvKeysUn[i].angle = kf_angle[i];
    if (kf_valid[i]) {
      pts[i] = make_point(kf_xyz + 3 * i, nullptr, kf_max_dist[i], kf_min_dist[i], kf_max_d[i], kf_desc + (size_t)i * 32, 1, false);
      KF.mvpMapPoints[i] = &pts[i];
    }
  }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int nm = matcher.SearchByProjection(Cur, &KF, std::set<MapPoint*>(), th, orb_dist);
  for (int i = 0; i < n_cur; ++i) {
    const int j = index_in(pts, Cur.mvpMapPoints[i]);
    if (j >= 0) cur_mp[i] = j;
    else if (!Cur.mvpMapPoints[i]) cur_mp[i] = -1;
  }
  return nm;
}

int omr_search_by_projection_sim3(const oo_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf, om_bounds b,
                                  const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam,
                                  const float* Scw, const float* calib, const int32_t* mp_valid, const float* mp_xyz,
                                  const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist,
                                  const float* mp_max_d, const uint8_t* mp_desc, int n_mp, int th, int32_t* matched) {
  set_bounds(b);
  KeyFrame KF;
  fill_rig(KF, kf_k, kf_desc, kf_cam, nullptr, n_kf, b);
  set_levels(KF, scale_factors, nlevels, log_scale_factor);
  set_camera(KF, cam, b);
  std::vector<MapPoint> held(n_kf), pts(n_mp);
  std::vector<MapPoint*> vpPoints(n_mp), vpMatched(n_kf, nullptr);
  for (int i = 0; i < n_kf; ++i)
    if (matched[i] >= 0) vpMatched[i] = &held[i];
  for (int i = 0; i < n_mp; ++i) {
    pts[i] = make_point(mp_xyz + 3 * i, mp_normal + 3 * i, mp_max_dist[i], mp_min_dist[i], mp_max_d[i], mp_desc + (size_t)i * 32, 1,
                        !mp_valid[i]);
    vpPoints[i] = &pts[i];
  }
  std::vector<int> cams(n_mp, 0);
  ORBmatcher matcher(0.75f, true);
#ifdef OMR_HOT_ONLY
  const int nm = matcher.SearchByProjection(&KF, fmat(Scw, 4, 4), vpPoints, cams, vpMatched, th, fmat(calib, 4, 3));
#else
  const int nm = g_cam1 ? matcher.SearchByProjection_cam1(&KF, fmat(Scw, 4, 4), vpPoints, vpMatched, th)
                        : matcher.SearchByProjection(&KF, fmat(Scw, 4, 4), vpPoints, cams, vpMatched, th, fmat(calib, 4, 3));
#endif
  for (int i = 0; i < n_kf; ++i) {
    const int j = index_in(pts, vpMatched[i]);
    if (j >= 0) matched[i] = j;
    else if (!vpMatched[i]) matched[i] = -1;
  }
  return nm;
}

#ifndef OMR_HOT_ONLY  // the same harness is linked against the product's drop-in translation unit, which defines the hot
                      // members only (multi_orb_slam_b200/dropin/ORBmatcher_b200.cc, tests/native/Makefile)
// variant 0: SearchByBoW(KeyFrame*, Frame&, ...) (valid2 ignored: the frame side has no validity test);
// variant 1: SearchByBoW(KeyFrame*, KeyFrame*, ...)
int omr_search_by_bow(int variant, const uint8_t* d1, const float* angle1, const int32_t* valid1, int n1, const int32_t* node1,
                      const int32_t* start1, const int32_t* items1, int nn1, const uint8_t* d2, const float* angle2,
                      const int32_t* valid2, int n2, const int32_t* node2, const int32_t* start2, const int32_t* items2, int nn2,
                      float nnratio, int check_ori, int32_t* matches12, int32_t* matches21) {
  const om_bounds b = {0.f, 640.f, 0.f, 480.f};
  std::vector<oo_keypoint> k1(std::max(n1, 1)), k2(std::max(n2, 1));
  for (int i = 0; i < n1; ++i) k1[i] = oo_keypoint{1.f, 1.f, 31.f, angle1[i], 0.f, 0};
  for (int i = 0; i < n2; ++i) k2[i] = oo_keypoint{1.f, 1.f, 31.f, angle2[i], 0.f, 0};
  KeyFrame KF1;
  fill_rig(KF1, k1.data(), d1, nullptr, nullptr, n1, b);
  KF1.mFeatVec = KF1.mFeatVec_cam1 = featvec(node1, start1, items1, nn1);
  std::vector<MapPoint> p1(n1), p2(n2);
  KF1.mvpMapPoints.assign(n1, nullptr);
  for (int i = 0; i < n1; ++i)
    if (!valid1 || valid1[i]) KF1.mvpMapPoints[i] = &p1[i];
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  for (int i = 0; i < n2; ++i) matches21[i] = -1;
  ORBmatcher matcher(nnratio, check_ori != 0);
  int nm;
  if (variant == 0) {
    Frame F;
    fill_rig(F, k2.data(), d2, nullptr, nullptr, n2, b);
    F.mFeatVec = F.mFeatVec_cam1 = featvec(node2, start2, items2, nn2);
    std::vector<MapPoint*> out;
    nm = g_cam1 ? matcher.SearchByBoW_cam1(&KF1, F, out) : matcher.SearchByBoW(&KF1, F, out);
    for (int i2 = 0; i2 < n2; ++i2) {
      const int i1 = index_in(p1, out[i2]);
      if (i1 >= 0) { matches21[i2] = i1; matches12[i1] = i2; }
    }
  } else {
    KeyFrame KF2;
    fill_rig(KF2, k2.data(), d2, nullptr, nullptr, n2, b);
    KF2.mFeatVec = KF2.mFeatVec_cam1 = featvec(node2, start2, items2, nn2);
    KF2.mvpMapPoints.assign(n2, nullptr);
    for (int i = 0; i < n2; ++i)
      if (!valid2 || valid2[i]) KF2.mvpMapPoints[i] = &p2[i];
    std::vector<MapPoint*> out;
    nm = g_cam1 ? matcher.SearchByBoW_cam1(&KF1, &KF2, out) : matcher.SearchByBoW(&KF1, &KF2, out);
    for (int i1 = 0; i1 < n1; ++i1) {
      const int i2 = index_in(p2, out[i1]);
      if (i2 >= 0) { matches12[i1] = i2; matches21[i2] = i1; }
    }
  }
  return nm;
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1364-1720).  Poses instead of the fundamental matrices (the
// reference recomputes them, :1376-1449): T1w / T2w = Tcw of camera 1 then camera 2 (2 x 16 floats), Cw1 = camera
// centres of key frame 1 (2 x 3).  F12s_out / epipoles_out return what those lines compute, evaluated here with the
// same expressions on the same shim, for feeding the flat-array implementations.
int omr_search_for_triangulation(const oo_keypoint* k1, const uint8_t* d1, const int32_t* has_mp1, const int32_t* cam1,
                                 const float* uright1, int n1, const int32_t* node1, const int32_t* start1,
                                 const int32_t* items1, int nn1, const oo_keypoint* k2, const uint8_t* d2,
                                 const int32_t* has_mp2, const int32_t* cam2, const float* uright2, int n2,
                                 const int32_t* node2, const int32_t* start2, const int32_t* items2, int nn2,
                                 const float* T1w, const float* T2w, const float* Cw1, om_camera cam, const float* scale_factors2,
                                 const float* level_sigma2_2, int nlevels, int only_stereo, const int32_t* cam_enabled,
                                 int check_ori, int32_t* matches12, float* F12s_out, float* epipoles_out) {
  const om_bounds b = {0.f, 640.f, 0.f, 480.f};
  KeyFrame KF1, KF2;
  fill_rig(KF1, k1, d1, cam1, uright1, n1, b);
  fill_rig(KF2, k2, d2, cam2, uright2, n2, b);
  KF1.mFeatVec = featvec(node1, start1, items1, nn1);
  KF2.mFeatVec = featvec(node2, start2, items2, nn2);
  std::vector<MapPoint> p1(n1), p2(n2);
  KF1.mvpMapPoints.assign(n1, nullptr);
  KF2.mvpMapPoints.assign(n2, nullptr);
  for (int i = 0; i < n1; ++i) if (has_mp1[i]) KF1.mvpMapPoints[i] = &p1[i];
  for (int i = 0; i < n2; ++i) if (has_mp2[i]) KF2.mvpMapPoints[i] = &p2[i];
  const float Kf[9] = {cam.fx, 0, cam.cx, 0, cam.fy, cam.cy, 0, 0, 1};
  KF1.mK = fmat(Kf, 3, 3); KF2.mK = fmat(Kf, 3, 3);
  set_camera(KF1, cam, b); set_camera(KF2, cam, b);
  KF1.Tcw = fmat(T1w, 4, 4); KF1.Tcw_cam2 = fmat(T1w + 16, 4, 4);
  KF2.Tcw = fmat(T2w, 4, 4); KF2.Tcw_cam2 = fmat(T2w + 16, 4, 4);
  KF1.Ow = fmat(Cw1, 3, 1); KF1.Ow_cam2 = fmat(Cw1 + 3, 3, 1);
  KF2.mvScaleFactors.assign(scale_factors2, scale_factors2 + nlevels);
  KF2.mvLevelSigma2.assign(level_sigma2_2, level_sigma2_2 + nlevels);
  // the quantities of :1376-1449, same expressions
  for (int c = 0; c < 2; ++c) {
    cv::Mat R1w = c ? KF1.GetRotation_cam2() : KF1.GetRotation(), t1w = c ? KF1.GetTranslation_cam2() : KF1.GetTranslation();
    cv::Mat R2w = c ? KF2.GetRotation_cam2() : KF2.GetRotation(), t2w = c ? KF2.GetTranslation_cam2() : KF2.GetTranslation();
    cv::Mat R12 = R1w * R2w.t();
    cv::Mat t12 = -R1w * R2w.t() * t2w + t1w;
    cv::Mat t12x = (cv::Mat_<float>(3, 3) << 0, -t12.at<float>(2), t12.at<float>(1), t12.at<float>(2), 0, -t12.at<float>(0),
                    -t12.at<float>(1), t12.at<float>(0), 0);
    cv::Mat F12 = KF1.mK.t().inv() * t12x * R12 * KF2.mK.inv();
    for (int i = 0; i < 9; ++i) F12s_out[9 * c + i] = F12.at<float>(i / 3, i % 3);
    cv::Mat Cw = c ? KF1.GetCameraCenter_cam2() : KF1.GetCameraCenter();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    epipoles_out[2 * c] = KF2.fx * C2.at<float>(0) * invz + KF2.cx;
    epipoles_out[2 * c + 1] = KF2.fy * C2.at<float>(1) * invz + KF2.cy;
  }
  std::vector<std::pair<size_t, size_t>> pairs;
  std::vector<bool> vbCam = {cam_enabled[0] != 0, cam_enabled[1] != 0};
  ORBmatcher matcher(0.6f, check_ori != 0);
  const int nm = matcher.SearchForTriangulation(&KF1, &KF2, cv::Mat(), pairs, only_stereo != 0, vbCam);
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  for (auto& pr : pairs) matches12[pr.first] = (int32_t)pr.second;
  return nm;
}

// ORBmatcher::Fuse, both two-camera overloads (src/ORBmatcher.cc:1986-2190 with Tcw / Ow; :2211-2441 with sim3 != 0,
// where pose = Scw and Ow is ignored).  kf_held[i] != 0: the key-frame feature already holds a map point (with more
// observations than the candidates when kf_held[i] == 2).  best_idx as om_fuse, read back from a call trace.
int omr_fuse(int sim3, const oo_keypoint* kf_k, const uint8_t* kf_desc, const float* kf_uright, const int32_t* kf_cam,
             const int32_t* kf_held, int n_kf, om_bounds b, const float* scale_factors, const float* inv_level_sigma2, int nlevels,
             float log_scale_factor, om_camera cam, const float* pose, const float* Ow, const float* calib,
             const int32_t* mp_valid, const float* mp_xyz, const float* mp_normal, const float* mp_max_dist,
             const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp, float th, int32_t* best_idx) {
  set_bounds(b);
  KeyFrame KF;
  fill_rig(KF, kf_k, kf_desc, kf_cam, kf_uright, n_kf, b);
  set_levels(KF, scale_factors, nlevels, log_scale_factor);
  KF.mvInvLevelSigma2.assign(nlevels, 1.f);
  if (inv_level_sigma2) KF.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
  set_camera(KF, cam, b);
  std::vector<MapPoint> held(n_kf), pts(n_mp);
  KF.mvpMapPoints.assign(n_kf, nullptr);
  for (int i = 0; i < n_kf; ++i)
    if (kf_held && kf_held[i]) { held[i].nObs = kf_held[i] == 2 ? 100 : 0; KF.mvpMapPoints[i] = &held[i]; }
  std::vector<MapPoint*> vp(n_mp);
  for (int i = 0; i < n_mp; ++i) {
    pts[i] = make_point(mp_xyz + 3 * i, mp_normal + 3 * i, mp_max_dist[i], mp_min_dist[i], mp_max_d[i], mp_desc + (size_t)i * 32, 3,
                        !mp_valid[i]);
    vp[i] = &pts[i];
  }
  for (int i = 0; i < 2 * n_mp; ++i) best_idx[i] = -1;
  MapPoint::trace().clear();
  ORBmatcher matcher(0.6f, true);
  int nf;
  std::vector<MapPoint*> vpReplace(n_mp, nullptr);
  if (!sim3) {
    KF.Tcw = fmat(pose, 4, 4);
    KF.Ow = fmat(Ow, 3, 1); KF.Ow_cam2 = fmat(Ow + 3, 3, 1);
    nf = matcher.Fuse(&KF, vp, fmat(calib, 4, 3), th);
  } else {
    std::vector<int> cams(n_mp, 0);
    nf = g_cam1 ? matcher.Fuse_cam1(&KF, fmat(pose, 4, 4), vp, th, vpReplace)
                : matcher.Fuse(&KF, fmat(pose, 4, 4), vp, cams, th, vpReplace, fmat(calib, 4, 3));
  }
  // read the fused (map point, feature) pairs back from the trace
  int cur = -1;
  for (const MapPoint::Trace& t : MapPoint::trace()) {
    if (t.kind == 0) cur = index_in(pts, static_cast<const MapPoint*>(t.p));
    else if (cur >= 0) best_idx[2 * cur + KF.keypoint_to_cam[(size_t)t.idx]] = (int32_t)t.idx;
  }
  return nf;
}

// ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:2814-3136); arguments as om_search_by_sim3.
int omr_search_by_sim3(const oo_keypoint* k1, const uint8_t* d1, const int32_t* cam1, int n1, const float* T1w,
                       const oo_keypoint* k2, const uint8_t* d2, const int32_t* cam2, int n2, const float* T2w, om_bounds b,
                       const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam, float s12,
                       const float* R12, const float* t12, const float* calib, const int32_t* mp1_valid, const float* mp1_xyz,
                       const float* mp1_max_dist, const float* mp1_min_dist, const float* mp1_max_d, const uint8_t* mp1_desc,
                       const int32_t* mp2_valid, const float* mp2_xyz, const float* mp2_max_dist, const float* mp2_min_dist,
                       const float* mp2_max_d, const uint8_t* mp2_desc, float th, int32_t* match12) {
  set_bounds(b);
  KeyFrame KF1, KF2;
  fill_rig(KF1, k1, d1, cam1, nullptr, n1, b);
  fill_rig(KF2, k2, d2, cam2, nullptr, n2, b);
  set_levels(KF1, scale_factors, nlevels, log_scale_factor);
  set_levels(KF2, scale_factors, nlevels, log_scale_factor);
  set_camera(KF1, cam, b); set_camera(KF2, cam, b);
  KF1.Tcw = fmat(T1w, 4, 4); KF2.Tcw = fmat(T2w, 4, 4);
  std::vector<MapPoint> p1(n1), p2(n2);
  KF1.mvpMapPoints.assign(n1, nullptr);
  KF2.mvpMapPoints.assign(n2, nullptr);
  for (int i = 0; i < n1; ++i)
    if (mp1_valid[i]) {
      p1[i] = make_point(mp1_xyz + 3 * i, nullptr, mp1_max_dist[i], mp1_min_dist[i], mp1_max_d[i], mp1_desc + (size_t)i * 32, 1, false);
      KF1.mvpMapPoints[i] = &p1[i];
    }
  for (int i = 0; i < n2; ++i)
    if (mp2_valid[i]) {
      p2[i] = make_point(mp2_xyz + 3 * i, nullptr, mp2_max_dist[i], mp2_min_dist[i], mp2_max_d[i], mp2_desc + (size_t)i * 32, 1, false);
      KF2.mvpMapPoints[i] = &p2[i];
    }
  std::vector<MapPoint*> vpMatches12(n1, nullptr);
  ORBmatcher matcher(0.75f, true);
  const int nf = g_cam1 ? matcher.SearchBySim3_cam1(&KF1, &KF2, vpMatches12, s12, fmat(R12, 3, 3), fmat(t12, 3, 1), th)
                        : matcher.SearchBySim3(&KF1, &KF2, vpMatches12, s12, fmat(R12, 3, 3), fmat(t12, 3, 1), th, fmat(calib, 4, 3));
  for (int i = 0; i < n1; ++i) match12[i] = index_in(p2, vpMatches12[i]);
  return nf;
}

// DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector, FeatureVector, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195) as Frame::ComputeBoW calls it, on a vocabulary read with the
// reference's own loadFromTextFile (:1339-1424, the ORBvoc.txt format).  Outputs as om_bow_transform's vectors.
int omr_bow_transform(const char* voc_path, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_value,
                      int32_t* n_bow, int32_t* fv_node, int32_t* fv_start, int32_t* fv_items, int32_t* n_fv) {
  typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
  ORBVocabulary voc;
  if (!voc.loadFromTextFile(voc_path) || voc.empty()) return -1;
  std::vector<cv::Mat> features(n);
  for (int i = 0; i < n; ++i) features[i] = desc_rows(desc + (size_t)i * 32, 1);
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  voc.transform(features, bv, fv, levelsup);
  int j = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++j) { bow_word[j] = (int32_t)it->first; bow_value[j] = it->second; }
  *n_bow = j;
  int a = 0, run = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++a) {
    fv_node[a] = (int32_t)it->first;
    fv_start[a] = run;
    for (size_t q = 0; q < it->second.size(); ++q) fv_items[run++] = (int32_t)it->second[q];
  }
  fv_start[a] = run;
  *n_fv = a;
  return (int)voc.size();
}

#endif  // OMR_HOT_ONLY

}  // extern "C"
