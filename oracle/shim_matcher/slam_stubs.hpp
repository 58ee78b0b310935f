// TEST INFRASTRUCTURE — stand-ins for ORB_SLAM2::Frame / KeyFrame / MapPoint so that the reference's
// src/ORBmatcher.cc compiles VERBATIM without the rest of the system (oracle/Makefile defines the three
// header guards FRAME_H / KEYFRAME_H / MAPPOINT_H and force-includes this file).  Written from scratch:
// only the members ORBmatcher.cc touches, with the field names and types of include/Frame.h,
// include/KeyFrame.h, include/MapPoint.h.  The grid queries restate src/Frame.cc:510-630 and
// src/KeyFrame.cc:888-977 (per-camera grids of global feature indices).
#pragma once
#include <cassert>
#include <cmath>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>
#include <vector>

#include "cvm.hpp"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;  // the reference's headers do (include/Frame.h, include/KeyFrame.h)

// The reference's progress prints (`cout << ... << nFound << endl`) are swallowed: the toolchain here links
// libstdc++ statically into the .so, whose iostream number formatting crashes when the process also holds the
// shared libstdc++ (GNU-unique locale ids), and the prints are noise in a test log anyway.  A macro, so that the
// reference source itself stays untouched.
namespace orb_shim {
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
inline NullStream& null_stream() { static NullStream s; return s; }
}  // namespace orb_shim
#define cout orb_shim::null_stream()

namespace ORB_SLAM2 {
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class MapPoint;
class KeyFrame;

struct GridSet {  // mGrids[cam][ix][iy] = global feature indices in insertion order
  std::vector<std::vector<std::vector<std::vector<size_t>>>> cells;
  float min_x = 0, min_y = 0, inv_w = 1, inv_h = 1;
  void build(const std::vector<cv::KeyPoint>& keys, const std::unordered_map<size_t, int>& cam_of, int n_cams, float mnMinX,
             float mnMaxX, float mnMinY, float mnMaxY) {
    min_x = mnMinX; min_y = mnMinY;
    inv_w = static_cast<float>(FRAME_GRID_COLS) / (mnMaxX - mnMinX);
    inv_h = static_cast<float>(FRAME_GRID_ROWS) / (mnMaxY - mnMinY);
    cells.assign(n_cams, std::vector<std::vector<std::vector<size_t>>>(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS)));
    for (size_t i = 0; i < keys.size(); ++i) {
      const int px = (int)round((keys[i].pt.x - min_x) * inv_w), py = (int)round((keys[i].pt.y - min_y) * inv_h);
      if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
      auto it = cam_of.find(i);
      cells[it == cam_of.end() ? 0 : it->second][px][py].push_back(i);
    }
  }
  template <class MinT>
  std::vector<size_t> query(const std::vector<cv::KeyPoint>& keys, int cam, float x, float y, float r, MinT minx, MinT miny,
                            int minLevel, int maxLevel, bool use_levels) const {
    std::vector<size_t> out;
    const int c0 = std::max(0, (int)floor((x - minx - r) * inv_w));
    if (c0 >= FRAME_GRID_COLS) return out;
    const int c1 = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - minx + r) * inv_w));
    if (c1 < 0) return out;
    const int r0 = std::max(0, (int)floor((y - miny - r) * inv_h));
    if (r0 >= FRAME_GRID_ROWS) return out;
    const int r1 = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - miny + r) * inv_h));
    if (r1 < 0) return out;
    const bool bCheckLevels = use_levels && ((minLevel > 0) || (maxLevel >= 0));
    for (int ix = c0; ix <= c1; ix++)
      for (int iy = r0; iy <= r1; iy++)
        for (size_t idx : cells[cam][ix][iy]) {
          const cv::KeyPoint& kp = keys[idx];
          if (bCheckLevels) {
            if (kp.octave < minLevel) continue;
            if (maxLevel >= 0 && kp.octave > maxLevel) continue;
          }
          if (fabs(kp.pt.x - x) < r && fabs(kp.pt.y - y) < r) out.push_back(idx);
        }
    return out;
  }
};

#ifndef ORB_REAL_FRAME  // the frame build uses the reference's own include/Frame.h
class Frame {
 public:
  int N = 0, N_total = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeys_total, mvKeysUn_total;
  std::vector<float> mvuRight, mvuRight_total, mvDepth;
  cv::Mat mDescriptors;
  std::vector<cv::Mat> mDescriptors_total;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  int mnScaleLevels = 8;
  float mfScaleFactor = 1.2f, mfLogScaleFactor = 0.f;
  cv::Mat mTcw;
  float mb = 0, mbf = 0, fx = 0, fy = 0, cx = 0, cy = 0;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec, mFeatVec_cam1;
  std::unordered_map<size_t, int> keypoint_to_cam, cont_idx_to_local_cam_idx;
  GridSet grid1, grids;  // camera-1 grid over mvKeysUn (mGrid), per-camera grids over mvKeysUn_total (mGrids)
  vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                   const int maxLevel = -1) const {
    return grid1.query(mvKeysUn, 0, x, y, r, mnMinX, mnMinY, minLevel, maxLevel, true);
  }
  vector<size_t> GetFeaturesInArea(const int cam, const float& x, const float& y, const float& r, const int minLevel = -1,
                                   const int maxLevel = -1) const {
    return grids.query(mvKeysUn_total, cam, x, y, r, mnMinX, mnMinY, minLevel, maxLevel, true);
  }
};

#else
class Frame;
#endif

class MapPoint {
 public:
  // the fields Frame::isInFrustum writes
  bool mbTrackInView = false;
  int mnTrackScaleLevel = 0;
  float mTrackViewCos = 0, mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
  // state behind the accessors
  bool bad = false;
  int nObs = 0;
  cv::Mat descriptor, worldPos, normal;
  float mfMinDistance = 0, mfMaxDistance = 0;
  float minInvariance = 0, maxInvariance = 0;  // what Get{Min,Max}DistanceInvariance return (0.8 / 1.2 x the above)
  std::map<KeyFrame*, size_t> observations;
  MapPoint* replaced = nullptr;
  bool isBad() { return bad; }
  cv::Mat GetDescriptor() { return descriptor.clone(); }
  cv::Mat GetWorldPos() { trace().push_back(Trace{0, this, -1}); return worldPos.clone(); }
  cv::Mat GetNormal() { return normal.clone(); }
  int Observations() { return nObs; }
  float GetMinDistanceInvariance() { return minInvariance; }
  float GetMaxDistanceInvariance() { return maxInvariance; }
  bool IsInKeyFrame(KeyFrame* pKF) { return observations.count(pKF) != 0; }
  int GetIndexInKeyFrame(KeyFrame* pKF) { return observations.count(pKF) ? (int)observations[pKF] : -1; }
  int GetIndexInKeyFrame_cam1(KeyFrame* pKF) { return GetIndexInKeyFrame(pKF); }
  // Fuse's effects on the map are not applied (AddObservation / Replace do nothing); instead a trace records which
  // candidate is being processed (GetWorldPos opens every candidate's iteration, :2015 / :2247) and which key-frame
  // features it asks for (GetMapPoint, :2164 / :2425), so a test can read back the fused (point, feature) pairs
  struct Trace { int kind; const void* p; long idx; };  // kind 0: candidate p; kind 1: feature idx
  static std::vector<Trace>& trace() { static std::vector<Trace> t; return t; }
  void AddObservation(KeyFrame*, size_t) {}
  void Replace(MapPoint*) {}
  int PredictScale(const float& currentDist, KeyFrame* pKF);  // src/MapPoint.cc:584-600
  int PredictScale(const float& currentDist, Frame* pF);     // src/MapPoint.cc:602-617 (defined once Frame is complete)
};

class KeyFrame {
 public:
  int N = 0, N_total = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeys_total, mvKeysUn_total;
  std::vector<float> mvuRight, mvuRight_total;
  cv::Mat mDescriptors;
  std::vector<cv::Mat> mDescriptors_total;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  int mnScaleLevels = 8;
  float mfScaleFactor = 1.2f, mfLogScaleFactor = 0.f;
  cv::Mat mK;
  float mbf = 0, mb = 0, fx = 0, fy = 0, cx = 0, cy = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;  // ints in the reference (include/KeyFrame.h:234-237)
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec, mFeatVec_cam1;
  std::unordered_map<size_t, int> keypoint_to_cam, cont_idx_to_local_cam_idx;
  std::vector<MapPoint*> mvpMapPoints;  // size N_total
  cv::Mat Tcw, Tcw_cam2, Ow, Ow_cam2;
  GridSet grid1, grids;
  vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const {
    return grid1.query(mvKeysUn_total, 0, x, y, r, mnMinX, mnMinY, -1, -1, false);
  }
  vector<size_t> GetFeaturesInArea(const int& cam, const float& x, const float& y, const float& r) const {
    return grids.query(mvKeysUn_total, cam, x, y, r, mnMinX, mnMinY, -1, -1, false);
  }
  bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
  vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  vector<MapPoint*> GetMapPointMatches_cam1() { return vector<MapPoint*>(mvpMapPoints.begin(), mvpMapPoints.begin() + N); }
  MapPoint* GetMapPoint(const size_t& idx) { MapPoint::trace().push_back(MapPoint::Trace{1, nullptr, (long)idx}); return mvpMapPoints[idx]; }
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints)
      if (p && !p->isBad()) s.insert(p);
    return s;
  }
  void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
  cv::Mat GetDescriptor(const int& cam, const size_t& idx) const { return mDescriptors_total[cam].row((int)idx); }
  cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
  cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
  cv::Mat GetRotation_cam2() { return Tcw_cam2.rowRange(0, 3).colRange(0, 3).clone(); }
  cv::Mat GetTranslation_cam2() { return Tcw_cam2.rowRange(0, 3).col(3).clone(); }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  cv::Mat GetCameraCenter_cam2() { return Ow_cam2.clone(); }
};

inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
  const float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(std::log(ratio) / pKF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
  return nScale;
}
#ifndef ORB_REAL_FRAME
inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
  const float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(std::log(ratio) / pF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
  return nScale;
}
#endif
}  // namespace ORB_SLAM2
