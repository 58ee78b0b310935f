// TEST INFRASTRUCTURE (compile-only) — the reference's own call sites of the hot path, shaped like
// src/Tracking.cc:870-871 (MonocularInitialization), :1267-1279 (TrackWithMotionModel), :1755-1764
// (SearchLocalPoints), :2099 (Relocalization), src/LoopClosing.cc:532-534 (the two-camera Sim3 search),
// src/MapPoint.cc:381 (static DescriptorDistance) and src/Frame.cc:397-403 (ExtractORB), compiled against the
// reference's REAL headers.  It proves that the members the drop-in translation units define are the members these
// call sites bind to: the snippet's undefined symbols must all be defined by ORBmatcher_b200.o / ORBextractor_b200.o
// (tests/test_dropin_build.py checks that with nm).
#include "ORBmatcher.h"
#include "ORBextractor.h"

using namespace ORB_SLAM2;

int monocular_initialization(Frame& mInitialFrame, Frame& mCurrentFrame, std::vector<cv::Point2f>& mvbPrevMatched,
                             std::vector<int>& mvIniMatches) {
  ORBmatcher matcher(0.9, true);
  return matcher.SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100);
}

int track_with_motion_model(Frame& mCurrentFrame, Frame& mLastFrame, cv::Mat mCaliMatrix, bool mono) {
  ORBmatcher matcher(0.9, true);
  int th = 15;
  int nmatches = matcher.SearchByProjection(mCurrentFrame, mLastFrame, th, mono, mCaliMatrix);
  if (nmatches < 20) {
    fill(mCurrentFrame.mvpMapPoints.begin(), mCurrentFrame.mvpMapPoints.end(), static_cast<MapPoint*>(NULL));
    nmatches = matcher.SearchByProjection(mCurrentFrame, mLastFrame, 2 * th, mono, mCaliMatrix);
  }
  return nmatches;
}

int search_local_points(Frame& mCurrentFrame, std::vector<MapPoint*>& mvpLocalMapPoints) {
  ORBmatcher matcher(0.8);
  int th = 3;
  return matcher.SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th);
}

int relocalization(Frame& mCurrentFrame, KeyFrame* pKF, std::set<MapPoint*>& sFound) {
  ORBmatcher matcher2(0.9, true);
  return matcher2.SearchByProjection(mCurrentFrame, pKF, sFound, 10, 100);
}

int compute_sim3(KeyFrame* mpCurrentKF, cv::Mat mScw, std::vector<MapPoint*>& mvpLoopMapPoints, std::vector<int>& vLoopMPCams,
                 std::vector<MapPoint*>& mvpCurrentMatchedPoints, cv::Mat mCalibMatrix) {
  ORBmatcher matcher(0.75, true);
  return matcher.SearchByProjection(mpCurrentKF, mScw, mvpLoopMapPoints, vLoopMPCams, mvpCurrentMatchedPoints, 10, mCalibMatrix);
}

int distinctive_descriptor_distance(const cv::Mat& a, const cv::Mat& b) { return ORBmatcher::DescriptorDistance(a, b); }

void extract_orb(ORBextractor* mpORBextractorLeft, const cv::Mat& im, std::vector<cv::KeyPoint>& mvKeys, cv::Mat& mDescriptors) {
  (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
}

ORBextractor* make_extractor(int nFeatures, float fScaleFactor, int nLevels, int fIniThFAST, int fMinThFAST) {
  return new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST);
}
