// TEST INFRASTRUCTURE — what the reference's src/Frame.cc needs around it to compile VERBATIM with its real
// include/Frame.h (oracle/Makefile: _ref/libframe_ref.so pre-defines the guards of MapPoint.h, KeyFrame.h,
// ORBextractor.h, Converter.h and force-includes this file).  Written from scratch:
//   * ORBextractor: returns the keypoints / descriptors a test queued for it (the extraction itself is pinned
//     elsewhere, oracle/_ref/liborb_ref.so), with the real class's scale-table getters;
//   * MapPoint / KeyFrame stand-ins (shared with the matcher build, oracle/shim_matcher/slam_stubs.hpp);
//   * Converter::toDescriptorVector (src/Converter.cc:27-45: the rows of the descriptor matrices);
//   * cv::undistortPoints -> the cv2-pinned restatement in oracle/cvprim.cc.
#pragma once
#define ORB_REAL_FRAME 1
#include "../shim_matcher/slam_stubs.hpp"
#include "../cvprim.h"

namespace cv {
#ifndef ORB_REAL_EXTRACTOR  // tests/native/shim_real provides InputArray / OutputArray for the reference's real ORBextractor.h
typedef const Mat& InputArray;
#endif
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat&, const Mat&) {
  // Frame.cc calls it in place on an N x 1 two-channel view of an N x 2 CV_32F matrix with R = Mat(), P = K
  const int n = src.rows;
  std::vector<float> in(2 * (size_t)n), out(2 * (size_t)n);
  for (int i = 0; i < n; ++i) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
  float d5[5] = {0, 0, 0, 0, 0};
  const int nd = dist.rows * dist.cols;
  for (int i = 0; i < nd && i < 5; ++i) d5[i] = dist.at<float>(i);
  cvp::undistort_points(in.data(), n, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d5, out.data());
  for (int i = 0; i < n; ++i) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
}
}  // namespace cv

namespace ORB_SLAM2 {
#ifndef ORB_REAL_EXTRACTOR  // tests/native/Makefile compiles Frame.cc against the reference's real include/ORBextractor.h
class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
  int nlevels = 8;
  float scaleFactor = 1.2f;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<cv::KeyPoint> next_keys;  // what the next operator() call returns
  cv::Mat next_desc;
  std::vector<cv::Mat> mvImagePyramid;
  ORBextractor(int, float sf, int nl, int, int) : nlevels(nl), scaleFactor(sf) {
    mvScaleFactor.resize(nl); mvLevelSigma2.resize(nl); mvInvScaleFactor.resize(nl); mvInvLevelSigma2.resize(nl);
    mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {  // src/ORBextractor.cc:417-433
      mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor;
      mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    for (int i = 0; i < nl; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
  }
  void operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) {
    keypoints = next_keys;
    descriptors = next_desc.clone();
  }
  int GetLevels() { return nlevels; }
  float GetScaleFactor() { return scaleFactor; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
};
#endif

class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> v;
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
  }
  static std::vector<cv::Mat> toDescriptorVector(const std::vector<cv::Mat>& Descriptors) {
    std::vector<cv::Mat> v;
    for (size_t c = 0; c < Descriptors.size(); ++c)
      for (int j = 0; j < Descriptors[c].rows; j++) v.push_back(Descriptors[c].row(j));
    return v;
  }
};
}  // namespace ORB_SLAM2
