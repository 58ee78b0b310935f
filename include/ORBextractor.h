/* ORBextractor.h — drop-in C++ class for the reference's ORB_SLAM2::ORBextractor
 * (reference: include/ORBextractor.h:45-112, src/ORBextractor.cc) backed by the B200 C-ABI
 * library (orb_b200.h).  Same constructor, call operator, getters and public mvImagePyramid, so
 * Frame::ExtractORB / ExtractORB_cam2 (src/Frame.cc:397-419) and Tracking (src/Tracking.cc:144-145)
 * compile unchanged; link liborb_b200.so instead of compiling src/ORBextractor.cc.
 *
 * Needs the OpenCV core types the reference already uses (cv::Mat, cv::KeyPoint, cv::InputArray,
 * cv::OutputArray).  Header-only; no CUDA headers leak through it.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <stdexcept>
#include <string>
#include <vector>

#include <opencv/cv.h>

#include "orb_b200.h"

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
      : nfeatures_(nfeatures), scaleFactor_(scaleFactor), nlevels_(nlevels), iniThFAST_(iniThFAST),
        minThFAST_(minThFAST), h_(nullptr), width_(0), height_(0) {
    // scale tables do not depend on the image size: a 64x64 probe handle is enough to read them
    mvImagePyramid.resize(nlevels);
    computeTables();
  }
  ~ORBextractor() { orbx_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  /* Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference. */
  void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                  cv::OutputArray descriptors) {
    if (_image.empty()) return;  // src/ORBextractor.cc:1047-1048
    cv::Mat image = _image.getMat();
    ensure(image.cols, image.rows);
    const int cap = orbx_max_keypoints(h_);
    kps_.resize(cap);
    desc_.resize((size_t)cap * 32);
    int n = 0;
    check(orbx_extract(h_, image.data, image.rows, image.cols, (size_t)image.step, kps_.data(), desc_.data(), cap, &n));
    keypoints.clear();
    keypoints.reserve(n);
    if (n == 0) {
      descriptors.release();
    } else {
      descriptors.create(n, 32, CV_8U);
      cv::Mat d = descriptors.getMat();
      for (int i = 0; i < n; ++i) {
        const orbx_keypoint& k = kps_[i];
        keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, -1));
        std::copy(desc_.begin() + (size_t)i * 32, desc_.begin() + (size_t)(i + 1) * 32, d.ptr(i));
      }
    }
    // mvImagePyramid is a public member some callers read (upstream ComputeStereoMatches, src/Frame.cc:879-896)
    for (int l = 0; l < nlevels_; ++l) {
      int w = 0, hh = 0;
      check(orbx_get_pyramid_level(h_, 0, l, 0, nullptr, 0, &w, &hh));
      mvImagePyramid[l].create(hh, w, CV_8UC1);
      check(orbx_get_pyramid_level(h_, 0, l, 0, mvImagePyramid[l].data, (size_t)mvImagePyramid[l].step, &w, &hh));
    }
  }

  int inline GetLevels() { return nlevels_; }
  float inline GetScaleFactor() { return scaleFactor_; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<cv::Mat> mvImagePyramid;

 protected:
  void check(int rc) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string("orb_b200: ") + orbx_last_error(h_));
  }
  void ensure(int w, int h) {
    if (h_ && w == width_ && h == height_) return;
    orbx_destroy(h_);
    h_ = nullptr;
    orbx_config cfg = {nfeatures_, scaleFactor_, nlevels_, iniThFAST_, minThFAST_, w, h, 1, -1};
    if (orbx_create(&cfg, &h_) != ORBX_OK) throw std::runtime_error(std::string("orb_b200: ") + orbx_last_error(nullptr));
    width_ = w;
    height_ = h;
  }
  void computeTables() {
    // src/ORBextractor.cc:416-432, same float/double arithmetic
    mvScaleFactor.assign(nlevels_, 1.0f);
    mvLevelSigma2.assign(nlevels_, 1.0f);
    const double sf = scaleFactor_;
    for (int i = 1; i < nlevels_; i++) {
      mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * sf);
      mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels_);
    mvInvLevelSigma2.resize(nlevels_);
    for (int i = 0; i < nlevels_; i++) {
      mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
      mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
  }

  int nfeatures_;
  float scaleFactor_;
  int nlevels_, iniThFAST_, minThFAST_;
  orbx_extractor* h_;
  int width_, height_;
  std::vector<orbx_keypoint> kps_;
  std::vector<uint8_t> desc_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM2

#endif
