// ORBextractor_b200.cc — drop-in replacement of the reference's src/ORBextractor.cc.
//
// Compile THIS file instead of src/ORBextractor.cc and link liborb_b200.so: it defines the members of the
// reference's own class ORB_SLAM2::ORBextractor exactly as include/ORBextractor.h:45-112 declares them (that
// header is used as is), on top of the C ABI (include/orb_b200.h).  Frame::ExtractORB / ExtractORB_cam2
// (src/Frame.cc:397-419), the Frame constructors (src/Frame.cc:148-346) and Tracking (src/Tracking.cc:144-145)
// compile and link unchanged.
//
// The class has no spare member for a device handle, so the handle lives in a side table keyed by the extractor's
// address (created on the first image, whose size sizes the device workspace; re-created when the image size
// changes).  There is no CPU fallback: without a CUDA device operator() throws.
#include "ORBextractor.h"  // the reference's header

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "orb_b200.h"
#include "orb_b200_dropin.h"

namespace ORB_SLAM2 {
namespace {

const int EDGE_THRESHOLD = 19;  // src/ORBextractor.cc:72

struct Slot {
  orbx_extractor* h = nullptr;
  orbx_config cfg = {};
  bool mirror = false;
  std::vector<orbx_keypoint> kps;
  std::vector<uint8_t> desc;
};
std::mutex g_mu;
std::map<const ORBextractor*, Slot> g_slots;
int g_device = -1;

[[noreturn]] void fail(const orbx_extractor* h) { throw std::runtime_error(std::string("orb_b200: ") + orbx_last_error(h)); }

}  // namespace

namespace b200 {
void SetDevice(int device) { std::lock_guard<std::mutex> lock(g_mu); g_device = device; }
void SetPyramidMirror(ORBextractor* e, bool mirror) { std::lock_guard<std::mutex> lock(g_mu); g_slots[e].mirror = mirror; }
void Release(ORBextractor* e) {
  std::lock_guard<std::mutex> lock(g_mu);
  std::map<const ORBextractor*, Slot>::iterator it = g_slots.find(e);
  if (it == g_slots.end()) return;
  orbx_destroy(it->second.h);
  g_slots.erase(it);
}
}  // namespace b200

// ORBextractor::ORBextractor, src/ORBextractor.cc:411-471: the members other code reads through the getters
// (scale tables, :416-432) and mnFeaturesPerLevel (:436-447) with the reference's float / double arithmetic.
// `pattern` and `umax` feed the CPU descriptor / orientation code only; the device holds its own copies.
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  mvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels);
  mvScaleFactor[0] = 1.0f;
  mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor;  // float * double -> float
    mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
  }
  mvInvScaleFactor.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
    mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
  }
  mvImagePyramid.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);
  const float factor = 1.0f / scaleFactor;
  float nDesiredFeaturesPerScale = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sumFeatures = 0;
  for (int level = 0; level < nlevels - 1; level++) {
    mnFeaturesPerLevel[level] = (int)std::lrint(nDesiredFeaturesPerScale);  // cvRound
    sumFeatures += mnFeaturesPerLevel[level];
    nDesiredFeaturesPerScale *= factor;
  }
  mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);
}

// ORBextractor::operator(), src/ORBextractor.cc:1044-1107.  The mask is ignored, as in the reference.
void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
  (void)_mask;
  if (_image.empty()) return;  // :1047-1048
  cv::Mat image = _image.getMat();
  if (image.type() != CV_8UC1) throw std::invalid_argument("ORBextractor: image must be CV_8UC1");  // assert at :1051

  Slot* s;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    s = &g_slots[this];  // std::map nodes are stable: the pointer stays valid outside the lock
    const orbx_config want = {nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, image.cols, image.rows, 1, g_device};
    if (!s->h || std::memcmp(&s->cfg, &want, sizeof(want)) != 0) {
      orbx_destroy(s->h);
      s->h = nullptr;
      if (orbx_create(&want, &s->h) != ORBX_OK) fail(nullptr);
      s->cfg = want;
      const int cap = orbx_max_keypoints(s->h);
      s->kps.resize(cap);
      s->desc.resize((size_t)cap * 32);
    }
  }
  int n = 0;
  if (orbx_extract(s->h, image.data, image.rows, image.cols, (size_t)image.step, s->kps.data(), s->desc.data(),
                   (int)s->kps.size(), &n) != ORBX_OK)
    fail(s->h);

  _keypoints.clear();
  _keypoints.reserve(n);
  if (n == 0) {
    _descriptors.release();  // :1065-1066
  } else {
    _descriptors.create(n, 32, CV_8U);  // :1069
    cv::Mat descriptors = _descriptors.getMat();
    for (int i = 0; i < n; ++i) {
      const orbx_keypoint& k = s->kps[i];
      cv::KeyPoint kp;  // class_id keeps its default (-1), as with the reference's KeyPoint(pt, size, angle, response, octave)
      kp.pt.x = k.x; kp.pt.y = k.y; kp.size = k.size; kp.angle = k.angle; kp.response = k.response; kp.octave = k.octave;
      _keypoints.push_back(kp);
      std::memcpy(descriptors.ptr(i), &s->desc[(size_t)i * 32], 32);
    }
  }
  if (s->mirror) {
    // mvImagePyramid[l] = ROI of a bordered parent, as ComputePyramid leaves it (:1109-1134)
    for (int l = 0; l < nlevels; ++l) {
      int w = 0, h = 0;
      if (orbx_get_pyramid_level(s->h, 0, l, 1, nullptr, 0, &w, &h) != ORBX_OK) fail(s->h);
      cv::Mat parent(h, w, CV_8UC1);
      if (orbx_get_pyramid_level(s->h, 0, l, 1, parent.data, (size_t)parent.step, &w, &h) != ORBX_OK) fail(s->h);
      mvImagePyramid[l] = parent.rowRange(EDGE_THRESHOLD, h - EDGE_THRESHOLD).colRange(EDGE_THRESHOLD, w - EDGE_THRESHOLD);
    }
  }
}

}  // namespace ORB_SLAM2
