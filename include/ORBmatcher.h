/* ORBmatcher.h — drop-in C++ class for the hot members of the reference's ORB_SLAM2::ORBmatcher
 * (reference: include/ORBmatcher.h:37-137, src/ORBmatcher.cc) backed by the B200 C-ABI library.
 *
 * The reference's matcher walks Frame / MapPoint object graphs; the device kernels take flat
 * arrays.  This class flattens exactly the fields the reference functions read and scatters the
 * results back, so call sites keep their shape:
 *   DescriptorDistance(a, b)                                       src/ORBmatcher.cc:3994-4010
 *   SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, w) src/ORBmatcher.cc:868-983
 *   SearchByProjection(F, vpMapPoints, th)                          src/ORBmatcher.cc:62-149
 *   SearchByBoW_cam1(pKF, F, vpMapPointMatches)                     src/ORBmatcher.cc:390-565
 * It is a template over the caller's Frame / MapPoint types so it compiles against the reference's
 * own include/Frame.h and include/MapPoint.h without modification (members used are listed at each
 * function).  The other SearchByBoW / SearchByProjection variants are reachable through the flat
 * entry points of orb_b200.h; Fuse and the Sim3 searches stay with the reference implementation.
 */
#ifndef ORBMATCHER_B200_H
#define ORBMATCHER_B200_H

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv/cv.h>

#include "orb_b200.h"

namespace ORB_SLAM2 {

class ORBmatcherB200 {
 public:
  static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;

  ORBmatcherB200(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri), m_(nullptr) {
    if (orbm_create(-1, &m_) != ORBX_OK) throw std::runtime_error(std::string("orb_b200: ") + orbm_last_error(nullptr));
  }
  ~ORBmatcherB200() { orbm_destroy(m_); }
  ORBmatcherB200(const ORBmatcherB200&) = delete;
  ORBmatcherB200& operator=(const ORBmatcherB200&) = delete;

  /* static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b) */
  int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    int32_t out = 0;
    check(orbm_distance_pairs_host(m_, a.ptr(0), b.ptr(0), 1, &out));
    return out;
  }

  /* int SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched,
   *                             vector<int> &vnMatches12, int windowSize=10)
   * reads F.mvKeysUn, F.mDescriptors, F2.mnMinX/mnMaxX/mnMinY/mnMaxY (static in the reference). */
  template <class FrameT>
  int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched,
                              std::vector<int>& vnMatches12, int windowSize = 10) {
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    const int cap = std::max(std::max(n1, n2), 1);
    std::vector<orbx_keypoint> k1(cap), k2(cap);
    std::vector<uint8_t> d1((size_t)cap * 32), d2((size_t)cap * 32);
    flatten(F1.mvKeysUn, F1.mDescriptors, k1, d1);
    flatten(F2.mvKeysUn, F2.mDescriptors, k2, d2);
    std::vector<float> prev((size_t)cap * 2, 0.f);
    for (int i = 0; i < n1; ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    std::vector<int32_t> m12(cap, -1);
    int32_t nmatches = 0;
    const orbm_bounds b = {FrameT::mnMinX, FrameT::mnMaxX, FrameT::mnMinY, FrameT::mnMaxY};
    check(orbm_search_for_initialization_host(m_, 1, cap, k1.data(), d1.data(), &n1, k2.data(), d2.data(), &n2, b,
                                              prev.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(),
                                              &nmatches));
    vnMatches12.assign(m12.begin(), m12.begin() + n1);
    for (int i = 0; i < n1; ++i) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    return nmatches;
  }

  /* int SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th=3)
   * reads F.mvKeysUn, F.mDescriptors, F.mvuRight, F.mvScaleFactors, F.mvpMapPoints (+Observations()),
   * and per map point mbTrackInView, isBad(), mnTrackScaleLevel, mTrackViewCos, mTrackProjX/Y/XR,
   * GetDescriptor(), Observations(); writes F.mvpMapPoints. */
  template <class FrameT, class MapPointT>
  int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th = 3) {
    const int n = (int)F.mvKeysUn.size(), nmp = (int)vpMapPoints.size();
    if (n == 0 || nmp == 0) return 0;
    std::vector<orbx_keypoint> k(n);
    std::vector<uint8_t> d((size_t)n * 32);
    flatten(F.mvKeysUn, F.mDescriptors, k, d);
    std::vector<float> ur(F.mvuRight.begin(), F.mvuRight.begin() + n);
    std::vector<int32_t> fmp(n, -1), fobs(n, 0), mobs(nmp, 0);
    std::vector<MapPointT*> initial(n, nullptr);
    for (int i = 0; i < n; ++i)
      if (F.mvpMapPoints[i]) {
        initial[i] = F.mvpMapPoints[i];
        fmp[i] = nmp;  // any value >= 0 other than a new index: "holds a point from before this call"
        fobs[i] = F.mvpMapPoints[i]->Observations() > 0;
      }
    std::vector<orbm_mappoint> mp(nmp);
    std::vector<uint8_t> md((size_t)nmp * 32);
    for (int i = 0; i < nmp; ++i) {
      MapPointT* p = vpMapPoints[i];
      mp[i].track_in_view = p->mbTrackInView ? 1 : 0;
      mp[i].bad = p->isBad() ? 1 : 0;
      mp[i].level = p->mnTrackScaleLevel;
      mp[i].view_cos = p->mTrackViewCos;
      mp[i].proj_x = p->mTrackProjX;
      mp[i].proj_y = p->mTrackProjY;
      mp[i].proj_xr = p->mTrackProjXR;
      mobs[i] = p->Observations() > 0;
      const cv::Mat desc = p->GetDescriptor();
      std::copy(desc.ptr(0), desc.ptr(0) + 32, md.begin() + (size_t)i * 32);
    }
    int nmatches = 0;
    const orbm_bounds b = {FrameT::mnMinX, FrameT::mnMaxX, FrameT::mnMinY, FrameT::mnMaxY};
    check(orbm_search_by_projection_points_host(m_, k.data(), d.data(), ur.data(), n, b, F.mvScaleFactors.data(),
                                                (int)F.mvScaleFactors.size(), mp.data(), md.data(), mobs.data(), nmp, th,
                                                mfNNratio, fmp.data(), fobs.data(), &nmatches));
    for (int i = 0; i < n; ++i)
      if (fmp[i] >= 0 && fmp[i] < nmp) F.mvpMapPoints[i] = vpMapPoints[fmp[i]];
    return nmatches;
  }

  /* int SearchByBoW_cam1(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)   (:390-565)
   * reads pKF->GetMapPointMatches_cam1(), mFeatVec_cam1, mDescriptors, mvKeysUn, N;
   *       F.mFeatVec_cam1, F.mDescriptors, F.mvKeys, F.N; isBad() of the key frame's map points.
   * FeatureVector = std::map<node id, std::vector<unsigned int>> (DBoW2), walked in map order. */
  template <class KeyFrameT, class FrameT, class MapPointT>
  int SearchByBoW_cam1(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches) {
    const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches_cam1();
    const int n1 = (int)pKF->N, n2 = (int)F.N;
    vpMapPointMatches.assign(n2, static_cast<MapPointT*>(nullptr));
    std::vector<int32_t> valid1(std::max(n1, 1), 0);
    for (int i = 0; i < n1; ++i) valid1[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
    std::vector<float> a1(std::max(n1, 1)), a2(std::max(n2, 1));
    std::vector<uint8_t> d1((size_t)std::max(n1, 1) * 32), d2((size_t)std::max(n2, 1) * 32);
    for (int i = 0; i < n1; ++i) {
      a1[i] = pKF->mvKeysUn[i].angle;
      std::copy(pKF->mDescriptors.ptr(i), pKF->mDescriptors.ptr(i) + 32, d1.begin() + (size_t)i * 32);
    }
    for (int i = 0; i < n2; ++i) {
      a2[i] = F.mvKeys[i].angle;
      std::copy(F.mDescriptors.ptr(i), F.mDescriptors.ptr(i) + 32, d2.begin() + (size_t)i * 32);
    }
    FlatFeatVec f1, f2;
    f1.fill(pKF->mFeatVec_cam1, n1);  // indices >= N (camera 2) are dropped, like `realIdxKF >= pKF->N` (:432)
    f2.fill(F.mFeatVec_cam1, n2);
    std::vector<int32_t> m12(std::max(n1, 1)), m21(std::max(n2, 1));
    int nmatches = 0;
    check(orbm_search_by_bow_host(m_, d1.data(), a1.data(), valid1.data(), n1, f1.view(), d2.data(), a2.data(), nullptr, n2,
                                  f2.view(), mfNNratio, mbCheckOrientation ? 1 : 0, TH_LOW, m12.data(), m21.data(),
                                  &nmatches));
    for (int i = 0; i < n2; ++i)
      if (m21[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[m21[i]];
    return nmatches;
  }

 protected:
  struct FlatFeatVec {  // DBoW2::FeatureVector -> CSR
    std::vector<int32_t> node, start, items;
    template <class MapT>
    void fill(const MapT& fv, int n_limit) {
      start.push_back(0);
      for (typename MapT::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        node.push_back((int32_t)it->first);
        for (size_t j = 0; j < it->second.size(); ++j)
          if ((int)it->second[j] < n_limit) items.push_back((int32_t)it->second[j]);
        start.push_back((int32_t)items.size());
      }
    }
    orbm_featvec view() const {
      orbm_featvec v = {node.data(), start.data(), items.data(), (int32_t)node.size()};
      return v;
    }
  };

  void check(int rc) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string("orb_b200: ") + orbm_last_error(m_));
  }
  static void flatten(const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc, std::vector<orbx_keypoint>& k,
                      std::vector<uint8_t>& d) {
    for (size_t i = 0; i < keys.size(); ++i) {
      const cv::KeyPoint& kp = keys[i];
      k[i].x = kp.pt.x; k[i].y = kp.pt.y; k[i].size = kp.size; k[i].angle = kp.angle;
      k[i].response = kp.response; k[i].octave = kp.octave;
      std::copy(desc.ptr((int)i), desc.ptr((int)i) + 32, d.begin() + i * 32);
    }
  }

  float mfNNratio;
  bool mbCheckOrientation;
  orbm_matcher* m_;
};

}  // namespace ORB_SLAM2

#endif
