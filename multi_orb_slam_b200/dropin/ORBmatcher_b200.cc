// ORBmatcher_b200.cc — drop-in definitions of the hot members of the reference's ORB_SLAM2::ORBmatcher.
//
// It is compiled against the reference's OWN include/ORBmatcher.h (and Frame.h / KeyFrame.h / MapPoint.h), which
// stay untouched, and defines, over the C ABI (include/orb_b200.h):
//
//   static int DescriptorDistance(const cv::Mat&, const cv::Mat&)                      src/ORBmatcher.cc:3994-4010
//   int SearchForInitialization(Frame&, Frame&, vbPrevMatched, vnMatches12, window)    :868-983
//   int SearchByProjection(Frame&, const vector<MapPoint*>&, th)                       :62-157
//   int SearchByProjection(Frame& Cur, const Frame& Last, th, bMono, CalibMatrix)      :3448-3641
//   int SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist)          :3809-3937
//   int SearchByProjection(KeyFrame*, Scw, vpPoints, vLoopMPCams, vpMatched, th, Calib) :566-752
//
// Two ways to build the reference with it:
//   (a) replace src/ORBmatcher.cc by this file + ORBmatcher_rest.cc-style CPU members you keep, or
//   (b) keep src/ORBmatcher.cc for the members not listed above and compile it with
//       -DDescriptorDistance=DescriptorDistance_cpu -DSearchForInitialization=SearchForInitialization_cpu
//       -DSearchByProjection=SearchByProjection_cpu, so that its definitions of the hot members move out of the way
//       (its own internal calls keep using the CPU versions) and every other translation unit links to these.
// With -DORB_B200_DROPIN_ALL_MEMBERS this file also defines the constructor and the three class constants, for
// builds that do not compile src/ORBmatcher.cc at all.
//
// Each function flattens exactly the fields the reference function reads (listed at each one) and scatters the
// result back into the caller's objects.  No CPU fallback: without a CUDA device the first call throws.
#include "ORBmatcher.h"  // the reference's header

#include <cstring>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "orb_b200.h"

namespace ORB_SLAM2 {
namespace {

std::mutex g_mu;
orbm_matcher* g_matcher = nullptr;

// One matcher handle per process, created on first use (the reference constructs ORBmatcher objects on the stack
// at every call site — src/Tracking.cc:870, 1238, 1267, 1756 — far too often to create device state each time).
// Calls are serialised: the handle's scratch buffers are not re-entrant.
struct Session {
  std::unique_lock<std::mutex> lock;
  orbm_matcher* m;
  Session() : lock(g_mu) {
    if (!g_matcher && orbm_create(-1, &g_matcher) != ORBX_OK)
      throw std::runtime_error(std::string("orb_b200: ") + orbm_last_error(nullptr));
    m = g_matcher;
  }
  void check(int rc) const {
    if (rc != ORBX_OK) throw std::runtime_error(std::string("orb_b200: ") + orbm_last_error(m));
  }
};

inline orbx_keypoint flat(const cv::KeyPoint& kp) {
  orbx_keypoint k;
  k.x = kp.pt.x; k.y = kp.pt.y; k.size = kp.size; k.angle = kp.angle; k.response = kp.response; k.octave = kp.octave;
  return k;
}

void flatten_keys(const std::vector<cv::KeyPoint>& keys, std::vector<orbx_keypoint>& out) {
  out.resize(std::max<size_t>(keys.size(), 1));
  for (size_t i = 0; i < keys.size(); ++i) out[i] = flat(keys[i]);
}

void flatten_rows(const cv::Mat& desc, int n, std::vector<uint8_t>& out) {
  out.assign((size_t)std::max(n, 1) * 32, 0);
  for (int i = 0; i < n; ++i) std::memcpy(&out[(size_t)i * 32], desc.ptr(i), 32);
}

// Multi-camera containers of Frame / KeyFrame: descriptor of global feature i =
// mDescriptors_total[keypoint_to_cam[i]].row(cont_idx_to_local_cam_idx[i]) (src/ORBmatcher.cc:3543-3545, :686-688).
template <class F>
void flatten_rig(const F& f, int n, std::vector<uint8_t>& desc, std::vector<int32_t>& cam) {
  desc.assign((size_t)std::max(n, 1) * 32, 0);
  cam.assign(std::max(n, 1), 0);
  for (int i = 0; i < n; ++i) {
    const int c = f.keypoint_to_cam.find((size_t)i)->second;
    const int local = f.cont_idx_to_local_cam_idx.find((size_t)i)->second;
    cam[i] = c;
    std::memcpy(&desc[(size_t)i * 32], f.mDescriptors_total[c].ptr(local), 32);
  }
}

void flatten_pose(const cv::Mat& T, float out[16]) {  // 4x4 CV_32F, row-major
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[4 * r + c] = T.at<float>(r, c);
}

void flatten_calib(const cv::Mat& C, float out[12]) {  // 4x3: rows 0-2 R_cam12, row 3 t_cam12
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 3; ++c) out[3 * r + c] = C.at<float>(r, c);
}

// MapPoint::mfMaxDistance is protected in the reference (include/MapPoint.h:153) and only reachable through
// PredictScale; the device needs the value itself.  A pointer to member formed in a derived class is the
// standard-conforming way to read a protected member of another object.
struct MapPointPeek : MapPoint {
  static float max_distance(MapPoint* p) { return p->*(&MapPointPeek::mfMaxDistance); }
};

orbm_bounds frame_bounds() {
  const orbm_bounds b = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
  return b;
}

}  // namespace

#ifdef ORB_B200_DROPIN_ALL_MEMBERS
const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
#endif

// ---- src/ORBmatcher.cc:3994-4010 ----------------------------------------------------------------------------
// One pair per call, as the reference's signature has it.  N x N users (MapPoint::ComputeDistinctiveDescriptors,
// src/MapPoint.cc:381) should call orbm_compute_distinctive_descriptors_host / orbm_distance_pairs_host once.
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
  Session s;
  int32_t out = 0;
  s.check(orbm_distance_pairs_host(s.m, a.ptr(0), b.ptr(0), 1, &out));
  return out;
}

// ---- src/ORBmatcher.cc:868-983 ------------------------------------------------------------------------------
// reads F1.mvKeysUn, F1.mDescriptors, F2.mvKeysUn, F2.mDescriptors, F2.GetFeaturesInArea (grid of F2.mvKeysUn over
// the static bounds); updates vbPrevMatched, fills vnMatches12.
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12,
                                        int windowSize) {
  const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
  vnMatches12 = vector<int>(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  const int cap = std::max(n1, n2);
  std::vector<orbx_keypoint> k1(cap), k2(cap);
  std::vector<uint8_t> d1((size_t)cap * 32), d2((size_t)cap * 32);
  for (int i = 0; i < n1; ++i) { k1[i] = flat(F1.mvKeysUn[i]); std::memcpy(&d1[(size_t)i * 32], F1.mDescriptors.ptr(i), 32); }
  for (int i = 0; i < n2; ++i) { k2[i] = flat(F2.mvKeysUn[i]); std::memcpy(&d2[(size_t)i * 32], F2.mDescriptors.ptr(i), 32); }
  std::vector<float> prev((size_t)cap * 2, 0.f);
  for (int i = 0; i < n1; ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
  std::vector<int32_t> m12(cap, -1);
  int32_t nmatches = 0;
  Session s;
  s.check(orbm_search_for_initialization_host(s.m, 1, cap, k1.data(), d1.data(), &n1, k2.data(), d2.data(), &n2, frame_bounds(),
                                              prev.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(),
                                              &nmatches));
  for (int i = 0; i < n1; ++i) {
    vnMatches12[i] = m12[i];
    vbPrevMatched[i].x = prev[2 * i];
    vbPrevMatched[i].y = prev[2 * i + 1];
  }
  return nmatches;
}

// ---- src/ORBmatcher.cc:62-157 -------------------------------------------------------------------------------
// reads F.mvKeysUn, F.mDescriptors, F.mvuRight, F.mvScaleFactors, F.mvpMapPoints (+Observations()), and per map
// point mbTrackInView, isBad(), mnTrackScaleLevel, mTrackViewCos, mTrackProjX/Y/XR, GetDescriptor(),
// Observations(); writes F.mvpMapPoints.
int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th) {
  const int n = (int)F.mvKeysUn.size(), nmp = (int)vpMapPoints.size();
  if (n == 0 || nmp == 0) return 0;
  std::vector<orbx_keypoint> k;
  std::vector<uint8_t> d, md((size_t)nmp * 32);
  flatten_keys(F.mvKeysUn, k);
  flatten_rows(F.mDescriptors, n, d);
  std::vector<float> ur(n, -1.f);
  for (int i = 0; i < n && i < (int)F.mvuRight.size(); ++i) ur[i] = F.mvuRight[i];
  std::vector<int32_t> fmp(n, -1), fobs(n, 0), mobs(nmp, 0);
  for (int i = 0; i < n; ++i)
    if (F.mvpMapPoints[i]) {
      fmp[i] = nmp;  // "holds a point from before this call": any value >= 0 that is not an index of vpMapPoints
      fobs[i] = F.mvpMapPoints[i]->Observations() > 0;
    }
  std::vector<orbm_mappoint> mp(nmp);
  for (int i = 0; i < nmp; ++i) {
    MapPoint* p = vpMapPoints[i];
    mp[i].track_in_view = p->mbTrackInView ? 1 : 0;
    mp[i].bad = p->isBad() ? 1 : 0;
    mp[i].level = p->mnTrackScaleLevel;
    mp[i].view_cos = p->mTrackViewCos;
    mp[i].proj_x = p->mTrackProjX;
    mp[i].proj_y = p->mTrackProjY;
    mp[i].proj_xr = p->mTrackProjXR;
    mobs[i] = p->Observations() > 0;
    const cv::Mat desc = p->GetDescriptor();
    std::memcpy(&md[(size_t)i * 32], desc.ptr(0), 32);
  }
  int nmatches = 0;
  Session s;
  s.check(orbm_search_by_projection_points_host(s.m, k.data(), d.data(), ur.data(), n, frame_bounds(), F.mvScaleFactors.data(),
                                                (int)F.mvScaleFactors.size(), mp.data(), md.data(), mobs.data(), nmp, th,
                                                mfNNratio, fmp.data(), fobs.data(), &nmatches));
  for (int i = 0; i < n; ++i)
    if (fmp[i] >= 0 && fmp[i] < nmp) F.mvpMapPoints[i] = vpMapPoints[fmp[i]];
  return nmatches;
}

// ---- src/ORBmatcher.cc:3448-3641 ----------------------------------------------------------------------------
// The per-frame tracking matcher (Tracking::TrackWithMotionModel, src/Tracking.cc:1267).
// reads Cur: mTcw, fx, fy, cx, cy, mb, mbf, mvScaleFactors, mvKeysUn_total, mvuRight_total, keypoint_to_cam,
//            cont_idx_to_local_cam_idx, mDescriptors_total, mvpMapPoints (+Observations()), per-camera grids;
//       Last: mTcw, N_total, mvpMapPoints (GetWorldPos, GetDescriptor), mvbOutlier, keypoint_to_cam,
//             mvKeys_total[i].octave, mvKeysUn_total[i].angle;  writes Cur.mvpMapPoints.
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono,
                                   cv::Mat CalibMatrix) {
  const int n_cur = (int)CurrentFrame.mvKeysUn_total.size(), n_last = LastFrame.N_total;
  if (n_cur == 0 || n_last == 0) return 0;
  std::vector<orbx_keypoint> ck, lk(n_last);
  std::vector<uint8_t> cd, ld((size_t)n_last * 32, 0);
  std::vector<int32_t> ccam, lcam(n_last, 0), lvalid(n_last, 0), lobs(n_last, 0);
  std::vector<float> lxyz((size_t)n_last * 3, 0.f);
  flatten_keys(CurrentFrame.mvKeysUn_total, ck);
  flatten_rig(CurrentFrame, n_cur, cd, ccam);
  for (int i = 0; i < n_last; ++i) {
    lk[i] = flat(LastFrame.mvKeysUn_total[i]);            // angle (:3561)
    lk[i].octave = LastFrame.mvKeys_total[i].octave;      // nLastOctave (:3512)
    lcam[i] = LastFrame.keypoint_to_cam.find((size_t)i)->second;
    MapPoint* p = LastFrame.mvpMapPoints[i];
    if (!p || LastFrame.mvbOutlier[i]) continue;
    lvalid[i] = 1;
    const cv::Mat x = p->GetWorldPos();
    for (int c = 0; c < 3; ++c) lxyz[3 * (size_t)i + c] = x.at<float>(c);
    const cv::Mat desc = p->GetDescriptor();
    std::memcpy(&ld[(size_t)i * 32], desc.ptr(0), 32);
    lobs[i] = p->Observations() > 0;
  }
  std::vector<float> ur(n_cur, -1.f);
  for (int i = 0; i < n_cur && i < (int)CurrentFrame.mvuRight_total.size(); ++i) ur[i] = CurrentFrame.mvuRight_total[i];
  std::vector<int32_t> cmp(n_cur, -1), cobs(n_cur, 0);
  for (int i = 0; i < n_cur; ++i)
    if (CurrentFrame.mvpMapPoints[i]) {
      cmp[i] = n_last;  // held from before: not an index into the last frame
      cobs[i] = CurrentFrame.mvpMapPoints[i]->Observations() > 0;
    }
  float Tc[16], Tl[16], calib[12];
  flatten_pose(CurrentFrame.mTcw, Tc);
  flatten_pose(LastFrame.mTcw, Tl);
  flatten_calib(CalibMatrix, calib);
  const orbm_camera cam = {CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy, CurrentFrame.mb, CurrentFrame.mbf};
  int nmatches = 0;
  Session s;
  s.check(orbm_search_by_projection_frame_host(s.m, ck.data(), cd.data(), ur.data(), ccam.data(), n_cur, frame_bounds(),
                                               CurrentFrame.mvScaleFactors.data(), (int)CurrentFrame.mvScaleFactors.size(), cam,
                                               Tc, Tl, lk.data(), lcam.data(), lvalid.data(), lxyz.data(), ld.data(), lobs.data(),
                                               n_last, calib, th, bMono ? 1 : 0, mbCheckOrientation ? 1 : 0, cmp.data(),
                                               cobs.data(), &nmatches));
  for (int i = 0; i < n_cur; ++i) {
    if (cmp[i] >= 0 && cmp[i] < n_last) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[cmp[i]];
    else if (cmp[i] < 0) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);  // cleared by the rotation check (:3631)
  }
  return nmatches;
}

// ---- src/ORBmatcher.cc:3809-3937 ----------------------------------------------------------------------------
// Relocalisation refinement (src/Tracking.cc:2099, 2116), camera 1 only.
// reads Cur: mTcw, fx, fy, cx, cy, mvScaleFactors, mfLogScaleFactor, mvKeysUn, mDescriptors, mvpMapPoints, grid;
//       pKF: GetMapPointMatches_cam1(), mvKeysUn[i].angle; per point isBad(), GetWorldPos, GetDescriptor,
//       Get{Max,Min}DistanceInvariance, PredictScale (mfMaxDistance);  writes Cur.mvpMapPoints.
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist) {
  const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches_cam1();
  const int n_cur = (int)CurrentFrame.mvKeysUn.size(), n_kf = (int)vpMPs.size();
  if (n_cur == 0 || n_kf == 0) return 0;
  std::vector<orbx_keypoint> ck;
  std::vector<uint8_t> cd, kd((size_t)n_kf * 32, 0);
  flatten_keys(CurrentFrame.mvKeysUn, ck);
  flatten_rows(CurrentFrame.mDescriptors, n_cur, cd);
  std::vector<int32_t> valid(n_kf, 0);
  std::vector<float> xyz((size_t)n_kf * 3, 0.f), maxd(n_kf, 0.f), mind(n_kf, 0.f), maxD(n_kf, 1.f), angle(n_kf, 0.f);
  for (int i = 0; i < n_kf; ++i) {
    MapPoint* p = vpMPs[i];
    angle[i] = pKF->mvKeysUn[i].angle;
    if (!p || p->isBad() || sAlreadyFound.count(p)) continue;
    valid[i] = 1;
    const cv::Mat x = p->GetWorldPos();
    for (int c = 0; c < 3; ++c) xyz[3 * (size_t)i + c] = x.at<float>(c);
    maxd[i] = p->GetMaxDistanceInvariance();
    mind[i] = p->GetMinDistanceInvariance();
    maxD[i] = MapPointPeek::max_distance(p);
    const cv::Mat desc = p->GetDescriptor();
    std::memcpy(&kd[(size_t)i * 32], desc.ptr(0), 32);
  }
  std::vector<int32_t> cmp(n_cur, -1);
  for (int i = 0; i < n_cur; ++i)
    if (CurrentFrame.mvpMapPoints[i]) cmp[i] = n_kf;
  float Tc[16];
  flatten_pose(CurrentFrame.mTcw, Tc);
  const orbm_camera cam = {CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy, CurrentFrame.mb, CurrentFrame.mbf};
  int nmatches = 0;
  Session s;
  s.check(orbm_search_by_projection_keyframe_host(s.m, ck.data(), cd.data(), n_cur, frame_bounds(), CurrentFrame.mvScaleFactors.data(),
                                                  (int)CurrentFrame.mvScaleFactors.size(), CurrentFrame.mfLogScaleFactor, cam, Tc,
                                                  valid.data(), xyz.data(), maxd.data(), mind.data(), maxD.data(), angle.data(),
                                                  kd.data(), n_kf, th, ORBdist, mbCheckOrientation ? 1 : 0, cmp.data(), &nmatches));
  for (int i = 0; i < n_cur; ++i) {
    if (cmp[i] >= 0 && cmp[i] < n_kf) CurrentFrame.mvpMapPoints[i] = vpMPs[cmp[i]];
    else if (cmp[i] < 0) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);
  }
  return nmatches;
}

// ---- src/ORBmatcher.cc:566-752 ------------------------------------------------------------------------------
// Loop-closing search (src/LoopClosing.cc:536): every point is projected through the Sim3 into both cameras.
// reads pKF: fx, fy, cx, cy, mnMinX..mnMaxY, mvScaleFactors, mfLogScaleFactor, mvKeysUn_total, keypoint_to_cam,
//            cont_idx_to_local_cam_idx, mDescriptors_total, per-camera grids; per point isBad(), GetWorldPos, GetNormal,
//            Get{Max,Min}DistanceInvariance, PredictScale (mfMaxDistance), GetDescriptor;  writes vpMatched.
//       (vLoopMPCams is not read by the reference's body either.)
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<int>& vLoopMPCams,
                                   vector<MapPoint*>& vpMatched, int th, const cv::Mat CalibMatrix) {
  (void)vLoopMPCams;
  const int n_kf = (int)pKF->mvKeysUn_total.size(), n_mp = (int)vpPoints.size();
  if (n_kf == 0 || n_mp == 0) return 0;
  std::vector<orbx_keypoint> kk;
  std::vector<uint8_t> kd, md((size_t)n_mp * 32, 0);
  std::vector<int32_t> kcam;
  flatten_keys(pKF->mvKeysUn_total, kk);
  flatten_rig(*pKF, n_kf, kd, kcam);
  std::set<MapPoint*> found(vpMatched.begin(), vpMatched.end());
  found.erase(static_cast<MapPoint*>(NULL));
  std::vector<int32_t> valid(n_mp, 0);
  std::vector<float> xyz((size_t)n_mp * 3, 0.f), normal((size_t)n_mp * 3, 0.f), maxd(n_mp, 0.f), mind(n_mp, 0.f), maxD(n_mp, 1.f);
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* p = vpPoints[i];
    if (p->isBad() || found.count(p)) continue;
    valid[i] = 1;
    const cv::Mat x = p->GetWorldPos(), nrm = p->GetNormal();
    for (int c = 0; c < 3; ++c) { xyz[3 * (size_t)i + c] = x.at<float>(c); normal[3 * (size_t)i + c] = nrm.at<float>(c); }
    maxd[i] = p->GetMaxDistanceInvariance();
    mind[i] = p->GetMinDistanceInvariance();
    maxD[i] = MapPointPeek::max_distance(p);
    const cv::Mat desc = p->GetDescriptor();
    std::memcpy(&md[(size_t)i * 32], desc.ptr(0), 32);
  }
  std::vector<int32_t> matched(n_kf, -1);
  for (int i = 0; i < n_kf; ++i)
    if (vpMatched[i]) matched[i] = n_mp;  // matched before this call
  float S[16], calib[12];
  flatten_pose(Scw, S);
  flatten_calib(CalibMatrix, calib);
  const orbm_camera cam = {pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mb, pKF->mbf};
  const orbm_bounds b = {(float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY};
  int nmatches = 0;
  Session s;
  s.check(orbm_search_by_projection_sim3_host(s.m, kk.data(), kd.data(), kcam.data(), n_kf, b, pKF->mvScaleFactors.data(),
                                              (int)pKF->mvScaleFactors.size(), pKF->mfLogScaleFactor, cam, S, calib, valid.data(),
                                              xyz.data(), normal.data(), maxd.data(), mind.data(), maxD.data(), md.data(), n_mp, th,
                                              matched.data(), &nmatches));
  for (int i = 0; i < n_kf; ++i)
    if (matched[i] >= 0 && matched[i] < n_mp) vpMatched[i] = vpPoints[matched[i]];
  return nmatches;
}

}  // namespace ORB_SLAM2
