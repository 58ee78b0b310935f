// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// C API of the oracle: a scalar CPU restatement of the reference's ORB front end
// (/root/reference/src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc grid functions).
// Two libraries export the extractor half of this API:
//   oracle/liborb_oracle.so      the restatement ("port"), built from oracle/*.cc
//   oracle/_ref/liborb_ref.so    the reference's own ORBextractor.cc compiled verbatim from
//                                /root/reference against oracle/shim (see oracle/Makefile)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load them.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Same field order as the product's orbx_keypoint (include/orb_b200.h) and as the
// reference's cv::KeyPoint minus class_id.
typedef struct {
  float x, y, size, angle, response;
  int32_t octave;
} oo_keypoint;

typedef struct oo_extractor oo_extractor;

oo_extractor* oo_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void oo_destroy(oo_extractor*);
// ORBextractor::operator() (src/ORBextractor.cc:1044-1107).  Returns the keypoint count
// (may exceed nfeatures by up to 3 per level), or -1 if cap is too small.
// level_counts (nlevels ints, may be NULL) receives the per-level counts.
int oo_extract(oo_extractor*, const uint8_t* img, int rows, int cols, size_t stride,
               oo_keypoint* kps, uint8_t* desc, int cap, int* level_counts);
// mvImagePyramid[level] of the last oo_extract call: pointer to the interior ROI (the 19-px
// border lies around it at negative offsets / beyond w,h).
int oo_pyramid_level(oo_extractor*, int level, const uint8_t** data, int* w, int* h, size_t* step);
void oo_scale_tables(oo_extractor*, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
void oo_features_per_level(oo_extractor*, int* out);

// --- stage taps, restatement only (liborb_oracle.so) ---
// FAST candidates of `level` fed to DistributeOctTree, (x,y) relative to (16,16), in
// vToDistributeKeys order; returns the count (fills at most cap).
int oo_stage_candidates(oo_extractor*, int level, int* x, int* y, int* score, int cap);
// Blurred working image of `level` (w*h bytes, stride w); returns 0 if the level had no keypoints.
int oo_stage_blurred(oo_extractor*, int level, uint8_t* out);
// Keypoints selected on `level` in level coordinates (before scaling), octree list order.
int oo_stage_level_keypoints(oo_extractor*, int level, int* x, int* y, float* response, float* angle, int cap);

// --- matcher restatement (liborb_oracle.so only) ---
// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:3994-4010).
int om_distance(const uint8_t* a, const uint8_t* b);

// Brute-force best/second-best scan with ratio test, the inner loop shared by the Search*
// functions (src/ORBmatcher.cc:895-924 without the window): for each query, scan targets in
// ascending index with strict-< updates; accepted iff best<=th_dist && best < ratio*second.
// out_idx = accepted target or -1; out_d1/out_d2 = best / second-best distance (256 if none).
void om_bruteforce(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int th_dist,
                   int* out_idx, int* out_d1, int* out_d2);

typedef struct {
  float min_x, max_x, min_y, max_y;  // mnMinX.. (src/Frame.cc:262-278); image bounds when undistorted
} om_bounds;

// Frame::GetFeaturesInArea over the 64x48 grid (src/Frame.cc:348-395,510-566,632-642).
// Returns the number of indices written to out (traversal order ix, iy, insertion).
int om_features_in_area(const float* kx, const float* ky, const int* koct, int n, om_bounds b,
                        float x, float y, float r, int min_level, int max_level, int* out, int cap);

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:868-983).  prev_xy (n1 x 2) is
// vbPrevMatched (in/out); matches12 (n1) out; returns nmatches.
int om_search_for_initialization(const oo_keypoint* k1, const uint8_t* d1, int n1,
                                 const oo_keypoint* k2, const uint8_t* d2, int n2, om_bounds b2,
                                 float* prev_xy, int window, float nnratio, int check_ori,
                                 int* matches12);

// Flattened MapPoint fields read by SearchByProjection(Frame&, vector<MapPoint*>&, th)
// (src/ORBmatcher.cc:62-149).
typedef struct {
  float proj_x, proj_y, proj_xr;  // mTrackProjX/Y/XR
  float view_cos;                 // mTrackViewCos
  int32_t level;                  // mnTrackScaleLevel
  int32_t track_in_view;          // mbTrackInView
  int32_t bad;                    // isBad()
} om_mappoint;

// frame_mp (n, in/out): index of the map point held by each keypoint (-1 = none), i.e.
// F.mvpMapPoints; frame_mp_obs (n): Observations()>0 of the point initially held (ignored
// when frame_mp<0).  mp_obs (nmp): Observations()>0 of each projected map point (needed
// because a keypoint assigned in this call is tested again by later iterations).
int om_search_by_projection_points(const oo_keypoint* k, const uint8_t* d, const float* u_right,
                                   int n, om_bounds b, const float* scale_factors, int nlevels,
                                   const om_mappoint* mp, const uint8_t* mp_desc, const int* mp_obs,
                                   int nmp, float th, float nnratio, int* frame_mp,
                                   const int* frame_mp_obs);

// Flattened inputs of the pose-based SearchByProjection overloads.
typedef struct {
  float fx, fy, cx, cy, mb, mbf;  // Frame::fx.., mb (baseline), mbf
} om_camera;

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono, CalibMatrix)
// (src/ORBmatcher.cc:3448-3641), the tracking matcher of the two-camera rig.
//  current frame: keypoints of all cameras concatenated (mvKeysUn_total), descriptor per global
//    index, u_right (mvuRight_total), cur_cam[i] (keypoint_to_cam), Tcw_cur (4x4 row-major),
//    cur_mp (n_cur, in/out: index into the last-frame arrays or -1), cur_mp_obs (Observations()>0
//    of the initially held points; may be NULL);
//  last frame: Tcw_last, last_k (octave and angle used), last_cam, last_valid[i] = has a map point
//    and is not an outlier, last_xyz (n_last x 3 world position), last_desc (MapPoint descriptor),
//    last_obs (Observations()>0 of that point);
//  calib: 4x3 row-major (rows 0-2 = R_cam12, row 3 = t_cam12; src/System.cc:63-72).
int om_search_by_projection_frame(const oo_keypoint* cur_k, const uint8_t* cur_desc, const float* cur_uright,
                                  const int32_t* cur_cam, int n_cur, om_bounds b, const float* scale_factors,
                                  int nlevels, om_camera cam, const float* Tcw_cur, const float* Tcw_last,
                                  const oo_keypoint* last_k, const int32_t* last_cam, const int32_t* last_valid,
                                  const float* last_xyz, const uint8_t* last_desc, const int32_t* last_obs,
                                  int n_last, const float* calib, float th, int mono, int check_ori,
                                  int32_t* cur_mp, const int32_t* cur_mp_obs);

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:3809-3937), relocalisation refinement (camera-1 data only).
//  kf_valid[i] = map point exists, !isBad(), not in sAlreadyFound; kf_xyz world position;
//  kf_max_dist / kf_min_dist = GetMaxDistanceInvariance() / GetMinDistanceInvariance();
//  kf_max_d = mfMaxDistance (PredictScale, src/MapPoint.cc:602-617); kf_angle = pKF->mvKeysUn[i].angle;
//  log_scale_factor = Frame::mfLogScaleFactor.  cur_mp (n_cur, in/out): any value >= 0 = occupied.
int om_search_by_projection_keyframe(const oo_keypoint* cur_k, const uint8_t* cur_desc, int n_cur, om_bounds b,
                                     const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam,
                                     const float* Tcw_cur, const int32_t* kf_valid, const float* kf_xyz,
                                     const float* kf_max_dist, const float* kf_min_dist, const float* kf_max_d,
                                     const float* kf_angle, const uint8_t* kf_desc, int n_kf, float th, int orb_dist,
                                     int check_ori, int32_t* cur_mp);

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vLoopMPCams, vpMatched, th, CalibMatrix)
// (src/ORBmatcher.cc:566-752): loop-closing search in BOTH cameras of the key frame, best over cameras.
//  key frame: concatenated keypoints (mvKeysUn_total), descriptor per global index, kf_cam
//    (keypoint_to_cam), matched (n_kf, in/out: index into the point arrays or -1 = vpMatched);
//  points: mp_valid[i] = !isBad() && not already in vpMatched; world position, normal
//    (GetNormal), GetMax/MinDistanceInvariance, mfMaxDistance, descriptor;
//  Scw: 4x4 row-major Sim3 (s*R | t); calib as above; log_scale_factor = KeyFrame::mfLogScaleFactor.
int om_search_by_projection_sim3(const oo_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf,
                                 om_bounds b, const float* scale_factors, int nlevels, float log_scale_factor,
                                 om_camera cam, const float* Scw, const float* calib, const int32_t* mp_valid,
                                 const float* mp_xyz, const float* mp_normal, const float* mp_max_dist,
                                 const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp,
                                 int th, int32_t* matched);

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:381-424), batched: descriptors of point p are rows
// offsets[p] .. offsets[p+1]-1 of desc; best_idx[p] = chosen row relative to offsets[p] (-1 for an empty set).
void om_compute_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx);

// Frame glue (src/Frame.cc): UndistortKeyPoints :673-706, ComputeImageBounds :743-779,
// ComputeStereoFromRGBD :959-985, AssignFeaturesToGrid :348-395.  dist5 = mDistCoef (k1, k2, p1, p2, k3).
void om_undistort_keypoints(const oo_keypoint* k, int n, float fx, float fy, float cx, float cy, const float* dist5,
                            oo_keypoint* k_un);
void om_compute_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float* dist5, om_bounds* b);
void om_compute_stereo_from_rgbd(const oo_keypoint* k, const oo_keypoint* k_un, int n, const float* depth, int cols,
                                 int rows, size_t stride_floats, float mbf, float* uright, float* depth_out);
// Frame::ComputeStereoMatches (src/Frame.cc:782-956; upstream ORB-SLAM2's rectified-stereo association, kept
// commented in this fork — SURVEY.md §8(f) rank 4).  pyr_l / pyr_r: mvImagePyramid[level] of the two extractors
// (data -> the level's ROI origin; the 19-px EDGE_THRESHOLD border around it is read by the 11x11 SAD windows).
// uright / depth: mvuRight / mvDepth (-1 where no match).  An empty match list skips the median filter (the
// reference would index an empty vector).
typedef struct { const uint8_t* data; int32_t w, h; size_t step; } om_image;
void om_compute_stereo_matches(const oo_keypoint* kl, const uint8_t* dl, int nl, const oo_keypoint* kr, const uint8_t* dr,
                               int nr, const om_image* pyr_l, const om_image* pyr_r, int nlevels, const float* scale_factors,
                               const float* inv_scale_factors, float mbf, float mb, float* uright, float* depth);
// CSR over cell = ix*48 + iy (cell_start: 64*48 + 1 entries), items in insertion order.
void om_assign_features_to_grid(const oo_keypoint* k_un, int n, om_bounds b, int32_t* cell_start, int32_t* items);

// ORBmatcher::SearchByBoW / SearchByBoW_cam1, Frame and KeyFrame variants (src/ORBmatcher.cc:206-388,
// 390-565, 996-1163, 1180-1363).  Side 1 = the key frame whose map points are searched, side 2 = the
// frame / second key frame.  Feature vectors are DBoW2::FeatureVector flattened to CSR: node ids
// ascending (std::map order), start[nn+1], items = feature indices in vector order.
//  valid1[i] = map point exists && !isBad() (&& i < N for the _cam1 variants); valid2 likewise for the
//  KeyFrame variants, (i < N) for SearchByBoW_cam1(KeyFrame*, Frame&), NULL = all valid.
//  max_dist = TH_LOW (Frame variants, `<=`) or TH_LOW-1 (KeyFrame variants, `<`).
//  matches12 (n1): matched side-2 feature or -1; matches21 (n2): matched side-1 feature or -1
//  (vpMapPointMatches[i2] = map point of side-1 feature matches21[i2]).  A feature index must occur at
//  most once per feature vector (DBoW2 guarantees it).
int om_search_by_bow(const uint8_t* d1, const float* angle1, const int32_t* valid1, int n1,
                     const int32_t* node1, const int32_t* start1, const int32_t* items1, int nn1,
                     const uint8_t* d2, const float* angle2, const int32_t* valid2, int n2,
                     const int32_t* node2, const int32_t* start2, const int32_t* items2, int nn2,
                     float nnratio, int check_ori, int max_dist, int32_t* matches12, int32_t* matches21);

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1364-1720) with CheckDistEpipolarLine (:167-184).
//  key frames: concatenated keypoints (mvKeysUn_total: pt, angle, octave used), descriptor per global
//    index, has_mp[i] = GetMapPoint(i) != NULL, cam[i] = keypoint_to_cam, uright[i] = mvuRight_total, CSR
//    feature vector (mFeatVec);
//  F12s: two row-major 3x3 fundamental matrices (camera 0, camera 1; :1421-1423); epipoles: (ex, ey) of
//    camera 0 then camera 1 (:1441-1449); scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors /
//    mvLevelSigma2; cam_enabled = vbCam; matches12 (n1) = vMatches12 (vMatchedPairs = its non-negative
//    entries in index order).
int om_search_for_triangulation(const oo_keypoint* k1, const uint8_t* d1, const int32_t* has_mp1, const int32_t* cam1,
                                const float* uright1, int n1, const int32_t* node1, const int32_t* start1,
                                const int32_t* items1, int nn1, const oo_keypoint* k2, const uint8_t* d2,
                                const int32_t* has_mp2, const int32_t* cam2, const float* uright2, int n2,
                                const int32_t* node2, const int32_t* start2, const int32_t* items2, int nn2,
                                const float* F12s, const float* epipoles, const float* scale_factors2,
                                const float* level_sigma2_2, int only_stereo, const int32_t* cam_enabled, int check_ori,
                                int32_t* matches12);

// ORBmatcher::Fuse(KeyFrame*, vpMapPoints, CalibMatrix, th) (src/ORBmatcher.cc:1986-2190), the search part:
// best_idx[2*i + cam] = key-frame feature that map point i fuses with in camera cam, or -1; returns nFused.
// Ow = {GetCameraCenter(), GetCameraCenter_cam2()}.  The Replace/AddObservation side effects stay with the caller.
int om_fuse(const oo_keypoint* kf_k, const uint8_t* kf_desc, const float* kf_uright, const int32_t* kf_cam, int n_kf,
            om_bounds b, const float* scale_factors, const float* inv_level_sigma2, int nlevels, float log_scale_factor,
            om_camera cam, const float* Tcw, const float* Ow, const float* calib, const int32_t* mp_valid,
            const float* mp_xyz, const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist,
            const float* mp_max_d, const uint8_t* mp_desc, int n_mp, float th, int32_t* best_idx);

// ORBmatcher::Fuse(KeyFrame*, Scw, vpPoints, vLoopMPCams, th, vpReplacePoint, CalibMatrix) (src/ORBmatcher.cc:2211-2441),
// the search part; Scw 4x4 row-major Sim3 (s*R | t).  mp_valid[i] = !isBad() && !spAlreadyFound.count(pMP).
int om_fuse_sim3(const oo_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf, om_bounds b,
                 const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam, const float* Scw,
                 const float* calib, const int32_t* mp_valid, const float* mp_xyz, const float* mp_normal,
                 const float* mp_max_dist, const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp,
                 float th, int32_t* best_idx);

// ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:2814-3136): mutual projection search between two key frames through a
// Sim3 (s12, R12 row-major 3x3, t12).  Map-point arrays are aligned with the keypoints of their key frame;
// mpx_valid[i] = point exists, !isBad(), not already matched.  match12 (n1): new mutual matches (idx2) or -1.
int om_search_by_sim3(const oo_keypoint* k1, const uint8_t* d1, const int32_t* cam1, int n1, const float* T1w,
                      const oo_keypoint* k2, const uint8_t* d2, const int32_t* cam2, int n2, const float* T2w, om_bounds b,
                      const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam, float s12,
                      const float* R12, const float* t12, const float* calib, const int32_t* mp1_valid, const float* mp1_xyz,
                      const float* mp1_max_dist, const float* mp1_min_dist, const float* mp1_max_d, const uint8_t* mp1_desc,
                      const int32_t* mp2_valid, const float* mp2_xyz, const float* mp2_max_dist, const float* mp2_min_dist,
                      const float* mp2_max_d, const uint8_t* mp2_desc, float th, int32_t* match12);

// DBoW2 vocabulary transform as Frame::ComputeBoW calls it (TemplatedVocabulary.h:1127-1195, 1218-1259; src/Frame.cc:657):
// per-feature word / node (level L - levelsup) / weight, the L1-normalised TF-IDF BowVector and the FeatureVector (CSR).
void om_bow_transform(const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc, const int32_t* word_id,
                      const double* node_weight, int n_nodes, int L, const uint8_t* desc, int n, int levelsup, int32_t* word,
                      int32_t* node, double* weight, int32_t* bow_word, double* bow_value, int32_t* n_bow, int32_t* fv_node,
                      int32_t* fv_start, int32_t* fv_items, int32_t* n_fv);

// Test hook for the cv::Mat 3x3 algebra emulation used by the pose-based searches (pinned against cv2.gemm).
void om_gemm3_probe(const float* A, const float* x, const float* c, float alpha, int transpose_a, float* out);

// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:3948-3989) on bin counts.
void om_three_maxima(const int* counts, int L, int* ind1, int* ind2, int* ind3);

#ifdef __cplusplus
}
#endif
