/* orb_b200.h — C ABI of the B200-native ORB front end (liborb_b200.so).
 *
 * Drop-in boundary for the per-frame hot path of AlterPang/Multi_ORB_SLAM: every entry point
 * names the reference interface it replaces (paths relative to the reference tree).  The
 * reference has no FFI of its own (single C++ process); include/ORBextractor.h and
 * include/ORBmatcher.h in this repo are the C++ classes, shaped like the reference's, that a
 * maintainer would compile in place of src/ORBextractor.cc / the hot members of
 * src/ORBmatcher.cc.  See INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes; int status return (0 = ok, <0 = error, see ORBX_E_*);
 * nothing throws across the boundary.  "_host" entry points take host buffers and include the
 * H2D/D2H copies; "_device" entry points take device pointers valid on the handle's GPU and are
 * asynchronous on the handle's stream (orbx_sync to wait).  There is no CPU fallback: without a
 * CUDA device every create call fails with ORBX_E_CUDA.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_E_INVALID (-1)  /* bad argument / unsupported geometry */
#define ORBX_E_CUDA (-2)     /* CUDA runtime error (see orbx_last_error) */
#define ORBX_E_CAPACITY (-3) /* caller-provided capacity too small */
#define ORBX_E_STATE (-4)    /* call order (e.g. stage tap before extract) */

#define ORBX_MAX_LEVELS 16

/* cv::KeyPoint without class_id (reference: include/ORBextractor.h:68-70 outputs
 * std::vector<cv::KeyPoint>; fields written at src/ORBextractor.cc:838-848,1096-1103). */
typedef struct {
  float x, y;     /* pt, level-0 pixel coordinates (scaled by mvScaleFactor[octave]) */
  float size;     /* (int)(31 * mvScaleFactor[octave]) */
  float angle;    /* IC_Angle, degrees [0,360) */
  float response; /* FAST score */
  int32_t octave; /* pyramid level */
} orbx_keypoint;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (src/ORBextractor.cc:411-471) plus the geometry the device workspace is sized for. */
typedef struct {
  int32_t nfeatures;
  float scale_factor;
  int32_t nlevels;
  int32_t ini_th_fast;
  int32_t min_th_fast;
  int32_t width, height; /* image size every frame of this handle has */
  int32_t max_batch;     /* frames processed per launch group (workspace size) */
  int32_t device;        /* CUDA device ordinal, -1 = current */
} orbx_config;

typedef struct orbx_extractor orbx_extractor;

/* ---- extractor: replaces class ORBextractor (include/ORBextractor.h:45-112) ---------------- */

/* Environment read once, when the handle is created (tuning / A-B runs; results are identical either way):
 *   ORB_B200_FAST=bands   run the FAST stage with the streaming band kernel instead of the per-cell kernel
 *   ORB_OT_SMEM_KEYS=n    octree keys kept in shared memory per CTA */
int orbx_create(const orbx_config* cfg, orbx_extractor** out);
void orbx_destroy(orbx_extractor* h);
/* Last error text of this handle (or of the failed create when h == NULL). */
const char* orbx_last_error(const orbx_extractor* h);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:71-90); each array nlevels floats, NULL to skip. */
int orbx_get_scale_tables(const orbx_extractor* h, float* scale, float* inv_scale, float* sigma2,
                          float* inv_sigma2);
/* mnFeaturesPerLevel (src/ORBextractor.cc:436-447). */
int orbx_get_features_per_level(const orbx_extractor* h, int32_t* out);
/* Safe per-frame output capacity: nfeatures + 3*nlevels (the octree may overshoot each level's
 * quota by up to 3, src/ORBextractor.cc:731-732). */
int orbx_max_keypoints(const orbx_extractor* h);

/* ORBextractor::operator()(image, mask, keypoints, descriptors) (src/ORBextractor.cc:1044-1107)
 * for ONE frame in host memory.  image: rows x cols u8, `stride` bytes per row.  kps / desc have
 * room for `cap` keypoints (desc: cap x 32 bytes).  *n receives the count. */
int orbx_extract(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t stride,
                 orbx_keypoint* kps, uint8_t* desc, int cap, int* n);

/* The same for a batch of frames (one camera stream, or many cameras with identical settings):
 * frame f starts at images + f*frame_stride.  Outputs are strided by `cap` keypoints per frame;
 * counts[f] receives frame f's count.  Host buffers; pinned memory makes the copies async. */
int orbx_extract_batch_host(orbx_extractor* h, const uint8_t* images, int n_frames, size_t frame_stride,
                            size_t row_stride, orbx_keypoint* kps, uint8_t* desc, int32_t* counts, int cap);

/* Device-resident variant: all pointers are device pointers; asynchronous on the handle's stream.
 * counts[f] may exceed cap (the count is exact, the arrays are clamped) — check after orbx_sync. */
int orbx_extract_batch_device(orbx_extractor* h, const uint8_t* d_images, int n_frames, size_t frame_stride,
                              size_t row_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_counts,
                              int cap);
int orbx_sync(orbx_extractor* h);
/* cudaStream_t of the handle as an opaque pointer (to order foreign work after it). */
void* orbx_stream(orbx_extractor* h);
/* Run this handle's work on a caller-owned cudaStream_t (NULL restores the handle's own stream),
 * e.g. to chain extractor -> matcher without host synchronisation. */
int orbx_set_stream(orbx_extractor* h, void* cuda_stream);

/* mvImagePyramid[level] (public member, include/ORBextractor.h:92) of frame `frame` of the last
 * batch, copied to host: with_border != 0 returns the (w+38) x (h+38) buffer including the
 * EDGE_THRESHOLD border, else the w x h interior.  dst may be NULL to query the size. */
int orbx_get_pyramid_level(orbx_extractor* h, int frame, int level, int with_border, uint8_t* dst,
                           size_t dst_stride, int* w, int* hgt);

/* Device view of the pyramid the handle holds for its LAST batch: what a consumer of the public member
 * mvImagePyramid (include/ORBextractor.h:92) reads, without leaving the GPU.  Levels are stored without the
 * EDGE_THRESHOLD border (a reader reflects out-of-range coordinates as BORDER_REFLECT_101 does); level 0 is the
 * caller's own frame buffer of that batch, which must still be alive.  Pixel (x, y) of level l of frame f is
 * base[l][f * frame_stride[l] + y * pitch[l] + x]. */
typedef struct orbx_pyramid_view {
  int32_t nlevels, n_frames;
  int32_t w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS], pitch[ORBX_MAX_LEVELS];
  const uint8_t* base[ORBX_MAX_LEVELS];
  size_t frame_stride[ORBX_MAX_LEVELS];
  float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS]; /* mvScaleFactor / mvInvScaleFactor */
  void* stream; /* cudaStream_t the batch was produced on: a consumer on another stream must order itself after it */
} orbx_pyramid_view;
int orbx_get_pyramid_view(orbx_extractor* h, orbx_pyramid_view* out);

/* Stage taps for parity tests (frame of the last batch): FAST candidates fed to the octree in
 * vToDistributeKeys order ((x,y) relative to (16,16)); the blurred working image. */
int orbx_debug_candidates(orbx_extractor* h, int frame, int level, int32_t* x, int32_t* y, int32_t* score,
                          int cap, int* n);
int orbx_debug_blurred(orbx_extractor* h, int frame, int level, uint8_t* dst, size_t dst_stride);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
long long orbx_launch_count(const orbx_extractor* h);
/* Per-stage device time (CUDA events recorded on the handle's stream around each stage's
 * launches).  orbx_set_profiling(h, 1) clears the sums and starts recording on every launch
 * group; orbx_stage_times_ms returns the summed ms per stage since then and the number of launch
 * groups (batches of <= max_batch frames) they cover.  stages: 0 pyramid, 1 fast, 2 octree,
 * 3 blur, 4 orientation+descriptor. */
int orbx_set_profiling(orbx_extractor* h, int enable);
int orbx_stage_times_ms(orbx_extractor* h, double* sum_ms5, long long* n_calls);

/* ---- matcher: replaces the hot members of class ORBmatcher (include/ORBmatcher.h:37-137) --- */

typedef struct orbm_matcher orbm_matcher;

int orbm_create(int device, orbm_matcher** out);
void orbm_destroy(orbm_matcher* m);
const char* orbm_last_error(const orbm_matcher* m);
int orbm_sync(orbm_matcher* m);
int orbm_set_stream(orbm_matcher* m, void* cuda_stream);
long long orbm_launch_count(const orbm_matcher* m);

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:3994-4010) for n independent pairs:
 * out[i] = Hamming(a[i], b[i]) over 32-byte rows.  Host buffers. */
int orbm_distance_pairs_host(orbm_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* out);

/* Brute-force best / second-best Hamming scan with the ratio test of the Search* functions
 * (src/ORBmatcher.cc:895-924 without the window): for each query, targets are scanned in
 * ascending index with strict-< updates; out_idx = best target if best <= th_dist and
 * (float)best < ratio*(float)second, else -1; out_d1/out_d2 = best / second-best distance
 * (256 when absent).  _device: device pointers, async. */
int orbm_bruteforce_device(orbm_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, float ratio,
                           int th_dist, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2);
int orbm_bruteforce_host(orbm_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio,
                         int th_dist, int32_t* idx, int32_t* d1, int32_t* d2);

/* The same scan for a batch of small independent pairs — cross-camera matching of a rig, the
 * per-camera loops the reference runs over mDescriptors_total (src/ORBmatcher.cc:628,2030,2269).
 * Pair p: query rows at d_q + p*q_stride bytes (d_nq[p] valid, <= cap), target rows at
 * d_t + p*t_stride (d_nt[p] valid); results at [p*cap + i].  Strides are multiples of 16 bytes.
 * Device pointers, asynchronous. */
int orbm_bruteforce_batch_device(orbm_matcher* m, int n_pairs, int cap, const uint8_t* d_q, const int32_t* d_nq,
                                 size_t q_stride, const uint8_t* d_t, const int32_t* d_nt, size_t t_stride, float ratio,
                                 int th_dist, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2);

/* The same scan for pairs whose rows lie anywhere inside ONE device buffer — the all-gathered per-camera blocks
 * of a multi-GPU rig (orbd_allgather_inplace below): pair p takes its query rows at d_base + d_q_off[p], its
 * target rows at d_base + d_t_off[p] and its two row counts (int32) at d_base + d_nq_off[p] / d_nt_off[p]; all
 * offsets in bytes, rows 16-byte aligned.  One launch for all camera pairs and rig-frames of a chunk
 * (src/ORBmatcher.cc:628,3582 loop over the cameras of mDescriptors_total).  Results at [p*cap + i]. */
int orbm_bruteforce_indexed_device(orbm_matcher* m, int n_pairs, int cap, const uint8_t* d_base, const int64_t* d_q_off,
                                   const int64_t* d_t_off, const int64_t* d_nq_off, const int64_t* d_nt_off, float ratio,
                                   int th_dist, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2);

/* Frame grid bounds mnMinX/mnMaxX/mnMinY/mnMaxY (src/Frame.cc:262-278). */
typedef struct {
  float min_x, max_x, min_y, max_y;
} orbm_bounds;

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
 * (src/ORBmatcher.cc:868-983) for a batch of independent frame pairs.  Pair p uses keypoints
 * k1 + p*cap (n1[p] valid), k2 + p*cap (n2[p] valid), descriptors likewise (cap x 32 bytes per
 * frame), prev_xy + p*cap*2 (vbPrevMatched, in/out), matches12 + p*cap (out, -1 = none);
 * nmatches[p] out.  mfNNratio = nnratio, mbCheckOrientation = check_ori.  Device pointers.
 * d_prev_xy == NULL means vbPrevMatched[i] = F1 keypoint i's position (how Tracking initialises it,
 * src/Tracking.cc:844-846) and nothing is written back. */
int orbm_search_for_initialization_device(orbm_matcher* m, int n_pairs, int cap, const orbx_keypoint* d_k1,
                                          const uint8_t* d_d1, const int32_t* d_n1, const orbx_keypoint* d_k2,
                                          const uint8_t* d_d2, const int32_t* d_n2, orbm_bounds bounds2,
                                          float* d_prev_xy, int window, float nnratio, int check_ori,
                                          int32_t* d_matches12, int32_t* d_nmatches);
int orbm_search_for_initialization_host(orbm_matcher* m, int n_pairs, int cap, const orbx_keypoint* k1,
                                        const uint8_t* d1, const int32_t* n1, const orbx_keypoint* k2,
                                        const uint8_t* d2, const int32_t* n2, orbm_bounds bounds2, float* prev_xy,
                                        int window, float nnratio, int check_ori, int32_t* matches12,
                                        int32_t* nmatches);

/* Flattened MapPoint fields read by SearchByProjection(Frame&, const vector<MapPoint*>&, th)
 * (src/ORBmatcher.cc:62-149): mTrackProjX/Y/XR, mTrackViewCos, mnTrackScaleLevel,
 * mbTrackInView, isBad(). */
typedef struct {
  float proj_x, proj_y, proj_xr;
  float view_cos;
  int32_t level;
  int32_t track_in_view;
  int32_t bad;
} orbm_mappoint;

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th)
 * (src/ORBmatcher.cc:62-157).  Frame side: keypoints k[n] (mvKeysUn), descriptors, u_right[n]
 * (mvuRight, may be NULL = all -1), scale_factors[nlevels] (mvScaleFactors), frame_mp[n] in/out
 * (index of the map point held by each keypoint, -1 = none; F.mvpMapPoints), frame_mp_obs[n]
 * (Observations()>0 of the point initially held; may be NULL).  Map-point side: mp[nmp],
 * mp_desc (nmp x 32), mp_obs[nmp] (Observations()>0; NULL = all 1).  *nmatches out.  Host buffers. */
int orbm_search_by_projection_points_host(orbm_matcher* m, const orbx_keypoint* k, const uint8_t* desc,
                                          const float* u_right, int n, orbm_bounds bounds,
                                          const float* scale_factors, int nlevels, const orbm_mappoint* mp,
                                          const uint8_t* mp_desc, const int32_t* mp_obs, int nmp, float th,
                                          float nnratio, int32_t* frame_mp, const int32_t* frame_mp_obs,
                                          int* nmatches);

/* The single-call projection searches size their candidate-row buffer from the need of earlier calls (first guess:
 * rows_per_query x queries, default 64) so that no host round trip separates counting from filling; a call whose rows
 * do not fit is repeated once with the exact size.  Test tap: a budget of 1 forces that path.  Resets the history. */
int orbm_debug_set_row_budget(orbm_matcher* m, int rows_per_query);

/* Camera intrinsics / stereo fields of Frame read by the pose-based overloads (fx, fy, cx, cy,
 * mb = baseline, mbf). */
typedef struct {
  float fx, fy, cx, cy, mb, mbf;
} orbm_camera;

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono, CalibMatrix)
 * (src/ORBmatcher.cc:3448-3641) — the tracking matcher of the two-camera rig
 * (Tracking::TrackWithMotionModel, src/Tracking.cc:1267).  Current frame: keypoints of all cameras
 * concatenated (mvKeysUn_total), one descriptor row per global index, cur_uright (mvuRight_total,
 * may be NULL), cur_cam (keypoint_to_cam, may be NULL = all camera 0), Tcw_cur (mTcw, 4x4
 * row-major); cur_mp (n_cur, in/out: index into the last-frame arrays or -1 = mvpMapPoints),
 * cur_mp_obs (Observations()>0 of points held on entry, may be NULL).  Last frame: Tcw_last,
 * last_k (octave + angle), last_cam, last_valid[i] = mvpMapPoints[i] && !mvbOutlier[i], last_xyz
 * (GetWorldPos, n_last x 3), last_desc (GetDescriptor), last_obs (Observations()>0, may be NULL).
 * calib: the 4x3 CalibMatrix row-major (rows 0-2 R_cam12, row 3 t_cam12).  The projection runs on
 * the host in the reference's float evaluation order; search and resolve run on the GPU. */
int orbm_search_by_projection_frame_host(orbm_matcher* m, const orbx_keypoint* cur_k, const uint8_t* cur_desc,
                                         const float* cur_uright, const int32_t* cur_cam, int n_cur, orbm_bounds b,
                                         const float* scale_factors, int nlevels, orbm_camera cam, const float* Tcw_cur,
                                         const float* Tcw_last, const orbx_keypoint* last_k, const int32_t* last_cam,
                                         const int32_t* last_valid, const float* last_xyz, const uint8_t* last_desc,
                                         const int32_t* last_obs, int n_last, const float* calib, float th, int mono,
                                         int check_ori, int32_t* cur_mp, const int32_t* cur_mp_obs, int* nmatches);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist)
 * (src/ORBmatcher.cc:3809-3937) — relocalisation refinement (src/Tracking.cc:2099,2116).
 * kf_valid[i] = map point exists, !isBad(), not in sAlreadyFound; kf_max_dist / kf_min_dist =
 * Get{Max,Min}DistanceInvariance(); kf_max_d = mfMaxDistance (PredictScale, src/MapPoint.cc:602-617);
 * kf_angle = pKF->mvKeysUn[i].angle; log_scale_factor = Frame::mfLogScaleFactor.  cur_mp (n_cur,
 * in/out): any value >= 0 means the keypoint already holds a point. */
int orbm_search_by_projection_keyframe_host(orbm_matcher* m, const orbx_keypoint* cur_k, const uint8_t* cur_desc, int n_cur,
                                            orbm_bounds b, const float* scale_factors, int nlevels, float log_scale_factor,
                                            orbm_camera cam, const float* Tcw_cur, const int32_t* kf_valid,
                                            const float* kf_xyz, const float* kf_max_dist, const float* kf_min_dist,
                                            const float* kf_max_d, const float* kf_angle, const uint8_t* kf_desc, int n_kf,
                                            float th, int orb_dist, int check_ori, int32_t* cur_mp, int* nmatches);

/* Note on bounds for the KeyFrame-side searches below (SearchByProjection with Scw, Fuse, SearchBySim3): the reference's
 * KeyFrame stores mnMinX/mnMinY/mnMaxX/mnMaxY as ints (include/KeyFrame.h:234-237, truncated copies of the Frame's floats)
 * while its grid cell sizes come from the Frame's float bounds.  These entry points take one orbm_bounds and are exact when
 * the bounds are integral — undistorted or pre-rectified input, where ComputeImageBounds returns (0, cols, 0, rows). */
/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vLoopMPCams, vpMatched, th, CalibMatrix)
 * (src/ORBmatcher.cc:566-752) — loop-closing search (src/LoopClosing.cc:536): every map point is
 * projected through the Sim3 into BOTH cameras of the key frame, best candidate over cameras,
 * TH_LOW.  Key frame: concatenated keypoints (mvKeysUn_total), descriptor per global index, kf_cam
 * (keypoint_to_cam); matched (n_kf, in/out: index into the point arrays or -1 = vpMatched).
 * Points: mp_valid[i] = !isBad() && not already in vpMatched; GetWorldPos, GetNormal,
 * Get{Max,Min}DistanceInvariance, mfMaxDistance, GetDescriptor.  Scw 4x4 row-major. */
int orbm_search_by_projection_sim3_host(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc,
                                        const int32_t* kf_cam, int n_kf, orbm_bounds b, const float* scale_factors,
                                        int nlevels, float log_scale_factor, orbm_camera cam, const float* Scw,
                                        const float* calib, const int32_t* mp_valid, const float* mp_xyz,
                                        const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist,
                                        const float* mp_max_d, const uint8_t* mp_desc, int n_mp, int th, int32_t* matched,
                                        int* nmatches);

/* ---- Frame glue between the extractor and the matchers (src/Frame.cc), device-resident batches laid out
 * like the extractor's batch outputs (n_frames x cap keypoints, counts per frame) ------------------------- */

/* Frame::UndistortKeyPoints{,_cam2} (src/Frame.cc:673-706, 708-741): cv::undistortPoints(mat, mat, mK,
 * mDistCoef, Mat(), mK), OpenCV 4.x, default criteria (5 iterations), double arithmetic; dist5 (host) =
 * mDistCoef (k1, k2, p1, p2, k3; k3 = 0 for the four-parameter model); k1 == 0 copies (:675-679).  Every other
 * keypoint field is carried over.  The _host variant takes one frame in host memory. */
int orbm_undistort_keypoints_device(orbm_matcher* m, int n_frames, int cap, const orbx_keypoint* d_kps,
                                    const int32_t* d_counts, float fx, float fy, float cx, float cy, const float* dist5,
                                    orbx_keypoint* d_kps_un);
int orbm_undistort_keypoints_host(orbm_matcher* m, const orbx_keypoint* k, int n, float fx, float fy, float cx, float cy,
                                  const float* dist5, orbx_keypoint* k_un);
/* Frame::ComputeImageBounds (src/Frame.cc:743-779): mnMinX/mnMaxX/mnMinY/mnMaxY from the undistorted corners. */
int orbm_compute_image_bounds_host(orbm_matcher* m, int cols, int rows, float fx, float fy, float cx, float cy,
                                   const float* dist5, orbm_bounds* out);
/* Frame::ComputeStereoFromRGBD{,_cam2} (src/Frame.cc:959-985, 987-1010): d_depth = CV_32F depth images
 * (n_frames x rows x row_stride floats, already scaled by mDepthMapFactor) read at the DISTORTED keypoint
 * truncated to int; writes mvuRight and mvDepth (-1 where depth <= 0 and beyond counts[f]). */
int orbm_compute_stereo_from_rgbd_device(orbm_matcher* m, int n_frames, int cap, const orbx_keypoint* d_kps,
                                         const orbx_keypoint* d_kps_un, const int32_t* d_counts, const float* d_depth,
                                         int cols, int rows, size_t row_stride_floats, size_t frame_stride_floats, float mbf,
                                         float* d_uright, float* d_depth_out);
/* Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:348-395, 632-642): per frame a CSR grid over
 * cell = ix*48 + iy (d_cell_start: n_frames x (64*48 + 1) ints; d_items: n_frames x cap u16 keypoint indices,
 * insertion order inside a cell = mGrid[ix][iy]). */
/* Frame::ComputeStereoMatches (src/Frame.cc:782-956 — upstream ORB-SLAM2's rectified-stereo association, kept
 * commented in the fork) for n_frames stereo pairs at once, on the device: left / right are the pyramid views of the
 * two extractors that produced the keypoints (mpORBextractorLeft / Right ->mvImagePyramid), d_kl/d_dl/d_nl and
 * d_kr/d_dr/d_nr their batch outputs (strides cap_l, cap_r; cap_r <= 4096).  mbf = baseline * fx, mb = mbf / fx.
 * Outputs d_uright / d_depth [n_frames][cap_l] = mvuRight / mvDepth (-1 where unmatched).  A pair with no accepted
 * match skips the median cut (the reference would index an empty vector).  The call orders the matcher's stream after
 * the two extractor streams recorded in the views (an event wait on the device, no host synchronisation). */
int orbm_compute_stereo_matches_device(orbm_matcher* m, const orbx_pyramid_view* left, const orbx_pyramid_view* right,
                                       int n_frames, int cap_l, const orbx_keypoint* d_kl, const uint8_t* d_dl,
                                       const int32_t* d_nl, int cap_r, const orbx_keypoint* d_kr, const uint8_t* d_dr,
                                       const int32_t* d_nr, float mbf, float mb, float* d_uright, float* d_depth);

int orbm_assign_features_to_grid_device(orbm_matcher* m, int n_frames, int cap, const orbx_keypoint* d_kps_un,
                                        const int32_t* d_counts, orbm_bounds bounds, int32_t* d_cell_start,
                                        uint16_t* d_items);

/* ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, CalibMatrix, th) (src/ORBmatcher.cc:1986-2190,
 * called from LocalMapping::SearchInNeighbors, src/LocalMapping.cc:741,772), the search part: for every map point and
 * both cameras of the key frame, best_idx[2*i + cam] = the key-frame feature the point fuses with, or -1; *n_fused = number
 * of entries >= 0 (= nFused).  The side effects (:2160-2186: Replace / AddObservation / AddMapPoint, applied in map-point
 * order, camera 0 before camera 1) do not feed back into the search and stay with the caller.
 * Key frame: concatenated keypoints (mvKeysUn_total), descriptor per global index, kf_uright (mvuRight_total), kf_cam
 * (keypoint_to_cam, NULL = one camera), bounds (mnMinX..), mvScaleFactors, mvInvLevelSigma2, mfLogScaleFactor, camera
 * (fx, fy, cx, cy, mbf), Tcw 4x4 row-major, Ow = {GetCameraCenter(), GetCameraCenter_cam2()} (6 floats), calib 4x3.
 * Map points: mp_valid[i] = pMP && !isBad() && !IsInKeyFrame(pKF); GetWorldPos, GetNormal,
 * Get{Max,Min}DistanceInvariance, mfMaxDistance, GetDescriptor. */
int orbm_fuse_host(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc, const float* kf_uright,
                   const int32_t* kf_cam, int n_kf, orbm_bounds b, const float* scale_factors, const float* inv_level_sigma2,
                   int nlevels, float log_scale_factor, orbm_camera cam, const float* Tcw, const float* Ow, const float* calib,
                   const int32_t* mp_valid, const float* mp_xyz, const float* mp_normal, const float* mp_max_dist,
                   const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp, float th,
                   int32_t* best_idx, int* n_fused);

/* ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, vpPoints, vLoopMPCams, th, vpReplacePoint, CalibMatrix)
 * (src/ORBmatcher.cc:2211-2441, called from LoopClosing::SearchAndFuse, src/LoopClosing.cc:841), the search part, same
 * outputs as orbm_fuse_host.  Scw = 4x4 row-major Sim3 (s*R | t); no reprojection gate in this overload;
 * mp_valid[i] = !isBad() && !spAlreadyFound.count(pMP).  The caller keeps :2420-2437 (vpReplacePoint / AddObservation). */
int orbm_fuse_sim3_host(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf,
                        orbm_bounds b, const float* scale_factors, int nlevels, float log_scale_factor, orbm_camera cam,
                        const float* Scw, const float* calib, const int32_t* mp_valid, const float* mp_xyz,
                        const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist, const float* mp_max_d,
                        const uint8_t* mp_desc, int n_mp, float th, int32_t* best_idx, int* n_fused);

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th, CalibMatrix) (src/ORBmatcher.cc:2814-3136,
 * called from LoopClosing::ComputeSim3, src/LoopClosing.cc:402): the map points of each key frame are projected into
 * the other through the Sim3 (s12, R12 3x3 row-major, t12), searched in the grid of their own camera within
 * th * mvScaleFactors[predicted level], best distance <= TH_HIGH, and only mutual pairs are kept.
 * Key frame x: concatenated keypoints / descriptors, camx (keypoint_to_cam, NULL = one camera), Txw (4x4 row-major).
 * Map-point arrays are aligned with the keypoints (GetMapPointMatches): mpx_valid[i] = point exists, !isBad() and not
 * already matched (vbAlreadyMatched1/2, :2846-2861); GetWorldPos, Get{Max,Min}DistanceInvariance, mfMaxDistance,
 * GetDescriptor.  match12 (n1) out: the key-frame-2 feature of every new mutual match, else -1
 * (vpMatches12[i1] = vpMapPoints2[match12[i1]]); *n_found = nFound. */
int orbm_search_by_sim3_host(orbm_matcher* m, const orbx_keypoint* k1, const uint8_t* d1, const int32_t* cam1, int n1,
                             const float* T1w, const orbx_keypoint* k2, const uint8_t* d2, const int32_t* cam2, int n2,
                             const float* T2w, orbm_bounds b, const float* scale_factors, int nlevels, float log_scale_factor,
                             orbm_camera cam, float s12, const float* R12, const float* t12, const float* calib,
                             const int32_t* mp1_valid, const float* mp1_xyz, const float* mp1_max_dist,
                             const float* mp1_min_dist, const float* mp1_max_d, const uint8_t* mp1_desc,
                             const int32_t* mp2_valid, const float* mp2_xyz, const float* mp2_max_dist,
                             const float* mp2_min_dist, const float* mp2_max_d, const uint8_t* mp2_desc, float th,
                             int32_t* match12, int* n_found);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:325-438), the arithmetic part (:381-424), batched over
 * map points: the observed descriptors of point p (vDescriptors, in std::map<KeyFrame*,size_t> iteration order, bad
 * key frames left out) are rows offsets[p] .. offsets[p+1]-1 of desc (offsets[0] = 0).  best_idx[p] = BestIdx relative
 * to offsets[p]: the first descriptor whose median distance to the set (element int(0.5*(N-1)) of the sorted row,
 * self-distance included) is least; -1 for an empty set.  mDescriptor = vDescriptors[BestIdx]. */
int orbm_compute_distinctive_descriptors_host(orbm_matcher* m, const uint8_t* desc, const int32_t* offsets, int n_points,
                                              int32_t* best_idx);

/* DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>, Thirdparty/DBoW2/DBoW2/FeatureVector.h)
 * flattened to CSR: node ids ascending (the map's order), start[n_nodes + 1], items = feature indices in
 * vector order.  A feature index occurs at most once per vector (DBoW2 guarantees it). */
typedef struct {
  const int32_t* node_id;
  const int32_t* start;
  const int32_t* items;
  int32_t n_nodes;
} orbm_featvec;

/* DBoW2 vocabulary transform, the producer of the feature vectors above: Frame::ComputeBoW / KeyFrame::ComputeBoW
 * (src/Frame.cc:649-671, src/KeyFrame.cc) call ORBVocabulary::transform(vCurrentDesc, mBowVec, mFeatVec, 4)
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195; tree descent :1218-1259; BowVector.cpp:34-46, 62-84).
 * orbm_set_vocabulary uploads the tree once: children of node i = child_ids[child_start[i] .. child_start[i+1]) in
 * m_nodes[i].children order (node 0 = root), node descriptors (32 bytes each), word id of the leaves (-1 for inner
 * nodes), node weights (WordValue), L = depth levels (m_L).  TF / TF_IDF weighting with L1 scoring (ORBvoc.txt).
 * orbm_bow_transform_host: per feature word / node at level L - levelsup / weight (each may be NULL), the BowVector
 * (bow_word ascending, bow_value L1-normalised, capacity n) and the FeatureVector as CSR (fv_node ascending, capacity n;
 * fv_start n + 1; fv_items n), features with weight 0 (stopped words) left out. */
int orbm_set_vocabulary(orbm_matcher* m, const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc,
                        const int32_t* word_id, const double* node_weight, int n_nodes, int L);
int orbm_bow_transform_host(orbm_matcher* m, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight,
                            int32_t* bow_word, double* bow_value, int32_t* n_bow, int32_t* fv_node, int32_t* fv_start,
                            int32_t* fv_items, int32_t* n_fv);

/* ORBmatcher::SearchByBoW, all four variants:
 *   SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)        src/ORBmatcher.cc:206-388   (Tracking.cc:1238, 2031)
 *   SearchByBoW_cam1(KeyFrame*, Frame&, vpMapPointMatches)   src/ORBmatcher.cc:390-565
 *   SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)           src/ORBmatcher.cc:996-1163  (LoopClosing.cc:329)
 *   SearchByBoW_cam1(KeyFrame*, KeyFrame*, vpMatches12)      src/ORBmatcher.cc:1180-1363
 * Side 1 = the key frame whose map points are searched for, side 2 = the frame / second key frame.
 *   valid1[i] = map point exists && !isBad() (&& i < N for the _cam1 variants; NULL = all valid);
 *   valid2[i] likewise for the KeyFrame variants, (i < N) for SearchByBoW_cam1(KeyFrame*, Frame&), NULL = all;
 *   angle = mvKeysUn(_total)[i].angle of side 1, mvKeys(_total)[i].angle of side 2;
 *   max_dist = TH_LOW (Frame variants, `bestDist1<=TH_LOW`) or TH_LOW-1 (KeyFrame variants, `bestDist1<TH_LOW`).
 * matches12 (n1): matched side-2 feature or -1 (vpMatches12[i1] = vpMapPoints2[matches12[i1]]);
 * matches21 (n2, may be NULL): matched side-1 feature or -1 (vpMapPointMatches[i2] = map point of it). */
int orbm_search_by_bow_host(orbm_matcher* m, const uint8_t* desc1, const float* angle1, const int32_t* valid1, int n1,
                            orbm_featvec fv1, const uint8_t* desc2, const float* angle2, const int32_t* valid2, int n2,
                            orbm_featvec fv2, float nnratio, int check_ori, int max_dist, int32_t* matches12,
                            int32_t* matches21, int* nmatches);

/* The same for a batch of independent pairs in one call (e.g. the relocalisation candidates of
 * Tracking::Relocalization, src/Tracking.cc:2010-2040, or the loop candidates of LoopClosing::ComputeSim3,
 * src/LoopClosing.cc:310-340, each against its own partner): the ordered part of the search runs
 * concurrently, one CTA per pair.  Fields as the arguments above; nmatches is written per pair. */
typedef struct {
  const uint8_t* desc1; const float* angle1; const int32_t* valid1; int32_t n1; orbm_featvec fv1;
  const uint8_t* desc2; const float* angle2; const int32_t* valid2; int32_t n2; orbm_featvec fv2;
  int32_t* matches12; int32_t* matches21; /* out (matches21 may be NULL) */
  int32_t nmatches;                       /* out */
} orbm_bow_pair;
int orbm_search_by_bow_batch_host(orbm_matcher* m, orbm_bow_pair* pairs, int n_pairs, float nnratio, int check_ori,
                                  int max_dist);

/* ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vMatchedPairs, bOnlyStereo, vbCam)
 * (src/ORBmatcher.cc:1364-1720, called from LocalMapping.cc:361) with CheckDistEpipolarLine (:167-184).
 * Key frames: concatenated keypoints (mvKeysUn_total; pt, angle, octave are read), descriptor per global index,
 * has_mp[i] = (GetMapPoint(i) != NULL), cam[i] = keypoint_to_cam (0/1), uright[i] = mvuRight_total, mFeatVec as CSR.
 * F12s: the two row-major 3x3 fundamental matrices F12s[cam] = K1^-T [t12]x R12 K2^-1 (:1421-1423) and
 * epipoles = {ex, ey, ex_cam2, ey_cam2} (:1441-1449): cv::Mat algebra on the key-frame poses, evaluated by the
 * caller with its own OpenCV (the reference takes F12 as an argument as well).  scale_factors2 / level_sigma2_2 =
 * pKF2->mvScaleFactors / mvLevelSigma2 (nlevels entries); cam_enabled = vbCam (2 entries).
 * matches12 (n1) = vMatches12: key-frame-2 feature or -1; vMatchedPairs = its entries >= 0 in index order.
 * The reference never marks key-frame-2 features as matched (:1452), so they may repeat. */
int orbm_search_for_triangulation_host(orbm_matcher* m, const orbx_keypoint* k1, const uint8_t* desc1, const int32_t* has_mp1,
                                       const int32_t* cam1, const float* uright1, int n1, orbm_featvec fv1,
                                       const orbx_keypoint* k2, const uint8_t* desc2, const int32_t* has_mp2,
                                       const int32_t* cam2, const float* uright2, int n2, orbm_featvec fv2, const float* F12s,
                                       const float* epipoles, const float* scale_factors2, const float* level_sigma2_2,
                                       int nlevels, int only_stereo, const int32_t* cam_enabled, int check_ori,
                                       int32_t* matches12, int* nmatches);

/* The same for a batch of independent key-frame pairs in one call (LocalMapping::CreateNewMapPoints matches the
 * current key frame against every neighbour, src/LocalMapping.cc:300-361).  Fields as the arguments above. */
typedef struct {
  const orbx_keypoint* k1; const uint8_t* desc1; const int32_t* has_mp1; const int32_t* cam1; const float* uright1;
  int32_t n1; orbm_featvec fv1;
  const orbx_keypoint* k2; const uint8_t* desc2; const int32_t* has_mp2; const int32_t* cam2; const float* uright2;
  int32_t n2; orbm_featvec fv2;
  const float* F12s; const float* epipoles; const float* scale_factors2; const float* level_sigma2_2;
  int32_t* matches12; /* out */
  int32_t nmatches;   /* out */
} orbm_tri_pair;
int orbm_search_for_triangulation_batch_host(orbm_matcher* m, orbm_tri_pair* pairs, int n_pairs, int nlevels, int only_stereo,
                                             const int32_t* cam_enabled, int check_ori);

/* ---- streaming front end of a multi-camera rig -------------------------------------------------------------
 * Frame::Frame runs one ORBextractor per camera on every rig-frame (src/Frame.cc:148-346, :182-185) and
 * MonocularInitialization matches consecutive frames of camera 1 (src/Tracking.cc:870-871).  A pipeline takes the
 * frames of `rig_frames` rig-frames per step from HOST memory (pinned memory makes the copies asynchronous) and
 * leaves keypoints / descriptors / counts of every camera and the initialisation matches of camera 0 in pinned host
 * memory it owns; the copies of step k+1 overlap the kernels of step k (`depth` steps in flight; four CUDA streams:
 * copy-in, extraction, matching, copy-out).  orbp_submit returns at once; orbp_wait blocks for one step. */
#define ORBP_MAX_CAMS 8
typedef struct {
  int32_t n_cams;
  int32_t nfeatures[ORBP_MAX_CAMS]; /* per camera (src/Tracking.cc:144-145: the second camera runs with half) */
  float scale_factor;
  int32_t nlevels, ini_th_fast, min_th_fast;
  int32_t width, height;
  int32_t rig_frames; /* rig-frames per step */
  int32_t depth;      /* steps in flight, >= 2 (0 = default 3) */
  int32_t match;      /* != 0: SearchForInitialization(frame t, frame t+1) of camera 0, vbPrevMatched = F1's keypoints */
  int32_t window;     /* windowSize of that search (100 at src/Tracking.cc:871) */
  float nnratio;      /* ORBmatcher(nnratio, check_ori) */
  int32_t check_ori;
  int32_t device;     /* CUDA device ordinal, -1 = current */
} orbp_config;
typedef struct orbp_pipeline orbp_pipeline;
/* Pinned host memory owned by the pipeline, valid until `depth` further submits.  Camera c: kps[c] / desc[c] hold
 * rig_frames x cap[c] entries (frame f at f*cap[c]), counts[c][f] valid ones.  matches12: (rig_frames-1) x cap[0]
 * (vnMatches12 of pair (f, f+1), -1 = none), nmatches[f]; NULL when the pipeline was created with match = 0. */
typedef struct {
  const orbx_keypoint* kps[ORBP_MAX_CAMS];
  const uint8_t* desc[ORBP_MAX_CAMS];
  const int32_t* counts[ORBP_MAX_CAMS];
  int32_t cap[ORBP_MAX_CAMS];
  const int32_t* matches12;
  const int32_t* nmatches;
  int32_t rig_frames;
} orbp_result;
int orbp_create(const orbp_config* cfg, orbp_pipeline** out);
void orbp_destroy(orbp_pipeline* p);
const char* orbp_last_error(const orbp_pipeline* p);
int orbp_capacity(const orbp_pipeline* p, int cam);
/* images[c]: rig_frames frames of camera c in host memory, frame f at images[c] + f*frame_stride, `row_stride` bytes
 * per row.  Returns the step's ticket (>= 0) or an ORBX_E_* code.  The host buffers may be reused once the step's
 * copy-in has run, i.e. after orbp_wait of this ticket (or orbp_drain). */
long long orbp_submit(orbp_pipeline* p, const uint8_t* const* images, size_t frame_stride, size_t row_stride);
int orbp_wait(orbp_pipeline* p, long long ticket, orbp_result* out);
int orbp_drain(orbp_pipeline* p);
/* cudaStream_t of the pipeline: 0 copy-in, 1 extraction, 2 matching, 3 copy-out (to time or to order foreign work). */
void* orbp_stream(orbp_pipeline* p, int which);
long long orbp_launch_count(const orbp_pipeline* p);

/* ---- multi-GPU: the one collective of the path ------------------------------------------------------------
 * One process per GPU.  Camera streams are dealt over the ranks; every rank extracts its cameras and the
 * per-camera blocks (counts, keypoints, descriptors of a chunk of rig-frames) of ALL cameras are all-gathered so
 * that each rank can match its share of rig-frames across cameras: the multi-GPU form of
 * Frame::mDescriptors_total (src/Frame.cc:170,191-194), which the reference builds by concatenating the cameras'
 * descriptors on its single device.  The extractor writes straight into the rank's slot of the gather buffer
 * (orbx_extract_batch_device takes any device pointers), so the exchange is ONE in-place NCCL all-gather per
 * chunk, asynchronous on the caller's stream.  NCCL is bound at run time (dlopen "libnccl.so.2", or the path in
 * $ORB_NCCL_LIB); ORBX_E_STATE when it cannot be loaded. */
#define ORBD_UNIQUE_ID_BYTES 128
typedef struct orbd_comm orbd_comm;
/* ncclGetUniqueId: call on one rank, distribute the 128 bytes to the others by the host's own means. */
int orbd_get_unique_id(uint8_t* id128);
/* ncclCommInitRank on `device` (-1 = current).  Collective: every rank of the job must call it. */
int orbd_comm_create(int rank, int world, const uint8_t* id128, int device, orbd_comm** out);
void orbd_comm_destroy(orbd_comm* c);
int orbd_rank(const orbd_comm* c);
int orbd_world(const orbd_comm* c);
/* d_buf holds world * bytes_per_rank bytes; rank r's block is at d_buf + r*bytes_per_rank and must be complete on
 * `cuda_stream` order.  After the call (stream order) every rank holds all blocks.  world == 1: no-op. */
int orbd_allgather_inplace(orbd_comm* c, void* d_buf, size_t bytes_per_rank, void* cuda_stream);
int orbd_nccl_version(void);
const char* orbd_last_error(const orbd_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
