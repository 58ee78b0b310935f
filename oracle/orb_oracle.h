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

// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:3948-3989) on bin counts.
void om_three_maxima(const int* counts, int L, int* ind1, int* ind2, int* ind3);

#ifdef __cplusplus
}
#endif
