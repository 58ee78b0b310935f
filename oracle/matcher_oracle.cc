// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// Flat-array restatement of the reference's matcher hot path (src/ORBmatcher.cc, src/Frame.cc grid functions,
// src/MapPoint.cc, DBoW2 transform); each function cites its source lines.
// The reference ships no tests or golden vectors for these functions.  Pins: the reference's
// own src/ORBmatcher.cc compiled VERBATIM (oracle/_ref/libmatcher_ref.so, oracle/matcher_ref_capi.cc) returns the
// same results as this restatement on the seeded scenes of tests/test_matcher_ref.py (DescriptorDistance,
// SearchForInitialization, the four SearchByProjection overloads, SearchByBoW x2, SearchForTriangulation,
// Fuse x2, SearchBySim3); plus hand-checkable known answers and an independent pure-Python transliteration on
// small cases (tests/test_matcher_oracle.py); cv2 for the OpenCV primitives; the Frame glue against the
// reference's own src/Frame.cc compiled verbatim (oracle/_ref/libframe_ref.so, tests/test_frame_ref.py).
// om_compute_stereo_matches restates code the fork keeps COMMENTED OUT (src/Frame.cc:782-956): parity unpinned
// against a reference build; pinned by a known-answer test and a pure-Python transliteration only.
#include "orb_oracle.h"
#include "cvprim.h"

#include <climits>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <utility>
#include <vector>

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;  // src/ORBmatcher.cc:37-39
const int GRID_COLS = 64, GRID_ROWS = 48;                 // include/Frame.h FRAME_GRID_COLS/ROWS

// Frame::AssignFeaturesToGrid + PosInGrid, src/Frame.cc:348-395, 632-642
struct Grid {
  float min_x, min_y, inv_w, inv_h;
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  const float *kx, *ky;
  const int* koct;
  Grid(const float* x, const float* y, const int* oct, int n, om_bounds b) : kx(x), ky(y), koct(oct) {
    min_x = b.min_x;
    min_y = b.min_y;
    inv_w = (float)GRID_COLS / (b.max_x - b.min_x);   // src/Frame.cc:271-272
    inv_h = (float)GRID_ROWS / (b.max_y - b.min_y);
    for (int i = 0; i < n; ++i) {
      const int px = (int)std::round((x[i] - min_x) * inv_w);
      const int py = (int)std::round((y[i] - min_y) * inv_h);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      cell[px][py].push_back(i);
    }
  }
  // Frame::GetFeaturesInArea, src/Frame.cc:510-566
  void query(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    const int cx0 = std::max(0, (int)std::floor((x - min_x - r) * inv_w));
    if (cx0 >= GRID_COLS) return;
    const int cx1 = std::min(GRID_COLS - 1, (int)std::ceil((x - min_x + r) * inv_w));
    if (cx1 < 0) return;
    const int cy0 = std::max(0, (int)std::floor((y - min_y - r) * inv_h));
    if (cy0 >= GRID_ROWS) return;
    const int cy1 = std::min(GRID_ROWS - 1, (int)std::ceil((y - min_y + r) * inv_h));
    if (cy1 < 0) return;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = cx0; ix <= cx1; ++ix)
      for (int iy = cy0; iy <= cy1; ++iy)
        for (int idx : cell[ix][iy]) {
          if (check) {
            if (koct[idx] < minLevel) continue;
            if (maxLevel >= 0 && koct[idx] > maxLevel) continue;
          }
          const float dx = kx[idx] - x, dy = ky[idx] - y;
          if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(idx);
        }
  }
};

struct SoA {
  std::vector<float> x, y;
  std::vector<int> oct;
  SoA(const oo_keypoint* k, int n) : x(n), y(n), oct(n) {
    for (int i = 0; i < n; ++i) { x[i] = k[i].x; y[i] = k[i].y; oct[i] = k[i].octave; }
  }
};

}  // namespace

extern "C" {

// src/ORBmatcher.cc:3994-4010 — the SWAR popcount is reproduced literally.
int om_distance(const uint8_t* a, const uint8_t* b) {
  int32_t pa[8], pb[8];
  std::memcpy(pa, a, 32);
  std::memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    unsigned int v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// Best / second-best scan of src/ORBmatcher.cc:895-924 applied to all targets.
void om_bruteforce(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int th_dist,
                   int* out_idx, int* out_d1, int* out_d2) {
  for (int i = 0; i < nq; ++i) {
    int best = 256, best2 = 256, bi = -1;
    for (int j = 0; j < nt; ++j) {
      const int d = om_distance(q + (size_t)i * 32, t + (size_t)j * 32);
      if (d < best) { best2 = best; best = d; bi = j; }
      else if (d < best2) best2 = d;
    }
    out_d1[i] = best;
    out_d2[i] = best2;
    out_idx[i] = (bi >= 0 && best <= th_dist && (float)best < (float)best2 * ratio) ? bi : -1;
  }
}

int om_features_in_area(const float* kx, const float* ky, const int* koct, int n, om_bounds b,
                        float x, float y, float r, int min_level, int max_level, int* out, int cap) {
  Grid g(kx, ky, koct, n, b);
  std::vector<int> v;
  g.query(x, y, r, min_level, max_level, v);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

// src/ORBmatcher.cc:3948-3989
void om_three_maxima(const int* counts, int L, int* ind1, int* ind2, int* ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  int i1 = -1, i2 = -1, i3 = -1;
  for (int i = 0; i < L; ++i) {
    const int s = counts[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
    else if (s > max3) { max3 = s; i3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
  else if (max3 < 0.1f * (float)max1) { i3 = -1; }
  *ind1 = i1; *ind2 = i2; *ind3 = i3;
}

// src/ORBmatcher.cc:868-983
int om_search_for_initialization(const oo_keypoint* k1, const uint8_t* d1, int n1,
                                 const oo_keypoint* k2, const uint8_t* d2, int n2, om_bounds b2,
                                 float* prev_xy, int window, float nnratio, int check_ori,
                                 int* matches12) {
  int nmatches = 0;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  std::vector<int> matchedDist(n2, INT_MAX), matches21(n2, -1);
  SoA s2(k2, n2);
  Grid grid(s2.x.data(), s2.y.data(), s2.oct.data(), n2, b2);
  std::vector<int> cand;
  for (int i1 = 0; i1 < n1; ++i1) {
    const int level1 = k1[i1].octave;
    if (level1 > 0) continue;
    grid.query(prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window, level1, level1, cand);
    if (cand.empty()) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (int i2 : cand) {
      const int dist = om_distance(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
      if (matchedDist[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
        matches12[i1] = bestIdx2;
        matches21[bestIdx2] = i1;
        matchedDist[bestIdx2] = bestDist;
        nmatches++;
        if (check_ori) {
          float rot = k1[i1].angle - k2[bestIdx2].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(i1);
        }
      }
    }
  }
  if (check_ori) {
    int counts[HISTO_LENGTH], ind1, ind2, ind3;
    for (int i = 0; i < HISTO_LENGTH; ++i) counts[i] = (int)rotHist[i].size();
    om_three_maxima(counts, HISTO_LENGTH, &ind1, &ind2, &ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < n1; ++i1)
    if (matches12[i1] >= 0) {
      prev_xy[2 * i1] = k2[matches12[i1]].x;
      prev_xy[2 * i1 + 1] = k2[matches12[i1]].y;
    }
  return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:325-438), the arithmetic part (:381-424): for one
// map point's observed descriptors, all pair distances, the median of every row (element int(0.5*(N-1)) of the
// sorted row, the zero self-distance included) and the first row with the least median.  Batched over points:
// the descriptors of point p are rows offsets[p] .. offsets[p+1]-1; best_idx[p] is relative to offsets[p],
// -1 for an empty set (:369-370 returns early).
void om_compute_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx) {
  std::vector<int> row;
  for (int p = 0; p < n_points; ++p) {
    const int N = offsets[p + 1] - offsets[p];
    best_idx[p] = -1;
    if (N <= 0) continue;
    const uint8_t* d = desc + (size_t)offsets[p] * 32;
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; ++i) {
      row.assign(N, 0);
      for (int j = 0; j < N; ++j) row[j] = i == j ? 0 : om_distance(d + (size_t)i * 32, d + (size_t)j * 32);
      std::sort(row.begin(), row.end());
      const int median = row[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best_idx[p] = BestIdx;
  }
}

// ---- Frame glue between the extractor and the matchers (src/Frame.cc) -----------------------
// Frame::UndistortKeyPoints (src/Frame.cc:673-706; _cam2 :708-741 is the same on camera 2's arrays).
void om_undistort_keypoints(const oo_keypoint* k, int n, float fx, float fy, float cx, float cy, const float* dist5,
                            oo_keypoint* k_un) {
  for (int i = 0; i < n; ++i) k_un[i] = k[i];
  if (dist5[0] == 0.0) return;  // :675-679
  std::vector<float> mat(2 * (size_t)n), out(2 * (size_t)n);
  for (int i = 0; i < n; ++i) { mat[2 * i] = k[i].x; mat[2 * i + 1] = k[i].y; }
  cvp::undistort_points(mat.data(), n, fx, fy, cx, cy, dist5, out.data());  // cv::undistortPoints(mat,mat,mK,mDistCoef,Mat(),mK)
  for (int i = 0; i < n; ++i) { k_un[i].x = out[2 * i]; k_un[i].y = out[2 * i + 1]; }
}

// Frame::ComputeImageBounds (src/Frame.cc:743-779).
void om_compute_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float* dist5, om_bounds* b) {
  if (dist5[0] != 0.0) {
    const float mat[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
    float o[8];
    cvp::undistort_points(mat, 4, fx, fy, cx, cy, dist5, o);
    b->min_x = std::min(o[0], o[4]);
    b->max_x = std::max(o[2], o[6]);
    b->min_y = std::min(o[1], o[3]);
    b->max_y = std::max(o[5], o[7]);
  } else {
    b->min_x = 0.0f;
    b->max_x = (float)cols;
    b->min_y = 0.0f;
    b->max_y = (float)rows;
  }
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:959-985; _cam2 :987-1010): depth is the CV_32F depth image
// (already scaled by mDepthMapFactor), indexed by the DISTORTED keypoint truncated to int.
void om_compute_stereo_from_rgbd(const oo_keypoint* k, const oo_keypoint* k_un, int n, const float* depth, int cols,
                                 int rows, size_t stride_floats, float mbf, float* uright, float* depth_out) {
  (void)cols; (void)rows;
  for (int i = 0; i < n; ++i) {
    uright[i] = -1;
    depth_out[i] = -1;
    const float v = k[i].y, u = k[i].x;
    const float d = depth[(size_t)(int)v * stride_floats + (int)u];
    if (d > 0) {
      depth_out[i] = d;
      uright[i] = k_un[i].x - mbf / d;
    }
  }
}

// Frame::ComputeStereoMatches (src/Frame.cc:782-956, commented upstream code of this fork): row-band Hamming
// association of left and right keypoints, 11x11 SAD refinement on the pyramid level of the left keypoint with a
// parabola fit, then the 2.1 x median SAD outlier cut.  Rows outside the image are clamped when a right keypoint's
// band is listed (the reference would index vRowIndices out of range); the SAD sums are integers, so every
// accumulation order agrees.
void om_compute_stereo_matches(const oo_keypoint* kl, const uint8_t* dl, int nl, const oo_keypoint* kr, const uint8_t* dr,
                               int nr, const om_image* pyr_l, const om_image* pyr_r, int nlevels, const float* scale_factors,
                               const float* inv_scale_factors, float mbf, float mb, float* uright, float* depth) {
  (void)nlevels;
  for (int i = 0; i < nl; ++i) { uright[i] = -1.0f; depth[i] = -1.0f; }           // :784-785
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                                     // :787
  const int nRows = pyr_l[0].h;                                                     // :789
  std::vector<std::vector<int> > vRowIndices(nRows);                                // :792-810
  for (int iR = 0; iR < nr; ++iR) {
    const float kpY = kr[iR].y;
    const float r = 2.0f * scale_factors[kr[iR].octave];
    const int maxr = (int)std::ceil(kpY + r);
    const int minr = (int)std::floor(kpY - r);
    for (int yi = std::max(minr, 0); yi <= std::min(maxr, nRows - 1); ++yi) vRowIndices[yi].push_back(iR);
  }
  const float minZ = mb, minD = 0, maxD = mbf / minZ;                               // :813-815
  std::vector<std::pair<int, int> > vDistIdx;
  for (int iL = 0; iL < nl; ++iL) {                                                 // :821
    const int levelL = kl[iL].octave;
    const float vL = kl[iL].y, uL = kl[iL].x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cand = vRowIndices[row];                                // :828
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;                                                         // :839
    int bestIdxR = 0;
    for (size_t iC = 0; iC < cand.size(); ++iC) {                                   // :845-866
      const int iR = cand[iC];
      if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
      const float uR = kr[iR].x;
      if (uR >= minU && uR <= maxU) {
        const int dist = om_distance(dl + 32 * (size_t)iL, dr + 32 * (size_t)iR);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (!(bestDist < thOrbDist)) continue;                                          // :869
    const float uR0 = kr[bestIdxR].x;
    const float scaleFactor = inv_scale_factors[levelL];
    const float scaleduL = std::round(uL * scaleFactor);                            // :873-875 (roundf)
    const float scaledvL = std::round(vL * scaleFactor);
    const float scaleduR0 = std::round(uR0 * scaleFactor);
    const int w = 5, L = 5;
    const om_image& IL = pyr_l[levelL];
    const om_image& IR = pyr_r[levelL];
    const int yl = (int)scaledvL, xl = (int)scaleduL, xr = (int)scaleduR0;
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;             // :888-891
    if (iniu < 0 || endu >= IR.w) continue;
    auto PL = [&](int y, int x) { return (int)IL.data[(ptrdiff_t)y * (ptrdiff_t)IL.step + x]; };
    auto PR = [&](int y, int x) { return (int)IR.data[(ptrdiff_t)y * (ptrdiff_t)IR.step + x]; };
    const int cl = PL(yl, xl);
    int bestSad = INT_MAX, bestincR = 0;
    float vDists[2 * 5 + 1];
    for (int incR = -L; incR <= L; ++incR) {                                        // :893-908
      const int cr = PR(yl, xr + incR);
      int sad = 0;
      for (int dy = -w; dy <= w; ++dy)
        for (int dx = -w; dx <= w; ++dx) sad += std::abs((PL(yl + dy, xl + dx) - cl) - (PR(yl + dy, xr + incR + dx) - cr));
      const float dist = (float)sad;
      if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
      vDists[L + incR] = dist;
    }
    if (bestincR == -L || bestincR == L) continue;                                  // :910
    const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2)); // :918
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = scale_factors[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);   // :924
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {                                    // :928-938
      if (disparity <= 0) {
        disparity = 0.01;
        bestuR = uL - 0.01;   // float - double, rounded once
      }
      depth[iL] = mbf / disparity;
      uright[iL] = bestuR;
      vDistIdx.push_back(std::make_pair(bestSad, iL));
    }
  }
  if (vDistIdx.empty()) return;
  std::sort(vDistIdx.begin(), vDistIdx.end());                                      // :942-955
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
    if (vDistIdx[i].first < thDist) break;
    uright[vDistIdx[i].second] = -1;
    depth[vDistIdx[i].second] = -1;
  }
}

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:348-395, 632-642): CSR over cell = ix*48 + iy,
// items in insertion (keypoint index) order.  cell_start has 64*48 + 1 entries.
void om_assign_features_to_grid(const oo_keypoint* k_un, int n, om_bounds b, int32_t* cell_start, int32_t* items) {
  std::vector<float> x(n), y(n);
  std::vector<int> oct(n, 0);
  for (int i = 0; i < n; ++i) { x[i] = k_un[i].x; y[i] = k_un[i].y; }
  Grid g(x.data(), y.data(), oct.data(), n, b);
  int run = 0;
  for (int ix = 0; ix < GRID_COLS; ++ix)
    for (int iy = 0; iy < GRID_ROWS; ++iy) {
      cell_start[ix * GRID_ROWS + iy] = run;
      for (int idx : g.cell[ix][iy]) items[run++] = idx;
    }
  cell_start[GRID_COLS * GRID_ROWS] = run;
}

// ORBmatcher::SearchByBoW — the four variants share one loop:
//   (KeyFrame*, Frame&, vpMapPointMatches)            src/ORBmatcher.cc:206-388   (all cameras)
//   SearchByBoW_cam1(KeyFrame*, Frame&, ...)           src/ORBmatcher.cc:390-565   (indices < N only)
//   (KeyFrame*, KeyFrame*, vpMatches12)                src/ORBmatcher.cc:996-1163  (all cameras)
//   SearchByBoW_cam1(KeyFrame*, KeyFrame*, ...)        src/ORBmatcher.cc:1180-1363
// Walk of the two DBoW2 feature vectors (std::map<node id, vector<feature index>>, flattened here
// to CSR with ascending node ids) with the map's lower_bound jumps (:350-359); inside a common
// node every valid feature of side 1 scans the node's side-2 features that are valid and not yet
// matched, strict-< best / second best starting from 256 (:286-321), accepted when
// best <= max_dist (TH_LOW for the Frame variants :324,487, TH_LOW-1 for the KeyFrame variants'
// `bestDist1<TH_LOW` :1107,1292) and best < ratio * second (:328); then the rotation histogram.
int om_search_by_bow(const uint8_t* d1, const float* angle1, const int32_t* valid1, int n1,
                     const int32_t* node1, const int32_t* start1, const int32_t* items1, int nn1,
                     const uint8_t* d2, const float* angle2, const int32_t* valid2, int n2,
                     const int32_t* node2, const int32_t* start2, const int32_t* items2, int nn2,
                     float nnratio, int check_ori, int max_dist, int32_t* matches12, int32_t* matches21) {
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  for (int i = 0; i < n2; ++i) matches21[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  int a = 0, b = 0;
  while (a < nn1 && b < nn2) {
    if (node1[a] == node2[b]) {
      for (int p = start1[a]; p < start1[a + 1]; ++p) {
        const int idx1 = items1[p];
        if (valid1 && !valid1[idx1]) continue;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int q = start2[b]; q < start2[b + 1]; ++q) {
          const int idx2 = items2[q];
          if (matches21[idx2] >= 0) continue;
          if (valid2 && !valid2[idx2]) continue;
          const int dist = om_distance(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= max_dist && (float)bestDist1 < nnratio * (float)bestDist2) {
          matches12[idx1] = bestIdx2;
          matches21[bestIdx2] = idx1;
          if (check_ori) {
            float rot = angle1[idx1] - angle2[bestIdx2];
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(idx1);
          }
          nmatches++;
        }
      }
      ++a;
      ++b;
    } else if (node1[a] < node2[b]) {
      a = (int)(std::lower_bound(node1 + a, node1 + nn1, node2[b]) - node1);
    } else {
      b = (int)(std::lower_bound(node2 + b, node2 + nn2, node1[a]) - node2);
    }
  }
  if (check_ori) {
    int counts[HISTO_LENGTH], i1, i2, i3;
    for (int i = 0; i < HISTO_LENGTH; ++i) counts[i] = (int)rotHist[i].size();
    om_three_maxima(counts, HISTO_LENGTH, &i1, &i2, &i3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != i1 && i != i2 && i != i3)
        for (int idx1 : rotHist[i]) {
          matches21[matches12[idx1]] = -1;
          matches12[idx1] = -1;
          nmatches--;
        }
  }
  return nmatches;
}

// ORBmatcher::CheckDistEpipolarLine (src/ORBmatcher.cc:167-184); F12 row-major 3x3.
static bool check_dist_epipolar_line(const oo_keypoint& kp1, const oo_keypoint& kp2, const float* F12,
                                     const float* level_sigma2) {
  const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
  const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
  const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
  const float num = a * kp2.x + b * kp2.y + c;
  const float den = a * a + b * b;
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * level_sigma2[kp2.octave];
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1364-1720), from the feature-vector walk
// on (:1433-1700).  The fundamental matrices F12s[cam] (:1421-1423) and the epipoles (ex, ey) of
// both cameras (:1441-1449) are inputs: they are cv::Mat algebra on the key-frame poses that the
// caller evaluates with its own OpenCV (the reference's F12 argument, which it recomputes).
// Note vbMatched2 is never set in the reference (:1452 only declares it), so key-frame-2 features
// can be matched by several key-frame-1 features.
int om_search_for_triangulation(const oo_keypoint* k1, const uint8_t* d1, const int32_t* has_mp1, const int32_t* cam1,
                                const float* uright1, int n1, const int32_t* node1, const int32_t* start1,
                                const int32_t* items1, int nn1, const oo_keypoint* k2, const uint8_t* d2,
                                const int32_t* has_mp2, const int32_t* cam2, const float* uright2, int n2,
                                const int32_t* node2, const int32_t* start2, const int32_t* items2, int nn2,
                                const float* F12s, const float* epipoles, const float* scale_factors2,
                                const float* level_sigma2_2, int only_stereo, const int32_t* cam_enabled, int check_ori,
                                int32_t* matches12) {
  (void)n2;
  int nmatches = 0;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  int a = 0, b = 0;
  while (a < nn1 && b < nn2) {
    if (node1[a] == node2[b]) {
      for (int p = start1[a]; p < start1[a + 1]; ++p) {
        const int idx1 = items1[p];
        if (has_mp1[idx1]) continue;
        const int camIdx1 = cam1[idx1];
        if (!cam_enabled[camIdx1]) continue;
        const bool bStereo1 = uright1[idx1] >= 0;
        if (only_stereo && !bStereo1) continue;
        const oo_keypoint& kp1 = k1[idx1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int q = start2[b]; q < start2[b + 1]; ++q) {
          const int idx2 = items2[q];
          if (has_mp2[idx2]) continue;
          const int camIdx2 = cam2[idx2];
          if (camIdx1 != camIdx2) continue;
          const bool bStereo2 = uright2[idx2] >= 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = om_distance(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > bestDist) continue;
          const oo_keypoint& kp2 = k2[idx2];
          if (!bStereo1 && !bStereo2) {
            const float distex = epipoles[2 * camIdx2] - kp2.x;
            const float distey = epipoles[2 * camIdx2 + 1] - kp2.y;
            if (distex * distex + distey * distey < 100 * scale_factors2[kp2.octave]) continue;
          }
          if (check_dist_epipolar_line(kp1, kp2, F12s + 9 * camIdx1, level_sigma2_2)) {
            bestIdx2 = idx2;
            bestDist = dist;
          }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          nmatches++;
          if (check_ori) {
            float rot = kp1.angle - k2[bestIdx2].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(idx1);
          }
        }
      }
      ++a;
      ++b;
    } else if (node1[a] < node2[b]) {
      a = (int)(std::lower_bound(node1 + a, node1 + nn1, node2[b]) - node1);
    } else {
      b = (int)(std::lower_bound(node2 + b, node2 + nn2, node1[a]) - node2);
    }
  }
  if (check_ori) {
    int counts[HISTO_LENGTH], i1, i2, i3;
    for (int i = 0; i < HISTO_LENGTH; ++i) counts[i] = (int)rotHist[i].size();
    om_three_maxima(counts, HISTO_LENGTH, &i1, &i2, &i3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != i1 && i != i2 && i != i3)
        for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
  }
  return nmatches;
}

// src/ORBmatcher.cc:62-157
int om_search_by_projection_points(const oo_keypoint* k, const uint8_t* d, const float* u_right,
                                   int n, om_bounds b, const float* scale_factors, int nlevels,
                                   const om_mappoint* mp, const uint8_t* mp_desc, const int* mp_obs,
                                   int nmp, float th, float nnratio, int* frame_mp,
                                   const int* frame_mp_obs) {
  (void)nlevels;
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  SoA s(k, n);
  Grid grid(s.x.data(), s.y.data(), s.oct.data(), n, b);
  // held[i]: does keypoint i currently hold a map point with Observations()>0 ?
  std::vector<char> held(n, 0);
  for (int i = 0; i < n; ++i) held[i] = (frame_mp[i] >= 0 && frame_mp_obs && frame_mp_obs[i] > 0);
  std::vector<int> cand;
  for (int i = 0; i < nmp; ++i) {
    const om_mappoint& p = mp[i];
    if (!p.track_in_view) continue;
    if (p.bad) continue;
    const int lvl = p.level;
    float r = p.view_cos > 0.998 ? 2.5f : 4.0f;       // RadiusByViewingCos :151-157
    if (bFactor) r *= th;
    grid.query(p.proj_x, p.proj_y, r * scale_factors[lvl], lvl - 1, lvl, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (held[idx]) continue;
      if (u_right && u_right[idx] > 0) {
        const float er = std::fabs(p.proj_xr - u_right[idx]);
        if (er > r * scale_factors[lvl]) continue;
      }
      const int dist = om_distance(mp_desc + (size_t)i * 32, d + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist;
        bestLevel2 = bestLevel; bestLevel = k[idx].octave;
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = k[idx].octave;
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      frame_mp[bestIdx] = i;
      held[bestIdx] = mp_obs ? (mp_obs[i] > 0) : 1;
      nmatches++;
    }
  }
  return nmatches;
}

// ---- pose-based overloads ------------------------------------------------------------------
// cv::Mat float algebra as OpenCV 4.13 evaluates it for these tiny matrices (probed against
// cv2.gemm, tools/make_golden.py header): a product row is accumulated in float32 left to right
// starting from 0, then the addend is added; cv::norm accumulates squares in double.
namespace {
inline void mat3_mul_vec_add(const float* R /*3x3 row-major, row stride rs*/, int rs, const float* x, const float* t,
                             float alpha, float* out) {
  for (int i = 0; i < 3; ++i) {
    float s = 0.f;
    for (int k = 0; k < 3; ++k) s += R[i * rs + k] * x[k];
    s *= alpha;
    out[i] = t ? s + t[i] : s;
  }
}
inline void mat3t_mul_vec(const float* R, int rs, const float* x, float alpha, float* out) {  // alpha * R^T x
  // cv::gemm with a transposed operand leaves the small-matrix float path: products and sums in double,
  // alpha applied in double, one rounding to float (probed against cv2.gemm(..., GEMM_1_T))
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += (double)R[k * rs + i] * (double)x[k];
    out[i] = (float)(s * (double)alpha);
  }
}
struct Grid2 {  // mGrids[c]: per-camera grids over the concatenated keypoints (src/Frame.cc:384-393)
  Grid g0, g1;
  Grid2(const float* x, const float* y, const int* oct, const int* cam, int n, om_bounds b)
      : g0(x, y, oct, 0, b), g1(x, y, oct, 0, b) {
    g0.kx = g1.kx = x; g0.ky = g1.ky = y; g0.koct = g1.koct = oct;
    for (int i = 0; i < n; ++i) {
      Grid& g = (cam && cam[i] == 1) ? g1 : g0;
      const int px = (int)std::round((x[i] - g.min_x) * g.inv_w);
      const int py = (int)std::round((y[i] - g.min_y) * g.inv_h);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      g.cell[px][py].push_back(i);
    }
  }
};
}  // namespace

// src/ORBmatcher.cc:3448-3641
int om_search_by_projection_frame(const oo_keypoint* cur_k, const uint8_t* cur_desc, const float* cur_uright,
                                  const int32_t* cur_cam, int n_cur, om_bounds b, const float* scale_factors,
                                  int nlevels, om_camera cam, const float* Tcw_cur, const float* Tcw_last,
                                  const oo_keypoint* last_k, const int32_t* last_cam, const int32_t* last_valid,
                                  const float* last_xyz, const uint8_t* last_desc, const int32_t* last_obs,
                                  int n_last, const float* calib, float th, int mono, int check_ori,
                                  int32_t* cur_mp, const int32_t* cur_mp_obs) {
  (void)nlevels;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  // mRcam21 = Rcam12.t(); mtcam21 = -mRcam21 * tcam12  (:3464-3471)
  float Rcam21[9], tcam21[3];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) Rcam21[i * 3 + k] = calib[k * 3 + i];
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  float tcw[3] = {Tcw_cur[3], Tcw_cur[7], Tcw_cur[11]}, tlw[3] = {Tcw_last[3], Tcw_last[7], Tcw_last[11]};
  float twc[3], tlc[3];
  mat3t_mul_vec(Tcw_cur, 4, tcw, -1.f, twc);             // twc = -Rcw.t()*tcw
  mat3_mul_vec_add(Tcw_last, 4, twc, tlw, 1.f, tlc);     // tlc = Rlw*twc+tlw
  const bool fwd[2] = {tlc[2] > cam.mb && !mono, tlc[0] > cam.mb && !mono};
  const bool bwd[2] = {-tlc[2] > cam.mb && !mono, -tlc[0] > cam.mb && !mono};
  SoA s(cur_k, n_cur);
  Grid2 grids(s.x.data(), s.y.data(), s.oct.data(), cur_cam, n_cur, b);
  std::vector<char> held(n_cur, 0);
  for (int i = 0; i < n_cur; ++i) held[i] = cur_mp[i] >= 0 && cur_mp_obs && cur_mp_obs[i] > 0;
  std::vector<int> cand;
  for (int i = 0; i < n_last; ++i) {
    if (!last_valid[i]) continue;
    const int c = last_cam ? last_cam[i] : 0;
    float x3Dc[3];
    mat3_mul_vec_add(Tcw_cur, 4, last_xyz + 3 * i, tcw, 1.f, x3Dc);  // Rcw*x3Dw+tcw
    if (c == 1) {
      float tmp[3];
      mat3_mul_vec_add(Rcam21, 3, x3Dc, tcam21, 1.f, tmp);
      x3Dc[0] = tmp[0]; x3Dc[1] = tmp[1]; x3Dc[2] = tmp[2];
    }
    const float xc = x3Dc[0], yc = x3Dc[1];
    const float invzc = (float)(1.0 / x3Dc[2]);
    if (invzc < 0) continue;
    const float u = cam.fx * xc * invzc + cam.cx;
    const float v = cam.fy * yc * invzc + cam.cy;
    if (u < b.min_x || u > b.max_x) continue;
    if (v < b.min_y || v > b.max_y) continue;
    const int nLastOctave = last_k[i].octave;
    const float radius = th * scale_factors[nLastOctave];
    const Grid& g = c == 1 ? grids.g1 : grids.g0;
    if (fwd[c]) g.query(u, v, radius, nLastOctave, -1, cand);           // default maxLevel = -1
    else if (bwd[c]) g.query(u, v, radius, 0, nLastOctave, cand);
    else g.query(u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (held[i2]) continue;
      if (cur_uright && cur_uright[i2] > 0) {
        const float ur = u - cam.mbf * invzc;
        const float er = std::fabs(ur - cur_uright[i2]);
        if (er > radius) continue;
      }
      const int dist = om_distance(last_desc + (size_t)i * 32, cur_desc + (size_t)i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      cur_mp[bestIdx2] = i;
      held[bestIdx2] = last_obs ? (last_obs[i] > 0) : 1;
      nmatches++;
      if (check_ori) {
        float rot = last_k[i].angle - cur_k[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (check_ori) {
    int counts[HISTO_LENGTH], i1, i2, i3;
    for (int i = 0; i < HISTO_LENGTH; ++i) counts[i] = (int)rotHist[i].size();
    om_three_maxima(counts, HISTO_LENGTH, &i1, &i2, &i3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != i1 && i != i2 && i != i3)
        for (int idx : rotHist[i]) { cur_mp[idx] = -1; nmatches--; }
  }
  return nmatches;
}

// src/ORBmatcher.cc:3809-3937 (+ MapPoint::PredictScale, src/MapPoint.cc:602-617)
int om_search_by_projection_keyframe(const oo_keypoint* cur_k, const uint8_t* cur_desc, int n_cur, om_bounds b,
                                     const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam,
                                     const float* Tcw_cur, const int32_t* kf_valid, const float* kf_xyz,
                                     const float* kf_max_dist, const float* kf_min_dist, const float* kf_max_d,
                                     const float* kf_angle, const uint8_t* kf_desc, int n_kf, float th, int orb_dist,
                                     int check_ori, int32_t* cur_mp) {
  int nmatches = 0;
  float tcw[3] = {Tcw_cur[3], Tcw_cur[7], Tcw_cur[11]}, Ow[3];
  mat3t_mul_vec(Tcw_cur, 4, tcw, -1.f, Ow);  // Ow = -Rcw.t()*tcw
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  SoA s(cur_k, n_cur);
  Grid grid(s.x.data(), s.y.data(), s.oct.data(), n_cur, b);
  std::vector<int> cand;
  for (int i = 0; i < n_kf; ++i) {
    if (!kf_valid[i]) continue;
    const float* x3Dw = kf_xyz + 3 * i;
    float x3Dc[3];
    mat3_mul_vec_add(Tcw_cur, 4, x3Dw, tcw, 1.f, x3Dc);
    const float xc = x3Dc[0], yc = x3Dc[1];
    const float invzc = (float)(1.0 / x3Dc[2]);
    const float u = cam.fx * xc * invzc + cam.cx;
    const float v = cam.fy * yc * invzc + cam.cy;
    if (u < b.min_x || u > b.max_x) continue;
    if (v < b.min_y || v > b.max_y) continue;
    double n2 = 0;
    for (int k = 0; k < 3; ++k) { const float d = x3Dw[k] - Ow[k]; n2 += (double)d * (double)d; }
    const float dist3D = (float)std::sqrt(n2);
    if (dist3D < kf_min_dist[i] || dist3D > kf_max_dist[i]) continue;
    const float ratio = kf_max_d[i] / dist3D;
    int nPredictedLevel = (int)std::ceil(std::log(ratio) / log_scale_factor);
    if (nPredictedLevel < 0) nPredictedLevel = 0;
    else if (nPredictedLevel >= nlevels) nPredictedLevel = nlevels - 1;
    const float radius = th * scale_factors[nPredictedLevel];
    grid.query(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (cur_mp[i2] >= 0) continue;
      const int dist = om_distance(kf_desc + (size_t)i * 32, cur_desc + (size_t)i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= orb_dist) {
      cur_mp[bestIdx2] = i;
      nmatches++;
      if (check_ori) {
        float rot = kf_angle[i] - cur_k[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (check_ori) {
    int counts[HISTO_LENGTH], i1, i2, i3;
    for (int i = 0; i < HISTO_LENGTH; ++i) counts[i] = (int)rotHist[i].size();
    om_three_maxima(counts, HISTO_LENGTH, &i1, &i2, &i3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != i1 && i != i2 && i != i3)
        for (int idx : rotHist[i]) { cur_mp[idx] = -1; nmatches--; }
  }
  return nmatches;
}

// src/ORBmatcher.cc:566-752.  cv::Mat::inv() of the 3x3 float matrix = OpenCV's closed form in
// double (pinned against cv2.invert); Mat::dot and cv::norm accumulate in double; A/s is taken as
// A * (float)(1/s) (OpenCV's scaled convertTo) — the last three are NOT pinnable from Python.
int om_search_by_projection_sim3(const oo_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf,
                                 om_bounds b, const float* scale_factors, int nlevels, float log_scale_factor,
                                 om_camera cam, const float* Scw, const float* calib, const int32_t* mp_valid,
                                 const float* mp_xyz, const float* mp_normal, const float* mp_max_dist,
                                 const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp,
                                 int th, int32_t* matched) {
  float Rcam21[9], tcam21[3];
  {
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = calib[i];
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1. / d;
    Rcam21[0] = (float)((S[4] * S[8] - S[5] * S[7]) * d); Rcam21[1] = (float)((S[2] * S[7] - S[1] * S[8]) * d);
    Rcam21[2] = (float)((S[1] * S[5] - S[2] * S[4]) * d); Rcam21[3] = (float)((S[5] * S[6] - S[3] * S[8]) * d);
    Rcam21[4] = (float)((S[0] * S[8] - S[2] * S[6]) * d); Rcam21[5] = (float)((S[2] * S[3] - S[0] * S[5]) * d);
    Rcam21[6] = (float)((S[3] * S[7] - S[4] * S[6]) * d); Rcam21[7] = (float)((S[1] * S[6] - S[0] * S[7]) * d);
    Rcam21[8] = (float)((S[0] * S[4] - S[1] * S[3]) * d);
  }
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  double ss = 0;
  for (int k = 0; k < 3; ++k) ss += (double)Scw[k] * (double)Scw[k];
  const float scw = (float)std::sqrt(ss);
  const float inv_s = (float)(1.0 / scw);
  float Rcw[16] = {0}, tcw[3], Ow[3];
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) Rcw[i * 4 + k] = Scw[i * 4 + k] * inv_s;
    tcw[i] = Scw[i * 4 + 3] * inv_s;
  }
  mat3t_mul_vec(Rcw, 4, tcw, -1.f, Ow);
  SoA s(kf_k, n_kf);
  Grid2 grids(s.x.data(), s.y.data(), s.oct.data(), kf_cam, n_kf, b);
  int nmatches = 0;
  std::vector<int> cand;
  for (int i = 0; i < n_mp; ++i) {
    if (!mp_valid[i]) continue;
    const float* p3Dw = mp_xyz + 3 * i;
    int bestDist = 256, bestIdxs = -1;
    for (int camidx = 0; camidx < 2; ++camidx) {
      float p3Dc[3];
      mat3_mul_vec_add(Rcw, 4, p3Dw, tcw, 1.f, p3Dc);
      if (camidx == 1) {
        float tmp[3];
        mat3_mul_vec_add(Rcam21, 3, p3Dc, tcam21, 1.f, tmp);
        p3Dc[0] = tmp[0]; p3Dc[1] = tmp[1]; p3Dc[2] = tmp[2];
      }
      if (p3Dc[2] < 0.0) continue;
      const float invz = 1 / p3Dc[2];
      const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
      const float u = cam.fx * x + cam.cx, v = cam.fy * y + cam.cy;
      if (!(u >= b.min_x && u < b.max_x && v >= b.min_y && v < b.max_y)) continue;  // KeyFrame::IsInImage
      float PO[3];
      double n2 = 0, dotn = 0;
      for (int k = 0; k < 3; ++k) {
        PO[k] = p3Dw[k] - Ow[k];
        n2 += (double)PO[k] * (double)PO[k];
        dotn += (double)PO[k] * (double)mp_normal[3 * i + k];
      }
      const float dist = (float)std::sqrt(n2);
      if (dist < mp_min_dist[i] || dist > mp_max_dist[i]) continue;
      if (dotn < 0.5 * dist) continue;
      const float ratio = mp_max_d[i] / dist;
      int lvl = (int)std::ceil(std::log(ratio) / log_scale_factor);
      if (lvl < 0) lvl = 0;
      else if (lvl >= nlevels) lvl = nlevels - 1;
      const float radius = th * scale_factors[lvl];
      const Grid& g = camidx == 1 ? grids.g1 : grids.g0;
      g.query(u, v, radius, -1, -1, cand);
      for (int idx : cand) {
        if (matched[idx] >= 0) continue;
        const int kpLevel = kf_k[idx].octave;
        if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
        const int dd = om_distance(mp_desc + (size_t)i * 32, kf_desc + (size_t)idx * 32);
        if (dd < bestDist) { bestDist = dd; bestIdxs = idx; }
      }
    }
    if (bestDist <= TH_LOW) { matched[bestIdxs] = i; nmatches++; }
  }
  return nmatches;
}

// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, CalibMatrix, th) (src/ORBmatcher.cc:1986-2190),
// the search part: for every map point and both cameras of the key frame, the key-frame feature it fuses with
// (best_idx[2*i + cam], -1 = none).  The side effects (:2160-2186: Replace / AddObservation / AddMapPoint, in map-point
// order, camera 0 before camera 1) do not feed back into the search and stay with the caller; nFused = number of
// entries >= 0.
//  key frame: concatenated keypoints (mvKeysUn_total), descriptors per global index, uright (mvuRight_total), kf_cam
//    (keypoint_to_cam), bounds (mnMinX..), mvScaleFactors, mvInvLevelSigma2, mfLogScaleFactor, fx/fy/cx/cy/mbf,
//    Tcw (4x4 row-major), Ow = {GetCameraCenter(), GetCameraCenter_cam2()} (6 floats), calib 4x3;
//  map points: mp_valid[i] = pMP && !isBad() && !IsInKeyFrame(pKF); world position, normal, distance invariance
//    limits, mfMaxDistance, descriptor.
int om_fuse(const oo_keypoint* kf_k, const uint8_t* kf_desc, const float* kf_uright, const int32_t* kf_cam, int n_kf,
            om_bounds b, const float* scale_factors, const float* inv_level_sigma2, int nlevels, float log_scale_factor,
            om_camera cam, const float* Tcw, const float* Ow, const float* calib, const int32_t* mp_valid,
            const float* mp_xyz, const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist,
            const float* mp_max_d, const uint8_t* mp_desc, int n_mp, float th, int32_t* best_idx) {
  // Rcam21 = Rcam12.inv() (3x3 CV_32F: closed form in double), tcam21 = -Rcam21 * tcam12   (:1996-2004)
  float Rcam21[9], tcam21[3];
  {
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = calib[i];
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1. / d;
    Rcam21[0] = (float)((S[4] * S[8] - S[5] * S[7]) * d); Rcam21[1] = (float)((S[2] * S[7] - S[1] * S[8]) * d);
    Rcam21[2] = (float)((S[1] * S[5] - S[2] * S[4]) * d); Rcam21[3] = (float)((S[5] * S[6] - S[3] * S[8]) * d);
    Rcam21[4] = (float)((S[0] * S[8] - S[2] * S[6]) * d); Rcam21[5] = (float)((S[2] * S[3] - S[0] * S[5]) * d);
    Rcam21[6] = (float)((S[3] * S[7] - S[4] * S[6]) * d); Rcam21[7] = (float)((S[1] * S[6] - S[0] * S[7]) * d);
    Rcam21[8] = (float)((S[0] * S[4] - S[1] * S[3]) * d);
  }
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  const float tcw[3] = {Tcw[3], Tcw[7], Tcw[11]};
  // camera 2 (:2031): p3Dc = Rcam21*Rcw*p3Dw + Rcam21*tcw + tcam21 = ((M*p3Dw) + (Rcam21*tcw)) + tcam21 with the 3x3
  // product M = Rcam21*Rcw and both matrix-vector products materialised (cv::MatExpr), all float32 left to right
  float M[9], Rt[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float acc = 0.f;
      for (int k = 0; k < 3; ++k) acc += Rcam21[i * 3 + k] * Tcw[k * 4 + j];
      M[i * 3 + j] = acc;
    }
  mat3_mul_vec_add(Rcam21, 3, tcw, nullptr, 1.f, Rt);
  SoA s(kf_k, n_kf);
  Grid2 grids(s.x.data(), s.y.data(), s.oct.data(), kf_cam, n_kf, b);
  int nFused = 0;
  std::vector<int> cand;
  for (int i = 0; i < n_mp; ++i) {
    best_idx[2 * i] = best_idx[2 * i + 1] = -1;
    if (!mp_valid[i]) continue;
    const float* p3Dw = mp_xyz + 3 * i;
    for (int c = 0; c < 2; ++c) {
      float p3Dc[3];
      if (c == 0) {
        mat3_mul_vec_add(Tcw, 4, p3Dw, tcw, 1.f, p3Dc);
      } else {
        float m1[3];
        mat3_mul_vec_add(M, 3, p3Dw, nullptr, 1.f, m1);
        for (int k = 0; k < 3; ++k) p3Dc[k] = (m1[k] + Rt[k]) + tcam21[k];
      }
      if (p3Dc[2] < 0.0f) continue;
      const float invz = 1 / p3Dc[2];
      const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
      const float u = cam.fx * x + cam.cx, v = cam.fy * y + cam.cy;
      if (!(u >= b.min_x && u < b.max_x && v >= b.min_y && v < b.max_y)) continue;  // KeyFrame::IsInImage
      const float ur = u - cam.mbf * invz;
      float PO[3];
      double n2 = 0, dotn = 0;
      for (int k = 0; k < 3; ++k) {
        PO[k] = p3Dw[k] - Ow[3 * c + k];
        n2 += (double)PO[k] * (double)PO[k];
        dotn += (double)PO[k] * (double)mp_normal[3 * i + k];
      }
      const float dist3D = (float)std::sqrt(n2);  // cv::norm
      if (dist3D < mp_min_dist[i] || dist3D > mp_max_dist[i]) continue;
      if (dotn < 0.5 * dist3D) continue;
      const float ratio = mp_max_d[i] / dist3D;  // MapPoint::PredictScale (src/MapPoint.cc:584-600)
      int lvl = (int)std::ceil(std::log(ratio) / log_scale_factor);
      if (lvl < 0) lvl = 0;
      else if (lvl >= nlevels) lvl = nlevels - 1;
      const float radius = th * scale_factors[lvl];
      const Grid& g = c == 1 ? grids.g1 : grids.g0;
      g.query(u, v, radius, -1, -1, cand);
      int bestDist = 256, bestIdx = -1;
      for (int idx : cand) {
        const oo_keypoint& kp = kf_k[idx];
        const int kpLevel = kp.octave;
        if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
        if (kf_uright[idx] >= 0) {
          const float ex = u - kp.x, ey = v - kp.y, er = ur - kf_uright[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
        } else {
          const float ex = u - kp.x, ey = v - kp.y;
          const float e2 = ex * ex + ey * ey;
          if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
        }
        const int dd = om_distance(mp_desc + (size_t)i * 32, kf_desc + (size_t)idx * 32);
        if (dd < bestDist) { bestDist = dd; bestIdx = idx; }
      }
      if (bestIdx < 0) continue;
      if (bestDist <= TH_LOW) { best_idx[2 * i + c] = bestIdx; nFused++; }
    }
  }
  return nFused;
}

// ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, vpPoints, vLoopMPCams, th, vpReplacePoint, CalibMatrix)
// (src/ORBmatcher.cc:2211-2441, called from LoopClosing::SearchAndFuse, src/LoopClosing.cc:841), the search part:
// best_idx[2*i + cam] as for om_fuse.  Differences from the pose-based overload: the pose is the Sim3 with its
// scale divided out (:2229-2233), one camera centre Ow = -Rcw^T*tcw with camera 2's at Ow + Rcw^T*tcam12
// (:2277-2279), no reprojection gate, bestDist starts at INT_MAX (:2334).  mp_valid[i] = !isBad() &&
// !spAlreadyFound.count(pMP).
int om_fuse_sim3(const oo_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf, om_bounds b,
                 const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam, const float* Scw,
                 const float* calib, const int32_t* mp_valid, const float* mp_xyz, const float* mp_normal,
                 const float* mp_max_dist, const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp,
                 float th, int32_t* best_idx) {
  float Rcam21[9], tcam21[3];
  {
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = calib[i];
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1. / d;
    Rcam21[0] = (float)((S[4] * S[8] - S[5] * S[7]) * d); Rcam21[1] = (float)((S[2] * S[7] - S[1] * S[8]) * d);
    Rcam21[2] = (float)((S[1] * S[5] - S[2] * S[4]) * d); Rcam21[3] = (float)((S[5] * S[6] - S[3] * S[8]) * d);
    Rcam21[4] = (float)((S[0] * S[8] - S[2] * S[6]) * d); Rcam21[5] = (float)((S[2] * S[3] - S[0] * S[5]) * d);
    Rcam21[6] = (float)((S[3] * S[7] - S[4] * S[6]) * d); Rcam21[7] = (float)((S[1] * S[6] - S[0] * S[7]) * d);
    Rcam21[8] = (float)((S[0] * S[4] - S[1] * S[3]) * d);
  }
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  // scw = sqrt(sRcw.row(0).dot(sRcw.row(0))) (double dot), Rcw = sRcw/scw, tcw = t/scw: cv::Mat / scalar multiplies
  // by the reciprocal taken in double and rounded to float
  double ss = 0;
  for (int k = 0; k < 3; ++k) ss += (double)Scw[k] * (double)Scw[k];
  const float scw = (float)std::sqrt(ss);
  const float inv_s = (float)(1.0 / scw);
  float Rcw[16] = {0}, tcw[3], Ow[3];
  double RtT12[3];  // Rcw.t() * tcam12, kept in double: `PO - Rcw.t()*tcam12` is ONE gemm(alpha=-1, C=PO, beta=1, GEMM_1_T)
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) Rcw[i * 4 + k] = Scw[i * 4 + k] * inv_s;
    tcw[i] = Scw[i * 4 + 3] * inv_s;
  }
  mat3t_mul_vec(Rcw, 4, tcw, -1.f, Ow);
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    for (int k = 0; k < 3; ++k) acc += (double)Rcw[k * 4 + i] * (double)calib[9 + k];
    RtT12[i] = acc;
  }
  float M[9], Rt[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float acc = 0.f;
      for (int k = 0; k < 3; ++k) acc += Rcam21[i * 3 + k] * Rcw[k * 4 + j];
      M[i * 3 + j] = acc;
    }
  mat3_mul_vec_add(Rcam21, 3, tcw, nullptr, 1.f, Rt);
  SoA s(kf_k, n_kf);
  Grid2 grids(s.x.data(), s.y.data(), s.oct.data(), kf_cam, n_kf, b);
  int nFused = 0;
  std::vector<int> cand;
  for (int i = 0; i < n_mp; ++i) {
    best_idx[2 * i] = best_idx[2 * i + 1] = -1;
    if (!mp_valid[i]) continue;
    const float* p3Dw = mp_xyz + 3 * i;
    for (int c = 0; c < 2; ++c) {
      float p3Dc[3];
      if (c == 0) {
        mat3_mul_vec_add(Rcw, 4, p3Dw, tcw, 1.f, p3Dc);
      } else {
        float m1[3];
        mat3_mul_vec_add(M, 3, p3Dw, nullptr, 1.f, m1);
        for (int k = 0; k < 3; ++k) p3Dc[k] = (m1[k] + Rt[k]) + tcam21[k];
      }
      if (p3Dc[2] < 0.0f) continue;
      const float invz = (float)(1.0 / p3Dc[2]);
      const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
      const float u = cam.fx * x + cam.cx, v = cam.fy * y + cam.cy;
      if (!(u >= b.min_x && u < b.max_x && v >= b.min_y && v < b.max_y)) continue;
      float PO[3];
      double n2 = 0, dotn = 0;
      for (int k = 0; k < 3; ++k) {
        PO[k] = p3Dw[k] - Ow[k];
        if (c == 1) PO[k] = (float)(-1.0 * RtT12[k] + (double)PO[k]);
        n2 += (double)PO[k] * (double)PO[k];
        dotn += (double)PO[k] * (double)mp_normal[3 * i + k];
      }
      const float dist3D = (float)std::sqrt(n2);
      if (dist3D < mp_min_dist[i] || dist3D > mp_max_dist[i]) continue;
      if (dotn < 0.5 * dist3D) continue;
      const float ratio = mp_max_d[i] / dist3D;
      int lvl = (int)std::ceil(std::log(ratio) / log_scale_factor);
      if (lvl < 0) lvl = 0;
      else if (lvl >= nlevels) lvl = nlevels - 1;
      const float radius = th * scale_factors[lvl];
      const Grid& g = c == 1 ? grids.g1 : grids.g0;
      g.query(u, v, radius, -1, -1, cand);
      int bestDist = INT_MAX, bestIdx = -1;
      for (int idx : cand) {
        const int kpLevel = kf_k[idx].octave;
        if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
        const int dd = om_distance(mp_desc + (size_t)i * 32, kf_desc + (size_t)idx * 32);
        if (dd < bestDist) { bestDist = dd; bestIdx = idx; }
      }
      if (bestIdx < 0) continue;
      if (bestDist <= TH_LOW) { best_idx[2 * i + c] = bestIdx; nFused++; }
    }
  }
  return nFused;
}

// Test hook: the cv::Mat small-matrix algebra helpers above, so that tests can pin them against cv2.gemm.
//   transpose_a == 0: out = alpha * (A x) + c   (float32 path of cv::gemm for tiny non-transposed operands)
//   transpose_a == 1: out = alpha * (A^T x)     (double-accumulating generic path; c must be NULL)
void om_gemm3_probe(const float* A, const float* x, const float* c, float alpha, int transpose_a, float* out) {
  if (transpose_a) mat3t_mul_vec(A, 3, x, alpha, out);
  else mat3_mul_vec_add(A, 3, x, c, alpha, out);
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th, CalibMatrix) (src/ORBmatcher.cc:2814-3136,
// called from LoopClosing::ComputeSim3, src/LoopClosing.cc:402): map points of each key frame are projected into the
// other one through the Sim3, searched in the per-camera grid of their own camera index within th * scale, best
// distance <= TH_HIGH, and only mutually consistent pairs are kept (:3107-3122).
//  key frame x: concatenated keypoints, descriptor per global index, cam (keypoint_to_cam), Txw (4x4 row-major);
//  map points are aligned with the keypoints (GetMapPointMatches): mpx_valid[i] = point exists, !isBad() and not
//  already matched (vbAlreadyMatched, :2846-2861); world position, distance invariance limits, mfMaxDistance, descriptor.
//  match12 (n1) out: key-frame-2 feature of every NEW mutual match, else -1 (vpMatches12[i1] = vpMapPoints2[match12[i1]]).
int om_search_by_sim3(const oo_keypoint* k1, const uint8_t* d1, const int32_t* cam1, int n1, const float* T1w,
                      const oo_keypoint* k2, const uint8_t* d2, const int32_t* cam2, int n2, const float* T2w, om_bounds b,
                      const float* scale_factors, int nlevels, float log_scale_factor, om_camera cam, float s12,
                      const float* R12, const float* t12, const float* calib, const int32_t* mp1_valid, const float* mp1_xyz,
                      const float* mp1_max_dist, const float* mp1_min_dist, const float* mp1_max_d, const uint8_t* mp1_desc,
                      const int32_t* mp2_valid, const float* mp2_xyz, const float* mp2_max_dist, const float* mp2_min_dist,
                      const float* mp2_max_d, const uint8_t* mp2_desc, float th, int32_t* match12) {
  float Rcam21[9], tcam21[3];
  {
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = calib[i];
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1. / d;
    Rcam21[0] = (float)((S[4] * S[8] - S[5] * S[7]) * d); Rcam21[1] = (float)((S[2] * S[7] - S[1] * S[8]) * d);
    Rcam21[2] = (float)((S[1] * S[5] - S[2] * S[4]) * d); Rcam21[3] = (float)((S[5] * S[6] - S[3] * S[8]) * d);
    Rcam21[4] = (float)((S[0] * S[8] - S[2] * S[6]) * d); Rcam21[5] = (float)((S[2] * S[3] - S[0] * S[5]) * d);
    Rcam21[6] = (float)((S[3] * S[7] - S[4] * S[6]) * d); Rcam21[7] = (float)((S[1] * S[6] - S[0] * S[7]) * d);
    Rcam21[8] = (float)((S[0] * S[4] - S[1] * S[3]) * d);
  }
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  // sR12 = s12*R12; sR21 = (1.0/s12)*R12.t(); t21 = -sR21*t12   (:2838-2840): scalings are float multiplies by the
  // scalar rounded to float, the last product takes gemm's float path
  float sR12[9], sR21[9], t21[3];
  const float inv_s12 = (float)(1.0 / (double)s12);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      sR12[i * 3 + j] = R12[i * 3 + j] * s12;
      sR21[i * 3 + j] = R12[j * 3 + i] * inv_s12;
    }
  mat3_mul_vec_add(sR21, 3, t12, nullptr, -1.f, t21);
  const float t1w[3] = {T1w[3], T1w[7], T1w[11]}, t2w[3] = {T2w[3], T2w[7], T2w[11]};
  SoA s1(k1, n1), s2(k2, n2);
  Grid2 g1(s1.x.data(), s1.y.data(), s1.oct.data(), cam1, n1, b), g2(s2.x.data(), s2.y.data(), s2.oct.data(), cam2, n2, b);
  std::vector<int> vnMatch1(n1, -1), vnMatch2(n2, -1), cand;
  // one direction: points of key frame A (pose TAw) into key frame B through (sRBA, tBA)
  auto pass = [&](int nA, const int32_t* camA, const float* TAw, const float* tAw, const float* sRBA, const float* tBA,
                  const int32_t* valid, const float* xyz, const float* maxd, const float* mind, const float* maxD,
                  const uint8_t* mpdesc, const oo_keypoint* kB, const uint8_t* dB, const Grid2& gB, std::vector<int>& out) {
    for (int i = 0; i < nA; ++i) {
      if (!valid[i]) continue;
      const int camIdx = camA ? camA[i] : 0;
      float pA[3], pB[3];
      mat3_mul_vec_add(TAw, 4, xyz + 3 * i, tAw, 1.f, pA);
      mat3_mul_vec_add(sRBA, 3, pA, tBA, 1.f, pB);
      if (camIdx == 1) {
        float tmp[3];
        mat3_mul_vec_add(Rcam21, 3, pB, tcam21, 1.f, tmp);
        pB[0] = tmp[0]; pB[1] = tmp[1]; pB[2] = tmp[2];
      }
      if (pB[2] < 0.0) continue;
      const float invz = (float)(1.0 / pB[2]);
      const float x = pB[0] * invz, y = pB[1] * invz;
      const float u = cam.fx * x + cam.cx, v = cam.fy * y + cam.cy;
      if (!(u >= b.min_x && u < b.max_x && v >= b.min_y && v < b.max_y)) continue;
      double n2s = 0;
      for (int k = 0; k < 3; ++k) n2s += (double)pB[k] * (double)pB[k];
      const float dist3D = (float)std::sqrt(n2s);
      if (dist3D < mind[i] || dist3D > maxd[i]) continue;
      const float ratio = maxD[i] / dist3D;
      int lvl = (int)std::ceil(std::log(ratio) / log_scale_factor);
      if (lvl < 0) lvl = 0;
      else if (lvl >= nlevels) lvl = nlevels - 1;
      const float radius = th * scale_factors[lvl];
      const Grid& g = camIdx == 1 ? gB.g1 : gB.g0;
      g.query(u, v, radius, -1, -1, cand);
      int bestDist = INT_MAX, bestIdx = -1;
      for (int idx : cand) {
        if (kB[idx].octave < lvl - 1 || kB[idx].octave > lvl) continue;
        const int dd = om_distance(mpdesc + (size_t)i * 32, dB + (size_t)idx * 32);
        if (dd < bestDist) { bestDist = dd; bestIdx = idx; }
      }
      if (bestDist <= TH_HIGH) out[i] = bestIdx;
    }
  };
  pass(n1, cam1, T1w, t1w, sR21, t21, mp1_valid, mp1_xyz, mp1_max_dist, mp1_min_dist, mp1_max_d, mp1_desc, k2, d2, g2, vnMatch1);
  pass(n2, cam2, T2w, t2w, sR12, t12, mp2_valid, mp2_xyz, mp2_max_dist, mp2_min_dist, mp2_max_d, mp2_desc, k1, d1, g1, vnMatch2);
  int nFound = 0;
  for (int i1 = 0; i1 < n1; ++i1) {
    match12[i1] = -1;
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && vnMatch2[idx2] == i1) { match12[i1] = idx2; nFound++; }
  }
  return nFound;
}

// DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195
// with the per-feature tree descent :1218-1259), as Frame::ComputeBoW calls it (src/Frame.cc:649-659, levelsup = 4;
// ORBvoc: k = 10, L = 6, TF_IDF weighting, L1 scoring).  The vocabulary tree is given as flat arrays: children of node
// i = child_ids[child_start[i] .. child_start[i+1]) in m_nodes[i].children order, node descriptors (32 bytes), word id
// of the leaves (-1 for inner nodes) and node weights (WordValue = double).
//  per feature: word[i], node[i] (the ancestor at level L - levelsup; 0 when that level is <= 0 or is never reached,
//  where the reference leaves nid uninitialised), weight[i].
//  BowVector (TF / TF_IDF + L1): bow_word/bow_value sorted by word id, *n_bow entries (capacity n): addWeight in
//  feature order (BowVector.cpp:34-46), then normalize(L1) (:62-84).
//  FeatureVector as CSR: fv_node (capacity n) ascending, fv_start (n+1), fv_items (n), *n_fv nodes.
void om_bow_transform(const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc, const int32_t* word_id,
                      const double* node_weight, int n_nodes, int L, const uint8_t* desc, int n, int levelsup, int32_t* word,
                      int32_t* node, double* weight, int32_t* bow_word, double* bow_value, int32_t* n_bow, int32_t* fv_node,
                      int32_t* fv_start, int32_t* fv_items, int32_t* n_fv) {
  (void)n_nodes;
  const int nid_level = L - levelsup;
  std::vector<std::pair<int, double>> bow;                 // BowVector = std::map<WordId, WordValue>
  std::vector<std::pair<int, std::vector<int>>> fv;        // FeatureVector = std::map<NodeId, vector<unsigned>>
  for (int i = 0; i < n; ++i) {
    const uint8_t* f = desc + (size_t)i * 32;
    int nid = 0;  // reference: uninitialised unless nid_level is reached
    int final_id = 0, current_level = 0;
    do {
      ++current_level;
      const int c0 = child_start[final_id], c1 = child_start[final_id + 1];
      final_id = child_ids[c0];
      double best_d = om_distance(f, node_desc + (size_t)final_id * 32);
      for (int c = c0 + 1; c < c1; ++c) {
        const int id = child_ids[c];
        const double d = om_distance(f, node_desc + (size_t)id * 32);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) nid = final_id;
    } while (child_start[final_id] != child_start[final_id + 1]);
    word[i] = word_id[final_id];
    weight[i] = node_weight[final_id];
    node[i] = nid;
    if (weight[i] > 0) {  // not stopped (:1158)
      auto vit = std::lower_bound(bow.begin(), bow.end(), word[i], [](const std::pair<int, double>& a, int b) { return a.first < b; });
      if (vit != bow.end() && vit->first == word[i]) vit->second += weight[i];
      else bow.insert(vit, std::make_pair(word[i], weight[i]));
      auto fit = std::lower_bound(fv.begin(), fv.end(), nid,
                                  [](const std::pair<int, std::vector<int>>& a, int b) { return a.first < b; });
      if (fit != fv.end() && fit->first == nid) fit->second.push_back(i);
      else fv.insert(fit, std::make_pair(nid, std::vector<int>(1, i)));
    }
  }
  double norm = 0.0;  // normalize(L1)
  for (auto& e : bow) norm += std::fabs(e.second);
  if (norm > 0.0)
    for (auto& e : bow) e.second /= norm;
  *n_bow = (int)bow.size();
  for (size_t j = 0; j < bow.size(); ++j) { bow_word[j] = bow[j].first; bow_value[j] = bow[j].second; }
  *n_fv = (int)fv.size();
  int run = 0;
  for (size_t j = 0; j < fv.size(); ++j) {
    fv_node[j] = fv[j].first;
    fv_start[j] = run;
    for (int idx : fv[j].second) fv_items[run++] = idx;
  }
  fv_start[fv.size()] = run;
}

}  // extern "C"
