// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// Flat-array restatement of the reference's matcher hot path.  The reference's
// src/ORBmatcher.cc cannot be compiled standalone (it drags in Frame/KeyFrame/MapPoint/DBoW2),
// so this file follows it line by line over plain arrays; each function cites its source.
// "Parity unpinned" by the reference itself: it ships no tests or golden vectors for these
// functions.  Pins we add: hand-checkable known-answer tests in
// tests/test_matcher_oracle.py.
#include "orb_oracle.h"

#include <climits>
#include <cmath>
#include <cstring>
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

}  // extern "C"
