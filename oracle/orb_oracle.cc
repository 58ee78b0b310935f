// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// Scalar restatement of the reference extractor, /root/reference/src/ORBextractor.cc, over
// the cvprim primitives.  Each function cites the lines it follows.  It is pinned against
// (a) the reference's own ORBextractor.cc compiled verbatim (oracle/_ref, built in the
// authoring container; tests/test_oracle_ref.py) and (b) the committed golden fixtures in
// tests/golden/ generated from that verbatim build (tools/make_golden.py).
// Parity pins: OpenCV 4.13 blur taps; octree size ties broken as
// "later-created node first" (= the reference built with a monotonic allocator); float math
// without FMA contraction (compile with -ffp-contract=off); libm cosf/sinf.
#include "orb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <vector>

#include "cvprim.h"
#include "../multi_orb_slam_b200/csrc/orb_pattern.h"

namespace {

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;  // ORBextractor.cc:70-72
const signed char kPattern[ORB_PATTERN_INTS] = ORB_PATTERN_INIT;

struct Level {
  int w = 0, h = 0;
  size_t step = 0;
  std::vector<uint8_t> buf;  // (w+38) x (h+38)
  uint8_t* roi() { return buf.data() + kEdge * step + kEdge; }
  const uint8_t* roi() const { return buf.data() + kEdge * step + kEdge; }
};

struct Cand { int x, y, score; };          // coordinates relative to (16,16)
struct LevelKP { int x, y; float response, angle; };  // level coordinates

}  // namespace

struct oo_extractor {
  int nfeatures, nlevels, ini_th, min_th;
  double scale_factor;  // the reference stores the ctor's float in a double member (ORBextractor.h:95)
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int> quota, umax;
  std::vector<Level> pyr;
  std::vector<std::vector<Cand>> cands;
  std::vector<std::vector<LevelKP>> level_kps;
  std::vector<std::vector<uint8_t>> blurred;

  // ORBextractor::ORBextractor, ORBextractor.cc:411-471
  oo_extractor(int nf, float sf, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), ini_th(ini), min_th(mn), scale_factor(sf) {
    scale.assign(nl, 1.f);
    sigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; ++i) {
      scale[i] = (float)(scale[i - 1] * scale_factor);  // float * double -> float
      sigma2[i] = scale[i] * scale[i];
    }
    inv_scale.resize(nl);
    inv_sigma2.resize(nl);
    for (int i = 0; i < nl; ++i) {
      inv_scale[i] = 1.0f / scale[i];
      inv_sigma2[i] = 1.0f / sigma2[i];
    }
    quota.resize(nl);
    float factor = (float)(1.0f / scale_factor);
    float desired = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
      quota[l] = cvp::round_f(desired);
      sum += quota[l];
      desired *= factor;
    }
    quota[nl - 1] = std::max(nf - sum, 0);
    // circular patch row half-widths, :455-470
    umax.assign(kHalfPatch + 1, 0);
    int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = cvp::round_d(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    pyr.resize(nl);
    cands.resize(nl);
    level_kps.resize(nl);
    blurred.resize(nl);
  }

  // ComputePyramid, ORBextractor.cc:1109-1134
  void compute_pyramid(const uint8_t* img, int rows, int cols, size_t stride) {
    for (int l = 0; l < nlevels; ++l) {
      Level& L = pyr[l];
      L.w = cvp::round_f((float)cols * inv_scale[l]);
      L.h = cvp::round_f((float)rows * inv_scale[l]);
      L.step = (size_t)L.w + 2 * kEdge;
      L.buf.assign(L.step * (L.h + 2 * kEdge), 0);
      if (l == 0) {
        for (int y = 0; y < rows; ++y) std::memcpy(L.roi() + y * L.step, img + y * stride, cols);
      } else {
        const Level& P = pyr[l - 1];
        cvp::resize_linear_u8(P.roi(), P.step, P.w, P.h, L.roi(), L.step, L.w, L.h);
      }
      cvp::border_reflect101_inplace(L.buf.data(), L.step, L.w, L.h, kEdge);
    }
  }

  // cell loop of ComputeKeyPointsOctTree, ORBextractor.cc:766-830
  void detect_level(int l) {
    const Level& L = pyr[l];
    std::vector<Cand>& out = cands[l];
    out.clear();
    const int minB = kEdge - 3;
    const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<cvp::FastKP> cell;
    for (int i = 0; i < nRows; ++i) {
      const float iniY = (float)(minB + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; ++j) {
        const float iniX = (float)(minB + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const uint8_t* sub = L.roi() + (size_t)(int)iniY * L.step + (int)iniX;
        const int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
        cvp::fast9_16(sub, L.step, cw, ch, ini_th, true, cell);
        if (cell.empty()) cvp::fast9_16(sub, L.step, cw, ch, min_th, true, cell);
        for (const cvp::FastKP& k : cell) out.push_back({k.x + j * wCell, k.y + i * hCell, k.score});
      }
    }
  }

  struct Node {
    int ulx, uly, brx, bry;  // UL and BR corners (UR.x==BR.x, BL.y==BR.y throughout)
    std::vector<int> keys;   // candidate indices, relative order preserved
    bool nomore = false;
    int seq = 0;             // creation order: stands in for the heap address in :685's sort
  };

  // ExtractorNode::DivideNode, ORBextractor.cc:482-538
  static void divide(const Node& n, const std::vector<Cand>& c, Node ch[4]) {
    const int halfX = (int)std::ceil((float)(n.brx - n.ulx) / 2);
    const int halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
    const int mx = n.ulx + halfX, my = n.uly + halfY;
    ch[0] = Node{n.ulx, n.uly, mx, my};
    ch[1] = Node{mx, n.uly, n.brx, my};
    ch[2] = Node{n.ulx, my, mx, n.bry};
    ch[3] = Node{mx, my, n.brx, n.bry};
    for (int k : n.keys) {
      const Cand& p = c[k];
      if (p.x < mx) ch[p.y < my ? 0 : 2].keys.push_back(k);
      else ch[p.y < my ? 1 : 3].keys.push_back(k);
    }
    for (int q = 0; q < 4; ++q) ch[q].nomore = ch[q].keys.size() == 1;
  }

  // DistributeOctTree, ORBextractor.cc:540-764.  Returns candidate indices in final list order.
  std::vector<int> distribute(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY,
                              int N) {
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> L;
    std::vector<Node*> roots(nIni);
    int seq = 0;
    for (int i = 0; i < nIni; ++i) {
      Node n{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY};
      n.seq = seq++;
      L.push_back(n);
      roots[i] = &L.back();
    }
    for (size_t k = 0; k < c.size(); ++k) roots[(size_t)((float)c[k].x / hX)]->keys.push_back((int)k);
    for (auto it = L.begin(); it != L.end();) {
      if (it->keys.size() == 1) { it->nomore = true; ++it; }
      else if (it->keys.empty()) it = L.erase(it);
      else ++it;
    }
    typedef std::list<Node>::iterator It;
    std::vector<It> pending;
    bool finish = false;
    auto push_children = [&](Node ch[4], std::vector<It>& pend, int* nexp) {
      for (int q = 0; q < 4; ++q) {
        if (ch[q].keys.empty()) continue;
        ch[q].seq = seq++;
        L.push_front(ch[q]);
        if (ch[q].keys.size() > 1) { pend.push_back(L.begin()); if (nexp) ++*nexp; }
      }
    };
    while (!finish) {
      const int prev = (int)L.size();
      int nToExpand = 0;
      pending.clear();
      for (It it = L.begin(); it != L.end();) {          // phase-1 pass, :607-666
        if (it->nomore) { ++it; continue; }
        Node ch[4];
        divide(*it, c, ch);
        push_children(ch, pending, &nToExpand);
        it = L.erase(it);
      }
      if ((int)L.size() >= N || (int)L.size() == prev) {
        finish = true;
      } else if ((int)L.size() + nToExpand * 3 > N) {    // phase 2, :675-741
        while (!finish) {
          const int prev2 = (int)L.size();
          std::vector<It> cur = pending;
          pending.clear();
          std::sort(cur.begin(), cur.end(), [](const It& a, const It& b) {
            if (a->keys.size() != b->keys.size()) return a->keys.size() < b->keys.size();
            return a->seq < b->seq;
          });
          for (int j = (int)cur.size() - 1; j >= 0; --j) {
            Node ch[4];
            divide(*cur[j], c, ch);
            push_children(ch, pending, nullptr);
            L.erase(cur[j]);
            if ((int)L.size() >= N) break;
          }
          if ((int)L.size() >= N || (int)L.size() == prev2) finish = true;
        }
      }
    }
    std::vector<int> result;
    for (const Node& n : L) {                            // :745-761
      int best = n.keys[0];
      for (size_t k = 1; k < n.keys.size(); ++k)
        if (c[n.keys[k]].score > c[best].score) best = n.keys[k];
      result.push_back(best);
    }
    return result;
  }

  // IC_Angle, ORBextractor.cc:77-104 (on the un-blurred level)
  float ic_angle(const Level& L, int x, int y) const {
    int m01 = 0, m10 = 0;
    const uint8_t* center = L.roi() + (size_t)y * L.step + x;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
    const int step = (int)L.step;
    for (int v = 1; v <= kHalfPatch; ++v) {
      int v_sum = 0;
      const int d = umax[v];
      for (int u = -d; u <= d; ++u) {
        const int plus = center[u + v * step], minus = center[u - v * step];
        v_sum += plus - minus;
        m10 += u * (plus + minus);
      }
      m01 += v * v_sum;
    }
    return cvp::fast_atan2((float)m01, (float)m10);
  }

  // computeOrbDescriptor, ORBextractor.cc:108-147 (on the blurred copy, stride w)
  static void descriptor(const uint8_t* img, int step, int x, int y, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t* center = img + (size_t)y * step + x;
    const signed char* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
      int val = 0;
      for (int k = 0; k < 8; ++k) {
        const float x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
        const int t0 = center[cvp::round_f(x0 * b + y0 * a) * step + cvp::round_f(x0 * a - y0 * b)];
        const int t1 = center[cvp::round_f(x1 * b + y1 * a) * step + cvp::round_f(x1 * a - y1 * b)];
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // operator(), ORBextractor.cc:1044-1107
  int extract(const uint8_t* img, int rows, int cols, size_t stride, oo_keypoint* kps, uint8_t* desc,
              int cap, int* level_counts) {
    if (level_counts) std::fill(level_counts, level_counts + nlevels, 0);
    if (!img || rows <= 0 || cols <= 0) return 0;
    compute_pyramid(img, rows, cols, stride);
    int total = 0;
    for (int l = 0; l < nlevels; ++l) {                  // ComputeKeyPointsOctTree :766-854
      detect_level(l);
      const Level& L = pyr[l];
      const int minB = kEdge - 3;
      std::vector<int> sel =
          distribute(cands[l], minB, L.w - kEdge + 3, minB, L.h - kEdge + 3, quota[l]);
      std::vector<LevelKP>& out = level_kps[l];
      out.clear();
      for (int k : sel) {
        LevelKP kp{cands[l][k].x + minB, cands[l][k].y + minB, (float)cands[l][k].score, -1.f};
        kp.angle = ic_angle(L, kp.x, kp.y);
        out.push_back(kp);
      }
      total += (int)out.size();
    }
    if (total > cap) return -1;
    int off = 0;
    for (int l = 0; l < nlevels; ++l) {
      const Level& L = pyr[l];
      const std::vector<LevelKP>& lk = level_kps[l];
      if (level_counts) level_counts[l] = (int)lk.size();
      blurred[l].clear();
      if (lk.empty()) continue;
      blurred[l].resize((size_t)L.w * L.h);
      cvp::gaussblur7_sigma2_u8(L.roi(), L.step, blurred[l].data(), L.w, L.w, L.h);
      const int patch = (int)(kPatch * scale[l]);
      for (const LevelKP& k : lk) {
        descriptor(blurred[l].data(), L.w, k.x, k.y, k.angle, desc + (size_t)off * 32);
        oo_keypoint& o = kps[off];
        o.x = (float)k.x;
        o.y = (float)k.y;
        if (l != 0) { o.x *= scale[l]; o.y *= scale[l]; }
        o.size = (float)patch;
        o.angle = k.angle;
        o.response = k.response;
        o.octave = l;
        ++off;
      }
    }
    return total;
  }
};

extern "C" {

oo_extractor* oo_create(int nf, float sf, int nl, int ini, int mn) {
  if (nf < 0 || nl < 1 || nl > 32 || !(sf > 1.f)) return nullptr;
  return new oo_extractor(nf, sf, nl, ini, mn);
}
void oo_destroy(oo_extractor* e) { delete e; }

int oo_extract(oo_extractor* e, const uint8_t* img, int rows, int cols, size_t stride,
               oo_keypoint* kps, uint8_t* desc, int cap, int* level_counts) {
  return e->extract(img, rows, cols, stride, kps, desc, cap, level_counts);
}

int oo_pyramid_level(oo_extractor* e, int level, const uint8_t** data, int* w, int* h, size_t* step) {
  if (level < 0 || level >= e->nlevels || e->pyr[level].buf.empty()) return -1;
  *data = e->pyr[level].roi();
  *w = e->pyr[level].w;
  *h = e->pyr[level].h;
  *step = e->pyr[level].step;
  return 0;
}

void oo_scale_tables(oo_extractor* e, float* s, float* is, float* s2, float* is2) {
  for (int i = 0; i < e->nlevels; ++i) {
    if (s) s[i] = e->scale[i];
    if (is) is[i] = e->inv_scale[i];
    if (s2) s2[i] = e->sigma2[i];
    if (is2) is2[i] = e->inv_sigma2[i];
  }
}
void oo_features_per_level(oo_extractor* e, int* out) {
  for (int i = 0; i < e->nlevels; ++i) out[i] = e->quota[i];
}

int oo_stage_candidates(oo_extractor* e, int level, int* x, int* y, int* score, int cap) {
  const auto& c = e->cands[level];
  for (int i = 0; i < (int)c.size() && i < cap; ++i) { x[i] = c[i].x; y[i] = c[i].y; score[i] = c[i].score; }
  return (int)c.size();
}
int oo_stage_blurred(oo_extractor* e, int level, uint8_t* out) {
  if (e->blurred[level].empty()) return 0;
  std::memcpy(out, e->blurred[level].data(), e->blurred[level].size());
  return 1;
}
int oo_stage_level_keypoints(oo_extractor* e, int level, int* x, int* y, float* response, float* angle, int cap) {
  const auto& k = e->level_kps[level];
  for (int i = 0; i < (int)k.size() && i < cap; ++i) {
    x[i] = k[i].x; y[i] = k[i].y; response[i] = k[i].response; angle[i] = k[i].angle;
  }
  return (int)k.size();
}

}  // extern "C"
