// TEST INFRASTRUCTURE — CPU oracle, not product code.  See cvprim.h for scope and pins.
#include "cvprim.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <algorithm>

namespace cvp {

int round_f(float v) { return (int)lrintf(v); }
int round_d(double v) { return (int)lrint(v); }

// OpenCV core/mathfuncs_core: atan2 via a 7th-order odd minimax polynomial evaluated in
// float32 left-to-right (compile this file with -ffp-contract=off).
float fast_atan2(float y, float x) {
  static const float scale = (float)(180.0 / 3.14159265358979323846);
  static const float p1 = 0.9997878412794807f * scale;
  static const float p3 = -0.3258083974640975f * scale;
  static const float p5 = 0.1555786518463281f * scale;
  static const float p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// OpenCV imgproc/resize.cpp, INTER_LINEAR, 8UC1: HResizeLinear into int32 rows scaled by
// 2048, then VResizeLinear's fixed-point combine.
void resize_linear_u8(const uint8_t* src, size_t sstep, int sw, int sh,
                      uint8_t* dst, size_t dstep, int dw, int dh) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1.0 / inv_scale_x, scale_y = 1.0 / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[2 * dx] = (short)round_f((1.f - fx) * 2048.f);
    ialpha[2 * dx + 1] = (short)round_f(fx * 2048.f);
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    ibeta[2 * dy] = (short)round_f((1.f - fy) * 2048.f);
    ibeta[2 * dy + 1] = (short)round_f(fy * 2048.f);
  }
  std::vector<int> row0(dw), row1(dw);
  auto hresize = [&](int sy, std::vector<int>& out) {
    sy = std::min(std::max(sy, 0), sh - 1);
    const uint8_t* S = src + (size_t)sy * sstep;
    for (int dx = 0; dx < dw; ++dx) {
      int sx = xofs[dx];
      int sx1 = std::min(sx + 1, sw - 1);
      out[dx] = S[sx] * ialpha[2 * dx] + S[sx1] * ialpha[2 * dx + 1];
    }
  };
  int have0 = INT32_MIN, have1 = INT32_MIN;
  for (int dy = 0; dy < dh; ++dy) {
    int sy = yofs[dy];
    if (have1 == sy && dy > 0) { row0.swap(row1); have0 = have1; have1 = INT32_MIN; }
    if (have0 != sy) { hresize(sy, row0); have0 = sy; }
    if (have1 != sy + 1) { hresize(sy + 1, row1); have1 = sy + 1; }
    const int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
    uint8_t* D = dst + (size_t)dy * dstep;
    for (int dx = 0; dx < dw; ++dx)
      D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

void border_reflect101_inplace(uint8_t* buf, size_t step, int w, int h, int b) {
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
  };
  // left/right columns of interior rows
  for (int y = 0; y < h; ++y) {
    uint8_t* row = buf + (size_t)(y + b) * step + b;
    for (int k = 1; k <= b; ++k) {
      row[-k] = row[refl(-k, w)];
      row[w - 1 + k] = row[refl(w - 1 + k, w)];
    }
  }
  // top/bottom rows (full width, so corners reflect both ways)
  for (int k = 1; k <= b; ++k) {
    std::memcpy(buf + (size_t)(b - k) * step, buf + (size_t)(b + refl(-k, h)) * step, w + 2 * b);
    std::memcpy(buf + (size_t)(b + h - 1 + k) * step, buf + (size_t)(b + refl(h - 1 + k, h)) * step,
                w + 2 * b);
  }
}

// OpenCV 4.x imgproc/smooth.dispatch.cpp fixed-point Gaussian for CV_8U: kernel from
// getGaussianKernelBitExact -> ufixedpoint16 taps {18,34,48,56,48,34,18} (8.8, sum 256);
// horizontal pass u8 x 8.8 -> 8.8 (16 bit), vertical pass 8.8 x 8.8 -> 16.16, rounded to u8.
void gaussblur7_sigma2_u8(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep,
                          int w, int h) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
  };
  std::vector<uint16_t> H((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* S = src + (size_t)y * sstep;
    uint16_t* Hr = &H[(size_t)y * w];
    for (int x = 0; x < w; ++x) {
      int s = 0;
      if (x >= 3 && x < w - 3) {
        for (int i = 0; i < 7; ++i) s += K[i] * S[x + i - 3];
      } else {
        for (int i = 0; i < 7; ++i) s += K[i] * S[refl(x + i - 3, w)];
      }
      Hr[x] = (uint16_t)s;
    }
  }
  for (int y = 0; y < h; ++y) {
    const uint16_t* R[7];
    for (int j = 0; j < 7; ++j) R[j] = &H[(size_t)refl(y + j - 3, h) * w];
    uint8_t* D = dst + (size_t)y * dstep;
    for (int x = 0; x < w; ++x) {
      uint32_t v = 0;
      for (int j = 0; j < 7; ++j) v += (uint32_t)K[j] * R[j][x];
      D[x] = (uint8_t)((v + 32768u) >> 16);
    }
  }
}

static inline void ring_offsets(size_t step, ptrdiff_t off[16]) {
  static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  for (int k = 0; k < 16; ++k) off[k] = (ptrdiff_t)dy[k] * (ptrdiff_t)step + dx[k];
}

static inline int arc_measure_off(const uint8_t* p, const ptrdiff_t off[16]) {
  int d[25];
  const int c = p[0];
  for (int k = 0; k < 16; ++k) d[k] = c - p[off[k]];
  for (int k = 0; k < 9; ++k) d[16 + k] = d[k];
  int best = -255;
  for (int s = 0; s < 16; ++s) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; ++k) {
      mn = std::min(mn, d[s + k]);
      mx = std::max(mx, d[s + k]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

int fast9_arc_measure(const uint8_t* p, size_t step) {
  ptrdiff_t off[16];
  ring_offsets(step, off);
  return arc_measure_off(p, off);
}

// OpenCV features2d/fast.cpp FAST_t<16>: corner iff some 9-arc is entirely brighter than
// v+th or darker than v-th; score (cornerScore<16>) = largest such th = m-1; 3x3 strict NMS
// over the score rows, rows/cols within 3 px of the (sub)image edge never tested (score 0).
void fast9_16(const uint8_t* img, size_t step, int w, int h, int threshold, bool nms,
              std::vector<FastKP>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  ptrdiff_t off[16];
  ring_offsets(step, off);
  threshold = std::min(std::max(threshold, 0), 255);
  std::vector<uint8_t> sbuf((size_t)3 * w, 0);
  std::vector<int> cbuf((size_t)3 * (w + 1), 0);
  uint8_t* rows[3] = {&sbuf[0], &sbuf[w], &sbuf[2 * (size_t)w]};
  int* cpos[3] = {&cbuf[0], &cbuf[w + 1], &cbuf[2 * (size_t)(w + 1)]};
  for (int i = 3; i < h - 2; ++i) {
    uint8_t* curr = rows[(i - 3) % 3];
    int* cornerpos = cpos[(i - 3) % 3];
    std::memset(curr, 0, w);
    int ncorners = 0;
    if (i < h - 3) {
      const uint8_t* ptr = img + (size_t)i * step;
      for (int j = 3; j < w - 3; ++j) {
        const uint8_t* p = ptr + j;
        const int v = p[0], hi = v + threshold, lo = v - threshold;
        // high-speed rejection: every opposite pair (k, k+8) must contain a ring pixel of
        // the arc's polarity
        bool br = true, dk = true;
        for (int k = 0; k < 8 && (br || dk); k += 2) {
          const int a = p[off[k]], b = p[off[k + 8]];
          br = br && (a > hi || b > hi);
          dk = dk && (a < lo || b < lo);
        }
        if (!br && !dk) continue;
        const int m = arc_measure_off(p, off);
        if (m > threshold) {
          if (nms) {
            cornerpos[ncorners++] = j;
            curr[j] = (uint8_t)(m - 1);
          } else {
            out.push_back({j, i, 0});
          }
        }
      }
    }
    cornerpos[w] = ncorners;  // stash count at the tail slot
    if (i == 3 || !nms) continue;
    const uint8_t* prev = rows[(i - 4 + 3) % 3];
    const uint8_t* pprev = rows[(i - 5 + 3) % 3];
    const int* pc = cpos[(i - 4 + 3) % 3];
    const int npc = pc[w];
    for (int k = 0; k < npc; ++k) {
      const int j = pc[k];
      const int score = prev[j];
      if (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] &&
          score > pprev[j] && score > pprev[j + 1] && score > curr[j - 1] &&
          score > curr[j] && score > curr[j + 1])
        out.push_back({j, i - 1, score});
    }
  }
}

}  // namespace cvp

namespace cvp {
void undistort_points(const float* src_xy, int n, float fxf, float fyf, float cxf, float cyf, const float* dist5,
                      float* dst_xy) {
  // A = K converted to double; k[0..4] = k1 k2 p1 p2 k3, k[5..13] = 0 (rational / thin-prism / tilt terms unused)
  const double fx = fxf, fy = fyf, cx = cxf, cy = cyf;
  const double ifx = 1. / fx, ify = 1. / fy;
  double k[14] = {0};
  for (int i = 0; i < 5; ++i) k[i] = dist5[i];
  // RR = P * R with R = I, P = K (cvMatMul in double; the products by the zeros and the one are exact)
  const double RR[3][3] = {{fx, 0, cx}, {0, fy, cy}, {0, 0, 1}};
  for (int i = 0; i < n; ++i) {
    double x = src_xy[2 * i], y = src_xy[2 * i + 1];
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    // tilt compensation with the identity matrix: vecUntilt = (x, y, 1), invProj = 1
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) {  // OpenCV regression 14583
        x = (u - cx) * ifx;
        y = (v - cy) * ify;
        break;
      }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    const double xx = RR[0][0] * x + RR[0][1] * y + RR[0][2];
    const double yy = RR[1][0] * x + RR[1][1] * y + RR[1][2];
    const double ww = 1. / (RR[2][0] * x + RR[2][1] * y + RR[2][2]);
    dst_xy[2 * i] = (float)(xx * ww);
    dst_xy[2 * i + 1] = (float)(yy * ww);
  }
}
}  // namespace cvp
