// Hand-written sm_100a kernels of the ORB extractor hot path.  Reference stages (all in
// /root/reference/src/ORBextractor.cc): ComputePyramid :1109-1134 (cv::resize INTER_LINEAR +
// copyMakeBorder REFLECT_101), per-cell cv::FAST with ini/min threshold fallback :790-830,
// DistributeOctTree :540-764, IC_Angle :77-104, GaussianBlur 7x7 s=2 :1087,
// computeOrbDescriptor :108-147.  All pixel work is integer and bit-exact by construction;
// the float pieces (fastAtan2, sin/cos, rotation) use explicit round-to-nearest intrinsics so
// nvcc cannot contract them into FMAs.
//
// Batch layout: every kernel covers ALL frames of a batch in one launch (blockIdx.y or .z =
// frame) so the small per-frame work fills the 148 SMs.
#include <cuda_runtime.h>
#include <cstdlib>
#include <mutex>
#include <stdint.h>
#include <stdio.h>

#ifdef ORB_OT_TIMING  // dev-only: clock64 marks inside the octree kernel (block 0), dumped by orbk::dump_octree_marks
__device__ long long g_ot_marks[2 * 4096];
__device__ int g_ot_nmarks;
#define OT_MARK(id) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && g_ot_nmarks < 4096) { g_ot_marks[2 * g_ot_nmarks] = (id); g_ot_marks[2 * g_ot_nmarks + 1] = clock64(); ++g_ot_nmarks; } } while (0)
#endif
#include "extract_kernels.h"
#include "orb_pattern.h"

namespace orbk {

__device__ __forceinline__ int reflect101(int p, int n) {
  p = p < 0 ? -p : p;
  return p >= n ? 2 * (n - 1) - p : p;
}

// ------------------------------------------------------------------------------------------
// Pyramid storage: level 0 is the caller's frame, read in place; levels >= 1 are stored WITHOUT
// the EDGE_THRESHOLD border (rows padded to 16 bytes).  Nothing inside the extractor reads the
// border (FAST cells stay in [16, w-16), the orientation disc and the descriptor pattern stay
// >= 4 px inside, the blur mirrors at the edge itself), so it is synthesised only when a caller
// asks for mvImagePyramid with its border (orbx_get_pyramid_level).
struct LevelView {
  const uint8_t* base;  // row 0, column 0 of this frame's level
  int pitch;
};
__device__ __forceinline__ LevelView level_view(const OrbGeom* __restrict__ g, const OrbLevel0& l0,
                                                const uint8_t* __restrict__ pyr, int level, int frame) {
  LevelView v;
  if (level == 0) {
    v.base = l0.base + (size_t)frame * l0.frame_stride;
    v.pitch = l0.pitch;
  } else {
    v.base = pyr + (size_t)frame * g->pyr_frame_bytes + g->lv[level].pyr_off;
    v.pitch = g->lv[level].pitch;
  }
  return v;
}

// K1 level l >= 1: cv::resize(level l-1 -> l, INTER_LINEAR) (:1122).  One WARP per tile of
// 32 words x `rows_per_tile` (<= 32) rows of the destination level, no shared memory, no barriers:
//   * a lane owns one aligned 32-bit word = four adjacent destination columns;
//   * its four horizontal taps live in registers; per source row it loads the three aligned words
//     that cover its taps and picks each tap pair with one byte-permute + dp2a
//     (h = a0*S[sx] + a1*S[sx+1]);
//   * the loop runs over the SOURCE rows the tile touches, each loaded and filtered exactly once,
//     the loads of the next four rows in flight while four rows are processed (the kernel is bound
//     by load latency, not by issue slots); an output row is emitted when its lower source row
//     arrives (sy1 = sy0 + 1 except at the clamped bottom edge, where both taps are the last row);
//   * the rows' vertical taps are fetched once per tile (lane r holds row r) and broadcast by
//     shuffle.
#define PYR_ROWS 32
#ifndef PYR_PREFETCH
#define PYR_PREFETCH 4  // source rows in flight per warp (even)
#endif
#ifndef PYR_MINB
#define PYR_MINB 8
#endif

__global__ void __launch_bounds__(128, PYR_MINB) k_pyr_resize(int level, int rows_per_tile, const OrbGeom* __restrict__ g,
                                                    OrbLevel0 l0, const OrbXTap* __restrict__ xtab,
                                                    const OrbYTap* __restrict__ ytab, uint8_t* __restrict__ pyr) {
  const OrbLevelGeom& D = g->lv[level];
  const int lane = threadIdx.x & 31;
  const int words = D.pitch >> 2, rows_total = D.h;
  const int tiles_x = (words + 31) >> 5;
  const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int ty_ = tile / tiles_x, tx_ = tile - ty_ * tiles_x;
  const int by0 = ty_ * rows_per_tile;
  if (by0 >= rows_total) return;                        // whole warp
  const int wq = min(tx_ * 32 + lane, words - 1);       // lanes past the row end duplicate the row's last word
  const LevelView S = level_view(g, l0, pyr, level - 1, blockIdx.y);
  uint8_t* drow = pyr + (size_t)blockIdx.y * g->pyr_frame_bytes + D.pyr_off;
  // horizontal taps of the four columns (row padding beyond w is written as 0)
  uint32_t a01[4];
  int sx[4];
  int lo_col = 1 << 30;
  uint32_t live_mask = 0;  // bytes of the output word that are image columns
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = 4 * wq + k;
    live_mask |= x < D.w ? 0xffu << (8 * k) : 0u;
    const OrbXTap t = xtab[D.xtab_off + min(x, D.w - 1)];  // padding columns reuse the last column's tap
    a01[k] = (uint32_t)t.a0 | (uint32_t)t.a1 << 16;
    sx[k] = t.sx;
    lo_col = min(lo_col, sx[k]);
  }
  const int w0 = min(lo_col >> 2, (S.pitch >> 2) - 3);  // three words from here cover all four tap pairs
  // Tap pair k starts at byte o = sx[k] - 4*w0 of the 12 staged bytes.  Unless the row end clamped
  // w0 (only the last lanes of a row), o - (lo_col & 3) <= 6: the warp then aligns the staged bytes to
  // lo_col with two funnel shifts and every pair is one byte-permute of the aligned 8 bytes.  Warps
  // with a clamped lane pick each pair from words (0,1) or (1,2) with selects instead.
  const int sh_bytes = lo_col - 4 * w0;
  const int span = max(max(sx[0], sx[1]), max(sx[2], sx[3])) - lo_col;  // <= 6 up to scale factor 2
  const bool aligned_path = __all_sync(0xffffffffu, sh_bytes < 4 && span <= 6);
  const uint32_t sh8 = 8u * (uint32_t)(sh_bytes & 3);
  uint32_t sel[4];
  bool upper[4];  // tap pair taken from words (1,2) instead of (0,1) (select path)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int o = sx[k] - 4 * w0;  // 0..11
    upper[k] = o > 6;
    o -= upper[k] ? 4 : 0;
    if (aligned_path) o = sx[k] - lo_col;  // 0..6
    // the right tap of the last source column has weight 0 (OpenCV clamps sx to sw-1 with fx = 0):
    // point it at the left tap so the window never has to reach past the row
    const int o1 = (a01[k] >> 16) ? o + 1 : o;
    sel[k] = (uint32_t)o | (uint32_t)o1 << 4 | 0x4400u;
  }
  const unsigned char* src = S.base + 4 * (size_t)w0;
  const uint32_t spitch = (uint32_t)S.pitch;
  const int nrows = min(rows_per_tile, rows_total - by0);
  uint2 my_t = make_uint2(0xffffffffu, 0);  // lanes without a row: sy1 = 0xffff never arrives
  if (lane < nrows) my_t = __ldg(reinterpret_cast<const uint2*>(ytab + D.ytab_off + by0 + lane));
  const uint32_t s_lo = __shfl_sync(0xffffffffu, my_t.x, 0) & 0xffffu;
  const uint32_t s_hi = __shfl_sync(0xffffffffu, my_t.x, nrows - 1) >> 16;
  // source rows are walked by a 32-bit byte offset that saturates at the tile's last row
  // (the prefetch runs up to seven rows ahead)
  const uint32_t off_last = s_hi * spitch;
  uint32_t off = s_lo * spitch;
  auto load = [&](uint32_t (&x)[3]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(src + off);
    x[0] = __ldg(r); x[1] = __ldg(r + 1); x[2] = __ldg(r + 2);
    off = min(off + spitch, off_last);
  };
  auto hcalc = [&](const uint32_t (&x)[3], uint32_t (&h)[4]) {
    if (aligned_path) {
      const uint32_t w_lo = __funnelshift_r(x[0], x[1], sh8), w_hi = __funnelshift_r(x[1], x[2], sh8);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __dp2a_lo(a01[k], __byte_perm(w_lo, w_hi, sel[k]), 0u) >> 4;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t pair = __byte_perm(upper[k] ? x[1] : x[0], upper[k] ? x[2] : x[1], sel[k]);
        h[k] = __dp2a_lo(a01[k], pair, 0u) >> 4;
      }
    }
  };
  uint32_t* dst = reinterpret_cast<uint32_t*>(drow) + (size_t)by0 * words + wq;
  // taps of the next output row to emit
  int r = 0;
  uint32_t tlo = __shfl_sync(0xffffffffu, my_t.x, 0), thi = __shfl_sync(0xffffffffu, my_t.y, 0);
  auto emit = [&](const uint32_t (&h0)[4], const uint32_t (&h1)[4]) {
    // vertical weights pre-shifted by 16: (b * h) >> 16 becomes the high word of (b << 16) * h
    const uint32_t b0 = thi << 16, b1 = thi & 0xffff0000u;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)  // <= 255 by construction of the weights
      v[k] = (__umulhi(b1, h1[k]) + (__umulhi(b0, h0[k]) + 2u)) >> 2;
    const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    *dst = out & live_mask;  // lanes past the row end repeat the last word's store (same address, same value)
    dst += words;
    ++r;  // r == 32 wraps to row 0, whose source rows lie behind: it never matches again
    tlo = __shfl_sync(0xffffffffu, my_t.x, r);
    thi = __shfl_sync(0xffffffffu, my_t.y, r);
  };
  uint32_t s = s_lo;
  // one source row: horizontal pass into HN, then every output row whose lower tap is this row
  // (at most two, at the clamped bottom edge, where the upper tap can be this row as well)
#define PYR_ROW(X, HP, HN)                                  \
  if (s <= s_hi) {                                          \
    hcalc(X, HN);                                           \
    while ((tlo >> 16) == s) {                              \
      if ((tlo & 0xffffu) == s) emit(HN, HN);               \
      else emit(HP, HN);                                    \
    }                                                       \
  }                                                         \
  ++s;
  uint32_t xa[PYR_PREFETCH][3], xb[PYR_PREFETCH][3];
  uint32_t hA[4] = {0, 0, 0, 0}, hB[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < PYR_PREFETCH; ++j) load(xa[j]);
  while (s <= s_hi) {
#pragma unroll
    for (int j = 0; j < PYR_PREFETCH; ++j) load(xb[j]);
#pragma unroll
    for (int j = 0; j < PYR_PREFETCH; j += 2) { PYR_ROW(xa[j], hB, hA) PYR_ROW(xa[j + 1], hA, hB) }
    if (s > s_hi) break;
#pragma unroll
    for (int j = 0; j < PYR_PREFETCH; ++j) load(xa[j]);
#pragma unroll
    for (int j = 0; j < PYR_PREFETCH; j += 2) { PYR_ROW(xb[j], hB, hA) PYR_ROW(xb[j + 1], hA, hB) }
  }
#undef PYR_ROW
}

// ------------------------------------------------------------------------------------------
// K3 per-cell FAST-9/16 with NMS and the iniThFAST -> minThFAST fallback (:790-830).
// One CTA (FAST_NT threads = one warp) per (cell, frame).  The arc measure m = max over the 16 cyclic 9-arcs of
// min(+-diff) is threshold independent: corner iff m > th, score m-1 (OpenCV cornerScore<16>);
// a corner survives the 3x3 NMS iff its m is strictly greater than its 8 neighbours' m
// (neighbours outside this cell's tested area count as 0 — NMS is per cell).
//
// Work is split the way the arithmetic cost falls (on the synthetic frames 16-30 % of the pixels
// of a level survive the rejection test at th=20 and 5-12 % are corners):
//   phase A  every tested pixel: opposite-pair rejection test on packed bytes, four pixels per
//            lane (~16 lane-instr/px), surviving words queued, then expanded to one entry per pixel;
//   phase B  queued pixels only: full arc measure on 16-bit pairs carrying both polarities
//            (VIMNMX3.S16x2), corners compacted in place;
//   phase C  3x3 NMS over the corner list.
// One warp walks the cell in row-major order and every compaction is ordered, so all lists stay in
// cv::FAST's output order.  The pass runs at iniThFAST; only a cell that kept nothing reruns at
// minThFAST (:813-817).  (Trimming the per-cell setup - float-reciprocal divisions, flat-indexed
// tile load - was measured and changes nothing: 1.126 vs 1.118 ms.)
#define FAST_NT 32  // threads per cell CTA: ONE warp per cell.  Measured with more warps per cell (128 frames, ms):
                    // 32 -> 0.42, 64 -> 0.48, 96 -> 0.49, 128 -> 0.50, 160 -> 0.56, 256 -> 0.74: a cell is ~1000 px,
                    // more warps only add barrier and tail idle time

#ifndef FAST_PAIR
#define FAST_PAIR 0  // 1: phase A handles two adjacent words per lane with 64-bit shared loads (11 LDS per 8 pixels
                     // instead of 22); the tile then starts one word further right so that the pairs are 8-byte aligned
#endif
#ifndef FAST_HVP
#define FAST_HVP 0   // bytes per row of the HALVED copy (0: same as the image copy).  Phase A's lanes cover four rows of
                     // eight words each: with 24 words per row the four row segments fall on disjoint banks (a 12-word
                     // pitch gives 2-way conflicts on every one of its 11 loads: ncu 169 wavefronts for 86 loads per cell)
#endif
#define FAST_COL0 (FAST_PAIR ? 5 : 1)  // shared-row byte of sub-image column 0: the tested area starts at byte FAST_COL0 + 3

struct FastSmem {
  // Three regions of shared memory, each reused once its first tenant is dead (6 KB per cell CTA
  // at 640x480, so the SM's limit of 32 resident CTAs is reached):
  uint8_t* img;        // [rows][PITCH] (PITCH = 48 or 80 bytes); sub-image column sx sits at byte 1 + sx, so the
                       // tested area (sx >= 3) starts on a word boundary
  uint16_t* kept;      // (aliases img) NMS survivors (byte offsets), unordered: written after the last read of
                       // img, and only by a pass that is not followed by another
  uint8_t* hv;         // img with every byte halved (p >> 1): input of the packed rejection test
  uint16_t* queue;     // (aliases hv) phase-A survivors (byte offsets into img / m), written after phase A
  uint8_t* m;          // arc measure map, 0 = not a corner at the pass threshold; zeroed after phase A
  uint16_t* eq;        // (aliases m) phase A's word entries
};

template <int PITCH>
__device__ __forceinline__ int fast_arc_measure(const uint8_t* p) {
  constexpr int RO[16] = {3 * PITCH,      3 * PITCH + 1,  2 * PITCH + 2,  PITCH + 3,
                          3,                   -PITCH + 3,     -2 * PITCH + 2, -3 * PITCH + 1,
                          -3 * PITCH,     -3 * PITCH - 1, -2 * PITCH - 2, -PITCH - 3,
                          -3,                  PITCH - 3,      2 * PITCH - 2,  3 * PITCH - 1};
  // Both polarities at once on 16-bit pairs: a ring value p travels as (p, 255 - p), so one
  // packed minimum over an arc yields min(p) (bright arc) and 255 - max(p) (dark arc), and the
  // centre is subtracted once at the end.  VIMNMX3.S16x2: 40 three-input ops per pixel.
  uint32_t d[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) d[k] = (uint32_t)p[RO[k]] * 0xFFFF0001u + 0x00FF0000u;
  uint32_t lo3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) lo3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  uint32_t best = 0;  // max over the 16 arcs of the packed arc minimum
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const uint32_t a0 = __vimin3_s16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
    const uint32_t a1 = __vimin3_s16x2(lo3[k + 1], lo3[(k + 4) & 15], lo3[(k + 7) & 15]);
    best = __vimax3_s16x2(best, a0, a1);
  }
  const int c = p[0];
  return max((int)(best & 0xffffu) - c, (int)(best >> 16) - 255 + c);
}

// One threshold pass over the tested area (tw x thh pixels starting at sub-image (3,3)).
// Returns the number of NMS survivors, left (unordered) in sm.kept.
// `bm_rows` != nullptr: the survivors of the rejection test come from the pass-bit bitmap k_fast_prefilter wrote (one
// 32-byte row per tested row of the cell's band, bit = column - band origin; `bm_bit0` = bit of the cell's first tested
// column) instead of phase A: a lane takes a row, counts its bits, and expands them behind the rows above it.
template <int PITCH>
__device__ __forceinline__ int fast_pass(const FastSmem& sm, int th, int tw, int thh, int ch,
                                         const uint32_t* __restrict__ bm_rows = nullptr, int bm_bit0 = 0) {
  static_assert(FAST_NT == 32, "one warp per cell");
  constexpr int PW = PITCH / 4;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int nq = 0;
  if (bm_rows) {
    const int k = bm_bit0 >> 5, sh = bm_bit0 & 31;
    const uint32_t keep0 = tw >= 32 ? 0xFFFFFFFFu : (1u << tw) - 1u;
    const uint32_t keep1 = tw > 32 ? (tw >= 64 ? 0xFFFFFFFFu : (1u << (tw - 32)) - 1u) : 0u;
    for (int r0 = 0; r0 < thh; r0 += 32) {
      const int r = r0 + lane;
      uint32_t m0 = 0, m1 = 0;
      if (r < thh) {
        const uint32_t* row = bm_rows + 8 * r;
        const uint32_t w0 = __ldg(row + k), w1 = k + 1 < 8 ? __ldg(row + k + 1) : 0u, w2 = k + 2 < 8 ? __ldg(row + k + 2) : 0u;
        m0 = __funnelshift_r(w0, w1, sh) & keep0;
        m1 = __funnelshift_r(w1, w2, sh) & keep1;
      }
      const int cnt = __popc(m0) + __popc(m1);
      int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      uint16_t* q = sm.queue + nq + incl - cnt;
      const int off0 = (r + 3) * PITCH + FAST_COL0 + 3;
      while (m0) {
        *q++ = (uint16_t)(off0 + __ffs(m0) - 1);
        m0 &= m0 - 1;
      }
      while (m1) {
        *q++ = (uint16_t)(off0 + 32 + __ffs(m1) - 1);
        m1 &= m1 - 1;
      }
      nq += __shfl_sync(0xffffffffu, incl, 31);
    }
  } else {
  // phase A: packed rejection test, four pixels (one word) per lane.  A 9-arc contains one pixel
  // of every opposite pair (k, k+8), so a corner has |p - c| > th for at least one pixel of each
  // of the pairs 0/8, 4/12, 2/10, 6/14.  On the halved bytes (a' = a >> 1) |a - c| > th implies
  // |a' - c'| >= th >> 1, one VABSDIFF4 + one add per ring word: the add carries into bit 7 of a
  // byte exactly when the halved difference reaches T (all sums stay below 256, so bytes do not
  // interact).  The test only has to be a superset of the corners: phase B measures exactly.
  const int T = th >> 1;
  const uint32_t K = 0x01010101u * (uint32_t)(128 - min(T, 127));
#if FAST_PAIR
  const int WPR = (tw + 3) >> 2;  // words per tested row
  const uint32_t last_mask = 0x80808080u >> (8 * (4 * WPR - tw));  // valid pixels of a row's last word
  const uint32_t* hw = reinterpret_cast<const uint32_t*>(sm.hv);
  uint16_t* eq = sm.eq;  // word entries: word offset << 4 | pixel nibble
  int ecount = 0;
  auto reject = [&](uint32_t c, uint32_t u3, uint32_t d3, uint32_t lw, uint32_t rw, uint32_t ul, uint32_t uc, uint32_t ur,
                    uint32_t dl, uint32_t dc, uint32_t dq) {
    const uint32_t f0 = __vabsdiffu4(d3, c) + K, f8 = __vabsdiffu4(u3, c) + K;
    const uint32_t f4 = __vabsdiffu4(__byte_perm(c, rw, 0x6543), c) + K;
    const uint32_t f12 = __vabsdiffu4(__byte_perm(lw, c, 0x4321), c) + K;
    const uint32_t f2 = __vabsdiffu4(__byte_perm(dc, dq, 0x5432), c) + K;
    const uint32_t f10 = __vabsdiffu4(__byte_perm(ul, uc, 0x5432), c) + K;
    const uint32_t f6 = __vabsdiffu4(__byte_perm(uc, ur, 0x5432), c) + K;
    const uint32_t f14 = __vabsdiffu4(__byte_perm(dl, dc, 0x5432), c) + K;
    return (f0 | f8) & (f4 | f12) & (f2 | f10) & (f6 | f14);
  };
  const int PPR = (WPR + 1) >> 1;  // word pairs per tested row
  const int items = thh * PPR;
  const int dr = 32 / PPR, dw = 32 - dr * PPR;
  int row = lane / PPR, w2 = lane - row * PPR;
  for (int i0 = 0; i0 < items; i0 += 32) {
    const bool valid = i0 + lane < items;
    const int woff = ((valid ? row : 0) + 3) * PW + (FAST_COL0 + 3) / 4 + 2 * w2;  // even: 8-byte aligned
    const uint32_t* p = hw + woff;
    const uint2 c = *reinterpret_cast<const uint2*>(p);
    const uint2 u3 = *reinterpret_cast<const uint2*>(p - 3 * PW), d3 = *reinterpret_cast<const uint2*>(p + 3 * PW);
    const uint2 uc = *reinterpret_cast<const uint2*>(p - 2 * PW), dc = *reinterpret_cast<const uint2*>(p + 2 * PW);
    const uint32_t lw = p[-1], rw = p[2], ul = p[-2 * PW - 1], ur = p[-2 * PW + 2], dl = p[2 * PW - 1], dq = p[2 * PW + 2];
    uint32_t pass0 = reject(c.x, u3.x, d3.x, lw, c.y, ul, uc.x, uc.y, dl, dc.x, dc.y);
    uint32_t pass1 = reject(c.y, u3.y, d3.y, c.x, rw, uc.x, uc.y, ur, dc.x, dc.y, dq);
    const int wi = 2 * w2;
    pass0 &= wi == WPR - 1 ? last_mask : 0x80808080u;
    pass1 &= wi + 1 == WPR - 1 ? last_mask : (wi + 1 < WPR ? 0x80808080u : 0u);
    if (!valid) pass0 = pass1 = 0;
    const unsigned b0 = __ballot_sync(0xffffffffu, pass0 != 0), b1 = __ballot_sync(0xffffffffu, pass1 != 0);
    const int pos = ecount + __popc(b0 & lt) + __popc(b1 & lt);
    if (pass0) eq[pos] = (uint16_t)(woff << 4 | ((pass0 >> 7) * 0x10204080u) >> 28);
    if (pass1) eq[pos + (pass0 != 0)] = (uint16_t)((woff + 1) << 4 | ((pass1 >> 7) * 0x10204080u) >> 28);
    ecount += __popc(b0) + __popc(b1);
    w2 += dw;
    row += dr;
    if (w2 >= PPR) { w2 -= PPR; ++row; }
  }
#else
  constexpr int PWH = (FAST_HVP ? FAST_HVP : PITCH) / 4;  // words per row of the halved copy
  const int WPR = (tw + 3) >> 2;  // words per tested row
  const int items = thh * WPR;
  const int dr = 32 / WPR, dw = 32 - dr * WPR;
  const uint32_t last_mask = 0x80808080u >> (8 * (4 * WPR - tw));  // valid pixels of a row's last word
  int row = lane / WPR, w = lane - row * WPR;
  const uint32_t* hw = reinterpret_cast<const uint32_t*>(sm.hv);
  uint16_t* eq = sm.eq;  // word entries: word offset (image copy) << 4 | pixel nibble
  int ecount = 0;
  for (int i0 = 0; i0 < items; i0 += 32) {
    const bool valid = i0 + lane < items;
    const int woff = ((valid ? row : 0) + 3) * PW + 1 + w;
    const uint32_t* p = hw + ((valid ? row : 0) + 3) * PWH + 1 + w;
    const uint32_t c = p[0];
    const uint32_t u3 = p[-3 * PWH], d3 = p[3 * PWH], lw = p[-1], rw = p[1];
    const uint32_t ul = p[-2 * PWH - 1], uc = p[-2 * PWH], ur = p[-2 * PWH + 1];
    const uint32_t dl = p[2 * PWH - 1], dc = p[2 * PWH], dq = p[2 * PWH + 1];
    const uint32_t f0 = __vabsdiffu4(d3, c) + K, f8 = __vabsdiffu4(u3, c) + K;
    const uint32_t f4 = __vabsdiffu4(__byte_perm(c, rw, 0x6543), c) + K;
    const uint32_t f12 = __vabsdiffu4(__byte_perm(lw, c, 0x4321), c) + K;
    const uint32_t f2 = __vabsdiffu4(__byte_perm(dc, dq, 0x5432), c) + K;
    const uint32_t f10 = __vabsdiffu4(__byte_perm(ul, uc, 0x5432), c) + K;
    const uint32_t f6 = __vabsdiffu4(__byte_perm(uc, ur, 0x5432), c) + K;
    const uint32_t f14 = __vabsdiffu4(__byte_perm(dl, dc, 0x5432), c) + K;
    uint32_t pass = (f0 | f8) & (f4 | f12) & (f2 | f10) & (f6 | f14) & (w == WPR - 1 ? last_mask : 0x80808080u);
    if (!valid) pass = 0;
    const unsigned bm = __ballot_sync(0xffffffffu, pass != 0);
    if (pass) eq[ecount + __popc(bm & lt)] = (uint16_t)(woff << 4 | ((pass >> 7) * 0x10204080u) >> 28);
    ecount += __popc(bm);
    w += dw;
    row += dr;
    if (w >= WPR) { w -= WPR; ++row; }
  }
#endif
  __syncwarp();
  // expand the word entries into one queue entry per surviving pixel
  for (int e0 = 0; e0 < ecount; e0 += 32) {
    const int e = e0 + lane < ecount ? eq[e0 + lane] : 0;
    const int nib = e & 15, cnt = __popc(nib);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    int pos = nq + incl - cnt;
    const int boff = (e >> 4) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (nib >> k & 1) sm.queue[pos++] = (uint16_t)(boff + k);
    nq += __shfl_sync(0xffffffffu, incl, 31);
  }
  }  // phase A in this kernel
  __syncwarp();
  // the entries are consumed: their region becomes the (zeroed) measure map
  for (int i = lane; i < ch * (PITCH / 16); i += 32) reinterpret_cast<uint4*>(sm.m)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  // phase B: full measure for the survivors (dense lanes); the corners are compacted in place at
  // the front of the queue (the write position never passes the read position)
  int nc = 0;
  for (int q0 = 0; q0 < nq; q0 += 32) {
    const bool valid = q0 + lane < nq;
    const int off = valid ? sm.queue[q0 + lane] : 3 * PITCH + FAST_COL0 + 3;  // idle lanes measure the first tested pixel
    const int m = fast_arc_measure<PITCH>(sm.img + off);
    const bool corner = valid && m > th;
    if (corner) sm.m[off] = (uint8_t)m;
    const unsigned bm = __ballot_sync(0xffffffffu, corner);
    __syncwarp();
    if (corner) sm.queue[nc + __popc(bm & lt)] = (uint16_t)off;
    nc += __popc(bm);
  }
  __syncwarp();
  // phase C: 3x3 NMS of the corners.  Neighbours outside the tested area hold 0; comparing raw m
  // values equals comparing thresholded scores because every stored m exceeds the pass threshold.
  // Queue, corner list and survivor list all stay in row-major order (one warp, ordered
  // compaction) = cv::FAST's output order.
  int nk = 0;
  for (int q0 = 0; q0 < nc; q0 += 32) {
    const int q = q0 + lane;
    bool keep = false;
    int off = 0;
    if (q < nc) {
      off = sm.queue[q];
      const uint8_t* c = sm.m + off;
      const int n1_ = __vimax3_s32((int)c[-1], (int)c[1], (int)c[-PITCH]);
      const int n2_ = __vimax3_s32((int)c[PITCH], (int)c[-PITCH - 1], (int)c[-PITCH + 1]);
      const int n3_ = __vimax3_s32((int)c[PITCH - 1], (int)c[PITCH + 1], n1_);
      keep = (int)c[0] > max(n2_, n3_);
    }
    const unsigned bm = __ballot_sync(0xffffffffu, keep);
    if (keep) sm.kept[nk + __popc(bm & lt)] = (uint16_t)off;
    nk += __popc(bm);
  }
  __syncwarp();
  return nk;
}

#ifndef FAST_WPC
#define FAST_WPC 1  // cells (= warps) per CTA; the warps of a CTA never talk to each other
#endif
template <int PITCH>
__global__ void __launch_bounds__(FAST_NT * FAST_WPC) k_fast_cells(const OrbCell* __restrict__ cells, OrbLevel0 l0,
                                                    const uint8_t* __restrict__ pyr,
                                                    uint32_t* __restrict__ cand, int* __restrict__ cell_count,
                                                    size_t pyr_frame_bytes, size_t cand_frame_u32, int n_cells,
                                                    int ini_th, int min_th, int rows_max, int t_max,
                                                    const uint8_t* __restrict__ bitmap = nullptr,
                                                    const uint2* __restrict__ cell_bm = nullptr, size_t bm_frame_bytes = 0) {
  extern __shared__ __align__(16) unsigned char fsm_all[];
  FastSmem sm;
  constexpr int HVP = FAST_HVP && !FAST_PAIR ? FAST_HVP : PITCH;
  const int r_img = rows_max * PITCH, r_q = max(rows_max * HVP, (2 * t_max + 31) & ~15);
  const int cell_idx = blockIdx.x * FAST_WPC + (threadIdx.x >> 5);
  if (cell_idx >= n_cells) return;  // whole warp; warps only ever synchronise with themselves
  unsigned char* fsm = fsm_all + (size_t)(threadIdx.x >> 5) * ((2 * r_img + r_q + 16 + 15) & ~15);
  sm.img = fsm;
  sm.kept = reinterpret_cast<uint16_t*>(fsm);
  sm.hv = fsm + r_img;
  sm.queue = reinterpret_cast<uint16_t*>(sm.hv);
  sm.m = sm.hv + r_q;
  sm.eq = reinterpret_cast<uint16_t*>(sm.m);

  // one 32-byte record tells the CTA everything about its cell
  const uint4* crec = reinterpret_cast<const uint4*>(cells + cell_idx);
  const uint4 c0 = __ldg(crec), c1 = __ldg(crec + 1);
  OrbCell cell;
  reinterpret_cast<uint4*>(&cell)[0] = c0;
  reinterpret_cast<uint4*>(&cell)[1] = c1;
  const int frame = blockIdx.y;
  const int cw = cell.cw, ch = cell.ch, a0 = cell.a0;
  const int tid = threadIdx.x & 31;
  // Word copy of the sub-image rows (16 lanes per row), shifted so that sub-image column 0 lands
  // on byte 1 of the shared row: global word grid -> shared word grid by one byte permute of two
  // adjacent global words.  Level-0 cells read the caller's frame.
  const int pitch = cell.level == 0 ? l0.pitch : cell.pitch;
  const uint8_t* base = cell.level == 0
                            ? l0.base + (size_t)frame * l0.frame_stride + (size_t)cell.ini_y * l0.pitch + (cell.ini_x - a0)
                            : pyr + (size_t)frame * pyr_frame_bytes + cell.tile_off;
  {
    const int shift = a0 - 1;  // -1..2 bytes between the two grids
    const int o = shift < 0 ? -1 : 0;
    const uint32_t sel = shift == 0 ? 0x3210u : (shift == 1 ? 0x4321u : (shift == 2 ? 0x5432u : 0x6543u));
    const int nsw = (cw + 4) >> 2, ngw = (a0 + cw + 3) >> 2;
    const int pw = pitch >> 2, y_first = tid >> 4;
    for (int wq = tid & 15; wq < nsw; wq += 16) {
      const int i0 = max(wq + o, 0), d1 = min(wq + o + 1, ngw - 1) - i0;
      const uint32_t* gp = reinterpret_cast<const uint32_t*>(base) + (size_t)y_first * pw + i0;
      uint32_t* sp = reinterpret_cast<uint32_t*>(sm.img) + y_first * (PITCH / 4) + wq + (FAST_COL0 - 1) / 4;
      uint32_t* hp = reinterpret_cast<uint32_t*>(sm.hv) + y_first * (HVP / 4) + wq + (FAST_COL0 - 1) / 4;
      for (int y = y_first; y < ch; y += FAST_NT / 16) {
        const uint32_t v = __byte_perm(__ldg(gp), __ldg(gp + d1), sel);
        sp[0] = v;
        if (!bitmap) hp[0] = (v >> 1) & 0x7f7f7f7fu;  // the halved copy feeds phase A only
        gp += 2 * pw;
        sp += 2 * (PITCH / 4);
        hp += 2 * (HVP / 4);
      }
    }
  }
  const int tw = cw - 6, thh = ch - 6;
  int total = 0, th = ini_th;
  __syncwarp();
  if (tw > 0 && thh > 0) {
    if (bitmap) {
      const uint2 bm = __ldg(cell_bm + cell_idx);
      total = fast_pass<PITCH>(sm, th, tw, thh, ch,
                               reinterpret_cast<const uint32_t*>(bitmap + (size_t)frame * bm_frame_bytes) + (size_t)bm.x * 8, (int)bm.y);
    } else {
      total = fast_pass<PITCH>(sm, th, tw, thh, ch);
    }
    if (total == 0 && min_th < th) {
      // Every corner of the first pass is a corner at the lower threshold too, so the second pass
      // re-measures it (same value) and its NMS sees the complete map.  The queue has overwritten
      // the halved image: rebuild it.
      for (int i = tid; i < ch * (PITCH / 4); i += FAST_NT) {
        const int y = i / (PITCH / 4), xw = i - y * (PITCH / 4);
        reinterpret_cast<uint32_t*>(sm.hv)[y * (HVP / 4) + xw] = (reinterpret_cast<const uint32_t*>(sm.img)[i] >> 1) & 0x7f7f7f7fu;
      }
      __syncwarp();
      th = min_th;
      total = fast_pass<PITCH>(sm, th, tw, thh, ch);
    }
  }
  // write-out: the survivor list is already in row-major order (byte offsets into the pitched
  // map are monotone in (y, x))
  total = min(total, (int)cell.cand_cap);
  uint32_t* out = cand + (size_t)frame * cand_frame_u32 + cell.cand_slot_off;
  for (int i = tid; i < total; i += FAST_NT) {
    const int off = sm.kept[i];
    const int y = off / PITCH, x = off - y * PITCH - FAST_COL0;
    out[i] = (uint32_t)(x + cell.off_x) | (uint32_t)(y + cell.off_y) << 12 | ((uint32_t)sm.m[off] - 1u) << 24;
  }
  if (tid == 0) cell_count[(size_t)frame * n_cells + cell_idx] = total;
}

// ------------------------------------------------------------------------------------------
// K3a (split form) the rejection test of the FAST stage on its own, one WARP per band (OrbBand), no shared memory: the
// phase A of k_fast_bands - a lane owns 8 adjacent pixels of every row and keeps the last seven rows (halved bytes, own
// words + neighbour words by shuffle) in registers - writing one byte of pass bits per lane and tested row into a bitmap
// (32 bytes per band row, coalesced).  k_fast_cells, launched behind it with the bitmap, then skips its own phase A (11
// shared loads per four pixels, the halved tile copy, the entry expansion) and starts at the arc measure.  Threshold
// iniThFAST only: a cell that keeps nothing reruns at minThFAST inside k_fast_cells with its own phase A (:813-817).
__global__ void __launch_bounds__(128) k_fast_prefilter(const OrbBand* __restrict__ bands, const unsigned* __restrict__ band_bm,
                                                        int n_bands, OrbLevel0 l0, const uint8_t* __restrict__ pyr,
                                                        uint8_t* __restrict__ bitmap, size_t pyr_frame_bytes,
                                                        size_t bm_frame_bytes, int ini_th) {
  const int lane = threadIdx.x & 31;
  const int bi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (bi >= n_bands) return;  // whole warp
  const unsigned FULL = 0xffffffffu;
  OrbBand band;
  {
    const uint4* br = reinterpret_cast<const uint4*>(bands + bi);
    reinterpret_cast<uint4*>(&band)[0] = __ldg(br);
    reinterpret_cast<uint4*>(&band)[1] = __ldg(br + 1);
  }
  const int nt = band.nt;
  if (nt <= 0) return;
  const int frame = blockIdx.y;
  const int xb = band.xb, x0 = band.x0, x1 = band.x1, y_first = band.y_first;
  const int pitch = band.level == 0 ? l0.pitch : band.pitch;
  const uint8_t* const src = band.level == 0
                                 ? l0.base + (size_t)frame * l0.frame_stride + (size_t)y_first * l0.pitch + xb
                                 : pyr + (size_t)frame * pyr_frame_bytes + band.src_off;
  const int pitch_w = pitch >> 2;
  const int R_load = nt + 6;
  const int nbytes = x1 + 3 - xb;
  const bool ldA = 8 * lane < nbytes, ldB = 8 * lane + 4 < nbytes;
  const uint32_t K = 0x01010101u * (uint32_t)(128 - min(ini_th >> 1, 127));
  uint32_t maskA = 0, maskB = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = xb + 8 * lane + k;
    if (x >= x0 && x < x1) {
      if (k < 4) maskA |= 0x80u << (8 * k);
      else maskB |= 0x80u << (8 * (k - 4));
    }
  }
  uint8_t* out = bitmap + (size_t)frame * bm_frame_bytes + (size_t)__ldg(band_bm + bi) * 32 + lane;

  const uint32_t* gp = reinterpret_cast<const uint32_t*>(src) + 2 * lane;
  int s_issue = 0;
  uint32_t pA0 = 0, pB0 = 0, pA1 = 0, pB1 = 0;
  auto issue = [&](uint32_t& a, uint32_t& b) {
    a = 0;
    b = 0;
    if (s_issue < R_load) {
      if (ldA) a = __ldg(gp);
      if (ldB) b = __ldg(gp + 1);
    }
    gp += pitch_w;
    ++s_issue;
  };
  issue(pA0, pB0);
  issue(pA1, pB1);
  uint32_t hA[7], hB[7], hL[7], hR[7];
  auto take_row = [&](int k) {
    const uint32_t a = pA0, b = pB0;
    pA0 = pA1;
    pB0 = pB1;
    issue(pA1, pB1);
    hA[k] = (a >> 1) & 0x7f7f7f7fu;
    hB[k] = (b >> 1) & 0x7f7f7f7fu;
    hL[k] = __shfl_up_sync(FULL, hB[k], 1);
    hR[k] = __shfl_down_sync(FULL, hA[k], 1);
  };
#pragma unroll
  for (int k = 0; k < 6; ++k) take_row(k);
  for (int s0 = 6; s0 < R_load; s0 += 7) {
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      if (s0 + u < R_load) {
        const int k = (6 + u) % 7;
        take_row(k);
        const int c = (k + 4) % 7, p2 = (k + 6) % 7, m2 = (k + 2) % 7, m3 = (k + 1) % 7;
        const uint32_t cA = hA[c], cB = hB[c];
        const uint32_t c_ab3 = __byte_perm(cA, cB, 0x6543), c_ab1 = __byte_perm(cA, cB, 0x4321);
        const uint32_t p_ab = __byte_perm(hA[p2], hB[p2], 0x5432), m_ab = __byte_perm(hA[m2], hB[m2], 0x5432);
        uint32_t passA, passB;
        {
          const uint32_t f0 = __vabsdiffu4(hA[k], cA) + K, f8 = __vabsdiffu4(hA[m3], cA) + K;
          const uint32_t f4 = __vabsdiffu4(c_ab3, cA) + K;
          const uint32_t f12 = __vabsdiffu4(__byte_perm(hL[c], cA, 0x4321), cA) + K;
          const uint32_t f2 = __vabsdiffu4(p_ab, cA) + K;
          const uint32_t f14 = __vabsdiffu4(__byte_perm(hL[p2], hA[p2], 0x5432), cA) + K;
          const uint32_t f6 = __vabsdiffu4(m_ab, cA) + K;
          const uint32_t f10 = __vabsdiffu4(__byte_perm(hL[m2], hA[m2], 0x5432), cA) + K;
          passA = (f0 | f8) & (f4 | f12) & (f2 | f10) & (f6 | f14) & maskA;
        }
        {
          const uint32_t f0 = __vabsdiffu4(hB[k], cB) + K, f8 = __vabsdiffu4(hB[m3], cB) + K;
          const uint32_t f4 = __vabsdiffu4(__byte_perm(cB, hR[c], 0x6543), cB) + K;
          const uint32_t f12 = __vabsdiffu4(c_ab1, cB) + K;
          const uint32_t f2 = __vabsdiffu4(__byte_perm(hB[p2], hR[p2], 0x5432), cB) + K;
          const uint32_t f14 = __vabsdiffu4(p_ab, cB) + K;
          const uint32_t f6 = __vabsdiffu4(__byte_perm(hB[m2], hR[m2], 0x5432), cB) + K;
          const uint32_t f10 = __vabsdiffu4(m_ab, cB) + K;
          passB = (f0 | f8) & (f4 | f12) & (f2 | f10) & (f6 | f14) & maskB;
        }
        *out = (uint8_t)((((passA >> 7) * 0x10204080u) >> 28) | ((((passB >> 7) * 0x10204080u) >> 28) << 4));
        out += 32;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3 (streaming form) per-cell FAST-9/16 with NMS and the iniThFAST -> minThFAST fallback (:790-830), one WARP per
// band (OrbBand: up to 8 whole cells of one cell row, <= 245 tested pixels wide).  Same three phases as k_fast_cells,
// but the warp STREAMS down the band's rows instead of staging a cell:
//   phase A  a lane owns 8 adjacent pixels (two words) of every row and keeps the last seven rows (halved bytes, own
//            words + the two neighbour words fetched by shuffle) in registers: the packed opposite-pair rejection test
//            runs entirely on registers - no shared-memory loads.  A lane whose 8 pixels hold a survivor pushes ONE
//            entry (ring rows, first column, 8-bit survivor mask) behind the entries of the lanes before it (ballot).
//   phase B  after every block of seven rows the entries are expanded, 32 at a time, into a staging list of
//            candidates (row-major) that is measured 32 candidates at a time: exact arc measure from a 13-row ring of
//            the raw rows in shared memory (mirror rows above and below the ring keep the 16 ring offsets compile-time
//            constants); corners write their measure into a 9-row measure ring and enter a corner FIFO;
//   phase C  3x3 NMS of the corners whose three measure rows are complete, in FIFO (= row-major) order, neighbours
//            across a cell boundary counting as 0 (NMS is per cell in the reference); survivors pass through a small
//            FIFO that a per-cell ordered split (one ballot per cell) empties into the cells' candidate slots.
// Every buffer is bounded by the row schedule (sized for "every pixel is a candidate / a corner"), the lists stay in
// cv::FAST's output order, and a band whose cells kept nothing at iniThFAST streams once more at minThFAST with only
// those cells unmasked (:813-817).
#define FB_BLOCK 7      // rows per block = the register window's period
#define FB_RP 264       // raw ring: bytes per row (256 + 8: rows two banks apart)
#define FB_RSLOTS 13    // block + 6 halo rows; slot s sits at physical row s + 3, rows 0..2 mirror the last three slots,
#define FB_RROWS 19     // rows 16..18 the first three
#define FB_MP 272       // measure ring: bytes per row
#define FB_MSLOTS 9     // block + the pending row and the row above it; slot s at physical row s + 1, row 0 mirrors the
#define FB_MROWS 11     // last slot, row 10 the first
#define FB_ENT_CAP 224  // entries of a block (7 rows x 32 lanes)
#define FB_STG_CAP 320  // staged candidates: < 32 left over + 32 entries x 8
#define FB_CF_CAP 1024  // corner FIFO (ring)
#define FB_KF_CAP 64    // NMS survivors waiting for the per-cell split
#define FB_RAW_BYTES ((FB_RROWS * FB_RP + 15) & ~15)
#define FB_SMEM (FB_RAW_BYTES + FB_MROWS * FB_MP + 4 * FB_ENT_CAP + 2 * FB_STG_CAP + 2 * FB_CF_CAP + 4 * FB_KF_CAP)

#ifndef FB_MINB
#define FB_MINB 18  // caps the kernel at 113 registers (it takes 128 uncapped for no gain): 18 warps per SM, the shared-memory limit
#endif
__global__ void __launch_bounds__(32, FB_MINB) k_fast_bands(const OrbBand* __restrict__ bands, OrbLevel0 l0,
                                                   const uint8_t* __restrict__ pyr, uint32_t* __restrict__ cand,
                                                   int* __restrict__ cell_count, size_t pyr_frame_bytes,
                                                   size_t cand_frame_u32, int n_cells_frame, int ini_th, int min_th) {
  __shared__ __align__(16) unsigned char fb[FB_SMEM];
  uint8_t* const raw = fb;
  uint8_t* const mr = fb + FB_RAW_BYTES;
  uint32_t* const ebuf = reinterpret_cast<uint32_t*>(mr + FB_MROWS * FB_MP);
  uint32_t* const kf = ebuf + FB_ENT_CAP;
  uint16_t* const stg = reinterpret_cast<uint16_t*>(kf + FB_KF_CAP);
  uint16_t* const cf = stg + FB_STG_CAP;
  const int lane = threadIdx.x;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned FULL = 0xffffffffu;

  OrbBand band;
  {
    const uint4* br = reinterpret_cast<const uint4*>(bands + blockIdx.x);
    reinterpret_cast<uint4*>(&band)[0] = __ldg(br);
    reinterpret_cast<uint4*>(&band)[1] = __ldg(br + 1);
  }
  const int frame = blockIdx.y;
  const int ncell = band.n_cells, nt = band.nt, cap = band.cand_cap;
  int* const cc = cell_count + (size_t)frame * n_cells_frame + band.cell_idx0;
  if (nt <= 0) {  // cells whose sub-image is too low to hold a tested pixel: cv::FAST finds nothing
    if (lane < ncell) cc[lane] = 0;
    return;
  }
  uint32_t* const out = cand + (size_t)frame * cand_frame_u32 + band.cand_slot_off;
  const int xb = band.xb, x0 = band.x0, x1 = band.x1, w_cell = band.w_cell, y_first = band.y_first;
  const uint32_t inv = (65536u + (uint32_t)w_cell - 1u) / (uint32_t)w_cell;  // (d * inv) >> 16 == d / w_cell for d < 256
  const int pitch = band.level == 0 ? l0.pitch : band.pitch;
  const uint8_t* const src = band.level == 0
                                 ? l0.base + (size_t)frame * l0.frame_stride + (size_t)y_first * l0.pitch + xb
                                 : pyr + (size_t)frame * pyr_frame_bytes + band.src_off;
  const int pitch_w = pitch >> 2;
  const int R_load = nt + 6;
  const int nbytes = x1 + 3 - xb;  // bytes of a row anybody reads
  const bool ldA = 8 * lane < nbytes, ldB = 8 * lane + 4 < nbytes;

  int cnt_cell = 0;                       // lane c < ncell: candidates kept in cell c
  unsigned active = (1u << ncell) - 1u;   // cells tested in this pass
  for (int pass = 0; pass < 2; ++pass) {
    const int th = pass ? min_th : ini_th;
    const uint32_t K = 0x01010101u * (uint32_t)(128 - min(th >> 1, 127));
    uint32_t maskA = 0, maskB = 0;  // bit 7 of every byte this lane tests
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int x = xb + 8 * lane + k;
      bool t = x >= x0 && x < x1;
      if (t) t = (active >> ((uint32_t)((x - x0) * inv) >> 16)) & 1u;
      if (t) {
        if (k < 4) maskA |= 0x80u << (8 * k);
        else maskB |= 0x80u << (8 * (k - 4));
      }
    }
    // the measure ring starts all zero (rows outside the tested area must read 0)
    for (int i = lane; i < FB_MROWS * FB_MP / 16; i += 32) reinterpret_cast<uint4*>(mr)[i] = make_uint4(0, 0, 0, 0);

    int n_e = 0;                // entries in ebuf
    int chead = 0, ctail = 0;   // corner FIFO
    int khead = 0, ktail = 0;   // survivor FIFO
    int sr = 0;                 // raw ring slot of the next row to store
    int er_c = 6, em_c = 1;     // physical ring rows (raw, measure) of the next row to test (local row 3 -> slots 3, 0)
    int t_tested = 2;

    // software pipeline of the global loads: two rows in flight
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(src) + 2 * lane;
    int s_issue = 0;
    uint32_t pA0 = 0, pB0 = 0, pA1 = 0, pB1 = 0;
    auto issue = [&](uint32_t& a, uint32_t& b) {
      a = 0;
      b = 0;
      if (s_issue < R_load) {
        if (ldA) a = __ldg(gp);
        if (ldB) b = __ldg(gp + 1);
      }
      gp += pitch_w;
      ++s_issue;
    };
    issue(pA0, pB0);
    issue(pA1, pB1);

    uint32_t hA[7], hB[7], hL[7], hR[7];  // window: halved bytes of the last seven rows (own two words, neighbour words)
    auto take_row = [&](int k) {
      const uint32_t a = pA0, b = pB0;
      pA0 = pA1;
      pB0 = pB1;
      issue(pA1, pB1);
      hA[k] = (a >> 1) & 0x7f7f7f7fu;
      hB[k] = (b >> 1) & 0x7f7f7f7fu;
      hL[k] = __shfl_up_sync(FULL, hB[k], 1);
      hR[k] = __shfl_down_sync(FULL, hA[k], 1);
      uint2* dst = reinterpret_cast<uint2*>(raw + (sr + 3) * FB_RP) + lane;
      *dst = make_uint2(a, b);
      if (sr >= FB_RSLOTS - 3) dst[-FB_RSLOTS * (FB_RP / 8)] = make_uint2(a, b);
      if (sr < 3) dst[FB_RSLOTS * (FB_RP / 8)] = make_uint2(a, b);
      sr = sr == FB_RSLOTS - 1 ? 0 : sr + 1;
    };

    // ---- phase B / C pieces ------------------------------------------------------------------
    // rows are recovered from ring slots: a measure-ring row em belongs to the row `age` rows before the last tested
    auto row_of = [&](int em) {
      int age = (em_c == 1 ? FB_MSLOTS : em_c - 1) - em;  // em_c - 1 (cyclic) = ring row of the last tested row
      if (age < 0) age += FB_MSLOTS;
      return t_tested - age;
    };
    auto emit_iter = [&]() {  // per-cell ordered split of up to 32 survivors
      const int ks = ktail - khead;
      const bool have = lane < ks;
      const uint32_t key = kf[(khead + lane) & (FB_KF_CAP - 1)];
      const int j = have ? (int)(((key & 0xfffu) + 16u - (uint32_t)x0) * inv >> 16) : -1;
      int rank = 0, add = 0;
      for (int c = 0; c < ncell; ++c) {
        const unsigned bm = __ballot_sync(FULL, j == c);
        if (j == c) rank = __popc(bm & lt);
        if (lane == c) add = __popc(bm);
      }
      const int idx = __shfl_sync(FULL, cnt_cell, j & 31) + rank;
      if (have && idx < cap) out[j * cap + idx] = key;
      cnt_cell += add;
      khead += min(ks, 32);
    };
    auto nms_iter = [&](int t_lim) -> int {  // NMS of the FIFO's leading corners with row <= t_lim; returns how many
      const int csize = ctail - chead;
      const bool have = lane < csize;
      const int ent = cf[(chead + (have ? lane : 0)) & (FB_CF_CAP - 1)];
      const int x = ent & 255, em = (ent >> 8) & 15;
      const int row = row_of(em);
      const bool elig = have && row <= t_lim;
      const int n_take = __popc(__ballot_sync(FULL, elig));  // the FIFO is in row order: the eligible ones lead
      bool keep = false;
      int m = 0;
      const int lx = xb + x;
      if (elig) {
        const uint8_t* c = mr + em * FB_MP + x;
        m = c[0];
        const int v = max((int)c[-FB_MP], (int)c[FB_MP]);
        const int l = __vimax3_s32((int)c[-1], (int)c[-FB_MP - 1], (int)c[FB_MP - 1]);
        const int r = __vimax3_s32((int)c[1], (int)c[-FB_MP + 1], (int)c[FB_MP + 1]);
        const int cx0 = x0 + (int)((uint32_t)((lx - x0) * inv) >> 16) * w_cell, cx1 = min(cx0 + w_cell, x1);
        keep = m > v && (lx == cx0 || m > l) && (lx + 1 == cx1 || m > r);
      }
      const unsigned kb = __ballot_sync(FULL, keep);
      if (keep)
        kf[(ktail + __popc(kb & lt)) & (FB_KF_CAP - 1)] =
            (uint32_t)(lx - 16) | (uint32_t)(y_first + row - 16) << 12 | ((uint32_t)m - 1u) << 24;
      ktail += __popc(kb);
      chead += n_take;
      __syncwarp();
      if (ktail - khead >= 32) {
        emit_iter();
        __syncwarp();
      }
      return n_take;
    };
    auto drain = [&]() {  // expand and measure every entry of the block, then NMS every corner above the last tested row
      __syncwarp();
      int e0 = 0, s_head = 0, s_tail = 0;
      while (true) {
        if (s_tail - s_head < 32 && e0 < n_e) {
          // the (< 32) staged candidates move to the front, 32 more entries expand behind them
          const int left = s_tail - s_head;
          const int keepv = lane < left ? stg[s_head + lane] : 0;
          const uint32_t e = e0 + lane < n_e ? ebuf[e0 + lane] : 0u;
          __syncwarp();
          if (lane < left) stg[lane] = (uint16_t)keepv;
          const uint32_t nib = e >> 16;
          const int cnt = __popc(nib);
          int incl = cnt;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
          }
          uint16_t* q = stg + left + incl - cnt;
          const uint32_t base = e & 0xffffu;
#pragma unroll
          for (int b = 0; b < 8; ++b)
            if (nib >> b & 1) *q++ = (uint16_t)(base + b);
          s_head = 0;
          s_tail = left + __shfl_sync(FULL, incl, 31);
          e0 += 32;
          __syncwarp();
          continue;
        }
        const int navail = s_tail - s_head;
        if (navail == 0) break;
        const bool valid = lane < navail;
        const int ent = stg[s_head + (valid ? lane : 0)];
        const int x = ent & 255, em = (ent >> 8) & 15, er = ent >> 12;
        const int m = fast_arc_measure<FB_RP>(raw + er * FB_RP + x);
        const bool corner = valid && m > th;
        if (corner) {
          mr[em * FB_MP + x] = (uint8_t)m;
          if (em == FB_MSLOTS) mr[x] = (uint8_t)m;
          if (em == 1) mr[(FB_MROWS - 1) * FB_MP + x] = (uint8_t)m;
        }
        const unsigned bm = __ballot_sync(FULL, corner);
        if (corner) cf[(ctail + __popc(bm & lt)) & (FB_CF_CAP - 1)] = (uint16_t)ent;
        ctail += __popc(bm);
        s_head += min(navail, 32);
        if (ctail - chead > FB_CF_CAP - 64) {
          // keep the FIFO inside its ring: every row above this iteration's first candidate is completely measured,
          // so at most two rows (490 corners) + this iteration's are not yet eligible - the rest can go now
          __syncwarp();
          const int t_first = row_of((__shfl_sync(FULL, ent, 0) >> 8) & 15);
          while (ctail - chead > FB_CF_CAP / 2 + 64 && nms_iter(t_first - 2) > 0) {
          }
        }
      }
      n_e = 0;
      __syncwarp();
      while (nms_iter(t_tested - 1) == 32) {
      }
      // the next block's measure rows: zero them (their slots held rows that are dead now)
      for (int i = lane; i < FB_BLOCK * (FB_MP / 16); i += 32) {
        const int r = i / (FB_MP / 16), c = i - r * (FB_MP / 16);
        int sl = em_c - 1 + r;
        if (sl >= FB_MSLOTS) sl -= FB_MSLOTS;
        reinterpret_cast<uint4*>(mr + (sl + 1) * FB_MP)[c] = make_uint4(0, 0, 0, 0);
        if (sl == FB_MSLOTS - 1) reinterpret_cast<uint4*>(mr)[c] = make_uint4(0, 0, 0, 0);
        if (sl == 0) reinterpret_cast<uint4*>(mr + (FB_MROWS - 1) * FB_MP)[c] = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
    };

    // ---- the stream --------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < 6; ++k) take_row(k);
    for (int s0 = 6; s0 < R_load; s0 += 7) {
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        if (s0 + u < R_load) {
          const int k = (6 + u) % 7;  // window slot of the new row; the tested row sits three slots back
          take_row(k);
          const int c = (k + 4) % 7, p2 = (k + 6) % 7, m2 = (k + 2) % 7, m3 = (k + 1) % 7;
          // Packed rejection test (see k_fast_cells): a corner has |p - c| > th for one pixel of every opposite
          // pair of the eight even ring positions.
          const uint32_t cA = hA[c], cB = hB[c];
          const uint32_t c_ab3 = __byte_perm(cA, cB, 0x6543), c_ab1 = __byte_perm(cA, cB, 0x4321);
          const uint32_t p_ab = __byte_perm(hA[p2], hB[p2], 0x5432), m_ab = __byte_perm(hA[m2], hB[m2], 0x5432);
          uint32_t passA, passB;
          {
            const uint32_t f0 = __vabsdiffu4(hA[k], cA) + K, f8 = __vabsdiffu4(hA[m3], cA) + K;
            const uint32_t f4 = __vabsdiffu4(c_ab3, cA) + K;
            const uint32_t f12 = __vabsdiffu4(__byte_perm(hL[c], cA, 0x4321), cA) + K;
            const uint32_t f2 = __vabsdiffu4(p_ab, cA) + K;
            const uint32_t f14 = __vabsdiffu4(__byte_perm(hL[p2], hA[p2], 0x5432), cA) + K;
            const uint32_t f6 = __vabsdiffu4(m_ab, cA) + K;
            const uint32_t f10 = __vabsdiffu4(__byte_perm(hL[m2], hA[m2], 0x5432), cA) + K;
            passA = (f0 | f8) & (f4 | f12) & (f2 | f10) & (f6 | f14) & maskA;
          }
          {
            const uint32_t f0 = __vabsdiffu4(hB[k], cB) + K, f8 = __vabsdiffu4(hB[m3], cB) + K;
            const uint32_t f4 = __vabsdiffu4(__byte_perm(cB, hR[c], 0x6543), cB) + K;
            const uint32_t f12 = __vabsdiffu4(c_ab1, cB) + K;
            const uint32_t f2 = __vabsdiffu4(__byte_perm(hB[p2], hR[p2], 0x5432), cB) + K;
            const uint32_t f14 = __vabsdiffu4(p_ab, cB) + K;
            const uint32_t f6 = __vabsdiffu4(__byte_perm(hB[m2], hR[m2], 0x5432), cB) + K;
            const uint32_t f10 = __vabsdiffu4(m_ab, cB) + K;
            passB = (f0 | f8) & (f4 | f12) & (f2 | f10) & (f6 | f14) & maskB;
          }
          // one entry per lane with a survivor: ring rows, first column, 8-bit mask (column order)
          const unsigned eb = __ballot_sync(FULL, (passA | passB) != 0);
          if (passA | passB) {
            const uint32_t nib = (((passA >> 7) * 0x10204080u) >> 28) | ((((passB >> 7) * 0x10204080u) >> 28) << 4);
            ebuf[n_e + __popc(eb & lt)] = nib << 16 | (uint32_t)(er_c << 12 | em_c << 8 | 8 * lane);
          }
          n_e += __popc(eb);
          er_c = er_c == FB_RSLOTS + 2 ? 3 : er_c + 1;
          em_c = em_c == FB_MSLOTS ? 1 : em_c + 1;
          ++t_tested;
        }
      }
      drain();
    }
    // the last tested row: the row below it is all zero in the measure ring
    while (nms_iter(t_tested) == 32) {
    }
    if (ktail - khead > 0) emit_iter();
    __syncwarp();

    if (pass == 0) {
      const unsigned empty = __ballot_sync(FULL, lane < ncell && cnt_cell == 0);
      if (empty == 0 || min_th >= ini_th) break;
      active = empty;
    }
  }
  if (lane < ncell) cc[lane] = min(cnt_cell, cap);
}

// ------------------------------------------------------------------------------------------
// K4 DistributeOctTree: one CTA per (level, frame); policy core in octree_core.h.
// Gather of the per-cell candidate slots into one key array: a thread per cell, eight loads in
// flight per thread (the slots sit in HBM: a load -> store loop would serialise the latency).
__device__ __forceinline__ void gather_cells(const uint32_t* __restrict__ cslots, int cand_cap, int n_cells,
                                             const int* cnt, const int* off, uint32_t* dst) {
  for (int c = threadIdx.x; c < n_cells; c += blockDim.x) {
    const int cn = cnt[c], o = off[c];
    const uint32_t* src = cslots + (size_t)c * cand_cap;
    for (int k = 0; k < cn; k += 8) {
      uint32_t r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = (k + u < cn) ? __ldg(src + k + u) : 0u;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k + u < cn) dst[o + k + u] = r[u];
    }
  }
}

__global__ void __launch_bounds__(256) k_octree(const OrbGeom* __restrict__ g, const uint32_t* __restrict__ cand,
                                                const int* __restrict__ cell_count, uint32_t* __restrict__ keys,
                                                uint16_t* __restrict__ knode, uint32_t* __restrict__ sel,
                                                int* __restrict__ sel_count, int smem_keys) {
  // level-major launch order: the big level-0 problems start first and the small top levels fill the tail
  const int level = blockIdx.y, frame = blockIdx.x;
  const OrbLevelGeom& L = g->lv[level];
  OtScratch s;
  const int scratch_bytes = ot_layout(s, g->ot_cap, g->ot_scan_cap, 256);
  int* sa = OT_INTS(s, s.a);
  int* sb = OT_INTS(s, s.b);
  int* vars = OT_INTS(s, s.vars);
  OT_MARK(20);

  // gather this level's candidates in vToDistributeKeys order: cells row-major, in-cell order
  const int* cc = cell_count + (size_t)frame * g->n_cells + L.cell_base;
  OT_FOR(i, L.n_cells) sa[i] = cc[i];
  OT_SYNC();
  ot_exclusive_scan(sa, sb, L.n_cells, &vars[OT_V_TOTAL], OT_INTS(s, s.part));
  const int M = vars[OT_V_TOTAL];
  OT_MARK(21);
  const uint32_t* cslots = cand + (size_t)frame * g->cand_frame_u32 + L.cand_off;
  uint32_t* out = sel + (size_t)frame * g->kp_cap_frame + L.sel_off;
  int n;
  // The policy replay re-labels every key once per round: keys and labels stay in shared memory
  // whenever the level's candidates fit (camera-like frames), else in the HBM workspace.
  if (M <= smem_keys) {
    uint32_t* skeys = reinterpret_cast<uint32_t*>(ot_smem + scratch_bytes);
    uint16_t* sknode = reinterpret_cast<uint16_t*>(ot_smem + scratch_bytes + 4 * smem_keys);
    gather_cells(cslots, L.cand_cap, L.n_cells, sa, sb, skeys);
    __syncthreads();
    OT_MARK(22);
    n = ot_distribute(skeys, sknode, M, L.roots, L.quota, s, out);
  } else {
    uint32_t* fkeys = keys + (size_t)frame * g->key_frame_u32 + L.key_off;
    uint16_t* fknode = knode + (size_t)frame * g->key_frame_u32 + L.key_off;
    gather_cells(cslots, L.cand_cap, L.n_cells, sa, sb, fkeys);
    __syncthreads();
    n = ot_distribute(fkeys, fknode, M, L.roots, L.quota, s, out);
  }
  if (threadIdx.x == 0) sel_count[frame * g->nlevels + level] = n;
  OT_MARK(23);
}

// ------------------------------------------------------------------------------------------
// K6 GaussianBlur 7x7 sigma 2, OpenCV 4.x fixed-point path: taps {18,34,48,56,48,34,18}/256 per
// pass, H pass u8 -> 8.8 (16 bit), V pass -> 16.16, (v + 32768) >> 16, BORDER_REFLECT_101.
// One WARP per 128 x 32 output tile, no shared memory, no barriers: a lane owns four adjacent
// columns and walks down the rows; the H pass of each new row (dp4a on packed bytes) enters a
// 7-row register window from which the V pass is taken.  Rows are fully unrolled (the window
// rotation is register renaming) and branch-free, so the loads of many rows are in flight at once.
// Rows mirror by index.  Columns: a lane needs image columns x-4..x+7 as three words (a, b, c).
// Warps that lie fully inside the image take them as three aligned loads; warps touching the
// left/right edge run the EDGE variant, where every lane loads four aligned words around its
// window and builds a, b, c with lane-constant word picks + byte permutes (the mirrored source
// columns of each word lie within <= 4 adjacent columns).
// VARIANT 0: warp fully inside the image; 1: warp at the left edge (only lane 0 mirrors: columns
// -4..-1 are columns 4..1, one byte permute of its own words); 2: generic edge (right edge, or
// levels narrower than a tile).
template <int VARIANT>
__device__ __forceinline__ void blur_rows(const uint32_t* __restrict__ src, int spw, int y0, int h, const int (&ia)[3],
                                          const uint32_t (&esel)[3], uint8_t* __restrict__ dst, int bpitch, int x,
                                          bool col_ok) {
  const uint32_t K0123 = 18u | 34u << 8 | 48u << 16 | 56u << 24, K456 = 48u | 34u << 8 | 18u << 16;
  // 7-row register window, rows processed in blocks of 7 so that the slot of row r is r % 7 at
  // compile time (ORB_BLUR_TH + 6 is a multiple of 7); the block loop stays rolled to keep the
  // code of all three variants inside the instruction cache.
  uint32_t win[7][4];
  for (int rb = 0; rb < ORB_BLUR_TH + 6; rb += 7) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int r = rb + j;
      // H pass of image row y0 - 3 + r (mirrored at the top/bottom edge)
      const int iy = reflect101(min(y0 - 3 + r, h + 2), h);
      const uint32_t* row = src + (size_t)iy * spw;
      uint32_t a, b, c;
      if (VARIANT == 0) {
        a = __ldg(row); b = __ldg(row + 1); c = __ldg(row + 2);
      } else if (VARIANT == 1) {
        // lane 0 (x == 0) was pointed at word 0: its q0, q1 are columns 0..7
        const uint32_t q0 = __ldg(row), q1 = __ldg(row + 1), q2 = __ldg(row + 2);
        const bool is_left = x == 0;
        a = is_left ? __byte_perm(q0, q1, 0x1234) : q0;
        b = is_left ? q0 : q1;
        c = is_left ? q1 : q2;
      } else {
        const uint32_t q0 = __ldg(row), q1 = __ldg(row + 1), q2 = __ldg(row + 2), q3 = __ldg(row + 3);
        auto lo = [&](int i) { return i == 0 ? q0 : (i == 1 ? q1 : q2); };
        auto hi = [&](int i) { return i == 0 ? q1 : (i == 1 ? q2 : q3); };
        a = __byte_perm(lo(ia[0]), hi(ia[0]), esel[0]);
        b = __byte_perm(lo(ia[1]), hi(ia[1]), esel[1]);
        c = __byte_perm(lo(ia[2]), hi(ia[2]), esel[2]);
      }
      win[j][0] = __dp4a(__byte_perm(a, b, 0x4321), K0123, __dp4a(__byte_perm(b, c, 0x4321), K456, 0u));
      win[j][1] = __dp4a(__byte_perm(a, b, 0x5432), K0123, __dp4a(__byte_perm(b, c, 0x5432), K456, 0u));
      win[j][2] = __dp4a(__byte_perm(a, b, 0x6543), K0123, __dp4a(__byte_perm(b, c, 0x6543), K456, 0u));
      win[j][3] = __dp4a(b, K0123, __dp4a(c, K456, 0u));
      if (r >= 6) {
        // rows r-6 .. r sit in slots (j+1)%7 .. (j+7)%7
        const int y = y0 + r - 6;
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t v = 18u * (win[(j + 1) % 7][i] + win[j][i]) + 34u * (win[(j + 2) % 7][i] + win[(j + 6) % 7][i]) +
                             48u * (win[(j + 3) % 7][i] + win[(j + 5) % 7][i]) + 56u * win[(j + 4) % 7][i] + 32768u;
          o[i] = v >> 16;
        }
        if (col_ok && y < h)
          reinterpret_cast<uint32_t*>(dst + (size_t)y * bpitch)[x >> 2] = o[0] | o[1] << 8 | o[2] << 16 | o[3] << 24;
      }
    }
  }
}

#ifndef BLUR_MINB
#define BLUR_MINB 1
#endif
__global__ void __launch_bounds__(128, BLUR_MINB) k_blur(const OrbGeom* __restrict__ g, OrbLevel0 l0, const uint8_t* __restrict__ pyr,
                                              uint8_t* __restrict__ blur) {
  const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (tile >= g->n_blur_tiles) return;
  const int lane = threadIdx.x & 31;
  int level = 0;
  while (level + 1 < g->nlevels && tile >= g->lv[level + 1].blur_tile_base) ++level;
  const OrbLevelGeom& L = g->lv[level];
  const int t = tile - L.blur_tile_base;
  const int ty = t / L.blur_tiles_x, tx = t - ty * L.blur_tiles_x;
  const int x = tx * ORB_BLUR_TW + 4 * lane, y0 = ty * ORB_BLUR_TH;
  const int frame = blockIdx.y;
  const LevelView S = level_view(g, l0, pyr, level, frame);
  const int w = L.w, h = L.h;
  uint8_t* dst = blur + (size_t)frame * g->blur_frame_bytes + L.blur_off;
  const bool col_ok = x < L.bpitch;
  const int spw = S.pitch >> 2;
  int ia[3] = {0, 1, 2};
  uint32_t esel[3] = {0x3210u, 0x3210u, 0x3210u};
  const bool right_ok = tx * ORB_BLUR_TW + ORB_BLUR_TW + 3 < w;  // warp-uniform: no lane reaches past the right edge
  if (right_ok) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(S.base) + (max(x - 4, 0) >> 2);
    if (tx > 0) blur_rows<0>(src, spw, y0, h, ia, esel, dst, L.bpitch, x, col_ok);
    else blur_rows<1>(src, spw, y0, h, ia, esel, dst, L.bpitch, x, col_ok);
    return;
  }
  // lane constants of the edge variant: four words from `wb` cover every (mirrored) source column
  int col[12], lo = 1 << 30;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    col[k] = reflect101(min(max(x - 4 + k, -18), w + 18), w);
    lo = min(lo, col[k]);
  }
  const int wb = min(lo >> 2, max(0, spw - 4));
#pragma unroll
  for (int gi = 0; gi < 3; ++gi) {
    int mo = 1 << 30;
#pragma unroll
    for (int k = 0; k < 4; ++k) mo = min(mo, col[4 * gi + k] - 4 * wb);
    ia[gi] = min(mo >> 2, 2);
    uint32_t sv = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) sv |= (uint32_t)((col[4 * gi + k] - 4 * wb - 4 * ia[gi]) & 7) << (4 * k);
    esel[gi] = sv;
  }
  blur_rows<2>(reinterpret_cast<const uint32_t*>(S.base) + wb, spw, y0, h, ia, esel, dst, L.bpitch, x, col_ok);
}

// ------------------------------------------------------------------------------------------
// K5 + K7: IC_Angle orientation (:77-104) and rotated-BRIEF descriptor (:108-147), one warp
// per keypoint.  Also the epilogue of operator() (:1060-1106): level order concatenation,
// pt *= mvScaleFactor[level], size, octave.
__constant__ int c_umax[ORB_HALF_PATCH + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
__device__ __align__(16) const signed char d_pattern[ORB_PATTERN_INTS] = ORB_PATTERN_INIT;

// cv::fastAtan2 (OpenCV core, scalar path): float32, evaluated left to right, no FMA.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
  const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
  const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// glibc >= 2.28 sinf/cosf (ARM optimized-routines algorithm): double-precision reduction and
// polynomial, rounded once to float.  Valid for |y| < 120; the caller's angles lie in [0, 2pi).
__device__ __forceinline__ double sincos_poly(double x, double x2, bool neg_cos, int n) {
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
         C4 = 0x1.99343027bf8c3p-16;
  if (neg_cos) { C0 = -C0; C1 = -C1; C2 = -C2; C3 = -C3; C4 = -C4; }
  if ((n & 1) == 0) {
    const double x3 = __dmul_rn(x, x2);
    const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
    const double x7 = __dmul_rn(x3, x2);
    const double s = __dadd_rn(x, __dmul_rn(x3, S1));
    return __dadd_rn(s, __dmul_rn(x7, s1));
  } else {
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(C3, __dmul_rn(x2, C4));
    const double c1 = __dadd_rn(C1, __dmul_rn(x2, C2));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(C0, __dmul_rn(x2, c1));
    return __dadd_rn(c, __dmul_rn(x6, c2));
  }
}

__device__ __forceinline__ void glibc_sincosf(float y, float* sinp, float* cosp) {
  const double x = (double)y;
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ffu;
  if (top < ((0x3f490fdbu >> 20) & 0x7ffu)) {  // |y| < pi/4 (top-12-bit compare as in glibc)
    const double x2 = __dmul_rn(x, x);
    if (top < ((0x39800000u >> 20) & 0x7ffu)) {  // |y| < 2^-12
      *sinp = y;
      *cosp = 1.0f;
      return;
    }
    *sinp = (float)sincos_poly(x, x2, false, 0);
    *cosp = (float)sincos_poly(x, x2, false, 1);
    return;
  }
  const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
  const int n = ((int)r + 0x800000) >> 24;
  const double xr = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
  const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  const bool neg = (n & 2) != 0;
  const double xs = __dmul_rn(xr, sgn), x2 = __dmul_rn(xr, xr);
  *sinp = (float)sincos_poly(xs, x2, neg, n);
  *cosp = (float)sincos_poly(xs, x2, neg, n ^ 1);
}

__global__ void __launch_bounds__(256) k_orient_describe(const OrbGeom* __restrict__ g, OrbLevel0 l0,
                                                         const uint8_t* __restrict__ pyr,
                                                         const uint8_t* __restrict__ blur,
                                                         const uint32_t* __restrict__ sel,
                                                         const int* __restrict__ sel_count,
                                                         orbx_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
                                                         int* __restrict__ counts, int cap) {
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // keypoint index in the frame
  const int* sc = sel_count + frame * g->nlevels;
  int level = -1, first = 0, total = 0;
  for (int l = 0; l < g->nlevels; ++l) {
    const int n = sc[l];
    if (level < 0 && gidx < total + n) { level = l; first = total; }
    total += n;
  }
  if (gidx == 0 && lane == 0) counts[frame] = total;
  if (level < 0 || gidx >= cap) return;
  const OrbLevelGeom& L = g->lv[level];
  const uint32_t key = sel[(size_t)frame * g->kp_cap_frame + L.sel_off + (gidx - first)];
  const int kx = OT_KEY_X(key) + ORB_MIN_BORDER, ky = OT_KEY_Y(key) + ORB_MIN_BORDER;  // :843-844

  // IC_Angle on the un-blurred level: lane = u + 15
  const LevelView S = level_view(g, l0, pyr, level, frame);
  const uint8_t* center = S.base + (size_t)ky * S.pitch + kx;
  const int u = lane - ORB_HALF_PATCH;
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    m10 = u * (int)center[u];
#pragma unroll
    for (int v = 1; v <= ORB_HALF_PATCH; ++v) {
      if (abs(u) <= c_umax[v]) {
        const int plus = center[u + v * S.pitch], minus = center[u - v * S.pitch];
        m10 += u * (plus + minus);
        m01 += v * (plus - minus);
      }
    }
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // rotated BRIEF on the blurred level: lane computes descriptor byte `lane`
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  float a, b;
  glibc_sincosf(__fmul_rn(angle, factorPI), &b, &a);
  const uint8_t* bc = blur + (size_t)frame * g->blur_frame_bytes + L.blur_off + (size_t)ky * L.bpitch + kx;
  const char4* pat = reinterpret_cast<const char4*>(d_pattern) + lane * 8;
  int val = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const char4 pp = pat[k];
    const float x0 = (float)pp.x, y0 = (float)pp.y, x1 = (float)pp.z, y1 = (float)pp.w;
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = bc[r0 * L.bpitch + c0], t1 = bc[r1 * L.bpitch + c1];
    val |= (t0 < t1) << k;
  }
  const size_t o = (size_t)frame * cap + gidx;
  desc[o * 32 + lane] = (uint8_t)val;
  if (lane < 6) {
    float f;
    if (lane == 0) f = level ? __fmul_rn((float)kx, L.scale) : (float)kx;
    else if (lane == 1) f = level ? __fmul_rn((float)ky, L.scale) : (float)ky;
    else if (lane == 2) f = L.patch_size;
    else if (lane == 3) f = angle;
    else if (lane == 4) f = (float)OT_KEY_SCORE(key);
    else f = __int_as_float(level);
    reinterpret_cast<float*>(kps + o)[lane] = f;
  }
}

// ------------------------------------------------------------------------------------------
// launch wrappers (host)
void launch_pyramid(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, uint8_t* d_pyr, cudaStream_t st,
                    long long* launches) {
  for (int l = 1; l < gh.g.nlevels; ++l) {
    const OrbLevelGeom& L = gh.g.lv[l];
    // tile height: the tallest of 32 / 16 / 8 rows that still gives every SM several waves of warps
    // (a level's launch ends with a partial wave; short tiles keep that tail small on the small levels)
    const int tiles_x = ((L.pitch >> 2) + 31) >> 5;
    int rows = PYR_ROWS;
    while (rows > 8 && (long long)tiles_x * ((L.h + rows - 1) / rows) * n_frames < 4LL * 148 * 24) rows >>= 1;
    const int tiles = tiles_x * ((L.h + rows - 1) / rows);
    k_pyr_resize<<<dim3((tiles + 3) / 4, n_frames), 128, 0, st>>>(l, rows, gh.d_geom, l0, gh.d_xtab, gh.d_ytab, d_pyr);
    ++*launches;
  }
}

void launch_fast(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, const uint8_t* d_pyr, uint32_t* d_cand,
                 int* d_cell_count, cudaStream_t st, long long* launches) {
  if (gh.fast_mode == 1) {
    k_fast_bands<<<dim3(gh.n_bands, n_frames), 32, 0, st>>>(gh.d_bands, l0, d_pyr, d_cand, d_cell_count, gh.g.pyr_frame_bytes,
                                                            gh.g.cand_frame_u32, gh.g.n_cells, gh.g.ini_th, gh.g.min_th);
    ++*launches;
    return;
  }
  const bool split = gh.fast_mode == 2;
  const size_t bm_frame_bytes = (size_t)gh.bm_rows_frame * 32;
  if (split) {
    k_fast_prefilter<<<dim3((gh.n_bands + 3) / 4, n_frames), 128, 0, st>>>(gh.d_bands, gh.d_band_bm, gh.n_bands, l0, d_pyr,
                                                                          gh.d_bitmap, gh.g.pyr_frame_bytes, bm_frame_bytes,
                                                                          gh.g.ini_th);
    ++*launches;
  }
  // shared memory sized for this geometry's largest cell (rows x pitch image + measure map,
  // survivor queue, kept list, overlaid): ~6 KB at 640x480, so 32 cell CTAs stay resident per SM
  int rows_max = 0, t_max = 1, cw_max = 0;
  for (int l = 0; l < gh.g.nlevels; ++l) {
    rows_max = max(rows_max, gh.g.lv[l].h_cell + 6);
    cw_max = max(cw_max, gh.g.lv[l].w_cell + 6);
    t_max = max(t_max, gh.g.lv[l].w_cell * gh.g.lv[l].h_cell);
  }
  const int pitch = cw_max + 3 + (FAST_COL0 - 1) <= 48 ? 48 : 80;  // + up to 3 bytes of word misalignment (+ the pair shift)
  const size_t r_img = (size_t)rows_max * pitch;
  const size_t r_hv = (size_t)rows_max * (FAST_HVP && !FAST_PAIR ? FAST_HVP : pitch);
  const size_t smem_cell = (2 * r_img + max(r_hv, (size_t)((2 * t_max + 31) & ~15)) + 16 + 15) & ~(size_t)15;
  const size_t smem = smem_cell * FAST_WPC;
  const dim3 grid((gh.g.n_cells + FAST_WPC - 1) / FAST_WPC, n_frames);
  if (pitch == 48)
    k_fast_cells<48><<<grid, FAST_NT * FAST_WPC, smem, st>>>(gh.d_cells, l0, d_pyr, d_cand, d_cell_count, gh.g.pyr_frame_bytes,
                                                  gh.g.cand_frame_u32, gh.g.n_cells, gh.g.ini_th, gh.g.min_th, rows_max, t_max,
                                                  split ? gh.d_bitmap : nullptr, gh.d_cell_bm, bm_frame_bytes);
  else
    k_fast_cells<80><<<grid, FAST_NT * FAST_WPC, smem, st>>>(gh.d_cells, l0, d_pyr, d_cand, d_cell_count, gh.g.pyr_frame_bytes,
                                                  gh.g.cand_frame_u32, gh.g.n_cells, gh.g.ini_th, gh.g.min_th, rows_max, t_max,
                                                  split ? gh.d_bitmap : nullptr, gh.d_cell_bm, bm_frame_bytes);
  ++*launches;
}

// Shared-memory budget of the octree kernel, resolved ONCE per handle (orbx_create) and kept in the
// handle's OrbGeomHost: the launch and the opt-in must agree, whatever happens to the environment later.
static int octree_smem_keys(const OrbGeom& g) {
  int m = 0;
  for (int l = 0; l < g.nlevels; ++l) m = max(m, g.lv[l].key_cap);
  int cap_keys = 6144;  // 36 KB of keys + labels keeps ~3 CTAs resident per SM
  if (const char* e = getenv("ORB_OT_SMEM_KEYS")) cap_keys = max(256, atoi(e));  // tuning knob (levels above it use HBM)
  return min(m, cap_keys);
}

cudaError_t prepare_pyramid(const OrbGeom&) { return cudaSuccess; }  // the pyramid kernels use no shared memory

// The opt-in limit is a property of the FUNCTION (per device), shared by every handle: only ever
// raise it, so a handle with a smaller nfeatures cannot shrink it under another handle's feet.
// The reference creates one extractor per camera, possibly from several threads: serialised.
cudaError_t prepare_octree(OrbGeomHost& gh) {
  static int raised[64] = {0};
  static std::mutex mu;
  OtScratch s;
  gh.ot_smem_keys = octree_smem_keys(gh.g);
  gh.ot_smem_bytes = (size_t)ot_layout(s, gh.g.ot_cap, gh.g.ot_scan_cap, 256) + (size_t)gh.ot_smem_keys * 6;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const int need = (int)gh.ot_smem_bytes;
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 64 && need <= raised[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
  if (e == cudaSuccess && dev < 64) raised[dev] = need;
  return e;
}

void launch_octree(const OrbGeomHost& gh, int n_frames, const uint32_t* d_cand, const int* d_cell_count,
                   uint32_t* d_keys, uint16_t* d_knode, uint32_t* d_sel, int* d_sel_count, cudaStream_t st,
                   long long* launches) {
  k_octree<<<dim3(n_frames, gh.g.nlevels), 256, gh.ot_smem_bytes, st>>>(gh.d_geom, d_cand, d_cell_count, d_keys, d_knode,
                                                                         d_sel, d_sel_count, gh.ot_smem_keys);
  ++*launches;
}

#ifdef ORB_OT_TIMING
void dump_octree_marks() {
  int nm = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&nm, g_ot_nmarks, 4);
  static long long marks[2 * 4096];
  cudaMemcpyFromSymbol(marks, g_ot_marks, sizeof(long long) * 2 * nm);
  long long agg[128] = {0};
  int cnt[128] = {0};
  for (int i = 1; i < nm; ++i) {
    const int id = (int)marks[2 * i];
    if (id == 20) continue;  // new launch
    agg[id] += marks[2 * i + 1] - marks[2 * i - 1];
    cnt[id]++;
  }
  printf("octree marks: %d\n", nm);
  for (int id = 0; id < 128; ++id)
    if (cnt[id]) printf("  phase ending at mark %3d: %4d times, %9lld cycles total, %7lld avg\n", id, cnt[id], agg[id], agg[id] / cnt[id]);
  const int zero = 0;
  cudaMemcpyToSymbol(g_ot_nmarks, &zero, 4);
}
#endif

void launch_blur(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, const uint8_t* d_pyr, uint8_t* d_blur,
                 cudaStream_t st, long long* launches) {
  k_blur<<<dim3((gh.g.n_blur_tiles + 3) / 4, n_frames), 128, 0, st>>>(gh.d_geom, l0, d_pyr, d_blur);
  ++*launches;
}

void launch_orient_describe(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, const uint8_t* d_pyr, const uint8_t* d_blur,
                            const uint32_t* d_sel, const int* d_sel_count, orbx_keypoint* d_kps, uint8_t* d_desc,
                            int* d_counts, int cap, cudaStream_t st, long long* launches) {
  const int max_kp = gh.g.kp_cap_frame;  // upper bound of keypoints per frame
  k_orient_describe<<<dim3((max_kp + 7) / 8, n_frames), 256, 0, st>>>(gh.d_geom, l0, d_pyr, d_blur, d_sel, d_sel_count,
                                                                        d_kps, d_desc, d_counts, cap);
  ++*launches;
}

}  // namespace orbk
