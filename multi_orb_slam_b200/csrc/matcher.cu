// Hand-written sm_100a kernels + C-ABI of the matcher hot path (include/orb_b200.h, orbm_*).
// Reference: /root/reference/src/ORBmatcher.cc — DescriptorDistance :3994-4010,
// SearchForInitialization :868-983, SearchByProjection(Frame&, vector<MapPoint*>&, th) :62-157,
// ComputeThreeMaxima :3948-3989; and the grid index of src/Frame.cc — AssignFeaturesToGrid
// :348-395, PosInGrid :632-642, GetFeaturesInArea :510-566.
//
// Descriptors are 32 bytes = 8 x u32; distance = sum of __popc(a ^ b).  Matching is bound by the
// integer POPC issue rate (16 lanes/clk/SM), not by HBM: descriptors are tiny and stay on chip.
// Order-dependent reference loops are split into a parallel phase (grid lookup + distances,
// candidates kept in the reference's traversal order) and an ordered resolve phase.
#include "device_guard.h"
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/orb_b200.h"

#define GRID_COLS 64  // FRAME_GRID_COLS (include/Frame.h)
#define GRID_ROWS 48  // FRAME_GRID_ROWS
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define TH_HIGH 100   // src/ORBmatcher.cc:37-39
#define TH_LOW 50
#define HISTO_LENGTH 30

namespace {

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ int hamming_rows(const uint8_t* a, const uint8_t* b) {
  const uint4* pa = reinterpret_cast<const uint4*>(a);
  const uint4* pb = reinterpret_cast<const uint4*>(b);
  return hamming256(pa[0], pa[1], pb[0], pb[1]);
}

// ---- DescriptorDistance for n independent pairs ------------------------------------------
__global__ void k_distance_pairs(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n,
                                 int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hamming_rows(a + (size_t)i * 32, b + (size_t)i * 32);
}

// ---- brute force ---------------------------------------------------------------------------
// Each thread owns BF_QPT query rows in registers; the CTA streams its chunk of targets through
// shared memory (every lane reads the same target word: broadcast).  Best / second-best are
// tracked on packed keys (dist << 16 | local target index): min over keys = lowest distance
// then lowest index, which is exactly the strict-< ascending scan of the reference; the second
// smallest key carries the second-best distance.  Target chunks (blockIdx.y) are merged by
// k_bf_merge in ascending chunk order.
#ifndef BF_THREADS
#define BF_THREADS 128
#endif
#ifndef BF_QPT
#define BF_QPT 2
#endif
#define BF_TILE 256

__global__ void __launch_bounds__(BF_THREADS) k_bruteforce(const uint8_t* __restrict__ q, int nq,
                                                           const uint8_t* __restrict__ t, int nt, int chunk,
                                                           uint2* __restrict__ partial) {
  __shared__ uint4 s_t[BF_TILE * 2];
  const int tid = threadIdx.x;
  const int t_begin = blockIdx.y * chunk, t_end = min(nt, t_begin + chunk);
  uint4 qa[BF_QPT], qb[BF_QPT];
  uint32_t best[BF_QPT], second[BF_QPT];
#pragma unroll
  for (int r = 0; r < BF_QPT; ++r) {
    const int qi = min((blockIdx.x * BF_QPT + r) * BF_THREADS + tid, nq - 1);
    const uint4* p = reinterpret_cast<const uint4*>(q + (size_t)qi * 32);
    qa[r] = __ldg(p);
    qb[r] = __ldg(p + 1);
    best[r] = second[r] = 0xFFFFFFFFu;
  }
  for (int base = t_begin; base < t_end; base += BF_TILE) {
    const int nt_tile = min(BF_TILE, t_end - base);
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(t + (size_t)base * 32);
    for (int i = tid; i < nt_tile * 2; i += BF_THREADS) s_t[i] = __ldg(src + i);
    __syncthreads();
    const uint32_t jbase = (uint32_t)(base - t_begin);
#pragma unroll 4
    for (int j = 0; j < nt_tile; ++j) {
      const uint4 ta = s_t[2 * j], tb = s_t[2 * j + 1];
#pragma unroll
      for (int r = 0; r < BF_QPT; ++r) {
        const uint32_t d = (uint32_t)hamming256(qa[r], qb[r], ta, tb);
        const uint32_t key = (d << 16) + (jbase + (uint32_t)j);
        second[r] = min(second[r], max(best[r], key));
        best[r] = min(best[r], key);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < BF_QPT; ++r) {
    const int qi = (blockIdx.x * BF_QPT + r) * BF_THREADS + tid;
    if (qi < nq) partial[(size_t)blockIdx.y * nq + qi] = make_uint2(best[r], second[r]);
  }
}

__global__ void k_bf_merge(const uint2* __restrict__ partial, int nq, int nsplit, int chunk, float ratio, int th_dist,
                           int32_t* __restrict__ out_idx, int32_t* __restrict__ out_d1, int32_t* __restrict__ out_d2) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  int best = 256, second = 256, idx = -1;
  for (int s = 0; s < nsplit; ++s) {
    const uint2 p = partial[(size_t)s * nq + qi];
    if (p.x == 0xFFFFFFFFu) continue;
    const int b = (int)(p.x >> 16), bi = s * chunk + (int)(p.x & 0xFFFFu);
    const int sc = p.y == 0xFFFFFFFFu ? 256 : (int)(p.y >> 16);
    if (b < best) {  // chunks ascend in target index: strict < keeps the earliest on ties
      second = min(best, sc);
      best = b;
      idx = bi;
    } else {
      second = min(second, b);
    }
  }
  out_d1[qi] = best;
  out_d2[qi] = second;
  out_idx[qi] = (idx >= 0 && best <= th_dist && (float)best < __fmul_rn((float)second, ratio)) ? idx : -1;
}

// Batched variant for many small independent pairs (cross-camera matching of a rig: one pair =
// (camera c, camera c') of one rig-frame).  Pair p matches rows q + p*q_stride (nq[p] valid)
// against rows t + p*t_stride (nt[p] valid); results are strided by `cap`.  One CTA per
// (256-query tile, pair); all targets of the pair stream through shared memory, so no merge pass.
__device__ __forceinline__ void bf_pair_body(const uint8_t* __restrict__ qp, int nq, const uint8_t* __restrict__ tp, int nt,
                                             int pair, int cap, float ratio, int th_dist, int32_t* __restrict__ out_idx,
                                             int32_t* __restrict__ out_d1, int32_t* __restrict__ out_d2) {
  __shared__ uint4 s_t[BF_TILE * 2];
  const int tid = threadIdx.x;
  if ((int)(blockIdx.x * BF_QPT * BF_THREADS) >= nq) return;
  uint4 qa[BF_QPT], qb[BF_QPT];
  uint32_t best[BF_QPT], second[BF_QPT];
#pragma unroll
  for (int r = 0; r < BF_QPT; ++r) {
    const int qi = min((blockIdx.x * BF_QPT + r) * BF_THREADS + tid, nq - 1);
    const uint4* p = reinterpret_cast<const uint4*>(qp + (size_t)qi * 32);
    qa[r] = __ldg(p);
    qb[r] = __ldg(p + 1);
    best[r] = second[r] = 0xFFFFFFFFu;
  }
  for (int base = 0; base < nt; base += BF_TILE) {
    const int nt_tile = min(BF_TILE, nt - base);
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(tp + (size_t)base * 32);
    for (int i = tid; i < nt_tile * 2; i += BF_THREADS) s_t[i] = __ldg(src + i);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nt_tile; ++j) {
      const uint4 ta = s_t[2 * j], tb = s_t[2 * j + 1];
#pragma unroll
      for (int r = 0; r < BF_QPT; ++r) {
        const uint32_t d = (uint32_t)hamming256(qa[r], qb[r], ta, tb);
        const uint32_t key = (d << 16) + (uint32_t)(base + j);
        second[r] = min(second[r], max(best[r], key));
        best[r] = min(best[r], key);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < BF_QPT; ++r) {
    const int qi = (blockIdx.x * BF_QPT + r) * BF_THREADS + tid;
    if (qi < nq) {
      const int b = best[r] == 0xFFFFFFFFu ? 256 : (int)(best[r] >> 16);
      const int sc = second[r] == 0xFFFFFFFFu ? 256 : (int)(second[r] >> 16);
      const int bi = best[r] == 0xFFFFFFFFu ? -1 : (int)(best[r] & 0xFFFFu);
      const size_t o = (size_t)pair * cap + qi;
      out_d1[o] = b;
      out_d2[o] = sc;
      out_idx[o] = (bi >= 0 && b <= th_dist && (float)b < __fmul_rn((float)sc, ratio)) ? bi : -1;
    }
  }
}

__global__ void __launch_bounds__(BF_THREADS) k_bruteforce_batch(const uint8_t* __restrict__ q, const int32_t* __restrict__ nq_arr,
                                                                 size_t q_stride, const uint8_t* __restrict__ t,
                                                                 const int32_t* __restrict__ nt_arr, size_t t_stride, int cap,
                                                                 float ratio, int th_dist, int32_t* __restrict__ out_idx,
                                                                 int32_t* __restrict__ out_d1, int32_t* __restrict__ out_d2) {
  const int pair = blockIdx.y;
  bf_pair_body(q + (size_t)pair * q_stride, min(nq_arr[pair], cap), t + (size_t)pair * t_stride, min(nt_arr[pair], cap), pair,
               cap, ratio, th_dist, out_idx, out_d1, out_d2);
}

// The same scan for pairs whose rows sit anywhere inside one buffer (the all-gathered per-camera blocks of a rig,
// multi_orb_slam_b200/dist.py): pair p reads its query / target rows and its two counts at byte offsets from `base`.
__global__ void __launch_bounds__(BF_THREADS) k_bruteforce_indexed(const uint8_t* __restrict__ base,
                                                                   const long long* __restrict__ q_off,
                                                                   const long long* __restrict__ t_off,
                                                                   const long long* __restrict__ nq_off,
                                                                   const long long* __restrict__ nt_off, int cap, float ratio,
                                                                   int th_dist, int32_t* __restrict__ out_idx,
                                                                   int32_t* __restrict__ out_d1, int32_t* __restrict__ out_d2) {
  const int pair = blockIdx.y;
  const int nq = *reinterpret_cast<const int32_t*>(base + nq_off[pair]);
  const int nt = *reinterpret_cast<const int32_t*>(base + nt_off[pair]);
  bf_pair_body(base + q_off[pair], min(nq, cap), base + t_off[pair], min(nt, cap), pair, cap, ratio, th_dist, out_idx, out_d1,
               out_d2);
}

// ---- frame grid ------------------------------------------------------------------------------
// Cells are stored column-major (cell = ix * GRID_ROWS + iy) so that the reference's traversal
// (ix outer, iy inner, insertion order inside a cell) of one grid column is ONE contiguous run
// of `items`.
struct GridView {
  const int* start;        // [GRID_CELLS + 1]
  const uint16_t* items;   // keypoint indices, cell-major, insertion (= index) order
  float min_x, min_y, inv_w, inv_h;
};

__device__ __forceinline__ int grid_cell_of(float x, float y, float min_x, float min_y, float inv_w, float inv_h) {
  // PosInGrid: round() = half away from zero (src/Frame.cc:634-635)
  const int px = (int)roundf(__fmul_rn(__fsub_rn(x, min_x), inv_w));
  const int py = (int)roundf(__fmul_rn(__fsub_rn(y, min_y), inv_h));
  if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
  return px * GRID_ROWS + py;
}

// One CTA per frame: counting sort of keypoints into grid cells, stable in keypoint index.
#define GRID_CELL_CACHE 4096  // keypoints whose grid cell is kept in shared memory between the count and the placement
__global__ void __launch_bounds__(256) k_build_grid(const orbx_keypoint* __restrict__ kps, const int32_t* __restrict__ n_arr,
                                                    int n_fixed, int cap, orbm_bounds b, int* __restrict__ start_out,
                                                    uint16_t* __restrict__ items_out,
                                                    const int32_t* __restrict__ cam_of = nullptr, int cam = 0,
                                                    int grid_per_camera = 0) {
  // grid_per_camera: one CTA per camera of ONE keypoint set (blockIdx.x = camera, outputs `cap` / GRID_CELLS + 1 apart);
  // otherwise one CTA per frame of a batch (keypoints `cap` apart)
  __shared__ int s_cnt[GRID_CELLS + 1];
  __shared__ int s_warp[8];
  __shared__ uint16_t s_cell[GRID_CELL_CACHE];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (grid_per_camera) cam = frame;
  const int n = n_arr ? min(n_arr[frame], cap) : n_fixed;
  const orbx_keypoint* k = grid_per_camera ? kps : kps + (size_t)frame * cap;
  const float inv_w = __fdiv_rn((float)GRID_COLS, __fsub_rn(b.max_x, b.min_x));
  const float inv_h = __fdiv_rn((float)GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
  auto cell_of = [&](int i) {
    int c = grid_cell_of(k[i].x, k[i].y, b.min_x, b.min_y, inv_w, inv_h);
    if (cam_of && cam_of[i] != cam) c = -1;  // mGrids[cam] holds that camera's keypoints only (src/Frame.cc:384-393)
    return c;
  };
  for (int i = tid; i <= GRID_CELLS; i += 256) s_cnt[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    const int c = cell_of(i);
    if (i < GRID_CELL_CACHE) s_cell[i] = (uint16_t)c;  // -1 -> 0xFFFF (GRID_CELLS < 0xFFFF)
    if (c >= 0) atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of GRID_CELLS counts: 12 per thread, warp scans, a scan of the 8 warp sums
  {
    const int lo = tid * 12;
    int s = 0;
    for (int i = lo; i < lo + 12; ++i) s += s_cnt[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    int run = incl - s;
    for (int ww = 0; ww < w; ++ww) run += s_warp[ww];
    for (int i = lo; i < lo + 12; ++i) { const int v = s_cnt[i]; s_cnt[i] = run; run += v; }
    if (tid == 255) s_cnt[GRID_CELLS] = run;
  }
  __syncthreads();
  int* start = start_out + (size_t)frame * (GRID_CELLS + 1);
  for (int i = tid; i <= GRID_CELLS; i += 256) start[i] = s_cnt[i];
  __syncthreads();
  // stable placement by warp 0: 32 keypoints at a time, in index order; s_cnt becomes the fill cursor
  if (tid < 32) {
    uint16_t* items = items_out + (size_t)frame * cap;
    for (int base = 0; base < n; base += 32) {
      const int i = base + tid;
      int c = -1;
      if (i < n) {
        if (i < GRID_CELL_CACHE) { const int cc = s_cell[i]; c = cc == 0xFFFF ? -1 : cc; }
        else c = cell_of(i);
      }
      const unsigned peers = __match_any_sync(0xffffffffu, c);
      if (c >= 0) {
        const int rank = __popc(peers & ((1u << tid) - 1u));
        items[s_cnt[c] + rank] = (uint16_t)i;
      }
      __syncwarp();
      if (c >= 0 && (peers >> tid) == 1u) s_cnt[c] += __popc(peers);  // highest peer lane updates the cursor
      __syncwarp();
    }
  }
}

// Frame::GetFeaturesInArea (src/Frame.cc:510-566), warp-cooperative.  Calls emit(idx) for every
// keypoint that passes the cell window, octave window and |dx|<r, |dy|<r tests, in the
// reference's traversal order; `emit` receives (lane_has, idx) and must be called by all lanes.
template <typename Emit>
__device__ __forceinline__ void grid_query(const GridView& gv, const orbx_keypoint* __restrict__ k, float x, float y,
                                           float r, int min_level, int max_level, Emit emit) {
  const int lane = threadIdx.x & 31;
  const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, gv.min_x), r), gv.inv_w)));
  if (cx0 >= GRID_COLS) return;
  const int cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, gv.min_x), r), gv.inv_w)));
  if (cx1 < 0) return;
  const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, gv.min_y), r), gv.inv_h)));
  if (cy0 >= GRID_ROWS) return;
  const int cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, gv.min_y), r), gv.inv_h)));
  if (cy1 < 0) return;
  const bool check = (min_level > 0) || (max_level >= 0);
  // One contiguous run of `items` per grid column (cells are stored column-major).  The runs of
  // all (<= 64) columns are flattened: lanes fetch the run bounds of two columns each, a warp
  // prefix sum gives every column's offset in the flattened order, and the entries are visited 32
  // at a time in exactly the reference's (ix, iy, insertion) order; a lane finds the column of its
  // entry by a shuffle binary search over the (sorted) offsets.
  const int ncol = cx1 - cx0 + 1;
  int r0a = 0, len_a = 0, r0b = 0, len_b = 0;
  if (lane < ncol) {
    r0a = gv.start[(cx0 + lane) * GRID_ROWS + cy0];
    len_a = gv.start[(cx0 + lane) * GRID_ROWS + cy1 + 1] - r0a;
  }
  if (lane + 32 < ncol) {
    r0b = gv.start[(cx0 + lane + 32) * GRID_ROWS + cy0];
    len_b = gv.start[(cx0 + lane + 32) * GRID_ROWS + cy1 + 1] - r0b;
  }
  int incl_a = len_a, incl_b = len_b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int va = __shfl_up_sync(0xffffffffu, incl_a, o), vb = __shfl_up_sync(0xffffffffu, incl_b, o);
    if (lane >= o) { incl_a += va; incl_b += vb; }
  }
  const int total_a = __shfl_sync(0xffffffffu, incl_a, 31);
  const int total = total_a + __shfl_sync(0xffffffffu, incl_b, 31);
  const int excl_a = incl_a - len_a, excl_b = incl_b - len_b;
  for (int e = 0; e < total; e += 32) {
    const int ei = e + lane;
    const bool second_half = ei >= total_a;
    const int key = second_half ? ei - total_a : ei;
    int ca = 0, cb = 0;  // largest column whose exclusive offset is <= key (empty columns precede it)
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int va = __shfl_sync(0xffffffffu, excl_a, (ca + step) & 31), vb = __shfl_sync(0xffffffffu, excl_b, (cb + step) & 31);
      if (ca + step < 32 && va <= key) ca += step;
      if (cb + step < 32 && vb <= key) cb += step;
    }
    const int ra = __shfl_sync(0xffffffffu, r0a, ca), xa = __shfl_sync(0xffffffffu, excl_a, ca);
    const int rb = __shfl_sync(0xffffffffu, r0b, cb), xb = __shfl_sync(0xffffffffu, excl_b, cb);
    bool ok = false;
    int idx = 0;
    if (ei < total) {
      idx = gv.items[second_half ? rb + (key - xb) : ra + (key - xa)];
      const orbx_keypoint kp = k[idx];
      ok = true;
      if (check) {
        if (kp.octave < min_level) ok = false;
        if (max_level >= 0 && kp.octave > max_level) ok = false;
      }
      ok = ok && fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
    }
    emit(ok, idx);
  }
}

// Warp-wide top-2 over packed keys (smaller = better); all lanes get the result.  Keys carry the
// candidate position, so they are unique except for the empty marker 0xFFFFFFFF: the warp's second
// best is the minimum over each lane's runner-up, where the lane that owns the best contributes
// its own second.  Two REDUX instructions instead of a ten-shuffle butterfly (this sits on the
// serial chain of the ordered resolve kernels).
__device__ __forceinline__ void warp_top2(uint32_t& best, uint32_t& second) {
  const uint32_t gb = __reduce_min_sync(0xffffffffu, best);
  second = __reduce_min_sync(0xffffffffu, best == gb ? second : best);
  best = gb;
}

__device__ __forceinline__ void three_maxima_keep(const int* hist, int* keep) {  // :3948-3989, one thread
  int max1 = 0, max2 = 0, max3 = 0, i1_ = -1, i2_ = -1, i3_ = -1;
  for (int i = 0; i < HISTO_LENGTH; ++i) {
    const int s = hist[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3_ = i2_; i2_ = i1_; i1_ = i; }
    else if (s > max2) { max3 = max2; max2 = s; i3_ = i2_; i2_ = i; }
    else if (s > max3) { max3 = s; i3_ = i; }
  }
  if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2_ = -1; i3_ = -1; }
  else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3_ = -1; }
  for (int i = 0; i < HISTO_LENGTH; ++i) keep[i] = (i == i1_ || i == i2_ || i == i3_);
}

// ---- SearchForInitialization ----------------------------------------------------------------
// Phase A (k_init_candidates, one warp per (pair, i1), whole GPU): for each level-0 keypoint i1
// of F1, the window query on F2's grid and the Hamming distances; candidates are stored in the
// reference's traversal order as (dist << 16 | i2).
// Phase B (k_init_resolve, one CTA per pair; warp 0 walks i1 ascending): the best/second scan with
// the vMatchedDistance filter (:907), acceptance, match stealing (:926-933) and histogram; then
// the three-maxima rotation filter and the vbPrevMatched update, in parallel.
__global__ void __launch_bounds__(256) k_init_candidates(int cap, const orbx_keypoint* __restrict__ k1_all,
                                                         const uint8_t* __restrict__ d1_all,
                                                         const int32_t* __restrict__ n1_arr,
                                                         const orbx_keypoint* __restrict__ k2_all,
                                                         const uint8_t* __restrict__ d2_all, orbm_bounds b2,
                                                         const int* __restrict__ grid_start,
                                                         const uint16_t* __restrict__ grid_items,
                                                         const float* __restrict__ prev_all, float window,
                                                         uint32_t* __restrict__ cand_all, int* __restrict__ cand_cnt_all) {
  const int pair = blockIdx.y, lane = threadIdx.x & 31;
  const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n1 = min(n1_arr[pair], cap);
  if (i1 >= n1) return;
  const orbx_keypoint* k1 = k1_all + (size_t)pair * cap;
  const orbx_keypoint* k2 = k2_all + (size_t)pair * cap;
  const uint8_t* d2 = d2_all + (size_t)pair * cap * 32;
  const orbx_keypoint kp1 = k1[i1];
  int cnt = 0;
  if (kp1.octave <= 0) {  // level1 > 0 -> continue (:885-887)
    GridView gv;
    gv.start = grid_start + (size_t)pair * (GRID_CELLS + 1);
    gv.items = grid_items + (size_t)pair * cap;
    gv.min_x = b2.min_x;
    gv.min_y = b2.min_y;
    gv.inv_w = __fdiv_rn((float)GRID_COLS, __fsub_rn(b2.max_x, b2.min_x));
    gv.inv_h = __fdiv_rn((float)GRID_ROWS, __fsub_rn(b2.max_y, b2.min_y));
    const uint4* q = reinterpret_cast<const uint4*>(d1_all + ((size_t)pair * cap + i1) * 32);
    const uint4 qa = __ldg(q), qb = __ldg(q + 1);
    uint32_t* row = cand_all + ((size_t)pair * cap + i1) * cap;
    const float* prev = prev_all ? prev_all + ((size_t)pair * cap + i1) * 2 : nullptr;
    const float px = prev ? prev[0] : kp1.x, py = prev ? prev[1] : kp1.y;
    grid_query(gv, k2, px, py, window, kp1.octave, kp1.octave, [&](bool ok, int idx) {
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint4* tp = reinterpret_cast<const uint4*>(d2 + (size_t)idx * 32);
        const int dist = hamming256(qa, qb, __ldg(tp), __ldg(tp + 1));
        row[cnt + __popc(m & ((1u << lane) - 1u))] = (uint32_t)dist << 16 | (uint32_t)idx;
      }
      cnt += __popc(m);
    });
  }
  if (lane == 0) cand_cnt_all[(size_t)pair * cap + i1] = cnt;
}

#define INIT_SMEM_CAND 12288  // candidate entries staged in shared memory per pair (48 KB)
#define INIT_NT 1024          // threads of the resolve CTA: 32 warps stage the candidate rows (a chain of global-load
                              // latencies per warp, a row per iteration) and then walk the queries 32 at a time

// Ordered walk of k_init_resolve, 32 queries per step (the scheme of k_proj_resolve_cta).  A query's outcome - accepted
// or not, and with which keypoint of F2 - is a function of its two least keys among the candidates that pass the
// vMatchedDistance filter (:907), and that filter only ever removes candidates as the walk proceeds (matched distances
// only shrink).  So every pending warp scans its query's row under the distances as they stand; warp 0 commits the
// longest prefix of the step whose answers cannot depend on each other: a query is in conflict when a LOWER query of
// the step that accepts a match takes the keypoint of its best or of its second candidate (claim bytes, lowest lane per
// keypoint); the lowest pending query never conflicts, so every round commits at least one, in order, and the rest scan
// again.  Stealing a keypoint from an earlier query (:926-933) touches only that earlier query, which is never one of
// the same prefix (two queries of a prefix never share a keypoint).

__global__ void __launch_bounds__(INIT_NT) k_init_resolve(int cap, const orbx_keypoint* __restrict__ k1_all,
                                                      const int32_t* __restrict__ n1_arr,
                                                      const orbx_keypoint* __restrict__ k2_all,
                                                      const int32_t* __restrict__ n2_arr, float* __restrict__ prev_all,
                                                      float nnratio, int check_ori, const uint32_t* __restrict__ cand_all,
                                                      const int* __restrict__ cand_cnt_all,
                                                      int32_t* __restrict__ matches12_all, int32_t* __restrict__ nmatches_out) {
  // matchedDist, matches21, bin_of, cand count, cand offset, query list, matches12: [cap] ints each;
  // angle1, angle2: [cap] floats; then INIT_SMEM_CAND staged candidate entries and [cap] claim bytes
  extern __shared__ int s_dyn[];
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_keep[HISTO_LENGTH];
  __shared__ int s_nmatch, s_total, s_nlist;
  __shared__ int s_wsum[INIT_NT / 32];
  __shared__ int s_xb[2][32], s_xs[2][32], s_xw[2][32], s_xd[2][32];  // per-round exchange, double-buffered by round parity
  __shared__ unsigned s_xdone[2], s_xleft[2];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n1 = min(n1_arr[pair], cap), n2 = min(n2_arr[pair], cap);
  const orbx_keypoint* k1 = k1_all + (size_t)pair * cap;
  const orbx_keypoint* k2 = k2_all + (size_t)pair * cap;
  float* prev = prev_all ? prev_all + (size_t)pair * cap * 2 : nullptr;
  const uint32_t* cand = cand_all + (size_t)pair * cap * cap;
  const int* cand_cnt = cand_cnt_all + (size_t)pair * cap;
  int32_t* matches12 = matches12_all + (size_t)pair * cap;
  int* s_mdist = s_dyn;
  int* s_m21 = s_dyn + cap;
  int* s_bin = s_dyn + 2 * cap;
  int* s_cnt = s_dyn + 3 * cap;
  int* s_off = s_dyn + 4 * cap;
  int* s_list = s_dyn + 5 * cap;
  int* s_m12 = s_dyn + 6 * cap;
  float* s_ang1 = reinterpret_cast<float*>(s_dyn + 7 * cap);
  float* s_ang2 = reinterpret_cast<float*>(s_dyn + 8 * cap);
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_dyn + 9 * cap);
  uint8_t* s_claim = reinterpret_cast<uint8_t*>(s_cand + INIT_SMEM_CAND);  // [cap] claim bytes of a round (0xFF = free)
  for (int i = tid; i < cap; i += INIT_NT) {
    s_mdist[i] = 0x7FFFFFFF;
    s_m21[i] = -1;
    s_bin[i] = -1;
    s_m12[i] = -1;
    s_cnt[i] = i < n1 ? cand_cnt[i] : 0;
    s_ang1[i] = i < n1 ? k1[i].angle : 0.f;
    s_ang2[i] = i < n2 ? k2[i].angle : 0.f;
  }
  if (tid < HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) s_nmatch = 0;
  __syncthreads();
  // exclusive scan of the candidate counts (each thread owns a contiguous chunk of queries)
  {
    const int chunk = (cap + INIT_NT - 1) / INIT_NT, lo = min(cap, tid * chunk), hi = min(cap, lo + chunk);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += s_cnt[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int run = incl - s;
    for (int w = 0; w < warp; ++w) run += s_wsum[w];
    if (tid == INIT_NT - 1) s_total = run + s;
    for (int i = lo; i < hi; ++i) { s_off[i] = run; run += s_cnt[i]; }
  }
  __syncthreads();
  const bool staged = s_total <= INIT_SMEM_CAND;
  if (staged) {  // all candidate rows of the pair into shared memory, a warp per row
    for (int i1 = warp; i1 < n1; i1 += INIT_NT / 32) {
      const int cnt = s_cnt[i1], off = s_off[i1];
      const uint32_t* row = cand + (size_t)i1 * cap;
      for (int c = lane; c < cnt; c += 32) s_cand[off + c] = row[c];
    }
  }
  __syncthreads();
  if (warp == 0) {
    // ordered compaction of the queries that have candidates (most keypoints are not level 0)
    int nlist = 0;
    for (int base = 0; base < n1; base += 32) {
      const int i = base + lane;
      const bool has = i < n1 && s_cnt[i] > 0;
      const unsigned bm = __ballot_sync(0xffffffffu, has);
      if (has) s_list[nlist + __popc(bm & ((1u << lane) - 1u))] = i;
      nlist += __popc(bm);
    }
    if (lane == 0) s_nlist = nlist;
  }
  for (int i = tid; i < cap; i += INIT_NT) s_claim[i] = 0xFF;
  __syncthreads();
  const int nlist = s_nlist;
  int nmatches = 0;  // warp 0's count
  int rp = 0;        // round parity of the exchange buffers
  for (int b0 = 0; b0 < nlist; b0 += 32) {
    const int i1 = b0 + warp < nlist ? s_list[b0 + warp] : -1;
    const int cnt = i1 >= 0 ? s_cnt[i1] : 0;
    const uint32_t* row = i1 >= 0 ? (staged ? s_cand + s_off[i1] : cand + (size_t)i1 * cap) : nullptr;
    bool pend = i1 >= 0;
    for (;;) {
      int bidx = -1, sidx = -1, will = 0, bdist = 0;
      if (pend) {
        // key = dist << 16 | traversal position: strict-< scan order (:910-919); every lane remembers the keypoints
        // of its own two least keys
        uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
        int best_i2 = -1, second_i2 = -1;
        for (int c = lane; c < cnt; c += 32) {
          const uint32_t e = row[c];
          const int dist = (int)(e >> 16), i2 = (int)(e & 0xFFFFu);
          if (s_mdist[i2] <= dist) continue;  // :907
          const uint32_t key = (uint32_t)dist << 16 | (uint32_t)c;
          if (key < best) { second = best; second_i2 = best_i2; best = key; best_i2 = i2; }
          else if (key < second) { second = key; second_i2 = i2; }
        }
        const uint32_t my_best = best, my_second = second;
        warp_top2(best, second);
        if (best == 0xFFFFFFFFu) {
          pend = false;  // nothing passes the filter any more: the query is finished without a match
        } else {
          bdist = (int)(best >> 16);
          bidx = __shfl_sync(0xffffffffu, best_i2, __ffs(__ballot_sync(0xffffffffu, my_best == best)) - 1);
          if (second != 0xFFFFFFFFu) {
            const int mine = my_best == second ? best_i2 : (my_second == second ? second_i2 : -1);
            sidx = __shfl_sync(0xffffffffu, mine, __ffs(__ballot_sync(0xffffffffu, mine >= 0)) - 1);
          }
          // bestDist2 stays INT_MAX when there is no second candidate (:896-897)
          const float second_f = second == 0xFFFFFFFFu ? (float)0x7FFFFFFF : (float)(int)(second >> 16);
          will = bdist <= TH_LOW && (float)bdist < __fmul_rn(second_f, nnratio);
        }
      }
      if (lane == 0) { s_xb[rp][warp] = bidx; s_xs[rp][warp] = sidx; s_xw[rp][warp] = will; s_xd[rp][warp] = bdist; }
      __syncthreads();
      if (warp == 0) {
        const int b = s_xb[rp][lane], sx = s_xs[rp][lane], wl = s_xw[rp][lane], bd = s_xd[rp][lane];
        const bool p = b >= 0;
        const bool wants = p && wl;
        const unsigned peers = __match_any_sync(0xffffffffu, wants ? b : -1 - lane);
        const bool claims = wants && (__ffs(peers) - 1 == lane);
        if (claims) s_claim[b] = (uint8_t)lane;
        __syncwarp();
        bool conflict = false;
        if (p) conflict = s_claim[b] < lane || (sx >= 0 && s_claim[sx] < lane);
        const unsigned cm = __ballot_sync(0xffffffffu, conflict);
        const int first = cm ? __ffs(cm) - 1 : 32;
        __syncwarp();
        if (claims) s_claim[b] = 0xFF;
        const bool done = p && lane < first;
        const bool commit = done && wl;
        bool stole = false;
        if (commit) {
          const int q1 = s_list[b0 + lane];
          const int old = s_m21[b];
          if (old >= 0) { s_m12[old] = -1; stole = true; }  // :926-933
          s_m12[q1] = b;
          s_m21[b] = q1;
          s_mdist[b] = bd;
          if (check_ori) {
            float rot = __fsub_rn(s_ang1[q1], s_ang2[b]);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
            if (bin == HISTO_LENGTH) bin = 0;
            s_bin[q1] = bin;
            atomicAdd(&s_hist[bin], 1);
          }
        }
        nmatches += __popc(__ballot_sync(0xffffffffu, commit)) - __popc(__ballot_sync(0xffffffffu, stole));
        const unsigned dm = __ballot_sync(0xffffffffu, done), pm = __ballot_sync(0xffffffffu, p);
        if (lane == 0) { s_xdone[rp] = dm; s_xleft[rp] = pm & ~dm; }
      }
      __syncthreads();
      if (s_xdone[rp] >> warp & 1u) pend = false;
      const unsigned left = s_xleft[rp];
      rp ^= 1;
      if (!left) break;
    }
  }
  if (tid == 0) {
    s_nmatch = nmatches;
    if (check_ori) three_maxima_keep(s_hist, s_keep);  // ComputeThreeMaxima (:3948-3989)
  }
  __syncthreads();
  for (int i1 = tid; i1 < n1; i1 += INIT_NT) {
    int m = s_m12[i1];
    if (check_ori) {
      const int bin = s_bin[i1];
      if (bin >= 0 && !s_keep[bin] && m >= 0) {
        m = -1;
        atomicSub(&s_nmatch, 1);
      }
    }
    matches12[i1] = m;
    if (prev && m >= 0) {  // :978-980
      prev[2 * i1] = k2[m].x;
      prev[2 * i1 + 1] = k2[m].y;
    }
  }
  __syncthreads();
  if (tid == 0) nmatches_out[pair] = s_nmatch;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:381-424) ------------------------
// One CTA per map point: the N x N distance matrix (u16) in shared memory (N <= DD_SMEM_N) or in a
// global scratch block; a warp per row finds the row median (element int(0.5*(N-1)) of the sorted row)
// from a 257-bin histogram of the row; the first row with the least median wins (:414-420).
#define DD_SMEM_N 160
__global__ void __launch_bounds__(128) k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ offsets,
                                                     uint16_t* __restrict__ big, const int* __restrict__ big_off,
                                                     int32_t* __restrict__ best_idx) {
  extern __shared__ uint16_t s_dd[];  // [N*N] if N <= DD_SMEM_N
  __shared__ int s_hist[4][264];
  __shared__ unsigned s_best;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o = offsets[p], N = offsets[p + 1] - o;
  if (N <= 0) {
    if (tid == 0) best_idx[p] = -1;
    return;
  }
  uint16_t* D = N <= DD_SMEM_N ? s_dd : big + big_off[p];
  if (tid == 0) s_best = 0xFFFFFFFFu;
  const uint8_t* d = desc + (size_t)o * 32;
  for (int e = tid; e < N * N; e += 128) {
    const int i = e / N, j = e - i * N;
    if (j < i) continue;
    int dist = 0;
    if (j > i) {
      const uint4* a = reinterpret_cast<const uint4*>(d + (size_t)i * 32);
      const uint4* b = reinterpret_cast<const uint4*>(d + (size_t)j * 32);
      dist = hamming256(__ldg(a), __ldg(a + 1), __ldg(b), __ldg(b + 1));
    }
    D[i * N + j] = (uint16_t)dist;
    D[j * N + i] = (uint16_t)dist;
  }
  __syncthreads();
  const int kth = (int)(0.5 * (N - 1));
  int* hist = s_hist[warp];
  for (int i = warp; i < N; i += 4) {
    for (int b = lane; b < 264; b += 32) hist[b] = 0;
    __syncwarp();
    for (int j = lane; j < N; j += 32) atomicAdd(&hist[D[i * N + j]], 1);
    __syncwarp();
    // smallest value v with #(row <= v) > kth: each lane owns 9 consecutive bins (257 = 28*9 + 5)
    int cnt = 0;
#pragma unroll
    for (int b = 0; b < 9; ++b) cnt += lane * 9 + b < 257 ? hist[lane * 9 + b] : 0;
    int incl = cnt;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += t;
    }
    const unsigned owner = __ballot_sync(0xffffffffu, incl > kth);
    const int ol = __ffs(owner) - 1;
    int median = 0;
    if (lane == ol) {
      int run = incl - cnt;
      for (int b = 0; b < 9; ++b) {
        run += hist[lane * 9 + b];
        if (run > kth) { median = lane * 9 + b; break; }
      }
      atomicMin(&s_best, (unsigned)median << 16 | (unsigned)i);
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0) best_idx[p] = (int)(s_best & 0xFFFFu);
}

// ---- DBoW2 vocabulary descent (TemplatedVocabulary.h:1218-1259) --------------------------------
// Warp per feature: at every level the lanes take the node's children (k = 10 in ORBvoc), the child with the
// least Hamming distance wins, the first one on ties (`d < best_d`) = min over dist << 8 | child position.
__global__ void __launch_bounds__(256) k_bow_descend(const int32_t* __restrict__ child_start, const int32_t* __restrict__ child_ids,
                                                     const uint8_t* __restrict__ node_desc, const uint8_t* __restrict__ desc, int n,
                                                     int nid_level, int32_t* __restrict__ leaf, int32_t* __restrict__ node) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const uint4* qp = reinterpret_cast<const uint4*>(desc + (size_t)i * 32);
  const uint4 qa = __ldg(qp), qb = __ldg(qp + 1);
  int final_id = 0, level = 0, nid = 0;
  int c0 = child_start[0], c1 = child_start[1];
  while (c0 != c1) {  // !isLeaf()
    ++level;
    uint32_t best = 0xFFFFFFFFu;
    for (int c = c0 + lane; c < c1; c += 32) {
      const int id = child_ids[c];
      const uint4* tp = reinterpret_cast<const uint4*>(node_desc + (size_t)id * 32);
      best = min(best, (uint32_t)hamming256(qa, qb, __ldg(tp), __ldg(tp + 1)) << 20 | (uint32_t)(c - c0));
    }
    best = __reduce_min_sync(0xffffffffu, best);
    final_id = child_ids[c0 + (int)(best & 0xFFFFFu)];
    if (level == nid_level) nid = final_id;
    c0 = child_start[final_id];
    c1 = child_start[final_id + 1];
  }
  if (lane == 0) {
    leaf[i] = final_id;
    node[i] = nid;
  }
}

// ---- Frame glue between extractor and matchers (src/Frame.cc) ---------------------------------
// Frame::UndistortKeyPoints (:673-706): cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK) with the
// default criteria = five fixed-point iterations of OpenCV 4.x cvUndistortPointsInternal, all in double
// without contraction (the restatement in oracle/cvprim.cc is pinned bit-exact against cv2 4.13).
// Terms that are exactly zero for a (k1, k2, p1, p2, k3) model are dropped: the rational numerator is 1,
// k8..k11 = 0, the tilt matrices are the identity, and RR = K * I = K so ww = 1.
struct UndistortParams {
  double fx, fy, cx, cy, ifx, ify, k0, k1, k2, k3, k4;
};

__global__ void __launch_bounds__(256) k_undistort(const orbx_keypoint* __restrict__ kps, const int32_t* __restrict__ counts,
                                                   int n_fixed, int cap, UndistortParams P, int identity,
                                                   orbx_keypoint* __restrict__ out) {
  const int frame = blockIdx.y;
  const int n = counts ? min(counts[frame], cap) : n_fixed;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  orbx_keypoint kp = kps[(size_t)frame * cap + i];
  if (!identity) {
    const double u = (double)kp.x, v = (double)kp.y;
    double x = __dmul_rn(__dsub_rn(u, P.cx), P.ifx), y = __dmul_rn(__dsub_rn(v, P.cy), P.ify);
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
      const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y);
      const double r2 = __dadd_rn(xx, yy);
      const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(P.k4, r2), P.k1), r2), P.k0), r2));
      const double icdist = __ddiv_rn(1.0, den);
      if (icdist < 0) {
        x = x0;  // (u - cx) * ifx, the value x0 holds
        y = y0;
        break;
      }
      // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x);  deltaY = k2*(r2 + 2*y*y) + 2*k3*x*y   (left to right)
      const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, P.k2), x), y),
                                  __dmul_rn(P.k3, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
      const double dY = __dadd_rn(__dmul_rn(P.k2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                  __dmul_rn(__dmul_rn(__dmul_rn(2.0, P.k3), x), y));
      x = __dmul_rn(__dsub_rn(x0, dX), icdist);
      y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    kp.x = (float)__dadd_rn(__dmul_rn(P.fx, x), P.cx);
    kp.y = (float)__dadd_rn(__dmul_rn(P.fy, y), P.cy);
  }
  out[(size_t)frame * cap + i] = kp;
}

// Frame::ComputeStereoFromRGBD (:959-985): depth at the distorted keypoint (coordinates truncated).
__global__ void __launch_bounds__(256) k_stereo_rgbd(const orbx_keypoint* __restrict__ kps,
                                                     const orbx_keypoint* __restrict__ kps_un,
                                                     const int32_t* __restrict__ counts, int cap,
                                                     const float* __restrict__ depth, int cols, int rows, size_t row_stride,
                                                     size_t frame_stride, float mbf, float* __restrict__ uright,
                                                     float* __restrict__ depth_out) {
  const int frame = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= cap) return;
  const size_t o = (size_t)frame * cap + i;
  float ur = -1.f, dz = -1.f;
  if (i < min(counts[frame], cap)) {
    const orbx_keypoint kp = kps[o];
    // a keypoint outside the depth image (a depth map smaller than the extraction image) has no depth; the
    // reference's cv::Mat::at would read out of bounds here
    const int px = (int)kp.x, py = (int)kp.y;
    const float d = (px >= 0 && px < cols && py >= 0 && py < rows)
                        ? depth[(size_t)frame * frame_stride + (size_t)py * row_stride + px] : -1.f;
    if (d > 0) {
      dz = d;
      ur = __fsub_rn(kps_un[o].x, __fdiv_rn(mbf, d));
    }
  }
  uright[o] = ur;
  depth_out[o] = dz;
}

// ---- Frame::ComputeStereoMatches (src/Frame.cc:782-956, commented upstream code of the fork) ----
// k_stereo_match: one CTA = 64 left keypoints of one stereo pair.  The right keypoints' row bands
// (:797-810: rows floor(y - 2s) .. ceil(y + 2s)), columns and octaves are staged in shared memory once; a
// warp then takes one left keypoint at a time: lanes scan the right keypoints (band, octave +-1, disparity
// range tests, then the Hamming distance), the first minimum in right-keypoint order wins (key = dist<<16 |
// iR), and 22 lanes evaluate the eleven 11x11 SAD windows from a warp-private copy of the two patches.
// k_stereo_median: one CTA per pair, exact radix select of the SAD median (the sums are < 2^16) and the
// 1.5 * 1.4 * median cut (:942-955).
#define STEREO_MAX_R 4096
#define STEREO_LPB 64

__device__ __forceinline__ int reflect101(int p, int n) {
  p = p < 0 ? -p : p;
  return p >= n ? 2 * (n - 1) - p : p;
}

__global__ void __launch_bounds__(256) k_stereo_match(orbx_pyramid_view pl, orbx_pyramid_view pr, int cap_l,
                                                      const orbx_keypoint* __restrict__ kl, const uint8_t* __restrict__ dl,
                                                      const int32_t* __restrict__ nl, int cap_r,
                                                      const orbx_keypoint* __restrict__ kr, const uint8_t* __restrict__ dr,
                                                      const int32_t* __restrict__ nr, float mbf, float mb,
                                                      float* __restrict__ uright, float* __restrict__ depth,
                                                      int32_t* __restrict__ sad_out) {
  __shared__ float s_x[STEREO_MAX_R];
  __shared__ uint32_t s_rows[STEREO_MAX_R];  // minr | maxr << 16 (an empty band is 1 | 0 << 16)
  __shared__ uint8_t s_oct[STEREO_MAX_R];
  __shared__ int16_t s_l[8][11 * 11];
  __shared__ uint8_t s_r[8][11 * 21 + 1];
  __shared__ float s_d[8][12];
  const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_r = min(nr[frame], cap_r), n_l = min(nl[frame], cap_l);
  const int n_rows = pl.h[0];
  const orbx_keypoint* fr = kr + (size_t)frame * cap_r;
  for (int i = threadIdx.x; i < n_r; i += 256) {
    const orbx_keypoint k = fr[i];
    const float r = __fmul_rn(2.0f, pl.scale[k.octave]);  // the frame's mvScaleFactors (left extractor)
    int maxr = (int)ceilf(__fadd_rn(k.y, r)), minr = (int)floorf(__fsub_rn(k.y, r));
    minr = max(minr, 0);
    maxr = min(maxr, n_rows - 1);
    s_x[i] = k.x;
    s_oct[i] = (uint8_t)k.octave;
    s_rows[i] = minr <= maxr ? ((uint32_t)minr | ((uint32_t)maxr << 16)) : 1u;
  }
  __syncthreads();
  const float max_d = __fdiv_rn(mbf, mb);  // :813-815
  const int i_end = min(n_l, (blockIdx.x + 1) * STEREO_LPB);
  for (int il = blockIdx.x * STEREO_LPB + warp; il < (blockIdx.x + 1) * STEREO_LPB && il < cap_l; il += 8) {
    const size_t o = (size_t)frame * cap_l + il;
    float out_u = -1.0f, out_z = -1.0f;
    int out_sad = -1;
    if (il < i_end) {
      const orbx_keypoint k = kl[o];
      const int level = k.octave, row = (int)k.y;
      const float min_u = __fsub_rn(k.x, max_d), max_u = k.x;
      uint32_t best = (uint32_t)TH_HIGH << 16;
      if (row >= 0 && row < n_rows && !(max_u < 0.0f)) {
        const uint4* pa = reinterpret_cast<const uint4*>(dl + 32 * o);
        const uint4 a0 = pa[0], a1 = pa[1];
        for (int ir = lane; ir < n_r; ir += 32) {
          const uint32_t rows = s_rows[ir];
          const int oc = s_oct[ir];
          const float ur = s_x[ir];
          if (row < (int)(rows & 0xffffu) || row > (int)(rows >> 16) || oc < level - 1 || oc > level + 1 ||
              !(ur >= min_u && ur <= max_u))
            continue;
          const uint4* pb = reinterpret_cast<const uint4*>(dr + 32 * ((size_t)frame * cap_r + ir));
          const uint32_t key = ((uint32_t)hamming256(a0, a1, pb[0], pb[1]) << 16) | (uint32_t)ir;
          best = min(best, key);
        }
      }
      best = __reduce_min_sync(0xffffffffu, best);
      if ((int)(best >> 16) < (TH_HIGH + TH_LOW) / 2 && (best >> 16) < (uint32_t)TH_HIGH) {  // :869
        const float ur0 = s_x[best & 0xffffu];
        const float isf = pl.inv_scale[level];
        const int xl = (int)roundf(__fmul_rn(k.x, isf)), yl = (int)roundf(__fmul_rn(k.y, isf));
        const int xr = (int)roundf(__fmul_rn(ur0, isf));
        const int wl = pl.w[level], hl = pl.h[level], wr = pr.w[level], hr = pr.h[level];
        if (!(xr < 0 || xr + 11 >= wr)) {  // iniu < 0 || endu >= cols (:888-891)
          const uint8_t* il0 = pl.base[level] + (size_t)frame * pl.frame_stride[level];
          const uint8_t* ir0 = pr.base[level] + (size_t)frame * pr.frame_stride[level];
          const int cl = il0[(size_t)reflect101(yl, hl) * pl.pitch[level] + reflect101(xl, wl)];
          for (int t = lane; t < 121; t += 32) {
            const int dy = t / 11, dx = t - dy * 11;
            s_l[warp][t] = (int16_t)((int)il0[(size_t)reflect101(yl + dy - 5, hl) * pl.pitch[level] +
                                              reflect101(xl + dx - 5, wl)] - cl);
          }
          for (int t = lane; t < 231; t += 32) {
            const int dy = t / 21, dx = t - dy * 21;
            s_r[warp][t] = ir0[(size_t)reflect101(yl + dy - 5, hr) * pr.pitch[level] + reflect101(xr + dx - 10, wr)];
          }
          __syncwarp();
          // lanes 0..10 take rows 0..5 of window lane, lanes 11..21 rows 6..10 of window lane - 11
          int sad = 0;
          if (lane < 22) {
            const int t = lane < 11 ? lane : lane - 11;
            const int y0 = lane < 11 ? 0 : 6, y1 = lane < 11 ? 6 : 11;
            const int cr = s_r[warp][5 * 21 + t + 5];
            for (int dy = y0; dy < y1; ++dy)
#pragma unroll
              for (int dx = 0; dx < 11; ++dx) sad += abs((int)s_l[warp][dy * 11 + dx] - ((int)s_r[warp][dy * 21 + t + dx] - cr));
          }
          sad += __shfl_down_sync(0xffffffffu, sad, 11);
          if (lane < 11) s_d[warp][lane] = (float)sad;
          __syncwarp();
          if (lane == 0) {
            int best_inc = 0;
            float best_sad = 2147483648.0f;
            for (int t = 0; t < 11; ++t)
              if (s_d[warp][t] < best_sad) { best_sad = s_d[warp][t]; best_inc = t - 5; }
            if (best_inc != -5 && best_inc != 5) {  // :910
              const float d1 = s_d[warp][5 + best_inc - 1], d2 = s_d[warp][5 + best_inc], d3 = s_d[warp][5 + best_inc + 1];
              const float delta = __fdiv_rn(__fsub_rn(d1, d3),
                                            __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));  // :918
              if (!(delta < -1.0f || delta > 1.0f)) {
                float best_ur = __fmul_rn(pl.scale[level], __fadd_rn(__fadd_rn((float)xr, (float)best_inc), delta));
                float disparity = __fsub_rn(k.x, best_ur);
                if (disparity >= 0.0f && disparity < max_d) {  // :928-938
                  if (disparity <= 0.0f) {
                    disparity = 0.01f;
                    best_ur = (float)__dsub_rn((double)k.x, 0.01);
                  }
                  out_z = __fdiv_rn(mbf, disparity);
                  out_u = best_ur;
                  out_sad = (int)best_sad;
                }
              }
            }
          }
          __syncwarp();
        }
      }
    }
    if (lane == 0) {
      uright[o] = out_u;
      depth[o] = out_z;
      sad_out[o] = out_sad;
    }
  }
}

__global__ void __launch_bounds__(256) k_stereo_median(int cap_l, const int32_t* __restrict__ nl,
                                                       const int32_t* __restrict__ sad, float* __restrict__ uright,
                                                       float* __restrict__ depth) {
  __shared__ int hist[256];
  __shared__ int s_sel[2];
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int n_l = min(nl[frame], cap_l);
  const int32_t* fs = sad + (size_t)frame * cap_l;
  // vDistIdx[size / 2].first of the sorted list: the rank-(n/2) SAD, by two 8-bit radix passes
  int prefix = 0, rank = 0;
  for (int pass = 0; pass < 2; ++pass) {
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n_l; i += 256) {
      const int v = fs[i];
      if (v < 0) continue;
      if (pass == 0) atomicAdd(&hist[(v >> 8) & 255], 1);
      else if ((v >> 8) == prefix) atomicAdd(&hist[v & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
      if (pass == 0) {
        int n = 0;
        for (int b = 0; b < 256; ++b) n += hist[b];
        rank = n / 2;
        s_sel[1] = n;
      }
      int acc = 0, b = 0;
      for (; b < 255; ++b) {
        if (acc + hist[b] > rank) break;
        acc += hist[b];
      }
      s_sel[0] = b;
      rank -= acc;
    }
    __syncthreads();
    prefix = pass == 0 ? s_sel[0] : (prefix << 8) | s_sel[0];
    __syncthreads();
  }
  if (s_sel[1] == 0) return;
  const float th = __fmul_rn(1.5f * 1.4f, (float)prefix);
  for (int i = tid; i < n_l; i += 256) {
    const int v = fs[i];
    if (v >= 0 && !((float)v < th)) {
      uright[(size_t)frame * cap_l + i] = -1.0f;
      depth[(size_t)frame * cap_l + i] = -1.0f;
    }
  }
}

// ---- SearchByBoW (:206-388, 390-565, 996-1163, 1180-1363) ------------------------------------
// The host walks the two feature vectors (node ids are a few hundred ints) and emits one query
// per valid side-1 feature of a common node: {idx1, first side-2 item, item count, row offset}.
// Phase A (k_bow_candidates, warp per query): distances to the node's side-2 features, stored in
// vector order as (dist << 16 | idx2), dist = 0xFFFF for features that are not valid.
// Phase B (k_bow_resolve, one CTA; warp 0 walks the queries in the reference's order): skip
// side-2 features already matched (:296-297, :1061), strict-< best / second from 256, acceptance,
// rotation histogram; then three maxima and the removal pass in parallel.
struct BowQuery {
  int32_t idx1, t0, cnt, off;  // side-1 feature, first side-2 item, item count, row offset (all batch-global)
  int32_t o2, pad0, pad1, pad2;  // first side-2 feature of the query's pair
};

__global__ void __launch_bounds__(256) k_bow_candidates(const BowQuery* __restrict__ q, int nq,
                                                        const uint8_t* __restrict__ d1, const uint8_t* __restrict__ d2,
                                                        const uint8_t* __restrict__ valid2,
                                                        const int32_t* __restrict__ items2, uint32_t* __restrict__ rows) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= nq) return;
  const BowQuery bq = q[i];
  const uint4* qp = reinterpret_cast<const uint4*>(d1 + (size_t)bq.idx1 * 32);
  const uint4 qa = __ldg(qp), qb = __ldg(qp + 1);
  for (int c = lane; c < bq.cnt; c += 32) {
    const int idx2 = items2[bq.t0 + c];  // batch-global
    uint32_t dist = 0xFFFFu;
    if (valid2[idx2]) {
      const uint4* tp = reinterpret_cast<const uint4*>(d2 + (size_t)idx2 * 32);
      dist = (uint32_t)hamming256(qa, qb, __ldg(tp), __ldg(tp + 1));
    }
    rows[bq.off + c] = dist << 16 | (uint32_t)(idx2 - bq.o2);  // index local to the pair: fits 16 bits
  }
}

#define BOW_SMEM_ROWS 52000  // candidate entries staged in shared memory (203 KB of the SM's 227 KB; one CTA per pair)

// One CTA per pair of the batch.  pair_info[p] = {first query, query count, first row, row count, o1, o2, n2, 0}.
#define BOW_NT 1024
__global__ void __launch_bounds__(BOW_NT) k_bow_resolve(const BowQuery* __restrict__ q_all, const int32_t* __restrict__ pair_info,
                                                     const uint32_t* __restrict__ rows_all, const float* __restrict__ angle1,
                                                     const float* __restrict__ angle2, float nnratio, int check_ori,
                                                     int max_dist, int max_nodes, int32_t* __restrict__ matches12,
                                                     int32_t* __restrict__ matches21, int32_t* __restrict__ q_bin_all,
                                                     int* __restrict__ nmatches_out) {
  // [max_nodes + 1] first query of every vocabulary node, [n2 / 32 + 1] matched bits of side 2, then the staged rows
  extern __shared__ uint32_t s_bow[];
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_keep[HISTO_LENGTH];
  __shared__ int s_nmatch, s_nn;
  __shared__ int s_wcount[BOW_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int32_t* info = pair_info + 8 * blockIdx.x;
  const int q0 = info[0], nq = info[1], r0 = info[2], total_rows = info[3], o2 = info[5], n2 = info[6];
  const BowQuery* q = q_all + q0;
  int32_t* q_bin = q_bin_all + q0;
  const int nwords = n2 / 32 + 1;
  int* s_nodes = reinterpret_cast<int*>(s_bow);
  uint32_t* s_taken = s_bow + max_nodes + 1;
  uint32_t* s_rows = s_taken + nwords;
  const bool staged = total_rows <= BOW_SMEM_ROWS;
  for (int i = tid; i < nwords; i += BOW_NT) s_taken[i] = 0;
  if (staged)
    for (int i = tid; i < total_rows; i += BOW_NT) s_rows[i] = rows_all[r0 + i];
  if (tid < HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) { s_nmatch = 0; s_nn = 0; }
  __syncthreads();
  const uint32_t* R = staged ? s_rows : rows_all + r0;  // indexed by row offsets relative to the pair
  // The queries arrive in node order (the host walks the two feature vectors) and a query's candidates are the side-2
  // features of ITS node.  The walk's only state is the matched bit of each side-2 feature (:270-271, :311), and a
  // feature belongs to exactly one node of the feature vector: queries of different nodes never see each other's
  // effects.  So the ordered walk decomposes into independent per-node walks - a warp takes a node and walks its
  // queries in order, 32 nodes at a time - with the same matches as the sequential loop over all nodes.
  for (int base = 0; base < nq; base += BOW_NT) {  // ordered list of the nodes' first queries (t0 = first side-2 item)
    const int j = base + tid;
    const bool first = j < nq && (j == 0 || q[j].t0 != q[j - 1].t0);
    const unsigned bal = __ballot_sync(0xffffffffu, first);
    if (lane == 0) s_wcount[w] = __popc(bal);
    __syncthreads();
    int pos = s_nn + __popc(bal & ((1u << lane) - 1u));
    for (int ww = 0; ww < w; ++ww) pos += s_wcount[ww];
    if (first) s_nodes[pos] = j;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int ww = 0; ww < BOW_NT / 32; ++ww) t += s_wcount[ww];
      s_nn += t;
    }
    __syncthreads();
  }
  const int n_nodes = s_nn;
  if (tid == 0) s_nodes[n_nodes] = nq;
  __syncthreads();
  {
    int nm = 0;
    for (int node = w; node < n_nodes; node += BOW_NT / 32) {
      const int j_end = s_nodes[node + 1];
      int j = s_nodes[node];
      int cnt = q[j].cnt, off = q[j].off - r0;
      for (; j < j_end; ++j) {
        int cnt_n = 0, off_n = 0;
        if (j + 1 < j_end) { cnt_n = q[j + 1].cnt; off_n = q[j + 1].off - r0; }  // in flight during this query
        // key = dist << 16 | position in the node's vector: the strict-< scan order (:311-321)
        uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
        for (int c = lane; c < cnt; c += 32) {
          const uint32_t e = R[off + c];
          const uint32_t idx2 = e & 0xFFFFu;
          if ((e >> 16) == 0xFFFFu || (s_taken[idx2 >> 5] >> (idx2 & 31) & 1u)) continue;
          const uint32_t key = (e & 0xFFFF0000u) | (uint32_t)c;
          second = min(second, max(best, key));
          best = min(best, key);
        }
        warp_top2(best, second);
        int res = -1;
        // nothing left, or the least distance above the gate: no match
        if (best != 0xFFFFFFFFu && (int)(best >> 16) <= max_dist) {
          const int bestDist1 = (int)(best >> 16);
          const int bestDist2 = second == 0xFFFFFFFFu ? 256 : (int)(second >> 16);
          if ((float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {
            res = (int)(R[off + (int)(best & 0xFFFFu)] & 0xFFFFu);
            if (lane == 0) atomicOr(&s_taken[res >> 5], 1u << (res & 31));  // other nodes' bits share the word
            ++nm;
          }
        }
        if (lane == 0) q_bin[j] = res;  // the matched side-2 index for now; the rotation bins follow below
        __syncwarp();
        cnt = cnt_n;
        off = off_n;
      }
    }
    if (lane == 0 && nm) atomicAdd(&s_nmatch, nm);
  }
  __syncthreads();
  // matches and their rotation bins (:330-341), a thread per query
  for (int j = tid; j < nq; j += BOW_NT) {
    const int idx2 = q_bin[j];
    if (idx2 < 0) continue;
    const int idx1 = q[j].idx1;
    matches12[idx1] = idx2;
    int bin = HISTO_LENGTH;  // matched, no orientation bin
    if (check_ori) {
      float rot = __fsub_rn(angle1[idx1], angle2[o2 + idx2]);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
      if (bin == HISTO_LENGTH) bin = 0;
      atomicAdd(&s_hist[bin], 1);
    }
    q_bin[j] = bin;
  }
  __syncthreads();
  if (tid == 0 && check_ori) three_maxima_keep(s_hist, s_keep);
  __syncthreads();
  // removal pass (:365-383) and the side-2 view of the matches
  const int o1 = info[4];
  for (int j = tid; j < nq; j += BOW_NT) {
    const int bin = q_bin[j];
    if (bin < 0) continue;
    const int idx1 = q[j].idx1;
    if (check_ori && !s_keep[bin]) {
      matches12[idx1] = -1;
      atomicSub(&s_nmatch, 1);
    } else {
      matches21[o2 + matches12[idx1]] = idx1 - o1;
    }
  }
  __syncthreads();
  if (tid == 0) nmatches_out[blockIdx.x] = s_nmatch;
}

// ---- SearchForTriangulation (:1364-1720) ------------------------------------------------------
// The reference never sets vbMatched2 (:1452), so every key-frame-1 feature is independent: a warp
// per query evaluates the node's key-frame-2 features in parallel (skip tests :1481-1500, distance
// gate :1505, epipole distance :1509-1520, CheckDistEpipolarLine :167-184) and the scan's result is
// the minimum distance among the passing candidates, the LAST one on ties (`dist>bestDist` lets
// equal distances through) = min over the key dist << 16 | (0xFFFF - position).
struct TriSide {
  const orbx_keypoint* k;
  const uint8_t* d;
  const int32_t* has_mp;
  const int32_t* cam;
  const float* uright;
};

__global__ void __launch_bounds__(256) k_tri_match(const BowQuery* __restrict__ q, int nq, TriSide s1, TriSide s2,
                                                   const int32_t* __restrict__ items2, const float* __restrict__ consts_all,
                                                   int only_stereo, int check_ori, int32_t* __restrict__ matches12,
                                                   int32_t* __restrict__ q_bin, int* __restrict__ hist_all) {
  // per pair: consts = F12s[18], epipoles[4], scale_factors2[16], level_sigma2_2[16] (54 floats, stride 56);
  // all feature indices are batch-global, bq.o2 = first key-frame-2 feature of the pair, bq.pad0 = pair
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= nq) return;
  const BowQuery bq = q[i];
  const float* consts = consts_all + 56 * bq.pad0;
  int* hist = hist_all + HISTO_LENGTH * bq.pad0;
  const int idx1 = bq.idx1;
  const orbx_keypoint kp1 = s1.k[idx1];
  const int camIdx1 = s1.cam[idx1];
  const bool bStereo1 = s1.uright[idx1] >= 0;
  const float* F12 = consts + 9 * camIdx1;
  // line of kp1 in image 2 (:170-172), float, left to right, no contraction
  const float a = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, F12[0]), __fmul_rn(kp1.y, F12[3])), F12[6]);
  const float b = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, F12[1]), __fmul_rn(kp1.y, F12[4])), F12[7]);
  const float c = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, F12[2]), __fmul_rn(kp1.y, F12[5])), F12[8]);
  const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
  const uint4* qp = reinterpret_cast<const uint4*>(s1.d + (size_t)idx1 * 32);
  const uint4 qa = __ldg(qp), qb = __ldg(qp + 1);
  uint32_t best = 0xFFFFFFFFu;
  for (int p = lane; p < bq.cnt; p += 32) {
    const int idx2 = items2[bq.t0 + p];
    if (s2.has_mp[idx2] || s2.cam[idx2] != camIdx1) continue;
    const bool bStereo2 = s2.uright[idx2] >= 0;
    if (only_stereo && !bStereo2) continue;
    const uint4* tp = reinterpret_cast<const uint4*>(s2.d + (size_t)idx2 * 32);
    const int dist = hamming256(qa, qb, __ldg(tp), __ldg(tp + 1));
    if (dist > TH_LOW) continue;
    const orbx_keypoint kp2 = s2.k[idx2];
    if (!bStereo1 && !bStereo2) {
      const float dx = __fsub_rn(consts[18 + 2 * camIdx1], kp2.x), dy = __fsub_rn(consts[19 + 2 * camIdx1], kp2.y);
      if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, consts[22 + kp2.octave])) continue;
    }
    if (den == 0.f) continue;
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, kp2.x), __fmul_rn(b, kp2.y)), c);
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    if (!((double)dsqr < __dmul_rn(3.84, (double)consts[38 + kp2.octave]))) continue;
    best = min(best, (uint32_t)dist << 16 | (uint32_t)(0xFFFF - p));
  }
  best = __reduce_min_sync(0xffffffffu, best);
  if (lane == 0) {
    int bin = -1;
    if (best != 0xFFFFFFFFu) {
      const int idx2 = items2[bq.t0 + (0xFFFF - (int)(best & 0xFFFFu))];
      matches12[idx1] = idx2 - bq.o2;
      bin = HISTO_LENGTH;
      if (check_ori) {
        float rot = __fsub_rn(kp1.angle, s2.k[idx2].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
        if (bin == HISTO_LENGTH) bin = 0;
        atomicAdd(&hist[bin], 1);
      }
    }
    q_bin[i] = bin;
  }
}

// rotation filter (:1701-1716) and the match count
__global__ void __launch_bounds__(256) k_tri_finish(const BowQuery* __restrict__ q_all, const int32_t* __restrict__ pair_q,
                                                    int check_ori, const int32_t* __restrict__ q_bin_all,
                                                    const int* __restrict__ hist_all, int32_t* __restrict__ matches12,
                                                    int* __restrict__ nmatches_out) {
  __shared__ int s_keep[HISTO_LENGTH];
  __shared__ int s_n;
  const int p = blockIdx.x, q0 = pair_q[p], nq = pair_q[p + 1] - q0;
  const BowQuery* q = q_all + q0;
  const int32_t* q_bin = q_bin_all + q0;
  if (threadIdx.x == 0) {
    s_n = 0;
    if (check_ori) three_maxima_keep(hist_all + HISTO_LENGTH * p, s_keep);
  }
  __syncthreads();
  int mine = 0;
  for (int j = threadIdx.x; j < nq; j += 256) {
    const int bin = q_bin[j];
    if (bin < 0) continue;
    if (check_ori && !s_keep[bin]) matches12[q[j].idx1] = -1;
    else ++mine;
  }
  atomicAdd(&s_n, mine);
  __syncthreads();
  if (threadIdx.x == 0) nmatches_out[p] = s_n;
}

// ---- SearchByProjection(Frame&, vector<MapPoint*>&, th) ------------------------------------
// Phase A (warp per map point, whole grid): window query at levels [pred-1, pred], stereo gate
// (:111-116), distances; candidates kept in traversal order as dist<<20 | octave<<16 | idx.
// `count_only` pass sizes the ragged rows, then a scan gives row offsets.
// :151-157 compares the float against the DOUBLE literal 0.998; (float)0.998 rounds up to 0.99800003, so
// view_cos == 0.998f is "greater" there: compare in double
__device__ __forceinline__ float radius_by_viewing_cos(float view_cos) { return (double)view_cos > 0.998 ? 2.5f : 4.0f; }

__global__ void __launch_bounds__(256) k_proj_candidates(const orbx_keypoint* __restrict__ k, const uint8_t* __restrict__ desc,
                                                         const float* __restrict__ u_right, orbm_bounds b,
                                                         const int* __restrict__ grid_start,
                                                         const uint16_t* __restrict__ grid_items,
                                                         const float* __restrict__ scale_factors,
                                                         const orbm_mappoint* __restrict__ mp,
                                                         const uint8_t* __restrict__ mp_desc, int nmp, float th,
                                                         int count_only, int* __restrict__ row_cnt,
                                                         const int* __restrict__ row_off, uint32_t* __restrict__ rows,
                                                         long long rows_cap = 0) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nmp) return;
  // rows_cap > 0: row buffer sized from earlier calls (see k_query_candidates); rows that would not fit are not written
  if (!count_only && rows_cap > 0 && (long long)row_off[i] + row_cnt[i] > rows_cap) return;
  const orbm_mappoint p = mp[i];
  int cnt = 0;
  if (p.track_in_view && !p.bad) {
    GridView gv;
    gv.start = grid_start;
    gv.items = grid_items;
    gv.min_x = b.min_x;
    gv.min_y = b.min_y;
    gv.inv_w = __fdiv_rn((float)GRID_COLS, __fsub_rn(b.max_x, b.min_x));
    gv.inv_h = __fdiv_rn((float)GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
    float r = radius_by_viewing_cos(p.view_cos);
    if (th != 1.0f) r = __fmul_rn(r, th);
    const float rs = __fmul_rn(r, scale_factors[p.level]);
    const uint4* q = reinterpret_cast<const uint4*>(mp_desc + (size_t)i * 32);
    const uint4 qa = __ldg(q), qb = __ldg(q + 1);
    uint32_t* row = count_only ? nullptr : rows + row_off[i];
    grid_query(gv, k, p.proj_x, p.proj_y, rs, p.level - 1, p.level, [&](bool ok, int idx) {
      if (ok && u_right) {
        const float ur = u_right[idx];
        if (ur > 0 && fabsf(__fsub_rn(p.proj_xr, ur)) > rs) ok = false;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok && !count_only) {
        const uint4* tp = reinterpret_cast<const uint4*>(desc + (size_t)idx * 32);
        const int dist = hamming256(qa, qb, __ldg(tp), __ldg(tp + 1));
        row[cnt + __popc(m & ((1u << lane) - 1u))] = (uint32_t)dist << 20 | (uint32_t)(k[idx].octave & 15) << 16 | (uint32_t)idx;
      }
      cnt += __popc(m);
    });
  }
  if (lane == 0 && count_only) row_cnt[i] = cnt;
}

// single-CTA exclusive scan (n up to a few hundred thousand)
__global__ void __launch_bounds__(1024) k_scan_exclusive(const int* __restrict__ in, int* __restrict__ out, int n,
                                                         int* __restrict__ total) {
  // one CTA: a contiguous chunk per thread, warp scans of the chunk sums, a scan of the 32 warp sums by warp 0
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chunk = (n + 1023) / 1024;
  const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += in[i];
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int ws = s_warp[lane];
    int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - ws;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  int run = s_warp[w] + incl - s;
  for (int i = lo; i < hi; ++i) { const int v = in[i]; out[i] = run; run += v; }
}

// Phase B: ordered resolve by one warp, map points ascending: occupancy skip (:107-109),
// best/second with levels (:122-135), TH_HIGH + same-level ratio (:138-141), assignment (:143).
// Ordered resolve of SearchByProjection(Frame, vector<MapPoint*>) (:91-147), 32 points per step.
// A point's outcome is a function of the two least keys among its candidates that hold no point, and
// occupancy only grows during the walk.  32 warps, a warp per point of the step: every pending warp
// scans its row under the occupancy as it stands; warp 0 then commits the longest prefix of the step
// whose answers do not depend on each other - committing lanes claim their keypoint (occupancy byte
// 2 + lane, lowest lane per keypoint), a lane is in conflict when one of its two keypoints is claimed
// by a lower lane - and the rest scan again.  The lowest pending point of a step never conflicts, so
// each round commits at least one point, in order; keypoints that stay free (Observations()==0) are
// overwritten by later points exactly as in the sequential loop.  (SearchLocalPoints projects several
// map points onto every keypoint, so a static best per point computed up front is mostly stale by the
// time the walk reaches it: that single-warp speculative form measured 8.9 ms on configs[3]'s 20 000
// points, the plain sequential warp 6.8 ms, this form 1.2 ms.)
__global__ void __launch_bounds__(1024) k_proj_resolve_cta(const int* __restrict__ row_cnt, const int* __restrict__ row_off,
                                                           const uint32_t* __restrict__ rows,
                                                           const int32_t* __restrict__ mp_obs, int nmp, int n, float nnratio,
                                                           int32_t* __restrict__ frame_mp, const int32_t* __restrict__ frame_mp_obs,
                                                           uint8_t* __restrict__ held_global, int held_in_smem,
                                                           int* __restrict__ nmatches_out,
                                                           const int* __restrict__ rows_needed = nullptr, long long rows_cap = 0) {
  extern __shared__ __align__(16) uint8_t s_held[];
  if (rows_needed && *rows_needed > rows_cap) return;  // the row buffer overflowed: the host repeats the call
  // per-round exchange, double-buffered by round parity: two barriers per round instead of three
  __shared__ int s_bidx[2][32], s_sidx[2][32], s_will[2][32];
  __shared__ unsigned s_done[2], s_left[2];
  int rp = 0;  // round parity
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint8_t* held = held_in_smem ? s_held : held_global;
  for (int i = tid; i < n; i += 1024) held[i] = (frame_mp[i] >= 0 && frame_mp_obs && frame_mp_obs[i] > 0) ? 1 : 0;
  __syncthreads();
  int nmatches = 0;  // warp 0's count
  // Nothing on the walk's critical path waits for HBM: a warp keeps the first PROJ_SLOTS * 32
  // candidates of its point in registers (longer rows read the rest in place), the entries of the
  // next step's point are fetched while this step resolves, and the counts / offsets one step
  // further ahead.
  constexpr int PROJ_SLOTS = 4;
  constexpr uint32_t NO_ENTRY = 0xFFFFFFFFu;  // entries are dist << 20 | level << 16 | index < 2^29
  auto entry_at = [&](const uint32_t (&ent)[PROJ_SLOTS], const uint32_t* row, int pos) -> uint32_t {
    if (pos >= PROJ_SLOTS * 32) return row[pos];
    uint32_t v = ent[0];
#pragma unroll
    for (int k = 1; k < PROJ_SLOTS; ++k)
      if ((pos >> 5) == k) v = ent[k];
    return __shfl_sync(0xffffffffu, v, pos & 31);
  };
  uint32_t ent[PROJ_SLOTS], ent_n[PROJ_SLOTS];
  int cnt = w < nmp ? row_cnt[w] : 0, off = w < nmp ? row_off[w] : 0;
  int cnt_n = 32 + w < nmp ? row_cnt[32 + w] : 0, off_n = 32 + w < nmp ? row_off[32 + w] : 0;
#pragma unroll
  for (int k = 0; k < PROJ_SLOTS; ++k) ent[k] = k * 32 + lane < cnt ? rows[off + k * 32 + lane] : NO_ENTRY;
  for (int b0 = 0; b0 < nmp; b0 += 32) {
#pragma unroll
    for (int k = 0; k < PROJ_SLOTS; ++k) ent_n[k] = k * 32 + lane < cnt_n ? rows[off_n + k * 32 + lane] : NO_ENTRY;
    const int i2 = b0 + 64 + w;
    const int cnt_nn = i2 < nmp ? row_cnt[i2] : 0, off_nn = i2 < nmp ? row_off[i2] : 0;
    const uint32_t* row = rows + off;
    bool pend = cnt > 0;
    // warp 0 keeps the per-point constants of the step, lane = point
    const int li = b0 + lane;
    const uint8_t hval = (w == 0 && li < nmp && mp_obs) ? (mp_obs[li] > 0 ? 1 : 0) : 1;
    for (;;) {
      int bidx = -1, sidx = -1, will = 0;
      if (pend) {
        uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < PROJ_SLOTS; ++k) {
          const uint32_t e = ent[k];
          if (e == NO_ENTRY || held[e & 0xFFFFu]) continue;
          const uint32_t key = (e >> 20) << 16 | (uint32_t)(k * 32 + lane);
          second = min(second, max(best, key));
          best = min(best, key);
        }
        for (int c = PROJ_SLOTS * 32 + lane; c < cnt; c += 32) {
          const uint32_t e = row[c];
          if (held[e & 0xFFFFu]) continue;
          const uint32_t key = (e >> 20) << 16 | (uint32_t)c;
          second = min(second, max(best, key));
          best = min(best, key);
        }
        warp_top2(best, second);
        // nothing free, or the least distance above TH_HIGH (it can only grow): the point is finished
        if (best == 0xFFFFFFFFu || (int)(best >> 16) > TH_HIGH) {
          pend = false;
        } else {
          const int bestDist = (int)(best >> 16);
          const uint32_t eb = entry_at(ent, row, (int)(best & 0xFFFFu));
          const int bestLevel = (int)(eb >> 16 & 15u);
          bidx = (int)(eb & 0xFFFFu);
          int bestDist2 = 256, bestLevel2 = -1;
          if (second != 0xFFFFFFFFu) {
            const uint32_t es = entry_at(ent, row, (int)(second & 0xFFFFu));
            bestDist2 = (int)(second >> 16);
            bestLevel2 = (int)(es >> 16 & 15u);
            sidx = (int)(es & 0xFFFFu);
          }
          will = !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2));
        }
      }
      if (lane == 0) { s_bidx[rp][w] = bidx; s_sidx[rp][w] = sidx; s_will[rp][w] = will; }
      __syncthreads();
      if (w == 0) {
        const int b = s_bidx[rp][lane], sx = s_sidx[rp][lane], wl = s_will[rp][lane];
        const bool p = b >= 0;  // pending points have a best keypoint, free under the current occupancy
        const bool wants = p && wl;
        const unsigned peers = __match_any_sync(0xffffffffu, wants ? b : -1 - lane);
        const bool claims = wants && (__ffs(peers) - 1 == lane);
        if (claims) held[b] = (uint8_t)(2 + lane);
        __syncwarp();
        bool conflict = false;
        if (p) {
          const int vb = held[b], vs = sx >= 0 ? held[sx] : 0;
          conflict = (vb >= 2 && vb - 2 < lane) || (vs >= 2 && vs - 2 < lane);
        }
        const unsigned cm = __ballot_sync(0xffffffffu, conflict);
        const int first = cm ? __ffs(cm) - 1 : 32;
        __syncwarp();
        if (claims && lane >= first) held[b] = 0;  // not this round: withdraw the claim
        const bool done = p && lane < first;
        const bool commit = done && wl;
        if (commit) {
          frame_mp[b] = li;
          held[b] = hval;
        }
        nmatches += __popc(__ballot_sync(0xffffffffu, commit));
        const unsigned dm = __ballot_sync(0xffffffffu, done), pm = __ballot_sync(0xffffffffu, p);
        if (lane == 0) { s_done[rp] = dm; s_left[rp] = pm & ~dm; }
      }
      __syncthreads();
      if (s_done[rp] >> w & 1u) pend = false;
      const unsigned left = s_left[rp];
      rp ^= 1;
      if (!left) break;
    }
    cnt = cnt_n; off = off_n;
    cnt_n = cnt_nn; off_n = off_nn;
#pragma unroll
    for (int k = 0; k < PROJ_SLOTS; ++k) ent[k] = ent_n[k];
  }
  if (tid == 0) *nmatches_out = nmatches;
}

constexpr int RESOLVE_WIN = 2048;  // candidate entries per shared-memory window of the ordered resolves


// ---- pose-based SearchByProjection overloads: generic projected queries ----------------------
// The host projects the source points with the reference's own float arithmetic (it is a few
// thousand flops and must match OpenCV's evaluation order); the device does what costs: window
// query on the per-camera grid, stereo gate, Hamming distances, ordered resolve with occupancy,
// rotation histogram.  One query = one projected source point.
struct ProjQuery {
  float u, v, radius;   // projection and search radius
  float ur;             // predicted right coordinate (u - mbf*invz); used when use_ur != 0
  float angle;          // source keypoint angle (rotation histogram)
  int32_t min_level, max_level;
  int32_t cam;          // which per-camera grid
  int32_t src;          // index written into frame_mp on acceptance
  int32_t obs;          // Observations()>0 of the source map point
  int32_t use_ur;
};

__global__ void __launch_bounds__(256) k_query_candidates(const orbx_keypoint* __restrict__ k, const uint8_t* __restrict__ desc,
                                                          const float* __restrict__ u_right, orbm_bounds b,
                                                          const int* __restrict__ grid_start,
                                                          const uint16_t* __restrict__ grid_items, int n_frame_kps,
                                                          const ProjQuery* __restrict__ q, const uint8_t* __restrict__ q_desc,
                                                          int nq, int count_only, int* __restrict__ row_cnt,
                                                          const int* __restrict__ row_off, uint32_t* __restrict__ rows,
                                                          long long rows_cap = 0) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nq) return;
  const ProjQuery p = q[i];
  GridView gv;
  gv.start = grid_start + (size_t)p.cam * (GRID_CELLS + 1);
  gv.items = grid_items + (size_t)p.cam * n_frame_kps;
  gv.min_x = b.min_x;
  gv.min_y = b.min_y;
  gv.inv_w = __fdiv_rn((float)GRID_COLS, __fsub_rn(b.max_x, b.min_x));
  gv.inv_h = __fdiv_rn((float)GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
  const uint4* qd = reinterpret_cast<const uint4*>(q_desc + (size_t)p.src * 32);
  const uint4 qa = __ldg(qd), qb = __ldg(qd + 1);
  // count_only 1: count; 0: fill rows + row_off[i].  rows_cap > 0: the row buffer was sized from a previous call's need
  // instead of this call's exact total (no host round trip between the count and the fill); a query whose rows would
  // not fit writes nothing and the host repeats the call with the exact size (the scan's total says so).
  if (!count_only && rows_cap > 0 && (long long)row_off[i] + row_cnt[i] > rows_cap) return;
  uint32_t* row = count_only ? nullptr : rows + row_off[i];
  int cnt = 0;
  grid_query(gv, k, p.u, p.v, p.radius, p.min_level, p.max_level, [&](bool ok, int idx) {
    if (ok && p.use_ur && u_right) {
      const float ur = u_right[idx];
      if (ur > 0 && fabsf(__fsub_rn(p.ur, ur)) > p.radius) ok = false;  // :3571-3577
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok && !count_only) {
      const uint4* tp = reinterpret_cast<const uint4*>(desc + (size_t)idx * 32);
      const int dist = hamming256(qa, qb, __ldg(tp), __ldg(tp + 1));
      row[cnt + __popc(m & ((1u << lane) - 1u))] = (uint32_t)dist << 16 | (uint32_t)idx;
    }
    cnt += __popc(m);
  });
  if (lane == 0 && count_only) row_cnt[i] = cnt;
}

// ORBmatcher::Fuse (:1986-2190), the search of one (map point, camera) query: window query at levels
// [predicted-1, predicted], reprojection chi-square gate (:2106-2131: 7.8 with a right coordinate, 5.99
// without), strict-< best Hamming distance in traversal order (:2139-2143), accepted when <= TH_LOW.
// Nothing here depends on other queries, so a warp per query and no resolve pass.
struct LevelTable {
  float v[ORBX_MAX_LEVELS];
};

__global__ void __launch_bounds__(256) k_fuse_match(const orbx_keypoint* __restrict__ k, const uint8_t* __restrict__ desc,
                                                    const float* __restrict__ u_right, orbm_bounds b,
                                                    const int* __restrict__ grid_start,
                                                    const uint16_t* __restrict__ grid_items, int n_kps,
                                                    const ProjQuery* __restrict__ q, const uint8_t* __restrict__ q_desc,
                                                    int nq, LevelTable inv_sigma2, int gate, int th_dist,
                                                    int32_t* __restrict__ best_idx) {
  // gate != 0: the pose-based overload (:1986); gate == 0: the Sim3 overload (:2211), which has no reprojection test
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= nq) return;
  const ProjQuery p = q[i];
  GridView gv;
  gv.start = grid_start + (size_t)p.cam * (GRID_CELLS + 1);
  gv.items = grid_items + (size_t)p.cam * n_kps;
  gv.min_x = b.min_x;
  gv.min_y = b.min_y;
  gv.inv_w = __fdiv_rn((float)GRID_COLS, __fsub_rn(b.max_x, b.min_x));
  gv.inv_h = __fdiv_rn((float)GRID_ROWS, __fsub_rn(b.max_y, b.min_y));
  const uint4* qd = reinterpret_cast<const uint4*>(q_desc + (size_t)p.src * 32);
  const uint4 qa = __ldg(qd), qb = __ldg(qd + 1);
  uint32_t best = 0xFFFFFFFFu;  // dist << 16 | traversal position
  int my_idx = -1, cnt = 0;
  grid_query(gv, k, p.u, p.v, p.radius, p.min_level, p.max_level, [&](bool ok, int idx) {
    if (ok && gate) {
      const orbx_keypoint kp = k[idx];
      const float ex = __fsub_rn(p.u, kp.x), ey = __fsub_rn(p.v, kp.y);
      float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
      const float kpr = u_right[idx];
      double limit = 5.99;
      if (kpr >= 0) {
        const float er = __fsub_rn(p.ur, kpr);
        e2 = __fadd_rn(e2, __fmul_rn(er, er));
        limit = 7.8;
      }
      if ((double)__fmul_rn(e2, inv_sigma2.v[kp.octave]) > limit) ok = false;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint4* tp = reinterpret_cast<const uint4*>(desc + (size_t)idx * 32);
      const uint32_t key = (uint32_t)hamming256(qa, qb, __ldg(tp), __ldg(tp + 1)) << 16 |
                           (uint32_t)(cnt + __popc(m & ((1u << lane) - 1u)));
      if (key < best) { best = key; my_idx = idx; }
    }
    cnt += __popc(m);
  });
  const uint32_t gb = __reduce_min_sync(0xffffffffu, best);
  const unsigned owner = __ballot_sync(0xffffffffu, best == gb && my_idx >= 0);
  const int idx = __shfl_sync(0xffffffffu, my_idx, owner ? __ffs(owner) - 1 : 0);
  if (lane == 0) best_idx[2 * p.src + p.cam] = (gb != 0xFFFFFFFFu && (int)(gb >> 16) <= th_dist) ? idx : -1;
}

// Ordered resolve, best match only (:3558-3637, :3886-3934): queries ascending; candidates that
// already hold a point are skipped (any_point_blocks: a15 blocks on any held point, a14 only on
// points with Observations()>0); accept best <= th_dist; rotation histogram + three maxima.
// The ordered resolve is one dependent chain over the queries, so everything a step needs sits on chip
// before the step starts: the query fields of 32 queries are fetched by the 32 lanes at once and
// broadcast by shuffle, the candidate rows (contiguous in `rows`, exclusive-scan offsets) stream
// through a shared-memory window, the occupancy bytes live in shared memory when they fit, and the
// rotation bins (which need the matched keypoint's angle from HBM) are computed after the walk.
__global__ void __launch_bounds__(32) k_query_resolve_best(const ProjQuery* __restrict__ q, const int* __restrict__ row_cnt,
                                                           const int* __restrict__ row_off, const uint32_t* __restrict__ rows,
                                                           int total_rows, int nq, int n, const orbx_keypoint* __restrict__ k,
                                                           int th_dist, int check_ori, int any_point_blocks,
                                                           int32_t* __restrict__ frame_mp, const int32_t* __restrict__ frame_mp_obs,
                                                           uint8_t* __restrict__ held_global, int held_in_smem,
                                                           int32_t* __restrict__ acc_idx, int32_t* __restrict__ acc_bin,
                                                           int* __restrict__ nmatches_out,
                                                           const int* __restrict__ rows_needed = nullptr, long long rows_cap = 0) {
  if (rows_needed && *rows_needed > rows_cap) return;  // see k_query_resolve_cta
  extern __shared__ __align__(16) uint8_t s_held[];
  __shared__ uint32_t s_rows[RESOLVE_WIN];
  __shared__ int s_hist[HISTO_LENGTH];
  const int lane = threadIdx.x;
  uint8_t* held = held_in_smem ? s_held : held_global;
  for (int i = lane; i < n; i += 32)
    held[i] = frame_mp[i] >= 0 && (any_point_blocks || (frame_mp_obs && frame_mp_obs[i] > 0)) ? 1 : 0;
  if (lane < HISTO_LENGTH) s_hist[lane] = 0;
  __syncwarp();
  int nmatches = 0, nacc = 0;
  int win_base = 0, win_end = 0;  // [win_base, win_end) of `rows` is resident in s_rows
  // Consecutive queries with the same source point form one group (the Sim3 overload projects a
  // point into both cameras and keeps the best over cameras, :629-735); elsewhere groups are single.
  int gsrc = 0, gobs = 0, gidx = -1;
  float gangle = 0.0f;
  uint32_t gbest = 0xFFFFFFFFu;  // dist << 16 | (anything): only the distance decides across rows (strict <)
  bool open = false;
  int l_src = 0, l_obs = 0, l_cnt = 0, l_off = 0;
  float l_angle = 0.0f;
  for (int j = 0; j <= nq; ++j) {
    if ((j & 31) == 0 && j < nq) {
      const int jj = j + lane;
      if (jj < nq) {
        l_src = q[jj].src; l_obs = q[jj].obs; l_angle = q[jj].angle;
        l_cnt = row_cnt[jj]; l_off = row_off[jj];
      }
    }
    int src = 0, cnt = 0, off = 0;
    if (j < nq) {
      src = __shfl_sync(0xffffffffu, l_src, j & 31);
      cnt = __shfl_sync(0xffffffffu, l_cnt, j & 31);
      off = __shfl_sync(0xffffffffu, l_off, j & 31);
    }
    if (open && (j == nq || src != gsrc)) {
      // the group is complete: accept its best (TH_HIGH / ORBdist / TH_LOW gate)
      open = false;
      if (gbest != 0xFFFFFFFFu && (int)(gbest >> 16) <= th_dist) {
        if (lane == 0) {
          frame_mp[gidx] = gsrc;
          held[gidx] = any_point_blocks ? 1 : (gobs ? 1 : 0);
          if (check_ori) {
            acc_idx[nacc] = gidx;
            acc_bin[nacc] = __float_as_int(gangle);  // the source angle; turned into the bin below
          }
        }
        ++nacc;
        ++nmatches;
        __syncwarp();
      }
    }
    if (j == nq) break;
    if (!open) {
      open = true;
      gsrc = src;
      gobs = __shfl_sync(0xffffffffu, l_obs, j & 31);
      gangle = __shfl_sync(0xffffffffu, l_angle, j & 31);
      gbest = 0xFFFFFFFFu;
      gidx = -1;
    }
    if (cnt == 0) continue;
    const uint32_t* row;
    if (cnt > RESOLVE_WIN) {
      row = rows + off;  // longer than the window: read in place
    } else {
      if (off < win_base || off + cnt > win_end) {
        __syncwarp();
        win_base = off;
        win_end = min(off + RESOLVE_WIN, total_rows);
        const int len = win_end - win_base;
#pragma unroll 8
        for (int c = lane; c < len; c += 32) s_rows[c] = rows[win_base + c];
        __syncwarp();
      }
      row = s_rows + (off - win_base);
    }
    uint32_t best = 0xFFFFFFFFu;
    int my_idx = -1;
    for (int c = lane; c < cnt; c += 32) {
      const uint32_t e = row[c];
      if (held[e & 0xFFFFu]) continue;
      const uint32_t key = (e >> 16) << 16 | (uint32_t)c;  // dist, then traversal position (strict <)
      if (key < best) { best = key; my_idx = (int)(e & 0xFFFFu); }
    }
    const uint32_t mine = best;
    best = __reduce_min_sync(0xffffffffu, best);
    if (best == 0xFFFFFFFFu) continue;
    const unsigned owner = __ballot_sync(0xffffffffu, mine == best);
    const int row_idx = __shfl_sync(0xffffffffu, my_idx, __ffs(owner) - 1);
    if ((best >> 16) < (gbest >> 16) || gbest == 0xFFFFFFFFu) { gbest = best; gidx = row_idx; }
  }
  if (check_ori) {
    __syncwarp();
    // rotation bins of the accepted matches (:3604-3615), lanes over the matches
    for (int e = lane; e < nacc; e += 32) {
      float rot = __fsub_rn(__int_as_float(acc_bin[e]), k[acc_idx[e]].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
      if (bin == HISTO_LENGTH) bin = 0;
      acc_bin[e] = bin;
      atomicAdd(&s_hist[bin], 1);
    }
    __syncwarp();
    int max1 = 0, max2 = 0, max3 = 0, i1_ = -1, i2_ = -1, i3_ = -1;  // every lane computes the same maxima
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      const int sv = s_hist[i];
      if (sv > max1) { max3 = max2; max2 = max1; max1 = sv; i3_ = i2_; i2_ = i1_; i1_ = i; }
      else if (sv > max2) { max3 = max2; max2 = sv; i3_ = i2_; i2_ = i; }
      else if (sv > max3) { max3 = sv; i3_ = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2_ = -1; i3_ = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3_ = -1; }
    int removed = 0;
    for (int e = lane; e < nacc; e += 32) {
      const int bin = acc_bin[e];
      if (bin != i1_ && bin != i2_ && bin != i3_) { frame_mp[acc_idx[e]] = -1; ++removed; }  // :3627-3633
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
  }
  if (lane == 0) *nmatches_out = nmatches;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
struct orbm_matcher {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  long long launches = 0;
  // growable device scratch
  void* buf[12] = {nullptr};
  size_t buf_bytes[12] = {0};
  // DBoW2 vocabulary tree (orbm_set_vocabulary)
  int voc_nodes = 0, voc_L = 0;
  int32_t* voc_child_start = nullptr;
  int32_t* voc_child_ids = nullptr;
  uint8_t* voc_desc = nullptr;
  std::vector<int32_t> voc_word;
  std::vector<double> voc_weight;

  bool check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
  // pinned host staging of the single-call host entry points: every input of a call is packed into ONE block and
  // goes up in ONE copy, every output comes back in one (pageable cudaMemcpyAsync calls cost ~10 us each)
  uint8_t* pin[2] = {nullptr, nullptr};
  size_t pin_bytes[2] = {0, 0};
  size_t rows_hint = 0;  // candidate-row capacity that was enough for the previous calls (run_projected)
  int rows_per_query = 64;  // first guess of that capacity (orbm_debug_set_row_budget)
  uint8_t* pinned(int slot, size_t bytes) {
    if (bytes > pin_bytes[slot]) {
      cudaStreamSynchronize(stream);
      if (pin[slot]) cudaFreeHost(pin[slot]);
      pin[slot] = nullptr;
      pin_bytes[slot] = 0;
      const size_t want = bytes + bytes / 2;
      if (!check(cudaHostAlloc((void**)&pin[slot], want, cudaHostAllocDefault), "cudaHostAlloc(matcher staging)")) return nullptr;
      pin_bytes[slot] = want;
    }
    return pin[slot];
  }
  template <typename T> T* scratch(int slot, size_t count) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (bytes > buf_bytes[slot]) {
      cudaStreamSynchronize(stream);
      cudaFree(buf[slot]);
      buf[slot] = nullptr;
      buf_bytes[slot] = 0;
      if (!check(cudaMalloc(&buf[slot], bytes), "cudaMalloc(matcher scratch)")) return nullptr;
      buf_bytes[slot] = bytes;
    }
    return reinterpret_cast<T*>(buf[slot]);
  }
};

namespace {
thread_local std::string g_mcreate_error;

int bf_launch(orbm_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, float ratio, int th_dist,
              int32_t* d_idx, int32_t* d_d1, int32_t* d_d2) {
  if (nq == 0) return ORBX_OK;
  const int qblocks = (nq + BF_THREADS * BF_QPT - 1) / (BF_THREADS * BF_QPT);
  // Enough target chunks for ~32 CTAs per SM (9 are resident at a time): with the 5.2 CTAs per SM of the first version
  // the SMs that drew 6 set the pace while the others idled (ncu: XU pipe — where POPC issues — 93 % busy while active
  // but only 81 % of the elapsed time).  Chunks stay >= 2 tiles so the query prologue amortises; <= 65536 targets
  // (16-bit local index).
  int nsplit = std::max(1, std::min((nt + 2 * BF_TILE - 1) / (2 * BF_TILE), (148 * 32 + qblocks - 1) / qblocks));
  nsplit = std::max(nsplit, (nt + 65535) / 65536);
  int chunk = nt > 0 ? (nt + nsplit - 1) / nsplit : 1;
  chunk = (chunk + BF_TILE - 1) / BF_TILE * BF_TILE;
  nsplit = nt > 0 ? (nt + chunk - 1) / chunk : 1;
  uint2* partial = m->scratch<uint2>(0, (size_t)nsplit * nq);
  if (!partial) return ORBX_E_CUDA;
  k_bruteforce<<<dim3(qblocks, nsplit), BF_THREADS, 0, m->stream>>>(d_q, nq, d_t, nt, chunk, partial);
  k_bf_merge<<<(nq + 255) / 256, 256, 0, m->stream>>>(partial, nq, nsplit, chunk, ratio, th_dist, d_idx, d_d1, d_d2);
  m->launches += 2;
  return m->check(cudaGetLastError(), "bruteforce launch") ? ORBX_OK : ORBX_E_CUDA;
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int orbm_create(int device, orbm_matcher** out) {
  if (!out) return ORBX_E_INVALID;
  *out = nullptr;
  orbm_matcher* m = new orbm_matcher();
  int ndev = 0;
  if (!m->check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0) {
    g_mcreate_error = m->err.empty() ? "no CUDA device (this library has no CPU fallback)" : m->err;
    delete m;
    return ORBX_E_CUDA;
  }
  if (device >= ndev) { g_mcreate_error = "no such CUDA device"; delete m; return ORBX_E_CUDA; }
  cudaGetDevice(&m->device);
  if (device >= 0) m->device = device;
  OrbDeviceGuard dev_guard(m->device);  // the caller's current device is restored on return
  if (!m->check(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
    g_mcreate_error = m->err;
    delete m;
    return ORBX_E_CUDA;
  }
  m->stream = m->own_stream;
  *out = m;
  return ORBX_OK;
}

void orbm_destroy(orbm_matcher* m) {
  if (m) { cudaFree(m->voc_child_start); cudaFree(m->voc_child_ids); cudaFree(m->voc_desc); }
  if (!m) return;
  cudaStreamSynchronize(m->stream);
  for (auto& b : m->buf) cudaFree(b);
  cudaStreamDestroy(m->own_stream);
  delete m;
}

const char* orbm_last_error(const orbm_matcher* m) { return m ? m->err.c_str() : g_mcreate_error.c_str(); }
int orbm_sync(orbm_matcher* m) {
  if (!m) return ORBX_E_INVALID;
  return m->check(cudaStreamSynchronize(m->stream), "stream synchronize") ? ORBX_OK : ORBX_E_CUDA;
}
long long orbm_launch_count(const orbm_matcher* m) { return m ? m->launches : 0; }
int orbm_set_stream(orbm_matcher* m, void* cuda_stream) {
  if (!m) return ORBX_E_INVALID;
  if (!m->check(cudaStreamSynchronize(m->stream), "stream synchronize")) return ORBX_E_CUDA;
  m->stream = cuda_stream ? (cudaStream_t)cuda_stream : m->own_stream;
  return ORBX_OK;
}

int orbm_distance_pairs_host(orbm_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* out) {
  if (!m || !a || !b || !out || n < 0) return ORBX_E_INVALID;
  if (n == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  uint8_t* da = m->scratch<uint8_t>(1, (size_t)n * 32);
  uint8_t* db = m->scratch<uint8_t>(2, (size_t)n * 32);
  int32_t* dout = m->scratch<int32_t>(3, n);
  if (!da || !db || !dout) return ORBX_E_CUDA;
  cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream);
  cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream);
  k_distance_pairs<<<(n + 255) / 256, 256, 0, m->stream>>>(da, db, n, dout);
  m->launches++;
  cudaMemcpyAsync(out, dout, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, m->stream);
  return m->check(cudaStreamSynchronize(m->stream), "distance_pairs") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_bruteforce_device(orbm_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, float ratio,
                           int th_dist, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2) {
  if (!m || nq < 0 || nt < 0 || (nq && (!d_q || !d_idx || !d_d1 || !d_d2)) || (nt && !d_t)) return ORBX_E_INVALID;
  OrbDeviceGuard dev_guard(m->device);
  return bf_launch(m, d_q, nq, d_t, nt, ratio, th_dist, d_idx, d_d1, d_d2);
}

int orbm_bruteforce_host(orbm_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int th_dist,
                         int32_t* idx, int32_t* d1, int32_t* d2) {
  if (!m || nq < 0 || nt < 0 || (nq && (!q || !idx || !d1 || !d2)) || (nt && !t)) return ORBX_E_INVALID;
  if (nq == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  uint8_t* dq = m->scratch<uint8_t>(1, (size_t)nq * 32);
  uint8_t* dt = m->scratch<uint8_t>(2, (size_t)nt * 32);
  int32_t* dres = m->scratch<int32_t>(3, (size_t)nq * 3);
  if (!dq || !dt || !dres) return ORBX_E_CUDA;
  cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream);
  if (nt) cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream);
  const int rc = bf_launch(m, dq, nq, dt, nt, ratio, th_dist, dres, dres + nq, dres + 2 * nq);
  if (rc != ORBX_OK) return rc;
  cudaMemcpyAsync(idx, dres, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, m->stream);
  cudaMemcpyAsync(d1, dres + nq, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, m->stream);
  cudaMemcpyAsync(d2, dres + 2 * nq, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, m->stream);
  return m->check(cudaStreamSynchronize(m->stream), "bruteforce") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_bruteforce_batch_device(orbm_matcher* m, int n_pairs, int cap, const uint8_t* d_q, const int32_t* d_nq,
                                 size_t q_stride, const uint8_t* d_t, const int32_t* d_nt, size_t t_stride, float ratio,
                                 int th_dist, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2) {
  if (!m || n_pairs < 0 || cap < 1 || cap > 65535 || (n_pairs && (!d_q || !d_t || !d_nq || !d_nt || !d_idx || !d_d1 || !d_d2)) ||
      (q_stride & 15) || (t_stride & 15))
    return ORBX_E_INVALID;
  if (n_pairs == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  const int qblocks = (cap + BF_THREADS * BF_QPT - 1) / (BF_THREADS * BF_QPT);
  for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
    const int np = std::min(65535, n_pairs - p0);
    k_bruteforce_batch<<<dim3(qblocks, np), BF_THREADS, 0, m->stream>>>(
        d_q + (size_t)p0 * q_stride, d_nq + p0, q_stride, d_t + (size_t)p0 * t_stride, d_nt + p0, t_stride, cap, ratio,
        th_dist, d_idx + (size_t)p0 * cap, d_d1 + (size_t)p0 * cap, d_d2 + (size_t)p0 * cap);
    m->launches++;
  }
  return m->check(cudaGetLastError(), "bruteforce batch launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_bruteforce_indexed_device(orbm_matcher* m, int n_pairs, int cap, const uint8_t* d_base, const int64_t* d_q_off,
                                   const int64_t* d_t_off, const int64_t* d_nq_off, const int64_t* d_nt_off, float ratio,
                                   int th_dist, int32_t* d_idx, int32_t* d_d1, int32_t* d_d2) {
  if (!m || n_pairs < 0 || cap < 1 || cap > 65535 ||
      (n_pairs && (!d_base || !d_q_off || !d_t_off || !d_nq_off || !d_nt_off || !d_idx || !d_d1 || !d_d2)) ||
      (reinterpret_cast<uintptr_t>(d_base) & 15))
    return ORBX_E_INVALID;
  if (n_pairs == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  const int qblocks = (cap + BF_THREADS * BF_QPT - 1) / (BF_THREADS * BF_QPT);
  static_assert(sizeof(long long) == sizeof(int64_t), "offset tables are 64-bit");
  for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
    const int np = std::min(65535, n_pairs - p0);
    k_bruteforce_indexed<<<dim3(qblocks, np), BF_THREADS, 0, m->stream>>>(
        d_base, reinterpret_cast<const long long*>(d_q_off) + p0, reinterpret_cast<const long long*>(d_t_off) + p0,
        reinterpret_cast<const long long*>(d_nq_off) + p0, reinterpret_cast<const long long*>(d_nt_off) + p0, cap, ratio, th_dist,
        d_idx + (size_t)p0 * cap, d_d1 + (size_t)p0 * cap, d_d2 + (size_t)p0 * cap);
    m->launches++;
  }
  return m->check(cudaGetLastError(), "bruteforce indexed launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_debug_set_row_budget(orbm_matcher* m, int rows_per_query) {
  if (!m || rows_per_query < 1) return ORBX_E_INVALID;
  m->rows_per_query = rows_per_query;
  m->rows_hint = 0;
  return ORBX_OK;
}

int orbm_search_for_initialization_device(orbm_matcher* m, int n_pairs, int cap, const orbx_keypoint* d_k1,
                                          const uint8_t* d_d1, const int32_t* d_n1, const orbx_keypoint* d_k2,
                                          const uint8_t* d_d2, const int32_t* d_n2, orbm_bounds bounds2,
                                          float* d_prev_xy, int window, float nnratio, int check_ori,
                                          int32_t* d_matches12, int32_t* d_nmatches) {
  if (!m || n_pairs < 0 || cap < 1 || cap > 65535) return ORBX_E_INVALID;
  if (n_pairs == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  const size_t smem = sizeof(int) * 9 * (size_t)cap + sizeof(uint32_t) * INIT_SMEM_CAND + (((size_t)cap + 15) & ~(size_t)15);
  if (smem > 200 * 1024) { m->err = "cap too large for SearchForInitialization"; return ORBX_E_INVALID; }
  // candidate rows are cap x cap per pair: process pairs in groups that keep the scratch <= ~2 GiB
  const size_t per_pair = (size_t)cap * cap * sizeof(uint32_t);
  const int group = (int)std::max<size_t>(1, std::min<size_t>(n_pairs, ((size_t)2 << 30) / per_pair));
  int* gstart = m->scratch<int>(4, (size_t)group * (GRID_CELLS + 1));
  uint16_t* gitems = m->scratch<uint16_t>(5, (size_t)group * cap);
  uint32_t* cand = m->scratch<uint32_t>(6, (size_t)group * cap * cap);
  int* cand_cnt = m->scratch<int>(7, (size_t)group * cap);
  if (!gstart || !gitems || !cand || !cand_cnt) return ORBX_E_CUDA;
  if (smem > 48 * 1024 &&
      !m->check(cudaFuncSetAttribute(k_init_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in"))
    return ORBX_E_CUDA;
  for (int p0 = 0; p0 < n_pairs; p0 += group) {
    const int np = std::min(group, n_pairs - p0);
    const size_t o = (size_t)p0 * cap;
    k_build_grid<<<np, 256, 0, m->stream>>>(d_k2 + o, d_n2 + p0, 0, cap, bounds2, gstart, gitems);
    float* prev = d_prev_xy ? d_prev_xy + o * 2 : nullptr;
    k_init_candidates<<<dim3((cap + 7) / 8, np), 256, 0, m->stream>>>(cap, d_k1 + o, d_d1 + o * 32, d_n1 + p0, d_k2 + o,
                                                                        d_d2 + o * 32, bounds2, gstart, gitems, prev,
                                                                        (float)window, cand, cand_cnt);
    k_init_resolve<<<np, INIT_NT, smem, m->stream>>>(cap, d_k1 + o, d_n1 + p0, d_k2 + o, d_n2 + p0, prev, nnratio, check_ori,
                                                 cand, cand_cnt, d_matches12 + o, d_nmatches + p0);
    m->launches += 3;
  }
  return m->check(cudaGetLastError(), "search_for_initialization launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_search_for_initialization_host(orbm_matcher* m, int n_pairs, int cap, const orbx_keypoint* k1, const uint8_t* d1,
                                        const int32_t* n1, const orbx_keypoint* k2, const uint8_t* d2, const int32_t* n2,
                                        orbm_bounds bounds2, float* prev_xy, int window, float nnratio, int check_ori,
                                        int32_t* matches12, int32_t* nmatches) {
  if (!m || n_pairs < 0 || cap < 1) return ORBX_E_INVALID;
  if (n_pairs == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  const size_t nk = (size_t)n_pairs * cap;
  // A few pairs (the per-frame call of Tracking::MonocularInitialization is ONE): ten pageable copies would cost more
  // than the kernels, so the inputs go up in one pinned block and the outputs come back in one.
  const size_t in_small = nk * (2 * sizeof(orbx_keypoint) + 64 + 8) + 8 * (size_t)n_pairs;
  if (in_small <= (1u << 20)) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // [k1 k2 | d1 d2 | n1 n2 | prev || nmatches | matches12]: the upload ends after prev (read AND updated), the download
    // starts at prev
    const size_t o_k = 0, o_d = al(2 * nk * sizeof(orbx_keypoint)), o_n = o_d + al(2 * nk * 32),
                 o_prev = o_n + al(8 * (size_t)n_pairs), o_nm = o_prev + al(8 * nk), o_m12 = o_nm + al(4 * (size_t)n_pairs),
                 dev_bytes = o_m12 + al(4 * nk), in_bytes = o_nm, o_out = o_prev, out_bytes = dev_bytes - o_out;
    uint8_t* hin = m->pinned(0, in_bytes);
    uint8_t* hout = m->pinned(1, out_bytes);
    uint8_t* dev = m->scratch<uint8_t>(8, dev_bytes);
    if (!hin || !hout || !dev) return ORBX_E_CUDA;
    cudaStream_t st = m->stream;
    std::memcpy(hin + o_k, k1, nk * sizeof(orbx_keypoint));
    std::memcpy(hin + o_k + nk * sizeof(orbx_keypoint), k2, nk * sizeof(orbx_keypoint));
    std::memcpy(hin + o_d, d1, nk * 32);
    std::memcpy(hin + o_d + nk * 32, d2, nk * 32);
    std::memcpy(hin + o_n, n1, 4 * (size_t)n_pairs);
    std::memcpy(hin + o_n + 4 * (size_t)n_pairs, n2, 4 * (size_t)n_pairs);
    std::memcpy(hin + o_prev, prev_xy, 8 * nk);
    cudaMemcpyAsync(dev, hin, in_bytes, cudaMemcpyHostToDevice, st);
    orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(dev + o_k);
    uint8_t* dd = dev + o_d;
    int32_t* dn = reinterpret_cast<int32_t*>(dev + o_n);
    int32_t* dnm = reinterpret_cast<int32_t*>(dev + o_nm);
    int32_t* dm12 = reinterpret_cast<int32_t*>(dev + o_m12);
    float* dprev = reinterpret_cast<float*>(dev + o_prev);
    const int rc = orbm_search_for_initialization_device(m, n_pairs, cap, dk, dd, dn, dk + nk, dd + nk * 32, dn + n_pairs,
                                                         bounds2, dprev, window, nnratio, check_ori, dm12, dnm);
    if (rc != ORBX_OK) return rc;
    cudaMemcpyAsync(hout, dev + o_out, out_bytes, cudaMemcpyDeviceToHost, st);
    if (!m->check(cudaStreamSynchronize(st), "search_for_initialization")) return ORBX_E_CUDA;
    std::memcpy(nmatches, hout + (o_nm - o_out), 4 * (size_t)n_pairs);
    std::memcpy(matches12, hout + (o_m12 - o_out), 4 * nk);
    std::memcpy(prev_xy, hout + (o_prev - o_out), 8 * nk);
    return ORBX_OK;
  }
  orbx_keypoint* dk = m->scratch<orbx_keypoint>(8, 2 * nk);
  uint8_t* dd = m->scratch<uint8_t>(9, 2 * nk * 32);
  int32_t* dn = m->scratch<int32_t>(10, 3 * (size_t)n_pairs + nk);
  float* dprev = m->scratch<float>(11, nk * 2);
  if (!dk || !dd || !dn || !dprev) return ORBX_E_CUDA;
  cudaStream_t st = m->stream;
  cudaMemcpyAsync(dk, k1, nk * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dk + nk, k2, nk * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dd, d1, nk * 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dd + nk * 32, d2, nk * 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dn, n1, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dn + n_pairs, n2, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dprev, prev_xy, nk * 2 * sizeof(float), cudaMemcpyHostToDevice, st);
  int32_t* dnm = dn + 2 * n_pairs;
  int32_t* dm12 = dn + 3 * n_pairs;
  const int rc = orbm_search_for_initialization_device(m, n_pairs, cap, dk, dd, dn, dk + nk, dd + nk * 32, dn + n_pairs,
                                                       bounds2, dprev, window, nnratio, check_ori, dm12, dnm);
  if (rc != ORBX_OK) return rc;
  cudaMemcpyAsync(matches12, dm12, nk * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(nmatches, dnm, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(prev_xy, dprev, nk * 2 * sizeof(float), cudaMemcpyDeviceToHost, st);
  return m->check(cudaStreamSynchronize(st), "search_for_initialization") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_search_by_projection_points_host(orbm_matcher* m, const orbx_keypoint* k, const uint8_t* desc,
                                          const float* u_right, int n, orbm_bounds bounds, const float* scale_factors,
                                          int nlevels, const orbm_mappoint* mp, const uint8_t* mp_desc,
                                          const int32_t* mp_obs, int nmp, float th, float nnratio, int32_t* frame_mp,
                                          const int32_t* frame_mp_obs, int* nmatches) {
  if (!m || !k || !desc || !scale_factors || !frame_mp || !nmatches || n < 0 || n > 65535 || nmp < 0 || nlevels < 1 ||
      nlevels > ORBX_MAX_LEVELS || (nmp && (!mp || !mp_desc)))
    return ORBX_E_INVALID;
  *nmatches = 0;
  if (n == 0 || nmp == 0) return ORBX_OK;
  for (int i = 0; i < nmp; ++i)
    if (mp[i].track_in_view && !mp[i].bad && (mp[i].level < 0 || mp[i].level >= nlevels)) {
      m->err = "map point level out of range";
      return ORBX_E_INVALID;
    }
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  // one device block [uploaded inputs | working arrays], inputs packed in one pinned block: ONE H2D copy (run_projected)
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_desc = 0, o_k = up(o_desc + (size_t)n * 32), o_ur = up(o_k + sizeof(orbx_keypoint) * n),
               o_fmp = up(o_ur + 4 * (size_t)n), o_fobs = up(o_fmp + 4 * (size_t)n), o_sf = up(o_fobs + 4 * (size_t)n),
               o_md = up(o_sf + 4 * 64), o_mp = up(o_md + (size_t)nmp * 32), o_mobs = up(o_mp + sizeof(orbm_mappoint) * nmp),
               o_misc = up(o_mobs + 4 * (size_t)nmp), in_bytes = up(o_misc + 8 * sizeof(int));
  const size_t o_held = in_bytes, o_cnt = up(o_held + n), o_off = up(o_cnt + 4 * (size_t)nmp), total_bytes = up(o_off + 4 * (size_t)nmp);
  uint8_t* dev = m->scratch<uint8_t>(8, total_bytes);
  uint8_t* hin = m->pinned(0, in_bytes);
  int* gstart = m->scratch<int>(4, GRID_CELLS + 1);
  uint16_t* gitems = m->scratch<uint16_t>(5, n);
  if (!dev || !hin || !gstart || !gitems) return ORBX_E_CUDA;
  size_t rows_cap = std::max<size_t>(m->rows_hint, (size_t)nmp * m->rows_per_query);
  int total = 0, got = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    uint32_t* rows = m->scratch<uint32_t>(6, rows_cap);
    uint8_t* hout = m->pinned(1, 4 * (size_t)n + 8 * sizeof(int));
    if (!rows || !hout) return ORBX_E_CUDA;
    std::memcpy(hin + o_desc, desc, (size_t)n * 32);
    std::memcpy(hin + o_k, k, sizeof(orbx_keypoint) * n);
    float* h_ur = reinterpret_cast<float*>(hin + o_ur);
    if (u_right) std::memcpy(h_ur, u_right, 4 * (size_t)n); else std::fill(h_ur, h_ur + n, -1.f);
    std::memcpy(hin + o_fmp, frame_mp, 4 * (size_t)n);
    if (frame_mp_obs) std::memcpy(hin + o_fobs, frame_mp_obs, 4 * (size_t)n); else std::memset(hin + o_fobs, 0, 4 * (size_t)n);
    std::memcpy(hin + o_sf, scale_factors, sizeof(float) * nlevels);
    std::memcpy(hin + o_md, mp_desc, (size_t)nmp * 32);
    std::memcpy(hin + o_mp, mp, sizeof(orbm_mappoint) * nmp);
    int32_t* h_mobs = reinterpret_cast<int32_t*>(hin + o_mobs);
    if (mp_obs) std::memcpy(h_mobs, mp_obs, 4 * (size_t)nmp); else std::fill(h_mobs, h_mobs + nmp, 1);
    std::memset(hin + o_misc, 0, 8 * sizeof(int));  // [0] total candidate rows (scan), [2] nmatches
    if (!m->check(cudaMemcpyAsync(dev, hin, in_bytes, cudaMemcpyHostToDevice, st), "H2D search_by_projection")) return ORBX_E_CUDA;
    uint8_t* dd = dev + o_desc;
    orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(dev + o_k);
    float* dur = reinterpret_cast<float*>(dev + o_ur);
    int32_t* dfmp = reinterpret_cast<int32_t*>(dev + o_fmp);
    int32_t* dfobs = reinterpret_cast<int32_t*>(dev + o_fobs);
    float* dsf = reinterpret_cast<float*>(dev + o_sf);
    uint8_t* dmd = dev + o_md;
    orbm_mappoint* dmp = reinterpret_cast<orbm_mappoint*>(dev + o_mp);
    int32_t* dmobs = reinterpret_cast<int32_t*>(dev + o_mobs);
    int* misc = reinterpret_cast<int*>(dev + o_misc);
    uint8_t* dheld = dev + o_held;
    int* drow_cnt = reinterpret_cast<int*>(dev + o_cnt);
    int* drow_off = reinterpret_cast<int*>(dev + o_off);
    k_build_grid<<<1, 256, 0, st>>>(dk, nullptr, n, n, bounds, gstart, gitems);
    const int blocks = (nmp + 7) / 8;
    k_proj_candidates<<<blocks, 256, 0, st>>>(dk, dd, dur, bounds, gstart, gitems, dsf, dmp, dmd, nmp, th, 1, drow_cnt, nullptr,
                                              nullptr);
    k_scan_exclusive<<<1, 1024, 0, st>>>(drow_cnt, drow_off, nmp, misc);  // misc[0] = total rows needed
    k_proj_candidates<<<blocks, 256, 0, st>>>(dk, dd, dur, bounds, gstart, gitems, dsf, dmp, dmd, nmp, th, 0, drow_cnt, drow_off,
                                              rows, (long long)rows_cap);
    const int held_in_smem = n <= 40960;  // occupancy bytes on chip when they fit
    k_proj_resolve_cta<<<1, 1024, held_in_smem ? (size_t)((n + 15) & ~15) : 0, st>>>(
        drow_cnt, drow_off, rows, dmobs, nmp, n, nnratio, dfmp, dfobs, dheld, held_in_smem, misc + 2, misc, (long long)rows_cap);
    m->launches += 5;
    cudaMemcpyAsync(hout, dfmp, 4 * (size_t)n, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(hout + 4 * (size_t)n, misc, 8 * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (!m->check(cudaStreamSynchronize(st), "search_by_projection")) return ORBX_E_CUDA;
    const int* r = reinterpret_cast<const int*>(hout + 4 * (size_t)n);
    total = r[0];
    got = r[2];
    if ((size_t)total <= rows_cap) {
      std::memcpy(frame_mp, hout, 4 * (size_t)n);
      break;
    }
    if (attempt == 1) { m->err = "candidate rows overflowed twice"; return ORBX_E_CAPACITY; }
    rows_cap = (size_t)total + (size_t)total / 4;
  }
  m->rows_hint = std::max(m->rows_hint, (size_t)total + (size_t)total / 4);
  *nmatches = got;
  return m->check(cudaGetLastError(), "search_by_projection launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_search_by_bow_batch_host(orbm_matcher* m, orbm_bow_pair* pairs, int n_pairs, float nnratio, int check_ori,
                                  int max_dist) {
  if (!m || n_pairs < 0 || (n_pairs && !pairs)) return ORBX_E_INVALID;
  std::vector<BowQuery> queries;
  std::vector<int32_t> info((size_t)n_pairs * 8, 0), items2;
  size_t G1 = 0, G2 = 0;
  int total = 0, max_n2 = 0, max_rows = 0;
  for (int p = 0; p < n_pairs; ++p) {
    orbm_bow_pair& P = pairs[p];
    const orbm_featvec &fv1 = P.fv1, &fv2 = P.fv2;
    if (!P.matches12 || P.n1 < 0 || P.n2 < 0 || P.n1 > 65535 || P.n2 > 65535 || fv1.n_nodes < 0 || fv2.n_nodes < 0 ||
        (P.n1 && (!P.desc1 || !P.angle1)) || (P.n2 && (!P.desc2 || !P.angle2)) ||
        (fv1.n_nodes && (!fv1.node_id || !fv1.start || !fv1.items)) || (fv2.n_nodes && (!fv2.node_id || !fv2.start || !fv2.items)))
      return ORBX_E_INVALID;
    P.nmatches = 0;
    for (int i = 0; i < P.n1; ++i) P.matches12[i] = -1;
    if (P.matches21)
      for (int i = 0; i < P.n2; ++i) P.matches21[i] = -1;
    int32_t* I = &info[(size_t)p * 8];
    I[0] = (int32_t)queries.size();
    I[2] = total;
    I[4] = (int32_t)G1;
    I[5] = (int32_t)G2;
    I[6] = P.n2;
    // the walk of the two std::map's with lower_bound jumps (:350-359), emitting the queries
    const int items_base = (int)items2.size();
    const int n_items2 = fv2.n_nodes ? fv2.start[fv2.n_nodes] : 0;
    for (int i = 0; i < n_items2; ++i) {
      if (fv2.items[i] < 0 || fv2.items[i] >= P.n2) { m->err = "feature vector 2: index out of range"; return ORBX_E_INVALID; }
      items2.push_back((int32_t)G2 + fv2.items[i]);
    }
    int a = 0, b = 0;
    while (a < fv1.n_nodes && b < fv2.n_nodes) {
      const int32_t na = fv1.node_id[a], nb = fv2.node_id[b];
      if (na == nb) {
        const int t0 = fv2.start[b], cnt = fv2.start[b + 1] - t0;
        for (int q = fv1.start[a]; q < fv1.start[a + 1]; ++q) {
          const int idx1 = fv1.items[q];
          if (idx1 < 0 || idx1 >= P.n1) { m->err = "feature vector 1: index out of range"; return ORBX_E_INVALID; }
          if (P.valid1 && !P.valid1[idx1]) continue;
          if (cnt > 0) {
            queries.push_back({(int32_t)G1 + idx1, items_base + t0, cnt, total, (int32_t)G2, 0, 0, 0});
            total += cnt;
          }
        }
        ++a; ++b;
      } else if (na < nb) {
        a = (int)(std::lower_bound(fv1.node_id + a, fv1.node_id + fv1.n_nodes, nb) - fv1.node_id);
      } else {
        b = (int)(std::lower_bound(fv2.node_id + b, fv2.node_id + fv2.n_nodes, na) - fv2.node_id);
      }
    }
    I[1] = (int32_t)queries.size() - I[0];
    I[3] = total - I[2];
    max_n2 = std::max(max_n2, P.n2);
    max_rows = std::max(max_rows, I[3] <= BOW_SMEM_ROWS ? I[3] : 0);
    G1 += P.n1;
    G2 += P.n2;
  }
  const int nq = (int)queries.size();
  if (nq == 0) return ORBX_OK;
  // One device block [uploaded inputs | working arrays]; the inputs are packed straight into ONE pinned host block and go
  // up in ONE copy, the outputs come back through a second pinned block (pageable cudaMemcpyAsync calls cost ~10 us each
  // and a batch of 32 pairs is 4 MB of descriptors).
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_desc = 0, o_valid = al((G1 + G2) * 32), o_ang = o_valid + al(G2), o_items = o_ang + al(4 * (G1 + G2)),
               o_q = o_items + al(4 * items2.size()), o_info = o_q + al(sizeof(BowQuery) * (size_t)nq),
               in_bytes = o_info + al(4 * info.size());
  const size_t o_m = in_bytes, o_bin = o_m + al(4 * (G1 + G2)), o_nm = o_bin + al(4 * (size_t)nq),
               dev_bytes = o_nm + al(4 * (size_t)n_pairs);
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  uint8_t* hin = m->pinned(0, in_bytes);
  uint8_t* hout = m->pinned(1, 4 * (G1 + G2) + 4 * (size_t)n_pairs);
  uint8_t* dev = m->scratch<uint8_t>(8, dev_bytes);
  uint32_t* rows = m->scratch<uint32_t>(6, (size_t)total);
  if (!hin || !hout || !dev || !rows) return ORBX_E_CUDA;
  {
    uint8_t* hd = hin + o_desc;
    uint8_t* hvalid2 = hin + o_valid;
    float* hang = reinterpret_cast<float*>(hin + o_ang);
    size_t o1 = 0, o2 = 0;
    for (int p = 0; p < n_pairs; ++p) {
      const orbm_bow_pair& P = pairs[p];
      if (P.n1) { std::memcpy(hd + o1 * 32, P.desc1, (size_t)P.n1 * 32); std::memcpy(hang + o1, P.angle1, sizeof(float) * P.n1); }
      if (P.n2) {
        std::memcpy(hd + (G1 + o2) * 32, P.desc2, (size_t)P.n2 * 32);
        std::memcpy(hang + G1 + o2, P.angle2, sizeof(float) * P.n2);
        for (int i = 0; i < P.n2; ++i) hvalid2[o2 + i] = P.valid2 ? (P.valid2[i] ? 1 : 0) : 1;
      }
      o1 += P.n1;
      o2 += P.n2;
    }
    std::memcpy(hin + o_items, items2.data(), 4 * items2.size());
    std::memcpy(hin + o_q, queries.data(), sizeof(BowQuery) * (size_t)nq);
    std::memcpy(hin + o_info, info.data(), 4 * info.size());
  }
  uint8_t *dd1 = dev + o_desc, *dd2 = dd1 + G1 * 32, *dvalid2 = dev + o_valid;
  float *da1 = reinterpret_cast<float*>(dev + o_ang), *da2 = da1 + G1;
  int32_t* ditems2 = reinterpret_cast<int32_t*>(dev + o_items);
  BowQuery* dq = reinterpret_cast<BowQuery*>(dev + o_q);
  int32_t* dinfo = reinterpret_cast<int32_t*>(dev + o_info);
  int32_t* dm12 = reinterpret_cast<int32_t*>(dev + o_m);
  int32_t* dm21 = dm12 + G1;
  int32_t* dbin = reinterpret_cast<int32_t*>(dev + o_bin);
  int* dnm = reinterpret_cast<int*>(dev + o_nm);
  cudaMemcpyAsync(dev, hin, in_bytes, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(dm12, 0xFF, sizeof(int32_t) * (G1 + G2), st);  // matches12 and matches21 = -1
  cudaMemsetAsync(dnm, 0, sizeof(int) * n_pairs, st);
  int max_nodes = 0;  // vocabulary nodes with queries, per pair (queries of a node are consecutive and share t0)
  for (int p = 0; p < n_pairs; ++p) {
    const int qa = info[8 * p], qn = info[8 * p + 1];
    int nn = 0;
    for (int j = qa; j < qa + qn; ++j) nn += j == qa || queries[j].t0 != queries[j - 1].t0;
    max_nodes = std::max(max_nodes, nn);
  }
  const size_t smem = sizeof(uint32_t) * ((size_t)(max_nodes + 1) + (size_t)(max_n2 / 32 + 1) + max_rows);
  if (smem > 48 * 1024 &&
      !m->check(cudaFuncSetAttribute(k_bow_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in"))
    return ORBX_E_CUDA;
  k_bow_candidates<<<(nq + 7) / 8, 256, 0, st>>>(dq, nq, dd1, dd2, dvalid2, ditems2, rows);
  k_bow_resolve<<<n_pairs, BOW_NT, smem, st>>>(dq, dinfo, rows, da1, da2, nnratio, check_ori, max_dist, max_nodes, dm12, dm21, dbin, dnm);
  m->launches += 2;
  int32_t* hm = reinterpret_cast<int32_t*>(hout);
  int32_t* hnm = hm + (G1 + G2);
  cudaMemcpyAsync(hm, dm12, sizeof(int32_t) * (G1 + G2), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(hnm, dnm, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost, st);
  if (!m->check(cudaStreamSynchronize(st), "search_by_bow")) return ORBX_E_CUDA;
  {
    size_t o1 = 0, o2 = 0;
    for (int p = 0; p < n_pairs; ++p) {
      orbm_bow_pair& P = pairs[p];
      std::copy(hm + o1, hm + o1 + P.n1, P.matches12);
      if (P.matches21) std::copy(hm + G1 + o2, hm + G1 + o2 + P.n2, P.matches21);
      P.nmatches = hnm[p];
      o1 += P.n1;
      o2 += P.n2;
    }
  }
  return m->check(cudaGetLastError(), "search_by_bow launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_search_by_bow_host(orbm_matcher* m, const uint8_t* desc1, const float* angle1, const int32_t* valid1, int n1,
                            orbm_featvec fv1, const uint8_t* desc2, const float* angle2, const int32_t* valid2, int n2,
                            orbm_featvec fv2, float nnratio, int check_ori, int max_dist, int32_t* matches12,
                            int32_t* matches21, int* nmatches) {
  if (!nmatches) return ORBX_E_INVALID;
  orbm_bow_pair P = {desc1, angle1, valid1, n1, fv1, desc2, angle2, valid2, n2, fv2, matches12, matches21, 0};
  const int rc = orbm_search_by_bow_batch_host(m, &P, 1, nnratio, check_ori, max_dist);
  *nmatches = P.nmatches;
  return rc;
}

int orbm_search_for_triangulation_batch_host(orbm_matcher* m, orbm_tri_pair* pairs, int n_pairs, int nlevels, int only_stereo,
                                             const int32_t* cam_enabled, int check_ori) {
  if (!m || n_pairs < 0 || (n_pairs && !pairs) || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !cam_enabled) return ORBX_E_INVALID;
  std::vector<BowQuery> queries;
  std::vector<int32_t> pair_q(n_pairs + 1, 0), items2;
  std::vector<float> hconst((size_t)n_pairs * 56, 0.f);
  size_t G1 = 0, G2 = 0;
  for (int p = 0; p < n_pairs; ++p) {
    orbm_tri_pair& P = pairs[p];
    const orbm_featvec &fv1 = P.fv1, &fv2 = P.fv2;
    if (!P.matches12 || P.n1 < 0 || P.n2 < 0 || P.n1 > 65535 || P.n2 > 65535 || fv1.n_nodes < 0 || fv2.n_nodes < 0 || !P.F12s ||
        !P.epipoles || !P.scale_factors2 || !P.level_sigma2_2 ||
        (P.n1 && (!P.k1 || !P.desc1 || !P.has_mp1 || !P.cam1 || !P.uright1)) ||
        (P.n2 && (!P.k2 || !P.desc2 || !P.has_mp2 || !P.cam2 || !P.uright2)) ||
        (fv1.n_nodes && (!fv1.node_id || !fv1.start || !fv1.items)) || (fv2.n_nodes && (!fv2.node_id || !fv2.start || !fv2.items)))
      return ORBX_E_INVALID;
    P.nmatches = 0;
    for (int i = 0; i < P.n1; ++i) P.matches12[i] = -1;
    for (int i = 0; i < P.n1; ++i)
      if (P.cam1[i] < 0 || P.cam1[i] > 1) { m->err = "key frame 1: camera index out of range"; return ORBX_E_INVALID; }
    for (int i = 0; i < P.n2; ++i)
      if (P.k2[i].octave < 0 || P.k2[i].octave >= nlevels) { m->err = "key frame 2: octave out of range"; return ORBX_E_INVALID; }
    float* hc = &hconst[(size_t)p * 56];
    std::copy(P.F12s, P.F12s + 18, hc);
    std::copy(P.epipoles, P.epipoles + 4, hc + 18);
    std::copy(P.scale_factors2, P.scale_factors2 + nlevels, hc + 22);
    std::copy(P.level_sigma2_2, P.level_sigma2_2 + nlevels, hc + 38);
    pair_q[p] = (int32_t)queries.size();
    const int items_base = (int)items2.size();
    const int n_items2 = fv2.n_nodes ? fv2.start[fv2.n_nodes] : 0;
    for (int i = 0; i < n_items2; ++i) {
      if (fv2.items[i] < 0 || fv2.items[i] >= P.n2) { m->err = "feature vector 2: index out of range"; return ORBX_E_INVALID; }
      items2.push_back((int32_t)G2 + fv2.items[i]);
    }
    // feature-vector walk (:1456-1700): one query per key-frame-1 feature that passes the side-1 tests (:1468-1480)
    int a = 0, b = 0;
    while (a < fv1.n_nodes && b < fv2.n_nodes) {
      const int32_t na = fv1.node_id[a], nb = fv2.node_id[b];
      if (na == nb) {
        const int t0 = fv2.start[b], cnt = fv2.start[b + 1] - t0;
        if (cnt > 65535) { m->err = "feature vector 2: node with more than 65535 features"; return ORBX_E_INVALID; }
        for (int q = fv1.start[a]; q < fv1.start[a + 1]; ++q) {
          const int idx1 = fv1.items[q];
          if (idx1 < 0 || idx1 >= P.n1) { m->err = "feature vector 1: index out of range"; return ORBX_E_INVALID; }
          if (P.has_mp1[idx1] || !cam_enabled[P.cam1[idx1]]) continue;
          if (only_stereo && !(P.uright1[idx1] >= 0)) continue;
          if (cnt > 0) queries.push_back({(int32_t)G1 + idx1, items_base + t0, cnt, 0, (int32_t)G2, p, 0, 0});
        }
        ++a; ++b;
      } else if (na < nb) {
        a = (int)(std::lower_bound(fv1.node_id + a, fv1.node_id + fv1.n_nodes, nb) - fv1.node_id);
      } else {
        b = (int)(std::lower_bound(fv2.node_id + b, fv2.node_id + fv2.n_nodes, na) - fv2.node_id);
      }
    }
    G1 += P.n1;
    G2 += P.n2;
  }
  pair_q[n_pairs] = (int32_t)queries.size();
  const int nq = (int)queries.size();
  if (nq == 0) return ORBX_OK;
  // key-frame-2 features follow all key-frame-1 features in the packed arrays
  for (int32_t& v : items2) v += (int32_t)G1;
  for (BowQuery& bq : queries) bq.o2 += (int32_t)G1;
  // One device block [uploaded inputs | working arrays]; both sides are packed into batch-global arrays (descriptors,
  // keypoints, has_mp, cam, uright) straight inside ONE pinned host block that goes up in ONE copy; the results come back
  // through a second pinned block.
  const size_t G = G1 + G2;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_d = 0, o_k = al(G * 32), o_mp = o_k + al(sizeof(orbx_keypoint) * G), o_cam = o_mp + al(4 * G),
               o_ur = o_cam + al(4 * G), o_const = o_ur + al(4 * G), o_items = o_const + al(4 * hconst.size()),
               o_q = o_items + al(4 * items2.size()), o_pq = o_q + al(sizeof(BowQuery) * (size_t)nq),
               in_bytes = o_pq + al(4 * (size_t)(n_pairs + 1));
  const size_t o_m12 = in_bytes, o_bin = o_m12 + al(4 * G1), o_hist = o_bin + al(4 * (size_t)nq),
               dev_bytes = o_hist + al(4 * ((size_t)n_pairs * (HISTO_LENGTH + 1)));
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  uint8_t* hin = m->pinned(0, in_bytes);
  uint8_t* hout = m->pinned(1, 4 * G1 + 4 * (size_t)n_pairs);
  uint8_t* dev = m->scratch<uint8_t>(8, dev_bytes);
  if (!hin || !hout || !dev) return ORBX_E_CUDA;
  {
    uint8_t* hd = hin + o_d;
    orbx_keypoint* hk = reinterpret_cast<orbx_keypoint*>(hin + o_k);
    int32_t* hmp = reinterpret_cast<int32_t*>(hin + o_mp);
    int32_t* hcam = reinterpret_cast<int32_t*>(hin + o_cam);
    float* hur = reinterpret_cast<float*>(hin + o_ur);
    size_t o1 = 0, o2 = G1;
    for (int p = 0; p < n_pairs; ++p) {
      const orbm_tri_pair& P = pairs[p];
      if (P.n1) {
        std::memcpy(hd + o1 * 32, P.desc1, (size_t)P.n1 * 32);
        std::copy(P.k1, P.k1 + P.n1, hk + o1);
        std::copy(P.has_mp1, P.has_mp1 + P.n1, hmp + o1);
        std::copy(P.cam1, P.cam1 + P.n1, hcam + o1);
        std::copy(P.uright1, P.uright1 + P.n1, hur + o1);
      }
      if (P.n2) {
        std::memcpy(hd + o2 * 32, P.desc2, (size_t)P.n2 * 32);
        std::copy(P.k2, P.k2 + P.n2, hk + o2);
        std::copy(P.has_mp2, P.has_mp2 + P.n2, hmp + o2);
        std::copy(P.cam2, P.cam2 + P.n2, hcam + o2);
        std::copy(P.uright2, P.uright2 + P.n2, hur + o2);
      }
      o1 += P.n1;
      o2 += P.n2;
    }
    std::memcpy(hin + o_const, hconst.data(), 4 * hconst.size());
    std::memcpy(hin + o_items, items2.data(), 4 * items2.size());
    std::memcpy(hin + o_q, queries.data(), sizeof(BowQuery) * (size_t)nq);
    std::memcpy(hin + o_pq, pair_q.data(), 4 * (size_t)(n_pairs + 1));
  }
  uint8_t* dd = dev + o_d;  // descriptors first: uint4 loads need 16-byte alignment
  orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(dev + o_k);
  int32_t* dmp = reinterpret_cast<int32_t*>(dev + o_mp);
  int32_t* dcam = reinterpret_cast<int32_t*>(dev + o_cam);
  float* dur = reinterpret_cast<float*>(dev + o_ur);
  float* dconst = reinterpret_cast<float*>(dev + o_const);
  int32_t* ditems2 = reinterpret_cast<int32_t*>(dev + o_items);
  BowQuery* dq = reinterpret_cast<BowQuery*>(dev + o_q);
  int32_t* dpq = reinterpret_cast<int32_t*>(dev + o_pq);
  int32_t* dm12 = reinterpret_cast<int32_t*>(dev + o_m12);
  int32_t* dbin = reinterpret_cast<int32_t*>(dev + o_bin);
  int* dhist = reinterpret_cast<int*>(dev + o_hist);
  int* dnm = dhist + (size_t)n_pairs * HISTO_LENGTH;
  cudaMemcpyAsync(dev, hin, in_bytes, cudaMemcpyHostToDevice, st);
  const TriSide s1 = {dk, dd, dmp, dcam, dur};  // both sides share the arrays (batch-global indices)
  cudaMemsetAsync(dm12, 0xFF, sizeof(int32_t) * G1, st);
  cudaMemsetAsync(dhist, 0, sizeof(int) * ((size_t)n_pairs * (HISTO_LENGTH + 1)), st);
  k_tri_match<<<(nq + 7) / 8, 256, 0, st>>>(dq, nq, s1, s1, ditems2, dconst, only_stereo, check_ori, dm12, dbin, dhist);
  k_tri_finish<<<n_pairs, 256, 0, st>>>(dq, dpq, check_ori, dbin, dhist, dm12, dnm);
  m->launches += 2;
  int32_t* hm = reinterpret_cast<int32_t*>(hout);
  int32_t* hnm = hm + G1;
  cudaMemcpyAsync(hm, dm12, sizeof(int32_t) * G1, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(hnm, dnm, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost, st);
  if (!m->check(cudaStreamSynchronize(st), "search_for_triangulation")) return ORBX_E_CUDA;
  {
    size_t o1 = 0;
    for (int p = 0; p < n_pairs; ++p) {
      std::copy(hm + o1, hm + o1 + pairs[p].n1, pairs[p].matches12);
      pairs[p].nmatches = hnm[p];
      o1 += pairs[p].n1;
    }
  }
  return m->check(cudaGetLastError(), "search_for_triangulation launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_search_for_triangulation_host(orbm_matcher* m, const orbx_keypoint* k1, const uint8_t* desc1, const int32_t* has_mp1,
                                       const int32_t* cam1, const float* uright1, int n1, orbm_featvec fv1,
                                       const orbx_keypoint* k2, const uint8_t* desc2, const int32_t* has_mp2,
                                       const int32_t* cam2, const float* uright2, int n2, orbm_featvec fv2, const float* F12s,
                                       const float* epipoles, const float* scale_factors2, const float* level_sigma2_2,
                                       int nlevels, int only_stereo, const int32_t* cam_enabled, int check_ori,
                                       int32_t* matches12, int* nmatches) {
  if (!nmatches) return ORBX_E_INVALID;
  orbm_tri_pair P = {k1, desc1, has_mp1, cam1, uright1, n1, fv1, k2, desc2, has_mp2, cam2, uright2, n2, fv2,
                     F12s, epipoles, scale_factors2, level_sigma2_2, matches12, 0};
  const int rc = orbm_search_for_triangulation_batch_host(m, &P, 1, nlevels, only_stereo, cam_enabled, check_ori);
  *nmatches = P.nmatches;
  return rc;
}

int orbm_compute_distinctive_descriptors_host(orbm_matcher* m, const uint8_t* desc, const int32_t* offsets, int n_points,
                                              int32_t* best_idx) {
  if (!m || n_points < 0 || (n_points && (!offsets || !best_idx))) return ORBX_E_INVALID;
  if (n_points == 0) return ORBX_OK;
  const int total = offsets[n_points];
  if (offsets[0] != 0 || total < 0 || (total && !desc)) return ORBX_E_INVALID;
  int n_max = 0;
  size_t big_total = 0;
  std::vector<int> big_off(n_points, 0);
  for (int p = 0; p < n_points; ++p) {
    const int N = offsets[p + 1] - offsets[p];
    if (N < 0 || N > 65535) { m->err = "descriptor set size out of range"; return ORBX_E_INVALID; }
    if (N > DD_SMEM_N) { big_off[p] = (int)big_total; big_total += (size_t)N * N; }
    if (big_total > 0x7FFFFFFFu) { m->err = "descriptor sets too large"; return ORBX_E_CAPACITY; }
    n_max = std::max(n_max, std::min(N, DD_SMEM_N));
  }
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  uint8_t* dd = m->scratch<uint8_t>(8, (size_t)std::max(total, 1) * 32);
  int32_t* dints = m->scratch<int32_t>(4, (size_t)3 * n_points + 2);
  uint16_t* dbig = m->scratch<uint16_t>(6, big_total);
  if (!dd || !dints || !dbig) return ORBX_E_CUDA;
  int32_t* doff = dints;
  int32_t* dbigoff = doff + n_points + 1;
  int32_t* dbest = dbigoff + n_points;
  cudaMemcpyAsync(dd, desc, (size_t)total * 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(doff, offsets, sizeof(int32_t) * (n_points + 1), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dbigoff, big_off.data(), sizeof(int32_t) * n_points, cudaMemcpyHostToDevice, st);
  const size_t smem = (size_t)n_max * n_max * sizeof(uint16_t);
  if (smem > 40 * 1024 &&
      !m->check(cudaFuncSetAttribute(k_distinctive, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     DD_SMEM_N * DD_SMEM_N * (int)sizeof(uint16_t)), "smem opt-in"))
    return ORBX_E_CUDA;
  k_distinctive<<<n_points, 128, smem, st>>>(dd, doff, dbig, dbigoff, dbest);
  m->launches += 1;
  cudaMemcpyAsync(best_idx, dbest, sizeof(int32_t) * n_points, cudaMemcpyDeviceToHost, st);
  if (!m->check(cudaStreamSynchronize(st), "compute_distinctive_descriptors")) return ORBX_E_CUDA;
  return m->check(cudaGetLastError(), "compute_distinctive_descriptors launch") ? ORBX_OK : ORBX_E_CUDA;
}

// ---- Frame glue -------------------------------------------------------------------------------
}  // extern "C"
#pragma GCC visibility pop
namespace {
UndistortParams undistort_params(float fx, float fy, float cx, float cy, const float* dist5) {
  UndistortParams P;
  P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
  P.ifx = 1. / P.fx; P.ify = 1. / P.fy;
  P.k0 = dist5[0]; P.k1 = dist5[1]; P.k2 = dist5[2]; P.k3 = dist5[3]; P.k4 = dist5[4];
  return P;
}
}  // namespace
extern "C" {
#pragma GCC visibility push(default)

int orbm_undistort_keypoints_device(orbm_matcher* m, int n_frames, int cap, const orbx_keypoint* d_kps,
                                    const int32_t* d_counts, float fx, float fy, float cx, float cy, const float* dist5,
                                    orbx_keypoint* d_kps_un) {
  if (!m || n_frames < 0 || cap < 1 || !d_kps || !d_counts || !dist5 || !d_kps_un || fx == 0.f || fy == 0.f)
    return ORBX_E_INVALID;
  if (n_frames == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  k_undistort<<<dim3((cap + 255) / 256, n_frames), 256, 0, m->stream>>>(d_kps, d_counts, 0, cap,
                                                                        undistort_params(fx, fy, cx, cy, dist5),
                                                                        dist5[0] == 0.0f, d_kps_un);  // :675-679
  m->launches += 1;
  return m->check(cudaGetLastError(), "undistort launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_undistort_keypoints_host(orbm_matcher* m, const orbx_keypoint* k, int n, float fx, float fy, float cx, float cy,
                                  const float* dist5, orbx_keypoint* k_un) {
  if (!m || n < 0 || (n && (!k || !k_un)) || !dist5 || fx == 0.f || fy == 0.f) return ORBX_E_INVALID;
  if (n == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  orbx_keypoint* d = m->scratch<orbx_keypoint>(8, 2 * (size_t)n);
  if (!d) return ORBX_E_CUDA;
  cudaMemcpyAsync(d, k, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice, st);
  k_undistort<<<dim3((n + 255) / 256, 1), 256, 0, st>>>(d, nullptr, n, n, undistort_params(fx, fy, cx, cy, dist5),
                                                        dist5[0] == 0.0f, d + n);
  m->launches += 1;
  cudaMemcpyAsync(k_un, d + n, sizeof(orbx_keypoint) * n, cudaMemcpyDeviceToHost, st);
  if (!m->check(cudaStreamSynchronize(st), "undistort")) return ORBX_E_CUDA;
  return m->check(cudaGetLastError(), "undistort launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_compute_image_bounds_host(orbm_matcher* m, int cols, int rows, float fx, float fy, float cx, float cy,
                                   const float* dist5, orbm_bounds* out) {
  if (!m || !dist5 || !out || cols < 1 || rows < 1) return ORBX_E_INVALID;
  if (dist5[0] == 0.0f) {  // :772-778
    out->min_x = 0.0f; out->max_x = (float)cols; out->min_y = 0.0f; out->max_y = (float)rows;
    return ORBX_OK;
  }
  orbx_keypoint c[4] = {}, u[4];
  c[1].x = (float)cols; c[2].y = (float)rows; c[3].x = (float)cols; c[3].y = (float)rows;  // :749-757
  const int rc = orbm_undistort_keypoints_host(m, c, 4, fx, fy, cx, cy, dist5, u);
  if (rc != ORBX_OK) return rc;
  out->min_x = std::min(u[0].x, u[2].x);  // :766-769
  out->max_x = std::max(u[1].x, u[3].x);
  out->min_y = std::min(u[0].y, u[1].y);
  out->max_y = std::max(u[2].y, u[3].y);
  return ORBX_OK;
}

int orbm_compute_stereo_from_rgbd_device(orbm_matcher* m, int n_frames, int cap, const orbx_keypoint* d_kps,
                                         const orbx_keypoint* d_kps_un, const int32_t* d_counts, const float* d_depth,
                                         int cols, int rows, size_t row_stride_floats, size_t frame_stride_floats, float mbf,
                                         float* d_uright, float* d_depth_out) {
  if (!m || n_frames < 0 || cap < 1 || !d_kps || !d_kps_un || !d_counts || !d_depth || !d_uright || !d_depth_out ||
      cols < 1 || rows < 1 || row_stride_floats < (size_t)cols ||
      (n_frames > 1 && frame_stride_floats < (size_t)rows * row_stride_floats))
    return ORBX_E_INVALID;
  if (n_frames == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  k_stereo_rgbd<<<dim3((cap + 255) / 256, n_frames), 256, 0, m->stream>>>(d_kps, d_kps_un, d_counts, cap, d_depth, cols, rows,
                                                                          row_stride_floats, frame_stride_floats, mbf,
                                                                          d_uright, d_depth_out);
  m->launches += 1;
  return m->check(cudaGetLastError(), "stereo-from-rgbd launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_compute_stereo_matches_device(orbm_matcher* m, const orbx_pyramid_view* left, const orbx_pyramid_view* right,
                                       int n_frames, int cap_l, const orbx_keypoint* d_kl, const uint8_t* d_dl,
                                       const int32_t* d_nl, int cap_r, const orbx_keypoint* d_kr, const uint8_t* d_dr,
                                       const int32_t* d_nr, float mbf, float mb, float* d_uright, float* d_depth) {
  if (!m || !left || !right || n_frames < 0 || cap_l < 1 || cap_r < 1 || cap_r > STEREO_MAX_R || !d_kl || !d_dl || !d_nl ||
      !d_kr || !d_dr || !d_nr || !d_uright || !d_depth || !(mb > 0.0f) || left->nlevels != right->nlevels ||
      left->nlevels < 1 || n_frames > left->n_frames || n_frames > right->n_frames)
    return ORBX_E_INVALID;
  if (n_frames == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  int32_t* sad = m->scratch<int32_t>(8, (size_t)n_frames * cap_l);
  if (!sad) return ORBX_E_CUDA;
  // the keypoints and pyramids come from the extractors' streams: order this stream after them
  for (const orbx_pyramid_view* v : {left, right}) {
    if ((cudaStream_t)v->stream == m->stream) continue;
    cudaEvent_t ev;
    if (!m->check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate") ||
        !m->check(cudaEventRecord(ev, (cudaStream_t)v->stream), "cudaEventRecord") ||
        !m->check(cudaStreamWaitEvent(m->stream, ev, 0), "cudaStreamWaitEvent"))
      return ORBX_E_CUDA;
    cudaEventDestroy(ev);  // released once the wait has been satisfied
  }
  k_stereo_match<<<dim3((cap_l + STEREO_LPB - 1) / STEREO_LPB, n_frames), 256, 0, m->stream>>>(
      *left, *right, cap_l, d_kl, d_dl, d_nl, cap_r, d_kr, d_dr, d_nr, mbf, mb, d_uright, d_depth, sad);
  k_stereo_median<<<n_frames, 256, 0, m->stream>>>(cap_l, d_nl, sad, d_uright, d_depth);
  m->launches += 2;
  return m->check(cudaGetLastError(), "stereo-matches launch") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_assign_features_to_grid_device(orbm_matcher* m, int n_frames, int cap, const orbx_keypoint* d_kps_un,
                                        const int32_t* d_counts, orbm_bounds bounds, int32_t* d_cell_start,
                                        uint16_t* d_items) {
  if (!m || n_frames < 0 || cap < 1 || cap > 65535 || !d_kps_un || !d_counts || !d_cell_start || !d_items ||
      !(bounds.max_x > bounds.min_x) || !(bounds.max_y > bounds.min_y))
    return ORBX_E_INVALID;
  if (n_frames == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  k_build_grid<<<n_frames, 256, 0, m->stream>>>(d_kps_un, d_counts, 0, cap, bounds, d_cell_start, d_items);
  m->launches += 1;
  return m->check(cudaGetLastError(), "grid launch") ? ORBX_OK : ORBX_E_CUDA;
}

// ---- pose-based SearchByProjection overloads ------------------------------------------------
}  // extern "C" (helpers below are internal)
#pragma GCC visibility pop
namespace {

// cv::Mat float algebra as OpenCV evaluates it for these tiny matrices (probed against cv2.gemm):
// a product row is accumulated in float32 left to right from 0, then the addend is added.
inline void mat3_mul_vec_add(const float* R, int rs, const float* x, const float* t, float alpha, float* out) {
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

// Ordered best-only resolve for query lists without multi-row groups (every overload except the
// two-camera Sim3 one), in the points overload's scheme: 32 warps, a warp per query of the 32-query
// step.  A query's answer is the least key among its candidates that hold no point, and occupancy
// only grows during the walk: every pending warp takes that least key under the occupancy as it
// stands, warp 0 commits the longest prefix of the step in which no two queries want the same
// keypoint, the rest scan again.  The lowest pending query never conflicts, so commit order,
// overwrites of keypoints that stay free (Observations()==0, :3565-3567) and the accepted list are
// those of the sequential loop.  (Tracking matcher, 1500 points: sequential warp 1.5 ms per call,
// single-warp speculation on a precomputed static best 0.49 ms, this form 0.34 ms.)
// Segments: queries interact only through the occupancy of the keypoints they can reach, and a query reaches the keypoints
// of ONE camera (its grid), so the queries of different cameras never see each other's effects: the host orders the queries
// by camera (stable) and every camera's segment is walked by its own CTA, in the reference's order within the camera.
// Only the rotation histogram is shared by the whole call: with more than one segment the CTAs add their bins to a global
// histogram and k_query_finish applies the three-maxima filter.
struct QuerySegs { int n; int start[9]; };
__global__ void __launch_bounds__(1024) k_query_resolve_cta(const ProjQuery* __restrict__ q, const int* __restrict__ row_cnt,
                                                            const int* __restrict__ row_off, const uint32_t* __restrict__ rows,
                                                            QuerySegs segs, int n, const orbx_keypoint* __restrict__ k, int th_dist,
                                                            int check_ori, int any_point_blocks, int32_t* __restrict__ frame_mp,
                                                            const int32_t* __restrict__ frame_mp_obs,
                                                            uint8_t* __restrict__ held_global, int held_in_smem,
                                                            int32_t* __restrict__ acc_idx, int32_t* __restrict__ acc_bin,
                                                            int* __restrict__ nmatches_out, int* __restrict__ ghist,
                                                            int* __restrict__ seg_nacc,
                                                            const int* __restrict__ rows_needed = nullptr, long long rows_cap = 0) {

  extern __shared__ __align__(16) uint8_t s_held[];
  __shared__ int s_hist[HISTO_LENGTH];
  if (rows_needed && *rows_needed > rows_cap) return;  // the row buffer overflowed: the host repeats the call (run_projected)
  // per-round exchange, double-buffered by round parity: two barriers per round instead of three
  __shared__ int s_bidx[2][32];
  __shared__ unsigned s_done[2], s_left[2];
  int rp = 0;  // round parity
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  uint8_t* held = held_in_smem ? s_held : held_global;
  for (int i = tid; i < n; i += 1024)
    held[i] = frame_mp[i] >= 0 && (any_point_blocks || (frame_mp_obs && frame_mp_obs[i] > 0)) ? 1 : 0;
  if (tid < HISTO_LENGTH) s_hist[tid] = 0;
  __syncthreads();
  const int qa = segs.start[blockIdx.x], nq = segs.start[blockIdx.x + 1];  // this CTA's queries [qa, nq)
  acc_idx += qa;  // the accepted matches of a segment are listed from its first query on
  acc_bin += qa;
  int nacc = 0;  // warp 0's count
  // Nothing on the walk's critical path waits for HBM (the scheme of k_proj_resolve_cta): a warp keeps the first
  // QR_SLOTS * 32 candidates of its query in registers (longer rows read the rest in place), the entries of the next
  // step's query are fetched while this step resolves, and the counts / offsets one step further ahead.
  constexpr int QR_SLOTS = 2;
  constexpr uint32_t QR_NONE = 0xFFFFFFFFu;  // entries are dist << 16 | index with dist <= 256
  uint32_t ent[QR_SLOTS], ent_n[QR_SLOTS];
  int cnt = qa + w < nq ? row_cnt[qa + w] : 0, off = qa + w < nq ? row_off[qa + w] : 0;
  int cnt_n = qa + 32 + w < nq ? row_cnt[qa + 32 + w] : 0, off_n = qa + 32 + w < nq ? row_off[qa + 32 + w] : 0;
#pragma unroll
  for (int kk = 0; kk < QR_SLOTS; ++kk) ent[kk] = kk * 32 + lane < cnt ? rows[off + kk * 32 + lane] : QR_NONE;
  for (int b0 = qa; b0 < nq; b0 += 32) {
#pragma unroll
    for (int kk = 0; kk < QR_SLOTS; ++kk) ent_n[kk] = kk * 32 + lane < cnt_n ? rows[off_n + kk * 32 + lane] : QR_NONE;
    const int j2 = b0 + 64 + w;
    const int cnt_nn = j2 < nq ? row_cnt[j2] : 0, off_nn = j2 < nq ? row_off[j2] : 0;
    const uint32_t* row = rows + off;
    bool pend = cnt > 0;
    // warp 0 keeps the per-query fields of the step, lane = query
    const int lj = b0 + lane;
    int src = 0;
    float angle = 0.0f;
    uint8_t hval = 1;
    if (w == 0 && lj < nq) {
      src = q[lj].src;
      angle = q[lj].angle;
      hval = any_point_blocks ? 1 : (q[lj].obs ? 1 : 0);
    }
    for (;;) {
      int bidx = -1;
      if (pend) {
        uint32_t best = 0xFFFFFFFFu;
        int my_idx = -1;
#pragma unroll
        for (int kk = 0; kk < QR_SLOTS; ++kk) {
          const uint32_t e = ent[kk];
          if (e == QR_NONE || held[e & 0xFFFFu]) continue;
          const uint32_t key = (e >> 16) << 16 | (uint32_t)(kk * 32 + lane);  // dist, then traversal position (strict <)
          if (key < best) { best = key; my_idx = (int)(e & 0xFFFFu); }
        }
        for (int c = QR_SLOTS * 32 + lane; c < cnt; c += 32) {
          const uint32_t e = row[c];
          if (held[e & 0xFFFFu]) continue;
          const uint32_t key = (e >> 16) << 16 | (uint32_t)c;
          if (key < best) { best = key; my_idx = (int)(e & 0xFFFFu); }
        }
        const uint32_t mine = best;
        best = __reduce_min_sync(0xffffffffu, best);
        // nothing free, or the least distance above the gate (it can only grow): the query is finished
        if (best == 0xFFFFFFFFu || (int)(best >> 16) > th_dist) {
          pend = false;
        } else {
          const unsigned owner = __ballot_sync(0xffffffffu, mine == best);
          bidx = __shfl_sync(0xffffffffu, my_idx, __ffs(owner) - 1);
        }
      }
      if (lane == 0) s_bidx[rp][w] = bidx;
      __syncthreads();
      if (w == 0) {
        const int b = s_bidx[rp][lane];
        const bool p = b >= 0;
        const unsigned peers = __match_any_sync(0xffffffffu, p ? b : -1 - lane);
        const bool conflict = p && (peers & lt) != 0;
        const unsigned cm = __ballot_sync(0xffffffffu, conflict);
        const int first = cm ? __ffs(cm) - 1 : 32;
        const bool commit = p && lane < first;
        const unsigned commits = __ballot_sync(0xffffffffu, commit);
        if (commit) {
          frame_mp[b] = src;
          held[b] = hval;
          if (check_ori) {
            const int e = nacc + __popc(commits & lt);
            acc_idx[e] = b;
            acc_bin[e] = __float_as_int(angle);  // the source angle; turned into the bin below
          }
        }
        nacc += __popc(commits);
        const unsigned pm = __ballot_sync(0xffffffffu, p);
        if (lane == 0) { s_done[rp] = commits; s_left[rp] = pm & ~commits; }
      }
      __syncthreads();
      if (s_done[rp] >> w & 1u) pend = false;
      const unsigned left = s_left[rp];
      rp ^= 1;
      if (!left) break;
    }
    cnt = cnt_n; off = off_n;
    cnt_n = cnt_nn; off_n = off_nn;
#pragma unroll
    for (int kk = 0; kk < QR_SLOTS; ++kk) ent[kk] = ent_n[kk];
  }
  if (w != 0) return;
  int nmatches = nacc;
  if (gridDim.x > 1) {
    // one of several segments: the bins go to the call's histogram, k_query_finish filters
    if (check_ori) {
      __syncwarp();
      for (int e = lane; e < nacc; e += 32) {
        float rot = __fsub_rn(__int_as_float(acc_bin[e]), k[acc_idx[e]].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
        if (bin == HISTO_LENGTH) bin = 0;
        acc_bin[e] = bin;
        atomicAdd(&ghist[bin], 1);
      }
    }
    if (lane == 0) seg_nacc[blockIdx.x] = nacc;
    return;
  }
  if (check_ori) {
    __syncwarp();
    // rotation bins of the accepted matches (:3604-3615), lanes over the matches
    for (int e = lane; e < nacc; e += 32) {
      float rot = __fsub_rn(__int_as_float(acc_bin[e]), k[acc_idx[e]].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
      if (bin == HISTO_LENGTH) bin = 0;
      acc_bin[e] = bin;
      atomicAdd(&s_hist[bin], 1);
    }
    __syncwarp();
    int max1 = 0, max2 = 0, max3 = 0, i1_ = -1, i2_ = -1, i3_ = -1;  // every lane computes the same maxima
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      const int sv = s_hist[i];
      if (sv > max1) { max3 = max2; max2 = max1; max1 = sv; i3_ = i2_; i2_ = i1_; i1_ = i; }
      else if (sv > max2) { max3 = max2; max2 = sv; i3_ = i2_; i2_ = i; }
      else if (sv > max3) { max3 = sv; i3_ = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2_ = -1; i3_ = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3_ = -1; }
    int removed = 0;
    for (int e = lane; e < nacc; e += 32) {
      const int bin = acc_bin[e];
      if (bin != i1_ && bin != i2_ && bin != i3_) { frame_mp[acc_idx[e]] = -1; ++removed; }  // :3627-3633
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
  }
  if (lane == 0) *nmatches_out = nmatches;
}

// Rotation-consistency filter of a call whose queries were walked as several segments (:3617-3636): one warp.
__global__ void __launch_bounds__(32) k_query_finish(QuerySegs segs, int check_ori, const int32_t* __restrict__ acc_idx,
                                                     const int32_t* __restrict__ acc_bin, const int* __restrict__ ghist,
                                                     const int* __restrict__ seg_nacc, int32_t* __restrict__ frame_mp,
                                                     int* __restrict__ nmatches_out) {
  const int lane = threadIdx.x;
  __shared__ int s_h[HISTO_LENGTH], s_na[8];
  if (lane < HISTO_LENGTH) s_h[lane] = ghist[lane];  // one round trip to HBM for everything the serial part reads
  if (lane < segs.n) s_na[lane] = seg_nacc[lane];
  __syncwarp();
  int nmatches = 0;
  for (int sg = 0; sg < segs.n; ++sg) nmatches += s_na[sg];
  if (check_ori) {
    int max1 = 0, max2 = 0, max3 = 0, i1_ = -1, i2_ = -1, i3_ = -1;  // every lane computes the same maxima
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      const int sv = s_h[i];
      if (sv > max1) { max3 = max2; max2 = max1; max1 = sv; i3_ = i2_; i2_ = i1_; i1_ = i; }
      else if (sv > max2) { max3 = max2; max2 = sv; i3_ = i2_; i2_ = i; }
      else if (sv > max3) { max3 = sv; i3_ = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2_ = -1; i3_ = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3_ = -1; }
    int removed = 0;
    for (int sg = 0; sg < segs.n; ++sg) {
      const int base = segs.start[sg], na = s_na[sg];
      for (int e = lane; e < na; e += 32) {
        const int bin = acc_bin[base + e];
        if (bin != i1_ && bin != i2_ && bin != i3_) { frame_mp[acc_idx[base + e]] = -1; ++removed; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
  }
  if (lane == 0) *nmatches_out = nmatches;
}

// Device side shared by the overloads: per-camera grids, candidates, ordered resolve.
int run_projected(orbm_matcher* m, const orbx_keypoint* k, const uint8_t* desc, const float* u_right, const int32_t* cam_of,
                  int n_cams, int n, orbm_bounds bounds, const std::vector<ProjQuery>& q_in, const uint8_t* src_desc, int n_src,
                  int th_dist, int check_ori, int any_point_blocks, int32_t* frame_mp, const int32_t* frame_mp_obs,
                  int* nmatches) {
  *nmatches = 0;
  const int nq = (int)q_in.size();
  if (n == 0 || nq == 0) return ORBX_OK;
  bool grouped = false;  // consecutive queries of one source point (Sim3 overload: best over both cameras)
  for (int i = 1; i < nq && !grouped; ++i) grouped = q_in[i].src == q_in[i - 1].src;
  // Queries of different cameras are independent (k_query_resolve_cta): ordered by camera, stable, one segment each.
  const bool held_fits = n <= 32768;
  const bool split = !grouped && n_cams > 1 && n_cams <= 8 && held_fits;
  std::vector<ProjQuery> q_sorted;
  QuerySegs segs;
  segs.n = 1;
  segs.start[0] = 0;
  segs.start[1] = nq;
  if (split) {
    q_sorted.reserve(nq);
    segs.n = n_cams;
    for (int c = 0; c < n_cams; ++c) {
      segs.start[c] = (int)q_sorted.size();
      for (const ProjQuery& e : q_in)
        if (e.cam == c) q_sorted.push_back(e);
    }
    segs.start[n_cams] = (int)q_sorted.size();
    if ((int)q_sorted.size() != nq) { m->err = "query with a camera index out of range"; return ORBX_E_INVALID; }
  }
  const std::vector<ProjQuery>& q = split ? q_sorted : q_in;
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  // One device block: [uploaded inputs | working arrays].  The inputs are packed into one pinned host block with the
  // same layout and go up in ONE copy; the outputs (frame_mp, nmatches, the row cursor) come back in one.
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_desc = 0, o_k = up(o_desc + (size_t)n * 32), o_ur = up(o_k + sizeof(orbx_keypoint) * n),
               o_cam = up(o_ur + 4 * (size_t)n), o_fmp = up(o_cam + 4 * (size_t)n), o_fobs = up(o_fmp + 4 * (size_t)n),
               o_sd = up(o_fobs + 4 * (size_t)n), o_q = up(o_sd + (size_t)n_src * 32), o_misc = up(o_q + sizeof(ProjQuery) * nq),
               in_bytes = up(o_misc + 64 * sizeof(int));  // misc: [0] rows needed, [2] nmatches, [8..38) histogram, [40..48) per segment
  const size_t o_held = in_bytes, o_cnt = up(o_held + n), o_off = up(o_cnt + 4 * (size_t)nq), o_ai = up(o_off + 4 * (size_t)nq),
               o_ab = up(o_ai + 4 * (size_t)nq), total_bytes = up(o_ab + 4 * (size_t)nq);
  uint8_t* dev = m->scratch<uint8_t>(8, total_bytes);
  uint8_t* hin = m->pinned(0, in_bytes);
  int* gstart = m->scratch<int>(4, (size_t)n_cams * (GRID_CELLS + 1));
  uint16_t* gitems = m->scratch<uint16_t>(5, (size_t)n_cams * n);
  if (!dev || !hin || !gstart || !gitems) return ORBX_E_CUDA;
  // candidate rows: a capacity that was enough before (starts at 64 per query); an overflow repeats the call once
  // (queries whose rows do not fit are skipped by the fill, so nothing is written out of bounds)
  size_t rows_cap = std::max<size_t>(m->rows_hint, (size_t)nq * m->rows_per_query);
  int total = 0, got = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    uint32_t* rows = m->scratch<uint32_t>(6, rows_cap);
    uint8_t* hout = m->pinned(1, 4 * (size_t)n + 8 * sizeof(int));
    if (!rows || !hout) return ORBX_E_CUDA;
    std::memcpy(hin + o_desc, desc, (size_t)n * 32);
    std::memcpy(hin + o_k, k, sizeof(orbx_keypoint) * n);
    float* h_ur = reinterpret_cast<float*>(hin + o_ur);
    if (u_right) std::memcpy(h_ur, u_right, 4 * (size_t)n); else std::fill(h_ur, h_ur + n, -1.f);
    if (cam_of) std::memcpy(hin + o_cam, cam_of, 4 * (size_t)n); else std::memset(hin + o_cam, 0, 4 * (size_t)n);
    std::memcpy(hin + o_fmp, frame_mp, 4 * (size_t)n);
    if (frame_mp_obs) std::memcpy(hin + o_fobs, frame_mp_obs, 4 * (size_t)n); else std::memset(hin + o_fobs, 0, 4 * (size_t)n);
    std::memcpy(hin + o_sd, src_desc, (size_t)n_src * 32);
    std::memcpy(hin + o_q, q.data(), sizeof(ProjQuery) * nq);
    int* h_misc = reinterpret_cast<int*>(hin + o_misc);  // [0] total candidate rows (scan), [2] nmatches
    std::memset(h_misc, 0, 64 * sizeof(int));
    if (!m->check(cudaMemcpyAsync(dev, hin, in_bytes, cudaMemcpyHostToDevice, st), "H2D projected search")) return ORBX_E_CUDA;
    uint8_t* dd = dev + o_desc;
    orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(dev + o_k);
    float* dur = reinterpret_cast<float*>(dev + o_ur);
    int32_t* dcam = reinterpret_cast<int32_t*>(dev + o_cam);
    int32_t* dfmp = reinterpret_cast<int32_t*>(dev + o_fmp);
    int32_t* dfobs = reinterpret_cast<int32_t*>(dev + o_fobs);
    uint8_t* dsd = dev + o_sd;
    ProjQuery* dq = reinterpret_cast<ProjQuery*>(dev + o_q);
    int* misc = reinterpret_cast<int*>(dev + o_misc);
    uint8_t* dheld = dev + o_held;
    int* drow_cnt = reinterpret_cast<int*>(dev + o_cnt);
    int* drow_off = reinterpret_cast<int*>(dev + o_off);
    int32_t* dacc_idx = reinterpret_cast<int32_t*>(dev + o_ai);
    int32_t* dacc_bin = reinterpret_cast<int32_t*>(dev + o_ab);
    k_build_grid<<<n_cams, 256, 0, st>>>(dk, nullptr, n, n, bounds, gstart, gitems, dcam, 0, 1);  // one CTA per camera
    m->launches++;
    const int blocks = (nq + 7) / 8;
    k_query_candidates<<<blocks, 256, 0, st>>>(dk, dd, dur, bounds, gstart, gitems, n, dq, dsd, nq, 1, drow_cnt, nullptr, nullptr);
    k_scan_exclusive<<<1, 1024, 0, st>>>(drow_cnt, drow_off, nq, misc);  // misc[0] = total rows needed
    k_query_candidates<<<blocks, 256, 0, st>>>(dk, dd, dur, bounds, gstart, gitems, n, dq, dsd, nq, 0, drow_cnt, drow_off, rows,
                                               (long long)rows_cap);
    m->launches += 2;
    const int held_in_smem = held_fits;  // occupancy bytes on chip when they fit beside the 8 KB row window
    const size_t held_bytes = held_in_smem ? (size_t)((n + 15) & ~15) : 0;
    if (grouped) {
      k_query_resolve_best<<<1, 32, held_bytes, st>>>(dq, drow_cnt, drow_off, rows, (int)std::min<size_t>(rows_cap, 0x7fffffff), nq,
                                                      n, dk, th_dist, check_ori, any_point_blocks, dfmp, dfobs, dheld, held_in_smem,
                                                      dacc_idx, dacc_bin, misc + 2, misc, (long long)rows_cap);
    } else {
      k_query_resolve_cta<<<segs.n, 1024, held_bytes, st>>>(dq, drow_cnt, drow_off, rows, segs, n, dk, th_dist, check_ori,
                                                            any_point_blocks, dfmp, dfobs, dheld, held_in_smem, dacc_idx,
                                                            dacc_bin, misc + 2, misc + 8, misc + 40, misc, (long long)rows_cap);
      if (segs.n > 1) {
        k_query_finish<<<1, 32, 0, st>>>(segs, check_ori, dacc_idx, dacc_bin, misc + 8, misc + 40, dfmp, misc + 2);
        m->launches++;
      }
    }
    m->launches += 2;
    // the outputs are adjacent on the device only in part: two small copies into one pinned block, one synchronisation
    cudaMemcpyAsync(hout, dfmp, 4 * (size_t)n, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(hout + 4 * (size_t)n, misc, 8 * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (!m->check(cudaStreamSynchronize(st), "projected search")) return ORBX_E_CUDA;
    const int* r = reinterpret_cast<const int*>(hout + 4 * (size_t)n);
    total = r[0];
    got = r[2];
    if ((size_t)total <= rows_cap) {
      std::memcpy(frame_mp, hout, 4 * (size_t)n);
      break;
    }
    if (attempt == 1) { m->err = "candidate rows overflowed twice"; return ORBX_E_CAPACITY; }
    rows_cap = (size_t)total + (size_t)total / 4;  // the scan counted every query's rows: this is the exact need
  }
  m->rows_hint = std::max(m->rows_hint, (size_t)total + (size_t)total / 4);
  *nmatches = got;
  return m->check(cudaGetLastError(), "projected search launch") ? ORBX_OK : ORBX_E_CUDA;
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int orbm_search_by_projection_frame_host(orbm_matcher* m, const orbx_keypoint* cur_k, const uint8_t* cur_desc,
                                         const float* cur_uright, const int32_t* cur_cam, int n_cur, orbm_bounds b,
                                         const float* scale_factors, int nlevels, orbm_camera cam, const float* Tcw_cur,
                                         const float* Tcw_last, const orbx_keypoint* last_k, const int32_t* last_cam,
                                         const int32_t* last_valid, const float* last_xyz, const uint8_t* last_desc,
                                         const int32_t* last_obs, int n_last, const float* calib, float th, int mono,
                                         int check_ori, int32_t* cur_mp, const int32_t* cur_mp_obs, int* nmatches) {
  if (!m || !cur_k || !cur_desc || !scale_factors || !Tcw_cur || !Tcw_last || !calib || !cur_mp || !nmatches || n_cur < 0 ||
      n_cur > 65535 || n_last < 0 || (n_last && (!last_k || !last_valid || !last_xyz || !last_desc)))
    return ORBX_E_INVALID;
  // host-side projection: src/ORBmatcher.cc:3464-3531, same float evaluation order
  float Rcam21[9], tcam21[3];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) Rcam21[i * 3 + k] = calib[k * 3 + i];
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  const float tcw[3] = {Tcw_cur[3], Tcw_cur[7], Tcw_cur[11]}, tlw[3] = {Tcw_last[3], Tcw_last[7], Tcw_last[11]};
  float twc[3], tlc[3];
  mat3t_mul_vec(Tcw_cur, 4, tcw, -1.f, twc);
  mat3_mul_vec_add(Tcw_last, 4, twc, tlw, 1.f, tlc);
  const bool fwd[2] = {tlc[2] > cam.mb && !mono, tlc[0] > cam.mb && !mono};
  const bool bwd[2] = {-tlc[2] > cam.mb && !mono, -tlc[0] > cam.mb && !mono};
  std::vector<ProjQuery> q;
  q.reserve(n_last);
  for (int i = 0; i < n_last; ++i) {
    if (!last_valid[i]) continue;
    const int c = last_cam ? last_cam[i] : 0;
    if (c < 0 || c > 1 || last_k[i].octave < 0 || last_k[i].octave >= nlevels) { m->err = "bad camera / octave in last frame"; return ORBX_E_INVALID; }
    float x3Dc[3];
    mat3_mul_vec_add(Tcw_cur, 4, last_xyz + 3 * i, tcw, 1.f, x3Dc);
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
    const int oct = last_k[i].octave;
    ProjQuery p;
    p.u = u; p.v = v;
    p.radius = th * scale_factors[oct];
    p.ur = u - cam.mbf * invzc;
    p.use_ur = 1;
    p.angle = last_k[i].angle;
    if (fwd[c]) { p.min_level = oct; p.max_level = -1; }
    else if (bwd[c]) { p.min_level = 0; p.max_level = oct; }
    else { p.min_level = oct - 1; p.max_level = oct + 1; }
    p.cam = c; p.src = i; p.obs = last_obs ? (last_obs[i] > 0) : 1;
    q.push_back(p);
  }
  return run_projected(m, cur_k, cur_desc, cur_uright, cur_cam, 2, n_cur, b, q, last_desc, n_last, TH_HIGH, check_ori, 0,
                       cur_mp, cur_mp_obs, nmatches);
}

int orbm_search_by_projection_keyframe_host(orbm_matcher* m, const orbx_keypoint* cur_k, const uint8_t* cur_desc, int n_cur,
                                            orbm_bounds b, const float* scale_factors, int nlevels, float log_scale_factor,
                                            orbm_camera cam, const float* Tcw_cur, const int32_t* kf_valid,
                                            const float* kf_xyz, const float* kf_max_dist, const float* kf_min_dist,
                                            const float* kf_max_d, const float* kf_angle, const uint8_t* kf_desc, int n_kf,
                                            float th, int orb_dist, int check_ori, int32_t* cur_mp, int* nmatches) {
  if (!m || !cur_k || !cur_desc || !scale_factors || !Tcw_cur || !cur_mp || !nmatches || n_cur < 0 || n_cur > 65535 ||
      n_kf < 0 || nlevels < 1 || (n_kf && (!kf_valid || !kf_xyz || !kf_max_dist || !kf_min_dist || !kf_max_d || !kf_angle || !kf_desc)))
    return ORBX_E_INVALID;
  // host-side projection and scale prediction: src/ORBmatcher.cc:3814-3866, src/MapPoint.cc:602-617
  const float tcw[3] = {Tcw_cur[3], Tcw_cur[7], Tcw_cur[11]};
  float Ow[3];
  mat3t_mul_vec(Tcw_cur, 4, tcw, -1.f, Ow);
  std::vector<ProjQuery> q;
  q.reserve(n_kf);
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
    int lvl = (int)std::ceil(std::log(ratio) / log_scale_factor);
    if (lvl < 0) lvl = 0;
    else if (lvl >= nlevels) lvl = nlevels - 1;
    ProjQuery p;
    p.u = u; p.v = v;
    p.radius = th * scale_factors[lvl];
    p.ur = 0.f; p.use_ur = 0;
    p.angle = kf_angle[i];
    p.min_level = lvl - 1; p.max_level = lvl + 1;
    p.cam = 0; p.src = i; p.obs = 1;
    q.push_back(p);
  }
  return run_projected(m, cur_k, cur_desc, nullptr, nullptr, 1, n_cur, b, q, kf_desc, n_kf, orb_dist, check_ori, 1, cur_mp,
                       nullptr, nmatches);
}

int orbm_search_by_projection_sim3_host(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc,
                                        const int32_t* kf_cam, int n_kf, orbm_bounds b, const float* scale_factors,
                                        int nlevels, float log_scale_factor, orbm_camera cam, const float* Scw,
                                        const float* calib, const int32_t* mp_valid, const float* mp_xyz,
                                        const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist,
                                        const float* mp_max_d, const uint8_t* mp_desc, int n_mp, int th, int32_t* matched,
                                        int* nmatches) {
  if (!m || !kf_k || !kf_desc || !scale_factors || !Scw || !calib || !matched || !nmatches || n_kf < 0 || n_kf > 65535 ||
      n_mp < 0 || nlevels < 1 ||
      (n_mp && (!mp_valid || !mp_xyz || !mp_normal || !mp_max_dist || !mp_min_dist || !mp_max_d || !mp_desc)))
    return ORBX_E_INVALID;
  // host-side projection into both cameras: src/ORBmatcher.cc:581-704, same evaluation order
  float Rcam21[9], tcam21[3];
  {
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = calib[i];
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1. / d;  // cv::Mat::inv() of a 3x3 CV_32F matrix: closed form in double
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
  std::vector<ProjQuery> q;
  q.reserve(2 * (size_t)n_mp);
  for (int i = 0; i < n_mp; ++i) {
    if (!mp_valid[i]) continue;
    const float* p3Dw = mp_xyz + 3 * i;
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
      if (!(u >= b.min_x && u < b.max_x && v >= b.min_y && v < b.max_y)) continue;
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
      ProjQuery p;
      p.u = u; p.v = v;
      p.radius = th * scale_factors[lvl];
      p.ur = 0.f; p.use_ur = 0; p.angle = 0.f;
      p.min_level = lvl - 1; p.max_level = lvl;  // the post-filter of :716-718 folded into the query
      p.cam = camidx; p.src = i; p.obs = 1;
      q.push_back(p);
    }
  }
  return run_projected(m, kf_k, kf_desc, nullptr, kf_cam, 2, n_kf, b, q, mp_desc, n_mp, TH_LOW, 0, 1, matched, nullptr,
                       nmatches);
}

}  // extern "C"
#pragma GCC visibility pop
namespace {
// device part shared by the two Fuse overloads: per-camera grids + one warp per (map point, camera) query
int fuse_run(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc, const float* kf_uright, const int32_t* kf_cam,
             int n_kf, orbm_bounds b, const std::vector<ProjQuery>& q, const uint8_t* mp_desc, int n_mp,
             const float* inv_level_sigma2, int nlevels, int gate, int32_t* best_idx, int* n_fused, int th_dist = TH_LOW) {
  const int nq = (int)q.size();
  if (nq == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  uint8_t* fb = m->scratch<uint8_t>(8, (size_t)n_kf * (32 + sizeof(orbx_keypoint) + 8) + 256);
  uint8_t* qb = m->scratch<uint8_t>(9, (size_t)n_mp * (32 + 8) + (size_t)nq * sizeof(ProjQuery) + 256);
  int* gstart = m->scratch<int>(4, (size_t)2 * (GRID_CELLS + 1));
  uint16_t* gitems = m->scratch<uint16_t>(5, (size_t)2 * n_kf);
  if (!fb || !qb || !gstart || !gitems) return ORBX_E_CUDA;
  uint8_t* dd = fb;
  orbx_keypoint* dk = reinterpret_cast<orbx_keypoint*>(dd + (size_t)n_kf * 32);
  float* dur = reinterpret_cast<float*>(dk + n_kf);
  int32_t* dcam = reinterpret_cast<int32_t*>(dur + n_kf);
  uint8_t* dmd = qb;
  int32_t* dbest = reinterpret_cast<int32_t*>(dmd + (size_t)n_mp * 32);
  ProjQuery* dq = reinterpret_cast<ProjQuery*>(dbest + 2 * (size_t)n_mp);
  cudaMemcpyAsync(dd, kf_desc, (size_t)n_kf * 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dk, kf_k, sizeof(orbx_keypoint) * n_kf, cudaMemcpyHostToDevice, st);
  if (kf_uright) cudaMemcpyAsync(dur, kf_uright, sizeof(float) * n_kf, cudaMemcpyHostToDevice, st);
  if (kf_cam) cudaMemcpyAsync(dcam, kf_cam, sizeof(int32_t) * n_kf, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dmd, mp_desc, (size_t)n_mp * 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dq, q.data(), sizeof(ProjQuery) * nq, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(dbest, 0xFF, sizeof(int32_t) * 2 * (size_t)n_mp, st);
  if (kf_cam) {
    k_build_grid<<<2, 256, 0, st>>>(dk, nullptr, n_kf, n_kf, b, gstart, gitems, dcam, 0, 1);  // one CTA per camera
    m->launches++;
  } else {
    for (int c = 0; c < 2; ++c) {
      k_build_grid<<<1, 256, 0, st>>>(dk, nullptr, n_kf, n_kf, b, gstart + (size_t)c * (GRID_CELLS + 1), gitems + (size_t)c * n_kf,
                                      nullptr, c);
      m->launches++;
    }
  }
  LevelTable lt;
  for (int i = 0; i < ORBX_MAX_LEVELS; ++i) lt.v[i] = (inv_level_sigma2 && i < nlevels) ? inv_level_sigma2[i] : 0.f;
  k_fuse_match<<<(nq + 7) / 8, 256, 0, st>>>(dk, dd, dur, b, gstart, gitems, n_kf, dq, dmd, nq, lt, gate, th_dist, dbest);
  m->launches++;
  cudaMemcpyAsync(best_idx, dbest, sizeof(int32_t) * 2 * (size_t)n_mp, cudaMemcpyDeviceToHost, st);
  if (!m->check(cudaStreamSynchronize(st), "fuse")) return ORBX_E_CUDA;
  int nf = 0;
  for (int i = 0; i < 2 * n_mp; ++i) nf += best_idx[i] >= 0;
  *n_fused = nf;
  return m->check(cudaGetLastError(), "fuse launch") ? ORBX_OK : ORBX_E_CUDA;
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int orbm_fuse_host(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc, const float* kf_uright,
                   const int32_t* kf_cam, int n_kf, orbm_bounds b, const float* scale_factors, const float* inv_level_sigma2,
                   int nlevels, float log_scale_factor, orbm_camera cam, const float* Tcw, const float* Ow, const float* calib,
                   const int32_t* mp_valid, const float* mp_xyz, const float* mp_normal, const float* mp_max_dist,
                   const float* mp_min_dist, const float* mp_max_d, const uint8_t* mp_desc, int n_mp, float th,
                   int32_t* best_idx, int* n_fused) {
  if (!m || !scale_factors || !inv_level_sigma2 || !Tcw || !Ow || !calib || !best_idx || !n_fused || n_kf < 0 || n_kf > 65535 ||
      n_mp < 0 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || (n_kf && (!kf_k || !kf_desc || !kf_uright)) ||
      (n_mp && (!mp_valid || !mp_xyz || !mp_normal || !mp_max_dist || !mp_min_dist || !mp_max_d || !mp_desc)))
    return ORBX_E_INVALID;
  *n_fused = 0;
  for (int i = 0; i < 2 * n_mp; ++i) best_idx[i] = -1;
  if (n_kf == 0 || n_mp == 0) return ORBX_OK;
  for (int i = 0; i < n_kf; ++i)
    if (kf_k[i].octave < 0 || kf_k[i].octave >= nlevels) { m->err = "key frame: octave out of range"; return ORBX_E_INVALID; }
  // host-side projection into both cameras (:1996-2071), float evaluation order of the cv::Mat expressions
  float Rcam21[9], tcam21[3];
  {
    double S[9];
    for (int i = 0; i < 9; ++i) S[i] = calib[i];
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    d = 1. / d;  // cv::Mat::inv() of a 3x3 CV_32F matrix: closed form in double
    Rcam21[0] = (float)((S[4] * S[8] - S[5] * S[7]) * d); Rcam21[1] = (float)((S[2] * S[7] - S[1] * S[8]) * d);
    Rcam21[2] = (float)((S[1] * S[5] - S[2] * S[4]) * d); Rcam21[3] = (float)((S[5] * S[6] - S[3] * S[8]) * d);
    Rcam21[4] = (float)((S[0] * S[8] - S[2] * S[6]) * d); Rcam21[5] = (float)((S[2] * S[3] - S[0] * S[5]) * d);
    Rcam21[6] = (float)((S[3] * S[7] - S[4] * S[6]) * d); Rcam21[7] = (float)((S[1] * S[6] - S[0] * S[7]) * d);
    Rcam21[8] = (float)((S[0] * S[4] - S[1] * S[3]) * d);
  }
  mat3_mul_vec_add(Rcam21, 3, calib + 9, nullptr, -1.f, tcam21);
  const float tcw[3] = {Tcw[3], Tcw[7], Tcw[11]};
  // camera 2 (:2031): ((Rcam21*Rcw)*p3Dw + Rcam21*tcw) + tcam21, products materialised by cv::MatExpr
  float M[9], Rt[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float acc = 0.f;
      for (int kk = 0; kk < 3; ++kk) acc += Rcam21[i * 3 + kk] * Tcw[kk * 4 + j];
      M[i * 3 + j] = acc;
    }
  mat3_mul_vec_add(Rcam21, 3, tcw, nullptr, 1.f, Rt);
  std::vector<ProjQuery> q;
  q.reserve(2 * (size_t)n_mp);
  for (int i = 0; i < n_mp; ++i) {
    if (!mp_valid[i]) continue;
    const float* p3Dw = mp_xyz + 3 * i;
    for (int c = 0; c < 2; ++c) {
      float p3Dc[3];
      if (c == 0) {
        mat3_mul_vec_add(Tcw, 4, p3Dw, tcw, 1.f, p3Dc);
      } else {
        float m1[3];
        mat3_mul_vec_add(M, 3, p3Dw, nullptr, 1.f, m1);
        for (int kk = 0; kk < 3; ++kk) p3Dc[kk] = (m1[kk] + Rt[kk]) + tcam21[kk];
      }
      if (p3Dc[2] < 0.0f) continue;
      const float invz = 1 / p3Dc[2];
      const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
      const float u = cam.fx * x + cam.cx, v = cam.fy * y + cam.cy;
      if (!(u >= b.min_x && u < b.max_x && v >= b.min_y && v < b.max_y)) continue;
      float PO[3];
      double n2 = 0, dotn = 0;
      for (int kk = 0; kk < 3; ++kk) {
        PO[kk] = p3Dw[kk] - Ow[3 * c + kk];
        n2 += (double)PO[kk] * (double)PO[kk];
        dotn += (double)PO[kk] * (double)mp_normal[3 * i + kk];
      }
      const float dist3D = (float)std::sqrt(n2);
      if (dist3D < mp_min_dist[i] || dist3D > mp_max_dist[i]) continue;
      if (dotn < 0.5 * dist3D) continue;
      const float ratio = mp_max_d[i] / dist3D;
      int lvl = (int)std::ceil(std::log(ratio) / log_scale_factor);
      if (lvl < 0) lvl = 0;
      else if (lvl >= nlevels) lvl = nlevels - 1;
      ProjQuery p;
      p.u = u; p.v = v;
      p.radius = th * scale_factors[lvl];
      p.ur = u - cam.mbf * invz;
      p.use_ur = 1; p.angle = 0.f;
      p.min_level = lvl - 1; p.max_level = lvl;  // the filter of :2100-2101 folded into the query
      p.cam = c; p.src = i; p.obs = 1;
      q.push_back(p);
    }
  }
  return fuse_run(m, kf_k, kf_desc, kf_uright, kf_cam, n_kf, b, q, mp_desc, n_mp, inv_level_sigma2, nlevels, 1, best_idx, n_fused);
}

int orbm_fuse_sim3_host(orbm_matcher* m, const orbx_keypoint* kf_k, const uint8_t* kf_desc, const int32_t* kf_cam, int n_kf,
                        orbm_bounds b, const float* scale_factors, int nlevels, float log_scale_factor, orbm_camera cam,
                        const float* Scw, const float* calib, const int32_t* mp_valid, const float* mp_xyz,
                        const float* mp_normal, const float* mp_max_dist, const float* mp_min_dist, const float* mp_max_d,
                        const uint8_t* mp_desc, int n_mp, float th, int32_t* best_idx, int* n_fused) {
  if (!m || !scale_factors || !Scw || !calib || !best_idx || !n_fused || n_kf < 0 || n_kf > 65535 || n_mp < 0 || nlevels < 1 ||
      nlevels > ORBX_MAX_LEVELS || (n_kf && (!kf_k || !kf_desc)) ||
      (n_mp && (!mp_valid || !mp_xyz || !mp_normal || !mp_max_dist || !mp_min_dist || !mp_max_d || !mp_desc)))
    return ORBX_E_INVALID;
  *n_fused = 0;
  for (int i = 0; i < 2 * n_mp; ++i) best_idx[i] = -1;
  if (n_kf == 0 || n_mp == 0) return ORBX_OK;
  // host-side projection (:2218-2300), float evaluation order of the cv::Mat expressions
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
  double RtT12[3];  // Rcw.t()*tcam12 in double: `PO - Rcw.t()*tcam12` is one gemm(alpha=-1, C=PO, beta=1, GEMM_1_T)
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
  std::vector<ProjQuery> q;
  q.reserve(2 * (size_t)n_mp);
  for (int i = 0; i < n_mp; ++i) {
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
      ProjQuery p;
      p.u = u; p.v = v;
      p.radius = th * scale_factors[lvl];
      p.ur = 0.f; p.use_ur = 0; p.angle = 0.f;
      p.min_level = lvl - 1; p.max_level = lvl;
      p.cam = c; p.src = i; p.obs = 1;
      q.push_back(p);
    }
  }
  return fuse_run(m, kf_k, kf_desc, nullptr, kf_cam, n_kf, b, q, mp_desc, n_mp, nullptr, nlevels, 0, best_idx, n_fused);
}

int orbm_search_by_sim3_host(orbm_matcher* m, const orbx_keypoint* k1, const uint8_t* d1, const int32_t* cam1, int n1,
                             const float* T1w, const orbx_keypoint* k2, const uint8_t* d2, const int32_t* cam2, int n2,
                             const float* T2w, orbm_bounds b, const float* scale_factors, int nlevels, float log_scale_factor,
                             orbm_camera cam, float s12, const float* R12, const float* t12, const float* calib,
                             const int32_t* mp1_valid, const float* mp1_xyz, const float* mp1_max_dist,
                             const float* mp1_min_dist, const float* mp1_max_d, const uint8_t* mp1_desc,
                             const int32_t* mp2_valid, const float* mp2_xyz, const float* mp2_max_dist,
                             const float* mp2_min_dist, const float* mp2_max_d, const uint8_t* mp2_desc, float th,
                             int32_t* match12, int* n_found) {
  if (!m || !T1w || !T2w || !scale_factors || !R12 || !t12 || !calib || !match12 || !n_found || n1 < 0 || n2 < 0 || n1 > 65535 ||
      n2 > 65535 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || s12 == 0.f ||
      (n1 && (!k1 || !d1 || !mp1_valid || !mp1_xyz || !mp1_max_dist || !mp1_min_dist || !mp1_max_d || !mp1_desc)) ||
      (n2 && (!k2 || !d2 || !mp2_valid || !mp2_xyz || !mp2_max_dist || !mp2_min_dist || !mp2_max_d || !mp2_desc)))
    return ORBX_E_INVALID;
  *n_found = 0;
  for (int i = 0; i < n1; ++i) match12[i] = -1;
  if (n1 == 0 || n2 == 0) return ORBX_OK;
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
  // sR12 = s12*R12; sR21 = (1.0/s12)*R12.t(); t21 = -sR21*t12   (:2838-2840)
  float sR12[9], sR21[9], t21[3];
  const float inv_s12 = (float)(1.0 / (double)s12);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      sR12[i * 3 + j] = R12[i * 3 + j] * s12;
      sR21[i * 3 + j] = R12[j * 3 + i] * inv_s12;
    }
  mat3_mul_vec_add(sR21, 3, t12, nullptr, -1.f, t21);
  const float t1w[3] = {T1w[3], T1w[7], T1w[11]}, t2w[3] = {T2w[3], T2w[7], T2w[11]};
  // queries of one direction (:2867-2900 / :2990-3023): points of key frame A into key frame B
  auto project = [&](int nA, const int32_t* camA, const float* TAw, const float* tAw, const float* sRBA, const float* tBA,
                     const int32_t* valid, const float* xyz, const float* maxd, const float* mind, const float* maxD,
                     std::vector<ProjQuery>& q) {
    q.clear();
    for (int i = 0; i < nA; ++i) {
      if (!valid[i]) continue;
      const int camIdx = camA ? camA[i] : 0;
      if (camIdx < 0 || camIdx > 1) continue;
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
      ProjQuery p;
      p.u = u; p.v = v;
      p.radius = th * scale_factors[lvl];
      p.ur = 0.f; p.use_ur = 0; p.angle = 0.f;
      p.min_level = lvl - 1; p.max_level = lvl;
      p.cam = camIdx; p.src = i; p.obs = 1;
      q.push_back(p);
    }
  };
  std::vector<ProjQuery> q;
  std::vector<int32_t> best1(2 * (size_t)n1, -1), best2(2 * (size_t)n2, -1);
  int dummy = 0;
  project(n1, cam1, T1w, t1w, sR21, t21, mp1_valid, mp1_xyz, mp1_max_dist, mp1_min_dist, mp1_max_d, q);
  int rc = fuse_run(m, k2, d2, nullptr, cam2, n2, b, q, mp1_desc, n1, nullptr, nlevels, 0, best1.data(), &dummy, TH_HIGH);
  if (rc != ORBX_OK) return rc;
  project(n2, cam2, T2w, t2w, sR12, t12, mp2_valid, mp2_xyz, mp2_max_dist, mp2_min_dist, mp2_max_d, q);
  rc = fuse_run(m, k1, d1, nullptr, cam1, n1, b, q, mp2_desc, n2, nullptr, nlevels, 0, best2.data(), &dummy, TH_HIGH);
  if (rc != ORBX_OK) return rc;
  // mutual consistency (:3107-3122); a point has one query, in the grid of its own camera
  auto match_of = [](const std::vector<int32_t>& best, const int32_t* camA, int i) {
    const int c = camA ? camA[i] : 0;
    return (c == 0 || c == 1) ? best[2 * (size_t)i + c] : -1;
  };
  int nf = 0;
  for (int i1 = 0; i1 < n1; ++i1) {
    const int idx2 = match_of(best1, cam1, i1);
    if (idx2 >= 0 && match_of(best2, cam2, idx2) == i1) { match12[i1] = idx2; nf++; }
  }
  *n_found = nf;
  return ORBX_OK;
}

int orbm_set_vocabulary(orbm_matcher* m, const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc,
                        const int32_t* word_id, const double* node_weight, int n_nodes, int L) {
  if (!m || !child_start || !node_desc || !word_id || !node_weight || n_nodes < 1 || L < 1 || child_start[0] != 0) return ORBX_E_INVALID;
  const int n_children = child_start[n_nodes];
  if (n_children < 0 || (n_children && !child_ids)) return ORBX_E_INVALID;
  for (int i = 0; i < n_nodes; ++i)
    if (child_start[i + 1] < child_start[i] || child_start[i + 1] - child_start[i] > 0xFFFFF) { m->err = "vocabulary: bad child table"; return ORBX_E_INVALID; }
  for (int c = 0; c < n_children; ++c)
    if (child_ids[c] <= 0 || child_ids[c] >= n_nodes) { m->err = "vocabulary: child id out of range"; return ORBX_E_INVALID; }
  if (child_start[1] == 0) { m->err = "vocabulary: the root has no children"; return ORBX_E_INVALID; }
  OrbDeviceGuard dev_guard(m->device);
  cudaStreamSynchronize(m->stream);
  cudaFree(m->voc_child_start); cudaFree(m->voc_child_ids); cudaFree(m->voc_desc);
  m->voc_child_start = m->voc_child_ids = nullptr; m->voc_desc = nullptr; m->voc_nodes = 0;
  if (!m->check(cudaMalloc(&m->voc_child_start, sizeof(int32_t) * (n_nodes + 1)), "cudaMalloc(vocabulary)") ||
      !m->check(cudaMalloc(&m->voc_child_ids, sizeof(int32_t) * std::max(n_children, 1)), "cudaMalloc(vocabulary)") ||
      !m->check(cudaMalloc(&m->voc_desc, (size_t)n_nodes * 32), "cudaMalloc(vocabulary)"))
    return ORBX_E_CUDA;
  cudaMemcpy(m->voc_child_start, child_start, sizeof(int32_t) * (n_nodes + 1), cudaMemcpyHostToDevice);
  if (n_children) cudaMemcpy(m->voc_child_ids, child_ids, sizeof(int32_t) * n_children, cudaMemcpyHostToDevice);
  cudaMemcpy(m->voc_desc, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice);
  m->voc_word.assign(word_id, word_id + n_nodes);
  m->voc_weight.assign(node_weight, node_weight + n_nodes);
  m->voc_nodes = n_nodes;
  m->voc_L = L;
  return m->check(cudaGetLastError(), "vocabulary upload") ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_bow_transform_host(orbm_matcher* m, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight,
                            int32_t* bow_word, double* bow_value, int32_t* n_bow, int32_t* fv_node, int32_t* fv_start,
                            int32_t* fv_items, int32_t* n_fv) {
  if (!m || n < 0 || (n && !desc) || !n_bow || !n_fv || !fv_start || (n && (!bow_word || !bow_value || !fv_node || !fv_items)))
    return ORBX_E_INVALID;
  if (!m->voc_nodes) { m->err = "no vocabulary set (orbm_set_vocabulary)"; return ORBX_E_STATE; }
  *n_bow = *n_fv = 0;
  fv_start[0] = 0;
  if (n == 0) return ORBX_OK;
  OrbDeviceGuard dev_guard(m->device);
  cudaStream_t st = m->stream;
  uint8_t* dd = m->scratch<uint8_t>(8, (size_t)n * 32);
  int32_t* dout = m->scratch<int32_t>(4, 2 * (size_t)n);
  if (!dd || !dout) return ORBX_E_CUDA;
  cudaMemcpyAsync(dd, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st);
  const int nid_level = m->voc_L - levelsup;
  k_bow_descend<<<(n + 7) / 8, 256, 0, st>>>(m->voc_child_start, m->voc_child_ids, m->voc_desc, dd, n, nid_level, dout, dout + n);
  m->launches += 1;
  std::vector<int32_t> h(2 * (size_t)n);
  cudaMemcpyAsync(h.data(), dout, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, st);
  if (!m->check(cudaStreamSynchronize(st), "bow transform")) return ORBX_E_CUDA;
  if (!m->check(cudaGetLastError(), "bow transform launch")) return ORBX_E_CUDA;
  // BowVector / FeatureVector assembly (TemplatedVocabulary.h:1149-1194, BowVector.cpp:34-46, 62-84): std::map
  // semantics on sorted vectors, weights summed in feature order, L1 norm summed in word order, all in double
  std::vector<std::pair<int, double>> bow;
  std::vector<std::pair<int, std::vector<int>>> fv;
  for (int i = 0; i < n; ++i) {
    const int leaf_id = h[i], nid = nid_level <= 0 ? 0 : h[n + i];
    const int w_id = m->voc_word[leaf_id];
    const double w = m->voc_weight[leaf_id];
    if (word) word[i] = w_id;
    if (node) node[i] = nid;
    if (weight) weight[i] = w;
    if (w > 0) {
      auto vit = std::lower_bound(bow.begin(), bow.end(), w_id, [](const std::pair<int, double>& a, int b) { return a.first < b; });
      if (vit != bow.end() && vit->first == w_id) vit->second += w;
      else bow.insert(vit, std::make_pair(w_id, w));
      auto fit = std::lower_bound(fv.begin(), fv.end(), nid,
                                  [](const std::pair<int, std::vector<int>>& a, int b) { return a.first < b; });
      if (fit != fv.end() && fit->first == nid) fit->second.push_back(i);
      else fv.insert(fit, std::make_pair(nid, std::vector<int>(1, i)));
    }
  }
  double norm = 0.0;
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
  return ORBX_OK;
}

#pragma GCC visibility pop
}  // extern "C"
