// C-ABI of the extractor (include/orb_b200.h): handle, geometry, HBM workspace, launches.
// Replaces class ORBextractor (reference include/ORBextractor.h:45-112, src/ORBextractor.cc).
#include "device_guard.h"
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "extract_kernels.h"

#ifdef ORB_OT_TIMING
namespace orbk { void dump_octree_marks(); }
#endif

namespace {

thread_local std::string g_create_error;

inline int cv_round(float v) { return (int)lrintf(v); }  // cvRound: round-half-even
inline int cv_round(double v) { return (int)lrint(v); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct orbx_extractor {
  orbx_config cfg;
  int device;
  cudaStream_t stream = nullptr;      // stream in use
  cudaStream_t own_stream = nullptr;  // created by the handle
  std::string err;
  long long launches = 0;
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int> quota;
  orbk::OrbGeomHost gh;
  // HBM workspace (max_batch frames)
  uint8_t* d_pyr = nullptr;
  uint8_t* d_blur = nullptr;
  uint32_t* d_cand = nullptr;
  int* d_cell_count = nullptr;
  uint32_t* d_keys = nullptr;
  uint16_t* d_knode = nullptr;
  uint32_t* d_sel = nullptr;
  int* d_sel_count = nullptr;
  // staging for the host entry points
  uint8_t* d_img = nullptr;
  orbx_keypoint* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int stage_cap = 0;
  // Small host calls (the per-image call of the drop-in: one frame at a time) are latency, not throughput: the 11 launches
  // of a batch are replayed from a CUDA graph (one per batch size, captured on first use), and when the whole staging
  // block is small its three result arrays come back in ONE copy through pinned memory.
  struct BatchGraph { int nb, cap; cudaGraphExec_t exec; long long launches; };
  std::vector<BatchGraph> graphs;
  bool fork_blur = false;       // set while a small batch is captured: the blur (it needs the pyramid only) runs on a side
  cudaStream_t side = nullptr;  // branch of the graph next to FAST + octree - at one frame every kernel is a latency chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint8_t* d_stage = nullptr;   // [kps | desc | counts] in one allocation (d_kps, d_desc, d_stage_counts point into it)
  int* d_stage_counts = nullptr;
  uint8_t* h_stage = nullptr;   // pinned mirror of d_stage when it is small
  size_t stage_bytes = 0;
  int last_batch = 0;  // frames resident in the workspace (for taps)
  orbk::OrbLevel0 last_l0 = {nullptr, 0, 0};  // level-0 view of the last batch
  uint8_t* d_l0 = nullptr;  // aligned copy of the frames, only when the caller's buffer is not 4-byte aligned
  int img_pitch = 0;         // row pitch of d_img / d_l0 (width rounded up to 16)
  // per-stage device timing: a ring of event sets, harvested into running sums
  static const int kRing = 32;
  bool profiling = false;
  cudaEvent_t ev[kRing][6] = {};
  bool ev_pending[kRing] = {};
  int ev_next = 0;
  double stage_sum_ms[5] = {0, 0, 0, 0, 0};
  long long stage_calls = 0;

  bool harvest(int slot) {
    if (!ev_pending[slot]) return true;
    if (!check(cudaEventSynchronize(ev[slot][5]), "event sync")) return false;
    for (int i = 0; i < 5; ++i) {
      float ms = 0;
      if (!check(cudaEventElapsedTime(&ms, ev[slot][i], ev[slot][i + 1]), "event elapsed")) return false;
      stage_sum_ms[i] += ms;
    }
    ++stage_calls;
    ev_pending[slot] = false;
    return true;
  }

  bool check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
};

namespace {

// ORBextractor ctor tables, src/ORBextractor.cc:411-447
void build_tables(orbx_extractor* h) {
  const int nl = h->cfg.nlevels;
  const double sf = (double)h->cfg.scale_factor;  // the reference keeps the float arg in a double member
  h->scale.assign(nl, 1.f);
  h->sigma2.assign(nl, 1.f);
  for (int i = 1; i < nl; ++i) {
    h->scale[i] = (float)(h->scale[i - 1] * sf);
    h->sigma2[i] = h->scale[i] * h->scale[i];
  }
  h->inv_scale.resize(nl);
  h->inv_sigma2.resize(nl);
  for (int i = 0; i < nl; ++i) {
    h->inv_scale[i] = 1.0f / h->scale[i];
    h->inv_sigma2[i] = 1.0f / h->sigma2[i];
  }
  h->quota.resize(nl);
  const float factor = (float)(1.0f / sf);
  float desired = h->cfg.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    h->quota[l] = cv_round(desired);
    sum += h->quota[l];
    desired *= factor;
  }
  h->quota[nl - 1] = std::max(h->cfg.nfeatures - sum, 0);
}

// cv::resize INTER_LINEAR tables (OpenCV imgproc/resize.cpp), level l <- l-1
void build_resize_tables(int sw, int sh, int dw, int dh, std::vector<OrbXTap>& xt, std::vector<OrbYTap>& yt) {
  const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    OrbXTap t;
    t.sx = (unsigned short)sx;
    t.a0 = (unsigned short)cv_round((1.f - fx) * 2048.f);
    t.a1 = (unsigned short)cv_round(fx * 2048.f);
    t.pad = 0;
    xt.push_back(t);
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    OrbYTap t;
    t.sy0 = (unsigned short)std::min(std::max(sy, 0), sh - 1);
    t.sy1 = (unsigned short)std::min(std::max(sy + 1, 0), sh - 1);
    t.b0 = (unsigned short)cv_round((1.f - fy) * 2048.f);
    t.b1 = (unsigned short)cv_round(fy * 2048.f);
    yt.push_back(t);
  }
}

// Level sizes (:1113-1117), FAST cell grid (:774-807), octree roots (:544-564), HBM offsets.
bool build_geometry(orbx_extractor* h, std::vector<OrbCell>& cells, std::vector<OrbXTap>& xt,
                    std::vector<OrbYTap>& yt) {
  OrbGeom& g = h->gh.g;
  std::memset(&g, 0, sizeof(g));
  const int nl = h->cfg.nlevels;
  g.nlevels = nl;
  g.width = h->cfg.width;
  g.height = h->cfg.height;
  g.ini_th = std::min(std::max(h->cfg.ini_th_fast, 0), 255);
  g.min_th = std::min(std::max(h->cfg.min_th_fast, 0), 255);
  if (g.min_th > g.ini_th) { h->err = "minThFAST > iniThFAST is not supported"; return false; }
  if (h->cfg.scale_factor > 2.5f) { h->err = "scaleFactor > 2.5 is not supported by the pyramid kernel"; return false; }
  size_t pyr_off = 0, blur_off = 0, cand_off = 0, key_off = 0;
  int sel_off = 0, tile_base = 0, max_cells_level = 0;
  g.ot_cap = 0;
  for (int l = 0; l < nl; ++l) {
    OrbLevelGeom& L = g.lv[l];
    L.w = cv_round((float)g.width * h->inv_scale[l]);
    L.h = cv_round((float)g.height * h->inv_scale[l]);
    if (L.w > 4000 || L.h > 4000) { h->err = "image larger than 4000 px is not supported"; return false; }
    L.pitch = (int)align_up(L.w, 16);  // levels are stored without the 19-px border (see extract_kernels.cu)
    L.bpitch = (int)align_up(L.w, 16);
    L.pyr_off = (unsigned)pyr_off;
    if (l > 0) pyr_off += align_up((size_t)L.pitch * L.h, 256);  // level 0 = the caller's frame, read in place
    L.blur_off = (unsigned)blur_off;
    blur_off += align_up((size_t)L.bpitch * L.h, 256);
    L.scale = h->scale[l];
    L.patch_size = (float)(int)(31 * h->scale[l]);  // PATCH_SIZE*mvScaleFactor[level] -> int (:838)
    L.quota = h->quota[l];
    // cell grid
    const int minB = ORB_MIN_BORDER, maxBX = L.w - ORB_EDGE + 3, maxBY = L.h - ORB_EDGE + 3;
    const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
    L.n_cols = (int)(width / 30.f);
    L.n_rows = (int)(height / 30.f);
    if (L.n_cols < 1 || L.n_rows < 1) {
      h->err = "pyramid level " + std::to_string(l) + " is too small for a 30-px FAST cell";
      return false;
    }
    L.w_cell = (int)std::ceil(width / L.n_cols);
    L.h_cell = (int)std::ceil(height / L.n_rows);
    if (L.w_cell + 6 > ORB_CELL_MAX || L.h_cell + 6 > ORB_CELL_MAX) { h->err = "FAST cell too large"; return false; }
    L.cell_base = (int)cells.size();
    for (int i = 0; i < L.n_rows; ++i) {
      const float iniY = (float)(minB + i * L.h_cell);
      float maxY = iniY + L.h_cell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < L.n_cols; ++j) {
        const float iniX = (float)(minB + j * L.w_cell);
        float maxX = iniX + L.w_cell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        OrbCell c;
        c.level = (short)l;
        c.ini_x = (short)iniX;
        c.ini_y = (short)iniY;
        c.cw = (short)((int)maxX - (int)iniX);
        c.ch = (short)((int)maxY - (int)iniY);
        c.off_x = (short)(j * L.w_cell);
        c.off_y = (short)(i * L.h_cell);
        c.slot = (short)((int)cells.size() - L.cell_base);
        c.pad = 0;
        cells.push_back(c);
      }
    }
    L.n_cells = (int)cells.size() - L.cell_base;
    max_cells_level = std::max(max_cells_level, L.n_cells);
    // NMS keeps no two 8-adjacent pixels: at most ceil(w/2)*ceil(h/2) per tested w_cell x h_cell area
    L.cand_cap = ((L.w_cell + 1) / 2) * ((L.h_cell + 1) / 2);
    L.cand_off = (unsigned)cand_off;
    L.key_off = (unsigned)key_off;
    L.key_cap = L.n_cells * L.cand_cap;
    for (int ci = L.cell_base; ci < (int)cells.size(); ++ci) {
      OrbCell& c = cells[ci];
      c.a0 = (short)(c.ini_x & 3);
      c.tile_off = L.pyr_off + (unsigned)c.ini_y * (unsigned)L.pitch + (unsigned)(c.ini_x - c.a0);  // levels >= 1
      c.pitch = (unsigned short)L.pitch;
      c.cand_cap = (unsigned short)L.cand_cap;
      c.cand_slot_off = (unsigned)cand_off + (unsigned)c.slot * (unsigned)L.cand_cap;
    }
    cand_off += (size_t)L.key_cap;
    key_off += (size_t)L.key_cap;
    // octree roots
    const int W = maxBX - minB, H = maxBY - minB;
    OtRoots& r = L.roots;
    r.n_ini = (int)std::round((float)W / (float)H);
    if (r.n_ini < 1 || r.n_ini > OT_MAX_ROOTS) { h->err = "unsupported aspect ratio for the octree roots"; return false; }
    r.hx = (float)W / r.n_ini;
    for (int i = 0; i <= r.n_ini; ++i) r.root_x[i] = (int)(r.hx * (float)i);
    r.height = H;
    L.sel_cap = std::max(L.quota + 3, 4 * r.n_ini) + 1;
    L.sel_off = sel_off;
    sel_off += L.sel_cap;
    g.ot_cap = std::max(g.ot_cap, L.sel_cap);
    // blur tiles
    L.blur_tiles_x = (L.w + ORB_BLUR_TW - 1) / ORB_BLUR_TW;
    L.blur_tiles_y = (L.h + ORB_BLUR_TH - 1) / ORB_BLUR_TH;
    L.blur_tile_base = tile_base;
    tile_base += L.blur_tiles_x * L.blur_tiles_y;
    // resize tables
    if (l > 0) {
      L.xtab_off = (int)xt.size();
      L.ytab_off = (int)yt.size();
      build_resize_tables(g.lv[l - 1].w, g.lv[l - 1].h, L.w, L.h, xt, yt);
    }
  }
  if (g.ot_cap > OT_MAX_NODES) { h->err = "nfeatures too large for the octree kernel"; return false; }
  g.n_cells = (int)cells.size();
  g.n_blur_tiles = tile_base;
  g.ot_scan_cap = std::max(g.ot_cap, max_cells_level) + 1;
  g.kp_cap_frame = sel_off;
  g.pyr_frame_bytes = pyr_off + 256;  // slack: the resize window may touch a few bytes past a level's last row
  g.blur_frame_bytes = blur_off;
  g.cand_frame_u32 = cand_off;
  g.key_frame_u32 = key_off;
  if (xt.empty()) { xt.push_back(OrbXTap{0, 0, 0, 0}); yt.push_back(OrbYTap{0, 0, 0, 0}); }
  return true;
}

// Work list of the FAST kernel: every row of cells of every level, cut into chunks of whole cells that fit the
// ORB_BAND_MAX_PX tested pixels a warp covers per image row (chunks of a row are balanced: 20 cells -> 7 + 7 + 6).
void build_bands(const OrbGeom& g, const std::vector<OrbCell>& cells, std::vector<OrbBand>& bands,
                 std::vector<unsigned>& band_bm, std::vector<uint2>& cell_bm, unsigned* bm_rows_frame) {
  unsigned bm_rows = 0;
  cell_bm.assign(cells.size(), make_uint2(0u, 0u));
  for (int l = 0; l < g.nlevels; ++l) {
    const OrbLevelGeom& L = g.lv[l];
    int p = L.cell_base;
    const int end = L.cell_base + L.n_cells;
    while (p < end) {
      int q = p;
      while (q < end && cells[q].ini_y == cells[p].ini_y) ++q;  // one row of cells
      const int n = q - p;
      const int per = std::max(1, std::min(8, ORB_BAND_MAX_PX / L.w_cell));
      const int nchunk = (n + per - 1) / per;
      int c0 = p;
      for (int k = 0; k < nchunk; ++k) {
        const int nc = n / nchunk + (k < n % nchunk ? 1 : 0);
        const OrbCell& first = cells[c0];
        const OrbCell& last = cells[c0 + nc - 1];
        OrbBand b;
        std::memset(&b, 0, sizeof(b));
        b.level = (short)l;
        b.n_cells = (short)nc;
        b.cell_idx0 = c0;
        b.cand_slot_off = first.cand_slot_off;
        b.cand_cap = first.cand_cap;
        b.pitch = (unsigned short)L.pitch;
        b.y_first = first.ini_y;
        b.nt = (short)(first.ch - 6);
        b.x0 = (short)(first.ini_x + 3);
        b.x1 = (short)(last.ini_x + last.cw - 3);
        b.xb = (short)((b.x0 - 4) & ~3);
        b.w_cell = (short)L.w_cell;
        b.src_off = L.pyr_off + (unsigned)b.y_first * (unsigned)L.pitch + (unsigned)b.xb;
        bands.push_back(b);
        band_bm.push_back(bm_rows);
        for (int ci = c0; ci < c0 + nc; ++ci) cell_bm[ci] = make_uint2(bm_rows, (unsigned)(cells[ci].ini_x + 3 - b.xb));
        bm_rows += (unsigned)std::max(0, (int)b.nt);
        c0 += nc;
      }
      p = q;
    }
  }
  *bm_rows_frame = bm_rows;
}

template <typename T> bool dev_alloc(orbx_extractor* h, T** p, size_t count, const char* what) {
  return h->check(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)), what);
}

void drop_graphs(orbx_extractor* h) {
  for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

constexpr int kGraphMaxFrames = 8;             // batches up to this size replay a captured graph
constexpr size_t kSmallStageBytes = 512 << 10;  // staging blocks up to this size come back in one pinned copy

bool ensure_staging(orbx_extractor* h, int cap) {
  if (h->d_stage && cap <= h->stage_cap) return true;
  drop_graphs(h);  // they hold the old staging pointers
  cudaFree(h->d_stage);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  h->d_stage = nullptr;
  h->h_stage = nullptr;
  h->d_kps = nullptr;
  h->d_desc = nullptr;
  const size_t n = (size_t)h->cfg.max_batch * cap;
  const size_t off_desc = align_up(n * sizeof(orbx_keypoint), 256), off_counts = off_desc + align_up(n * 32, 256);
  h->stage_bytes = off_counts + align_up(sizeof(int32_t) * (size_t)h->cfg.max_batch, 256);
  if (!dev_alloc(h, &h->d_stage, h->stage_bytes, "cudaMalloc(result staging)")) return false;
  h->d_kps = reinterpret_cast<orbx_keypoint*>(h->d_stage);
  h->d_desc = h->d_stage + off_desc;
  h->d_stage_counts = reinterpret_cast<int*>(h->d_stage + off_counts);
  if (h->stage_bytes <= kSmallStageBytes &&
      !h->check(cudaHostAlloc((void**)&h->h_stage, h->stage_bytes, cudaHostAllocDefault), "cudaHostAlloc(result staging)"))
    return false;
  h->stage_cap = cap;
  return true;
}

// One batch (<= max_batch frames) of the full pipeline on the handle's stream.
int run_batch(orbx_extractor* h, const uint8_t* d_images, int n_frames, size_t frame_stride, size_t row_stride,
              orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_counts, int cap) {
  cudaStream_t st = h->stream;
  const int W = h->cfg.width, H = h->cfg.height;
  // level 0 is read in place when the frames are 4-byte aligned; otherwise through an aligned copy
  orbk::OrbLevel0 l0 = {d_images, frame_stride, (int)row_stride};
  if ((reinterpret_cast<uintptr_t>(d_images) & 3) || (row_stride & 3) || (frame_stride & 3)) {
    if (!h->d_l0 && !h->check(cudaMalloc((void**)&h->d_l0, (size_t)h->cfg.max_batch * h->img_pitch * H), "cudaMalloc(aligned level 0)"))
      return ORBX_E_CUDA;
    for (int f = 0; f < n_frames; ++f)
      if (!h->check(cudaMemcpy2DAsync(h->d_l0 + (size_t)f * h->img_pitch * H, h->img_pitch, d_images + (size_t)f * frame_stride,
                                      row_stride, W, H, cudaMemcpyDeviceToDevice, st), "align level 0"))
        return ORBX_E_CUDA;
    l0 = {h->d_l0, (size_t)h->img_pitch * H, h->img_pitch};
  }
  const bool prof = h->profiling;
  cudaEvent_t* ev = h->ev[h->ev_next];
  if (prof) {
    if (!h->harvest(h->ev_next)) return ORBX_E_CUDA;
    cudaEventRecord(ev[0], st);
  }
  orbk::launch_pyramid(h->gh, l0, n_frames, h->d_pyr, st, &h->launches);
  if (prof) cudaEventRecord(ev[1], st);
  const bool fork = h->fork_blur && !prof && h->side;
  if (fork) {
    cudaEventRecord(h->ev_fork, st);
    cudaStreamWaitEvent(h->side, h->ev_fork, 0);
    orbk::launch_blur(h->gh, l0, n_frames, h->d_pyr, h->d_blur, h->side, &h->launches);
    cudaEventRecord(h->ev_join, h->side);
  }
  orbk::launch_fast(h->gh, l0, n_frames, h->d_pyr, h->d_cand, h->d_cell_count, st, &h->launches);
  if (prof) cudaEventRecord(ev[2], st);
  orbk::launch_octree(h->gh, n_frames, h->d_cand, h->d_cell_count, h->d_keys, h->d_knode, h->d_sel, h->d_sel_count, st,
                      &h->launches);
  if (prof) cudaEventRecord(ev[3], st);
  if (fork) cudaStreamWaitEvent(st, h->ev_join, 0);
  else orbk::launch_blur(h->gh, l0, n_frames, h->d_pyr, h->d_blur, st, &h->launches);
  if (prof) cudaEventRecord(ev[4], st);
  orbk::launch_orient_describe(h->gh, l0, n_frames, h->d_pyr, h->d_blur, h->d_sel, h->d_sel_count, d_kps, d_desc, d_counts,
                               cap, st, &h->launches);
  if (prof) {
    cudaEventRecord(ev[5], st);
    h->ev_pending[h->ev_next] = true;
    h->ev_next = (h->ev_next + 1) % orbx_extractor::kRing;
  }
  h->last_batch = n_frames;
  h->last_l0 = l0;
  if (!h->check(cudaGetLastError(), "kernel launch")) return ORBX_E_CUDA;
  return ORBX_OK;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int orbx_create(const orbx_config* cfg, orbx_extractor** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return ORBX_E_INVALID; }
  *out = nullptr;
  if (cfg->nfeatures < 1 || cfg->nlevels < 1 || cfg->nlevels > ORBX_MAX_LEVELS || !(cfg->scale_factor > 1.0f) ||
      cfg->width < 1 || cfg->height < 1 || cfg->max_batch < 1) {
    g_create_error = "invalid configuration";
    return ORBX_E_INVALID;
  }
  orbx_extractor* h = new orbx_extractor();
  h->cfg = *cfg;
  auto fail = [&](int code) {
    g_create_error = h->err;
    orbx_destroy(h);
    return code;
  };
  build_tables(h);
  std::vector<OrbCell> cells;
  std::vector<OrbXTap> xt;
  std::vector<OrbYTap> yt;
  if (!build_geometry(h, cells, xt, yt)) return fail(ORBX_E_INVALID);
  std::vector<OrbBand> bands;
  std::vector<unsigned> band_bm;
  std::vector<uint2> cell_bm;
  build_bands(h->gh.g, cells, bands, band_bm, cell_bm, &h->gh.bm_rows_frame);
  h->gh.n_bands = (int)bands.size();
  {
    const char* e = getenv("ORB_B200_FAST");
    const std::string mode = e ? e : "";
    h->gh.fast_mode = mode == "bands" ? 1 : (mode == "split" ? 2 : 0);
  }
  int ndev = 0;
  if (!h->check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev == 0) {
    if (h->err.empty()) h->err = "no CUDA device (this library has no CPU fallback)";
    return fail(ORBX_E_CUDA);
  }
  if (cfg->device >= ndev) { h->err = "no such CUDA device"; return fail(ORBX_E_CUDA); }
  if (!h->check(cudaGetDevice(&h->device), "cudaGetDevice")) return fail(ORBX_E_CUDA);
  if (cfg->device >= 0) h->device = cfg->device;
  OrbDeviceGuard dev_guard(h->device);  // the caller's current device is restored on return
  if (!h->check(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
      !h->check(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking), "cudaStreamCreate") ||
      !h->check(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming), "cudaEventCreate") ||
      !h->check(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming), "cudaEventCreate"))
    return fail(ORBX_E_CUDA);
  h->stream = h->own_stream;
  for (auto& set : h->ev)
    for (auto& e : set)
      if (!h->check(cudaEventCreate(&e), "cudaEventCreate")) return fail(ORBX_E_CUDA);
  const OrbGeom& g = h->gh.g;
  const size_t B = (size_t)cfg->max_batch;
  h->img_pitch = (int)align_up(cfg->width, 16);
  bool ok = dev_alloc(h, &h->gh.d_geom, 1, "cudaMalloc(geom)") &&
            dev_alloc(h, &h->gh.d_cells, cells.size(), "cudaMalloc(cells)") &&
            dev_alloc(h, &h->gh.d_bands, bands.size(), "cudaMalloc(bands)") &&
            dev_alloc(h, &h->gh.d_band_bm, band_bm.size(), "cudaMalloc(band bitmap rows)") &&
            dev_alloc(h, &h->gh.d_cell_bm, cell_bm.size(), "cudaMalloc(cell bitmap rows)") &&
            dev_alloc(h, &h->gh.d_bitmap, h->gh.fast_mode == 2 ? B * h->gh.bm_rows_frame * 32 : 1, "cudaMalloc(FAST bitmap)") &&
            dev_alloc(h, &h->gh.d_xtab, xt.size(), "cudaMalloc(xtab)") &&
            dev_alloc(h, &h->gh.d_ytab, yt.size(), "cudaMalloc(ytab)") &&
            dev_alloc(h, &h->d_pyr, B * g.pyr_frame_bytes, "cudaMalloc(pyramid)") &&
            dev_alloc(h, &h->d_blur, B * g.blur_frame_bytes, "cudaMalloc(blur)") &&
            dev_alloc(h, &h->d_cand, B * g.cand_frame_u32, "cudaMalloc(candidates)") &&
            dev_alloc(h, &h->d_cell_count, B * g.n_cells, "cudaMalloc(cell counts)") &&
            dev_alloc(h, &h->d_keys, B * g.key_frame_u32, "cudaMalloc(keys)") &&
            dev_alloc(h, &h->d_knode, B * g.key_frame_u32, "cudaMalloc(key nodes)") &&
            dev_alloc(h, &h->d_sel, B * g.kp_cap_frame, "cudaMalloc(selection)") &&
            dev_alloc(h, &h->d_sel_count, B * g.nlevels, "cudaMalloc(selection counts)") &&
            dev_alloc(h, &h->d_img, B * (size_t)h->img_pitch * cfg->height, "cudaMalloc(image staging)");
  if (!ok) return fail(ORBX_E_CUDA);
  ok = h->check(cudaMemcpy(h->gh.d_geom, &g, sizeof(g), cudaMemcpyHostToDevice), "copy geom") &&
       h->check(cudaMemcpy(h->gh.d_cells, cells.data(), cells.size() * sizeof(OrbCell), cudaMemcpyHostToDevice), "copy cells") &&
       h->check(cudaMemcpy(h->gh.d_bands, bands.data(), bands.size() * sizeof(OrbBand), cudaMemcpyHostToDevice), "copy bands") &&
       h->check(cudaMemcpy(h->gh.d_band_bm, band_bm.data(), band_bm.size() * sizeof(unsigned), cudaMemcpyHostToDevice), "copy band rows") &&
       h->check(cudaMemcpy(h->gh.d_cell_bm, cell_bm.data(), cell_bm.size() * sizeof(uint2), cudaMemcpyHostToDevice), "copy cell rows") &&
       h->check(cudaMemcpy(h->gh.d_xtab, xt.data(), xt.size() * sizeof(OrbXTap), cudaMemcpyHostToDevice), "copy xtab") &&
       h->check(cudaMemcpy(h->gh.d_ytab, yt.data(), yt.size() * sizeof(OrbYTap), cudaMemcpyHostToDevice), "copy ytab") &&
       h->check(orbk::prepare_octree(h->gh), "octree shared-memory opt-in") &&
       h->check(orbk::prepare_pyramid(g), "pyramid shared-memory opt-in");
  if (!ok) return fail(ORBX_E_CUDA);
  *out = h;
  return ORBX_OK;
}

void orbx_destroy(orbx_extractor* h) {
  if (!h) return;
#ifdef ORB_OT_TIMING
  orbk::dump_octree_marks();
#endif
  if (h->stream) cudaStreamSynchronize(h->stream);
  drop_graphs(h);
  cudaFree(h->d_stage);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  cudaFree(h->gh.d_geom); cudaFree(h->gh.d_cells); cudaFree(h->gh.d_bands); cudaFree(h->gh.d_band_bm); cudaFree(h->gh.d_cell_bm); cudaFree(h->gh.d_bitmap); cudaFree(h->gh.d_xtab); cudaFree(h->gh.d_ytab);
  cudaFree(h->d_pyr); cudaFree(h->d_blur); cudaFree(h->d_cand); cudaFree(h->d_cell_count);
  cudaFree(h->d_keys); cudaFree(h->d_knode); cudaFree(h->d_sel); cudaFree(h->d_sel_count);
  cudaFree(h->d_img); cudaFree(h->d_l0);
  for (auto& set : h->ev)
    for (auto& e : set)
      if (e) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
}

const char* orbx_last_error(const orbx_extractor* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int orbx_get_scale_tables(const orbx_extractor* h, float* s, float* is, float* s2, float* is2) {
  if (!h) return ORBX_E_INVALID;
  for (int i = 0; i < h->cfg.nlevels; ++i) {
    if (s) s[i] = h->scale[i];
    if (is) is[i] = h->inv_scale[i];
    if (s2) s2[i] = h->sigma2[i];
    if (is2) is2[i] = h->inv_sigma2[i];
  }
  return ORBX_OK;
}

int orbx_get_features_per_level(const orbx_extractor* h, int32_t* out) {
  if (!h || !out) return ORBX_E_INVALID;
  for (int i = 0; i < h->cfg.nlevels; ++i) out[i] = h->quota[i];
  return ORBX_OK;
}

int orbx_max_keypoints(const orbx_extractor* h) { return h ? h->cfg.nfeatures + 3 * h->cfg.nlevels : ORBX_E_INVALID; }

int orbx_extract_batch_device(orbx_extractor* h, const uint8_t* d_images, int n_frames, size_t frame_stride,
                              size_t row_stride, orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_counts, int cap) {
  if (!h) return ORBX_E_INVALID;
  if (!d_images || !d_kps || !d_desc || !d_counts || n_frames < 0 || cap < 1 || row_stride < (size_t)h->cfg.width) {
    h->err = "invalid argument";
    return ORBX_E_INVALID;
  }
  OrbDeviceGuard dev_guard(h->device);
  for (int f0 = 0; f0 < n_frames; f0 += h->cfg.max_batch) {
    const int nb = std::min(h->cfg.max_batch, n_frames - f0);
    const int rc = run_batch(h, d_images + (size_t)f0 * frame_stride, nb, frame_stride, row_stride,
                             d_kps + (size_t)f0 * cap, d_desc + (size_t)f0 * cap * 32, d_counts + f0, cap);
    if (rc != ORBX_OK) return rc;
  }
  return ORBX_OK;
}

int orbx_extract_batch_host(orbx_extractor* h, const uint8_t* images, int n_frames, size_t frame_stride,
                            size_t row_stride, orbx_keypoint* kps, uint8_t* desc, int32_t* counts, int cap) {
  if (!h) return ORBX_E_INVALID;
  if (!images || !kps || !desc || !counts || n_frames < 0 || cap < 1 || row_stride < (size_t)h->cfg.width) {
    h->err = "invalid argument";
    return ORBX_E_INVALID;
  }
  OrbDeviceGuard dev_guard(h->device);
  if (!ensure_staging(h, cap)) return ORBX_E_CUDA;
  const int W = h->cfg.width, H = h->cfg.height;
  int status = ORBX_OK;
  for (int f0 = 0; f0 < n_frames; f0 += h->cfg.max_batch) {
    const int nb = std::min(h->cfg.max_batch, n_frames - f0);
    const uint8_t* src = images + (size_t)f0 * frame_stride;
    // H2D into the staging buffer (row pitch rounded up to 16: level 0 is then read in place);
    // one 2D copy per batch when the frames are contiguous rows
    const int P = h->img_pitch;
    if (frame_stride == row_stride * (size_t)H) {
      if (!h->check(cudaMemcpy2DAsync(h->d_img, P, src, row_stride, W, (size_t)H * nb, cudaMemcpyHostToDevice, h->stream),
                    "H2D images"))
        return ORBX_E_CUDA;
    } else {
      for (int f = 0; f < nb; ++f)
        if (!h->check(cudaMemcpy2DAsync(h->d_img + (size_t)f * P * H, P, src + (size_t)f * frame_stride, row_stride, W, H,
                                        cudaMemcpyHostToDevice, h->stream), "H2D image"))
          return ORBX_E_CUDA;
    }
    int rc;
    if (nb <= kGraphMaxFrames && !h->profiling) {
      orbx_extractor::BatchGraph* bg = nullptr;
      for (auto& g : h->graphs)
        if (g.nb == nb && g.cap == cap) bg = &g;
      if (!bg) {
        const long long before = h->launches;
        if (!h->check(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed), "graph capture")) return ORBX_E_CUDA;
        h->fork_blur = true;
        rc = run_batch(h, h->d_img, nb, (size_t)P * H, P, h->d_kps, h->d_desc, h->d_stage_counts, cap);
        h->fork_blur = false;
        cudaGraph_t graph = nullptr;
        const bool ended = h->check(cudaStreamEndCapture(h->stream, &graph), "graph capture end");
        if (rc != ORBX_OK || !ended) {
          if (graph) cudaGraphDestroy(graph);
          return rc != ORBX_OK ? rc : ORBX_E_CUDA;
        }
        orbx_extractor::BatchGraph g = {nb, cap, nullptr, h->launches - before};
        const bool inst = h->check(cudaGraphInstantiate(&g.exec, graph, 0), "graph instantiate");
        cudaGraphDestroy(graph);
        if (!inst) return ORBX_E_CUDA;
        h->launches = before;
        h->graphs.push_back(g);
        bg = &h->graphs.back();
      }
      if (!h->check(cudaGraphLaunch(bg->exec, h->stream), "graph launch")) return ORBX_E_CUDA;
      h->launches += bg->launches;
      h->last_batch = nb;
      h->last_l0 = {h->d_img, (size_t)P * H, P};
      rc = ORBX_OK;
    } else {
      rc = run_batch(h, h->d_img, nb, (size_t)P * H, P, h->d_kps, h->d_desc, h->d_stage_counts, cap);
    }
    if (rc != ORBX_OK) return rc;
    if (h->h_stage && nb == h->cfg.max_batch && cap == h->stage_cap) {
      // the whole staging block in one pinned copy, then out of the mirror
      if (!h->check(cudaMemcpyAsync(h->h_stage, h->d_stage, h->stage_bytes, cudaMemcpyDeviceToHost, h->stream), "D2H results") ||
          !h->check(cudaStreamSynchronize(h->stream), "extract batch"))
        return ORBX_E_CUDA;
      const size_t n = (size_t)nb * cap;
      std::memcpy(counts + f0, h->h_stage + ((uint8_t*)h->d_stage_counts - h->d_stage), sizeof(int32_t) * nb);
      std::memcpy(kps + (size_t)f0 * cap, h->h_stage, sizeof(orbx_keypoint) * n);
      std::memcpy(desc + (size_t)f0 * cap * 32, h->h_stage + (h->d_desc - h->d_stage), n * 32);
    } else if (!h->check(cudaMemcpyAsync(counts + f0, h->d_stage_counts, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, h->stream), "D2H counts") ||
               !h->check(cudaMemcpyAsync(kps + (size_t)f0 * cap, h->d_kps, sizeof(orbx_keypoint) * (size_t)nb * cap, cudaMemcpyDeviceToHost, h->stream), "D2H keypoints") ||
               !h->check(cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, h->d_desc, (size_t)nb * cap * 32, cudaMemcpyDeviceToHost, h->stream), "D2H descriptors") ||
               !h->check(cudaStreamSynchronize(h->stream), "extract batch"))
      return ORBX_E_CUDA;
    for (int f = 0; f < nb; ++f)
      if (counts[f0 + f] > cap) {
        h->err = "keypoint capacity too small: frame " + std::to_string(f0 + f) + " has " + std::to_string(counts[f0 + f]);
        status = ORBX_E_CAPACITY;
      }
  }
  return status;
}

int orbx_extract(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t stride, orbx_keypoint* kps,
                 uint8_t* desc, int cap, int* n) {
  if (!h) return ORBX_E_INVALID;
  if (n) *n = 0;
  if (!image || rows == 0 || cols == 0) return ORBX_OK;  // empty image: silent return (:1047-1048)
  if (rows != h->cfg.height || cols != h->cfg.width) {
    h->err = "image size differs from the size this handle was created for";
    return ORBX_E_INVALID;
  }
  int32_t count = 0;
  const int rc = orbx_extract_batch_host(h, image, 1, stride * (size_t)rows, stride, kps, desc, &count, cap);
  if (n) *n = std::min(count, cap);
  return rc;
}

int orbx_sync(orbx_extractor* h) {
  if (!h) return ORBX_E_INVALID;
  return h->check(cudaStreamSynchronize(h->stream), "stream synchronize") ? ORBX_OK : ORBX_E_CUDA;
}

void* orbx_stream(orbx_extractor* h) { return h ? (void*)h->stream : nullptr; }

int orbx_set_stream(orbx_extractor* h, void* cuda_stream) {
  if (!h) return ORBX_E_INVALID;
  if (!h->check(cudaStreamSynchronize(h->stream), "stream synchronize")) return ORBX_E_CUDA;
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return ORBX_OK;
}

int orbx_get_pyramid_view(orbx_extractor* h, orbx_pyramid_view* out) {
  if (!h || !out) return ORBX_E_INVALID;
  if (!h->last_l0.base || h->last_batch < 1) { h->err = "no batch has been extracted yet"; return ORBX_E_STATE; }
  std::memset(out, 0, sizeof(*out));
  out->nlevels = h->cfg.nlevels;
  out->n_frames = h->last_batch;
  out->stream = (void*)h->stream;
  for (int l = 0; l < h->cfg.nlevels; ++l) {
    const OrbLevelGeom& L = h->gh.g.lv[l];
    out->w[l] = L.w;
    out->h[l] = L.h;
    out->scale[l] = h->scale[l];
    out->inv_scale[l] = h->inv_scale[l];
    if (l == 0) {
      out->base[0] = h->last_l0.base;
      out->pitch[0] = h->last_l0.pitch;
      out->frame_stride[0] = h->last_l0.frame_stride;
    } else {
      out->base[l] = h->d_pyr + L.pyr_off;
      out->pitch[l] = L.pitch;
      out->frame_stride[l] = (size_t)h->gh.g.pyr_frame_bytes;
    }
  }
  return ORBX_OK;
}

int orbx_get_pyramid_level(orbx_extractor* h, int frame, int level, int with_border, uint8_t* dst, size_t dst_stride,
                           int* w, int* hgt) {
  if (!h || level < 0 || level >= h->cfg.nlevels) return ORBX_E_INVALID;
  const OrbLevelGeom& L = h->gh.g.lv[level];
  const int B = with_border ? ORB_EDGE : 0;
  const int ow = L.w + 2 * B, oh = L.h + 2 * B;
  if (w) *w = ow;
  if (hgt) *hgt = oh;
  if (!dst) return ORBX_OK;
  if (frame < 0 || frame >= h->last_batch) { h->err = "no such frame in the last batch"; return ORBX_E_STATE; }
  if (dst_stride < (size_t)ow) return ORBX_E_INVALID;
  OrbDeviceGuard dev_guard(h->device);
  // The levels are stored without the EDGE_THRESHOLD border (level 0 is the caller's frame): copy the
  // interior into place, then mirror it outwards exactly as copyMakeBorder(BORDER_REFLECT_101) does
  // (src/ORBextractor.cc:1124-1130).  This accessor is not on the hot path.
  const uint8_t* src;
  size_t spitch;
  if (level == 0) {
    if (!h->last_l0.base) return ORBX_E_STATE;
    src = h->last_l0.base + (size_t)frame * h->last_l0.frame_stride;
    spitch = (size_t)h->last_l0.pitch;
  } else {
    src = h->d_pyr + (size_t)frame * h->gh.g.pyr_frame_bytes + L.pyr_off;
    spitch = (size_t)L.pitch;
  }
  uint8_t* interior = dst + (size_t)B * dst_stride + B;
  if (!h->check(cudaStreamSynchronize(h->stream), "sync") ||
      !h->check(cudaMemcpy2D(interior, dst_stride, src, spitch, L.w, L.h, cudaMemcpyDeviceToHost), "D2H pyramid level"))
    return ORBX_E_CUDA;
  if (B) {
    auto refl = [](int p, int n) { p = p < 0 ? -p : p; return p >= n ? 2 * (n - 1) - p : p; };
    for (int y = 0; y < L.h; ++y) {
      uint8_t* row = interior + (size_t)y * dst_stride;
      for (int k = 1; k <= B; ++k) {
        row[-k] = row[refl(-k, L.w)];
        row[L.w - 1 + k] = row[refl(L.w - 1 + k, L.w)];
      }
    }
    for (int k = 1; k <= B; ++k) {
      std::memcpy(dst + (size_t)(B - k) * dst_stride, dst + (size_t)(B + refl(-k, L.h)) * dst_stride, ow);
      std::memcpy(dst + (size_t)(B + L.h - 1 + k) * dst_stride, dst + (size_t)(B + refl(L.h - 1 + k, L.h)) * dst_stride, ow);
    }
  }
  return ORBX_OK;
}

int orbx_debug_candidates(orbx_extractor* h, int frame, int level, int32_t* x, int32_t* y, int32_t* score, int cap,
                          int* n) {
  if (!h || level < 0 || level >= h->cfg.nlevels || !n) return ORBX_E_INVALID;
  if (frame < 0 || frame >= h->last_batch) return ORBX_E_STATE;
  OrbDeviceGuard dev_guard(h->device);
  const OrbGeom& g = h->gh.g;
  const OrbLevelGeom& L = g.lv[level];
  std::vector<int> cc(L.n_cells);
  std::vector<uint32_t> slots((size_t)L.key_cap);
  if (!h->check(cudaStreamSynchronize(h->stream), "sync") ||
      !h->check(cudaMemcpy(cc.data(), h->d_cell_count + (size_t)frame * g.n_cells + L.cell_base, sizeof(int) * L.n_cells, cudaMemcpyDeviceToHost), "D2H cell counts") ||
      !h->check(cudaMemcpy(slots.data(), h->d_cand + (size_t)frame * g.cand_frame_u32 + L.cand_off, sizeof(uint32_t) * slots.size(), cudaMemcpyDeviceToHost), "D2H candidates"))
    return ORBX_E_CUDA;
  int k = 0;
  for (int c = 0; c < L.n_cells; ++c)
    for (int i = 0; i < cc[c]; ++i, ++k) {
      if (k < cap) {
        const uint32_t key = slots[(size_t)c * L.cand_cap + i];
        x[k] = OT_KEY_X(key); y[k] = OT_KEY_Y(key); score[k] = OT_KEY_SCORE(key);
      }
    }
  *n = k;
  return k <= cap ? ORBX_OK : ORBX_E_CAPACITY;
}

int orbx_debug_blurred(orbx_extractor* h, int frame, int level, uint8_t* dst, size_t dst_stride) {
  if (!h || level < 0 || level >= h->cfg.nlevels || !dst) return ORBX_E_INVALID;
  if (frame < 0 || frame >= h->last_batch) return ORBX_E_STATE;
  OrbDeviceGuard dev_guard(h->device);
  const OrbLevelGeom& L = h->gh.g.lv[level];
  if (!h->check(cudaStreamSynchronize(h->stream), "sync") ||
      !h->check(cudaMemcpy2D(dst, dst_stride, h->d_blur + (size_t)frame * h->gh.g.blur_frame_bytes + L.blur_off, L.bpitch,
                             L.w, L.h, cudaMemcpyDeviceToHost), "D2H blurred level"))
    return ORBX_E_CUDA;
  return ORBX_OK;
}

long long orbx_launch_count(const orbx_extractor* h) { return h ? h->launches : 0; }

int orbx_set_profiling(orbx_extractor* h, int enable) {
  if (!h) return ORBX_E_INVALID;
  if (!h->check(cudaStreamSynchronize(h->stream), "stream synchronize")) return ORBX_E_CUDA;
  h->profiling = enable != 0;
  for (int i = 0; i < orbx_extractor::kRing; ++i) h->ev_pending[i] = false;
  for (double& v : h->stage_sum_ms) v = 0;
  h->stage_calls = 0;
  return ORBX_OK;
}

int orbx_stage_times_ms(orbx_extractor* h, double* sum_ms5, long long* n_calls) {
  if (!h || !sum_ms5 || !n_calls) return ORBX_E_INVALID;
  for (int i = 0; i < orbx_extractor::kRing; ++i)
    if (!h->harvest(i)) return ORBX_E_CUDA;
  for (int i = 0; i < 5; ++i) sum_ms5[i] = h->stage_sum_ms[i];
  *n_calls = h->stage_calls;
  return ORBX_OK;
}

#pragma GCC visibility pop
}  // extern "C"
