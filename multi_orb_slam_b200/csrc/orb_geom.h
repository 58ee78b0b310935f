// Device-visible geometry of one extractor handle: pyramid level layout in HBM, FAST cell grid,
// octree roots, workspace offsets.  Built once on the host (api.cu: build_geometry) with the
// reference's own float arithmetic (src/ORBextractor.cc:411-471, 766-788, 1109-1117), stored in
// device memory and read by every kernel.
#pragma once
#include <stdint.h>

#include "octree_core.h"

#define ORB_MAX_LEVELS 16
#define ORB_EDGE 19        // EDGE_THRESHOLD (src/ORBextractor.cc:72)
#define ORB_HALF_PATCH 15  // HALF_PATCH_SIZE (:71)
#define ORB_MIN_BORDER 16  // EDGE_THRESHOLD - 3 (:774)
#define ORB_CELL_MAX 72    // max FAST sub-image side (cell + 6) the cell kernel stages

struct OrbLevelGeom {
  int w, h;                 // interior (mvImagePyramid[l]) size
  int pitch;                // bytes per row of the bordered buffer ((w+38) rounded up to 16)
  int bpitch;               // bytes per row of the blurred image (w rounded up to 16)
  unsigned pyr_off;         // byte offset of the bordered buffer inside a frame's pyramid block
  unsigned blur_off;        // byte offset of the blurred image inside a frame's blur block
  // FAST cell grid (:782-788)
  int n_cols, n_rows, w_cell, h_cell;
  int cell_base;            // index of this level's first cell in the frame's cell list
  int n_cells;              // live cells (rows/cols skipped by :795,804 are not listed)
  int cand_cap;             // candidate slots per cell
  unsigned cand_off;        // u32 offset of this level's slots inside a frame's candidate block
  unsigned key_off;         // u32 offset of this level's compacted keys inside a frame's key block
  int key_cap;              // n_cells * cand_cap
  int quota;                // mnFeaturesPerLevel[l]
  int sel_off;              // offset of this level's selected keys inside a frame's selection block
  int sel_cap;              // max(quota + 3, 4 * n_ini) + 1
  int blur_tile_base, blur_tiles_x, blur_tiles_y;
  int xtab_off, ytab_off;   // resize coefficient tables (level l <- l-1), entries in OrbGeom tables
  float scale;              // mvScaleFactor[l]
  float patch_size;         // (float)(int)(31 * mvScaleFactor[l])
  OtRoots roots;
};

struct __align__(16) OrbCell {  // one live FAST cell (:790-830); self-contained so the cell kernel
                                // needs ONE 32-byte load before it can fetch its tile
  unsigned tile_off;        // byte offset, inside a frame's pyramid block, of the word-aligned tile origin
  unsigned cand_slot_off;   // u32 offset, inside a frame's candidate block, of this cell's slots
  unsigned short pitch;     // bytes per row of the level's bordered buffer
  unsigned short cand_cap;  // candidate slots of this cell
  short cw, ch;             // sub-image size (cell + 6, clipped)
  short off_x, off_y;       // j*wCell, i*hCell added to cell-local coordinates (:823-824)
  short a0;                 // misalignment of the sub-image's first column inside its first word
  short level;
  short ini_x, ini_y;       // top-left of the sub-image in level coordinates
  short slot;               // cell index within its level
  short pad;
};

// One unit of work of the FAST kernel: a horizontal band = one row of FAST cells of one level, cut into chunks of
// whole cells no wider than the 30 lanes x 8 px a warp tests per image row.  The tested areas of the reference's
// cells (:790-830: sub-image minus its 3-px border) tile [19, w-19) x [19, h-19) of the level, so a band's tested
// area is the rectangle [x0, x1) x [y_first + 3, y_first + 3 + nt).
struct __align__(16) OrbBand {
  unsigned src_off;         // levels >= 1: byte offset, inside a frame's pyramid block, of (row y_first, column xb)
  unsigned cand_slot_off;   // u32 offset, inside a frame's candidate block, of the first cell's slots
  int cell_idx0;            // index of the first cell in the frame's cell list
  unsigned short pitch;     // bytes per row of the level (levels >= 1)
  unsigned short cand_cap;  // candidate slots per cell
  short level, n_cells;     // n_cells <= 8
  short y_first;            // level row of the first row the band loads (first tested row - 3)
  short nt;                 // tested rows (<= 0: the cells exist but their tested area is empty)
  short xb;                 // level column of byte 0 of lane 0 (multiple of 4); lanes 1..30 hold the tested pixels
  short x0, x1;             // tested columns [x0, x1); cell c of the band tests [x0 + c*w_cell, min(x0 + (c+1)*w_cell, x1))
  short w_cell;
};
#define ORB_BAND_MAX_PX 245  // tested pixels per band row: 256 bytes - (up to 7 bytes of alignment) - one word of halo

struct OrbGeom {
  int nlevels;
  int width, height;
  int n_cells;              // live cells over all levels
  int n_blur_tiles;
  int ini_th, min_th;
  int ot_cap;               // octree node capacity (max sel_cap over levels)
  int ot_scan_cap;          // max(ot_cap, max n_cells per level) + 1
  int kp_cap_frame;         // sum of sel_cap: selection block size per frame
  unsigned long long pyr_frame_bytes, blur_frame_bytes;
  unsigned long long cand_frame_u32, key_frame_u32;
  OrbLevelGeom lv[ORB_MAX_LEVELS];
};

// cv::resize INTER_LINEAR coefficient table entries, computed on the host exactly as OpenCV
// does (double scale, float fraction, cvRound to 11-bit fixed point).
struct OrbXTap { unsigned short sx, a0, a1, pad; };     // dst x -> src x, weights (sum ~2048)
struct OrbYTap { unsigned short sy0, sy1, b0, b1; };    // dst y -> src rows (clipped), weights

#define ORB_BLUR_TW 128
#define ORB_BLUR_TH 36  // + 6 halo rows = 42 = 6 blocks of 7 (the blur's register window period)
