// Host-side launch interface of the extractor kernels (extract_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_geom.h"

namespace orbk {

// Level 0 of the pyramid is the caller's frame itself, read in place (mvImagePyramid[0] is the
// input image; the reference only copies it to add a border nobody inside the extractor reads).
struct OrbLevel0 {
  const uint8_t* base;   // frame 0, row 0 (4-byte aligned)
  size_t frame_stride;   // bytes between frames
  int pitch;             // bytes per row (multiple of 4)
};

struct OrbGeomHost {
  OrbGeom g;               // host copy
  OrbGeom* d_geom;         // device copy
  OrbCell* d_cells;        // [g.n_cells]
  OrbBand* d_bands;        // [n_bands] work list of the FAST kernel
  int n_bands;
  OrbXTap* d_xtab;
  OrbYTap* d_ytab;
  int ot_smem_keys;        // octree: keys kept in shared memory (prepare_octree)
  size_t ot_smem_bytes;    // octree: dynamic shared memory per CTA
  int fast_mode;           // FAST stage of this handle, resolved once in orbx_create (ORB_B200_FAST=cells|bands|split):
                           // 0 = one warp per cell (k_fast_cells), 1 = one warp per band (k_fast_bands), 2 = split:
                           // k_fast_prefilter (band rejection test on registers -> pass-bit bitmap) + k_fast_cells fed by it
  unsigned* d_band_bm;     // [n_bands] first bitmap row of every band (split mode)
  uint2* d_cell_bm;        // [g.n_cells] (first bitmap row of the cell's band, bit of the cell's first tested column)
  uint8_t* d_bitmap;       // [max_batch][bm_rows_frame][32] pass bits of the rejection test, one row per tested band row
  unsigned bm_rows_frame;
};

void launch_pyramid(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, uint8_t* d_pyr, cudaStream_t st,
                    long long* launches);
void launch_fast(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, const uint8_t* d_pyr, uint32_t* d_cand,
                 int* d_cell_count, cudaStream_t st, long long* launches);
cudaError_t prepare_octree(OrbGeomHost& gh);  // resolves gh.ot_smem_* and raises the kernel's opt-in limit
cudaError_t prepare_pyramid(const OrbGeom& g);
void launch_octree(const OrbGeomHost& gh, int n_frames, const uint32_t* d_cand, const int* d_cell_count,
                   uint32_t* d_keys, uint16_t* d_knode, uint32_t* d_sel, int* d_sel_count, cudaStream_t st,
                   long long* launches);
void launch_blur(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, const uint8_t* d_pyr, uint8_t* d_blur,
                 cudaStream_t st, long long* launches);
void launch_orient_describe(const OrbGeomHost& gh, OrbLevel0 l0, int n_frames, const uint8_t* d_pyr, const uint8_t* d_blur,
                            const uint32_t* d_sel, const int* d_sel_count, orbx_keypoint* d_kps, uint8_t* d_desc,
                            int* d_counts, int cap, cudaStream_t st, long long* launches);

}  // namespace orbk
