// Host build of the product's octree policy core (multi_orb_slam_b200/csrc/octree_core.h) under
// its sequential thread emulation, so tests/test_octree_host.py can check the parallel
// reformulation against the oracle without a GPU.  This is a TEST of product code, not a
// fallback: nothing in the package loads it.
#include <cmath>
#include <cstdlib>
#include <vector>
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

#include "../../multi_orb_slam_b200/csrc/octree_core.h"

extern "C" int octree_host_distribute(const int* x, const int* y, const int* score, int M, int W, int H,
                                      int N, int* out_x, int* out_y, int* out_score, int cap_out) {
  OtRoots roots;
  roots.n_ini = (int)std::round((float)W / (float)H);
  if (roots.n_ini < 1 || roots.n_ini > OT_MAX_ROOTS) return -1;
  roots.hx = (float)W / roots.n_ini;
  for (int i = 0; i <= roots.n_ini; ++i) roots.root_x[i] = (int)(roots.hx * (float)i);
  roots.height = H;
  int cap = (N + 3 > 4 * roots.n_ini ? N + 3 : 4 * roots.n_ini) + 1;
  OtScratch s;
  const int bytes = ot_layout(s, cap, cap + 1, OT_NTHREADS);
  std::vector<unsigned long long> arena(bytes / 8 + 2);
  s.host_base = reinterpret_cast<unsigned char*>(arena.data());
  std::vector<uint32_t> keys(M), out(cap);
  std::vector<uint16_t> knode(M);
  for (int i = 0; i < M; ++i) keys[i] = (uint32_t)x[i] | (uint32_t)y[i] << 12 | (uint32_t)score[i] << 24;
  int n = ot_distribute(keys.data(), knode.data(), M, roots, N, s, out.data());
  if (n > cap_out) return -2;
  for (int i = 0; i < n; ++i) { out_x[i] = OT_KEY_X(out[i]); out_y[i] = OT_KEY_Y(out[i]); out_score[i] = OT_KEY_SCORE(out[i]); }
  return n;
}
