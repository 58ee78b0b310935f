// Host build of the product's octree policy core (multi_orb_slam_b200/csrc/octree_core.h) under
// its sequential thread emulation, so tests/test_octree_host.py can check the parallel
// reformulation against the oracle without a GPU.  This is a TEST of product code, not a
// fallback: nothing in the package loads it.
#include <cmath>
#include <cstdlib>
#include <vector>

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
  std::vector<OtNode> nodes(cap);
  std::vector<int> o0(cap), o1(cap), P(cap), sp0(cap), sp1(cap), c40(4 * cap), c41(4 * cap), ch0(4 * cap), ch1(4 * cap),
      a(cap + 1), b(cap + 1), c(cap + 1), d(cap + 1), part(OT_NTHREADS + 1), vars(8);
  std::vector<unsigned long long> best(cap);
  OtScratch s;
  s.nodes = nodes.data();
  s.order[0] = o0.data(); s.order[1] = o1.data();
  s.P = P.data();
  s.split[0] = sp0.data(); s.split[1] = sp1.data();
  s.cnt4[0] = c40.data(); s.cnt4[1] = c41.data();
  s.child[0] = ch0.data(); s.child[1] = ch1.data();
  s.a = a.data(); s.b = b.data(); s.c = c.data(); s.d = d.data();
  s.part = part.data(); s.best = best.data(); s.vars = vars.data();
  std::vector<uint32_t> keys(M), out(cap);
  std::vector<uint16_t> knode(M);
  for (int i = 0; i < M; ++i) keys[i] = (uint32_t)x[i] | (uint32_t)y[i] << 12 | (uint32_t)score[i] << 24;
  int n = ot_distribute(keys.data(), knode.data(), M, roots, N, s, out.data());
  if (n > cap_out) return -2;
  for (int i = 0; i < n; ++i) { out_x[i] = OT_KEY_X(out[i]); out_y[i] = OT_KEY_Y(out[i]); out_score[i] = OT_KEY_SCORE(out[i]); }
  return n;
}
