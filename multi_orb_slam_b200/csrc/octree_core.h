// Parallel reformulation of the reference's keypoint culling, ORBextractor::DistributeOctTree
// (src/ORBextractor.cc:540-764) + ExtractorNode::DivideNode (:482-538), reproducing the exact
// survivor set AND list order of the sequential std::list algorithm.
//
// Idea: node geometry is a pure function of the root box and the split path, so the
// sequential policy (phase-1 "split everything" passes, then phase-2 "largest first with early
// break") can be replayed in ROUNDS.  One round = a set of nodes P split "simultaneously":
//   1. P = nodes with >1 key, in list order (phase 1) or sorted by (count desc, creation seq
//      desc) (phase 2: std::sort ascending on (size, node address) walked from the back; the
//      address is pinned to creation order);
//   2. every key of a P node finds its quadrant -> per-child counts (shared-memory atomics);
//   3. prefix sums over P give the list size after each split, hence the early-break index
//      (:731-732), the children's creation sequence and the new list order (children are
//      push_front'ed, so they appear reversed in front; untouched nodes keep their order);
//   4. keys are re-labelled with their node's new list position.
// One CTA handles one (frame, level).  The same source compiles for the host with the OT_*
// macros expanding to a sequential thread emulation (tests/native/octree_host.cc), which is
// how the logic is validated against the oracle without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define OT_DEV __device__ __forceinline__
#define OT_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define OT_FOR_TID(t, T) for (int t = threadIdx.x, _once = 1; _once; _once = 0)
#define OT_NTHREADS ((int)blockDim.x)
#define OT_SYNC() __syncthreads()
#define OT_SINGLE if (threadIdx.x == 0)
#define OT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define OT_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define OT_ATOMIC_MAX64(p, v) atomicMax((p), (v))
#define OT_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define OT_DEV static inline
#define OT_FOR(i, n) for (int i = 0; i < (n); ++i)
#define OT_FOR_TID(t, T) for (int t = 0; t < (T); ++t)
#define OT_NTHREADS 256
#define OT_SYNC() ((void)0)
#define OT_SINGLE
#define OT_ATOMIC_ADD(p, v) (*(p) += (v))
#define OT_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#define OT_ATOMIC_MAX64(p, v) (*(p) = (*(p) > (v) ? *(p) : (v)))
#define OT_FDIV(a, b) ((a) / (b))
#endif

#define OT_MAX_ROOTS 16
#define OT_POS_MASK 0x3FFFu
#define OT_MAX_NODES 0x3FFF

// Candidate key: x | y<<12 | score<<24, (x,y) relative to (minBorderX, minBorderY) = (16,16).
#define OT_KEY_X(k) ((int)((k) & 0xFFFu))
#define OT_KEY_Y(k) ((int)(((k) >> 12) & 0xFFFu))
#define OT_KEY_SCORE(k) ((int)((k) >> 24))

struct OtNode {
  short ulx, uly, brx, bry;  // UL and BR corners; UR.x == BR.x and BL.y == BR.y throughout
  int count;                 // keys inside
  int seq;                   // creation sequence (stands in for the heap address in :685)
};

// Per (frame, level) root geometry, precomputed on the host with the reference's float
// arithmetic (:544-564).
struct OtRoots {
  int n_ini;                       // round((float)W / H)
  float hx;                        // (float)W / n_ini
  int root_x[OT_MAX_ROOTS + 1];    // (int)(hx * i)
  int height;                      // maxY - minY
};

// Scratch (shared memory on the device).  cap = node capacity >= max(N + 3, 4 * n_ini) + 1.
struct OtScratch {
  OtNode* nodes[2];   // [cap] x2   current / next list, index = list position (0 = front)
  int* P;             // [cap]      processing order -> list position
  int* rankP;         // [cap]      list position -> index in P, or -1
  int* cnt4;          // [4*cap]    per P entry, keys per child n1..n4
  int* a;             // [cap + 1]  scan input / scratch
  int* b;             // [cap + 1]  scan output
  int* c;             // [cap + 1]  second scan output
  int* newpos;        // [cap]      old list position -> new list position (untouched nodes)
  int* childpos;      // [4*cap]    P entry, child -> new list position
  int* part;          // [OT_NTHREADS + 1] scan partials
  unsigned long long* best;  // [cap] arg-max accumulator of the final stage
  int* vars;          // [8]
};
enum { OT_V_S = 0, OT_V_NP, OT_V_JSTOP, OT_V_TOTAL, OT_V_TOTAL2, OT_V_SEQ };

// Block-wide exclusive scan: out[i] = sum(in[0..i)), *total = sum(in[0..n)).  in != out.
// Each thread owns `chunk` consecutive elements.  Device: warp-shuffle scan of the per-thread
// sums + one shared-memory hop across warps (2 barriers).  Host emulation: the same data flow
// with the cross-thread step done serially.
OT_DEV void ot_exclusive_scan(const int* in, int* out, int n, int* total, int* part) {
  const int T = OT_NTHREADS;
  const int chunk = (n + T - 1) / T;
#ifdef __CUDACC__
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = T >> 5;
  const int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
  int s = 0;
  for (int i = lo; i < hi; ++i) s += in[i];
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) part[warp] = incl;
  __syncthreads();
  int run = incl - s;
  for (int w = 0; w < warp; ++w) run += part[w];
  if (t == T - 1) *total = run + s;
  for (int i = lo; i < hi; ++i) { const int v = in[i]; out[i] = run; run += v; }
  (void)nwarp;
  __syncthreads();
#else
  for (int t = 0; t < T; ++t) {
    int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n, s = 0;
    for (int i = lo; i < hi; ++i) s += in[i];
    part[t] = s;
  }
  int run = 0;
  for (int t = 0; t < T; ++t) { int v = part[t]; part[t] = run; run += v; }
  *total = run;
  for (int t = 0; t < T; ++t) {
    int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n, r = part[t];
    for (int i = lo; i < hi; ++i) { int v = in[i]; out[i] = r; r += v; }
  }
#endif
}

// One split round over the nodes flagged in rankP/P (np entries).  early_break_n < 0 disables
// the early break (phase 1).  Returns the new list size; flips *cur.
OT_DEV int ot_round(const uint32_t* keys, uint16_t* knode, int M, OtScratch& s, int* cur, int S, int np,
                    int early_break_n) {
  const OtNode* L = s.nodes[*cur];
  OtNode* Lnew = s.nodes[*cur ^ 1];
  OT_FOR(i, 4 * np) s.cnt4[i] = 0;
  OT_SYNC();
  // 2. quadrant of every key that lives in a node being split (DivideNode :521-535)
  OT_FOR(k, M) {
    const unsigned pos = knode[k] & OT_POS_MASK;
    const int j = s.rankP[pos];
    if (j >= 0) {
      const OtNode nd = L[pos];
      const int mx = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1);  // UL.x + ceil(w/2)
      const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
      const uint32_t key = keys[k];
      const int q = (OT_KEY_X(key) < mx ? 0 : 1) + (OT_KEY_Y(key) < my ? 0 : 2);
      OT_ATOMIC_ADD(&s.cnt4[4 * j + q], 1);
      knode[k] = (uint16_t)(pos | (unsigned)q << 14);
    }
  }
  OT_SYNC();
  // 3. non-empty children per split; prefix sums in processing order
  OT_FOR(j, np) {
    int nne = 0;
    for (int q = 0; q < 4; ++q) nne += s.cnt4[4 * j + q] > 0;
    s.a[j] = nne;
  }
  OT_SINGLE s.vars[OT_V_JSTOP] = np - 1;
  OT_SYNC();
  ot_exclusive_scan(s.a, s.b, np, &s.vars[OT_V_TOTAL], s.part);  // b[j] = children created before j
  if (early_break_n >= 0) {
    // list size after splitting P[0..j] = S + (b[j] + a[j]) - (j + 1); first j reaching N stops
    OT_FOR(j, np) {
      if (S + s.b[j] + s.a[j] - (j + 1) >= early_break_n) OT_ATOMIC_MIN(&s.vars[OT_V_JSTOP], j);
    }
    OT_SYNC();
  }
  const int jstop = s.vars[OT_V_JSTOP];
  const int C = np > 0 ? s.b[jstop] + s.a[jstop] : 0;  // children pushed this round
  const int seq_base = s.vars[OT_V_SEQ];
  // 4a. untouched nodes keep their relative order behind the new children
  OT_FOR(p, S) {
    const int j = s.rankP[p];
    s.c[p] = !(j >= 0 && j <= jstop);
  }
  OT_SYNC();
  ot_exclusive_scan(s.c, s.newpos, S, &s.vars[OT_V_TOTAL2], s.part);
  OT_FOR(p, S) {
    if (s.c[p]) {
      const int np_ = C + s.newpos[p];
      s.newpos[p] = np_;
      Lnew[np_] = L[p];
    }
  }
  // 4b. children: creation order (j asc, n1..n4), pushed to the front => reversed positions
  OT_FOR(j, jstop + 1) {
    const OtNode nd = L[s.P[j]];
    const int mx = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1);
    const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
    int cr = s.b[j];
    for (int q = 0; q < 4; ++q) {
      const int cnt = s.cnt4[4 * j + q];
      if (cnt == 0) continue;
      OtNode ch;
      ch.ulx = (short)((q & 1) ? mx : nd.ulx);
      ch.brx = (short)((q & 1) ? nd.brx : mx);
      ch.uly = (short)((q & 2) ? my : nd.uly);
      ch.bry = (short)((q & 2) ? nd.bry : my);
      ch.count = cnt;
      ch.seq = seq_base + cr;
      const int pos = C - 1 - cr;
      Lnew[pos] = ch;
      s.childpos[4 * j + q] = pos;
      ++cr;
    }
  }
  OT_SYNC();
  // 5. re-label keys with their node's new list position
  OT_FOR(k, M) {
    const unsigned v = knode[k];
    const unsigned pos = v & OT_POS_MASK;
    const int j = s.rankP[pos];
    knode[k] = (uint16_t)((j >= 0 && j <= jstop) ? s.childpos[4 * j + (v >> 14)] : s.newpos[pos]);
  }
  OT_SYNC();
  const int S_new = S + C - (np > 0 ? jstop + 1 : 0);
  OT_SINGLE s.vars[OT_V_SEQ] = seq_base + C;
  *cur ^= 1;
  OT_SYNC();
  return S_new;
}

// Build P = nodes with count > 1.  sorted == 0: list order (phase-1 pass, :607-666).
// sorted == 1: (count desc, seq desc), i.e. the reference's ascending sort on
// (size, address) walked from the back (:685-686).  Returns np.
OT_DEV int ot_build_P(OtScratch& s, int cur, int S, int sorted) {
  const OtNode* L = s.nodes[cur];
  if (!sorted) {
    OT_FOR(p, S) s.a[p] = L[p].count > 1;
    OT_SYNC();
    ot_exclusive_scan(s.a, s.b, S, &s.vars[OT_V_NP], s.part);
    OT_FOR(p, S) {
      if (s.a[p]) { s.rankP[p] = s.b[p]; s.P[s.b[p]] = p; }
      else s.rankP[p] = -1;
    }
    OT_SYNC();
  } else {
    OT_SINGLE s.vars[OT_V_NP] = 0;
    OT_SYNC();
    OT_FOR(p, S) {
      const int cp = L[p].count, sp = L[p].seq;
      if (cp > 1) {
        int r = 0;
        for (int o = 0; o < S; ++o) {
          const int co = L[o].count;
          r += (co > 1) && (co > cp || (co == cp && L[o].seq > sp));
        }
        s.rankP[p] = r;
        s.P[r] = p;
        OT_ATOMIC_ADD(&s.vars[OT_V_NP], 1);
      } else {
        s.rankP[p] = -1;
      }
    }
    OT_SYNC();
  }
  return s.vars[OT_V_NP];
}

// Full culling of one (frame, level).  keys[M] in vToDistributeKeys order; knode[M] scratch.
// Writes selected keys to out[] in final list order; returns their count (<= cap).
OT_DEV int ot_distribute(const uint32_t* keys, uint16_t* knode, int M, const OtRoots& roots, int N,
                         OtScratch& s, uint32_t* out) {
  // roots (:546-586): key -> root by float division, empty roots dropped
  const int nIni = roots.n_ini;
  OT_FOR(i, nIni) s.a[i] = 0;
  OT_SINGLE s.vars[OT_V_SEQ] = nIni;
  OT_SYNC();
  OT_FOR(k, M) {
    const int r = (int)OT_FDIV((float)OT_KEY_X(keys[k]), roots.hx);
    knode[k] = (uint16_t)r;
    OT_ATOMIC_ADD(&s.a[r], 1);
  }
  OT_SYNC();
  OT_FOR(i, nIni) s.c[i] = s.a[i] > 0;
  OT_SYNC();
  ot_exclusive_scan(s.c, s.newpos, nIni, &s.vars[OT_V_S], s.part);
  int cur = 0;
  OT_FOR(i, nIni) {
    if (s.c[i]) {
      OtNode nd;
      nd.ulx = (short)roots.root_x[i];
      nd.brx = (short)roots.root_x[i + 1];
      nd.uly = 0;
      nd.bry = (short)roots.height;
      nd.count = s.a[i];
      nd.seq = i;
      s.nodes[0][s.newpos[i]] = nd;
    }
  }
  OT_SYNC();
  OT_FOR(k, M) knode[k] = (uint16_t)s.newpos[knode[k]];
  OT_SYNC();
  int S = s.vars[OT_V_S];

  // policy replay (:598-741)
  bool finish = false;
  while (!finish) {
    const int prev = S;
    int np = ot_build_P(s, cur, S, 0);
    S = ot_round(keys, knode, M, s, &cur, S, np, -1);
    // nToExpand = children with >1 key = all nodes with >1 key after a full pass
    OT_FOR(p, S) s.a[p] = s.nodes[cur][p].count > 1;
    OT_SYNC();
    ot_exclusive_scan(s.a, s.b, S, &s.vars[OT_V_TOTAL], s.part);
    const int nToExpand = s.vars[OT_V_TOTAL];
    OT_SYNC();
    if (S >= N || S == prev) {
      finish = true;
    } else if (S + nToExpand * 3 > N) {
      while (!finish) {
        const int prev2 = S;
        np = ot_build_P(s, cur, S, 1);
        S = ot_round(keys, knode, M, s, &cur, S, np, N);
        if (S >= N || S == prev2) finish = true;
      }
    }
  }

  // retain the best key per node (:745-761): max response, earliest key on ties
  OT_FOR(p, S) s.best[p] = 0ull;
  OT_SYNC();
  OT_FOR(k, M) {
    const unsigned long long v =
        ((unsigned long long)(OT_KEY_SCORE(keys[k]) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
    OT_ATOMIC_MAX64(&s.best[knode[k] & OT_POS_MASK], v);
  }
  OT_SYNC();
  OT_FOR(p, S) out[p] = keys[0xFFFFFFFFu - (uint32_t)(s.best[p] & 0xFFFFFFFFull)];
  OT_SYNC();
  return S;
}
