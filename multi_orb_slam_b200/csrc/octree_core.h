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
//   2. ONE pass over the keys per round: a key first picks up the child it fell into when its
//      node was split in the previous round, then — if its (new) node is in P — finds its
//      quadrant and bumps that child's counter (shared-memory atomics);
//   3. prefix sums over P give the list size after each split, hence the early-break index
//      (:731-732), the children's creation sequence and the new list order (children are
//      push_front'ed, so they appear reversed in front; untouched nodes keep their order).
// Nodes live in stable slots (the first non-empty child reuses its parent's slot), the list is
// an array of slots.  One CTA handles one (frame, level).  The same source compiles for the
// host with the OT_* macros expanding to a sequential thread emulation
// (tests/native/octree_host.cc), which is how the logic is validated against the oracle
// without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define OT_DEV __device__ __forceinline__
#define OT_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
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

// Scratch (shared memory on the device).  cap = node capacity >= max(N + 3, 4 * n_ini) + 1;
// the scan arrays a..d hold max(cap, cells of the level) + 1 entries (the caller's gather
// reuses a/b).
struct OtScratch {
  OtNode* nodes;      // [cap]      stable slots
  int* order[2];      // [cap] x2   list position -> slot (current / next)
  int* P;             // [cap]      processing order -> slot
  int* split[2];      // [cap] x2   slot -> index in P (this round / previous round), -1 = not split
  int* cnt4[2];       // [4*cap] x2 per P entry, keys per child n1..n4 (this / previous round)
  int* child[2];      // [4*cap] x2 per P entry, child -> slot (this / previous round)
  int* a;             // scan scratch
  int* b;
  int* c;
  int* d;
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
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
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

// Build P = nodes with count > 1 and split[cur][slot] = index in P (or -1) over the current
// list.  sorted == 0: list order (phase-1 pass, :607-666).  sorted == 1: (count desc, seq desc),
// i.e. the reference's ascending sort on (size, address) walked from the back (:685-686).
// Also zeroes the child counters of the round.  Returns np.
OT_DEV int ot_build_P(OtScratch& s, int cur, int lst, int S, int sorted) {
  const int* L = s.order[lst];
  int* split = s.split[cur];
  if (!sorted) {
    OT_FOR(p, S) s.a[p] = s.nodes[L[p]].count > 1;
    OT_SYNC();
    ot_exclusive_scan(s.a, s.b, S, &s.vars[OT_V_NP], s.part);
    OT_FOR(p, S) {
      const int slot = L[p];
      if (s.a[p]) { split[slot] = s.b[p]; s.P[s.b[p]] = slot; }
      else split[slot] = -1;
    }
  } else {
    OT_SINGLE s.vars[OT_V_NP] = 0;
    OT_SYNC();
    OT_FOR(p, S) {
      const int slot = L[p];
      const int cp = s.nodes[slot].count, sp = s.nodes[slot].seq;
      if (cp > 1) {
        int r = 0;
        for (int o = 0; o < S; ++o) {
          const OtNode& other = s.nodes[L[o]];
          r += (other.count > 1) && (other.count > cp || (other.count == cp && other.seq > sp));
        }
        split[slot] = r;
        s.P[r] = slot;
        OT_ATOMIC_ADD(&s.vars[OT_V_NP], 1);
      } else {
        split[slot] = -1;
      }
    }
  }
  OT_SYNC();
  const int np = s.vars[OT_V_NP];
  OT_FOR(i, 4 * np) s.cnt4[cur][i] = 0;
  OT_SYNC();
  return np;
}

// The per-round key pass.  prev >= 0: first move every key of a node split in the previous
// round (P index <= jstop_prev) into the child it fell into.  cur >= 0: then, if the key's node
// is about to be split, find its quadrant (DivideNode :521-535) and count it.
OT_DEV void ot_key_pass(const uint32_t* keys, uint16_t* knode, int M, OtScratch& s, int prev, int jstop_prev, int cur) {
  OT_FOR(k, M) {
    unsigned v = knode[k];
    int slot = (int)(v & OT_POS_MASK);
    if (prev >= 0) {
      const int jp = s.split[prev][slot];
      if (jp >= 0 && jp <= jstop_prev) slot = s.child[prev][4 * jp + (int)(v >> 14)];
      v = (unsigned)slot;
    }
    if (cur >= 0) {
      const int j = s.split[cur][slot];
      if (j >= 0) {
        const OtNode nd = s.nodes[slot];
        const int mx = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1);  // UL.x + ceil(w/2)
        const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
        const uint32_t key = keys[k];
        const int q = (OT_KEY_X(key) < mx ? 0 : 1) + (OT_KEY_Y(key) < my ? 0 : 2);
        OT_ATOMIC_ADD(&s.cnt4[cur][4 * j + q], 1);
        v = (unsigned)slot | (unsigned)q << 14;
      }
    }
    knode[k] = (uint16_t)v;
  }
  OT_SYNC();
}

// Commit one round: given P (np entries), its child counters and the early-break limit
// (early_break_n < 0: none), create the children, rebuild the list.  Returns the new list size;
// *jstop_out = last processed P index.
OT_DEV int ot_commit_round(OtScratch& s, int cur, int lst, int S, int np, int early_break_n, int* jstop_out) {
  const int* L = s.order[lst];
  int* Lnew = s.order[lst ^ 1];
  const int* cnt4 = s.cnt4[cur];
  // non-empty children per split; prefix sums in processing order
  OT_FOR(j, np) {
    int nne = 0;
    for (int q = 0; q < 4; ++q) nne += cnt4[4 * j + q] > 0;
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
  // untouched nodes keep their relative order behind the new children
  OT_FOR(p, S) {
    const int j = s.split[cur][L[p]];
    s.c[p] = !(j >= 0 && j <= jstop);
  }
  OT_SYNC();
  ot_exclusive_scan(s.c, s.d, S, &s.vars[OT_V_TOTAL2], s.part);
  OT_FOR(p, S) {
    if (s.c[p]) Lnew[C + s.d[p]] = L[p];
  }
  // children: creation order (j asc, n1..n4), pushed to the front => reversed positions.  The
  // first non-empty child reuses the parent's slot, the others take fresh slots from S upward.
  OT_FOR(j, jstop + 1) {
    const int pslot = s.P[j];
    const OtNode nd = s.nodes[pslot];
    const int mx = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1);
    const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
    int cr = s.b[j];
    int fresh = S + s.b[j] - j;  // slots handed out to earlier splits: sum(nne - 1)
    bool first = true;
    for (int q = 0; q < 4; ++q) {
      const int cnt = cnt4[4 * j + q];
      if (cnt == 0) continue;
      OtNode ch;
      ch.ulx = (short)((q & 1) ? mx : nd.ulx);
      ch.brx = (short)((q & 1) ? nd.brx : mx);
      ch.uly = (short)((q & 2) ? my : nd.uly);
      ch.bry = (short)((q & 2) ? nd.bry : my);
      ch.count = cnt;
      ch.seq = seq_base + cr;
      const int slot = first ? pslot : fresh++;
      first = false;
      s.nodes[slot] = ch;
      s.child[cur][4 * j + q] = slot;
      Lnew[C - 1 - cr] = slot;
      ++cr;
    }
  }
  OT_SYNC();
  OT_SINGLE s.vars[OT_V_SEQ] = seq_base + C;
  *jstop_out = jstop;
  OT_SYNC();
  return S + C - (np > 0 ? jstop + 1 : 0);
}

// Full culling of one (frame, level).  keys[M] in vToDistributeKeys order; knode[M] scratch.
// Writes selected keys to out[] in final list order; returns their count.
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
  ot_exclusive_scan(s.c, s.d, nIni, &s.vars[OT_V_S], s.part);
  int lst = 0, cur = 0;
  // Non-empty root i gets slot == list position d[i].  The keys still carry the root index; they
  // are re-labelled by the first key pass through a pseudo round "-1" stored in split[1]/child[1].
  OT_FOR(i, nIni) {
    s.split[1][i] = s.c[i] ? i : -1;
    if (s.c[i]) {
      OtNode nd;
      nd.ulx = (short)roots.root_x[i];
      nd.brx = (short)roots.root_x[i + 1];
      nd.uly = 0;
      nd.bry = (short)roots.height;
      nd.count = s.a[i];
      nd.seq = i;
      s.nodes[s.d[i]] = nd;
      s.order[0][s.d[i]] = s.d[i];
      s.child[1][4 * i] = s.d[i];
    }
  }
  OT_SYNC();
  int S = s.vars[OT_V_S];

  // policy replay (:598-741)
  int prev = 1, jstop_prev = nIni;
  int phase2 = 0;
  bool finish = false;
  while (!finish) {
    const int before = S;
    const int np = ot_build_P(s, cur, lst, S, phase2);
    ot_key_pass(keys, knode, M, s, prev, jstop_prev, cur);
    int jstop = -1;
    S = ot_commit_round(s, cur, lst, S, np, phase2 ? N : -1, &jstop);
    lst ^= 1;
    prev = cur;
    jstop_prev = jstop;
    cur ^= 1;
    if (S >= N || S == before) {
      finish = true;
    } else if (!phase2) {
      // nToExpand = children with >1 key = all nodes with >1 key after a full pass (:670-676)
      OT_FOR(p, S) s.a[p] = s.nodes[s.order[lst][p]].count > 1;
      OT_SYNC();
      ot_exclusive_scan(s.a, s.b, S, &s.vars[OT_V_TOTAL], s.part);
      const int nToExpand = s.vars[OT_V_TOTAL];
      OT_SYNC();
      if (S + nToExpand * 3 > N) phase2 = 1;
    }
  }
  // move the keys of the last round's split nodes into their children
  ot_key_pass(keys, knode, M, s, prev, jstop_prev, -1);

  // retain the best key per node (:745-761): max response, earliest key on ties
  OT_FOR(p, S) s.best[s.order[lst][p]] = 0ull;
  OT_SYNC();
  OT_FOR(k, M) {
    const unsigned long long v =
        ((unsigned long long)(OT_KEY_SCORE(keys[k]) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
    OT_ATOMIC_MAX64(&s.best[knode[k] & OT_POS_MASK], v);
  }
  OT_SYNC();
  OT_FOR(p, S) out[p] = keys[0xFFFFFFFFu - (uint32_t)(s.best[s.order[lst][p]] & 0xFFFFFFFFull)];
  OT_SYNC();
  return S;
}
