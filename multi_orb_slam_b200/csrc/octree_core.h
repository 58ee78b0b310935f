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
// an array of slots; a per-round "move table" mv[4*slot+q] sends a key to the child it fell into
// (identity for nodes that were not split), so the pass is three dependent shared-memory loads
// per key: label -> mv -> node.  All scratch is addressed as offsets from the CTA's dynamic
// shared memory so the compiler emits LDS/STS/ATOMS (a struct of generic pointers ended up in
// local memory with generic atomics — 5x slower, see profiles/).  One CTA handles one (frame, level).  The same source compiles for the
// host with the OT_* macros expanding to a sequential thread emulation
// (tests/native/octree_host.cc), which is how the logic is validated against the oracle
// without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define OT_DEV __device__ __forceinline__
#define OT_HD __host__ __device__ inline
#define OT_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define OT_NTHREADS ((int)blockDim.x)
#define OT_SYNC() __syncthreads()
#define OT_SINGLE if (threadIdx.x == 0)
#define OT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define OT_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define OT_ATOMIC_MAX64(p, v) atomicMax((p), (v))
#define OT_ATOMIC_MAX32(p, v) atomicMax((p), (v))
#define OT_FDIV(a, b) __fdiv_rn((a), (b))
extern __shared__ __align__(16) unsigned char ot_smem[];  // aliases the kernel's dynamic shared memory
#define OT_BASE(s) ot_smem
#else
#define OT_DEV static inline
#define OT_HD static inline
#define OT_FOR(i, n) for (int i = 0; i < (n); ++i)
#define OT_NTHREADS 256
#define OT_SYNC() ((void)0)
#define OT_SINGLE
#define OT_ATOMIC_ADD(p, v) (*(p) += (v))
#define OT_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#define OT_ATOMIC_MAX64(p, v) (*(p) = (*(p) > (v) ? *(p) : (v)))
#define OT_ATOMIC_MAX32(p, v) (*(p) = (*(p) > (v) ? *(p) : (v)))
#define OT_FDIV(a, b) ((a) / (b))
#define OT_BASE(s) ((s).host_base)
#endif
#define OT_INTS(s, off) (reinterpret_cast<int*>(OT_BASE(s) + (off)))
#ifndef OT_MARK
#define OT_MARK(id) ((void)0)  /* dev timing hook, see tools/dev/octree_bench.cu */
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

// Scratch layout: BYTE OFFSETS into the CTA's dynamic shared memory (host emulation: into
// host_base).  cap = node capacity >= max(N + 3, 4 * n_ini) + 1; scap = max(cap, cells of the
// level) + 1 entries for the scan arrays (the caller's gather reuses a/b).
struct OtScratch {
  int best;     // u64 [cap]    arg-max accumulator of the final stage (8-byte aligned)
  int nodes;    // OtNode [cap] stable slots
  int cnt4;     // int [4*cap]  keys per child n1..n4 of slot
  int mv;       // int [4*cap]  move table of the previous round
  int order0;   // int [cap]    list position -> slot (two buffers)
  int order1;
  int rankP;    // int [cap]    slot -> index in P (valid for slots with count > 1)
  int P;        // int [cap]    processing order -> slot
  int a, b, c, d;  // int [scap] scan scratch
  int part;     // int [OT_NTHREADS + 1]
  int vars;     // int [8]
  int total;    // bytes
  unsigned char* host_base;
};
enum { OT_V_S = 0, OT_V_NP, OT_V_JSTOP, OT_V_TOTAL, OT_V_TOTAL2, OT_V_SEQ };

OT_HD int ot_layout(OtScratch& s, int cap, int scap, int nthreads) {
  int o = 0;
#define OT_TAKE(field, bytes) s.field = o; o = (o + (bytes) + 15) & ~15  /* every array 16-byte aligned */
  OT_TAKE(best, 8 * cap);
  OT_TAKE(nodes, 16 * cap);
  OT_TAKE(cnt4, 16 * cap);
  OT_TAKE(mv, 16 * cap);
  OT_TAKE(order0, 4 * cap);
  OT_TAKE(order1, 4 * cap);
  OT_TAKE(rankP, 4 * cap);
  OT_TAKE(P, 4 * cap);
  OT_TAKE(a, 4 * scap);
  OT_TAKE(b, 4 * scap);
  OT_TAKE(c, 4 * scap);
  OT_TAKE(d, 4 * scap);
  OT_TAKE(part, 4 * (nthreads + 1));
  OT_TAKE(vars, 32);
#undef OT_TAKE
  s.total = o;
  s.host_base = 0;
  return o;
}

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

// Build P = nodes with count > 1 and rankP[slot] = index in P over the current list L.
// sorted == 0: list order (phase-1 pass, :607-666).  sorted == 1: (count desc, seq desc), i.e.
// the reference's ascending sort on (size, address) walked from the back (:685-686).
// Also zeroes the child counters of the listed slots.  Returns np.
OT_DEV int ot_build_P(const OtScratch& s, const int* L, int S, int sorted) {
  const OtNode* nodes = reinterpret_cast<const OtNode*>(OT_BASE(s) + s.nodes);
  int* rankP = OT_INTS(s, s.rankP);
  int* P = OT_INTS(s, s.P);
  int* a = OT_INTS(s, s.a);
  int* b = OT_INTS(s, s.b);
  int* vars = OT_INTS(s, s.vars);
  int4* cnt4 = reinterpret_cast<int4*>(OT_BASE(s) + s.cnt4);
  if (!sorted) {
    OT_FOR(p, S) {
      a[p] = nodes[L[p]].count > 1;
      cnt4[L[p]] = make_int4(0, 0, 0, 0);
    }
    OT_SYNC();
    ot_exclusive_scan(a, b, S, &vars[OT_V_NP], OT_INTS(s, s.part));
    OT_FOR(p, S) {
      if (a[p]) { rankP[L[p]] = b[p]; P[b[p]] = L[p]; }
    }
  } else {
    OT_SINGLE vars[OT_V_NP] = 0;
    // stage (count, seq) of the multi-key nodes so the O(S^2) ranking reads two flat arrays
    OT_FOR(p, S) {
      const OtNode nd = nodes[L[p]];
      a[p] = nd.count > 1 ? nd.count : 0;
      b[p] = nd.seq;
      cnt4[L[p]] = make_int4(0, 0, 0, 0);
    }
    OT_SYNC();
    OT_FOR(p, S) {
      const int cp = a[p], sp = b[p];
      if (cp > 1) {
        int r = 0;
        for (int o = 0; o < S; ++o) r += (a[o] > cp) || (a[o] == cp && b[o] > sp);
        rankP[L[p]] = r;
        P[r] = L[p];
        OT_ATOMIC_ADD(&vars[OT_V_NP], 1);
      }
    }
  }
  OT_SYNC();
  return vars[OT_V_NP];
}

// The per-round key pass: a key first follows the previous round's move table into the child it
// fell into; then, if its node holds more than one key (= is in P), it finds its quadrant
// (DivideNode :521-535) and counts itself.  count_next == 0: relabel only.
OT_DEV void ot_key_pass(const uint32_t* keys, uint16_t* knode, int M, const OtScratch& s, int count_next) {
  const OtNode* nodes = reinterpret_cast<const OtNode*>(OT_BASE(s) + s.nodes);
  const int* mv = OT_INTS(s, s.mv);
  int* cnt4 = OT_INTS(s, s.cnt4);
  // four keys per thread and iteration: the dependent chain label -> mv -> node is issued for all
  // four before the first atomic, so the shared-memory latencies overlap
  const int T = OT_NTHREADS;
#ifdef __CUDACC__
  for (int k0 = threadIdx.x; k0 < M; k0 += 4 * T) {
#else
  for (int k0 = 0; k0 < M; k0 = (k0 % T == T - 1) ? k0 + 3 * T + 1 : k0 + 1) {
#endif
    unsigned v[4];
    int slot[4];
    OtNode nd[4];
    uint32_t key[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (k0 + u * T < M) ? knode[k0 + u * T] : 0u;
#pragma unroll
    // (lanes past the last key must not follow the table: their label 0 may name a root that holds no key, whose entry
    // was never written - an arbitrary slot would index the node array out of bounds)
    for (int u = 0; u < 4; ++u)
      slot[u] = (k0 + u * T < M) ? mv[4 * (int)(v[u] & OT_POS_MASK) + (int)(v[u] >> 14)] : 0;
    if (count_next) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        nd[u] = nodes[slot[u]];
        key[u] = (k0 + u * T < M) ? keys[k0 + u * T] : 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k0 + u * T >= M) continue;
      unsigned nv = (unsigned)slot[u];
      if (count_next && nd[u].count > 1) {
        const int mx = nd[u].ulx + ((nd[u].brx - nd[u].ulx + 1) >> 1);  // UL.x + ceil(w/2)
        const int my = nd[u].uly + ((nd[u].bry - nd[u].uly + 1) >> 1);
        const int q = (OT_KEY_X(key[u]) < mx ? 0 : 1) + (OT_KEY_Y(key[u]) < my ? 0 : 2);
        OT_ATOMIC_ADD(&cnt4[4 * slot[u] + q], 1);
        nv |= (unsigned)q << 14;
      }
      knode[k0 + u * T] = (uint16_t)nv;
    }
  }
  OT_SYNC();
}

// Commit one round: given P (np entries), the child counters and the early-break limit
// (early_break_n < 0: none), create the children, rebuild the list (L -> Lnew) and the move
// table.  Returns the new list size.
OT_DEV int ot_commit_round(const OtScratch& s, const int* L, int* Lnew, int S, int np, int early_break_n) {
  OtNode* nodes = reinterpret_cast<OtNode*>(OT_BASE(s) + s.nodes);
  const int* cnt4 = OT_INTS(s, s.cnt4);
  int* mv = OT_INTS(s, s.mv);
  const int* rankP = OT_INTS(s, s.rankP);
  const int* P = OT_INTS(s, s.P);
  int* a = OT_INTS(s, s.a);
  int* b = OT_INTS(s, s.b);
  int* c = OT_INTS(s, s.c);
  int* d = OT_INTS(s, s.d);
  int* vars = OT_INTS(s, s.vars);
  int* part = OT_INTS(s, s.part);
  // non-empty children per split; prefix sums in processing order
  OT_FOR(j, np) {
    int nne = 0;
    for (int q = 0; q < 4; ++q) nne += cnt4[4 * P[j] + q] > 0;
    a[j] = nne;
  }
  OT_SINGLE vars[OT_V_JSTOP] = np - 1;
  OT_SYNC();
  ot_exclusive_scan(a, b, np, &vars[OT_V_TOTAL], part);  // b[j] = children created before j
  if (early_break_n >= 0) {
    // list size after splitting P[0..j] = S + (b[j] + a[j]) - (j + 1); first j reaching N stops
    OT_FOR(j, np) {
      if (S + b[j] + a[j] - (j + 1) >= early_break_n) OT_ATOMIC_MIN(&vars[OT_V_JSTOP], j);
    }
    OT_SYNC();
  }
  const int jstop = vars[OT_V_JSTOP];
  const int C = np > 0 ? b[jstop] + a[jstop] : 0;  // children pushed this round
  const int seq_base = vars[OT_V_SEQ];
  // untouched nodes keep their relative order behind the new children; their keys stay put
  OT_FOR(p, S) {
    const int slot = L[p];
    const bool touched = nodes[slot].count > 1 && rankP[slot] <= jstop;
    c[p] = !touched;
    if (!touched) {
      mv[4 * slot] = slot; mv[4 * slot + 1] = slot; mv[4 * slot + 2] = slot; mv[4 * slot + 3] = slot;
    }
  }
  OT_SYNC();
  ot_exclusive_scan(c, d, S, &vars[OT_V_TOTAL2], part);
  OT_FOR(p, S) {
    if (c[p]) Lnew[C + d[p]] = L[p];
  }
  // children: creation order (j asc, n1..n4), pushed to the front => reversed positions.  The
  // first non-empty child reuses the parent's slot, the others take fresh slots from S upward.
  OT_FOR(j, jstop + 1) {
    const int pslot = P[j];
    const OtNode nd = nodes[pslot];
    const int mx = nd.ulx + ((nd.brx - nd.ulx + 1) >> 1);
    const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
    int cr = b[j];
    int fresh = S + b[j] - j;  // slots handed out to earlier splits: sum(nne - 1)
    bool first = true;
    int cn[4];
    for (int q = 0; q < 4; ++q) cn[q] = cnt4[4 * pslot + q];
    for (int q = 0; q < 4; ++q) {
      if (cn[q] == 0) continue;
      OtNode ch;
      ch.ulx = (short)((q & 1) ? mx : nd.ulx);
      ch.brx = (short)((q & 1) ? nd.brx : mx);
      ch.uly = (short)((q & 2) ? my : nd.uly);
      ch.bry = (short)((q & 2) ? nd.bry : my);
      ch.count = cn[q];
      ch.seq = seq_base + cr;
      const int slot = first ? pslot : fresh++;
      first = false;
      nodes[slot] = ch;
      mv[4 * pslot + q] = slot;
      Lnew[C - 1 - cr] = slot;
      ++cr;
    }
  }
  OT_SYNC();
  OT_SINGLE vars[OT_V_SEQ] = seq_base + C;
  OT_SYNC();
  return S + C - (np > 0 ? jstop + 1 : 0);
}

// Full culling of one (frame, level).  keys[M] in vToDistributeKeys order; knode[M] scratch.
// Writes selected keys to out[] in final list order; returns their count.
OT_DEV int ot_distribute(const uint32_t* keys, uint16_t* knode, int M, const OtRoots& roots, int N,
                         const OtScratch& s, uint32_t* out) {
  OtNode* nodes = reinterpret_cast<OtNode*>(OT_BASE(s) + s.nodes);
  int* order0 = OT_INTS(s, s.order0);
  int* order1 = OT_INTS(s, s.order1);
  int* mv = OT_INTS(s, s.mv);
  int* a = OT_INTS(s, s.a);
  int* b = OT_INTS(s, s.b);
  int* c = OT_INTS(s, s.c);
  int* d = OT_INTS(s, s.d);
  int* vars = OT_INTS(s, s.vars);
  int* part = OT_INTS(s, s.part);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(OT_BASE(s) + s.best);
  // roots (:546-586): key -> root by float division, empty roots dropped
  const int nIni = roots.n_ini;
  OT_FOR(i, nIni) a[i] = 0;
  OT_SINGLE vars[OT_V_SEQ] = nIni;
  OT_SYNC();
  OT_FOR(k, M) {
    const int r = (int)OT_FDIV((float)OT_KEY_X(keys[k]), roots.hx);
    knode[k] = (uint16_t)r;
    OT_ATOMIC_ADD(&a[r], 1);
  }
  OT_SYNC();
  OT_FOR(i, nIni) c[i] = a[i] > 0;
  OT_SYNC();
  ot_exclusive_scan(c, d, nIni, &vars[OT_V_S], part);
  // Non-empty root i gets slot == list position d[i]; the keys still carry the root index and are
  // re-labelled by the first key pass through the move table (a pseudo round "-1").
  OT_FOR(i, nIni) {
    if (c[i]) {
      OtNode nd;
      nd.ulx = (short)roots.root_x[i];
      nd.brx = (short)roots.root_x[i + 1];
      nd.uly = 0;
      nd.bry = (short)roots.height;
      nd.count = a[i];
      nd.seq = i;
      nodes[d[i]] = nd;
      order0[d[i]] = d[i];
      mv[4 * i] = d[i];
    }
  }
  OT_SYNC();
  int S = vars[OT_V_S];
  OT_MARK(10);

  // policy replay (:598-741)
  int lst = 0, phase2 = 0;
  bool finish = false;
  while (!finish) {
    const int before = S;
    int* L = lst ? order1 : order0;
    int* Lnew = lst ? order0 : order1;
    OT_MARK(1);
    const int np = ot_build_P(s, L, S, phase2);
    OT_MARK(phase2 ? 3 : 2);
    ot_key_pass(keys, knode, M, s, 1);
    OT_MARK(4);
    S = ot_commit_round(s, L, Lnew, S, np, phase2 ? N : -1);
    OT_MARK(5);
    lst ^= 1;
    if (S >= N || S == before) {
      finish = true;
    } else if (!phase2) {
      // nToExpand = children with >1 key = all nodes with >1 key after a full pass (:670-676)
      OT_FOR(p, S) a[p] = nodes[Lnew[p]].count > 1;
      OT_SYNC();
      ot_exclusive_scan(a, b, S, &vars[OT_V_TOTAL], part);
      const int nToExpand = vars[OT_V_TOTAL];
      OT_SYNC();
      if (S + nToExpand * 3 > N) phase2 = 1;
    }
  }
  const int* Lf = lst ? order1 : order0;
  OT_MARK(6);
  // move the keys of the last round's split nodes into their children
  ot_key_pass(keys, knode, M, s, 0);
  OT_MARK(7);

  // retain the best key per node (:745-761): max response, earliest key on ties.  Packed
  // (score+1, ~index) arg-max: one native 32-bit shared atomic when the index fits 16 bits.
  if (M <= 0xFFFF) {
    unsigned* best32 = reinterpret_cast<unsigned*>(best);
    OT_FOR(p, S) best32[Lf[p]] = 0u;
    OT_SYNC();
    OT_FOR(k, M) {
      const unsigned v = ((unsigned)(OT_KEY_SCORE(keys[k]) + 1) << 16) | (0xFFFFu - (unsigned)k);
      OT_ATOMIC_MAX32(&best32[knode[k] & OT_POS_MASK], v);
    }
    OT_SYNC();
    OT_MARK(8);
    OT_FOR(p, S) out[p] = keys[0xFFFFu - (best32[Lf[p]] & 0xFFFFu)];
  } else {
    OT_FOR(p, S) best[Lf[p]] = 0ull;
    OT_SYNC();
    OT_FOR(k, M) {
      const unsigned long long v =
          ((unsigned long long)(OT_KEY_SCORE(keys[k]) + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
      OT_ATOMIC_MAX64(&best[knode[k] & OT_POS_MASK], v);
    }
    OT_SYNC();
    OT_MARK(8);
    OT_FOR(p, S) out[p] = keys[0xFFFFFFFFu - (uint32_t)(best[Lf[p]] & 0xFFFFFFFFull)];
  }
  OT_SYNC();
  return S;
}
