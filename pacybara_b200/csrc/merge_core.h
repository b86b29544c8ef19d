// merge_core.h -- the per-component cluster replay (seed step, greedy merge, consensus).
//
// One connected component of the "link" graph (buckets joined by a row with
// 1 <= dist <= maxDiff) is an independent unit of the reference's sequential algorithm
// (src/pacybara_cluster.py:385-455; the reference's own abandoned plan says the same,
// src/legacy/pacybara_parstrat.py:63-87).  process_component() replays that algorithm for
// one component, exactly, in the reference's visiting order.  On the GPU one thread runs
// one component (merge.cu); the function is __host__ __device__ so that the identical code
// can be exercised without a GPU by tests/host_emul (a unit-test harness, never a product
// path: the shipped library only ever launches it as a kernel).
//
// Formulation (equivalent to the reference's read-pair loops, proven against the oracle):
//  * a cluster is a linked list of reads (rd_next) hanging off a slot; slots form a
//    union-find so that "relabel every member" (:198-199) is O(alpha); the root carries the
//    reference-visible attributes (creation key of the surviving cid, head, tail, size);
//  * addLink(rid1, rid2) keeps rid1's cluster id and appends rid2's list (:193-202);
//  * the size rule |log2(s1/s2)| >= ed is the integer test max >= min << ed (exact for
//    every size below 2^26);
//  * the transitive rule (:442) only depends on the buckets of the members, so it runs over
//    bucket runs of the two member lists and looks distances up in the sorted adjacency;
//  * the genotype rule (:447-452) needs, per mutation, (#reads with Q>0, sum Q) of each side:
//    accumulated in a small open-addressing table in per-component scratch;
//  * a failed test between two clusters is remembered until the next merge in the component
//    (the state it depended on cannot change before that), which turns the |b1| x |b2|
//    read-pair expansion of a row into one test per distinct cluster pair.
#pragma once
#include <stdint.h>

#include "myers.cuh"

namespace pcb {

struct GenoSlot {          // 32 bytes; an all-zero slot is empty, so a memset(0) makes a table
  int32_t key;             // mutation id (or barcode id when counting) PLUS ONE, 0 = empty
  int32_t obs1, obs2;      // #reads with Q>0 on each side / #reads carrying the key
  int32_t pad;
  int64_t sum1, sum2;      // sum of Q on each side
};

struct MergeParams {
  int32_t maxDiff, minQual, minMatches;
  double minJaccard;
  // Jaccard as an integer threshold: jthr[u] = the smallest `inters` with (double)inters / (double)u > minJaccard
  // (u + 1 if there is none), for a union of u = 1..32 mutations.  The quotient grows with inters, so
  // "inters / uni > minJaccard" (:358-367, :447-452) is exactly "inters >= jthr[uni]" -- the same IEEE divide,
  // done once per call on the host (jaccard_thresholds) instead of once per test.
  int32_t jthr[33];
};

inline void jaccard_thresholds(MergeParams &p) {
  p.jthr[0] = 0;
  for (int32_t u = 1; u <= 32; u++) {
    p.jthr[u] = u + 1;
    for (int32_t i = u; i >= 0; i--) { if ((double)i / (double)u > p.minJaccard) p.jthr[u] = i; else break; }
  }
}

#ifndef PCB_FILL                        // same values as include/pacybara_b200.h
#define PCB_FILL INT32_MIN              // "not produced by this shard"
#define PCB_NEEDS_ALIGNMENT (-2)        // reference would call muscle (:530)
#endif

struct MergeCtx {
  // problem (read-only)
  int32_t R, U;
  const int32_t *bucket_ptr, *bucket_reads, *bucket_of_read, *lexrank, *bc_id, *up_id;
  const int32_t *geno_ptr, *geno_mut, *geno_q, *mut_rank;
  // unique bucket pairs: symmetric adjacency sorted by (a, nbr), and the per-component
  // visiting lists (pairs with 1 <= ed <= maxDiff, grouped by component, visiting order kept)
  const int32_t *adj_ptr, *adj_nbr, *adj_d;
  // optional on-demand distances: when compute_k >= 0 the adjacency is only complete up to
  // maxDiff and a pair it does not hold is scored on the spot (planes[2b], planes[2b+1] = bit
  // planes L, H of bucket b); the value is used when <= compute_k, else the pair is "absent"
  const uint64_t *planes;
  const int32_t *blen;
  int32_t compute_k, wide;
  const int32_t *comp_ptr, *comp_buckets;            // component -> buckets (ascending)
  const int32_t *cpair_ptr, *cpair_q, *cpair_r, *cpair_ed;
  // per-component scratch (global-memory variant; see MergeState)
  const int64_t *scr_slot_off;   // into slots[]  (capacity = scr_slot_cap[c], a power of two)
  const int32_t *scr_slot_cap;
  const int64_t *scr_list_off;   // into lists[]  (capacity 2 * (#reads + #geno entries) of the component)
  const int64_t *call_off;       // into call_mut[] (capacity = #geno entries of the component); nullptr: regions are taken from
  int64_t *call_cursor;          //   this shared cursor as the components come by (merge_small.cuh, merge.cu)
  const int32_t *bucket_gid;     // optional: bucket number -> bucket id in the adjacency (components replayed on a local copy)
  int32_t adj_forward;           // the adjacency holds every pair once, in the row of its smaller bucket id
  GenoSlot *slots;
  int32_t *lists;
  // cluster state in global memory (size R)
  int32_t *rd_slot, *rd_next;
  int32_t *sl_parent, *sl_head, *sl_tail, *sl_size;
  int64_t *sl_key;
  // outputs (size R unless noted)
  int32_t *read_root, *read_pos;                     // representative read of the row, position in it
  int32_t *row_size, *row_bc, *row_up, *row_call_n;  // valid at representative reads
  int64_t *row_key, *row_call_off;
  int32_t *call_mut;                                 // [nnz]
  int32_t *status;                                   // [4]: error code, detail, #links, #tests
  int32_t *cell_nnz;                                 // optional (merge_huge.cuh): genotype entries per cluster, kept at its root slot
  MergeParams prm;
};

// The mutable working set of ONE component: cluster state of its reads (indexed by read - r0) and its
// scratch table / lists.  It lives either in the global arrays of MergeCtx (r0 = 0) or, for components
// small enough, in shared memory (r0 = first read of the component) -- every read-modify-write of the
// replay then costs a shared-memory round trip instead of an L2 one.
struct MergeState {
  int32_t r0;
  int32_t *rd_slot, *rd_next;
  int32_t *sl_parent, *sl_head, *sl_tail, *sl_size;
  int64_t *sl_key;
  GenoSlot *slots;
  int32_t slot_mask;
  int32_t *lists;
  int64_t list_cap;          // lists holds 8 + 2 * list_cap ints
};

PCB_HD MergeState global_state(const MergeCtx &c, int32_t comp) {
  MergeState s;
  s.r0 = 0;
  s.rd_slot = c.rd_slot; s.rd_next = c.rd_next; s.sl_parent = c.sl_parent; s.sl_head = c.sl_head;
  s.sl_tail = c.sl_tail; s.sl_size = c.sl_size; s.sl_key = c.sl_key;
  s.slots = c.slots + c.scr_slot_off[comp];
  s.slot_mask = c.scr_slot_cap[comp] - 1;
  s.lists = c.lists + c.scr_list_off[comp];
  s.list_cap = (c.scr_list_off[comp + 1] - c.scr_list_off[comp]) / 2 - 4;     // region = 8 + 2 * list_cap ints
  return s;
}

PCB_HD int32_t slot_find(const MergeCtx &c, const MergeState &st, int32_t s) {
  while (st.sl_parent[(s) - st.r0] != s) {
    int32_t g = st.sl_parent[(st.sl_parent[(s) - st.r0]) - st.r0];
    st.sl_parent[(s) - st.r0] = g;
    s = g;
  }
  return s;
}
PCB_HD int32_t cluster_of(const MergeCtx &c, const MergeState &st, int32_t r) {
  int32_t s = st.rd_slot[(r) - st.r0];
  return s < 0 ? -1 : slot_find(c, st, s);
}

// src/pacybara_cluster.py:185-220
PCB_HD void add_link(const MergeCtx &c, const MergeState &st, int32_t rid1, int32_t rid2, int64_t new_key) {
  int32_t c1 = cluster_of(c, st, rid1), c2 = cluster_of(c, st, rid2);
  if (c1 >= 0) {
    if (c2 >= 0) {
      if (c1 == c2) return;
      int32_t head = st.sl_head[(c1) - st.r0], tail = st.sl_tail[(c2) - st.r0], size = st.sl_size[(c1) - st.r0] + st.sl_size[(c2) - st.r0];
      int64_t key = st.sl_key[(c1) - st.r0];
      st.rd_next[(st.sl_tail[(c1) - st.r0]) - st.r0] = st.sl_head[(c2) - st.r0];
      int32_t root = st.sl_size[(c1) - st.r0] >= st.sl_size[(c2) - st.r0] ? c1 : c2;
      int32_t child = root == c1 ? c2 : c1;
      st.sl_parent[(child) - st.r0] = root;
      st.sl_head[(root) - st.r0] = head; st.sl_tail[(root) - st.r0] = tail; st.sl_size[(root) - st.r0] = size; st.sl_key[(root) - st.r0] = key;
    } else {
      st.rd_next[(st.sl_tail[(c1) - st.r0]) - st.r0] = rid2; st.rd_next[(rid2) - st.r0] = -1;
      st.sl_tail[(c1) - st.r0] = rid2; st.sl_size[(c1) - st.r0] += 1; st.rd_slot[(rid2) - st.r0] = c1;
    }
  } else if (c2 >= 0) {
    st.rd_next[(st.sl_tail[(c2) - st.r0]) - st.r0] = rid1; st.rd_next[(rid1) - st.r0] = -1;
    st.sl_tail[(c2) - st.r0] = rid1; st.sl_size[(c2) - st.r0] += 1; st.rd_slot[(rid1) - st.r0] = c2;
  } else {
    int32_t s = rid1;
    st.sl_parent[(s) - st.r0] = s; st.sl_head[(s) - st.r0] = rid1; st.sl_tail[(s) - st.r0] = rid2; st.sl_size[(s) - st.r0] = 2; st.sl_key[(s) - st.r0] = new_key;
    st.rd_next[(rid1) - st.r0] = rid2; st.rd_next[(rid2) - st.r0] = -1;
    st.rd_slot[(rid1) - st.r0] = s; st.rd_slot[(rid2) - st.r0] = s;
  }
}

// Fibonacci hashing: one multiply, the table index is the top bits
PCB_HD uint32_t hash_i32(int32_t k, int32_t shift) { return ((uint32_t)k * 0x9E3779B1u) >> shift; }

struct Table {
  GenoSlot *slots; int32_t mask; int32_t shift; int32_t *touched; int32_t n_touched;
};
PCB_HD void table_init(Table &t, GenoSlot *slots, int32_t mask, int32_t *touched) {
  t.slots = slots; t.mask = mask; t.touched = touched; t.n_touched = 0;
  int32_t bits = 0;
  while ((1 << bits) <= mask) bits++;          // mask = 2^bits - 1, bits >= 1
  t.shift = 32 - bits;
}
PCB_HD GenoSlot *table_slot(Table &t, int32_t key) {
  uint32_t i = hash_i32(key, t.shift);
  for (;;) {
    GenoSlot *s = &t.slots[i];
    if (s->key == key + 1) return s;
    if (s->key == 0) {
      s->key = key + 1;                          // the other fields of an empty slot are already zero
      t.touched[t.n_touched++] = (int32_t)i;
      return s;
    }
    i = (i + 1) & (uint32_t)t.mask;
  }
}
PCB_HD const GenoSlot *table_find(const Table &t, int32_t key) {
  uint32_t i = hash_i32(key, t.shift);
  for (;;) {
    const GenoSlot *s = &t.slots[i];
    if (s->key == key + 1) return s;
    if (s->key == 0) return nullptr;
    i = (i + 1) & (uint32_t)t.mask;
  }
}
PCB_HD void table_clear(Table &t) {
  for (int32_t i = 0; i < t.n_touched; i++) {
    GenoSlot *s = &t.slots[t.touched[i]];
    s->key = 0; s->obs1 = s->obs2 = 0; s->sum1 = s->sum2 = 0;
  }
  t.n_touched = 0;
}

// a distance the pair list does not hold, scored on the spot (kept out of line: it is rare and two Myers loops long)
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
int32_t on_demand_dist(const uint64_t *planes, const int32_t *blen, int32_t compute_k, int32_t wide, int32_t a, int32_t b) {
  int32_t m = blen[a], n = blen[b];
  if (m > 64 || n > 64) return -1;               // an opaque barcode (csrc/common.cuh: PCB_LEN_OPAQUE) is near nothing
  int32_t dl = m > n ? m - n : n - m;
  if (dl > compute_k) return -1;
  int32_t d = wide ? myers64(planes[2 * a], planes[2 * a + 1], m, planes[2 * b], planes[2 * b + 1], n)
                   : myers32((uint32_t)planes[2 * a], (uint32_t)planes[2 * a + 1], m, (uint32_t)planes[2 * b], (uint32_t)planes[2 * b + 1], n);
  return d <= compute_k ? d : -1;
}

// distance between two buckets as edistLookup holds it (:236,:266,:294); -1 = absent (inf)
PCB_HD int32_t lookup_dist(const MergeCtx &c, int32_t a, int32_t b) {
  if (a == b) return 0;
  if (c.bucket_gid) { a = c.bucket_gid[a]; b = c.bucket_gid[b]; }
  if (c.adj_forward && a > b) { const int32_t t = a; a = b; b = t; }
  int32_t lo = c.adj_ptr[a], hi = c.adj_ptr[a + 1];
  while (lo < hi) {
    int32_t mid = (lo + hi) >> 1;
    if (c.adj_nbr[mid] < b) lo = mid + 1; else hi = mid;
  }
  if (lo < c.adj_ptr[a + 1] && c.adj_nbr[lo] == b) return c.adj_d[lo];
  if (c.compute_k < 0) return -1;
  return on_demand_dist(c.planes, c.blen, c.compute_k, c.wide, a, b);
}

// SEED CLUSTERING of one bucket (:385-400)
PCB_HD void seed_bucket(const MergeCtx &c, const MergeState &st, int32_t b, Table &tab) {
  int32_t lo = c.bucket_ptr[b], n = c.bucket_ptr[b + 1] - lo;
  if (n < 2) return;
  const int32_t *rids = c.bucket_reads ? c.bucket_reads + lo : nullptr;     // null: reads are numbered lo, lo+1, ...
#define PCB_RID(i) (rids ? rids[i] : lo + (i))
  // The usual bucket: every read carries the same calls (same mutations, same order, all Q > 0).  Then every
  // mutation is kept (seen in n > 1 reads), every pair has Jaccard g/g, and the loop below would link (r1, r0),
  // then (r2, r0), (r3, r0), ...: one cluster r1 -> r0 -> r2 -> ... -> r(n-1), created first in this bucket.
  {
    const int32_t r0 = PCB_RID(0);
    const int32_t g0 = c.geno_ptr[r0], g = c.geno_ptr[r0 + 1] - g0;
    bool uniform = true;
    for (int32_t e = 0; e < g && uniform; e++) uniform = c.geno_q[g0 + e] > 0;
    for (int32_t i = 1; i < n && uniform; i++) {
      const int32_t ri = PCB_RID(i), gi = c.geno_ptr[ri];
      uniform = c.geno_ptr[ri + 1] - gi == g;
      for (int32_t e = 0; e < g && uniform; e++) uniform = c.geno_mut[gi + e] == c.geno_mut[g0 + e] && c.geno_q[gi + e] > 0;
    }
    if (uniform) {
      if (!(g == 0 || (1.0 > c.prm.minJaccard && g >= c.prm.minMatches))) return;     // no pair of this bucket links
      const int32_t r1 = PCB_RID(1), s = r1;
      int32_t prev = r0;
      st.rd_next[(r1) - st.r0] = r0;
      st.rd_slot[(r0) - st.r0] = s; st.rd_slot[(r1) - st.r0] = s;
      for (int32_t i = 2; i < n; i++) {
        const int32_t ri = PCB_RID(i);
        st.rd_next[(prev) - st.r0] = ri; st.rd_slot[(ri) - st.r0] = s;
        prev = ri;
      }
      st.rd_next[(prev) - st.r0] = -1;
      st.sl_parent[(s) - st.r0] = s; st.sl_head[(s) - st.r0] = r1; st.sl_tail[(s) - st.r0] = prev; st.sl_size[(s) - st.r0] = n;
      st.sl_key[(s) - st.r0] = (int64_t)b << 32;
      return;
    }
  }
  // varMatrix + filterSeqError (:338-356): per mutation #reads with Q>0 and sum of Q over the bucket
  for (int32_t i = 0; i < n; i++)
    for (int32_t e = c.geno_ptr[PCB_RID(i)]; e < c.geno_ptr[PCB_RID(i) + 1]; e++) {
      GenoSlot *s = table_slot(tab, c.geno_mut[e]);
      int32_t q = c.geno_q[e];
      if (q > 0) s->obs1++;
      s->sum1 += q;
    }
  int32_t n_created = 0;
  for (int32_t i = 1; i < n; i++) {
    int32_t ri = PCB_RID(i);
    int32_t ci = -1;                                             // read i is in no cluster when its row of pairs starts
    for (int32_t j = 0; j < i; j++) {
      int32_t rj = PCB_RID(j);
      if (ci >= 0 && ci == cluster_of(c, st, rj)) continue;     // addLink would be a no-op
      // calcJaccard on presence over the kept mutations (:358-367)
      int32_t ni = 0, nj = 0, inters = 0;
      for (int32_t e = c.geno_ptr[ri]; e < c.geno_ptr[ri + 1]; e++) {
        if (c.geno_q[e] <= 0) continue;
        const GenoSlot *s = table_find(tab, c.geno_mut[e]);
        if (!(s->obs1 > 1 || s->sum1 > c.prm.minQual)) continue;
        ni++;
        for (int32_t f = c.geno_ptr[rj]; f < c.geno_ptr[rj + 1]; f++)
          if (c.geno_mut[f] == c.geno_mut[e]) { if (c.geno_q[f] > 0) inters++; break; }
      }
      for (int32_t f = c.geno_ptr[rj]; f < c.geno_ptr[rj + 1]; f++) {
        if (c.geno_q[f] <= 0) continue;
        const GenoSlot *s = table_find(tab, c.geno_mut[f]);
        if (s->obs1 > 1 || s->sum1 > c.prm.minQual) nj++;
      }
      int32_t uni = ni + nj - inters;
      bool ok = uni == 0 || ((double)inters / (double)uni > c.prm.minJaccard && inters >= c.prm.minMatches);
      if (ok) {
        bool fresh = st.rd_slot[(ri) - st.r0] < 0 && st.rd_slot[(rj) - st.r0] < 0;
        add_link(c, st, ri, rj, ((int64_t)b << 32) | (int64_t)n_created);
        if (fresh) n_created++;
        ci = cluster_of(c, st, ri);
        // during the seed step a cluster only holds reads of this bucket: once it holds all of 0..i,
        // every remaining pair (i, j) is a no-op
        if (st.sl_size[(ci) - st.r0] == i + 1) break;
      }
    }
  }
  table_clear(tab);
#undef PCB_RID
}

// accumulate one side of the genotype rule; side 0 -> (obs1,sum1), side 1 -> (obs2,sum2)
PCB_HD void accumulate_side(const MergeCtx &c, const MergeState &st, Table &tab, int32_t cl, int32_t single, int side) {
  int32_t r = cl >= 0 ? st.sl_head[(cl) - st.r0] : single;
  while (r >= 0) {
    for (int32_t e = c.geno_ptr[r]; e < c.geno_ptr[r + 1]; e++) {
      GenoSlot *s = table_slot(tab, c.geno_mut[e]);
      int32_t q = c.geno_q[e];
      if (side == 0) { if (q > 0) s->obs1++; s->sum1 += q; }
      else           { if (q > 0) s->obs2++; s->sum2 += q; }
    }
    r = cl >= 0 ? st.rd_next[(r) - st.r0] : -1;
  }
}

// one visit of the main loop body (:419-455) for the ordered read pair (rid1, rid2)
// returns 1 if the pair was linked
// `cache_root`: the cluster whose genotype sums currently sit in side 1 of the table (-1: none).  A
// cluster's sums only change when it merges, and the merged sums are exactly side 1 + side 2 of the test
// that allowed the merge, so the usual "many satellites join one big cluster" sequence accumulates the
// big side once instead of once per test.
PCB_HD int visit_pair(const MergeCtx &c, const MergeState &st, Table &tab, int32_t *runs, int32_t rid1, int32_t rid2, int32_t ed,
                      int32_t c1, int32_t c2, int32_t &cache_root) {
  int32_t size1 = c1 >= 0 ? st.sl_size[(c1) - st.r0] : 1, size2 = c2 >= 0 ? st.sl_size[(c2) - st.r0] : 1;
  int32_t mx = size1 > size2 ? size1 : size2, mn = size1 > size2 ? size2 : size1;
  if ((int64_t)mx < ((int64_t)mn << ed)) return 0;                         // :435-440
  // transitive rule (:442-445): bucket runs of cluster 2, then every bucket run of cluster 1 against them
  int32_t n_runs = 0, last = -1;
  for (int32_t r = c2 >= 0 ? st.sl_head[(c2) - st.r0] : rid2; r >= 0; r = c2 >= 0 ? st.rd_next[(r) - st.r0] : -1) {
    int32_t b = c.bucket_of_read[r];
    if (b != last) { runs[n_runs++] = b; last = b; }
  }
  int32_t maxED = 0; last = -1;
  for (int32_t r = c1 >= 0 ? st.sl_head[(c1) - st.r0] : rid1; r >= 0; r = c1 >= 0 ? st.rd_next[(r) - st.r0] : -1) {
    int32_t a = c.bucket_of_read[r];
    if (a == last) continue;
    last = a;
    for (int32_t k = 0; k < n_runs; k++) {
      int32_t d = lookup_dist(c, a, runs[k]);
      if (d < 0) return 0;                                                 // inf
      if (d > maxED) maxED = d;
    }
  }
  if (maxED > 2 * ed) return 0;
  // genotype rule (:447-452); the test is symmetric in the two sides, so whichever is cached stays side 1
  if (c1 >= 0 && c1 == cache_root) {
    accumulate_side(c, st, tab, c2, rid2, 1);
  } else if (c2 >= 0 && c2 == cache_root) {
    accumulate_side(c, st, tab, c1, rid1, 1);
  } else {
    table_clear(tab);
    if (c1 < 0 && c2 >= 0) { accumulate_side(c, st, tab, c2, rid2, 0); accumulate_side(c, st, tab, c1, rid1, 1); cache_root = c2; }
    else                   { accumulate_side(c, st, tab, c1, rid1, 0); accumulate_side(c, st, tab, c2, rid2, 1); cache_root = c1; }
  }
  int32_t inters = 0, uni = 0;
  for (int32_t i = 0; i < tab.n_touched; i++) {
    const GenoSlot *s = &tab.slots[tab.touched[i]];
    if (!(s->obs1 + s->obs2 > 1 || s->sum1 + s->sum2 > c.prm.minQual)) continue;   // filterSeqError
    bool p1 = s->sum1 > 0, p2 = s->sum2 > 0;
    if (p1 && p2) inters++;
    if (p1 || p2) uni++;
  }
  bool link = uni == 0 || (double)inters / (double)uni > c.prm.minJaccard;
  if (link) add_link(c, st, rid1, rid2, -1);
  if (cache_root < 0) {
    table_clear(tab);                                  // side 1 was a lone read: nothing worth keeping
  } else {
    for (int32_t i = 0; i < tab.n_touched; i++) {      // fold (merged) or drop (rejected) side 2
      GenoSlot *s = &tab.slots[tab.touched[i]];
      if (link) { s->obs1 += s->obs2; s->sum1 += s->sum2; }
      s->obs2 = 0; s->sum2 = 0;
    }
    if (link) cache_root = cluster_of(c, st, rid1);
  }
  return link ? 1 : 0;
}

// consensus of one id column (barcode or uptag) over a member list (:515-532)
PCB_HD int32_t consensus_id(const MergeCtx &c, const MergeState &st, Table &tab, const int32_t *ids, int32_t head, int32_t size, int32_t *err) {
  int32_t best = -1, bestCnt = 0, nTop = 0;
  {                                      // all members agree (the usual case): that id is the unique top
    int32_t first = ids[head];
    bool same = first >= 0;
    for (int32_t r = st.rd_next[(head) - st.r0]; same && r >= 0; r = st.rd_next[(r) - st.r0]) same = ids[r] == first;
    if (same) return first;
  }
  for (int32_t r = head; r >= 0; r = st.rd_next[(r) - st.r0]) {
    int32_t id = ids[r];
    if (id < 0) { *err = 1; id = 0x7ffffffe; }
    table_slot(tab, id)->obs1++;
  }
  for (int32_t i = 0; i < tab.n_touched; i++) {
    const GenoSlot *s = &tab.slots[tab.touched[i]];
    if (s->obs1 > bestCnt) { bestCnt = s->obs1; best = s->key - 1; nTop = 1; }
    else if (s->obs1 == bestCnt) nTop++;
  }
  table_clear(tab);
  if (nTop == 1) return best;
  if (size == 2) return ids[head];
  return PCB_NEEDS_ALIGNMENT;
}

// The replay of one component in three phases.  They only communicate through the cluster state (MergeState),
// so they can run back to back in one thread (process_component, host emulation) or as three kernels whose
// code each fits the instruction cache and whose threads are all in the same loop nest (merge.cu).

// ---- phase 1: seed clustering, bucket by bucket
PCB_HD void component_seed(const MergeCtx &c, const MergeState &st, int32_t comp) {
  Table tab;
  table_init(tab, st.slots, st.slot_mask, st.lists);     // lists: first half touched slots, second half bucket runs
  int32_t b_lo = c.comp_ptr[comp], b_hi = c.comp_ptr[comp + 1];
  for (int32_t bi = b_lo; bi < b_hi; bi++) seed_bucket(c, st, c.comp_buckets ? c.comp_buckets[bi] : bi, tab);
}

// ---- phase 2: main clustering: the component's rows, pass by pass, in visiting order
PCB_HD void component_main(const MergeCtx &c, const MergeState &st, int32_t comp) {
  if (c.cpair_ptr[comp] == c.cpair_ptr[comp + 1]) return;
  Table tab;
  table_init(tab, st.slots, st.slot_mask, st.lists);
  int32_t *runs = st.lists + st.list_cap;
  int32_t n_links = 0, n_tests = 0, cache_root = -1;
  int32_t memo_a[4], memo_b[4], memo_n = 0, memo_w = 0;
  for (int32_t p = c.cpair_ptr[comp]; p < c.cpair_ptr[comp + 1]; p++) {
    int32_t bq = c.cpair_q[p], br = c.cpair_r[p], ed = c.cpair_ed[p];
    for (int32_t x = c.bucket_ptr[bq]; x < c.bucket_ptr[bq + 1]; x++) {
      int32_t rx = c.bucket_reads ? c.bucket_reads[x] : x;
      for (int32_t y = c.bucket_ptr[br]; y < c.bucket_ptr[br + 1]; y++) {
        int32_t ry = c.bucket_reads ? c.bucket_reads[y] : y;
        int32_t rid1 = rx, rid2 = ry;                                       // makePID order (:152-157)
        if (!(c.lexrank[rx] < c.lexrank[ry])) { rid1 = ry; rid2 = rx; }
        int32_t c1 = cluster_of(c, st, rid1), c2 = cluster_of(c, st, rid2);
        if (c1 >= 0 && c1 == c2) continue;                                  // :424
        int32_t ida = c1 >= 0 ? c1 : ~rid1, idb = c2 >= 0 ? c2 : ~rid2;
        bool seen = false;
        for (int32_t k = 0; k < memo_n; k++) if (memo_a[k] == ida && memo_b[k] == idb) { seen = true; break; }
        if (seen) continue;
        n_tests++;
        if (visit_pair(c, st, tab, runs, rid1, rid2, ed, c1, c2, cache_root)) { n_links++; memo_n = 0; memo_w = 0; }
        else { memo_a[memo_w] = ida; memo_b[memo_w] = idb; memo_w = (memo_w + 1) & 3; if (memo_n < 4) memo_n++; }
      }
    }
    memo_n = 0; memo_w = 0;   // ids of a failed pair are only compared within one row
  }

  table_clear(tab);                                    // drop whatever the main loop left cached
#if defined(__CUDA_ARCH__)
  if (n_links) atomicAdd(&c.status[2], n_links);
  if (n_tests) atomicAdd(&c.status[3], n_tests);
#else
  c.status[2] += n_links; c.status[3] += n_tests;
#endif
}

// Where the calls of a component's rows go: its fixed region (call_off), or as many entries as its reads carry -- an
// upper bound of its calls -- taken from the shared cursor (regions are then laid out in the order the components finish).
PCB_HD int64_t component_call_base(const MergeCtx &c, int32_t comp) {
  if (c.call_off) return c.call_off[comp];
  int64_t ne = 0;
  for (int32_t bi = c.comp_ptr[comp]; bi < c.comp_ptr[comp + 1]; bi++) {
    const int32_t b = c.comp_buckets ? c.comp_buckets[bi] : bi;
    for (int32_t x = c.bucket_ptr[b]; x < c.bucket_ptr[b + 1]; x++) {
      const int32_t r = c.bucket_reads ? c.bucket_reads[x] : x;
      ne += c.geno_ptr[r + 1] - c.geno_ptr[r];
    }
  }
#if defined(__CUDA_ARCH__)
  return (int64_t)atomicAdd((unsigned long long *)c.call_cursor, (unsigned long long)ne);
#else
  const int64_t at = *c.call_cursor;
  *c.call_cursor = at + ne;
  return at;
#endif
}

// ---- phase 3: consensus + rows (:464-477, :515-537, :546-567)
PCB_HD void component_consensus(const MergeCtx &c, const MergeState &st, int32_t comp) {
  Table tab;
  table_init(tab, st.slots, st.slot_mask, st.lists);
  int32_t b_lo = c.comp_ptr[comp], b_hi = c.comp_ptr[comp + 1];
  int64_t call_pos = component_call_base(c, comp);
  int32_t err = 0;
  for (int32_t bi = b_lo; bi < b_hi; bi++) {
    int32_t b = c.comp_buckets ? c.comp_buckets[bi] : bi;
    for (int32_t x = c.bucket_ptr[b]; x < c.bucket_ptr[b + 1]; x++) {
      int32_t r = c.bucket_reads ? c.bucket_reads[x] : x;
      int32_t cl = cluster_of(c, st, r);
      if (cl < 0) {
        // singleton (:560-567): calls with Q >= minQual in input order
        c.read_root[r] = r; c.read_pos[r] = 0; c.row_size[r] = 1; c.row_key[r] = -1;
        c.row_bc[r] = c.bc_id[r]; c.row_up[r] = c.up_id ? c.up_id[r] : c.bc_id[r];
        c.row_call_off[r] = call_pos;
        int32_t n = 0;
        for (int32_t e = c.geno_ptr[r]; e < c.geno_ptr[r + 1]; e++)
          if (c.geno_q[e] >= c.prm.minQual) { c.call_mut[call_pos + n] = c.geno_mut[e]; n++; }
        c.row_call_n[r] = n; call_pos += n;
        continue;
      }
      int32_t head = st.sl_head[(cl) - st.r0];
      if (head != r) continue;                     // each cluster is emitted when its head comes by
      int32_t size = st.sl_size[(cl) - st.r0], pos = 0;
      for (int32_t m = head; m >= 0; m = st.rd_next[(m) - st.r0]) {
        c.read_root[m] = head; c.read_pos[m] = pos++;
        if (m != head) c.row_size[m] = 0;
        // genotype vote (:468-470): Q where the read has the key, -minQual where it has not
        for (int32_t e = c.geno_ptr[m]; e < c.geno_ptr[m + 1]; e++) {
          GenoSlot *s = table_slot(tab, c.geno_mut[e]);
          s->obs1++; s->sum1 += c.geno_q[e];
        }
      }
      c.row_size[head] = size; c.row_key[head] = st.sl_key[(cl) - st.r0];
      int32_t n = 0;
      int32_t *calls = c.call_mut + call_pos;
      for (int32_t i = 0; i < tab.n_touched; i++) {
        const GenoSlot *s = &tab.slots[tab.touched[i]];
        if (s->sum1 - (int64_t)c.prm.minQual * (int64_t)(size - s->obs1) > 0) {
          int32_t v = s->key - 1, k = n - 1;       // insertion sort by the (digit key, string) rank (:475)
          while (k >= 0 && c.mut_rank[calls[k]] > c.mut_rank[v]) { calls[k + 1] = calls[k]; k--; }
          calls[k + 1] = v; n++;
        }
      }
      table_clear(tab);
      c.row_call_off[head] = call_pos; c.row_call_n[head] = n; call_pos += n;
      c.row_bc[head] = consensus_id(c, st, tab, c.bc_id, head, size, &err);
      int32_t uerr = 0;
      c.row_up[head] = c.up_id ? consensus_id(c, st, tab, c.up_id, head, size, &uerr) : c.row_bc[head];
      if (uerr) { c.status[0] = 3; c.status[1] = head; }     // clustered read without uptag: KeyError in the reference
    }
  }
}

PCB_HD void process_component(const MergeCtx &c, const MergeState &st, int32_t comp) {
  component_seed(c, st, comp);
  component_main(c, st, comp);
  component_consensus(c, st, comp);
}

}  // namespace pcb
