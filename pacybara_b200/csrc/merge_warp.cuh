// merge_warp.cuh -- warp-cooperative version of the per-component replay.
//
// Same algorithm, same visiting order and the same outputs as process_component() in merge_core.h
// (which stays the host-testable statement of the logic and the PCB_REPLAY=thread fallback); here ONE
// WARP runs one unit of the replay with uniform control flow:
//   * everything order dependent (addLink, list surgery, memo, the i/j and read-pair visiting order)
//     is decided uniformly by all lanes and written by lane 0;
//   * everything order independent runs across the lanes: genotype sums (atomics into the scratch
//     table, one lane per genotype entry), the Jaccard tests of read i against reads j < i (one lane
//     per j), the same-cluster / memo screening of a row's read pairs (one lane per pair), the
//     transitive-rule lookups (one lane per bucket pair), the evaluation / folding / clearing of the
//     touched table slots, the consensus vote and its ordering.
// The units -- seed step of one bucket (w_seed_bucket), one row of the main loop (w_run_row), one output
// row (w_emit_singleton, w_cluster_vote / w_cluster_write) -- are composed twice: by process_component_warp
// below (one warp walks a whole component) and by merge_huge.cuh (many warps inside ONE giant component).
// Written against simt.h so that tests/host_emul runs this very source on the CPU (unit tests only).
#pragma once
#include "merge_core.h"
#include "simt.h"

namespace pcb {

#define PCB_FULL 0xFFFFFFFFu

struct WTable {
  GenoSlot *slots;
  int32_t mask, shift;
  int32_t *touched;
  int32_t *n_touched;       // shared counter (scratch)
};

// the table and the run list of a warp from its scratch (MergeState::slots / lists); the counter is zeroed
SIMT_FN WTable w_make_table(const MergeState &st, int32_t *&runs, int lane) {
  WTable t;
  t.slots = st.slots; t.mask = st.slot_mask;
  { int32_t bits = 0; while ((1 << bits) <= st.slot_mask) bits++; t.shift = 32 - bits; }
  t.n_touched = st.lists;                                  // lists[0]: shared counter (kept zero between uses)
  t.touched = st.lists + 4;
  runs = st.lists + 4 + st.list_cap;
  if (lane == 0) *t.n_touched = 0;
  simt::sync(PCB_FULL);
  return t;
}

SIMT_FN int32_t w_find_ro(const MergeState &st, int32_t s) {     // no path halving: safe to run on all lanes
  for (;;) {
    int32_t p = ((volatile int32_t *)st.sl_parent)[s - st.r0];
    if (p == s) return s;
    s = p;
  }
}
SIMT_FN int32_t w_cluster_of(const MergeState &st, int32_t r) {
  int32_t s = ((volatile int32_t *)st.rd_slot)[r - st.r0];
  return s < 0 ? -1 : w_find_ro(st, s);
}

// concurrent insert-or-find; empty slots are all zero, so a claimed slot needs no initialisation
SIMT_FN GenoSlot *w_table_slot(const WTable &t, int32_t key) {
  uint32_t i = hash_i32(key, t.shift);
  for (;;) {
    GenoSlot *s = &t.slots[i];
    int32_t k = *(volatile int32_t *)&s->key;
    if (k == key + 1) return s;
    if (k == 0) {
      int32_t old = simt::atomic_cas(&s->key, 0, key + 1);
      if (old == 0) { t.touched[simt::atomic_add(t.n_touched, 1)] = (int32_t)i; return s; }
      if (old == key + 1) return s;
    }
    i = (i + 1) & (uint32_t)t.mask;
  }
}
SIMT_FN const GenoSlot *w_table_find(const WTable &t, int32_t key) {
  uint32_t i = hash_i32(key, t.shift);
  for (;;) {
    const GenoSlot *s = &t.slots[i];
    int32_t k = *(volatile const int32_t *)&s->key;
    if (k == key + 1) return s;
    if (k == 0) return nullptr;
    i = (i + 1) & (uint32_t)t.mask;
  }
}

SIMT_FN void w_table_clear(const WTable &t, int lane) {
  simt::sync(PCB_FULL);
  int32_t n = *(volatile int32_t *)t.n_touched;
  for (int32_t i = lane; i < n; i += 32) {
    GenoSlot *s = &t.slots[t.touched[i]];
    s->key = 0; s->obs1 = 0; s->obs2 = 0; s->sum1 = 0; s->sum2 = 0;
  }
  simt::sync(PCB_FULL);
  if (lane == 0) *t.n_touched = 0;
  simt::sync(PCB_FULL);
}

// entries of read r into side 0 / 1 of the table, one lane per entry
SIMT_FN void w_accumulate_read(const MergeCtx &c, const WTable &t, int32_t r, int side, int lane) {
  int32_t lo = c.geno_ptr[r], hi = c.geno_ptr[r + 1];
  for (int32_t e = lo + lane; e < hi; e += 32) {
    GenoSlot *s = w_table_slot(t, c.geno_mut[e]);
    int32_t q = c.geno_q[e];
    if (side == 0) { if (q > 0) simt::atomic_add(&s->obs1, 1); simt::atomic_add64(&s->sum1, q); }
    else           { if (q > 0) simt::atomic_add(&s->obs2, 1); simt::atomic_add64(&s->sum2, q); }
  }
}
SIMT_FN void w_accumulate_side(const MergeCtx &c, const MergeState &st, const WTable &t, int32_t cl, int32_t single, int side, int lane) {
  int32_t r = cl >= 0 ? st.sl_head[cl - st.r0] : single;
  while (r >= 0) {                                                   // uniform walk of the member list
    w_accumulate_read(c, t, r, side, lane);
    r = cl >= 0 ? st.rd_next[r - st.r0] : -1;
  }
  simt::sync(PCB_FULL);
}

// calcJaccard on presence over the kept mutations (:358-367) and the seed acceptance rule (:398); the table
// holds the bucket-wide (#reads with Q>0, sum Q) per mutation that filterSeqError needs
SIMT_NOINLINE bool w_seed_pair_ok(const MergeCtx &c, const WTable &t, int32_t ri, int32_t rj) {
  int32_t ni = 0, nj = 0, inters = 0;
  for (int32_t e = c.geno_ptr[ri]; e < c.geno_ptr[ri + 1]; e++) {
    if (c.geno_q[e] <= 0) continue;
    const GenoSlot *s = w_table_find(t, c.geno_mut[e]);
    if (!(s->obs1 > 1 || s->sum1 > c.prm.minQual)) continue;
    ni++;
    for (int32_t f = c.geno_ptr[rj]; f < c.geno_ptr[rj + 1]; f++)
      if (c.geno_mut[f] == c.geno_mut[e]) { if (c.geno_q[f] > 0) inters++; break; }
  }
  for (int32_t f = c.geno_ptr[rj]; f < c.geno_ptr[rj + 1]; f++) {
    if (c.geno_q[f] <= 0) continue;
    const GenoSlot *s = w_table_find(t, c.geno_mut[f]);
    if (s->obs1 > 1 || s->sum1 > c.prm.minQual) nj++;
  }
  const int32_t uni = ni + nj - inters;
  return uni == 0 || ((double)inters / (double)uni > c.prm.minJaccard && inters >= c.prm.minMatches);
}

// SEED CLUSTERING of one bucket (:385-400)
SIMT_FN void w_seed_bucket(const MergeCtx &c, const MergeState &st, const WTable &t, int32_t b, int lane) {
  const int32_t lo = c.bucket_ptr[b], n = c.bucket_ptr[b + 1] - lo;
  if (n < 2) return;
  const int32_t *rids = c.bucket_reads ? c.bucket_reads + lo : nullptr;
#define PCB_RID(i) (rids ? rids[i] : lo + (i))
  // per mutation: #reads with Q>0 and sum of Q over the bucket (varMatrix + filterSeqError, :338-356)
  for (int32_t i = 0; i < n; i++) w_accumulate_read(c, t, PCB_RID(i), 0, lane);
  simt::sync(PCB_FULL);
  int32_t n_created = 0;
  for (int32_t i = 1; i < n; i++) {
    const int32_t ri = PCB_RID(i);
    int32_t ci = -1;
    bool done = false;
    // Usual case: reads 0..i-1 already form one cluster and read i is compatible with read 0 -- then the row
    // ends after its first pair.  Test that pair on its own (uniformly) before fanning the rest out.
    int32_t j_first = 0;
    {
      const int32_t r0 = PCB_RID(0);
      const int32_t c0 = w_cluster_of(st, r0);
      if (c0 >= 0 && ((volatile int32_t *)st.sl_size)[c0 - st.r0] == i) {
        if (w_seed_pair_ok(c, t, ri, r0)) {
          simt::sync(PCB_FULL);                            // every lane has looked at the state lane 0 is about to change
          if (lane == 0) add_link(c, st, ri, r0, -1);      // joins cluster c0 (no new cluster id), which then holds 0..i
          simt::sync(PCB_FULL);
          continue;
        }
        j_first = 1;                                       // pair (i, 0) is done and was rejected
      }
    }
    for (int32_t j0 = j_first; j0 < i && !done; j0 += 32) {
      const int32_t j = j0 + lane;
      const bool act = j < i;
      const int32_t rj = act ? PCB_RID(j) : -1;
      bool ok = false;
      if (act && !(ci >= 0 && w_cluster_of(st, rj) == ci)) {
        ok = w_seed_pair_ok(c, t, ri, rj);
      }
      uint32_t mask = simt::ballot(PCB_FULL, ok);
      while (mask) {                                                 // apply in ascending j, as the reference does
        const int l = simt::ffs(mask) - 1;
        mask &= mask - 1;
        const int32_t rjl = simt::shfl(PCB_FULL, rj, l);
        if (ci >= 0 && w_cluster_of(st, rjl) == ci) continue;        // became a no-op through an earlier link of this row
        simt::sync(PCB_FULL);
        if (lane == 0) {
          bool fresh = st.rd_slot[ri - st.r0] < 0 && st.rd_slot[rjl - st.r0] < 0;
          add_link(c, st, ri, rjl, ((int64_t)b << 32) | (int64_t)n_created);
          if (fresh) n_created++;
        }
        simt::sync(PCB_FULL);
        n_created = simt::shfl(PCB_FULL, n_created, 0);
        ci = w_cluster_of(st, ri);
        if (((volatile int32_t *)st.sl_size)[ci - st.r0] == i + 1) { done = true; break; }   // all of 0..i in one cluster
      }
    }
  }
  w_table_clear(t, lane);
#undef PCB_RID
}

// one visit of the main loop body (:419-455); uniform across the warp.  Returns 1 if linked.
SIMT_FN int w_visit_pair(const MergeCtx &c, const MergeState &st, const WTable &t, int32_t *runs, int32_t rid1, int32_t rid2, int32_t ed,
                         int32_t c1, int32_t c2, int32_t &cache_root, int lane) {
  const int32_t size1 = c1 >= 0 ? st.sl_size[c1 - st.r0] : 1, size2 = c2 >= 0 ? st.sl_size[c2 - st.r0] : 1;
  const int32_t mx = size1 > size2 ? size1 : size2, mn = size1 > size2 ? size2 : size1;
  if ((int64_t)mx < ((int64_t)mn << ed)) return 0;                                 // :435-440
  // transitive rule (:442-445)
  int32_t n_runs = 0, last = -1;
  for (int32_t r = c2 >= 0 ? st.sl_head[c2 - st.r0] : rid2; r >= 0; r = c2 >= 0 ? st.rd_next[r - st.r0] : -1) {
    int32_t b = c.bucket_of_read[r];
    if (b != last) { if (lane == 0) runs[n_runs] = b; n_runs++; last = b; }
  }
  simt::sync(PCB_FULL);
  int32_t maxED = 0;
  bool missing = false;
  last = -1;
  for (int32_t r = c1 >= 0 ? st.sl_head[c1 - st.r0] : rid1; r >= 0; r = c1 >= 0 ? st.rd_next[r - st.r0] : -1) {
    int32_t a = c.bucket_of_read[r];
    if (a == last) continue;
    last = a;
    for (int32_t k = lane; k < n_runs; k += 32) {
      int32_t d = lookup_dist(c, a, runs[k]);
      if (d < 0) missing = true; else if (d > maxED) maxED = d;
    }
  }
  if (simt::ballot(PCB_FULL, missing) != 0u) return 0;                              // inf
  maxED = simt::reduce_max(PCB_FULL, maxED);
  if (maxED > 2 * ed) return 0;
  // genotype rule (:447-452); symmetric in the two sides, the cached cluster stays side 1
  if (c1 >= 0 && c1 == cache_root) {
    w_accumulate_side(c, st, t, c2, rid2, 1, lane);
  } else if (c2 >= 0 && c2 == cache_root) {
    w_accumulate_side(c, st, t, c1, rid1, 1, lane);
  } else {
    w_table_clear(t, lane);
    if (c1 < 0 && c2 >= 0) { w_accumulate_side(c, st, t, c2, rid2, 0, lane); w_accumulate_side(c, st, t, c1, rid1, 1, lane); cache_root = c2; }
    else                   { w_accumulate_side(c, st, t, c1, rid1, 0, lane); w_accumulate_side(c, st, t, c2, rid2, 1, lane); cache_root = c1; }
  }
  const int32_t n_t = *(volatile int32_t *)t.n_touched;
  int32_t inters = 0, uni = 0;
  for (int32_t i = lane; i < n_t; i += 32) {
    const GenoSlot *s = &t.slots[t.touched[i]];
    if (!(s->obs1 + s->obs2 > 1 || s->sum1 + s->sum2 > c.prm.minQual)) continue;   // filterSeqError
    bool p1 = s->sum1 > 0, p2 = s->sum2 > 0;
    if (p1 && p2) inters++;
    if (p1 || p2) uni++;
  }
  inters = simt::reduce_add(PCB_FULL, inters); uni = simt::reduce_add(PCB_FULL, uni);
  const bool link = uni == 0 || (double)inters / (double)uni > c.prm.minJaccard;
  if (link && lane == 0) {
    const int32_t id1 = c1 >= 0 ? c1 : rid1, id2 = c2 >= 0 ? c2 : rid2;            // (a lone read's cell is the read itself)
    const int32_t nnz = c.cell_nnz ? c.cell_nnz[id1] + c.cell_nnz[id2] : 0;
    add_link(c, st, rid1, rid2, -1);
    if (c.cell_nnz) c.cell_nnz[cluster_of(c, st, rid1)] = nnz;
  }
  simt::sync(PCB_FULL);
  if (cache_root < 0) {
    w_table_clear(t, lane);
  } else {
    for (int32_t i = lane; i < n_t; i += 32) {                                      // fold (merged) or drop (rejected) side 2
      GenoSlot *s = &t.slots[t.touched[i]];
      if (link) { s->obs1 += s->obs2; s->sum1 += s->sum2; }
      s->obs2 = 0; s->sum2 = 0;
    }
    simt::sync(PCB_FULL);
    if (link) cache_root = w_cluster_of(st, rid1);
  }
  return link ? 1 : 0;
}

// the memo of failed tests since the last merge (merge_core.h) and the cached side of the genotype table
struct RowMemo {
  int32_t a[4], b[4], n, w, cache_root;
};

// One row of the main loop (:408-455): the read pairs of buckets (bq, br) in visiting order, screened 32 at a time.
SIMT_FN void w_run_row(const MergeCtx &c, const MergeState &st, const WTable &t, int32_t *runs, int32_t bq, int32_t br, int32_t ed,
                       RowMemo &memo, int32_t &n_links, int32_t &n_tests, int lane) {
  const int32_t q_lo = c.bucket_ptr[bq], n1 = c.bucket_ptr[bq + 1] - q_lo;
  const int32_t r_lo = c.bucket_ptr[br], n2 = c.bucket_ptr[br + 1] - r_lo;
  const int64_t total = (int64_t)n1 * n2;
  int64_t at = 0;
  while (at < total) {
    const int64_t idx = at + lane;
    int32_t rid1 = -1, rid2 = -1, c1 = -1, c2 = -1;
    bool cand = false;
    if (idx < total) {
      const int32_t x = q_lo + (int32_t)(idx / n2), y = r_lo + (int32_t)(idx % n2);
      const int32_t rx = c.bucket_reads ? c.bucket_reads[x] : x, ry = c.bucket_reads ? c.bucket_reads[y] : y;
      rid1 = rx; rid2 = ry;                                                       // makePID order (:152-157)
      if (!(c.lexrank[rx] < c.lexrank[ry])) { rid1 = ry; rid2 = rx; }
      c1 = w_cluster_of(st, rid1); c2 = w_cluster_of(st, rid2);
      cand = !(c1 >= 0 && c1 == c2);                                              // :424
      if (cand) {
        const int32_t ida = c1 >= 0 ? c1 : ~rid1, idb = c2 >= 0 ? c2 : ~rid2;
        for (int32_t k = 0; k < memo.n; k++) if (memo.a[k] == ida && memo.b[k] == idb) { cand = false; break; }
      }
    }
    const uint32_t mask = simt::ballot(PCB_FULL, cand);
    if (mask == 0) { at += 32; continue; }
    const int l = simt::ffs(mask) - 1;                                            // the first pair the reference would test
    rid1 = simt::shfl(PCB_FULL, rid1, l); rid2 = simt::shfl(PCB_FULL, rid2, l);
    c1 = simt::shfl(PCB_FULL, c1, l); c2 = simt::shfl(PCB_FULL, c2, l);
    n_tests++;
    if (w_visit_pair(c, st, t, runs, rid1, rid2, ed, c1, c2, memo.cache_root, lane)) { n_links++; memo.n = 0; memo.w = 0; }
    else {
      memo.a[memo.w] = c1 >= 0 ? c1 : ~rid1; memo.b[memo.w] = c2 >= 0 ? c2 : ~rid2;
      memo.w = (memo.w + 1) & 3; if (memo.n < 4) memo.n++;
    }
    at += l + 1;                                                                  // state may have changed: re-screen what follows
  }
  memo.n = 0; memo.w = 0;                                                         // ids of a failed pair are only compared within one row
}

// consensus of one id column over a member list (:515-532); uniform.
SIMT_NOINLINE int32_t w_consensus_id(const MergeCtx &c, const MergeState &st, const WTable &t, const int32_t *ids, int32_t head, int32_t size,
                                     int32_t *err, int lane) {
  const int32_t first = ids[head];
  bool same = first >= 0;
  for (int32_t r = st.rd_next[head - st.r0]; same && r >= 0; r = st.rd_next[r - st.r0]) same = ids[r] == first;
  if (same) return first;
  // counts per id, then the reference's tie rules
  for (int32_t r = head; r >= 0; r = st.rd_next[r - st.r0]) {
    int32_t id = ids[r];
    if (id < 0) { *err = 1; id = 0x7ffffffe; }
    if (lane == 0) simt::atomic_add(&w_table_slot(t, id)->obs1, 1);
  }
  simt::sync(PCB_FULL);
  const int32_t n_t = *(volatile int32_t *)t.n_touched;
  int32_t best = -1, bestCnt = 0, nTop = 0;
  for (int32_t i = 0; i < n_t; i++) {                       // few distinct ids: uniform scan
    const GenoSlot *s = &t.slots[t.touched[i]];
    if (s->obs1 > bestCnt) { bestCnt = s->obs1; best = s->key - 1; nTop = 1; }
    else if (s->obs1 == bestCnt) nTop++;
  }
  w_table_clear(t, lane);
  if (nTop == 1) return best;
  if (size == 2) return ids[head];
  return PCB_NEEDS_ALIGNMENT;
}

// the row of an unclustered read (:560-567): calls with Q >= minQual in input order
SIMT_FN int32_t w_singleton_calls(const MergeCtx &c, int32_t r, int lane) {
  int32_t n = 0;
  for (int32_t e = c.geno_ptr[r] + lane; e < c.geno_ptr[r + 1]; e += 32) n += c.geno_q[e] >= c.prm.minQual ? 1 : 0;
  return simt::reduce_add(PCB_FULL, n);
}
SIMT_FN int32_t w_emit_singleton(const MergeCtx &c, int32_t r, int64_t call_pos, int lane) {
  const int32_t lo = c.geno_ptr[r], hi = c.geno_ptr[r + 1];
  int32_t n = 0;
  for (int32_t e0 = lo; e0 < hi; e0 += 32) {
    const int32_t e = e0 + lane;
    const bool keep = e < hi && c.geno_q[e] >= c.prm.minQual;
    const uint32_t m = simt::ballot(PCB_FULL, keep);
    if (keep) c.call_mut[call_pos + n + simt::popc(m & ((1u << lane) - 1u))] = c.geno_mut[e];
    n += simt::popc(m);
  }
  if (lane == 0) {
    c.read_root[r] = r; c.read_pos[r] = 0; c.row_size[r] = 1; c.row_key[r] = -1;
    c.row_bc[r] = c.bc_id[r]; c.row_up[r] = c.up_id ? c.up_id[r] : c.bc_id[r];
    c.row_call_off[r] = call_pos; c.row_call_n[r] = n;
  }
  return n;
}

// The row of cluster cl (:464-477, :515-537), in two steps so that the caller can place the calls once it knows how many
// there are.  Vote: members get their root / position, the called mutations are left in runs[0 .. n).
SIMT_FN int32_t w_cluster_vote(const MergeCtx &c, const MergeState &st, const WTable &t, int32_t *runs, int32_t cl, int lane) {
  const int32_t head = st.sl_head[cl - st.r0], size = st.sl_size[cl - st.r0];
  int32_t pos = 0;
  for (int32_t m = head; m >= 0; m = st.rd_next[m - st.r0]) {
    if (lane == 0) { c.read_root[m] = head; c.read_pos[m] = pos; if (m != head) c.row_size[m] = 0; }
    pos++;
    // genotype vote (:468-470): Q where the read has the key, -minQual where it has not
    for (int32_t e = c.geno_ptr[m] + lane; e < c.geno_ptr[m + 1]; e += 32) {
      GenoSlot *s = w_table_slot(t, c.geno_mut[e]);
      simt::atomic_add(&s->obs1, 1); simt::atomic_add64(&s->sum1, c.geno_q[e]);
    }
  }
  simt::sync(PCB_FULL);
  const int32_t n_t = *(volatile int32_t *)t.n_touched;
  int32_t n = 0;
  for (int32_t i0 = 0; i0 < n_t; i0 += 32) {              // collect the called mutations ...
    const int32_t i = i0 + lane;
    int32_t v = -1;
    bool called = false;
    if (i < n_t) {
      const GenoSlot *s = &t.slots[t.touched[i]];
      called = s->sum1 - (int64_t)c.prm.minQual * (int64_t)(size - s->obs1) > 0;
      v = s->key - 1;
    }
    const uint32_t m = simt::ballot(PCB_FULL, called);
    if (called) runs[n + simt::popc(m & ((1u << lane) - 1u))] = v;
    n += simt::popc(m);
  }
  simt::sync(PCB_FULL);
  return n;
}
// ... and write: each call at its (digit key, string) rank (:475), then the id columns and the row fields
SIMT_FN void w_cluster_write(const MergeCtx &c, const MergeState &st, const WTable &t, const int32_t *runs, int32_t cl, int32_t n,
                             int64_t call_pos, int lane) {
  const int32_t head = st.sl_head[cl - st.r0], size = st.sl_size[cl - st.r0];
  int32_t *calls = c.call_mut + call_pos;
  for (int32_t i = lane; i < n; i += 32) {
    const int32_t v = runs[i], rk = c.mut_rank[v];
    int32_t before = 0;
    for (int32_t k = 0; k < n; k++) before += c.mut_rank[runs[k]] < rk ? 1 : 0;
    calls[before] = v;
  }
  w_table_clear(t, lane);
  int32_t err = 0, uerr = 0;
  const int32_t bc = w_consensus_id(c, st, t, c.bc_id, head, size, &err, lane);
  const int32_t up = c.up_id ? w_consensus_id(c, st, t, c.up_id, head, size, &uerr, lane) : bc;
  if (lane == 0) {
    c.row_size[head] = size; c.row_key[head] = st.sl_key[cl - st.r0];
    c.row_call_off[head] = call_pos; c.row_call_n[head] = n;
    c.row_bc[head] = bc; c.row_up[head] = up;
    if (uerr) { c.status[0] = 3; c.status[1] = head; }     // clustered read without uptag: KeyError in the reference
  }
}

SIMT_FN void process_component_warp(const MergeCtx &c, const MergeState &st, int32_t comp, int lane) {
  int32_t *runs;
  const WTable t = w_make_table(st, runs, lane);
  const int32_t b_lo = c.comp_ptr[comp], b_hi = c.comp_ptr[comp + 1];

  // ---- seed clustering
  for (int32_t bi = b_lo; bi < b_hi; bi++) w_seed_bucket(c, st, t, c.comp_buckets ? c.comp_buckets[bi] : bi, lane);

  // ---- main clustering: rows in visiting order
  int32_t n_links = 0, n_tests = 0;
  RowMemo memo;
  memo.n = 0; memo.w = 0; memo.cache_root = -1;
  for (int32_t p = c.cpair_ptr[comp]; p < c.cpair_ptr[comp + 1]; p++)
    w_run_row(c, st, t, runs, c.cpair_q[p], c.cpair_r[p], c.cpair_ed[p], memo, n_links, n_tests, lane);

  // ---- consensus + rows (:464-477, :515-537, :546-567)
  w_table_clear(t, lane);
  int64_t call_pos = 0;
  {
    int32_t lo32 = 0, hi32 = 0;
    if (lane == 0) { const int64_t at = component_call_base(c, comp); lo32 = (int32_t)(at & 0xFFFFFFFFll); hi32 = (int32_t)(at >> 32); }
    lo32 = simt::shfl(PCB_FULL, lo32, 0); hi32 = simt::shfl(PCB_FULL, hi32, 0);
    call_pos = ((int64_t)hi32 << 32) | (int64_t)(uint32_t)lo32;
  }
  for (int32_t bi = b_lo; bi < b_hi; bi++) {
    const int32_t b = c.comp_buckets ? c.comp_buckets[bi] : bi;
    for (int32_t x = c.bucket_ptr[b]; x < c.bucket_ptr[b + 1]; x++) {
      const int32_t r = c.bucket_reads ? c.bucket_reads[x] : x;
      const int32_t cl = w_cluster_of(st, r);
      if (cl < 0) { call_pos += w_emit_singleton(c, r, call_pos, lane); continue; }
      if (st.sl_head[cl - st.r0] != r) continue;               // each cluster is emitted when its head comes by
      const int32_t n = w_cluster_vote(c, st, t, runs, cl, lane);
      w_cluster_write(c, st, t, runs, cl, n, call_pos, lane);
      call_pos += n;
    }
  }
  if (lane == 0) {
    if (n_links) simt::atomic_add(&c.status[2], n_links);
    if (n_tests) simt::atomic_add(&c.status[3], n_tests);
  }
}

}  // namespace pcb
