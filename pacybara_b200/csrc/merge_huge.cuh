// merge_huge.cuh -- the replay INSIDE one giant component, spread over the whole device.
//
// Dense barcode spaces (templates/pacybara_parameters.txt:13, a 25-nt SWSW... alphabet = 2^25 codes) join most buckets
// of a library into one connected component of the link graph: 5.8*10^5 buckets and 1.1*10^6 rows at 3*10^6 reads.
// One warp replaying that component alone (merge_warp.cuh) takes tens of seconds.  The reference's algorithm
// (src/pacybara_cluster.py:385-455) is sequential, but what one step reads and writes is narrow:
//   seed step (:385-400)    per bucket: only the reads of that bucket                  -> one warp per bucket
//   main loop (:408-455)    per row (bucket pair): only the clusters that hold a read of either bucket -- their
//                           member lists, sizes, genotype sums; `known` marks concern the pair itself, the creation
//                           keys of cluster ids are made in the seed step (two lone reads never link: |log2(1/1)| < ed)
//   consensus (:464-537)    per cluster / per lone read                                 -> one warp per cluster
// So the main loop is replayed with DETERMINISTIC RESERVATIONS: the pending rows (a window of the visiting order) each
// write their rank into a reservation cell of every cluster they could touch (atomicMin); a row that holds all its
// cells has no earlier pending row that could still touch its clusters, so it sees exactly the state the sequential
// loop would show it and is executed now -- all such rows of a round in parallel, on disjoint clusters; the others
// wait for the next round.  The earliest pending row always wins, so every round makes progress, and the outcome is
// the sequential one whatever the schedule (tests: the goldens and the fuzz libraries through this path, in the
// emulator with shuffled item order and on the device).
//   cell of a read = root slot of its cluster, or the read itself while it is in no cluster.  A merge keeps one of the
//   two roots, a lone read joins a root: whatever a winning row turns its clusters into stays inside its own cells.
//   stamp = (2^31 - 1 - round) << 32 | row: newer rounds beat the stale stamps of older ones, so cells are never reset.
// Scratch (genotype table, run list) is per warp and fixed (cap slots); a unit that needs more -- the sum of the
// genotype entries of its clusters, kept per root in cell_nnz -- stays pending (rows) or goes on a list (buckets,
// clusters) for a few warps with scratch sized on demand by the host.
// Written against simt.h: tests/host_emul runs these very functions on the CPU, one emulated warp per unit, the units
// of a "launch" in shuffled order (unit tests only, never a product path).
#pragma once
#include "merge_warp.cuh"

namespace pcb {

struct HugeCounters {
  int32_t round, parity;                 // carry[parity] holds the rows that lost the previous round
  int32_t total, next_fresh;             // rows of all giant components / the first one no window has held yet
  int32_t n_window, n_carry_in, fresh_lo;
  int32_t n_carry_out, n_heavy;
  int32_t ticket, ticket_heavy;
  int32_t max_need;                      // largest scratch a declined unit asked for (slots)
  int32_t done;
  int32_t links, tests;
  int32_t n_heads, n_big;                // consensus: cluster heads / seed + consensus units for the big-scratch warps
  int32_t n_seed;                        // buckets the seed step has work for
  int32_t tk[4];                         // tickets: seed list, big seed list, heads, big heads
};

struct HugeDev {
  unsigned long long *resv;              // [R] reservation cells, all ones = free
  const int32_t *comp_of_b;              // [U] bucket -> component (this shard's), -1 = another shard's
  const int32_t *huge_of_comp;           // [n_comp] component -> index among the giant ones, -1 = not giant
  const int32_t *hcomp;                  // [n_huge] giant components
  const int32_t *hrow_ptr;               // [n_huge + 1] their rows, as a prefix
  unsigned long long *cursor;            // [n_huge] calls handed out inside each one's call region
  int32_t n_huge, W;
  int32_t *win;                          // [W] rows of this round
  int32_t *carry[2];                     // [W] each
  int32_t *heavy;                        // [W] winners that need the big scratch
  int32_t *heads;                        // [R] cluster heads of the giant components
  int32_t *big;                          // [R] buckets / heads for the big-scratch warps
  HugeCounters *k;
};

SIMT_FN int32_t huge_cell(const MergeState &st, int32_t r) {
  const int32_t s = ((volatile int32_t *)st.rd_slot)[r - st.r0];
  return s < 0 ? r : w_find_ro(st, s);
}
SIMT_FN unsigned long long huge_stamp(int32_t round, int32_t p) {
  return ((unsigned long long)(uint32_t)(0x7FFFFFFF - round) << 32) | (unsigned long long)(uint32_t)p;
}
// slots a unit needs: a table at most half full over `nnz` keys, lists of `nnz` and `reads` entries
SIMT_FN int32_t huge_need(int32_t nnz, int32_t reads) { const int32_t m = nnz > reads ? nnz : reads; return m > (1 << 29) ? (1 << 30) : 2 * m; }

// window position -> row (index into cpair_*): the losers of the previous round, then fresh rows in visiting order
SIMT_FN int32_t huge_window_row(const MergeCtx &c, const HugeDev &h, int32_t idx) {
  const HugeCounters &k = *h.k;
  if (idx < k.n_carry_in) return h.carry[k.parity][idx];
  const int32_t f = k.fresh_lo + (idx - k.n_carry_in);
  int32_t lo = 0, hi = h.n_huge;                              // last giant component whose first row is <= f
  while (hi - lo > 1) { const int32_t mid = (lo + hi) >> 1; if (h.hrow_ptr[mid] <= f) lo = mid; else hi = mid; }
  return c.cpair_ptr[h.hcomp[lo]] + (f - h.hrow_ptr[lo]);
}

// between two rounds (one thread)
SIMT_FN void huge_advance(const HugeDev &h) {
  HugeCounters &k = *h.k;
  k.parity ^= 1;                                              // what the last round carried out is carried in
  k.n_carry_in = k.n_carry_out; k.n_carry_out = 0; k.n_heavy = 0; k.ticket = 0; k.ticket_heavy = 0;
  int32_t fresh = h.W - k.n_carry_in;
  if (fresh > k.total - k.next_fresh) fresh = k.total - k.next_fresh;
  k.fresh_lo = k.next_fresh; k.next_fresh += fresh;
  k.n_window = k.n_carry_in + fresh;
  k.round++;
  if (k.n_window == 0) k.done = 1;
}

// every read of the row's two buckets, 32 at a time: f(read, active)
#define PCB_ROW_READS(c, p, lane, BODY)                                                           \
  for (int side_ = 0; side_ < 2; side_++) {                                                       \
    const int32_t b_ = side_ ? (c).cpair_r[p] : (c).cpair_q[p];                                   \
    const int32_t x0_ = (c).bucket_ptr[b_], x1_ = (c).bucket_ptr[b_ + 1];                         \
    for (int32_t xb_ = x0_; xb_ < x1_; xb_ += 32) {                                               \
      const int32_t x_ = xb_ + (lane);                                                            \
      const bool act = x_ < x1_;                                                                  \
      const int32_t r = act ? ((c).bucket_reads ? (c).bucket_reads[x_] : x_) : -1;                \
      BODY                                                                                        \
    }                                                                                             \
  }

// round, step 1: the row puts its stamp on every cell it could touch
SIMT_FN void huge_row_reserve(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t p, int lane) {
  const unsigned long long stamp = huge_stamp(h.k->round, p);
  PCB_ROW_READS(c, p, lane, { if (act) simt::atomic_min_u64(&h.resv[huge_cell(st, r)], stamp); })
}

// ... for window position idx, which is also where the row is noted for step 2
SIMT_FN void huge_reserve_item(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t idx, int lane) {
  const int32_t p = huge_window_row(c, h, idx);
  if (lane == 0) h.win[idx] = p;
  huge_row_reserve(c, st, h, p, lane);
}

// round, step 2: does it hold them all?  need = scratch slots its visit can use
SIMT_FN bool huge_row_check(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t p, int lane, int32_t &need) {
  const unsigned long long stamp = huge_stamp(h.k->round, p);
  bool lost = false;
  int32_t nnz = 0, reads = 0;
  PCB_ROW_READS(c, p, lane, {
    const int32_t cell = act ? huge_cell(st, r) : -1 - lane;
    if (act && *(volatile unsigned long long *)&h.resv[cell] != stamp) lost = true;
    const uint32_t same = simt::match_any(PCB_FULL, cell);
    if (act && simt::ffs(same) - 1 == lane) {                 // one lane per distinct cell of these 32 reads
      nnz += c.cell_nnz[cell];
      reads += ((volatile int32_t *)st.rd_slot)[r - st.r0] < 0 ? 1 : st.sl_size[cell - st.r0];
    }
  })
  nnz = simt::reduce_add(PCB_FULL, nnz); reads = simt::reduce_add(PCB_FULL, reads);
  need = huge_need(nnz, reads);
  return simt::ballot(PCB_FULL, lost) == 0u;
}

SIMT_FN void huge_row_execute(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t p, int lane) {
  int32_t *runs;
  const WTable t = w_make_table(st, runs, lane);
  RowMemo memo;
  memo.n = 0; memo.w = 0; memo.cache_root = -1;
  int32_t n_links = 0, n_tests = 0;
  w_run_row(c, st, t, runs, c.cpair_q[p], c.cpair_r[p], c.cpair_ed[p], memo, n_links, n_tests, lane);
  w_table_clear(t, lane);
  if (lane == 0) {
    if (n_links) simt::atomic_add(&h.k->links, n_links);
    if (n_tests) simt::atomic_add(&h.k->tests, n_tests);
  }
}

SIMT_FN void huge_push(int32_t *list, int32_t *n, int32_t v, int lane) {
  if (lane == 0) list[simt::atomic_add(n, 1)] = v;
}

// round, step 2 for one window position, by a warp whose scratch (st.slots / st.lists) holds `cap` slots
SIMT_FN void huge_row_item(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t cap, bool have_big, int32_t idx, int lane) {
  const int32_t p = h.win[idx];
  int32_t need = 0;
  if (!huge_row_check(c, st, h, p, lane, need)) { huge_push(h.carry[h.k->parity ^ 1], &h.k->n_carry_out, p, lane); return; }
  if (need <= cap) { huge_row_execute(c, st, h, p, lane); return; }
  if (have_big) { huge_push(h.heavy, &h.k->n_heavy, p, lane); return; }
  if (lane == 0) simt::atomic_max(&h.k->max_need, need);      // stays pending (and keeps holding its cells) until the host has sized the big scratch
  huge_push(h.carry[h.k->parity ^ 1], &h.k->n_carry_out, p, lane);
}
// ... and for a winner the ordinary warps passed on, by a big-scratch warp
SIMT_FN void huge_heavy_item(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t cap, int32_t idx, int lane) {
  const int32_t p = h.heavy[idx];
  int32_t need = 0;
  huge_row_check(c, st, h, p, lane, need);                    // (it won; this is for `need`)
  if (need <= cap) { huge_row_execute(c, st, h, p, lane); return; }
  if (lane == 0) simt::atomic_max(&h.k->max_need, need);
  huge_push(h.carry[h.k->parity ^ 1], &h.k->n_carry_out, p, lane);
}

// seed step of bucket b of a giant component; a bucket too big for `cap` goes on the big list (have_big: this IS a
// big-scratch warp and the host has sized it for every listed bucket)
SIMT_FN void huge_seed_item(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t cap, bool is_big, int32_t b, int lane) {
  const int32_t x0 = c.bucket_ptr[b], n = c.bucket_ptr[b + 1] - x0;
  if (n < 2) return;
  int32_t nnz = 0;
  for (int32_t x = x0 + lane; x < x0 + n; x += 32) { const int32_t r = c.bucket_reads ? c.bucket_reads[x] : x; nnz += c.geno_ptr[r + 1] - c.geno_ptr[r]; }
  nnz = simt::reduce_add(PCB_FULL, nnz);
  const int32_t need = huge_need(nnz, n);
  if (need > cap && !is_big) {
    if (lane == 0) simt::atomic_max(&h.k->max_need, need);
    huge_push(h.big, &h.k->n_big, b, lane);
    return;
  }
  int32_t *runs;
  const WTable t = w_make_table(st, runs, lane);
  w_seed_bucket(c, st, t, b, lane);
}

// after the seed step, one THREAD per read of a giant component: genotype entries per cell
SIMT_FN void huge_count_read(const MergeCtx &c, const MergeState &st, int32_t r) {
  const int32_t n = c.geno_ptr[r + 1] - c.geno_ptr[r];
  if (n) simt::atomic_add(&c.cell_nnz[huge_cell(st, r)], n);
}

// where the calls of a row of giant component `hi` go: the next n entries of its call region
SIMT_FN int64_t huge_take_calls(const MergeCtx &c, const HugeDev &h, int32_t hi, int32_t n) {
  return c.call_off[h.hcomp[hi]] + (int64_t)simt::atomic_add_u64(&h.cursor[hi], (unsigned long long)n);
}

// consensus, one THREAD per read of a giant component: a lone read writes its row (:560-567), the head of a
// cluster goes on the list of the warps
SIMT_FN void huge_finish_read(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t hi, int32_t r) {
  const int32_t s = st.rd_slot[r - st.r0];
  if (s >= 0) {
    const int32_t cl = w_find_ro(st, s);
    if (st.sl_head[cl - st.r0] == r) h.heads[simt::atomic_add(&h.k->n_heads, 1)] = r;
    return;
  }
  int32_t n = 0;
  for (int32_t e = c.geno_ptr[r]; e < c.geno_ptr[r + 1]; e++) n += c.geno_q[e] >= c.prm.minQual ? 1 : 0;
  const int64_t at = huge_take_calls(c, h, hi, n);
  int32_t k = 0;
  for (int32_t e = c.geno_ptr[r]; e < c.geno_ptr[r + 1]; e++)
    if (c.geno_q[e] >= c.prm.minQual) c.call_mut[at + k++] = c.geno_mut[e];
  c.read_root[r] = r; c.read_pos[r] = 0; c.row_size[r] = 1; c.row_key[r] = -1;
  c.row_bc[r] = c.bc_id[r]; c.row_up[r] = c.up_id ? c.up_id[r] : c.bc_id[r];
  c.row_call_off[r] = at; c.row_call_n[r] = n;
}

// consensus of the cluster whose head is read `head` (:464-477, :515-537)
SIMT_FN void huge_head_item(const MergeCtx &c, const MergeState &st, const HugeDev &h, int32_t cap, bool is_big, int32_t head, int lane) {
  const int32_t cl = w_cluster_of(st, head);
  const int32_t need = huge_need(c.cell_nnz[cl], st.sl_size[cl - st.r0]);
  if (need > cap && !is_big) {
    if (lane == 0) simt::atomic_max(&h.k->max_need, need);
    huge_push(h.big, &h.k->n_big, head, lane);
    return;
  }
  int32_t *runs;
  const WTable t = w_make_table(st, runs, lane);
  const int32_t n = w_cluster_vote(c, st, t, runs, cl, lane);
  int32_t lo32 = 0, hi32 = 0;
  if (lane == 0) {
    const int64_t at = huge_take_calls(c, h, h.huge_of_comp[h.comp_of_b[c.bucket_of_read[head]]], n);
    lo32 = (int32_t)(at & 0xFFFFFFFFll); hi32 = (int32_t)(at >> 32);
  }
  lo32 = simt::shfl(PCB_FULL, lo32, 0); hi32 = simt::shfl(PCB_FULL, hi32, 0);
  w_cluster_write(c, st, t, runs, cl, n, ((int64_t)hi32 << 32) | (int64_t)(uint32_t)lo32, lane);
}

}  // namespace pcb
