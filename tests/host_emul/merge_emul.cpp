// merge_emul.cpp -- UNIT-TEST HARNESS, not a product path.
// Compiles pacybara_b200/csrc/merge_core.h for the host and runs process_component() over
// every component sequentially, with the preprocessing (adjacency, components, grouping)
// done by plain std:: code, so that the sequential replay logic the GPU threads execute can
// be checked against the oracle in the GPU-less build container.  The shipped library never
// runs this: it launches the same function as a CUDA kernel (csrc/merge.cu).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "merge_core.h"
#include "merge_small.cuh"
#include "merge_huge.cuh"
#include "simt_emul.h"

using namespace pcb;

static int32_t uf_find(std::vector<int32_t> &p, int32_t x) {
  while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
  return x;
}

// ---- the giant tier (merge_huge.cuh) with the host side of csrc/merge.cu (run_huge) restated: every "launch" runs its
// units one after the other on an emulated warp, in shuffled order
namespace {
struct Scratch {
  std::vector<GenoSlot> slots;
  std::vector<int32_t> lists;
  int32_t cap = 0;
  void make(int32_t c) { cap = c; slots.assign((size_t)c, GenoSlot{0, 0, 0, 0, 0, 0}); lists.assign(8 + 2 * (size_t)c, 0); }
  MergeState state(const MergeCtx &c) {
    MergeState st;
    st.r0 = 0; st.rd_slot = c.rd_slot; st.rd_next = c.rd_next; st.sl_parent = c.sl_parent; st.sl_head = c.sl_head;
    st.sl_tail = c.sl_tail; st.sl_size = c.sl_size; st.sl_key = c.sl_key;
    st.slots = slots.data(); st.slot_mask = cap - 1; st.lists = lists.data(); st.list_cap = cap;
    return st;
  }
};
int32_t pow2_at_least(int64_t v) { int64_t c = 16; while (c < v) c <<= 1; return (int32_t)c; }
std::vector<int32_t> shuffled(int32_t n, uint32_t &rng) {
  std::vector<int32_t> o(n);
  std::iota(o.begin(), o.end(), 0);
  for (int32_t i = n - 1; i > 0; i--) { rng = rng * 1664525u + 1013904223u; std::swap(o[i], o[(rng >> 8) % (uint32_t)(i + 1)]); }
  return o;
}
}  // namespace

static void emul_huge(MergeCtx &c, const std::vector<int32_t> &mine, int32_t n_comp, int32_t R, int32_t U, int32_t cap, int32_t W, unsigned seed) {
  uint32_t rng = seed * 2654435761u + 99u;
  const int32_t n_huge = (int32_t)mine.size();
  if (n_huge == 0) return;
  std::vector<int32_t> comp_of_b(U, -1), huge_of_comp(n_comp, -1), hcomp(mine), hrow_ptr(n_huge + 1, 0), cell_nnz(R, 0);
  for (int32_t h = 0; h < n_huge; h++) {
    const int32_t comp = mine[h];
    huge_of_comp[comp] = h;
    for (int32_t i = c.comp_ptr[comp]; i < c.comp_ptr[comp + 1]; i++) comp_of_b[c.comp_buckets[i]] = comp;
    hrow_ptr[h + 1] = hrow_ptr[h] + (c.cpair_ptr[comp + 1] - c.cpair_ptr[comp]);
  }
  std::vector<unsigned long long> resv(R, ~0ull), cursor(n_huge, 0ull);
  std::vector<int32_t> win(W), carry0(W), carry1(W), heavy(W), heads(R), big(std::max(R, U));
  HugeCounters k;
  memset(&k, 0, sizeof(k));
  k.total = hrow_ptr[n_huge];
  HugeDev h;
  h.resv = resv.data(); h.comp_of_b = comp_of_b.data(); h.huge_of_comp = huge_of_comp.data(); h.hcomp = hcomp.data(); h.hrow_ptr = hrow_ptr.data();
  h.cursor = cursor.data(); h.n_huge = n_huge; h.W = W; h.win = win.data(); h.carry[0] = carry0.data(); h.carry[1] = carry1.data();
  h.heavy = heavy.data(); h.heads = heads.data(); h.big = big.data(); h.k = &k;
  c.cell_nnz = cell_nnz.data();
  Scratch ord, bigs;
  ord.make(cap);
  // seed step
  std::vector<int32_t> seedl;
  for (int32_t comp : mine)
    for (int32_t i = c.comp_ptr[comp]; i < c.comp_ptr[comp + 1]; i++) {
      const int32_t b = c.comp_buckets[i];
      if (c.bucket_ptr[b + 1] - c.bucket_ptr[b] >= 2) seedl.push_back(b);
    }
  for (int32_t i : shuffled((int32_t)seedl.size(), rng))
    simt_emul::run_warp([&](int lane) { huge_seed_item(c, ord.state(c), h, cap, false, seedl[i], lane); }, rng + (unsigned)i);
  if (k.n_big) {
    bigs.make(pow2_at_least(k.max_need));
    for (int32_t i : shuffled(k.n_big, rng))
      simt_emul::run_warp([&](int lane) { huge_seed_item(c, bigs.state(c), h, bigs.cap, true, big[i], lane); }, rng + (unsigned)i);
  }
  k.n_big = 0; k.max_need = 0;
  for (int32_t r = 0; r < R; r++)
    if (comp_of_b[c.bucket_of_read[r]] >= 0) huge_count_read(c, ord.state(c), r);
  // main loop
  for (int64_t round = 0; k.total > 0; round++) {
    huge_advance(h);
    if (k.done) break;
    for (int32_t idx : shuffled(k.n_window, rng))
      simt_emul::run_warp([&](int lane) { huge_reserve_item(c, ord.state(c), h, idx, lane); }, rng + (unsigned)idx);
    for (int32_t idx : shuffled(k.n_window, rng))
      simt_emul::run_warp([&](int lane) { huge_row_item(c, ord.state(c), h, cap, bigs.cap > 0, idx, lane); }, rng + (unsigned)idx);
    for (int32_t idx : shuffled(k.n_heavy, rng))
      simt_emul::run_warp([&](int lane) { huge_heavy_item(c, bigs.state(c), h, bigs.cap, idx, lane); }, rng + (unsigned)idx);
    if (round % 3 == 2 && k.max_need > bigs.cap) bigs.make(pow2_at_least(k.max_need));
  }
  // consensus
  k.max_need = 0;
  for (int32_t r : shuffled(R, rng)) {
    const int32_t comp = comp_of_b[c.bucket_of_read[r]];
    if (comp >= 0) huge_finish_read(c, ord.state(c), h, huge_of_comp[comp], r);
  }
  for (int32_t i : shuffled(k.n_heads, rng))
    simt_emul::run_warp([&](int lane) { huge_head_item(c, ord.state(c), h, cap, false, heads[i], lane); }, rng + (unsigned)i);
  if (k.n_big) {
    if (k.max_need > bigs.cap) bigs.make(pow2_at_least(k.max_need));
    for (int32_t i : shuffled(k.n_big, rng))
      simt_emul::run_warp([&](int lane) { huge_head_item(c, bigs.state(c), h, bigs.cap, true, big[i], lane); }, rng + (unsigned)i);
  }
  c.status[2] += k.links; c.status[3] += k.tests;
  c.cell_nnz = nullptr;
}

extern "C" int emul_merge(int32_t R, int32_t U, const int32_t *bucket_ptr, const int32_t *bucket_reads,
                          const int32_t *lexrank, const int32_t *bc_id, const int32_t *up_id,
                          const int32_t *geno_ptr, const int32_t *geno_mut, const int32_t *geno_q,
                          const int32_t *mut_rank, int64_t n_pairs, const int32_t *pair_q, const int32_t *pair_r,
                          const int32_t *pair_d, const int32_t *pair_ed, int32_t max_diff, int32_t min_qual,
                          double min_jaccard, int32_t min_matches, int32_t shard_rank, int32_t shard_count,
                          int32_t *read_root, int32_t *read_pos, int32_t *row_size, int64_t *row_key,
                          int32_t *row_bc, int32_t *row_up, int64_t *row_call_off, int32_t *row_call_n,
                          int32_t *call_mut, int32_t *status,
                          const uint64_t *planes, const int32_t *blen, int32_t compute_k, int32_t wide, int32_t use_small,
                          int32_t mode, int32_t huge_cap, int32_t huge_window) {
  // mode 2: every component through process_component_warp() on an emulated warp (merge_warp.cuh); mode 3: every component
  // through the giant tier (merge_huge.cuh) -- units of a "launch" in shuffled order, scratch of huge_cap slots, windows of
  // huge_window rows, the big scratch sized by the "host" between batches of 3 rounds, as csrc/merge.cu does.
  // use_small != 0: every component goes through process_component_small() on an emulated warp first (use_small =
  // lane-order seed), and through process_component() only if that declines.  status[1] returns how many it took.
  std::vector<int32_t> bucket_of_read(R);
  for (int32_t b = 0; b < U; b++)
    for (int32_t x = bucket_ptr[b]; x < bucket_ptr[b + 1]; x++) bucket_of_read[bucket_reads[x]] = b;
  // adjacency
  struct E { int32_t a, b, d; };
  std::vector<E> adj;
  for (int64_t p = 0; p < n_pairs; p++) {
    adj.push_back({pair_q[p], pair_r[p], pair_d[p]});
    adj.push_back({pair_r[p], pair_q[p], pair_d[p]});
  }
  std::sort(adj.begin(), adj.end(), [](const E &x, const E &y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
  std::vector<int32_t> adj_ptr(U + 1, 0), adj_nbr(adj.size()), adj_d(adj.size());
  for (size_t i = 0; i < adj.size(); i++) { adj_ptr[adj[i].a + 1]++; adj_nbr[i] = adj[i].b; adj_d[i] = adj[i].d; }
  for (int32_t b = 0; b < U; b++) adj_ptr[b + 1] += adj_ptr[b];
  // components of the link graph
  std::vector<int32_t> par(U);
  std::iota(par.begin(), par.end(), 0);
  for (int64_t p = 0; p < n_pairs; p++)
    if (pair_ed[p] >= 1) {
      int32_t a = uf_find(par, pair_q[p]), b = uf_find(par, pair_r[p]);
      if (a != b) par[std::max(a, b)] = std::min(a, b);
    }
  std::vector<int32_t> label(U), comp_of_label(U, -1), comp_buckets(U), comp_ptr;
  for (int32_t b = 0; b < U; b++) label[b] = uf_find(par, b);
  std::vector<int32_t> order(U);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return label[x] < label[y]; });
  int32_t n_comp = 0;
  for (int32_t i = 0; i < U; i++) {
    if (i == 0 || label[order[i]] != label[order[i - 1]]) { comp_of_label[label[order[i]]] = n_comp++; comp_ptr.push_back(i); }
    comp_buckets[i] = order[i];
  }
  comp_ptr.push_back(U);
  std::vector<int32_t> cpair_ptr(n_comp + 1, 0), cq, cr, ce;
  std::vector<int64_t> pidx;
  for (int64_t p = 0; p < n_pairs; p++) if (pair_ed[p] >= 1) pidx.push_back(p);
  std::stable_sort(pidx.begin(), pidx.end(), [&](int64_t x, int64_t y) {
    return comp_of_label[label[pair_q[x]]] < comp_of_label[label[pair_q[y]]]; });
  for (int64_t p : pidx) { cpair_ptr[comp_of_label[label[pair_q[p]]] + 1]++; cq.push_back(pair_q[p]); cr.push_back(pair_r[p]); ce.push_back(pair_ed[p]); }
  for (int32_t c = 0; c < n_comp; c++) cpair_ptr[c + 1] += cpair_ptr[c];
  // scratch
  std::vector<int64_t> slot_off(n_comp + 1, 0), list_off(n_comp + 1, 0), call_off(n_comp + 1, 0);
  std::vector<int32_t> slot_cap(n_comp);
  for (int32_t c = 0; c < n_comp; c++) {
    int64_t reads = 0, nnz = 0;
    for (int32_t i = comp_ptr[c]; i < comp_ptr[c + 1]; i++) {
      int32_t b = comp_buckets[i];
      reads += bucket_ptr[b + 1] - bucket_ptr[b];
      for (int32_t x = bucket_ptr[b]; x < bucket_ptr[b + 1]; x++) nnz += geno_ptr[bucket_reads[x] + 1] - geno_ptr[bucket_reads[x]];
    }
    int64_t need = 2 * std::max<int64_t>(std::max(reads, nnz), 1);
    int32_t cap = 2; while (cap < need) cap <<= 1;
    slot_cap[c] = cap;
    slot_off[c + 1] = slot_off[c] + cap;
    list_off[c + 1] = list_off[c] + 2 * (reads + nnz) + 8;
    call_off[c + 1] = call_off[c] + nnz;
  }
  std::vector<GenoSlot> slots(slot_off[n_comp]);
  memset(slots.data(), 0, slots.size() * sizeof(GenoSlot));
  std::vector<int32_t> lists(list_off[n_comp] + 1);
  std::vector<int32_t> rd_slot(R, -1), rd_next(R, -1), sl_parent(R, -1), sl_head(R, -1), sl_tail(R, -1), sl_size(R, 0);
  std::vector<int64_t> sl_key(R, -1);
  int64_t nnz_tot = geno_ptr[R];
  for (int32_t r = 0; r < R; r++) {
    read_root[r] = read_pos[r] = row_size[r] = row_bc[r] = row_up[r] = row_call_n[r] = PCB_FILL;
    row_key[r] = INT64_MIN; row_call_off[r] = INT64_MIN;
  }
  for (int64_t i = 0; i < nnz_tot; i++) call_mut[i] = PCB_FILL;
  status[0] = status[1] = status[2] = status[3] = 0;

  MergeCtx c;
  memset(&c, 0, sizeof(c));
  c.R = R; c.U = U;
  c.bucket_ptr = bucket_ptr; c.bucket_reads = bucket_reads; c.bucket_of_read = bucket_of_read.data();
  c.lexrank = lexrank; c.bc_id = bc_id; c.up_id = up_id;
  c.geno_ptr = geno_ptr; c.geno_mut = geno_mut; c.geno_q = geno_q; c.mut_rank = mut_rank;
  c.adj_ptr = adj_ptr.data(); c.adj_nbr = adj_nbr.data(); c.adj_d = adj_d.data();
  c.planes = planes; c.blen = blen; c.compute_k = planes ? compute_k : -1; c.wide = wide;
  c.comp_ptr = comp_ptr.data(); c.comp_buckets = comp_buckets.data();
  c.cpair_ptr = cpair_ptr.data(); c.cpair_q = cq.data(); c.cpair_r = cr.data(); c.cpair_ed = ce.data();
  c.scr_slot_off = slot_off.data(); c.scr_slot_cap = slot_cap.data(); c.scr_list_off = list_off.data();
  c.call_off = call_off.data(); c.slots = slots.data(); c.lists = lists.data();
  c.rd_slot = rd_slot.data(); c.rd_next = rd_next.data(); c.sl_parent = sl_parent.data();
  c.sl_head = sl_head.data(); c.sl_tail = sl_tail.data(); c.sl_size = sl_size.data(); c.sl_key = sl_key.data();
  c.read_root = read_root; c.read_pos = read_pos; c.row_size = row_size; c.row_key = row_key;
  c.row_bc = row_bc; c.row_up = row_up; c.row_call_off = row_call_off; c.row_call_n = row_call_n;
  c.call_mut = call_mut; c.status = status;
  c.prm.maxDiff = max_diff; c.prm.minQual = min_qual; c.prm.minMatches = min_matches; c.prm.minJaccard = min_jaccard;
  jaccard_thresholds(c.prm);
  int32_t n_small = 0;
  int64_t call_cursor = 0;
  if (use_small && mode != 3) { c.call_off = nullptr; c.call_cursor = &call_cursor; }     // regions from the shared cursor, like csrc/merge.cu
  static SmallSmem smem;
  std::vector<int32_t> mine;
  for (int32_t comp = 0; comp < n_comp; comp++)
    if (label[comp_buckets[comp_ptr[comp]]] % shard_count == shard_rank) mine.push_back(comp);
  if (mode == 3) {
    emul_huge(c, mine, n_comp, R, U, huge_cap, huge_window, (unsigned)use_small + 17u);
    return n_comp;
  }
  for (int32_t comp : mine) {
    if (mode == 2) {
      simt_emul::run_warp([&](int lane) { process_component_warp(c, global_state(c, comp), comp, lane); }, (unsigned)(7 + comp));
      continue;
    }
    if (use_small) {
      int rc[32];
      simt_emul::run_warp([&](int lane) { rc[lane] = process_component_small(c, comp, lane, &smem, c.prm.jthr, &c.status[2], 1 << 30, 1 << 30); }, (unsigned)(use_small + comp));
      for (int l = 1; l < 32; l++) if (rc[l] != rc[0]) return -2;
      if (rc[0] == 0) { n_small++; continue; }
    }
    process_component(c, global_state(c, comp), comp);
  }
  if (use_small && status[0] == 0) status[1] = n_small;
  return n_comp;
}

// ---- self test of the emulator (tests/test_merge_small_host.py): neighbour exchange through a shared array
static int selftest(bool with_sync) {
  int wrong = 0;
  for (unsigned seed = 1; seed <= 20; seed++) {
    static int32_t buf[32];
    int got[32];
    for (int l = 0; l < 32; l++) buf[l] = -1;
    simt_emul::run_warp([&](int lane) {
      buf[lane] = lane * 7;
      if (with_sync) simt::sync(0xFFFFFFFFu);
      got[lane] = buf[(lane + 1) & 31];
      simt::sync(0xFFFFFFFFu);
    }, seed);
    for (int l = 0; l < 32; l++) if (got[l] != ((l + 1) & 31) * 7) wrong = 1;
  }
  return wrong;
}
extern "C" int emul_selftest_missing_sync() { return selftest(false); }
extern "C" int emul_selftest_with_sync() { return selftest(true); }
