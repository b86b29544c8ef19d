// merge_small.cuh -- the per-component replay for components of at most 32 reads: ONE WARP per component,
// one LANE per read, the whole cluster state in registers.
//
// Same algorithm, same visiting order and the same rows as process_component() in merge_core.h (seed step
// src/pacybara_cluster.py:385-400, main loop :408-455, consensus :464-477,:515-537), reformulated so that the
// lanes of a warp do data-parallel work instead of one lane chasing pointers:
//
//   cluster state    per read (= lane): cid (row of its cluster, -1 = in no cluster), pos (place in the member
//                    list), key (creation key of the surviving cluster id).  "Relabel every member and append
//                    its list" (:193-202) is one predicated update: lanes of c2 add size(c1) to pos and take
//                    c1's cid and key.  Sizes and member sets are ballots.
//   genotype sums    a dense matrix in shared memory: row = cluster (or lone read), column = one of the <= 32
//                    distinct mutations of the component, LANE m OWNS COLUMN m.  cell = (sum of Q, #reads with
//                    Q > 0, #reads naming the mutation).  A merge adds two rows; the genotype rule (:447-452) is
//                    two row reads, one filterSeqError test per lane and two ballots -- no hash table, no
//                    accumulate-per-test.  After the fill pass a lane only ever touches its own column, so no
//                    further synchronisation is needed.
//   seed step        per bucket: the bucket-wide filter (:350-356) is a column sum; presence masks of the reads
//                    come out of the same loop (one ballot per read); row i of the pair loop (:391-400) is ONE
//                    parallel Jaccard test of read i against every j < i (its verdict depends on the filter and
//                    the two reads only, not on link order), then addLink is replayed over the set bits.
//   main loop        a row's read pairs (x, y) are screened 32 at a time (same cluster / already failed since the
//                    last merge), the first live one is visited, then the rest are screened again.
//   transitive rule  the buckets within 2e of each bucket of the component as bit masks (one lookup per bucket
//                    pair, spread over the lanes, on-demand Myers included); the rule (:442) is then
//                    "no bucket of cluster 1 has a bucket of cluster 2 outside its mask".
//   consensus        vote per column from the cluster's row; order of the calls by counting smaller ranks;
//                    barcode majority with match_any.
//
// Eligibility: <= 32 reads, <= 32 LIVE mutations (a mutation only one read of the component names, with Q <= minQual,
// is dead: it can pass no filter and win no vote, and gets no column), <= 128 distinct mutations, |Q| <= 2^24.
// Anything else returns non-zero before a single output is written and is replayed by the generic kernel.
// The caller's numbering is read in place: component -> comp_buckets -> bucket_ptr / bucket_reads -> read.
// Written against simt.h so that tests/host_emul runs this very source on the CPU (unit tests only).
#pragma once
#include "merge_core.h"
#include "simt.h"

namespace pcb {

constexpr int SMALL_READS = 32;     // one lane per read
constexpr int SMALL_MUTS = 32;      // one matrix column per lane
constexpr int SMALL_STRIDE = 33;    // row stride: conflict-free along a row (lane = column) and down a column (lane = row)
constexpr int SMALL_TAB = 128;      // open-addressing table over the component's mutations (live ones get a column)

struct alignas(16) SmallSmem {
  int32_t sum[SMALL_READS * SMALL_STRIDE];    // sum of Q
  uint16_t cnt[SMALL_READS * SMALL_STRIDE];   // low byte: #reads with Q > 0; high byte: #reads naming the mutation
  int32_t tab[SMALL_TAB];                     // mutation id + 1, 0 = empty
  uint32_t tflag[SMALL_TAB];                  // 2: named by two reads or more; 4: some Q > minQual; then: column + 1 (0 = dead)
  int32_t colmut[SMALL_MUTS];                 // column -> mutation id
  uint32_t near[3][32];                       // near[e-1][a]: buckets b of the component with dist(a, b) <= 2e
  int32_t bid[32];                            // local bucket -> bucket id
};

SIMT_FN uint32_t small_hash(int32_t mut) { return ((uint32_t)mut * 0x9E3779B1u) >> 25; }      // 7 bits

// Returns 0 when the component was replayed; otherwise nothing was written: 3 = more than 32 reads (not this kernel's
// tier), 1 = not eligible (more than 32 live mutations, huge qualities), 2 = not eligible AND more than med_calls
// genotype entries or med_pairs rows (too big for the medium kernel's local copy as well).
// jthr: MergeParams::jthr somewhere a lane can index cheaply (shared memory on the device).
// tally: two counters (links, tests) the caller sums up (shared memory per block on the device: same-address global
// atomics from every component would serialise in L2).
SIMT_FN int process_component_small(const MergeCtx &c, int32_t comp, int lane, SmallSmem *sm, const int32_t *jthr, int32_t *tally,
                                    int32_t med_calls, int32_t med_pairs) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lt = (1u << lane) - 1u;
  // The component's buckets are comp_buckets[b_lo .. b_hi) (ascending bucket id), bucket b's reads are
  // bucket_reads[bucket_ptr[b] .. bucket_ptr[b + 1]) in arrival order -- the caller's numbering, nothing is gathered
  // beforehand.  Lane L is the L-th read of that sequence; lane bi < B also keeps bucket bi (id, first lane, size).
  const int32_t b_lo = c.comp_ptr[comp], b_hi = c.comp_ptr[comp + 1];
  const int32_t B = b_hi - b_lo;
  if (B > SMALL_READS) return 3;
  int32_t myb = -1, mysz = 0, myx0 = 0;
  if (lane < B) {
    myb = c.comp_buckets ? c.comp_buckets[b_lo + lane] : b_lo + lane;
    myx0 = c.bucket_ptr[myb]; mysz = c.bucket_ptr[myb + 1] - myx0;
  }
  int32_t n = 0, mystart = 0;
  uint32_t starts = 0;                     // bit L: a bucket starts at lane L (buckets are never empty)
  for (int32_t bi = 0; bi < B; bi++) {
    const int32_t sz = simt::shfl(FULL, mysz, bi);
    if (lane == bi) mystart = n;
    if (n < 32) starts |= 1u << n;
    n += sz;
  }
  if (n > SMALL_READS) return 3;
  const bool has = lane < n;
  const int32_t minQual = c.prm.minQual, minMatches = c.prm.minMatches;
  const int32_t bkt = simt::popc(starts & (2u * (1u << lane) - 1u)) - 1;              // my bucket: last start <= lane
  const int32_t bsrc = has ? bkt : 0;
  const int32_t k_in_b = lane - simt::shfl(FULL, mystart, bsrc), x0_of_b = simt::shfl(FULL, myx0, bsrc);
  int32_t r = 0, lex = 0, bc = 0, up = 0, g0 = 0, g1 = 0;
  if (has) {
    r = c.bucket_reads ? c.bucket_reads[x0_of_b + k_in_b] : x0_of_b + k_in_b;
    lex = c.lexrank[r]; bc = c.bc_id[r]; up = c.up_id ? c.up_id[r] : 0;
    g0 = c.geno_ptr[r]; g1 = c.geno_ptr[r + 1];
  }
  sm->bid[lane] = myb;

  // ---- number the mutations of the component (column = rank of the table slot), fill the matrix
  {
    struct alignas(16) Q16 { int32_t a, b, c, d; };             // 16-byte stores: rows 0..n-1 of both matrices
    const Q16 zero = {0, 0, 0, 0};
    Q16 *zs = (Q16 *)sm->sum, *zc = (Q16 *)sm->cnt;
    for (int32_t i = lane; i < (n * SMALL_STRIDE + 3) / 4; i += 32) zs[i] = zero;
    for (int32_t i = lane; i < (n * SMALL_STRIDE + 7) / 8; i += 32) zc[i] = zero;
  }
  for (int32_t i = lane; i < SMALL_TAB; i += 32) { sm->tab[i] = 0; sm->tflag[i] = 0u; }
  simt::sync(FULL);
  // A mutation that a single read of the whole component names, with Q <= minQual, can never pass filterSeqError
  // (:350-356: it is seen once and its sum stays <= minQual in any pair of clusters) nor be called in a cluster
  // (Q - minQual * (size - 1) <= 0): such DEAD mutations -- the private sequencing-error calls -- get no column.
  bool bad = false;
  for (int32_t e = g0; e < g1; e++) {
    const int32_t mut = c.geno_mut[e], q = c.geno_q[e];
    if (q > (1 << 24) || q < -(1 << 24)) bad = true;
    uint32_t h = small_hash(mut);
    for (int probes = 0;; probes++) {
      if (probes == SMALL_TAB) { bad = true; break; }
      const int32_t old = simt::atomic_cas(&sm->tab[h], 0, mut + 1);
      if (old == 0 || old == mut + 1) {
        const uint32_t f = (old != 0 ? 2u : 0u) | (q > minQual ? 4u : 0u);
        if (f) simt::atomic_or(&sm->tflag[h], f);
        break;
      }
      h = (h + 1) & (SMALL_TAB - 1);
    }
  }
  simt::sync(FULL);
  uint32_t live[SMALL_TAB / 32];
  int32_t M = 0;
#pragma unroll
  for (int w = 0; w < SMALL_TAB / 32; w++) { live[w] = simt::ballot(FULL, sm->tflag[32 * w + lane] != 0u); M += simt::popc(live[w]); }
  if (simt::ballot(FULL, bad) != 0u || M > SMALL_MUTS) {
    const int32_t ne = simt::reduce_add(FULL, g1 - g0);
    return (ne > med_calls || c.cpair_ptr[comp + 1] - c.cpair_ptr[comp] > med_pairs) ? 2 : 1;
  }
  {
    int32_t base = 0;                      // live slots get the columns 0 .. M-1 in slot order
#pragma unroll
    for (int w = 0; w < SMALL_TAB / 32; w++) {
      if ((live[w] >> lane) & 1u) {
        const int32_t col = base + simt::popc(live[w] & lt);
        sm->colmut[col] = sm->tab[32 * w + lane] - 1;
        sm->tflag[32 * w + lane] = (uint32_t)col + 1u;
      }
      base += simt::popc(live[w]);
    }
  }
  simt::sync(FULL);
  for (int32_t e = g0; e < g1; e++) {
    const int32_t mut = c.geno_mut[e], q = c.geno_q[e];
    uint32_t h = small_hash(mut);
    while (sm->tab[h] != mut + 1) h = (h + 1) & (SMALL_TAB - 1);
    const uint32_t col1 = sm->tflag[h];
    if (col1 == 0u) continue;                                                    // dead
    sm->sum[lane * SMALL_STRIDE + (col1 - 1u)] = q;                              // mutation ids are distinct within a read
    sm->cnt[lane * SMALL_STRIDE + (col1 - 1u)] = (uint16_t)(0x100 | (q > 0 ? 1 : 0));
  }
  simt::sync(FULL);
  const int32_t mymut = lane < M ? sm->colmut[lane] : -1;
  const int32_t myrank = lane < M ? c.mut_rank[mymut] : 0x7fffffff;

  // ---- transitive rule: which buckets lie within 2e of each bucket (only components that have rows to visit)
  const int32_t p_lo = c.cpair_ptr[comp], p_hi = c.cpair_ptr[comp + 1];
  uint32_t near1 = 0, near2 = 0, near3 = 0;
  if (p_hi > p_lo) {
    if (lane < B) { sm->near[0][lane] = 1u << lane; sm->near[1][lane] = 1u << lane; sm->near[2][lane] = 1u << lane; }
    simt::sync(FULL);
    const int32_t n_bp = B * (B - 1) / 2;
    for (int32_t t = lane; t < n_bp; t += 32) {
      int32_t b = (int32_t)((1.0f + simt::sqrt_f(1.0f + 8.0f * (float)t)) * 0.5f);       // t = b(b-1)/2 + a, a < b
      while (b * (b - 1) / 2 > t) b--;
      while ((b + 1) * b / 2 <= t) b++;
      const int32_t a = t - b * (b - 1) / 2;
      const int32_t d = lookup_dist(c, sm->bid[a], sm->bid[b]);
      if (d >= 0)
        for (int32_t e = 1; e <= 3; e++)
          if (d <= 2 * e) { simt::atomic_or(&sm->near[e - 1][a], 1u << b); simt::atomic_or(&sm->near[e - 1][b], 1u << a); }
    }
    simt::sync(FULL);
    if (lane < B) { near1 = sm->near[0][lane]; near2 = sm->near[1][lane]; near3 = sm->near[2][lane]; }
  }

  // ---- cluster state
  int32_t cid = -1, pos = 0, keyb = -1, keyn = -1;

#define PCB_ROW_ADD(dst, src)                                                                   \
  do {                                                                                          \
    sm->sum[(dst) * SMALL_STRIDE + lane] += sm->sum[(src) * SMALL_STRIDE + lane];               \
    sm->cnt[(dst) * SMALL_STRIDE + lane] = (uint16_t)(sm->cnt[(dst) * SMALL_STRIDE + lane] + sm->cnt[(src) * SMALL_STRIDE + lane]); \
  } while (0)
  // addLink(R1, R2) (:185-220) with uniform arguments: C1, C2 the clusters of the two reads (-1: none);
  // (NKB, NKN) the creation key if a new cluster is opened
#define PCB_LINK(R1, R2, C1, C2, NKB, NKN)                                                      \
  do {                                                                                          \
    if ((C1) >= 0 && (C2) >= 0) {                                                               \
      const uint32_t m1_ = simt::ballot(FULL, cid == (C1));                                     \
      const int src_ = simt::ffs(m1_) - 1;                                                      \
      const int32_t kb_ = simt::shfl(FULL, keyb, src_), kn_ = simt::shfl(FULL, keyn, src_);     \
      if (cid == (C2)) { pos += simt::popc(m1_); cid = (C1); keyb = kb_; keyn = kn_; }          \
      PCB_ROW_ADD((C1), (C2));                                                                  \
    } else if ((C1) >= 0) {                                                                     \
      const uint32_t m1_ = simt::ballot(FULL, cid == (C1));                                     \
      const int src_ = simt::ffs(m1_) - 1;                                                      \
      const int32_t kb_ = simt::shfl(FULL, keyb, src_), kn_ = simt::shfl(FULL, keyn, src_);     \
      if (lane == (R2)) { pos = simt::popc(m1_); cid = (C1); keyb = kb_; keyn = kn_; }          \
      PCB_ROW_ADD((C1), (R2));                                                                  \
    } else if ((C2) >= 0) {                                                                     \
      const uint32_t m2_ = simt::ballot(FULL, cid == (C2));                                     \
      const int src_ = simt::ffs(m2_) - 1;                                                      \
      const int32_t kb_ = simt::shfl(FULL, keyb, src_), kn_ = simt::shfl(FULL, keyn, src_);     \
      if (lane == (R1)) { pos = simt::popc(m2_); cid = (C2); keyb = kb_; keyn = kn_; }          \
      PCB_ROW_ADD((C2), (R1));                                                                  \
    } else {                                                                                    \
      if (lane == (R1)) { cid = (R1); pos = 0; keyb = (NKB); keyn = (NKN); }                    \
      if (lane == (R2)) { cid = (R1); pos = 1; keyb = (NKB); keyn = (NKN); }                    \
      PCB_ROW_ADD((R1), (R2));                                                                  \
    }                                                                                           \
  } while (0)

  // ---- seed clustering, bucket by bucket (:385-400)
  for (int32_t bi = 0; bi < B; bi++) {
    const int32_t lo = simt::shfl(FULL, mystart, bi), hi = lo + simt::shfl(FULL, mysz, bi);
    const int32_t bucket = simt::shfl(FULL, myb, bi);
    if (hi - lo < 2) continue;
    int32_t colcnt = 0, colsum = 0;
    uint32_t colall = 0, mymask = 0;
    for (int32_t x = lo; x < hi; x++) {
      const uint32_t cv = sm->cnt[x * SMALL_STRIDE + lane];
      colall += cv; colcnt += (int32_t)(cv & 0xFFu); colsum += sm->sum[x * SMALL_STRIDE + lane];
      const uint32_t pm = simt::ballot(FULL, (cv & 0xFFu) != 0u);           // mutations read x carries with Q > 0
      if (lane == x) mymask = pm;
    }
    const uint32_t kept = simt::ballot(FULL, colcnt > 1 || colsum > minQual);     // filterSeqError over the bucket (:350-356)
    mymask &= kept;
    const bool inb = lane >= lo && lane < hi;
    const uint32_t m0 = simt::shfl(FULL, mymask, lo);
    if (simt::ballot(FULL, inb && mymask != m0) == 0u) {
      // Every read of the bucket carries the same kept calls: every pair has Jaccard g/g, so either no pair links or
      // the loop below links (r1, r0), then appends r2, r3, ...: one cluster r1, r0, r2, ..., created first in this bucket,
      // whose sums are the column sums just taken.
      const int32_t g = simt::popc(m0);
      const int32_t tg = c.prm.jthr[g];
      if (!(g == 0 || (g >= tg && g >= minMatches))) continue;
      if (inb) { cid = lo + 1; pos = lane == lo + 1 ? 0 : lane == lo ? 1 : lane - lo; keyb = bucket; keyn = 0; }
      sm->sum[(lo + 1) * SMALL_STRIDE + lane] = colsum;
      sm->cnt[(lo + 1) * SMALL_STRIDE + lane] = (uint16_t)colall;
      continue;
    }
    int32_t n_created = 0;
    for (int32_t ri = lo + 1; ri < hi; ri++) {
      const uint32_t mi = simt::shfl(FULL, mymask, ri);
      const int32_t inters = simt::popc(mi & mymask), uni = simt::popc(mi | mymask);        // calcJaccard on presence (:358-367)
      const int32_t tu = jthr[uni];
      const bool ok = lane >= lo && lane < ri && (uni == 0 || (inters >= tu && inters >= minMatches));
      uint32_t J = simt::ballot(FULL, ok);
      int32_t ci = -1;                                             // read i is in no cluster when its row of pairs starts
      while (J) {
        const uint32_t live = ci < 0 ? J : (J & simt::ballot(FULL, cid != ci));   // addLink inside one cluster is a no-op
        if (!live) break;
        const int j = simt::ffs(live) - 1;
        J &= ~((2u << j) - 1u);
        const int32_t cj = simt::shfl(FULL, cid, j);
        const bool fresh = ci < 0 && cj < 0;
        PCB_LINK(ri, j, ci, cj, bucket, n_created);
        if (fresh) n_created++;
        ci = simt::shfl(FULL, cid, ri);
      }
    }
  }

  // ---- main clustering: the component's rows in visiting order (:408-455)
  int32_t n_links = 0, n_tests = 0;
  for (int32_t p = p_lo; p < p_hi; p++) {
    const int32_t ba = c.cpair_q[p], bb = c.cpair_r[p], ed = c.cpair_ed[p];
    const int ia = simt::ffs(simt::ballot(FULL, myb == ba)) - 1, ib = simt::ffs(simt::ballot(FULL, myb == bb)) - 1;
    const int32_t xlo = simt::shfl(FULL, mystart, ia), nx = simt::shfl(FULL, mysz, ia);
    const int32_t ylo = simt::shfl(FULL, mystart, ib), ny = simt::shfl(FULL, mysz, ib);
    const uint32_t nearE = ed == 1 ? near1 : ed == 2 ? near2 : near3;
    // cluster pairs that failed since the last merge (within this row): the last four, -1 never matches an id pair
    int32_t ma0 = 0, mb0 = 0, ma1 = 0, mb1 = 0, ma2 = 0, mb2 = 0, ma3 = 0, mb3 = 0, memo_n = 0;
    const int32_t total = nx * ny;
    for (int32_t base = 0; base < total; base += 32) {
      const int32_t t = base + lane;
      const bool valid = t < total;
      const int32_t x = valid ? xlo + t / ny : 0, y = valid ? ylo + t % ny : 0;
      const int32_t lx = simt::shfl(FULL, lex, x), ly = simt::shfl(FULL, lex, y);
      int32_t rid1 = x, rid2 = y;                                  // makePID order (:152-157)
      if (!(lx < ly)) { rid1 = y; rid2 = x; }
      int cursor = 0;
      for (;;) {
        const int32_t c1 = simt::shfl(FULL, cid, rid1), c2 = simt::shfl(FULL, cid, rid2);
        const int32_t ida = c1 >= 0 ? c1 : ~rid1, idb = c2 >= 0 ? c2 : ~rid2;
        bool live = valid && lane >= cursor && !(c1 >= 0 && c1 == c2);            // :424
        if ((memo_n > 0 && ma0 == ida && mb0 == idb) || (memo_n > 1 && ma1 == ida && mb1 == idb) ||
            (memo_n > 2 && ma2 == ida && mb2 == idb) || (memo_n > 3 && ma3 == ida && mb3 == idb)) live = false;
        const uint32_t L = simt::ballot(FULL, live);
        if (!L) break;
        const int l = simt::ffs(L) - 1;
        cursor = l + 1;
        const int32_t R1 = simt::shfl(FULL, rid1, l), R2 = simt::shfl(FULL, rid2, l);
        const int32_t C1 = simt::shfl(FULL, c1, l), C2 = simt::shfl(FULL, c2, l);
        n_tests++;
        // one visit of the loop body (:419-455)
        bool linked = false;
        do {
          const uint32_t m1 = C1 >= 0 ? simt::ballot(FULL, cid == C1) : (1u << R1);
          const uint32_t m2 = C2 >= 0 ? simt::ballot(FULL, cid == C2) : (1u << R2);
          const int32_t s1 = simt::popc(m1), s2 = simt::popc(m2);
          const int32_t mx = s1 > s2 ? s1 : s2, mn = s1 > s2 ? s2 : s1;
          if (mx < (mn << ed)) break;                                            // size rule (:435-440)
          const uint32_t S1 = simt::reduce_or(FULL, ((m1 >> lane) & 1u) ? (1u << bkt) : 0u);
          const uint32_t S2 = simt::reduce_or(FULL, ((m2 >> lane) & 1u) ? (1u << bkt) : 0u);
          const bool far = lane < B && ((S1 >> lane) & 1u) && (S2 & ~nearE) != 0u;
          if (simt::ballot(FULL, far) != 0u) break;                              // transitive rule (:442-445)
          const int32_t row1 = C1 >= 0 ? C1 : R1, row2 = C2 >= 0 ? C2 : R2;
          const int32_t v1 = sm->sum[row1 * SMALL_STRIDE + lane], v2 = sm->sum[row2 * SMALL_STRIDE + lane];
          const int32_t o1 = sm->cnt[row1 * SMALL_STRIDE + lane] & 0xFF, o2 = sm->cnt[row2 * SMALL_STRIDE + lane] & 0xFF;
          const bool keep = o1 + o2 > 1 || v1 + v2 > minQual;                    // filterSeqError across both clusters
          const int32_t inters = simt::popc(simt::ballot(FULL, keep && v1 > 0 && v2 > 0));
          const int32_t uni = simt::popc(simt::ballot(FULL, keep && (v1 > 0 || v2 > 0)));
          if (!(uni == 0 || inters >= c.prm.jthr[uni])) break;                      // genotype rule (:447-452)
          PCB_LINK(R1, R2, C1, C2, -1, -1);
          linked = true;
        } while (0);
        if (linked) { n_links++; memo_n = 0; }
        else {
          ma3 = ma2; mb3 = mb2; ma2 = ma1; mb2 = mb1; ma1 = ma0; mb1 = mb0;
          ma0 = C1 >= 0 ? C1 : ~R1; mb0 = C2 >= 0 ? C2 : ~R2;
          if (memo_n < 4) memo_n++;
        }
      }
    }
  }

  // ---- consensus and rows (:464-477, :515-537, :546-567)
  // call region: as many entries as the component's reads carry (an upper bound of its calls), taken from the shared
  // cursor when the caller lays regions out dynamically (call_off == nullptr)
  int64_t call_base = 0;
  if (c.call_off) call_base = c.call_off[comp];
  else {
    const int32_t ne = simt::reduce_add(FULL, g1 - g0);
    int32_t lo32 = 0, hi32 = 0;
    if (lane == 0) { const int64_t at = simt::atomic_add64(c.call_cursor, (int64_t)ne); lo32 = (int32_t)(at & 0xFFFFFFFFll); hi32 = (int32_t)(at >> 32); }
    lo32 = simt::shfl(FULL, lo32, 0); hi32 = simt::shfl(FULL, hi32, 0);
    call_base = ((int64_t)hi32 << 32) | (int64_t)(uint32_t)lo32;
  }
  int32_t up_err_head = -1;
  uint32_t heads = simt::ballot(FULL, cid >= 0 && pos == 0);
  while (heads) {
    const int h = simt::ffs(heads) - 1;
    heads &= heads - 1u;
    const int32_t ch = simt::shfl(FULL, cid, h);
    const uint32_t mem = simt::ballot(FULL, cid == ch);
    const int32_t size = simt::popc(mem);
    const bool in = ((mem >> lane) & 1u) != 0u;
    const int32_t r_head = simt::shfl(FULL, r, h);
    if (in) { c.read_root[r] = r_head; c.read_pos[r] = pos; if (lane != h) c.row_size[r] = 0; }
    // genotype vote (:468-470): Q where a member names the mutation, -minQual where it does not
    const int32_t sv = sm->sum[ch * SMALL_STRIDE + lane], any = sm->cnt[ch * SMALL_STRIDE + lane] >> 8;
    const bool called = lane < M && (int64_t)sv - (int64_t)minQual * (int64_t)(size - any) > 0;
    const uint32_t cm = simt::ballot(FULL, called);
    int32_t at = 0;                                               // calls ordered by the (digit key, string) rank (:475)
    for (uint32_t tm = cm; tm; tm &= tm - 1u) {
      const int s = simt::ffs(tm) - 1;
      const int32_t rk = simt::shfl(FULL, myrank, s);
      if (rk < myrank || (rk == myrank && s < lane)) at++;
    }
    if (called) c.call_mut[call_base + at] = mymut;
    // majority barcode / uptag (:515-532)
    int32_t cons[2] = {0, 0};
    for (int w = 0; w < 2; w++) {
      if (w == 1 && !c.up_id) { cons[1] = cons[0]; break; }
      const int32_t id = w == 0 ? bc : up;
      if (w == 1 && simt::ballot(FULL, in && id < 0) != 0u) up_err_head = r_head;     // clustered read without uptag: KeyError (:519)
      const int32_t v = id < 0 ? 0x7ffffffe : id;
      int32_t votes = 0;
      if (in) votes = simt::popc(simt::match_any(mem, v));
      const int32_t best = simt::reduce_max(FULL, votes);
      const uint32_t top = simt::ballot(FULL, in && votes == best);
      if (simt::popc(top) == best) cons[w] = simt::shfl(FULL, v, simt::ffs(top) - 1);      // one barcode holds all the top votes
      else if (size == 2) cons[w] = simt::shfl(FULL, id, h);
      else cons[w] = PCB_NEEDS_ALIGNMENT;
    }
    if (lane == h) {
      c.row_size[r] = size; c.row_key[r] = ((int64_t)keyb << 32) | (int64_t)keyn;
      c.row_call_off[r] = call_base; c.row_call_n[r] = simt::popc(cm);
      c.row_bc[r] = cons[0]; c.row_up[r] = cons[1];
    }
    call_base += simt::popc(cm);
  }
  {
    // unclustered reads (:560-567): calls with Q >= minQual in input order
    const bool single = has && cid < 0;
    int32_t ncs = 0;
    if (single) for (int32_t e = g0; e < g1; e++) if (c.geno_q[e] >= minQual) ncs++;
    int32_t incl = ncs;
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t tv = simt::shfl(FULL, incl, lane >= o ? lane - o : lane);
      if (lane >= o) incl += tv;
    }
    if (single) {
      const int64_t off = call_base + (int64_t)(incl - ncs);
      c.read_root[r] = r; c.read_pos[r] = 0; c.row_size[r] = 1; c.row_key[r] = -1;
      c.row_bc[r] = bc; c.row_up[r] = c.up_id ? up : bc;
      c.row_call_off[r] = off; c.row_call_n[r] = ncs;
      int32_t k = 0;
      for (int32_t e = g0; e < g1; e++) if (c.geno_q[e] >= minQual) c.call_mut[off + k++] = c.geno_mut[e];
    }
  }
  if (lane == 0) {
    if (n_links) simt::atomic_add(&tally[0], n_links);
    if (n_tests) simt::atomic_add(&tally[1], n_tests);
    if (up_err_head >= 0) { c.status[0] = 3; c.status[1] = up_err_head; }
  }
#undef PCB_LINK
#undef PCB_ROW_ADD
  return 0;
}

}  // namespace pcb
