// probe.cuh -- the pair search's probe for barcodes of at most 32 nt within one set (the clustering path): candidate
// generation from the seed index, two cheap exact filters, and the Myers/Hyyro bit-vector distance with an early exit,
// organised so that the lanes of a warp are always busy with a pair of their own.
//
//   filter     every lane walks the bin of its probe (query i, piece t, shift s); a target j is kept when i is the
//              pair's query side (every unordered pair has exactly one), the lengths differ by at most k, and the
//              base-composition bound holds: an edit changes the count of at most two bases by one each, so
//              dist >= max(surplus, deficit) of the four base counts (two popc-built 4-byte vectors, two SIMD
//              subtractions on a 4-byte sketch of the target that sits in the index entry itself, so a candidate
//              costs 8 streamed bytes and no gather).  Survivors go to a per-warp queue.
//   scoring    a lane that has no pair pops one from the queue; all lanes then advance THEIR pair by a few text
//              columns; at the checkpoint a finished pair emits its hit and a pair whose cell on the final diagonal
//              already exceeds k is dropped (D[m][n] >= D[j + m - n][j], Ukkonen's cut-off) -- either way the lane is
//              free for the next pair, so rejected pairs cost a few columns and the warp stays full.
//   direction  the seed that produced the candidate matches exactly, so the errors sit on the other side of it: the
//              recurrence runs over the reversed strings when the longer unseeded part is the tail (dist(x, y) =
//              dist(rev x, rev y)), which puts the mismatches first and the cut-off early.
// Same distances and the same hit set as the plain kernel (edit_pairs.cu), which stays for 33..64-nt barcodes and for
// the two-set matcher.  Written against simt.h: tests/host_emul runs this source on an emulated warp (unit tests).
#pragma once
#include "myers.cuh"
#include "simt.h"

namespace pcb {

struct Planes16 { uint64_t x, y; };       // layout of ulonglong2: plane L, plane H

struct SeedGeom {
  int32_t q;         // seed length (0: single bin)
  int32_t npos;      // target positions indexed: 0 .. npos-1
  int32_t k;
  uint32_t nbins;
  uint64_t pos_mask; // positions a probe can ask for (piece start + shift of some query length): only those are indexed
  int32_t cbits;     // every bin is split into 2^cbits sub-bins by the low bits of the target (its CLASS), see owns_pair()
  int32_t rbits;     // the ownership rule of the whole search (all its indexes): 0 = parity, else classes of rbits bits;
                     // an index with sub-bins has cbits == rbits, one without walks whole bins under the same rule
  // Two indexes (edit_pairs.cu): barcodes shorter than q * (k + 1) cannot probe the main index (their pieces are shorter
  // than its seeds).  A pair of a SHORT and a long barcode is therefore owned by the long one, which finds it through the
  // main index like any other (the pigeonhole argument holds from either side); the second index only holds the short
  // barcodes and only they probe it -- it costs what they cost, not a second pass over every target.
  int32_t short_below;   // main index: a target shorter than this belongs to the (long) query whatever the rule says; 0: none
  int32_t only_below;    // second index: only targets shorter than this are indexed; 0: all
};

// Which side of an unordered pair {i, j} probes for it.  cbits == 0: the parity rule of myers.cuh.  cbits > 0 (big
// libraries, long bins): class = low bits of the index; between classes the smaller class is the query side, inside a
// class the parity rule on the remaining bits decides.  A bin's entries are grouped by class, so a query of class c starts its walk at sub-bin c: what it
// skips are exactly pairs it does not own -- on average (C - 1) / 2C of every bin (44 % with 8 classes), where the
// parity rule reads a whole bin to keep half of it.  Classes interleave, so any index range (a shard) holds all alike.
PCB_HD bool owns_pair(const SeedGeom &g, int32_t i, int32_t j, int32_t target_len) {
  if (target_len < g.short_below) return true;
  if (g.rbits == 0) return j != i && is_query_side(i, j);
  const int32_t mask = (1 << g.rbits) - 1;
  // (inside a class the parity rule on the remaining bits: "the smaller index" would hand the first of two query shards
  // every same-class pair that straddles them -- 55 : 45 at 4 GPUs, profiles/r02w_bench_n4.json)
  return (j & mask) != (i & mask) ? (j & mask) > (i & mask) : (j != i && is_query_side(i >> g.rbits, j >> g.rbits));
}
// sub-bin of target j (of length len_j) in bin `bin`; short targets sit in the last class, where every probe passes
PCB_HD uint32_t seed_slot(const SeedGeom &g, uint32_t bin, int32_t j, int32_t len_j) {
  const uint32_t mask = (1u << g.cbits) - 1u;
  return (bin << g.cbits) | (len_j < g.short_below ? mask : ((uint32_t)j & mask));
}
// ... and where the walk of query i starts in it
PCB_HD uint32_t seed_slot_of_query(const SeedGeom &g, uint32_t bin, int32_t i) {
  return (bin << g.cbits) | ((uint32_t)i & ((1u << g.cbits) - 1u));
}

PCB_HD uint32_t seed_bin(const SeedGeom g, uint64_t L, uint64_t H, int src_pos, int bin_pos) {
  if (g.q == 0) return 0u;
  return ((uint32_t)bin_pos << (2 * g.q)) | qgram_key(L, H, src_pos, g.q);
}

struct alignas(8) SeedEnt { int32_t j; uint32_t sk; };      // an index entry: the target and its sketch (so the filters need no gather)

struct ProbeArgs {
  const Planes16 *planes;          // the barcodes (queries and targets are the same set)
  const int32_t *len;
  const uint32_t *sketch;          // per barcode: length << 24 | #T << 16 | #G << 8 | #C -- all the filters read
  int32_t q_begin, q_end;          // this call's queries
  SeedGeom g;
  const uint32_t *bin_off;
  const SeedEnt *ent;              // the bins: (target, its sketch), streamed
  uint64_t *hit_key;
  uint32_t *hit_d;
  unsigned long long cap;
  int key_shift;
  unsigned long long *counters;    // [0] hits, [1] pairs that entered the recurrence, [2] cells computed, [3] next block of probes,
                                   // [4] bin entries walked, [5] entries that passed the query-side rule (the filters' input)
  int32_t len_lo, len_hi;          // query lengths this index serves
  const int32_t *qlist;            // optional: the queries of this launch as a list (the few short barcodes that probe the second
  const int32_t *qlist_n;          //   index), its length on the device; nullptr: the range [q_begin, q_end)
};

constexpr int PROBE_QUEUE = 96;    // survivors waiting for a lane (at most 63 + 32)
constexpr int PROBE_CHUNK = 5;     // text columns between checkpoints
struct ProbeQueue {
  int32_t qi[PROBE_QUEUE], tj[PROBE_QUEUE];     // query, target (bit 31 of tj: run the recurrence over the reversed strings)
};

// Staging of long bins (the north star's "candidate barcodes staged through shared memory, TMA bulk copies where the
// bucket is large"): when the warp walks bins together, their entries are fetched PROBE_STAGE entries at a time by
// cp.async.bulk into one of two buffers (completion on an mbarrier), the next chunk in flight while the lanes filter
// the current one.  A compile-time variant of probe_warp, measured against the direct streaming loads
// (profiles/r02k_probe_staging.md).
constexpr int PROBE_STAGE = 256;   // entries per chunk (2 KB)
struct alignas(16) ProbeStage {
  SeedEnt ent[2][PROBE_STAGE];
  uint64_t bar[2];
};

PCB_HD uint32_t barcode_sketch(uint32_t L, uint32_t H, int32_t n) {
  return ((uint32_t)n << 24) | ((uint32_t)popc32(L & H) << 16) | ((uint32_t)popc32(~L & H) << 8) | (uint32_t)popc32(L & ~H);
}
// counts of A, C, G, T as four bytes, from a sketch
SIMT_FN uint32_t sketch_counts(uint32_t sk) {
  const uint32_t n = sk >> 24, nc = sk & 0xFFu, ng = (sk >> 8) & 0xFFu, nt = (sk >> 16) & 0xFFu;
  return (n - nc - ng - nt) | ((sk & 0xFFFFFFu) << 8);
}

template <bool STAGED>
SIMT_FN void probe_warp(const ProbeArgs &a, int lane, ProbeQueue *qu, ProbeStage *stg) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lt = (1u << lane) - 1u;
  const SeedGeom g = a.g;
  const int k = g.k;
  const int n_shift = g.q == 0 ? 1 : 2 * k + 1;
  const int per_query = g.q == 0 ? 1 : (k + 1) * n_shift;
  const unsigned long long n_probe = (unsigned long long)(a.qlist ? *a.qlist_n : a.q_end - a.q_begin) * per_query;
  const unsigned long long n_blocks = (n_probe + 31) / 32;
  uint32_t n_started = 0, n_cells = 0, n_walked = 0, n_sided = 0;      // per lane; widened when the warp is done
  int cnt = 0;                       // queued survivors (uniform)
  bool exhausted = false;            // no more probe blocks (uniform)
  // my probe; and, when the block's bins are long, the probe the whole warp is walking together
  uint32_t lo = 0, hi = 0;
  int32_t qi = -1, qm = 0;
  uint32_t qcounts = 0, qrev = 0;
  int mode = 0;                      // 0: fetch a block of probes; 1: every lane walks its own bin; 2: the warp walks bin after bin
  int coop_p = -1;
  uint32_t c_at = 0, c_hi = 0, c_counts = 0, c_rev = 0;
  int32_t c_qi = 0, c_qm = 0;
  // staged walk: the chunk generator (bins of the block's probes, in lane order) and the two buffers
  uint32_t g_left = 0, g_at = 0, g_hi = 0, g_lo = 0;
  int g_p = 0, cur = 0;
  uint32_t s_base[2] = {0, 0}, s_n[2] = {0, 0}, s_lo[2] = {0, 0}, s_hi[2] = {0, 0}, s_phase = 0, s_off = 0;
  int s_p[2] = {-1, -1};
  if (STAGED) {
    if (lane == 0) { simt::bar_init(&stg->bar[0], 1); simt::bar_init(&stg->bar[1], 1); }
    simt::sync(FULL);
  }
  // next chunk of the bins into buffer b (uniform; lane 0 issues the copy); s_p[b] = -1 when the bins are used up
  auto fill = [&](int b) {
    if (g_at >= g_hi) {
      if (g_left == 0u) { s_p[b] = -1; return; }
      g_p = simt::ffs(g_left) - 1;
      g_left &= g_left - 1u;
      g_lo = simt::shfl(FULL, lo, g_p); g_hi = simt::shfl(FULL, hi, g_p);
      g_at = g_lo & ~1u;                                   // 16-byte aligned source
    }
    uint32_t n = ((g_hi - g_at) + 1u) & ~1u;               // 16-byte multiple (the index is padded by one entry)
    if (n > (uint32_t)PROBE_STAGE) n = PROBE_STAGE;
    if (lane == 0) simt::bulk_load(stg->ent[b], &a.ent[g_at], n * (uint32_t)sizeof(SeedEnt), &stg->bar[b]);
    s_base[b] = g_at; s_n[b] = n; s_lo[b] = g_lo; s_hi[b] = g_hi; s_p[b] = g_p;
    g_at += n;
  };
  // my pair
  bool active = false;
  uint32_t Pv = 0, Mv = 0, rl = 0, rh = 0, pL = 0, pH = 0;
  int32_t pj = 0, pn = 0, pm = 0, pi = 0, ptj = 0;
  for (;;) {
    // ---- filter: until a warp's worth of survivors waits (or the probes run out)
    while (cnt < 32 && !exhausted) {
      if (mode == 0) {
        unsigned long long blk = 0;
        if (lane == 0) blk = simt::atomic_add_u64(&a.counters[3], 1ull);
        const uint32_t b_lo = simt::shfl(FULL, (uint32_t)(blk & 0xFFFFFFFFull), 0), b_hi = simt::shfl(FULL, (uint32_t)(blk >> 32), 0);
        blk = ((unsigned long long)b_hi << 32) | b_lo;
        if (blk >= n_blocks) { exhausted = true; break; }
        const unsigned long long pid = blk * 32 + (unsigned)lane;
        lo = hi = 0; qi = -1;
        if (pid < n_probe) {
          qi = a.qlist ? a.qlist[pid / per_query] : a.q_begin + (int32_t)(pid / per_query);
          const int32_t rem = (int32_t)(pid % per_query);
          const int32_t t = rem / n_shift, s = rem % n_shift - (g.q == 0 ? 0 : k);
          qm = a.len[qi];
          const int32_t st = g.q == 0 ? 0 : (int32_t)((int64_t)t * qm / (k + 1));
          const int32_t p = st + s;
          if (p >= 0 && p < g.npos && qm >= a.len_lo && qm <= a.len_hi) {     // queries of other lengths use the other index
            const Planes16 pl = a.planes[qi];
            const uint32_t bin = seed_bin(g, pl.x, pl.y, st, p);
            lo = a.bin_off[seed_slot_of_query(g, bin, qi)]; hi = a.bin_off[(bin + 1u) << g.cbits];      // my class and the ones above
            qcounts = sketch_counts(a.sketch[qi]);
            // the seed matches exactly: the errors are before it or after it -- start the recurrence on the longer side
            qrev = (st < qm - (st + g.q)) ? 0x80000000u : 0u;
          }
        }
        // long bins (a big library): the warp walks them one after the other, 32 entries per step, coalesced;
        // short bins: every lane walks its own
        mode = simt::reduce_add(FULL, (int32_t)(hi - lo)) >= 32 * 12 ? 2 : 1;
        coop_p = -1; c_at = c_hi = 0;
        if (STAGED && mode == 2) {
          g_left = simt::ballot(FULL, lo < hi); g_at = g_hi = 0;
          cur = 0; s_off = 0;
          fill(0); fill(1);
        }
        continue;
      }
      bool keep = false;
      int32_t j = 0, kq = 0;
      uint32_t krev = 0;
      if (mode == 1) {
        if (simt::ballot(FULL, lo < hi) == 0u) { mode = 0; continue; }
        if (lo < hi) {
          const SeedEnt en = a.ent[lo++];
          j = en.j;
          n_walked++;
          if (owns_pair(g, qi, j, (int32_t)(en.sk >> 24))) {
            n_sided++;
            const uint32_t sk = en.sk;
            const int32_t dl = qm - (int32_t)(sk >> 24);
            if (dl <= k && -dl <= k) {
              const uint32_t tc = sketch_counts(sk);
              const uint32_t up = simt::surplus4(qcounts, tc), down = simt::surplus4(tc, qcounts);
              keep = (int32_t)(up > down ? up : down) <= k;
            }
          }
        }
        kq = qi; krev = qrev;
      } else if (STAGED) {
        if (s_p[cur] < 0) { mode = 0; continue; }            // every bin of the block has been walked
        if (coop_p != s_p[cur]) {                            // (the probe of this chunk: its query and filters)
          coop_p = s_p[cur];
          c_qi = simt::shfl(FULL, qi, coop_p); c_qm = simt::shfl(FULL, qm, coop_p);
          c_counts = simt::shfl(FULL, qcounts, coop_p); c_rev = simt::shfl(FULL, qrev, coop_p);
        }
        if (s_off == 0) simt::bar_wait(&stg->bar[cur], (s_phase >> cur) & 1u);
        const uint32_t e = s_base[cur] + s_off + (uint32_t)lane;
        if (s_off + (uint32_t)lane < s_n[cur] && e >= s_lo[cur] && e < s_hi[cur]) {
          const SeedEnt en = stg->ent[cur][s_off + (uint32_t)lane];
          j = en.j;
          n_walked++;
          if (owns_pair(g, c_qi, j, (int32_t)(en.sk >> 24))) {
            n_sided++;
            const uint32_t sk = en.sk;
            const int32_t dl = c_qm - (int32_t)(sk >> 24);
            if (dl <= k && -dl <= k) {
              const uint32_t tc = sketch_counts(sk);
              const uint32_t up = simt::surplus4(c_counts, tc), down = simt::surplus4(tc, c_counts);
              keep = (int32_t)(up > down ? up : down) <= k;
            }
          }
        }
        kq = c_qi; krev = c_rev;
        s_off += 32u;
        if (s_off >= s_n[cur]) {                             // chunk used up: refill this buffer, go on with the other
          simt::sync(FULL);
          s_phase ^= 1u << cur;
          fill(cur);
          cur ^= 1; s_off = 0;
        }
      } else {
        if (c_at >= c_hi) {                                  // next probe of the block that has entries
          const uint32_t left = simt::ballot(FULL, lo < hi);
          if (left == 0u) { mode = 0; continue; }
          coop_p = simt::ffs(left) - 1;
          c_qi = simt::shfl(FULL, qi, coop_p); c_qm = simt::shfl(FULL, qm, coop_p);
          c_at = simt::shfl(FULL, lo, coop_p); c_hi = simt::shfl(FULL, hi, coop_p);
          c_counts = simt::shfl(FULL, qcounts, coop_p); c_rev = simt::shfl(FULL, qrev, coop_p);
          if (lane == coop_p) lo = hi;
        }
        const uint32_t e = c_at + (uint32_t)lane;
        if (e < c_hi) {
          const SeedEnt en = simt::ld_stream(&a.ent[e]);
          j = en.j;
          n_walked++;
          if (owns_pair(g, c_qi, j, (int32_t)(en.sk >> 24))) {
            n_sided++;
            const uint32_t sk = en.sk;
            const int32_t dl = c_qm - (int32_t)(sk >> 24);
            if (dl <= k && -dl <= k) {
              const uint32_t tc = sketch_counts(sk);
              const uint32_t up = simt::surplus4(c_counts, tc), down = simt::surplus4(tc, c_counts);
              keep = (int32_t)(up > down ? up : down) <= k;
            }
          }
        }
        c_at += 32u;
        kq = c_qi; krev = c_rev;
      }
      const uint32_t mask = simt::ballot(FULL, keep);
      if (keep) {
        const int at = cnt + simt::popc(mask & lt);
        qu->qi[at] = kq; qu->tj[at] = (int32_t)((uint32_t)j | krev);
      }
      cnt += simt::popc(mask);
    }
    simt::sync(FULL);
    // ---- lanes without a pair take one
    {
      const uint32_t idle = simt::ballot(FULL, !active);
      const int n_pop = simt::popc(idle) < cnt ? simt::popc(idle) : cnt;
      if (!active && simt::popc(idle & lt) < n_pop) {
        const int at = cnt - 1 - simt::popc(idle & lt);
        pi = qu->qi[at];
        const uint32_t tj = (uint32_t)qu->tj[at];
        ptj = (int32_t)(tj & 0x7FFFFFFFu);
        const Planes16 tp = a.planes[ptj], pl = a.planes[pi];
        const uint32_t tL = (uint32_t)tp.x, tH = (uint32_t)tp.y;
        pn = a.len[ptj]; pm = a.len[pi];
        if (tj & 0x80000000u) {              // reversed strings: text from its last base, pattern mirrored
          rl = tL << (32 - pn); rh = tH << (32 - pn);
          pL = simt::brev((uint32_t)pl.x) >> (32 - pm); pH = simt::brev((uint32_t)pl.y) >> (32 - pm);
        } else {
          rl = simt::brev(tL); rh = simt::brev(tH);
          pL = (uint32_t)pl.x; pH = (uint32_t)pl.y;
        }
        Pv = 0xFFFFFFFFu; Mv = 0u; pj = 0;
        active = true;
        n_started++;
      }
      cnt -= n_pop;
      simt::sync(FULL);
      if (simt::ballot(FULL, active) == 0u) {
        if (exhausted && cnt == 0) break;
        continue;
      }
    }
    // ---- a few columns of every lane's own pair (myers.cuh: the recurrence; the text column is the sign bit of rl / rh)
    const int32_t pj_before = pj;
#pragma unroll
    for (int c = 0; c < PROBE_CHUNK; c++) {
      if (active && pj < pn) {
        const uint32_t cl = (uint32_t)((int32_t)rl >> 31), ch = (uint32_t)((int32_t)rh >> 31);
        rl += rl; rh += rh;
        const uint32_t Eq = ~((pL ^ cl) | (pH ^ ch));
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint32_t Ph = Mv | ~(Xh | Pv);
        uint32_t Mh = Pv & Xh;
        Ph = Ph + Ph + 1u;
        Mh = Mh + Mh;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
        pj++;
      }
    }
    // ---- checkpoint
    if (active) {
      n_cells += (uint32_t)((pj - pj_before) * pm);
      if (pj >= pn) {
        const uint32_t rows = pm >= 32 ? 0xFFFFFFFFu : ((1u << pm) - 1u);
        const int d = pn + simt::popc(Pv & rows) - simt::popc(Mv & rows);
        if (d <= k) {
          const unsigned long long at = simt::atomic_add_u64(&a.counters[0], 1ull);
          if (at < a.cap) {
            const uint32_t x = ptj < pi ? (uint32_t)ptj : (uint32_t)pi, y = ptj < pi ? (uint32_t)pi : (uint32_t)ptj;
            a.hit_key[at] = ((uint64_t)x << a.key_shift) | y;
            a.hit_d[at] = (uint32_t)d;
          }
        }
        active = false;
      } else {
        // the cell of column pj on the diagonal that ends in (m, n) bounds the distance from below
        int32_t i = pj + pm - pn, extra = 0;
        if (i < 0) { extra = -i; i = 0; }
        if (i > pm) { extra = i - pm; i = pm; }
        const uint32_t below = i >= 32 ? 0xFFFFFFFFu : ((1u << i) - 1u);
        if (pj + simt::popc(Pv & below) - simt::popc(Mv & below) + extra > k) active = false;
      }
    }
  }
  // one set of atomics per warp (a warp's totals fit 32 bits: it is one of ~10^4 warps of a persistent grid)
  const uint32_t w_started = (uint32_t)simt::reduce_add(FULL, (int32_t)n_started), w_cells = (uint32_t)simt::reduce_add(FULL, (int32_t)n_cells);
  const uint32_t w_walked = (uint32_t)simt::reduce_add(FULL, (int32_t)n_walked), w_sided = (uint32_t)simt::reduce_add(FULL, (int32_t)n_sided);
  if (lane == 0 && w_walked) {
    simt::atomic_add_u64(&a.counters[1], w_started); simt::atomic_add_u64(&a.counters[2], w_cells);
    simt::atomic_add_u64(&a.counters[4], w_walked); simt::atomic_add_u64(&a.counters[5], w_sided);
  }
}

}  // namespace pcb
