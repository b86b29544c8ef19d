// edit_pairs.cu -- all pairs of unique barcodes within Levenshtein distance k (rows a2+a3).
//
// Replaces the bowtie2 all-vs-all alignment and src/pacybara_calcEdits.R (src/pacybara.sh:448-468).
//
// Candidate generation (pigeonhole): split the query into k+1 pieces; if lev(query, target) <= k
// one piece occurs unchanged in the target, displaced by at most k.  So:
//   1. index   every q-gram of every barcode into the bin (position, q-gram)      [seed_count/fill]
//   2. probe   for each query piece t and shift s, walk the bin (start_t + s, q-gram of piece t),
//              and score each entry with the bit-parallel Myers kernel             [probe_kernel]
//   3. order   sort the hits by (i, j), drop the duplicates a pair found through several
//              pieces produces                                                     [radix sort + compact]
// q = floor(min length / (k+1)), capped so that the bin table stays small; q = 0 (barcodes shorter
// than k+1) degenerates to one bin, i.e. all pairs.
//
// Roofline: the probe kernel is integer-ALU bound (~14 LOP3/IADD/SHF per text column per pair,
// nothing to feed a tensor core with); its HBM traffic is the bin entries (4 B) plus the 16-byte
// planes of each scored target, mostly served by L2.  See DESIGN.md for the op and byte counts.
#include "common.cuh"
#include "myers.cuh"
#include "probe.cuh"

namespace pcb {

// the targets indexed are [t_begin, t_begin + U)
__global__ void seed_count_kernel(const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int32_t t_begin, int32_t U,
                                  SeedGeom g, uint32_t *__restrict__ bin_cnt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)U * g.npos) return;
  int32_t j = t_begin + (int32_t)(i / g.npos), p = (int32_t)(i % g.npos);
  if (len[j] > 64) return;                       // an opaque barcode: not a target
  if (g.only_below && len[j] >= g.only_below) return;
  if (g.q > 0 && (p + g.q > len[j] || !((g.pos_mask >> p) & 1ull))) return;
  ulonglong2 pl = planes[j];
  atomicAdd(&bin_cnt[seed_slot(g, seed_bin(g, pl.x, pl.y, p, p), j, len[j])], 1u);
}

__global__ void seed_fill_kernel(const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int32_t t_begin, int32_t U,
                                 SeedGeom g, uint32_t *__restrict__ cursor, int32_t *__restrict__ ent) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)U * g.npos) return;
  int32_t j = t_begin + (int32_t)(i / g.npos), p = (int32_t)(i % g.npos);
  if (len[j] > 64) return;
  if (g.only_below && len[j] >= g.only_below) return;
  if (g.q > 0 && (p + g.q > len[j] || !((g.pos_mask >> p) & 1ull))) return;
  ulonglong2 pl = planes[j];
  uint32_t at = atomicAdd(&cursor[seed_slot(g, seed_bin(g, pl.x, pl.y, p, p), j, len[j])], 1u);
  ent[at] = len[j] < g.short_below ? (int32_t)((uint32_t)j | 0x80000000u) : j;      // sign bit: a short target (owned by the query)
}

// the same with (target, sketch) entries: the index of the lane-per-pair probe (probe.cuh)
__global__ void seed_fill8_kernel(const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int32_t t_begin, int32_t U,
                                  SeedGeom g, uint32_t *__restrict__ cursor, SeedEnt *__restrict__ ent) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)U * g.npos) return;
  int32_t j = t_begin + (int32_t)(i / g.npos), p = (int32_t)(i % g.npos);
  if (len[j] > 64) return;
  if (g.only_below && len[j] >= g.only_below) return;
  if (g.q > 0 && (p + g.q > len[j] || !((g.pos_mask >> p) & 1ull))) return;
  ulonglong2 pl = planes[j];
  uint32_t at = atomicAdd(&cursor[seed_slot(g, seed_bin(g, pl.x, pl.y, p, p), j, len[j])], 1u);
  SeedEnt e;
  e.j = j; e.sk = barcode_sketch((uint32_t)pl.x, (uint32_t)pl.y, len[j]);
  ent[at] = e;
}

// Device-side Myers for one-word barcodes, tuned for the ALU/FMA pipe split of sm_100a: the text
// planes are bit-reversed once so that the current column's bits sit in the sign position (one
// arithmetic shift gives the all-ones/all-zero mask, one add moves to the next column), and the
// two carries-in are written as adds so that ptxas can place them on the FMA pipe (IMAD.IADD).
// Same recurrence and same result as pcb::myers32 (myers.cuh), which the host tests pin to the DP.
__device__ __forceinline__ int myers32_dev(uint32_t pL, uint32_t pH, int m, uint32_t tL, uint32_t tH, int n, uint32_t other = 0u) {
  uint32_t Pv = 0xFFFFFFFFu, Mv = 0u;
  uint32_t rl = __brev(tL), rh = __brev(tH);
#pragma unroll 5
  for (int j = 0; j < n; j++) {
    uint32_t cl = (uint32_t)((int32_t)rl >> 31), ch = (uint32_t)((int32_t)rh >> 31);
    rl += rl; rh += rh;
    uint32_t Eq = ~((pL ^ cl) | (pH ^ ch)) & ~other;
    uint32_t Xv = Eq | Mv;
    uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint32_t Ph = Mv | ~(Xh | Pv);
    uint32_t Mh = Pv & Xh;
    asm("mad.lo.u32 %0, %0, 2, 1;" : "+r"(Ph));   // (Ph << 1) | 1 as one IMAD (FMA pipe)
    Mh = Mh + Mh;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
  }
  uint32_t rows = m >= 32 ? 0xFFFFFFFFu : ((1u << m) - 1u);
  return n + __popc(Pv & rows) - __popc(Mv & rows);
}

// Probe kernel: one warp per block of 32 probes (query i, piece t, shift s), candidates compacted
// through a per-warp shared-memory queue so that the Myers recurrence always runs on full warps.
//   filter phase   every lane walks the bin of its own probe (coalesced 4-byte entries), keeps the
//                  targets j for which i is the query side of {i, j}, and appends (i, j) to the
//                  warp's queue (ballot + popc prefix);
//   score phase    whenever 32 pairs are queued, each lane takes one: 16-byte plane loads for both
//                  barcodes (L2-resident table), length filter, bit-parallel distance, hit emission.
// Blocks of probes are handed out by an atomic counter (bins differ in size).
// counters: [0] hits emitted, [1] pairs scored, [2] cells (sum len_i * len_j), [3] next block of probes.
constexpr int PROBE_WARPS = 8;

// SELF: queries and targets are the same set (the clustering path); otherwise query i comes from (qplanes, qlen)
// and the hit key is (query, target) as it stands (f3: barcode -> library matching).
template <bool WIDE, bool SELF>
__device__ __forceinline__ void score_pair(const ulonglong2 *__restrict__ qplanes, const int32_t *__restrict__ qlen,
                                           const uint64_t *__restrict__ qother, const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int2 pr,
                                           int k, uint64_t *__restrict__ hit_key, uint32_t *__restrict__ hit_d,
                                           unsigned long long cap, int key_shift, unsigned long long *counters,
                                           unsigned long long &scored, unsigned long long &cells) {
  int32_t i = pr.x, j = pr.y;
  int32_t m = qlen[i], n = len[j];
  int32_t dl = m - n;
  if (dl > k || -dl > k) return;
  ulonglong2 a = qplanes[i], b = planes[j];
  const uint64_t x = (!SELF && qother) ? qother[i] : 0ull;                // query characters that are no base never match
  int d = WIDE ? myers64(a.x, a.y, m, b.x, b.y, n, x)
               : myers32_dev((uint32_t)a.x, (uint32_t)a.y, m, (uint32_t)b.x, (uint32_t)b.y, n, (uint32_t)x);
  scored++; cells += (unsigned long long)(m * n);
  if (d <= k) {
    unsigned long long at = atomicAdd(&counters[0], 1ull);
    if (at < cap) {
      uint32_t x = (SELF && j < i) ? j : i, y = (SELF && j < i) ? i : j;
      hit_key[at] = ((uint64_t)x << key_shift) | y;
      hit_d[at] = (uint32_t)d;
    }
  }
}

template <bool WIDE, bool SELF>
__global__ void __launch_bounds__(PROBE_WARPS * 32) probe_kernel(const ulonglong2 *__restrict__ qplanes, const int32_t *__restrict__ qlen,
                                                                  const uint64_t *__restrict__ qother,
                                                                  const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len,
                                                                  int32_t q_begin, int32_t q_end, SeedGeom g,
                                                                  const uint32_t *__restrict__ bin_off, const int32_t *__restrict__ ent,
                                                                  uint64_t *__restrict__ hit_key, uint32_t *__restrict__ hit_d,
                                                                  unsigned long long cap, int key_shift, unsigned long long *counters,
                                                                  int32_t len_lo, int32_t len_hi) {
  __shared__ int2 queue[PROBE_WARPS][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int n_shift = g.q == 0 ? 1 : 2 * g.k + 1;
  const int per_query = g.q == 0 ? 1 : (g.k + 1) * n_shift;
  const unsigned long long n_probe = (unsigned long long)(q_end - q_begin) * per_query;
  const unsigned long long n_blocks = (n_probe + 31) / 32;
  unsigned long long scored = 0, cells = 0;
  int cnt = 0;                                   // queued pairs (warp-uniform, < 32 between steps)
  for (;;) {
    unsigned long long blk = 0;
    if (lane == 0) blk = atomicAdd(&counters[3], 1ull);
    blk = __shfl_sync(0xFFFFFFFFu, blk, 0);
    if (blk >= n_blocks) break;
    unsigned long long pid = blk * 32 + lane;
    uint32_t lo = 0, hi = 0;
    int32_t qi = -1;
    if (pid < n_probe) {
      qi = q_begin + (int32_t)(pid / per_query);
      int32_t rem = (int32_t)(pid % per_query);
      int32_t t = rem / n_shift, s = rem % n_shift - (g.q == 0 ? 0 : g.k);
      int32_t m = qlen[qi];
      int32_t st = g.q == 0 ? 0 : (int32_t)((int64_t)t * m / (g.k + 1));
      int32_t p = st + s;
      if (p >= 0 && p < g.npos && m >= len_lo && m <= len_hi) {     // queries of other lengths use the other index
        ulonglong2 a = qplanes[qi];
        uint32_t bin = seed_bin(g, a.x, a.y, st, p);
        // a seed that holds a non-base cannot occur in a target; the error-free piece the pigeonhole promises has none
        bool clean = SELF || !qother || g.q == 0 || ((qother[qi] >> st) & ((1ull << g.q) - 1ull)) == 0ull;
        if (clean) { lo = bin_off[bin]; hi = bin_off[bin + 1]; }
      }
    }
    while (__any_sync(0xFFFFFFFFu, lo < hi)) {
      bool keep = false;
      int32_t j = 0;
      if (lo < hi) {
        j = ent[lo++];
        const bool t_short = j < 0;                          // (only the self-join's main index marks short targets)
        j &= 0x7FFFFFFF;
        keep = !SELF || t_short || (j != qi && is_query_side(qi, j));
      }
      uint32_t mask = __ballot_sync(0xFFFFFFFFu, keep);
      if (keep) queue[warp][cnt + __popc(mask & lt_mask)] = make_int2(qi, j);
      cnt += __popc(mask);
      __syncwarp();
      if (cnt >= 32) {
        score_pair<WIDE, SELF>(qplanes, qlen, qother, planes, len, queue[warp][lane], g.k, hit_key, hit_d, cap, key_shift, counters, scored, cells);
        int rest = cnt - 32;
        int2 carry = make_int2(0, 0);
        if (lane < rest) carry = queue[warp][32 + lane];
        __syncwarp();
        if (lane < rest) queue[warp][lane] = carry;
        cnt = rest;
        __syncwarp();
      }
    }
  }
  if (lane < cnt) score_pair<WIDE, SELF>(qplanes, qlen, qother, planes, len, queue[warp][lane], g.k, hit_key, hit_d, cap, key_shift, counters, scored, cells);
  for (int o = 16; o > 0; o >>= 1) {
    scored += __shfl_down_sync(0xFFFFFFFFu, scored, o);
    cells += __shfl_down_sync(0xFFFFFFFFu, cells, o);
  }
  if (lane == 0 && scored) { atomicAdd(&counters[1], scored); atomicAdd(&counters[2], cells); }
}

// The clustering path's probe for barcodes of at most 32 nt (probe.cuh): composition filter on a 4-byte sketch, lanes
// that each carry their own pair through the recurrence, Ukkonen cut-off at checkpoints.
__global__ void __launch_bounds__(PROBE_WARPS * 32) probe_kernel_lanes(ProbeArgs a) {
  __shared__ ProbeQueue queue[PROBE_WARPS];
  probe_warp<false>(a, threadIdx.x & 31, &queue[threadIdx.x >> 5], nullptr);
}
// the same with long bins staged through shared memory by TMA bulk copies (PCB_OPT_PROBE_STAGING; probe.cuh)
__global__ void __launch_bounds__(PROBE_WARPS * 32) probe_kernel_staged(ProbeArgs a) {
  __shared__ ProbeQueue queue[PROBE_WARPS];
  __shared__ ProbeStage stage[PROBE_WARPS];
  probe_warp<true>(a, threadIdx.x & 31, &queue[threadIdx.x >> 5], &stage[threadIdx.x >> 5]);
}

// the queries shorter than `below` in [q_begin, q_end), in any order (they probe the second index)
__global__ void short_queries_kernel(const int32_t *__restrict__ len, int32_t q_begin, int32_t q_end, int32_t below, int32_t *__restrict__ list,
                                     int32_t *__restrict__ n) {
  const int32_t i = q_begin + (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < q_end && len[i] < below) list[atomicAdd(n, 1)] = i;
}

__global__ void sketch_kernel(const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int32_t U, uint32_t *__restrict__ sketch) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < U) sketch[i] = barcode_sketch((uint32_t)planes[i].x, (uint32_t)planes[i].y, len[i]);
}

__global__ void len_hist_kernel(const int32_t *__restrict__ len, int32_t U, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[66];
  if (threadIdx.x < 66) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < U; i += gridDim.x * blockDim.x)
    if (len[i] <= 64) atomicAdd(&sh[len[i]], 1u);   // (opaque barcodes carry the marker length and are not counted)
  __syncthreads();
  if (threadIdx.x < 66 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void flag_unique_kernel(const uint64_t *__restrict__ key, size_t n, uint32_t *__restrict__ flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

__global__ void compact_edges_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ d, const uint32_t *__restrict__ rank,
                                     size_t n, int key_shift, int32_t *__restrict__ ei, int32_t *__restrict__ ej,
                                     int32_t *__restrict__ ed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || key[i] != key[i - 1]) {
    uint32_t at = rank[i];
    ei[at] = (int32_t)(key[i] >> key_shift);
    ej[at] = (int32_t)(key[i] & ((1ull << key_shift) - 1ull));
    ed[at] = (int32_t)d[i];
  }
}

static void free_edges(pcb_ctx *ctx) {
  pcb_edges &e = ctx->edges;
  if (e.ei) cudaFreeAsync(e.ei, ctx->stream);
  if (e.ej) cudaFreeAsync(e.ej, ctx->stream);
  if (e.ed) cudaFreeAsync(e.ed, ctx->stream);
  e = pcb_edges();
}

// best match of every query among the (query, target, d) rows sorted by (query, target)
__global__ void match_reduce_kernel(const int32_t *__restrict__ eq, const int32_t *__restrict__ et, const int32_t *__restrict__ ed,
                                    int64_t n, int32_t Q, int32_t *__restrict__ best_d, int32_t *__restrict__ n_best,
                                    int32_t *__restrict__ first_hit) {
  int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (eq[mid] < q) lo = mid + 1; else hi = mid;
  }
  int32_t best = -1, cnt = 0, first = -1;
  for (int64_t x = lo; x < n && eq[x] == q; x++) {
    int32_t d = ed[x];
    if (best < 0 || d < best) { best = d; cnt = 1; first = et[x]; }
    else if (d == best) cnt++;
  }
  best_d[q] = best; n_best[q] = cnt; first_hit[q] = first;
}

static void pair_search(pcb_ctx *ctx, const PackedBarcodes &bc, const PackedBarcodes *queries, int32_t k, int32_t q_begin, int32_t q_end,
                        int32_t t_begin, int32_t t_end);

void edit_pairs_dev(pcb_ctx *ctx, const PackedBarcodes &bc, int32_t k, int32_t q_begin, int32_t q_end, int32_t t_begin, int32_t t_end) {
  pair_search(ctx, bc, nullptr, k, q_begin, q_end, t_begin, t_end < 0 ? bc.n : t_end);
}

void match_barcodes_dev(pcb_ctx *ctx, const PackedBarcodes &library, const PackedBarcodes &queries, int32_t k, int32_t *best_d,
                        int32_t *n_best, int32_t *first_hit) {
  pair_search(ctx, library, &queries, k, 0, queries.n, 0, library.n);
  const pcb_edges &e = ctx->edges;
  if (queries.n)
    PCB_LAUNCH(ctx, match_reduce_kernel, grid_for((size_t)queries.n, 256), 256, 0, e.ei, e.ej, e.ed, e.n, queries.n, best_d, n_best, first_hit);
}

// Pairs (query, target) within k.  queries == nullptr: the self-join of the clustering path (every unordered pair
// once, through its query side); otherwise the rows are (query index, target index) as they stand.
// [t_begin, t_end): the targets that are indexed -- a pair is found by the call whose target range holds its target side.
static void pair_search(pcb_ctx *ctx, const PackedBarcodes &bc, const PackedBarcodes *queries, int32_t k, int32_t q_begin, int32_t q_end,
                        int32_t t_begin, int32_t t_end) {
  free_edges(ctx);
  const bool self = queries == nullptr;
  const PackedBarcodes &qs = self ? bc : *queries;
  const int32_t U = bc.n;
  ctx->stats[PCB_STAT_PAIRS_SCORED] = ctx->stats[PCB_STAT_CELLS] = ctx->stats[PCB_STAT_EDGES] = ctx->stats[PCB_STAT_CANDIDATES] = 0;
  ctx->stats[PCB_STAT_PROBE_NS] = 0;
  if (k < 0 || k > 8) throw Error(PCB_E_INPUT, "k must be in 0..8");
  if (q_begin < 0 || q_end > qs.n || q_begin > q_end) throw Error(PCB_E_INPUT, "query range outside [0, U)");
  if (t_begin < 0 || t_end > bc.n || t_begin > t_end) throw Error(PCB_E_INPUT, "target range outside [0, U)");
  const int32_t T = t_end - t_begin;
  pcb_edges &res = ctx->edges;
  res.valid = true;
  if (U < (self ? 2 : 1) || q_begin == q_end || T == 0) return;
  if (bc.max_len == 0 || qs.max_len == 0) return;           // nothing but opaque barcodes

  // ---- seed lengths.  A query of length m splits into k+1 pieces of at least floor(m/(k+1)) bases, so it can
  // probe an index of q-grams with q up to that.  Most barcodes share one length; the few shorter ones (deletions)
  // would force the short q on everybody, so there are up to two indexes over all targets: q_hi for the queries
  // long enough for it, q_lo for the rest.  Either way every pair is found through its query side.
  uint32_t h_hist[66];
  if (qs.has_hist) {                                          // counted while packing: no round trip
    for (int l = 0; l < 66; l++) h_hist[l] = qs.len_hist[l];
  } else {
    DevBuf<uint32_t> d_hist(ctx, 66);
    PCB_CUDA(cudaMemsetAsync(d_hist.p, 0, 66 * sizeof(uint32_t), ctx->stream));
    PCB_LAUNCH(ctx, len_hist_kernel, (unsigned)ctx->sm_count, 256, 0, qs.len.p, qs.n, d_hist.p);
    PCB_CUDA(cudaMemcpyAsync(h_hist, d_hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, ctx->stream));
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int32_t mode_len = qs.min_len;
  for (int32_t l = 1; l <= 64; l++) if (h_hist[l] > h_hist[mode_len]) mode_len = l;
  auto make_geom = [&](int32_t q) {
    SeedGeom g;
    g.k = k;
    g.q = q > 12 ? 12 : q;
    while (g.q > 0 && ((uint64_t)(bc.max_len - g.q + 1) << (2 * g.q)) > (1ull << ctx->seed_bins_log2)) g.q--;
    g.npos = g.q == 0 ? 1 : bc.max_len - g.q + 1;
    g.nbins = g.q == 0 ? 1u : (uint32_t)g.npos << (2 * g.q);
    g.pos_mask = ~0ull;
    g.cbits = 0; g.rbits = 0; g.short_below = 0; g.only_below = 0;
    return g;
  };
  // the positions the probes of query lengths [lo, hi] can ask for: start of piece t (t * m / (k+1)) plus a shift in [-k, k]
  auto probe_positions = [&](const SeedGeom &g, int32_t lo, int32_t hi) -> uint64_t {
    if (g.q == 0) return ~0ull;
    uint64_t mask = 0;
    if (lo < qs.min_len) lo = qs.min_len;
    if (hi > qs.max_len) hi = qs.max_len;
    for (int32_t m = lo; m <= hi; m++)
      for (int32_t t = 0; t <= k; t++)
        for (int32_t sft = -k; sft <= k; sft++) {
          int32_t pp = (int32_t)((int64_t)t * m / (k + 1)) + sft;
          if (pp >= 0 && pp < g.npos) mask |= 1ull << pp;
        }
    return mask;
  };
  struct SeedIndex { SeedGeom g; DevBuf<uint32_t> bin_off; DevBuf<int32_t> ent; DevBuf<SeedEnt> ent8; int32_t len_lo, len_hi; };
  const bool lanes = self && !(bc.max_len > 32);          // the clustering path on one-word barcodes: probe.cuh
  SeedIndex idx[2];
  int n_idx = 1;
  idx[0].g = make_geom(mode_len / (k + 1));
  idx[0].len_lo = idx[0].g.q * (k + 1); idx[0].len_hi = 64;
  if (qs.min_len < idx[0].len_lo) {
    idx[1].g = make_geom(qs.min_len / (k + 1));
    idx[1].len_lo = 0; idx[1].len_hi = idx[0].len_lo - 1;
    if (idx[1].g.q == idx[0].g.q) idx[0].len_lo = 0; else n_idx = 2;
  } else {
    idx[0].len_lo = 0;
  }
  if (n_idx == 2 && self) {           // the self-join: short x long pairs belong to the long side, the second index holds the short barcodes only
    idx[0].g.short_below = idx[0].len_lo;
    idx[1].g.only_below = idx[0].len_lo;
  }
  ctx->stats[PCB_STAT_SEED_LEN] = idx[0].g.q;
  for (int t = 0; t < n_idx; t++) {
    idx[t].g.pos_mask = probe_positions(idx[t].g, idx[t].len_lo, idx[t].len_hi);
    // long bins (about 8 entries or more, counted over ALL barcodes so that every shard of a multi-GPU search decides
    // alike): 8 classes per bin, so that a probe skips most of what it does not own (probe.cuh: owns_pair).  The first
    // index decides the rule for the whole search -- every pair must have ONE owner whichever index its sides probe.
    const bool long_bins = idx[t].g.q > 0 && ((uint64_t)idx[t].g.nbins << 3) <= (1ull << 25) &&
        (uint64_t)U * (uint64_t)__builtin_popcountll(idx[t].g.pos_mask & (idx[t].g.npos >= 64 ? ~0ull : (1ull << idx[t].g.npos) - 1ull)) >= 8ull * idx[t].g.nbins;
    if (t == 0) idx[0].g.rbits = (lanes && ctx->probe_classes && long_bins) ? 3 : 0;
    idx[t].g.rbits = idx[0].g.rbits;
    idx[t].g.cbits = (long_bins && !idx[t].g.only_below) ? idx[t].g.rbits : 0;
    const SeedGeom &g = idx[t].g;
    const size_t n_sub = (size_t)g.nbins << g.cbits;
    DevBuf<uint32_t> cursor(ctx, n_sub + 1);
    idx[t].bin_off.alloc(ctx, n_sub + 1);
    if (lanes) idx[t].ent8.alloc(ctx, (size_t)T * g.npos + 2); else idx[t].ent.alloc(ctx, (size_t)T * g.npos);
    PCB_CUDA(cudaMemsetAsync(idx[t].bin_off.p, 0, (n_sub + 1) * sizeof(uint32_t), ctx->stream));
    size_t n_slots = (size_t)T * g.npos;
    PCB_LAUNCH(ctx, seed_count_kernel, grid_for(n_slots, 256), 256, 0, bc.planes.p, bc.len.p, t_begin, T, g, idx[t].bin_off.p);
    exclusive_scan_u32(ctx, idx[t].bin_off.p, idx[t].bin_off.p, n_sub + 1);
    PCB_CUDA(cudaMemcpyAsync(cursor.p, idx[t].bin_off.p, (n_sub + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (lanes) PCB_LAUNCH(ctx, seed_fill8_kernel, grid_for(n_slots, 256), 256, 0, bc.planes.p, bc.len.p, t_begin, T, g, cursor.p, idx[t].ent8.p);
    else PCB_LAUNCH(ctx, seed_fill_kernel, grid_for(n_slots, 256), 256, 0, bc.planes.p, bc.len.p, t_begin, T, g, cursor.p, idx[t].ent.p);
  }

  stage_end(ctx, PCB_STAT_T_INDEX_NS);
  // ---- probe (re-run with a larger hit buffer if the first guess overflows)
  const int key_shift = bits_for((uint64_t)(U > 1 ? U - 1 : 1));
  unsigned long long cap = 8ull * (unsigned long long)(q_end - q_begin) + (1ull << 16);
  DevBuf<unsigned long long> counters(ctx, 6);
  DevBuf<uint32_t> sketch;
  if (lanes) {
    sketch.alloc(ctx, (size_t)U);
    PCB_LAUNCH(ctx, sketch_kernel, grid_for(U, 256), 256, 0, bc.planes.p, bc.len.p, U, sketch.p);
  }
  DevBuf<uint64_t> hk0, hk1;
  DevBuf<uint32_t> hv0, hv1;
  unsigned long long h_counters[6];
  const bool wide = bc.max_len > 32 || qs.max_len > 32;
  auto launch = [&](auto kernel, unsigned n_cta, const SeedIndex &ix) {
    PCB_LAUNCH(ctx, kernel, n_cta, PROBE_WARPS * 32, 0, qs.planes.p, qs.len.p, (const uint64_t *)(self ? nullptr : qs.other.p), bc.planes.p, bc.len.p, q_begin, q_end, ix.g,
               ix.bin_off.p, ix.ent.p, hk0.p, hv0.p, cap, key_shift, counters.p, ix.len_lo, ix.len_hi);
  };
  // the short barcodes as a list: the second index is probed by them alone (a probe loop over every barcode to find the few
  // that qualify cost 70 us of the 270 us probe stage at 10^6 reads)
  DevBuf<int32_t> short_list, short_n;
  if (lanes && n_idx == 2 && idx[1].g.only_below) {
    short_list.alloc(ctx, (size_t)(q_end - q_begin));
    short_n.alloc(ctx, 1);
    PCB_CUDA(cudaMemsetAsync(short_n.p, 0, sizeof(int32_t), ctx->stream));
    PCB_LAUNCH(ctx, short_queries_kernel, grid_for((size_t)(q_end - q_begin), 256), 256, 0, bc.len.p, q_begin, q_end, idx[1].g.only_below,
               short_list.p, short_n.p);
  }
  for (int attempt = 0;; attempt++) {
    hk0.alloc(ctx, cap); hv0.alloc(ctx, cap);
    PCB_CUDA(cudaMemsetAsync(counters.p, 0, 6 * sizeof(unsigned long long), ctx->stream));
    PCB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int t = 0; t < n_idx; t++) {
      const SeedGeom &g = idx[t].g;
      const int per_query = g.q == 0 ? 1 : (k + 1) * (2 * k + 1);
      size_t n_probe = (size_t)(q_end - q_begin) * per_query;
      // persistent grid: a multiple of the SM count, blocks of probes handed out by counters[3]
      unsigned n_cta = (unsigned)ctx->sm_count * 8u;
      unsigned need = grid_for((n_probe + 31) / 32, PROBE_WARPS);
      if (need < n_cta) n_cta = need;
      if (t > 0) PCB_CUDA(cudaMemsetAsync(counters.p + 3, 0, sizeof(unsigned long long), ctx->stream));
      if (lanes) {
        ProbeArgs a;
        a.planes = (const Planes16 *)bc.planes.p; a.len = bc.len.p; a.sketch = sketch.p; a.q_begin = q_begin; a.q_end = q_end; a.g = g;
        a.bin_off = idx[t].bin_off.p; a.ent = idx[t].ent8.p; a.hit_key = hk0.p; a.hit_d = hv0.p; a.cap = cap; a.key_shift = key_shift;
        a.counters = counters.p; a.len_lo = idx[t].len_lo; a.len_hi = idx[t].len_hi;
        a.qlist = nullptr; a.qlist_n = nullptr;
        if (idx[t].g.only_below && short_list.p) {            // only the short barcodes probe this index
          a.qlist = short_list.p; a.qlist_n = short_n.p;
          if (n_cta > (unsigned)ctx->sm_count) n_cta = (unsigned)ctx->sm_count;
        }
        if (ctx->probe_staging) PCB_LAUNCH(ctx, probe_kernel_staged, n_cta, PROBE_WARPS * 32, 0, a);
        else PCB_LAUNCH(ctx, probe_kernel_lanes, n_cta, PROBE_WARPS * 32, 0, a);
      }
      else if (wide && self) launch(probe_kernel<true, true>, n_cta, idx[t]);
      else if (wide) launch(probe_kernel<true, false>, n_cta, idx[t]);
      else if (self) launch(probe_kernel<false, true>, n_cta, idx[t]);
      else launch(probe_kernel<false, false>, n_cta, idx[t]);
    }
    PCB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PCB_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, ctx->stream));
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    PCB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats[PCB_STAT_PROBE_NS] = (int64_t)(ms * 1e6);
    if (h_counters[0] <= cap) break;
    if (attempt >= 2) throw Error(PCB_E_CUDA, "hit buffer overflowed twice");
    cap = h_counters[0];
  }
  stage_end(ctx, -1);      // the probe interval itself is PCB_STAT_PROBE_NS (ev0/ev1 above)
  ctx->stats[PCB_STAT_PAIRS_SCORED] = (int64_t)h_counters[1];
  ctx->stats[PCB_STAT_CELLS] = (int64_t)h_counters[2];
  ctx->stats[PCB_STAT_CANDIDATES] = lanes ? (int64_t)h_counters[5] : (int64_t)h_counters[1];
  size_t n_hits = (size_t)h_counters[0];
  if (n_hits == 0) return;

  // ---- canonical order, duplicates removed
  hk1.alloc(ctx, n_hits); hv1.alloc(ctx, n_hits);
  int where = radix_sort_pairs(ctx, hk0.p, hv0.p, hk1.p, hv1.p, n_hits, key_shift + bits_for((uint64_t)(qs.n > 1 ? qs.n - 1 : 1)));
  uint64_t *sk = where ? hk1.p : hk0.p;
  uint32_t *sv = where ? hv1.p : hv0.p;
  DevBuf<uint32_t> rank(ctx, n_hits + 1);
  PCB_LAUNCH(ctx, flag_unique_kernel, grid_for(n_hits, 256), 256, 0, sk, n_hits, rank.p);
  PCB_CUDA(cudaMemsetAsync(rank.p + n_hits, 0, sizeof(uint32_t), ctx->stream));
  exclusive_scan_u32(ctx, rank.p, rank.p, n_hits + 1);                  // one past the end: the total
  size_t n_edges = (size_t)read_scalar(ctx, rank.p + n_hits);
  PCB_CUDA(cudaMallocAsync((void **)&res.ei, n_edges * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMallocAsync((void **)&res.ej, n_edges * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMallocAsync((void **)&res.ed, n_edges * sizeof(int32_t), ctx->stream));
  PCB_LAUNCH(ctx, compact_edges_kernel, grid_for(n_hits, 256), 256, 0, sk, sv, rank.p, n_hits, key_shift, res.ei, res.ej, res.ed);
  res.n = (int64_t)n_edges;
  ctx->stats[PCB_STAT_EDGES] = res.n;
  stage_end(ctx, PCB_STAT_T_HITS_NS);
}

}  // namespace pcb
