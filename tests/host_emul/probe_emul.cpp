// probe_emul.cpp -- UNIT-TEST HARNESS, not a product path: runs pacybara_b200/csrc/probe.cuh (the probe the GPU executes
// for <= 32-nt barcodes) on emulated warps over a seed index built with plain loops, so that its candidate filters,
// queue handling and early exit can be checked against the exhaustive DP without a GPU.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "probe.cuh"
#include "simt_emul.h"

using namespace pcb;

// planes: [U][2] uint64 (L, H); returns the number of distinct pairs found (ei < ej, sorted by (ei, ej)); -1: capacity
extern "C" int64_t emul_probe(const uint64_t *planes, const int32_t *len, int32_t U, int32_t k, int32_t q_begin, int32_t q_end,
                              int32_t n_warps, uint32_t seed, int32_t *ei, int32_t *ej, int32_t *ed, int64_t cap, uint64_t *stats) {
  int32_t min_len = 64, max_len = 0;
  for (int32_t i = 0; i < U; i++) { min_len = std::min(min_len, len[i]); max_len = std::max(max_len, len[i]); }
  if (U == 0) return 0;
  SeedGeom g;
  g.k = k;
  g.q = std::min(12, min_len / (k + 1));
  g.npos = g.q == 0 ? 1 : max_len - g.q + 1;
  g.nbins = g.q == 0 ? 1u : (uint32_t)g.npos << (2 * g.q);
  g.pos_mask = ~0ull;
  g.short_below = 0; g.only_below = 0;
  g.rbits = (seed & 0x40000000u) ? 3 : 0;              // bit 30 of the seed: the class rule, bins split into 8 classes (owns_pair)
  g.cbits = g.q > 0 && !(seed & 0x20000000u) ? g.rbits : 0;      // (bit 29: whole bins under the class rule, as a second index has them)
  const size_t n_sub = (size_t)g.nbins << g.cbits;
  std::vector<uint32_t> bin_off(n_sub + 1, 0);
  std::vector<std::pair<uint32_t, int32_t>> items;
  for (int32_t j = 0; j < U; j++)
    for (int32_t p = 0; p < g.npos; p++) {
      if (g.q > 0 && p + g.q > len[j]) continue;
      items.push_back({seed_slot(g, seed_bin(g, planes[2 * j], planes[2 * j + 1], p, p), j, len[j]), j});
    }
  std::stable_sort(items.begin(), items.end(), [](auto &x, auto &y) { return x.first < y.first; });
  std::vector<SeedEnt> ent(items.size() + 2);
  for (size_t i = 0; i < items.size(); i++) {
    const int32_t j = items[i].second;
    bin_off[items[i].first + 1]++;
    ent[i].j = j; ent[i].sk = barcode_sketch((uint32_t)planes[2 * j], (uint32_t)planes[2 * j + 1], len[j]);
  }
  for (size_t b = 0; b < n_sub; b++) bin_off[b + 1] += bin_off[b];
  unsigned long long hit_cap = 64ull * (unsigned long long)U * (unsigned long long)U + 4096;      // a pair can be found through every piece and shift
  if (hit_cap > (1ull << 24)) hit_cap = 1ull << 24;
  std::vector<uint64_t> hk(hit_cap);
  std::vector<uint32_t> hd(hit_cap);
  unsigned long long counters[6] = {0, 0, 0, 0, 0, 0};
  std::vector<uint32_t> sketch(U);
  for (int32_t j = 0; j < U; j++) sketch[j] = barcode_sketch((uint32_t)planes[2 * j], (uint32_t)planes[2 * j + 1], len[j]);
  ProbeArgs a;
  a.planes = (const Planes16 *)planes; a.len = len; a.sketch = sketch.data(); a.q_begin = q_begin; a.q_end = q_end; a.g = g;
  a.bin_off = bin_off.data(); a.ent = ent.data(); a.hit_key = hk.data(); a.hit_d = hd.data(); a.cap = hit_cap;
  a.key_shift = 24; a.counters = counters; a.len_lo = 0; a.len_hi = 64; a.qlist = nullptr; a.qlist_n = nullptr;
  static ProbeQueue qu;
  static ProbeStage stg;
  const bool staged = (seed & 0x80000000u) != 0;        // top bit of the seed: long bins through the staging buffers
  for (int32_t w = 0; w < n_warps; w++) {               // warps one after the other: they only share the block counter
    if (staged) simt_emul::run_warp([&](int lane) { probe_warp<true>(a, lane, &qu, &stg); }, seed + (uint32_t)w);
    else simt_emul::run_warp([&](int lane) { probe_warp<false>(a, lane, &qu, nullptr); }, seed + (uint32_t)w);
  }
  if (counters[0] > hit_cap) return -1;
  std::vector<std::pair<uint64_t, uint32_t>> hits(counters[0]);
  for (size_t i = 0; i < hits.size(); i++) hits[i] = {hk[i], hd[i]};
  std::sort(hits.begin(), hits.end());
  int64_t n = 0;
  for (size_t i = 0; i < hits.size(); i++) {
    if (i && hits[i].first == hits[i - 1].first) { if (hits[i].second != hits[i - 1].second) return -2; continue; }
    if (n >= cap) return -1;
    ei[n] = (int32_t)(hits[i].first >> 24); ej[n] = (int32_t)(hits[i].first & 0xFFFFFF); ed[n] = (int32_t)hits[i].second;
    n++;
  }
  for (int i = 0; i < 6; i++) stats[i] = counters[i];
  return n;
}
