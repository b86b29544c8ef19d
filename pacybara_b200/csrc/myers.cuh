// myers.cuh -- bit-parallel unit-cost Levenshtein distance (Myers 1999, Hyyro's global-distance
// formulation) on barcodes packed as two bit planes.
//
// Replaces, for one pair, what the reference gets out of a bowtie2 alignment plus
// src/pacybara_calcEdits.R:144-162; the distance is DEFINED as Levenshtein.
//
// Packing: base codes A=0 C=1 G=2 T=3; plane L holds bit 0 of every base, plane H bit 1,
// base i at bit i.  A barcode of <= 32 nt is one 64-bit word (H << 32 | L), <= 64 nt two.
//
// Recurrence per text column j (pattern = one barcode in registers, all m rows at once):
//     Eq = rows whose base equals text[j]
//     Xv = Eq | Mv
//     Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq
//     Ph = Mv | ~(Xh | Pv);  Mh = Pv & Xh
//     Ph = (Ph << 1) | 1;    Mh <<= 1            (row 0 of the DP grows by 1 per column)
//     Pv = Mh | ~(Xv | Ph);  Mv = Ph & Xv
// and D[m][n] = n + popc(Pv & rows) - popc(Mv & rows) after the last column, so no per-column
// score bookkeeping is needed.  Bits above row m-1 never influence the rows below them
// (carries and shifts only travel upwards), so no masking is needed inside the loop.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PCB_HD __host__ __device__ __forceinline__
#else
#define PCB_HD inline
#endif

namespace pcb {

PCB_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
PCB_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

// ASCII base -> 2-bit code, 4 for anything that is not ACGT
PCB_HD uint32_t base_code(uint8_t c) {
  return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

// Distance between a pattern (pL,pH; m rows) and a text (tL,tH; n columns), both <= 32 nt.
PCB_HD int myers32(uint32_t pL, uint32_t pH, int m, uint32_t tL, uint32_t tH, int n) {
  uint32_t Pv = 0xFFFFFFFFu, Mv = 0u;
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    uint32_t cl = 0u - ((tL >> j) & 1u), ch = 0u - ((tH >> j) & 1u);
    uint32_t Eq = ~((pL ^ cl) | (pH ^ ch));
    uint32_t Xv = Eq | Mv;
    uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint32_t Ph = Mv | ~(Xh | Pv);
    uint32_t Mh = Pv & Xh;
    Ph = (Ph << 1) | 1u;
    Mh <<= 1;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
  }
  uint32_t rows = m >= 32 ? 0xFFFFFFFFu : ((1u << m) - 1u);
  return n + popc32(Pv & rows) - popc32(Mv & rows);
}

// Same for barcodes of up to 64 nt.
// `other`: pattern rows holding a character that equals no base (N, ...): they never match (f3 queries)
PCB_HD int myers64(uint64_t pL, uint64_t pH, int m, uint64_t tL, uint64_t tH, int n, uint64_t other = 0ull) {
  uint64_t Pv = ~0ull, Mv = 0ull;
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    uint64_t cl = 0ull - ((tL >> j) & 1ull), ch = 0ull - ((tH >> j) & 1ull);
    uint64_t Eq = ~((pL ^ cl) | (pH ^ ch)) & ~other;
    uint64_t Xv = Eq | Mv;
    uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    Ph = (Ph << 1) | 1ull;
    Mh <<= 1;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
  }
  uint64_t rows = m >= 64 ? ~0ull : ((1ull << m) - 1ull);
  return n + popc64(Pv & rows) - popc64(Mv & rows);
}

// q-gram starting at base `pos` as a 2q-bit key (low q bits: plane L, high q bits: plane H)
PCB_HD uint32_t qgram_key(uint64_t L, uint64_t H, int pos, int q) {
  uint64_t mask = (1ull << q) - 1ull;
  return (uint32_t)(((L >> pos) & mask) | (((H >> pos) & mask) << q));
}

// Which side of an unordered pair {a, b} is its "query" (pcb_edit_pairs): balanced so that every
// barcode is the query of about half of its candidate pairs whatever its index.
PCB_HD bool is_query_side(int32_t a, int32_t b) {
  return ((a ^ b) & 1) ? (a > b) : (a < b);
}

}  // namespace pcb
