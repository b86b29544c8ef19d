// myers_emul.cpp -- UNIT-TEST HARNESS, not a product path: compiles csrc/myers.cuh for the host
// so the bit-parallel recurrence can be checked against the DP oracle without a GPU.
#include <cstdint>
#include "myers.cuh"
using namespace pcb;

static void pack(const uint8_t *s, int n, uint64_t *L, uint64_t *H) {
  *L = 0; *H = 0;
  for (int i = 0; i < n; i++) { uint64_t c = base_code(s[i]); *L |= (c & 1) << i; *H |= ((c >> 1) & 1) << i; }
}

extern "C" void emul_myers(const uint8_t *seq, const int32_t *len, int32_t stride, int64_t n_pairs,
                           const int32_t *pi, const int32_t *pj, int32_t use64, int32_t *out) {
  for (int64_t p = 0; p < n_pairs; p++) {
    uint64_t aL, aH, bL, bH;
    int m = len[pi[p]], n = len[pj[p]];
    pack(seq + (int64_t)pi[p] * stride, m, &aL, &aH);
    pack(seq + (int64_t)pj[p] * stride, n, &bL, &bH);
    out[p] = use64 ? myers64(aL, aH, m, bL, bH, n)
                   : myers32((uint32_t)aL, (uint32_t)aH, m, (uint32_t)bL, (uint32_t)bH, n);
  }
}
