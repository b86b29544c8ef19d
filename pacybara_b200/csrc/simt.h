// simt.h -- the warp-collective vocabulary of the warp-synchronous kernels (merge_small.cuh).
//
// Under nvcc these are the sm_100a intrinsics, one instruction each.  Under a plain host compiler
// (tests/host_emul only -- a unit-test harness, never a product path) they are declared here and
// implemented by tests/host_emul/simt_emul.cpp, which runs the 32 lanes of a warp as coroutines that
// meet at every collective.  That way the very source the GPU executes is exercised against the oracle
// in the GPU-less build container, with the lane interleaving randomised so that a missing
// __syncwarp shows up as a failing test rather than as a rare flake on the device.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SIMT_FN __device__ __forceinline__
#define SIMT_NOINLINE __device__ __noinline__

namespace simt {
SIMT_FN uint32_t ballot(uint32_t mask, bool p) { return __ballot_sync(mask, p); }
SIMT_FN int32_t shfl(uint32_t mask, int32_t v, int src) { return __shfl_sync(mask, v, src); }
SIMT_FN uint32_t shfl(uint32_t mask, uint32_t v, int src) { return __shfl_sync(mask, v, src); }
SIMT_FN uint32_t match_any(uint32_t mask, int32_t v) { return __match_any_sync(mask, v); }
SIMT_FN uint32_t reduce_or(uint32_t mask, uint32_t v) { return __reduce_or_sync(mask, v); }
SIMT_FN int32_t reduce_max(uint32_t mask, int32_t v) { return __reduce_max_sync(mask, v); }
SIMT_FN int32_t reduce_add(uint32_t mask, int32_t v) { return __reduce_add_sync(mask, v); }
SIMT_FN void sync(uint32_t mask) { __syncwarp(mask); }
SIMT_FN int32_t atomic_cas(int32_t *p, int32_t cmp, int32_t v) { return atomicCAS(p, cmp, v); }
SIMT_FN uint32_t atomic_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
SIMT_FN int32_t atomic_add(int32_t *p, int32_t v) { return atomicAdd(p, v); }
SIMT_FN int64_t atomic_add64(int64_t *p, int64_t v) { return (int64_t)atomicAdd((unsigned long long *)p, (unsigned long long)v); }
SIMT_FN int popc(uint32_t x) { return __popc(x); }
SIMT_FN int ffs(uint32_t x) { return __ffs((int)x); }      // 1-based, 0 when x == 0
SIMT_FN float sqrt_f(float x) { return sqrtf(x); }
SIMT_FN uint32_t brev(uint32_t x) { return __brev(x); }
// read once: do not keep it in the caches (8-byte items)
template <typename T>
SIMT_FN T ld_stream(const T *p) {
  static_assert(sizeof(T) == 8, "8-byte items");
  const longlong1 v = {__ldcs((const long long *)p)};
  T out;
  memcpy(&out, &v, 8);
  return out;
}
SIMT_FN unsigned long long atomic_add_u64(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
SIMT_FN unsigned long long atomic_min_u64(unsigned long long *p, unsigned long long v) { return atomicMin(p, v); }
SIMT_FN int32_t atomic_max(int32_t *p, int32_t v) { return atomicMax(p, v); }
// sum over the four bytes of max(0, a_byte - b_byte)
SIMT_FN uint32_t surplus4(uint32_t a, uint32_t b) { return __vsadu4(__vsubus4(a, b), 0u); }
// Bulk copy global -> shared through the TMA unit, completion counted in bytes on an mbarrier (one thread issues, any
// thread waits).  dst / src / bytes must be multiples of 16.
SIMT_FN uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SIMT_FN void bar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
SIMT_FN void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
               "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
SIMT_FN void bar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
}
}  // namespace simt

#else
#include <math.h>
#define SIMT_FN inline
#define SIMT_NOINLINE inline

namespace simt {
// implemented by tests/host_emul/simt_emul.cpp; `lane` is implicit (the coroutine that is running)
uint32_t ballot(uint32_t mask, bool p);
int32_t shfl(uint32_t mask, int32_t v, int src);
uint32_t shfl(uint32_t mask, uint32_t v, int src);
uint32_t match_any(uint32_t mask, int32_t v);
uint32_t reduce_or(uint32_t mask, uint32_t v);
int32_t reduce_max(uint32_t mask, int32_t v);
int32_t reduce_add(uint32_t mask, int32_t v);
void sync(uint32_t mask);
// lanes run one at a time between collectives, so plain read-modify-writes are atomic
inline int32_t atomic_cas(int32_t *p, int32_t cmp, int32_t v) { int32_t o = *p; if (o == cmp) *p = v; return o; }
inline uint32_t atomic_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
inline int32_t atomic_add(int32_t *p, int32_t v) { int32_t o = *p; *p = o + v; return o; }
inline int64_t atomic_add64(int64_t *p, int64_t v) { int64_t o = *p; *p = o + v; return o; }
inline int popc(uint32_t x) { return __builtin_popcount(x); }
inline int ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline float sqrt_f(float x) { return sqrtf(x); }
template <typename T> inline T ld_stream(const T *p) { return *p; }
inline uint32_t brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned long long atomic_add_u64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomic_min_u64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
inline int32_t atomic_max(int32_t *p, int32_t v) { int32_t o = *p; if (v > o) *p = v; return o; }
inline uint32_t surplus4(uint32_t a, uint32_t b) {
  uint32_t s = 0;
  for (int i = 0; i < 4; i++) { uint32_t x = (a >> (8 * i)) & 0xFFu, y = (b >> (8 * i)) & 0xFFu; if (x > y) s += x - y; }
  return s;
}
// the emulated bulk copy completes at once
inline void bar_init(uint64_t *bar, int) { *bar = 0; }
inline void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
inline void bar_wait(uint64_t *, uint32_t) {}
}  // namespace simt
#endif
