// merge.cu -- cluster merge on the device (rows a4..a13): a few streaming kernels that turn the bucket-pair list
// into per-component work, then the replay of the reference's sequential algorithm, one connected component per
// warp (merge_small.cuh for components of at most 32 reads, merge_warp.cuh / merge_core.h for the rest).
//
// Nothing is renumbered or gathered: the replay kernels read the caller's arrays through two small indirections
// (component -> its buckets, bucket -> its reads) and write the caller's output arrays directly.
//
//   components    lock-free union-find over the pairs a pass visits (1 <= ed <= maxDiff); label = smallest bucket of
//                 the component; a component belongs to shard (label mod shard_count)
//   grouping      buckets sorted by (not mine, label) -- ONE radix sort over the U bucket labels: comp_buckets
//                 (ascending bucket id inside a component, the sort is stable) and comp_ptr
//   adjacency     the pair list itself when the caller's pairs are sorted by (q < r) -- which is how the pair search
//                 delivers them -- plus row pointers; a sorted copy otherwise.  Every pair is held once, in the row of
//                 its smaller bucket (MergeCtx::adj_forward); this is the lookup table of :236,:266
//   visiting      the pairs a pass visits, per component, in the reference's order: pass by pass, rows in the caller's
//                 order.  Sorted pairs: a count + one walk per component over the rows of its buckets (no sort);
//                 otherwise a STABLE sort on (component, ed)
//   tiers         <= 32 reads: component_kernel_small.  33..128 reads: component_kernel_medium -- the generic warp
//                 replay on a LOCAL copy of the component in shared memory (reads, calls, state, scratch), the
//                 handful of such components of a library would otherwise be the tail of the step; it runs on a
//                 second stream beside the small kernel, and takes whatever that kernel declines afterwards.
//                 Bigger: the generic replay on global state (scratch sized per component).  Beyond huge_reads reads (or
//                 huge_buckets buckets): the giant tier, merge_huge.cuh -- the whole device inside one component.
//   call regions  a component's rows put their calls into a region as long as its reads' call lists, taken from a
//                 shared cursor as components come by: the layout of call_mut is not reproducible from run to
//                 run, the rows (row_call_off / row_call_n / the calls themselves) are.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "merge_core.h"
#include "merge_small.cuh"
#include "merge_warp.cuh"
#include "merge_huge.cuh"

namespace pcb {

// ------------------------------------------------------------------------------------------ components
__device__ __forceinline__ int32_t uf_find(int32_t *parent, int32_t x) {
  for (;;) {
    int32_t p = ((volatile int32_t *)parent)[x];
    if (p == x) return x;
    int32_t gp = ((volatile int32_t *)parent)[p];
    if (gp != p) parent[x] = gp;                // path halving: still an ancestor
    x = p;
  }
}

__global__ void uf_init_kernel(int32_t *parent, int32_t U) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < U) parent[b] = b;
}

__global__ void uf_union_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ pr, const int32_t *__restrict__ ped,
                                int64_t n_pairs, int32_t max_diff, int32_t *parent) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || ped[p] < 1 || ped[p] > max_diff) return;
  int32_t a = pq[p], b = pr[p];
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    int32_t hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(&parent[hi], hi, lo) == hi) return;       // roots only ever point to a smaller root
  }
}

// counters: [0] buckets of this shard, [1] components of this shard, [2] medium list, [3] large list, [4] visiting pairs
// of this shard, [5] components the small kernel declined, [6] declined components too big for the medium kernel
// [7] giant components (merge_huge.cuh)
enum { CNT_BUCKETS = 0, CNT_COMPS, CNT_MEDIUM, CNT_LARGE, CNT_VISITS, CNT_DECLINED, CNT_LATE, CNT_HUGE, CNT_N };
// status: [0] error code, [1] detail, [2] links, [3] tests (generic kernels), [4 ..] 32 slots of (links, tests) from the small kernel
constexpr int STATUS_N = 4 + 64;

// key = (not mine, label): this shard's components first, buckets of a component together in ascending order
__global__ void comp_keys_kernel(int32_t *parent, int32_t U, int bbits, int32_t shard_rank, int32_t shard_count,
                                 int32_t *__restrict__ label, uint64_t *__restrict__ key, uint32_t *__restrict__ val,
                                 int32_t *__restrict__ counters) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  const uint32_t l = (uint32_t)uf_find(parent, b);
  label[b] = (int32_t)l;
  const bool mine = (int32_t)(l % (uint32_t)shard_count) == shard_rank;
  key[b] = ((uint64_t)(mine ? 0 : 1) << bbits) | l;
  val[b] = (uint32_t)b;
  if (mine) atomicAdd(&counters[CNT_BUCKETS], 1);
}

__global__ void comp_flag_kernel(const uint64_t *__restrict__ key, int32_t U, uint32_t *__restrict__ flag) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= U) flag[i] = (i < U && (i == 0 || key[i] != key[i - 1])) ? 1u : 0u;
}

// sorted position i: bucket val[i] of component rank[i] (rank = exclusive scan of the head flags; rank[U] = #components)
__global__ void comp_assign_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, const uint32_t *__restrict__ rank,
                                   int32_t U, int bbits, int32_t *__restrict__ comp_buckets, int32_t *__restrict__ comp_ptr,
                                   int32_t *__restrict__ comp_of_b, int32_t *__restrict__ counters) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  const bool head = i == 0 || key[i] != key[i - 1];
  const bool mine = (key[i] >> bbits) == 0;
  const int32_t c = (int32_t)rank[i] + (head ? 1 : 0) - 1;
  const int32_t b = (int32_t)val[i];
  comp_buckets[i] = b;
  comp_of_b[b] = mine ? c : -1;
  if (head) {
    comp_ptr[c] = i;                             // the first bucket of another shard closes this shard's last component
    if (mine) atomicAdd(&counters[CNT_COMPS], 1);
  }
  if (i == U - 1) comp_ptr[rank[U]] = U;
}

// PCB_MERGE_SHARD_LAYOUT: lab_nnz[label] = genotype entries of the whole component, so that the call region of a
// component sits where its label puts it -- which every shard computes identically
__global__ void label_nnz_kernel(const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads,
                                 const int32_t *__restrict__ geno_ptr, const int32_t *__restrict__ label, int32_t U,
                                 uint32_t *__restrict__ lab_nnz) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  uint32_t nnz = 0;
  for (int32_t x = bucket_ptr[b]; x < bucket_ptr[b + 1]; x++) {
    int32_t r = bucket_reads[x];
    nnz += (uint32_t)(geno_ptr[r + 1] - geno_ptr[r]);
  }
  if (nnz) atomicAdd(&lab_nnz[label[b]], nnz);
}
__global__ void comp_call_off_kernel(const int32_t *__restrict__ comp_ptr, const int32_t *__restrict__ comp_buckets,
                                     const int32_t *__restrict__ counters, const uint32_t *__restrict__ lab_off, int64_t *__restrict__ call_off) {
  const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < counters[1]) call_off[c] = (int64_t)lab_off[comp_buckets[comp_ptr[c]]];      // a component's label is its smallest bucket
}

// ------------------------------------------------------------------------------------------ adjacency
// row pointers of pairs sorted by q: ptr[b] = first pair with q >= b
__global__ void row_ptr_kernel(const int32_t *__restrict__ q, int64_t n, int32_t U, int32_t *__restrict__ ptr) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  const int32_t lo = p == 0 ? 0 : q[p - 1] + 1, hi = p == n ? U : q[p];
  for (int32_t b = lo; b <= hi; b++) ptr[b] = (int32_t)p;
}

__global__ void pair_keys_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ pr, int64_t n, int shift,
                                 uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t a = (uint32_t)pq[p], b = (uint32_t)pr[p];
  key[p] = a < b ? ((uint64_t)a << shift) | b : ((uint64_t)b << shift) | a;
  val[p] = (uint32_t)p;
}
__global__ void pair_unpack_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ order, const int32_t *__restrict__ pd,
                                   int64_t n, int shift, int32_t *__restrict__ sq, int32_t *__restrict__ sr, int32_t *__restrict__ sd) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sq[i] = (int32_t)(key[i] >> shift);
  sr[i] = (int32_t)(key[i] & ((1ull << shift) - 1ull));
  sd[i] = pd[order[i]];
}

// ------------------------------------------------------------------------------------------ visiting lists
// cnt[4 * comp + ed] = pairs pass `ed` visits in this shard's component comp
__global__ void visit_count_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ ped, int64_t n, int32_t max_diff,
                                   const int32_t *__restrict__ comp_of_b, uint32_t *__restrict__ cnt, int32_t *__restrict__ counters) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int32_t ed = ped[p];
  if (ed < 1 || ed > max_diff) return;
  const int32_t c = comp_of_b[pq[p]];
  if (c < 0) return;
  atomicAdd(&cnt[4 * (size_t)c + ed], 1u);
  atomicAdd(&counters[CNT_VISITS], 1);
}

// Pairs sorted by (q, r): the rows a component's passes visit, in order, are the visiting pairs found by walking the
// adjacency rows of its buckets in ascending bucket order -- once per pass.  One thread per BUCKET (in component order):
// g[e][i] = rows at distance e + 1 in the adjacency row of the i-th bucket; after an exclusive scan of each g[e] over
// the bucket list, the rows of bucket i in pass e of component c (buckets i0 .. i1) start at
//   sum_e' G[e'][i0]  +  sum_{e' < e} (G[e'][i1] - G[e'][i0])  +  G[e][i] - G[e][i0]
// (everything before the component, the earlier passes of the component, the earlier buckets of this pass) -- however
// many buckets the component has: a giant component costs what its buckets cost.
__global__ void visit_cnt_kernel(const int32_t *__restrict__ comp_buckets, const int32_t *__restrict__ comp_of_b,
                                 const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ ped, int32_t U, int32_t max_diff,
                                 uint32_t *__restrict__ g) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > U) return;
  uint32_t n[3] = {0u, 0u, 0u};
  if (i < U) {
    const int32_t b = comp_buckets[i];
    if (comp_of_b[b] >= 0)
      for (int32_t e = row_ptr[b]; e < row_ptr[b + 1]; e++) {
        const int32_t ed = ped[e];
        if (ed == 1) n[0]++; else if (ed == 2) n[1]++; else if (ed == 3) n[2]++;
      }
  }
  for (int32_t e = 0; e < max_diff; e++) g[(size_t)e * ((size_t)U + 1) + i] = n[e];
}
__global__ void visit_fill_kernel(const int32_t *__restrict__ comp_buckets, const int32_t *__restrict__ comp_of_b,
                                  const int32_t *__restrict__ comp_ptr, const int32_t *__restrict__ counters,
                                  const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ pr, const int32_t *__restrict__ ped,
                                  int32_t U, int32_t max_diff, const uint32_t *__restrict__ g, int32_t *__restrict__ cpair_ptr,
                                  int32_t *__restrict__ cq, int32_t *__restrict__ cr, int32_t *__restrict__ ce) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  const int32_t b = comp_buckets[i], comp = comp_of_b[b];
  if (comp < 0) return;
  const int32_t i0 = comp_ptr[comp], i1 = comp_ptr[comp + 1];
  const size_t ld = (size_t)U + 1;
  uint32_t base = 0;
  for (int32_t e = 0; e < max_diff; e++) base += g[e * ld + i0];
  uint32_t at[3] = {0u, 0u, 0u}, acc = base;
  for (int32_t e = 0; e < max_diff; e++) {
    at[e] = acc + (g[e * ld + i] - g[e * ld + i0]);
    acc += g[e * ld + i1] - g[e * ld + i0];
  }
  if (i == i0) cpair_ptr[comp] = (int32_t)base;
  if (i == i1 - 1 && comp == counters[CNT_COMPS] - 1) cpair_ptr[comp + 1] = (int32_t)acc;      // end of this shard's last component
  for (int32_t e = row_ptr[b]; e < row_ptr[b + 1]; e++) {
    const int32_t ed = ped[e];
    uint32_t k;
    if (ed == 1) k = at[0]++; else if (ed == 2 && max_diff >= 2) k = at[1]++; else if (ed == 3 && max_diff >= 3) k = at[2]++; else continue;
    cq[k] = b; cr[k] = pr[e]; ce[k] = ed;
  }
}

// pairs in arbitrary order (pcb_merge on a caller's pair list): stable sort on (component, ed)
__global__ void cpair_keys_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ ped, int64_t n_pairs,
                                  const int32_t *__restrict__ comp_of_b, int32_t U, int32_t max_diff,
                                  uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int32_t ed = ped[p];
  const int32_t c = (ed >= 1 && ed <= max_diff) ? comp_of_b[pq[p]] : -1;
  key[p] = c >= 0 ? (((uint64_t)(uint32_t)c << 2) | (uint64_t)ed) : (((uint64_t)(uint32_t)U << 2) | 3ull);      // others sort behind
  val[p] = (uint32_t)p;
}
__global__ void cpair_gather_kernel(const uint32_t *__restrict__ order, const int32_t *__restrict__ counters, const int32_t *__restrict__ pq,
                                    const int32_t *__restrict__ pr, const int32_t *__restrict__ ped,
                                    int32_t *__restrict__ cq, int32_t *__restrict__ cr, int32_t *__restrict__ ce, int64_t n_pairs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs || i >= (int64_t)counters[CNT_VISITS]) return;
  const uint32_t p = order[i];
  cq[i] = pq[p]; cr[i] = pr[p]; ce[i] = ped[p];
}
__global__ void cpair_ptr_kernel(const uint32_t *__restrict__ off, const int32_t *__restrict__ counters, int32_t *__restrict__ cpair_ptr) {
  const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= counters[CNT_COMPS]) cpair_ptr[c] = (int32_t)off[4 * (size_t)c];
}

// bucket of every read from the CSR: a warp takes 32 buckets, a lane writes the reads of its own bucket when they are
// few, the warp together those of the bigger ones
__global__ void bucket_of_read_kernel(const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads, int32_t R, int32_t U,
                                      int32_t *__restrict__ bor) {
  const int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int32_t lo = 0, n = 0;
  if (b < U) { lo = bucket_ptr[b]; n = bucket_ptr[b + 1] - lo; }
  if (n <= 16) for (int32_t x = 0; x < n; x++) bor[bucket_reads[lo + x]] = b;
  uint32_t big = __ballot_sync(0xFFFFFFFFu, n > 16);
  while (big) {
    const int l = __ffs(big) - 1;
    big &= big - 1u;
    const int32_t blo = __shfl_sync(0xFFFFFFFFu, lo, l), bn = __shfl_sync(0xFFFFFFFFu, n, l), bb = __shfl_sync(0xFFFFFFFFu, b, l);
    for (int32_t x = lane; x < bn; x += 32) bor[bucket_reads[blo + x]] = bb;
  }
}

// ------------------------------------------------------------------------------------------ tiers
constexpr int MED_READS = 128, MED_BUCKETS = 128, MED_CALLS = 256, MED_PAIRS = 256;

// reads and genotype entries per component: one thread per bucket in component order, one atomic per run of equal
// components inside a warp (the list is grouped by component; a giant component is 1/32 of its buckets in atomics)
__global__ void comp_stats_kernel(const int32_t *__restrict__ comp_buckets, const int32_t *__restrict__ comp_of_b,
                                  const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads,
                                  const int32_t *__restrict__ geno_ptr, int32_t U, int32_t *__restrict__ comp_reads,
                                  int64_t *__restrict__ comp_calls) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int32_t comp = -1, reads = 0, nnz = 0;
  if (i < U) {
    const int32_t b = comp_buckets[i];
    comp = comp_of_b[b];
    if (comp >= 0) {
      reads = bucket_ptr[b + 1] - bucket_ptr[b];
      for (int32_t x = bucket_ptr[b]; x < bucket_ptr[b + 1]; x++) { const int32_t r = bucket_reads[x]; nnz += geno_ptr[r + 1] - geno_ptr[r]; }
    }
  }
  for (int o = 1; o < 32; o <<= 1) {                          // segmented sums towards the first lane of every run
    const int32_t c2 = __shfl_down_sync(PCB_FULL, comp, o), r2 = __shfl_down_sync(PCB_FULL, reads, o), n2 = __shfl_down_sync(PCB_FULL, nnz, o);
    if (lane + o < 32 && c2 == comp) { reads += r2; nnz += n2; }
  }
  const int32_t prev = __shfl_up_sync(PCB_FULL, comp, 1);
  if (comp >= 0 && (lane == 0 || prev != comp)) {
    atomicAdd(&comp_reads[comp], reads);
    atomicAdd((unsigned long long *)&comp_calls[comp], (unsigned long long)nnz);
  }
}

// Components of more than 32 reads: giant if above huge_reads (replayed by the whole device, merge_huge.cuh), else medium
// if a local copy fits the medium kernel's shared memory, else large.  force: FORCE_LARGE puts every component on the
// large list (the PCB_REPLAY_THREAD / PCB_REPLAY_WARP shapes), FORCE_HUGE on the giant list (PCB_REPLAY_HUGE).
// comp_calls[comp] = genotype entries of the component's reads is also the length of its call region (an upper bound
// of its rows' calls); a scan turns the counts into the regions' offsets.
enum { FORCE_NONE = 0, FORCE_LARGE = 1, FORCE_HUGE = 2 };
__global__ void classify_kernel(MergeCtx c, const int32_t *__restrict__ comp_reads, const int64_t *__restrict__ comp_calls, int32_t force,
                                int32_t huge_reads, int32_t huge_buckets, int32_t *__restrict__ counters, int32_t *__restrict__ medium,
                                int32_t *__restrict__ large, int32_t *__restrict__ hcomp, int32_t *__restrict__ huge_of_comp) {
  const int32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp >= counters[CNT_COMPS]) return;
  const int32_t reads = comp_reads[comp];
  const int64_t nnz = comp_calls[comp];
  huge_of_comp[comp] = -1;
  if (force == FORCE_NONE && reads <= SMALL_READS) return;
  // giant: many reads, or many BUCKETS (a chain of small buckets is what the rounds parallelise; a few big buckets -- collisions
  // -- is what one warp with shared-memory state does well)
  const int32_t n_b = c.comp_ptr[comp + 1] - c.comp_ptr[comp];
  if (force == FORCE_HUGE || (force == FORCE_NONE && (reads > huge_reads || (huge_buckets > 0 && n_b > huge_buckets && reads > MED_READS)))) {
    const int32_t h = atomicAdd(&counters[CNT_HUGE], 1);
    hcomp[h] = comp; huge_of_comp[comp] = h;
    return;
  }
  const bool fits = force == FORCE_NONE && reads <= MED_READS && c.comp_ptr[comp + 1] - c.comp_ptr[comp] <= MED_BUCKETS &&
                    c.cpair_ptr[comp + 1] - c.cpair_ptr[comp] <= MED_PAIRS && nnz <= MED_CALLS;
  if (fits) medium[atomicAdd(&counters[CNT_MEDIUM], 1)] = comp;
  else large[atomicAdd(&counters[CNT_LARGE], 1)] = comp;
}

// scratch of the large components (one thread each): table capacity, list length
__global__ void large_scratch_kernel(const int32_t *__restrict__ comp_reads, const int64_t *__restrict__ comp_calls,
                                     const int32_t *__restrict__ list, int32_t n, int32_t *__restrict__ slot_cap,
                                     int64_t *__restrict__ slot_sz, int64_t *__restrict__ list_sz) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { slot_sz[i] = 0; list_sz[i] = 0; return; }
  const unsigned long long reads = (unsigned long long)comp_reads[list[i]], nnz = (unsigned long long)comp_calls[list[i]];
  unsigned long long need = 2ull * (reads > nnz ? reads : nnz);
  if (need < 2) need = 2;
  unsigned long long cap = 2;
  while (cap < need) cap <<= 1;
  slot_cap[i] = (int32_t)cap;
  slot_sz[i] = (int64_t)cap;
  list_sz[i] = (int64_t)(2ull * (reads + nnz) + 8);      // 4 counters, touched list, run list (merge_warp.cuh)
}

// ------------------------------------------------------------------------------------------ replay kernels
// One warp per component of at most 32 reads, one lane per read (merge_small.cuh).  A component it is not eligible
// for (more than 32 live mutations, huge qualities) goes on the declined list -- or, when a local copy would not fit
// the medium kernel either, on the late list the host looks at after the step.
// WARPS per block: an SM releases registers and shared memory per BLOCK, so a block whose other warps are done keeps their
// share until its slowest component is through (25 % of the warp slots busy with 4 warps per block, profiles/r02q);
// PCB_OPT_SMALL_WARPS picks the instantiation.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) component_kernel_small(MergeCtx c, int32_t n_comp, int32_t *__restrict__ counters,
                                                                     int32_t *__restrict__ declined, int32_t *__restrict__ late) {
  __shared__ SmallSmem sm_all[WARPS];
  __shared__ int32_t jthr[33];
  __shared__ int32_t tally[3];                       // links, tests, warps that are done
  for (int i = threadIdx.x; i < 33; i += WARPS * 32) jthr[i] = c.prm.jthr[i];
  if (threadIdx.x < 3) tally[threadIdx.x] = 0;
  if (WARPS > 1) __syncthreads(); else __syncwarp();
  const int32_t comp = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (comp < n_comp) {
    const int rc = process_component_small(c, comp, lane, &sm_all[threadIdx.x >> 5], jthr, tally, MED_CALLS, MED_PAIRS);
    if (lane == 0 && rc == 1) declined[atomicAdd(&counters[CNT_DECLINED], 1)] = comp;
    if (lane == 0 && rc == 2) late[atomicAdd(&counters[CNT_LATE], 1)] = comp;
  }
  // links / tests: one atomic per block and counter, spread over 32 slots (status[4 + 2 * slot + which]), added by
  // whichever warp finishes last (no barrier at the end)
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    if (atomicAdd(&tally[2], 1) == WARPS - 1) {
      const int32_t links = atomicAdd(&tally[0], 0), tests = atomicAdd(&tally[1], 0);
      if (links) atomicAdd(&c.status[4 + 2 * (blockIdx.x & 31)], links);
      if (tests) atomicAdd(&c.status[5 + 2 * (blockIdx.x & 31)], tests);
    }
  }
}

// The generic warp replay (merge_warp.cuh) on a LOCAL copy of one component: reads, calls, buckets and visiting pairs
// are numbered 0.. inside the component and live in shared memory together with the cluster state and the scratch
// table, so the pointer chasing of the replay costs shared-memory round trips; rows are written out through the map
// back to the caller's read numbers.  One warp per block.
struct __align__(16) MediumSmem {
  GenoSlot slots[2 * MED_CALLS];
  int64_t sl_key[MED_READS], o_key[MED_READS], o_calloff[MED_READS];
  int32_t lists[8 + 2 * (MED_READS + MED_CALLS)];
  int32_t rd_slot[MED_READS], rd_next[MED_READS], sl_parent[MED_READS], sl_head[MED_READS], sl_tail[MED_READS], sl_size[MED_READS];
  int32_t rid[MED_READS], lex[MED_READS], bc[MED_READS], up[MED_READS], bor[MED_READS], gptr[MED_READS + 1];
  int32_t gmut[MED_CALLS], gq[MED_CALLS];
  int32_t bptr[MED_BUCKETS + 1], gid[MED_BUCKETS];
  int32_t cq[MED_PAIRS], cr[MED_PAIRS], ce[MED_PAIRS];
  int32_t o_root[MED_READS], o_pos[MED_READS], o_size[MED_READS], o_bc[MED_READS], o_up[MED_READS], o_calln[MED_READS];
  int32_t comp_ptr[2], cpair_ptr[2];
  int64_t call_off[1];
};

__global__ void __launch_bounds__(32) component_kernel_medium(MergeCtx c, const int32_t *__restrict__ list, const int32_t *__restrict__ n_list) {
  extern __shared__ __align__(16) unsigned char medium_raw[];
  MediumSmem *sm = (MediumSmem *)medium_raw;
  const int lane = threadIdx.x;
  const int32_t n = *n_list;
  for (int32_t it = (int32_t)blockIdx.x; it < n; it += (int32_t)gridDim.x) {
    const int32_t comp = list[it];
    const int32_t b_lo = c.comp_ptr[comp], B = c.comp_ptr[comp + 1] - b_lo;
    // buckets and reads
    int32_t n_reads = 0;
    for (int32_t bi = 0; bi < B; bi++) {                      // (B <= 128: a short uniform loop)
      const int32_t b = c.comp_buckets[b_lo + bi];
      const int32_t x0 = c.bucket_ptr[b], sz = c.bucket_ptr[b + 1] - x0;
      if (lane == 0) { sm->bptr[bi] = n_reads; sm->gid[bi] = b; }
      for (int32_t k = lane; k < sz; k += 32) {
        const int32_t r = c.bucket_reads[x0 + k], i = n_reads + k;
        sm->rid[i] = r; sm->lex[i] = c.lexrank[r]; sm->bc[i] = c.bc_id[r]; sm->up[i] = c.up_id ? c.up_id[r] : 0; sm->bor[i] = bi;
        sm->rd_slot[i] = -1; sm->rd_next[i] = -1;
        sm->o_size[i] = PCB_FILL; sm->o_key[i] = INT64_MIN; sm->o_bc[i] = PCB_FILL; sm->o_up[i] = PCB_FILL;
        sm->o_calloff[i] = INT64_MIN; sm->o_calln[i] = PCB_FILL;
      }
      n_reads += sz;
    }
    if (lane == 0) { sm->bptr[B] = n_reads; sm->comp_ptr[0] = 0; sm->comp_ptr[1] = B; }
    __syncwarp();
    // calls: prefix of the per-read counts, then the entries
    int32_t run = 0;
    for (int32_t i0 = 0; i0 < n_reads; i0 += 32) {
      const int32_t i = i0 + lane;
      int32_t cnt = 0;
      if (i < n_reads) { const int32_t r = sm->rid[i]; cnt = c.geno_ptr[r + 1] - c.geno_ptr[r]; }
      int32_t inc = cnt;
      for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(PCB_FULL, inc, o); if (lane >= o) inc += t; }
      if (i < n_reads) sm->gptr[i] = run + inc - cnt;
      run += __shfl_sync(PCB_FULL, inc, 31);
    }
    if (lane == 0) sm->gptr[n_reads] = run;
    __syncwarp();
    for (int32_t i = 0; i < n_reads; i++) {
      const int32_t r = sm->rid[i], g0 = c.geno_ptr[r], cnt = sm->gptr[i + 1] - sm->gptr[i];
      for (int32_t e = lane; e < cnt; e += 32) { sm->gmut[sm->gptr[i] + e] = c.geno_mut[g0 + e]; sm->gq[sm->gptr[i] + e] = c.geno_q[g0 + e]; }
    }
    // visiting pairs in local bucket numbers
    const int32_t p_lo = c.cpair_ptr[comp], P = c.cpair_ptr[comp + 1] - p_lo;
    for (int32_t p = lane; p < P; p += 32) {
      const int32_t ga = c.cpair_q[p_lo + p], gb = c.cpair_r[p_lo + p];
      int32_t la = 0, lb = 0;
      for (int32_t bi = 0; bi < B; bi++) { if (sm->gid[bi] == ga) la = bi; if (sm->gid[bi] == gb) lb = bi; }
      sm->cq[p] = la; sm->cr[p] = lb; sm->ce[p] = c.cpair_ed[p_lo + p];
    }
    for (int32_t i = lane; i < 2 * MED_CALLS; i += 32) { GenoSlot z; z.key = 0; z.obs1 = z.obs2 = z.pad = 0; z.sum1 = z.sum2 = 0; sm->slots[i] = z; }
    if (lane == 0) {
      sm->cpair_ptr[0] = 0; sm->cpair_ptr[1] = P;
      sm->call_off[0] = c.call_off ? c.call_off[comp] : (int64_t)atomicAdd((unsigned long long *)c.call_cursor, (unsigned long long)run);
    }
    __syncwarp();
    MergeCtx lc = c;
    lc.bucket_ptr = sm->bptr; lc.bucket_reads = nullptr; lc.bucket_of_read = sm->bor; lc.lexrank = sm->lex; lc.bc_id = sm->bc;
    lc.up_id = c.up_id ? sm->up : nullptr;
    lc.geno_ptr = sm->gptr; lc.geno_mut = sm->gmut; lc.geno_q = sm->gq;
    lc.comp_ptr = sm->comp_ptr; lc.comp_buckets = nullptr;
    lc.cpair_ptr = sm->cpair_ptr; lc.cpair_q = sm->cq; lc.cpair_r = sm->cr; lc.cpair_ed = sm->ce;
    lc.call_off = sm->call_off; lc.bucket_gid = sm->gid;
    lc.read_root = sm->o_root; lc.read_pos = sm->o_pos; lc.row_size = sm->o_size; lc.row_key = sm->o_key;
    lc.row_bc = sm->o_bc; lc.row_up = sm->o_up; lc.row_call_off = sm->o_calloff; lc.row_call_n = sm->o_calln;
    MergeState st;
    st.r0 = 0;
    st.rd_slot = sm->rd_slot; st.rd_next = sm->rd_next; st.sl_parent = sm->sl_parent; st.sl_head = sm->sl_head;
    st.sl_tail = sm->sl_tail; st.sl_size = sm->sl_size; st.sl_key = sm->sl_key;
    st.slots = sm->slots; st.slot_mask = 2 * MED_CALLS - 1;
    st.lists = sm->lists; st.list_cap = MED_READS + MED_CALLS;
    process_component_warp(lc, st, 0, lane);
    __syncwarp();
    // rows out, through the map back to the caller's numbering (creation keys name local buckets)
    for (int32_t i = lane; i < n_reads; i += 32) {
      const int32_t r = sm->rid[i];
      c.read_root[r] = sm->rid[sm->o_root[i]];
      c.read_pos[r] = sm->o_pos[i];
      c.row_size[r] = sm->o_size[i];
      int64_t key = sm->o_key[i];
      if (key >= 0) key = ((int64_t)sm->gid[(int32_t)(key >> 32)] << 32) | (key & 0xFFFFFFFFll);
      c.row_key[r] = key;
      c.row_bc[r] = sm->o_bc[i]; c.row_up[r] = sm->o_up[i];
      c.row_call_off[r] = sm->o_calloff[i]; c.row_call_n[r] = sm->o_calln[i];
    }
    __syncwarp();
  }
}

// scratch of large component i of a list (global memory)
struct LargeScratch {
  const int32_t *list;
  const int32_t *slot_cap;
  const int64_t *slot_off, *list_off;
};
__device__ __forceinline__ MergeState large_state(const MergeCtx &c, const LargeScratch &s, int32_t i) {
  MergeState st;
  st.r0 = 0;
  st.rd_slot = c.rd_slot; st.rd_next = c.rd_next; st.sl_parent = c.sl_parent; st.sl_head = c.sl_head;
  st.sl_tail = c.sl_tail; st.sl_size = c.sl_size; st.sl_key = c.sl_key;
  st.slots = c.slots + s.slot_off[i];
  st.slot_mask = s.slot_cap[i] - 1;
  st.lists = c.lists + s.list_off[i];
  st.list_cap = (s.list_off[i + 1] - s.list_off[i]) / 2 - 4;     // region = 8 + 2 * list_cap ints
  return st;
}

constexpr int COMP_BLOCK = 128;
// the generic warp replay on global state, the caller's numbering: one warp per listed component
__global__ void __launch_bounds__(COMP_BLOCK) component_kernel_large(MergeCtx c, LargeScratch s, int32_t n) {
  const int32_t i = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (i >= n) return;
  process_component_warp(c, large_state(c, s, i), s.list[i], threadIdx.x & 31);
}

// PCB_REPLAY_THREAD: one THREAD per component running the host-testable process_component() phases
// (merge_core.h); divergent lanes serialise, so only every (32 / per_warp)-th lane is used.
constexpr int THREAD_BLOCK = 128;
template <int PHASE>
__global__ void __launch_bounds__(THREAD_BLOCK) component_kernel_thread(MergeCtx c, LargeScratch s, int32_t n, int32_t per_warp) {
  const int32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t stride = 32 / per_warp;
  if ((tid & 31) % stride != 0) return;
  const int32_t i = (tid >> 5) * per_warp + (tid & 31) / stride;
  if (i >= n) return;
  const MergeState st = large_state(c, s, i);
  if (PHASE == 1) component_seed(c, st, s.list[i]);
  if (PHASE == 2) component_main(c, st, s.list[i]);
  if (PHASE == 3) component_consensus(c, st, s.list[i]);
}

// ------------------------------------------------------------------------------------------ giant components
// Kernels around merge_huge.cuh.  Warps take units (buckets, window rows, cluster heads) from a list by ticket; each warp
// owns one scratch of a pool (genotype table of `cap` slots + touched / run lists).
constexpr int HUGE_BLOCK = 128;
struct HugePool {
  GenoSlot *slots;
  int32_t *lists;
  int32_t cap, n_warps;
};
__device__ __forceinline__ MergeState huge_state(const MergeCtx &c, const HugePool &pool, int32_t w) {
  MergeState st;
  st.r0 = 0;
  st.rd_slot = c.rd_slot; st.rd_next = c.rd_next; st.sl_parent = c.sl_parent; st.sl_head = c.sl_head;
  st.sl_tail = c.sl_tail; st.sl_size = c.sl_size; st.sl_key = c.sl_key;
  st.slots = pool.slots + (size_t)w * (size_t)pool.cap;
  st.slot_mask = pool.cap - 1;
  st.lists = pool.lists + (size_t)w * (size_t)(8 + 2 * (size_t)pool.cap);
  st.list_cap = pool.cap;
  return st;
}
__device__ __forceinline__ int32_t huge_ticket(int32_t *counter, int lane) {
  int32_t i = 0;
  if (lane == 0) i = atomicAdd(counter, 1);
  return __shfl_sync(PCB_FULL, i, 0);
}

__global__ void huge_rows_kernel(const int32_t *__restrict__ cpair_ptr, const int32_t *__restrict__ hcomp, int32_t n_huge, uint32_t *__restrict__ hrows) {
  const int32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h <= n_huge) hrows[h] = h < n_huge ? (uint32_t)(cpair_ptr[hcomp[h] + 1] - cpair_ptr[hcomp[h]]) : 0u;
}
// the buckets the seed step has work for; the largest scratch one of them needs
__global__ void huge_seedlist_kernel(MergeCtx c, HugeDev h, int32_t U, int32_t cap, const uint32_t *__restrict__ hrow_ptr, int32_t *__restrict__ seedl) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) h.k->total = (int32_t)hrow_ptr[h.n_huge];
  if (i >= U) return;
  const int32_t b = c.comp_buckets[i], comp = h.comp_of_b[b];
  if (comp < 0 || h.huge_of_comp[comp] < 0) return;
  const int32_t n = c.bucket_ptr[b + 1] - c.bucket_ptr[b];
  if (n < 2) return;
  seedl[atomicAdd(&h.k->n_seed, 1)] = b;
  int32_t nnz = 0;
  for (int32_t x = c.bucket_ptr[b]; x < c.bucket_ptr[b + 1]; x++) { const int32_t r = c.bucket_reads[x]; nnz += c.geno_ptr[r + 1] - c.geno_ptr[r]; }
  const int32_t need = huge_need(nnz, n);
  if (need > cap) atomicMax(&h.k->max_need, need);
}
// which: 0 = the seed list with ordinary scratch, 1 = the big list with big scratch
__global__ void __launch_bounds__(HUGE_BLOCK) huge_seed_kernel(MergeCtx c, HugeDev h, HugePool pool, const int32_t *__restrict__ list, int which) {
  const int32_t w = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const MergeState st = huge_state(c, pool, w);
  const int32_t n = which ? h.k->n_big : h.k->n_seed;
  for (;;) {
    const int32_t i = huge_ticket(&h.k->tk[which], lane);
    if (i >= n) break;
    huge_seed_item(c, st, h, pool.cap, which != 0, list[i], lane);
  }
}
// one thread per read: which == 0 after the seed step (genotype entries per cell), which == 1 after the main loop
// (rows of the lone reads, list of the cluster heads)
__global__ void huge_reads_kernel(MergeCtx c, HugeDev h, int which) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  const int32_t comp = h.comp_of_b[c.bucket_of_read[r]];
  if (comp < 0) return;
  const int32_t hi = h.huge_of_comp[comp];
  if (hi < 0) return;
  MergeState st;
  st.r0 = 0; st.rd_slot = c.rd_slot; st.rd_next = c.rd_next; st.sl_parent = c.sl_parent; st.sl_head = c.sl_head;
  st.sl_tail = c.sl_tail; st.sl_size = c.sl_size; st.sl_key = c.sl_key; st.slots = nullptr; st.slot_mask = 0; st.lists = nullptr; st.list_cap = 0;
  if (which == 0) huge_count_read(c, st, r);
  else huge_finish_read(c, st, h, hi, r);
}
__global__ void huge_advance_kernel(HugeDev h) { huge_advance(h); }
__global__ void __launch_bounds__(HUGE_BLOCK) huge_reserve_kernel(MergeCtx c, HugeDev h) {
  const int32_t idx = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (idx >= h.k->n_window) return;
  HugePool none{nullptr, nullptr, 0, 0};
  huge_reserve_item(c, huge_state(c, none, 0), h, idx, threadIdx.x & 31);
}
// which: 0 = the window with ordinary scratch, 1 = the winners those warps passed on, with big scratch
__global__ void __launch_bounds__(HUGE_BLOCK) huge_exec_kernel(MergeCtx c, HugeDev h, HugePool pool, int which, int have_big) {
  const int32_t w = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const MergeState st = huge_state(c, pool, w);
  const int32_t n = which ? h.k->n_heavy : h.k->n_window;
  for (;;) {
    const int32_t i = huge_ticket(which ? &h.k->ticket_heavy : &h.k->ticket, lane);
    if (i >= n) break;
    if (which) huge_heavy_item(c, st, h, pool.cap, i, lane);
    else huge_row_item(c, st, h, pool.cap, have_big != 0, i, lane);
  }
}
__global__ void __launch_bounds__(HUGE_BLOCK) huge_head_kernel(MergeCtx c, HugeDev h, HugePool pool, int which) {
  const int32_t w = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const MergeState st = huge_state(c, pool, w);
  const int32_t n = which ? h.k->n_big : h.k->n_heads;
  const int32_t *list = which ? h.big : h.heads;
  for (;;) {
    const int32_t i = huge_ticket(&h.k->tk[2 + which], lane);
    if (i >= n) break;
    huge_head_item(c, st, h, pool.cap, which != 0, list[i], lane);
  }
}
__global__ void huge_tally_kernel(HugeDev h, int32_t *status) {
  atomicAdd(&status[2], h.k->links);
  atomicAdd(&status[3], h.k->tests);
}

__global__ void init_outputs_kernel(int32_t R, int64_t nnz, MergeOut out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) {
    out.read_root[i] = out.read_pos[i] = out.row_size[i] = out.row_bc[i] = out.row_up[i] = out.row_call_n[i] = PCB_FILL;
    out.row_key[i] = INT64_MIN; out.row_call_off[i] = INT64_MIN;
  }
  if (i < nnz) out.call_mut[i] = PCB_FILL;
}

// ------------------------------------------------------------------------------------------ host
namespace {

// cluster state in global memory (size R), shared by the large and the giant tier: allocated when first needed
struct ClusterArrays {
  DevBuf<int32_t> rd_slot, rd_next, sl_parent, sl_head, sl_tail, sl_size;
  DevBuf<int64_t> sl_key;
  bool ready = false;
  void attach(pcb_ctx *ctx, MergeCtx &c, int32_t R) {
    if (!ready) {
      rd_slot.alloc(ctx, R); rd_next.alloc(ctx, R); sl_parent.alloc(ctx, R); sl_head.alloc(ctx, R); sl_tail.alloc(ctx, R); sl_size.alloc(ctx, R);
      sl_key.alloc(ctx, R);
      PCB_CUDA(cudaMemsetAsync(rd_slot.p, 0xFF, (size_t)R * sizeof(int32_t), ctx->stream));
      PCB_CUDA(cudaMemsetAsync(rd_next.p, 0xFF, (size_t)R * sizeof(int32_t), ctx->stream));
      ready = true;
    }
    c.rd_slot = rd_slot.p; c.rd_next = rd_next.p; c.sl_parent = sl_parent.p; c.sl_head = sl_head.p;
    c.sl_tail = sl_tail.p; c.sl_size = sl_size.p; c.sl_key = sl_key.p;
  }
};

// The generic replay on global state for a list of components (the large tier; every component under the forced shapes).
void run_large(pcb_ctx *ctx, MergeCtx c, ClusterArrays &cl, const int32_t *comp_reads, const int64_t *comp_calls, const int32_t *list, int32_t n,
               int32_t R, int shape) {
  if (n <= 0) return;
  DevBuf<int32_t> slot_cap(ctx, (size_t)n + 1);
  DevBuf<int64_t> slot_off(ctx, (size_t)n + 1), list_off(ctx, (size_t)n + 1);
  PCB_LAUNCH(ctx, large_scratch_kernel, grid_for((size_t)n + 1, 128), 128, 0, comp_reads, comp_calls, list, n, slot_cap.p, slot_off.p, list_off.p);
  exclusive_scan_i64(ctx, slot_off.p, slot_off.p, (size_t)n + 1);
  exclusive_scan_i64(ctx, list_off.p, list_off.p, (size_t)n + 1);
  int64_t n_slots = 0, n_list = 0;
  PCB_CUDA(cudaMemcpyAsync(&n_slots, slot_off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(&n_list, list_off.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  DevBuf<GenoSlot> slots(ctx, (size_t)n_slots);
  DevBuf<int32_t> lists(ctx, (size_t)n_list + 1);
  PCB_CUDA(cudaMemsetAsync(slots.p, 0, (size_t)n_slots * sizeof(GenoSlot), ctx->stream));     // all-zero = empty
  cl.attach(ctx, c, R);
  c.slots = slots.p; c.lists = lists.p;
  LargeScratch s{list, slot_cap.p, slot_off.p, list_off.p};
  if (shape == PCB_REPLAY_THREAD) {
    const int per_warp = 4;               // best of a 1..32 sweep (divergent lanes serialise)
    size_t n_threads = ((size_t)n + per_warp - 1) / per_warp * 32;
    // three launches, one per phase: each kernel's code fits the instruction cache and its threads share a loop nest
    PCB_LAUNCH(ctx, component_kernel_thread<1>, grid_for(n_threads, THREAD_BLOCK), THREAD_BLOCK, 0, c, s, n, per_warp);
    PCB_LAUNCH(ctx, component_kernel_thread<2>, grid_for(n_threads, THREAD_BLOCK), THREAD_BLOCK, 0, c, s, n, per_warp);
    PCB_LAUNCH(ctx, component_kernel_thread<3>, grid_for(n_threads, THREAD_BLOCK), THREAD_BLOCK, 0, c, s, n, per_warp);
  } else {
    PCB_LAUNCH(ctx, component_kernel_large, grid_for((size_t)n * 32, COMP_BLOCK), COMP_BLOCK, 0, c, s, n);
  }
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));          // the scratch above is released when this returns
}

// a pool of warp scratches (zeroed tables)
struct HugePoolBuf {
  DevBuf<GenoSlot> slots;
  DevBuf<int32_t> lists;
  HugePool pool{nullptr, nullptr, 0, 0};
  void make(pcb_ctx *ctx, int32_t cap, int32_t n_warps) {
    slots.alloc(ctx, (size_t)cap * (size_t)n_warps);
    lists.alloc(ctx, (size_t)(8 + 2 * (size_t)cap) * (size_t)n_warps);
    PCB_CUDA(cudaMemsetAsync(slots.p, 0, (size_t)cap * (size_t)n_warps * sizeof(GenoSlot), ctx->stream));
    pool = HugePool{slots.p, lists.p, cap, n_warps};
  }
  unsigned grid() const { return (unsigned)((pool.n_warps * 32 + HUGE_BLOCK - 1) / HUGE_BLOCK); }
};
int32_t pow2_at_least(int64_t v) { int64_t c = 16; while (c < v) c <<= 1; return (int32_t)(c > (1ll << 30) ? (1ll << 30) : c); }

// The giant tier (merge_huge.cuh): seed step per bucket, main loop in rounds of deterministic reservations, consensus per
// cluster -- all giant components of the call together.
void run_huge(pcb_ctx *ctx, MergeCtx c, ClusterArrays &cl, const int32_t *comp_of_b, const int32_t *huge_of_comp, const int32_t *hcomp,
              int32_t n_huge, int32_t R, int32_t U) {
  if (n_huge <= 0) return;
  HostTrace tr(ctx, "giant tier");
  const int32_t cap = ctx->huge_cap, W = ctx->huge_window;
  cl.attach(ctx, c, R);
  DevBuf<uint32_t> hrow_ptr(ctx, (size_t)n_huge + 1);
  DevBuf<int32_t> seedl(ctx, (size_t)U + 1), cell_nnz(ctx, R), win(ctx, W), carry0(ctx, W), carry1(ctx, W), heavy(ctx, W), heads(ctx, R), big(ctx, R);
  DevBuf<unsigned long long> resv(ctx, R), cursor(ctx, n_huge);
  DevBuf<HugeCounters> k(ctx, 1);
  PCB_CUDA(cudaMemsetAsync(k.p, 0, sizeof(HugeCounters), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(resv.p, 0xFF, (size_t)R * sizeof(unsigned long long), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(cell_nnz.p, 0, (size_t)R * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)n_huge * sizeof(unsigned long long), ctx->stream));
  c.cell_nnz = cell_nnz.p;
  HugeDev h;
  h.resv = resv.p; h.comp_of_b = comp_of_b; h.huge_of_comp = huge_of_comp; h.hcomp = hcomp; h.hrow_ptr = (const int32_t *)hrow_ptr.p;
  h.cursor = cursor.p; h.n_huge = n_huge; h.W = W; h.win = win.p; h.carry[0] = carry0.p; h.carry[1] = carry1.p; h.heavy = heavy.p;
  h.heads = heads.p; h.big = big.p; h.k = k.p;
  PCB_LAUNCH(ctx, huge_rows_kernel, grid_for((size_t)n_huge + 1, 256), 256, 0, c.cpair_ptr, hcomp, n_huge, hrow_ptr.p);
  exclusive_scan_u32(ctx, hrow_ptr.p, hrow_ptr.p, (size_t)n_huge + 1);
  PCB_LAUNCH(ctx, huge_seedlist_kernel, grid_for(U, 256), 256, 0, c, h, U, cap, hrow_ptr.p, seedl.p);
  HugeCounters hk;
  auto read_counters = [&] {
    PCB_CUDA(cudaMemcpyAsync(&hk, k.p, sizeof(hk), cudaMemcpyDeviceToHost, ctx->stream));
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  };
  read_counters();
  // scratch: one per resident warp; a few big ones sized for the largest unit that asked (re-made when one asks for more)
  HugePoolBuf pool, bigpool;
  pool.make(ctx, cap, ctx->sm_count * ctx->huge_warps_per_sm);
  auto size_big = [&](int32_t need) {
    const int32_t bcap = pow2_at_least(need);
    int64_t nw = ((int64_t)1 << 24) / bcap;
    bigpool.make(ctx, bcap, (int32_t)(nw < 1 ? 1 : nw > 64 ? 64 : nw));
  };
  if (hk.max_need > cap) size_big(hk.max_need);
  tr.lap("lists");

  // ---- seed step
  PCB_LAUNCH(ctx, huge_seed_kernel, pool.grid(), HUGE_BLOCK, 0, c, h, pool.pool, seedl.p, 0);
  if (bigpool.pool.cap) PCB_LAUNCH(ctx, huge_seed_kernel, bigpool.grid(), HUGE_BLOCK, 0, c, h, bigpool.pool, big.p, 1);
  PCB_CUDA(cudaMemsetAsync(&k.p->n_big, 0, sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(&k.p->max_need, 0, sizeof(int32_t), ctx->stream));
  PCB_LAUNCH(ctx, huge_reads_kernel, grid_for(R, 256), 256, 0, c, h, 0);

  // ---- main loop: rounds; the host looks at the counters once per batch
  const int batch = 16;
  int64_t rounds = 0;
  if (hk.total > 0)
    for (;;) {
      for (int r = 0; r < batch; r++) {
        PCB_LAUNCH(ctx, huge_advance_kernel, 1, 1, 0, h);
        PCB_LAUNCH(ctx, huge_reserve_kernel, grid_for((size_t)W * 32, HUGE_BLOCK), HUGE_BLOCK, 0, c, h);
        PCB_LAUNCH(ctx, huge_exec_kernel, pool.grid(), HUGE_BLOCK, 0, c, h, pool.pool, 0, bigpool.pool.cap ? 1 : 0);
        if (bigpool.pool.cap) PCB_LAUNCH(ctx, huge_exec_kernel, bigpool.grid(), HUGE_BLOCK, 0, c, h, bigpool.pool, 1, 1);
      }
      read_counters();
      rounds = hk.round;
      if (hk.done) break;
      if (hk.max_need > bigpool.pool.cap) size_big(hk.max_need);       // rows waiting for more scratch than any warp has
      // (the earliest pending row wins every round, or waits at most one batch for the big scratch)
      if (rounds > 2 * (int64_t)hk.total + 4 * batch) throw Error(PCB_E_STATE, "giant tier: the main loop makes no progress");
    }
  tr.lap("rounds");

  // ---- consensus
  PCB_CUDA(cudaMemsetAsync(&k.p->max_need, 0, sizeof(int32_t), ctx->stream));
  PCB_LAUNCH(ctx, huge_reads_kernel, grid_for(R, 256), 256, 0, c, h, 1);
  PCB_LAUNCH(ctx, huge_head_kernel, pool.grid(), HUGE_BLOCK, 0, c, h, pool.pool, 0);
  read_counters();
  if (hk.n_big > 0) {
    if (hk.max_need > bigpool.pool.cap) size_big(hk.max_need);
    PCB_LAUNCH(ctx, huge_head_kernel, bigpool.grid(), HUGE_BLOCK, 0, c, h, bigpool.pool, 1);
  }
  PCB_LAUNCH(ctx, huge_tally_kernel, 1, 1, 0, h, c.status);
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));          // the scratch above is released when this returns
  ctx->stats[PCB_STAT_HUGE_ROUNDS] = rounds;
  ctx->stats[PCB_STAT_HUGE_COMPONENTS] = n_huge;
  tr.lap("consensus");
}

}  // namespace

void merge_dev(pcb_ctx *ctx, const MergeIn &in, const MergeOut &out) {
  const int32_t R = in.R, U = in.U;
  const int64_t P = in.n_pairs;
  if (in.max_diff < 0 || in.max_diff > 3) throw Error(PCB_E_INPUT, "maxDiff must be between 0 and 3");
  if (!(in.min_jaccard > 0.0 && in.min_jaccard <= 1.0)) throw Error(PCB_E_INPUT, "minJaccard must be between 0 and 1");
  if (in.min_matches < 1) throw Error(PCB_E_INPUT, "minMatches must be at least 1");
  if (in.min_qual < 1) throw Error(PCB_E_INPUT, "minQual must be a positive integer");
  if (in.shard_count < 1 || in.shard_rank < 0 || in.shard_rank >= in.shard_count) throw Error(PCB_E_INPUT, "bad shard");
  ctx->stats[PCB_STAT_HUGE_COMPONENTS] = ctx->stats[PCB_STAT_HUGE_ROUNDS] = 0;
  if (R == 0 || U == 0) return;
  const int64_t nnz = in.nnz >= 0 ? in.nnz : (int64_t)read_scalar(ctx, in.geno_ptr + R);
  const int bbits = bits_for((uint64_t)(U > 1 ? U - 1 : 1));
  const bool on_demand = in.planes != nullptr && in.compute_k >= 0;
  const int shape = ctx->replay_shape;
  if (!ctx->medium_attr_set) {
    PCB_CUDA(cudaFuncSetAttribute(component_kernel_medium, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MediumSmem)));
    ctx->medium_attr_set = true;
  }

  // ---- components of the link graph; this shard's components first
  DevBuf<int32_t> counters(ctx, CNT_N + 1), label(ctx, U), comp_buckets(ctx, U), comp_ptr(ctx, (size_t)U + 2), comp_of_b(ctx, U);
  DevBuf<int64_t> cursor(ctx, 1);
  PCB_CUDA(cudaMemsetAsync(counters.p, 0, (CNT_N + 1) * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(cursor.p, 0, sizeof(int64_t), ctx->stream));
  {
    DevBuf<int32_t> parent(ctx, U);
    DevBuf<uint32_t> rank(ctx, (size_t)U + 1);
    DevBuf<uint64_t> k0(ctx, U), k1(ctx, U);
    DevBuf<uint32_t> v0(ctx, U), v1(ctx, U);
    PCB_LAUNCH(ctx, uf_init_kernel, grid_for(U, 256), 256, 0, parent.p, U);
    if (P > 0) PCB_LAUNCH(ctx, uf_union_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, in.pair_ed, P, in.max_diff, parent.p);
    PCB_LAUNCH(ctx, comp_keys_kernel, grid_for(U, 256), 256, 0, parent.p, U, bbits, in.shard_rank, in.shard_count, label.p, k0.p, v0.p,
               counters.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, U, bbits + (in.shard_count > 1 ? 1 : 0));
    uint64_t *sk = w ? k1.p : k0.p;
    uint32_t *sv = w ? v1.p : v0.p;
    PCB_LAUNCH(ctx, comp_flag_kernel, grid_for((size_t)U + 1, 256), 256, 0, sk, U, rank.p);
    exclusive_scan_u32(ctx, rank.p, rank.p, (size_t)U + 1);              // one past the end: the total
    PCB_LAUNCH(ctx, comp_assign_kernel, grid_for(U, 256), 256, 0, sk, sv, rank.p, U, bbits, comp_buckets.p, comp_ptr.p, comp_of_b.p,
               counters.p);
  }

  // ---- adjacency: every pair once, in the row of its smaller bucket
  DevBuf<int32_t> adj_ptr(ctx, (size_t)U + 1), sq, sr, sd;
  const int32_t *adj_nbr = in.pair_r, *adj_d = in.pair_d;
  if (!in.pairs_sorted && P > 0) {
    DevBuf<uint64_t> k0(ctx, P), k1(ctx, P);
    DevBuf<uint32_t> v0(ctx, P), v1(ctx, P);
    sq.alloc(ctx, P); sr.alloc(ctx, P); sd.alloc(ctx, P);
    PCB_LAUNCH(ctx, pair_keys_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, P, bbits, k0.p, v0.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, P, 2 * bbits);
    PCB_LAUNCH(ctx, pair_unpack_kernel, grid_for(P, 256), 256, 0, w ? k1.p : k0.p, w ? v1.p : v0.p, in.pair_d, P, bbits, sq.p, sr.p, sd.p);
    adj_nbr = sr.p; adj_d = sd.p;
  }
  PCB_LAUNCH(ctx, row_ptr_kernel, grid_for((size_t)P + 1, 256), 256, 0, (in.pairs_sorted || P == 0) ? in.pair_q : sq.p, P, U, adj_ptr.p);

  // ---- visiting pairs per component
  DevBuf<int32_t> cpair_ptr(ctx, (size_t)U + 2), cq(ctx, (size_t)P), cr(ctx, (size_t)P), ce(ctx, (size_t)P);
  PCB_CUDA(cudaMemsetAsync(cpair_ptr.p, 0, sizeof(int32_t), ctx->stream));            // (a shard that owns no component)
  if (in.pairs_sorted || P == 0) {
    const size_t ld = (size_t)U + 1;
    DevBuf<uint32_t> g(ctx, ld * (size_t)(in.max_diff > 0 ? in.max_diff : 1));
    PCB_LAUNCH(ctx, visit_cnt_kernel, grid_for(ld, 256), 256, 0, comp_buckets.p, comp_of_b.p, adj_ptr.p, in.pair_ed, U, in.max_diff, g.p);
    for (int32_t e = 0; e < in.max_diff; e++) exclusive_scan_u32(ctx, g.p + e * ld, g.p + e * ld, ld);
    PCB_LAUNCH(ctx, visit_fill_kernel, grid_for(U, 256), 256, 0, comp_buckets.p, comp_of_b.p, comp_ptr.p, counters.p, adj_ptr.p, in.pair_r,
               in.pair_ed, U, in.max_diff, g.p, cpair_ptr.p, cq.p, cr.p, ce.p);
  } else {
    DevBuf<uint32_t> cnt(ctx, 4 * ((size_t)U + 1) + 1);
    PCB_CUDA(cudaMemsetAsync(cnt.p, 0, (4 * ((size_t)U + 1) + 1) * sizeof(uint32_t), ctx->stream));
    PCB_LAUNCH(ctx, visit_count_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_ed, P, in.max_diff, comp_of_b.p, cnt.p, counters.p);
    exclusive_scan_u32(ctx, cnt.p, cnt.p, 4 * ((size_t)U + 1) + 1);
    DevBuf<uint64_t> k0(ctx, P), k1(ctx, P);
    DevBuf<uint32_t> v0(ctx, P), v1(ctx, P);
    PCB_LAUNCH(ctx, cpair_keys_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_ed, P, comp_of_b.p, U, in.max_diff, k0.p, v0.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, P, bbits + 3);
    PCB_LAUNCH(ctx, cpair_gather_kernel, grid_for(P, 256), 256, 0, w ? v1.p : v0.p, counters.p, in.pair_q, in.pair_r, in.pair_ed, cq.p,
               cr.p, ce.p, P);
    PCB_LAUNCH(ctx, cpair_ptr_kernel, grid_for((size_t)U + 1, 256), 256, 0, cnt.p, counters.p, cpair_ptr.p);
  }

  // ---- what the replay reads
  DevBuf<int32_t> bor_own;
  const int32_t *bor = in.bucket_of_read;
  if (!bor) {
    bor_own.alloc(ctx, R);
    PCB_LAUNCH(ctx, bucket_of_read_kernel, grid_for(U, 256), 256, 0, in.bucket_ptr, in.bucket_reads, R, U, bor_own.p);
    bor = bor_own.p;
  }
  DevBuf<int32_t> status(ctx, STATUS_N), medium(ctx, (size_t)U + 1), large(ctx, (size_t)U + 1), declined(ctx, (size_t)U + 1), late(ctx, (size_t)U + 1);
  DevBuf<int32_t> hcomp(ctx, (size_t)U + 1), huge_of_comp(ctx, (size_t)U + 1), comp_reads(ctx, (size_t)U + 1);
  PCB_CUDA(cudaMemsetAsync(status.p, 0, STATUS_N * sizeof(int32_t), ctx->stream));
  MergeCtx c;
  memset(&c, 0, sizeof(c));
  c.R = R; c.U = U;
  c.bucket_ptr = in.bucket_ptr; c.bucket_reads = in.bucket_reads; c.bucket_of_read = bor;
  c.lexrank = in.lexrank; c.bc_id = in.bc_id; c.up_id = in.up_id;
  c.geno_ptr = in.geno_ptr; c.geno_mut = in.geno_mut; c.geno_q = in.geno_q; c.mut_rank = in.mut_rank;
  c.adj_ptr = adj_ptr.p; c.adj_nbr = adj_nbr; c.adj_d = adj_d; c.adj_forward = 1;
  c.planes = on_demand ? (const uint64_t *)in.planes : nullptr; c.blen = on_demand ? in.blen : nullptr;
  c.compute_k = on_demand ? in.compute_k : -1; c.wide = in.wide;
  c.comp_ptr = comp_ptr.p; c.comp_buckets = comp_buckets.p;
  c.cpair_ptr = cpair_ptr.p; c.cpair_q = cq.p; c.cpair_r = cr.p; c.cpair_ed = ce.p;
  c.call_off = nullptr; c.call_cursor = cursor.p;
  DevBuf<int64_t> call_off;
  if (in.shard_layout) {
    DevBuf<uint32_t> lab_off(ctx, (size_t)U + 1);
    call_off.alloc(ctx, (size_t)U + 1);
    PCB_CUDA(cudaMemsetAsync(lab_off.p, 0, ((size_t)U + 1) * sizeof(uint32_t), ctx->stream));
    PCB_LAUNCH(ctx, label_nnz_kernel, grid_for(U, 256), 256, 0, in.bucket_ptr, in.bucket_reads, in.geno_ptr, label.p, U, lab_off.p);
    exclusive_scan_u32(ctx, lab_off.p, lab_off.p, (size_t)U + 1);
    PCB_LAUNCH(ctx, comp_call_off_kernel, grid_for(U, 256), 256, 0, comp_ptr.p, comp_buckets.p, counters.p, lab_off.p, call_off.p);
    c.call_off = call_off.p;
  }
  c.read_root = out.read_root; c.read_pos = out.read_pos; c.row_size = out.row_size; c.row_key = out.row_key;
  c.row_bc = out.row_bc; c.row_up = out.row_up; c.row_call_off = out.row_call_off; c.row_call_n = out.row_call_n;
  c.call_mut = out.call_mut; c.status = status.p;
  c.prm.maxDiff = in.max_diff; c.prm.minQual = in.min_qual; c.prm.minMatches = in.min_matches; c.prm.minJaccard = in.min_jaccard;
  jaccard_thresholds(c.prm);
  const int32_t force = shape == PCB_REPLAY_AUTO ? FORCE_NONE : shape == PCB_REPLAY_HUGE ? FORCE_HUGE : FORCE_LARGE;
  DevBuf<int64_t> comp_calls(ctx, (size_t)U + 2), comp_call_off;
  PCB_CUDA(cudaMemsetAsync(comp_calls.p, 0, ((size_t)U + 2) * sizeof(int64_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(comp_reads.p, 0, ((size_t)U + 1) * sizeof(int32_t), ctx->stream));
  PCB_LAUNCH(ctx, comp_stats_kernel, grid_for(U, 256), 256, 0, comp_buckets.p, comp_of_b.p, in.bucket_ptr, in.bucket_reads, in.geno_ptr, U,
             comp_reads.p, comp_calls.p);
  PCB_LAUNCH(ctx, classify_kernel, grid_for((size_t)U + 1, 128), 128, 0, c, comp_reads.p, comp_calls.p, force, ctx->huge_reads, ctx->huge_buckets, counters.p,
             medium.p, large.p, hcomp.p, huge_of_comp.p);
  if (!in.shard_layout) {                                  // regions in component order
    comp_call_off.alloc(ctx, (size_t)U + 2);
    exclusive_scan_i64(ctx, comp_calls.p, comp_call_off.p, (size_t)U + 1);     // (entries past the last component are never read)
    c.call_off = comp_call_off.p;
  }
  // every output starts as "not produced": entries of reads other shards own, the row fields of non-representative reads
  PCB_LAUNCH(ctx, init_outputs_kernel, grid_for((size_t)(R > nnz ? R : nnz) + 1, 256), 256, 0, R, nnz, out);
  int32_t h_cnt[CNT_N];                                                 // the one round trip of the preprocessing
  PCB_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  const int32_t n_comp = h_cnt[CNT_COMPS], n_medium = h_cnt[CNT_MEDIUM], n_large = h_cnt[CNT_LARGE], n_huge = h_cnt[CNT_HUGE];
  ctx->stats[PCB_STAT_COMPONENTS] = n_comp;
  stage_end(ctx, PCB_STAT_T_PREP_NS);

  // ---- replay
  const unsigned medium_grid = (unsigned)ctx->sm_count * 4u;
  if (n_medium > 0) {                     // the few components of 33..128 reads on a second stream, beside everything else
    PCB_CUDA(cudaEventRecord(ctx->ev_alloc, ctx->stream));
    PCB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_alloc, 0));
    component_kernel_medium<<<(unsigned)n_medium < medium_grid ? (unsigned)n_medium : medium_grid, 32, sizeof(MediumSmem), ctx->copy_stream>>>(
        c, medium.p, counters.p + CNT_MEDIUM);
    ctx->stats[PCB_STAT_KERNEL_LAUNCHES]++;
    PCB_CUDA(cudaGetLastError());
    PCB_CUDA(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
  }
  if (shape == PCB_REPLAY_AUTO && n_comp > 0) {
    if (ctx->small_warps == 1)
      PCB_LAUNCH(ctx, component_kernel_small<1>, grid_for((size_t)n_comp * 32, 32), 32, (size_t)ctx->small_pad, c, n_comp, counters.p, declined.p, late.p);
    else if (ctx->small_warps == 2)
      PCB_LAUNCH(ctx, component_kernel_small<2>, grid_for((size_t)n_comp * 32, 64), 64, 0, c, n_comp, counters.p, declined.p, late.p);
    else
      PCB_LAUNCH(ctx, component_kernel_small<4>, grid_for((size_t)n_comp * 32, 128), 128, 0, c, n_comp, counters.p, declined.p, late.p);
    // whatever it declined (the count is only known on the device): a fixed grid walks the list
    PCB_LAUNCH(ctx, component_kernel_medium, medium_grid / 4u, 32, sizeof(MediumSmem), c, declined.p, counters.p + CNT_DECLINED);
  }
  if (n_medium > 0) PCB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
  ClusterArrays cl;
  run_large(ctx, c, cl, comp_reads.p, comp_calls.p, large.p, n_large, R, shape);
  run_huge(ctx, c, cl, comp_of_b.p, huge_of_comp.p, hcomp.p, n_huge, R, U);
  stage_end(ctx, PCB_STAT_T_REPLAY_NS);

  int32_t h_status[STATUS_N], h_late = 0;
  PCB_CUDA(cudaMemcpyAsync(h_status, status.p, sizeof(h_status), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(&h_late, counters.p + CNT_LATE, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_late > 0) {                       // <= 32 reads, but too many calls or rows for either warp kernel's shared memory
    run_large(ctx, c, cl, comp_reads.p, comp_calls.p, late.p, h_late, R, PCB_REPLAY_WARP);
    PCB_CUDA(cudaMemcpyAsync(h_status, status.p, sizeof(h_status), cudaMemcpyDeviceToHost, ctx->stream));
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  stage_end(ctx, PCB_STAT_T_SCATTER_NS);
  for (int slot = 0; slot < 32; slot++) { h_status[2] += h_status[4 + 2 * slot]; h_status[3] += h_status[5 + 2 * slot]; }
  ctx->stats[PCB_STAT_LINKS] = h_status[2];
  ctx->stats[PCB_STAT_TESTS] = h_status[3];
  if (h_status[0] == 3)
    throw Error(PCB_E_UPTAG, "a clustered read (row of read " + std::to_string(h_status[1]) +
                                 ") has no uptag record: the reference raises KeyError at pacybara_cluster.py:519");
}

}  // namespace pcb
