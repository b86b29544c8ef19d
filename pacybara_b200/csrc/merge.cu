// merge.cu -- cluster merge on the device (rows a4..a13): preprocessing kernels that turn the
// bucket-pair list into per-component work, then one thread per connected component replaying
// the reference's sequential algorithm (merge_core.h).
//
//   components    lock-free union-find over the pairs a pass visits (1 <= ed <= maxDiff);
//                 label = smallest bucket of the component; a component belongs to shard (label mod shard_count)
//                 and everything below only touches this shard's buckets, reads and pairs
//   renumbering   buckets are sorted by (component work descending, component label), reads by
//                 (new bucket, arrival order), and every per-read / per-bucket array the replay touches
//                 is gathered into that order: a component's working set becomes a handful of
//                 contiguous cache lines instead of one sector per read scattered over HBM, and
//                 neighbouring threads (= neighbouring components of similar size) follow similar
//                 control flow.  The renumbering is an identity relabelling: nothing in the algorithm
//                 depends on the numeric value of a read or bucket id (string order of read ids comes
//                 in through `lexrank`, creation keys are mapped back to the old bucket numbers).
//   adjacency     symmetric (a, nbr, d) sorted by (a, nbr) in the new numbering: the lookup table of :236,:266
//   visiting      pairs a pass visits, grouped per component by a STABLE sort on (component, ed),
//                 which keeps the reference's visiting order inside every component
//   replay        process_component(): seed step, greedy merge, consensus
//   scatter       outputs back to the caller's read numbering
// Everything except the replay is HBM-bound streaming; the replay is latency-bound pointer work
// whose parallelism is the number of components (about 10^5 at cfg 2).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "merge_core.h"
#include "merge_small.cuh"
#include "merge_warp.cuh"

namespace pcb {

constexpr uint32_t WORK_CAP = 1023;     // component work classes used for ordering (10 bits)
constexpr int HEAVY_READS = SMALL_READS + 1;   // components too big for the one-lane-per-read kernel (they are numbered first)

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

// label[b] = root; work[root] += reads of b
__global__ void uf_label_kernel(int32_t *parent, const int32_t *__restrict__ bucket_ptr, int32_t U,
                                int32_t *__restrict__ label, uint32_t *__restrict__ work) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  int32_t l = uf_find(parent, b);
  label[b] = l;
  atomicAdd(&work[l], (uint32_t)(bucket_ptr[b + 1] - bucket_ptr[b]));
}

// lab_nnz[root] = genotype entries of the whole component: the call region of a component is laid out by its label,
// which every shard computes identically (so shards can be combined element-wise)
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

// counters: [0] buckets of this shard, [1] components of this shard, [2] heavy components of this shard, [3] reads of this shard
// A component belongs to shard (label mod shard_count); the buckets of other shards sort behind everything and are
// ignored from here on, so the preprocessing below only touches this shard's part of the library.
__global__ void comp_keys_kernel(const int32_t *__restrict__ label, const uint32_t *__restrict__ work, int32_t U, int bbits,
                                 int32_t shard_rank, int32_t shard_count, uint64_t *__restrict__ key, uint32_t *__restrict__ val,
                                 uint32_t *__restrict__ counters) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  uint32_t l = (uint32_t)label[b];
  uint32_t w = work[l] < WORK_CAP ? work[l] : WORK_CAP;
  bool mine = (int32_t)(l % (uint32_t)shard_count) == shard_rank;
  key[b] = ((uint64_t)(mine ? 0 : 1) << (bbits + 10)) | ((uint64_t)(WORK_CAP - w) << bbits) | l;   // mine first, heavy first
  val[b] = (uint32_t)b;
  if (mine) atomicAdd(&counters[0], 1u);
}

__global__ void comp_flag_kernel(const uint64_t *__restrict__ key, int32_t U, uint32_t *__restrict__ flag) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < U) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// sorted position i = new bucket id (this shard's buckets come first).  rank = exclusive scan of the head flags.
__global__ void comp_assign_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, const uint32_t *__restrict__ rank,
                                   const int32_t *__restrict__ bucket_ptr, int32_t U, int32_t n_comp,
                                   int32_t *__restrict__ old_of_new_b, int32_t *__restrict__ new_of_old_b,
                                   int32_t *__restrict__ comp_ptr, int32_t *__restrict__ comp_of_new_b, uint32_t *__restrict__ nb_cnt,
                                   int32_t *__restrict__ comp_label, int bbits, uint32_t heavy_reads, uint32_t *__restrict__ counters) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  const bool head = i == 0 || key[i] != key[i - 1];
  const bool mine = (key[i] >> (bbits + 10)) == 0;
  const int32_t c = (int32_t)rank[i] + (head ? 1 : 0) - 1;
  const int32_t b = (int32_t)val[i];
  if (head) comp_ptr[c] = i;                     // the first bucket of another shard closes this shard's last component
  if (mine) {
    old_of_new_b[i] = b;
    new_of_old_b[b] = i;
    comp_of_new_b[i] = c;
    nb_cnt[i] = (uint32_t)(bucket_ptr[b + 1] - bucket_ptr[b]);
    if (head) {
      comp_label[c] = (int32_t)(key[i] & ((1ull << bbits) - 1ull));
      atomicAdd(&counters[1], 1u);
      if (WORK_CAP - (uint32_t)((key[i] >> bbits) & WORK_CAP) >= heavy_reads) atomicAdd(&counters[2], 1u);   // heaviest first
    }
  } else {
    new_of_old_b[b] = -1;
    nb_cnt[i] = 0;
  }
  if (i == U - 1) { comp_ptr[n_comp] = U; nb_cnt[U] = 0; }
}

__global__ void shard_reads_kernel(const uint32_t *__restrict__ nb_ptr, uint32_t *__restrict__ counters) {
  counters[3] = nb_ptr[counters[0]];
}

// one thread per position x of the caller's bucket_reads: where does that read go?
__global__ void read_renumber_kernel(const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads, int32_t R, int32_t U,
                                     const int32_t *__restrict__ new_of_old_b, const uint32_t *__restrict__ nb_ptr,
                                     const int32_t *__restrict__ lexrank, const int32_t *__restrict__ bc_id, const int32_t *__restrict__ up_id,
                                     const int32_t *__restrict__ geno_ptr,
                                     int32_t *__restrict__ old_of_new_r, int32_t *__restrict__ n_bor, int32_t *__restrict__ n_lex,
                                     int32_t *__restrict__ n_bc, int32_t *__restrict__ n_up, uint32_t *__restrict__ n_gcnt) {
  int32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= R) return;
  int32_t lo = 0, hi = U;                       // last b with bucket_ptr[b] <= x
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (bucket_ptr[mid] <= x) lo = mid; else hi = mid;
  }
  int32_t nb = new_of_old_b[lo];
  if (nb < 0) return;                            // a bucket of another shard
  int32_t nr = (int32_t)nb_ptr[nb] + (x - bucket_ptr[lo]);
  int32_t r = bucket_reads[x];
  old_of_new_r[nr] = r;
  n_bor[nr] = nb;
  n_lex[nr] = lexrank[r];
  n_bc[nr] = bc_id[r];
  if (up_id) n_up[nr] = up_id[r];
  n_gcnt[nr] = (uint32_t)(geno_ptr[r + 1] - geno_ptr[r]);
}

__global__ void geno_gather_kernel(const int32_t *__restrict__ old_of_new_r, const int32_t *__restrict__ geno_ptr,
                                   const int32_t *__restrict__ geno_mut, const int32_t *__restrict__ geno_q, int32_t R,
                                   const uint32_t *__restrict__ n_gptr, int32_t *__restrict__ n_gmut, int32_t *__restrict__ n_gq,
                                   const uint32_t *__restrict__ counters) {
  int32_t nr = blockIdx.x * blockDim.x + threadIdx.x;
  if (nr >= R || (uint32_t)nr >= counters[3]) return;
  int32_t r = old_of_new_r[nr];
  int32_t src = geno_ptr[r], n = geno_ptr[r + 1] - src;
  uint32_t dst = n_gptr[nr];
  for (int32_t e = 0; e < n; e++) { n_gmut[dst + e] = geno_mut[src + e]; n_gq[dst + e] = geno_q[src + e]; }
}

__global__ void planes_gather_kernel(const int32_t *__restrict__ old_of_new_b, const ulonglong2 *__restrict__ planes,
                                     const int32_t *__restrict__ blen, int32_t U, ulonglong2 *__restrict__ n_planes,
                                     int32_t *__restrict__ n_blen, const uint32_t *__restrict__ counters) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U || (uint32_t)i >= counters[0]) return;
  n_planes[i] = planes[old_of_new_b[i]];
  n_blen[i] = blen[old_of_new_b[i]];
}

// the pairs whose two buckets belong to this shard, in their original order, in the new bucket numbering
__global__ void pair_flag_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ pr, int64_t n_pairs,
                                 const int32_t *__restrict__ new_of_old_b, uint32_t *__restrict__ flag) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pairs) flag[p] = (new_of_old_b[pq[p]] >= 0 && new_of_old_b[pr[p]] >= 0) ? 1u : 0u;
}
__global__ void pair_compact_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ pr, const int32_t *__restrict__ pd,
                                    const int32_t *__restrict__ ped, int64_t n_pairs, const int32_t *__restrict__ new_of_old_b,
                                    const uint32_t *__restrict__ pos, int32_t *__restrict__ mq, int32_t *__restrict__ mr,
                                    int32_t *__restrict__ md, int32_t *__restrict__ me) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  int32_t a = new_of_old_b[pq[p]], b = new_of_old_b[pr[p]];
  if (a < 0 || b < 0) return;
  uint32_t at = pos[p];
  mq[at] = a; mr[at] = b; md[at] = pd[p]; me[at] = ped[p];
}

__global__ void adj_keys_kernel(const int32_t *__restrict__ mq, const int32_t *__restrict__ mr, const int32_t *__restrict__ md,
                                int64_t n_pairs, int shift, uint64_t *__restrict__ key, uint32_t *__restrict__ val,
                                uint32_t *__restrict__ adj_cnt) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  uint32_t a = (uint32_t)mq[p], b = (uint32_t)mr[p];
  key[2 * p] = ((uint64_t)a << shift) | b;
  key[2 * p + 1] = ((uint64_t)b << shift) | a;
  val[2 * p] = val[2 * p + 1] = (uint32_t)md[p];
  atomicAdd(&adj_cnt[a], 1u);
  atomicAdd(&adj_cnt[b], 1u);
}

__global__ void adj_unpack_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, int64_t n, int shift,
                                  int32_t *__restrict__ nbr, int32_t *__restrict__ d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  nbr[i] = (int32_t)(key[i] & ((1ull << shift) - 1ull));
  d[i] = (int32_t)val[i];
}

__global__ void cpair_keys_kernel(const int32_t *__restrict__ mq, const int32_t *__restrict__ me, int64_t n_pairs,
                                  const int32_t *__restrict__ comp_of_new_b, int32_t n_comp, int32_t max_diff,
                                  uint64_t *__restrict__ key, uint32_t *__restrict__ val, uint32_t *__restrict__ cnt) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  int32_t ed = me[p];
  bool visit = ed >= 1 && ed <= max_diff;
  uint32_t c = visit ? (uint32_t)comp_of_new_b[mq[p]] : (uint32_t)n_comp;
  key[p] = ((uint64_t)c << 2) | (uint64_t)(visit ? ed : 0);
  val[p] = (uint32_t)p;
  if (visit) atomicAdd(&cnt[c], 1u);
}

__global__ void cpair_gather_kernel(const uint32_t *__restrict__ order, int64_t n, const int32_t *__restrict__ mq,
                                    const int32_t *__restrict__ mr, const int32_t *__restrict__ me,
                                    int32_t *__restrict__ cq, int32_t *__restrict__ cr, int32_t *__restrict__ ce) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t p = order[i];
  cq[i] = mq[p]; cr[i] = mr[p]; ce[i] = me[p];
}

// components are contiguous in the new numbering: sizes come straight from the prefix arrays; the call region of a
// component sits where its label puts it (label_nnz_kernel)
__global__ void comp_scratch_kernel(const int32_t *__restrict__ comp_ptr, const uint32_t *__restrict__ nb_ptr,
                                    const uint32_t *__restrict__ n_gptr, int32_t n_comp, const int32_t *__restrict__ comp_label,
                                    const uint32_t *__restrict__ lab_call_off, int32_t *__restrict__ slot_cap,
                                    int64_t *__restrict__ slot_sz, int64_t *__restrict__ list_sz, int64_t *__restrict__ call_off) {
  int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_comp) return;
  if (c == n_comp) { slot_sz[c] = list_sz[c] = 0; call_off[c] = 0; return; }
  uint32_t r_lo = nb_ptr[comp_ptr[c]], r_hi = nb_ptr[comp_ptr[c + 1]];
  unsigned long long reads = r_hi - r_lo, nnz = n_gptr[r_hi] - n_gptr[r_lo];
  unsigned long long need = 2ull * (reads > nnz ? reads : nnz);
  if (need < 2) need = 2;
  unsigned long long cap = 2;
  while (cap < need) cap <<= 1;
  slot_cap[c] = (int32_t)cap;
  slot_sz[c] = (int64_t)cap;
  list_sz[c] = (int64_t)(2ull * (reads + nnz) + 8);      // 4 counters, touched list, run list (merge_warp.cuh)
  call_off[c] = (int64_t)lab_call_off[comp_label[c]];
}

// Two launch shapes for the replay (merge_core.h / merge_warp.cuh):
//  * component_kernel_thread: one THREAD per component running the host-testable
//    process_component().  Divergent lanes serialise, so only every (32 / per_warp)-th lane is used.
//    Components are numbered heaviest first, so the long ones start first.
//  * component_kernel_warp: one WARP per component, uniform control flow, lanes share the order-independent
//    work (merge_warp.cuh), mutable state of small components in shared memory.  Instruction-fetch bound
//    (5.3 k SASS instructions, no_instruction is the top stall in profiles/r01f), so it loses to the thread
//    kernel on the 10-read components of cfg 2 and wins wherever components carry more work.
// merge_dev() picks by mean bucket size; PCB_REPLAY=thread|warp overrides.
constexpr int COMP_BLOCK = 128;

// Mutable working set of one small component in shared memory: cluster lists, union-find, scratch table,
// lists.  Every read-modify-write of the replay then costs a shared-memory round trip (~30 cycles) instead
// of an L2 one; at one component per warp this fits at full occupancy (8 blocks x 4 warps x 3.6 KB per SM).
#ifndef PCB_SM_READS
#define PCB_SM_READS 32
#endif
constexpr int SM_READS = PCB_SM_READS;
constexpr int SM_SLOTS = 2 * SM_READS;
static_assert((SM_SLOTS & (SM_SLOTS - 1)) == 0, "the scratch table size must be a power of two");
struct __align__(16) CompSmem {
  GenoSlot slots[SM_SLOTS];
  int64_t sl_key[SM_READS];
  int32_t lists[8 + 4 * SM_READS];
  int32_t rd_slot[SM_READS], rd_next[SM_READS], sl_parent[SM_READS], sl_head[SM_READS], sl_tail[SM_READS], sl_size[SM_READS];
};

__device__ __forceinline__ void replay_component_warp(const MergeCtx &c, int32_t comp, int lane, CompSmem *sm) {
  const int32_t r_lo = c.bucket_ptr[c.comp_ptr[comp]];
  const int32_t reads = c.bucket_ptr[c.comp_ptr[comp + 1]] - r_lo;
  const int32_t nnz = c.geno_ptr[r_lo + reads] - c.geno_ptr[r_lo];
  MergeState st = global_state(c, comp);
  if (reads <= SM_READS && nnz <= SM_READS) {
    for (int32_t i = lane; i < SM_SLOTS; i += 32) { GenoSlot z; z.key = 0; z.obs1 = z.obs2 = z.pad = 0; z.sum1 = z.sum2 = 0; sm->slots[i] = z; }
    for (int32_t i = lane; i < reads; i += 32) { sm->rd_slot[i] = -1; sm->rd_next[i] = -1; }
    __syncwarp();
    st.r0 = r_lo;
    st.rd_slot = sm->rd_slot; st.rd_next = sm->rd_next; st.sl_parent = sm->sl_parent; st.sl_head = sm->sl_head;
    st.sl_tail = sm->sl_tail; st.sl_size = sm->sl_size; st.sl_key = sm->sl_key;
    st.slots = sm->slots; st.slot_mask = SM_SLOTS - 1;
    st.lists = sm->lists; st.list_cap = 2 * SM_READS;
  }
  process_component_warp(c, st, comp, lane);        // one call site: the same code runs on either memory
}

__global__ void __launch_bounds__(COMP_BLOCK) component_kernel_warp(MergeCtx c, int32_t comp_lo, int32_t n_comp) {
  __shared__ CompSmem sm_all[COMP_BLOCK / 32];
  const int32_t comp = comp_lo + (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (comp >= n_comp) return;
  replay_component_warp(c, comp, threadIdx.x & 31, &sm_all[threadIdx.x >> 5]);
}

// the components the small-component kernel declined (list filled on the device, its length known only there):
// a fixed grid of warps walks the list
__global__ void __launch_bounds__(COMP_BLOCK) component_kernel_warp_list(MergeCtx c, const int32_t *__restrict__ list,
                                                                         const int32_t *__restrict__ n_list) {
  __shared__ CompSmem sm_all[COMP_BLOCK / 32];
  const int32_t n = *n_list;
  const int32_t n_warps = (int32_t)(gridDim.x * (blockDim.x >> 5));
  for (int32_t i = (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5); i < n; i += n_warps) {
    replay_component_warp(c, list[i], threadIdx.x & 31, &sm_all[threadIdx.x >> 5]);
    __syncwarp();
  }
}

// One warp per component of at most 32 reads, one lane per read (merge_small.cuh).  A component it is not
// eligible for (more than 32 distinct mutations, huge qualities) goes on the list for the generic kernel.
constexpr int SMALL_BLOCK = 128;
__global__ void __launch_bounds__(SMALL_BLOCK) component_kernel_small(MergeCtx c, int32_t comp_lo, int32_t n_comp,
                                                                      int32_t *__restrict__ declined, int32_t *__restrict__ n_declined) {
  __shared__ SmallSmem sm_all[SMALL_BLOCK / 32];
  const int32_t comp = comp_lo + (int32_t)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (comp >= n_comp) return;
  const int lane = threadIdx.x & 31;
  if (process_component_small(c, comp, lane, &sm_all[threadIdx.x >> 5]) != 0 && lane == 0) declined[atomicAdd(n_declined, 1)] = comp;
}

// Block size of the thread kernels: warps of one block finish at very different times, and a block keeps its
// registers until its slowest warp is done.
#ifndef PCB_THREAD_BLOCK
#define PCB_THREAD_BLOCK 128
#endif
constexpr int THREAD_BLOCK = PCB_THREAD_BLOCK;
template <int PHASE>
__global__ void __launch_bounds__(THREAD_BLOCK) component_kernel_thread(MergeCtx c, int32_t comp_lo, int32_t n_comp, int32_t shard_rank,
                                                                      int32_t shard_count, int32_t per_warp) {
  const int32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t stride = 32 / per_warp;
  if ((tid & 31) % stride != 0) return;
  const int32_t comp = comp_lo + (tid >> 5) * per_warp + (tid & 31) / stride;
  if (comp >= n_comp) return;
  if (comp % shard_count != shard_rank) return;
  const MergeState st = global_state(c, comp);
  if (PHASE == 1) component_seed(c, st, comp);
  if (PHASE == 2) component_main(c, st, comp);
  if (PHASE == 3) component_consensus(c, st, comp);
}

__global__ void init_outputs_kernel(int32_t R, int64_t nnz, int32_t *t_root, int32_t *t_pos, int32_t *t_size, int64_t *t_key,
                                    int32_t *t_bc, int32_t *t_up, int64_t *t_calloff, int32_t *t_calln, int32_t *call_mut) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) {
    t_root[i] = t_pos[i] = t_size[i] = t_bc[i] = t_up[i] = t_calln[i] = PCB_FILL;
    t_key[i] = INT64_MIN; t_calloff[i] = INT64_MIN;
  }
  if (i < nnz) call_mut[i] = PCB_FILL;
}

// new numbering -> the caller's numbering.  Rows of components another shard owns stay PCB_FILL.
__global__ void scatter_outputs_kernel(int32_t R, const int32_t *__restrict__ old_of_new_r, const int32_t *__restrict__ old_of_new_b,
                                       const int32_t *__restrict__ t_root, const int32_t *__restrict__ t_pos,
                                       const int32_t *__restrict__ t_size, const int64_t *__restrict__ t_key,
                                       const int32_t *__restrict__ t_bc, const int32_t *__restrict__ t_up,
                                       const int64_t *__restrict__ t_calloff, const int32_t *__restrict__ t_calln, MergeOut out) {
  int32_t nr = blockIdx.x * blockDim.x + threadIdx.x;
  if (nr >= R) return;                              // R = reads of this shard
  int32_t r = old_of_new_r[nr];
  int32_t root = t_root[nr];
  out.read_root[r] = root == PCB_FILL ? PCB_FILL : old_of_new_r[root];
  out.read_pos[r] = t_pos[nr];
  out.row_size[r] = t_size[nr];
  int64_t key = t_key[nr];
  if (key >= 0) key = ((int64_t)old_of_new_b[(int32_t)(key >> 32)] << 32) | (key & 0xFFFFFFFFll);   // creation order is by caller bucket
  out.row_key[r] = key;
  out.row_bc[r] = t_bc[nr];
  out.row_up[r] = t_up[nr];
  out.row_call_off[r] = t_calloff[nr];
  out.row_call_n[r] = t_calln[nr];
}

void merge_dev(pcb_ctx *ctx, const MergeIn &in, const MergeOut &out) {
  const int32_t R = in.R, U = in.U;
  const int64_t P = in.n_pairs;
  if (in.max_diff < 0 || in.max_diff > 3) throw Error(PCB_E_INPUT, "maxDiff must be between 0 and 3");
  if (!(in.min_jaccard > 0.0 && in.min_jaccard <= 1.0)) throw Error(PCB_E_INPUT, "minJaccard must be between 0 and 1");
  if (in.min_matches < 1) throw Error(PCB_E_INPUT, "minMatches must be at least 1");
  if (in.min_qual < 1) throw Error(PCB_E_INPUT, "minQual must be a positive integer");
  if (in.shard_count < 1 || in.shard_rank < 0 || in.shard_rank >= in.shard_count) throw Error(PCB_E_INPUT, "bad shard");
  if (R == 0 || U == 0) return;
  const int64_t nnz = in.nnz >= 0 ? in.nnz : (int64_t)read_scalar(ctx, in.geno_ptr + R);
  const int bbits = bits_for((uint64_t)(U > 1 ? U - 1 : 1));
  const bool on_demand = in.planes != nullptr && in.compute_k >= 0;

  // ---- components of the link graph; this shard's components first, heaviest first
  DevBuf<int32_t> label(ctx, U), old_of_new_b(ctx, U), new_of_old_b(ctx, U), comp_of_new_b(ctx, U), comp_ptr(ctx, (size_t)U + 2),
      comp_label(ctx, (size_t)U + 1);
  DevBuf<uint32_t> nb_ptr(ctx, (size_t)U + 1), lab_call_off(ctx, (size_t)U + 1), counters(ctx, 4);
  int32_t n_comp_all = 0;
  {
    DevBuf<int32_t> parent(ctx, U);
    DevBuf<uint32_t> work(ctx, U), rank(ctx, (size_t)U + 1);
    DevBuf<uint64_t> k0(ctx, U), k1(ctx, U);
    DevBuf<uint32_t> v0(ctx, U), v1(ctx, U);
    PCB_CUDA(cudaMemsetAsync(work.p, 0, (size_t)U * sizeof(uint32_t), ctx->stream));
    PCB_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(uint32_t), ctx->stream));
    PCB_CUDA(cudaMemsetAsync(lab_call_off.p, 0, ((size_t)U + 1) * sizeof(uint32_t), ctx->stream));
    PCB_LAUNCH(ctx, uf_init_kernel, grid_for(U, 256), 256, 0, parent.p, U);
    if (P > 0) PCB_LAUNCH(ctx, uf_union_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, in.pair_ed, P, in.max_diff, parent.p);
    PCB_LAUNCH(ctx, uf_label_kernel, grid_for(U, 256), 256, 0, parent.p, in.bucket_ptr, U, label.p, work.p);
    PCB_LAUNCH(ctx, label_nnz_kernel, grid_for(U, 256), 256, 0, in.bucket_ptr, in.bucket_reads, in.geno_ptr, label.p, U, lab_call_off.p);
    exclusive_scan_u32(ctx, lab_call_off.p, lab_call_off.p, (size_t)U + 1);
    PCB_LAUNCH(ctx, comp_keys_kernel, grid_for(U, 256), 256, 0, label.p, work.p, U, bbits, in.shard_rank, in.shard_count, k0.p, v0.p,
               counters.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, U, bbits + 11);
    uint64_t *sk = w ? k1.p : k0.p;
    uint32_t *sv = w ? v1.p : v0.p;
    PCB_LAUNCH(ctx, comp_flag_kernel, grid_for(U, 256), 256, 0, sk, U, rank.p);
    PCB_CUDA(cudaMemsetAsync(rank.p + U, 0, sizeof(uint32_t), ctx->stream));
    exclusive_scan_u32(ctx, rank.p, rank.p, (size_t)U + 1);              // one past the end: the total
    n_comp_all = (int32_t)read_scalar(ctx, rank.p + U);
    PCB_LAUNCH(ctx, comp_assign_kernel, grid_for(U, 256), 256, 0, sk, sv, rank.p, in.bucket_ptr, U, n_comp_all, old_of_new_b.p,
               new_of_old_b.p, comp_ptr.p, comp_of_new_b.p, nb_ptr.p, comp_label.p, bbits, (uint32_t)HEAVY_READS, counters.p);
    exclusive_scan_u32(ctx, nb_ptr.p, nb_ptr.p, (size_t)U + 1);
    PCB_LAUNCH(ctx, shard_reads_kernel, 1, 1, 0, nb_ptr.p, counters.p);
  }
  ctx->stats[PCB_STAT_COMPONENTS] = n_comp_all;

  // ---- this shard's pairs, compacted (original order kept) and renumbered
  DevBuf<uint32_t> ppos(ctx, (size_t)P + 1);
  PCB_CUDA(cudaMemsetAsync(ppos.p, 0, ((size_t)P + 1) * sizeof(uint32_t), ctx->stream));
  if (P > 0) {
    PCB_LAUNCH(ctx, pair_flag_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, P, new_of_old_b.p, ppos.p);
    exclusive_scan_u32(ctx, ppos.p, ppos.p, (size_t)P + 1);
  }
  uint32_t h_counters[4], h_pm = 0;                                     // one round trip for all the shard sizes
  PCB_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(&h_pm, ppos.p + P, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  const int32_t Um = (int32_t)h_counters[0], n_comp = (int32_t)h_counters[1], Rm = (int32_t)h_counters[3];
  const uint32_t h_n_heavy = h_counters[2];
  const int64_t Pm = (int64_t)h_pm;
  (void)Um;
  DevBuf<int32_t> mq(ctx, (size_t)Pm), mr(ctx, (size_t)Pm), md(ctx, (size_t)Pm), me(ctx, (size_t)Pm);
  if (Pm > 0)
    PCB_LAUNCH(ctx, pair_compact_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, in.pair_d, in.pair_ed, P, new_of_old_b.p,
               ppos.p, mq.p, mr.p, md.p, me.p);

  // ---- reads and genotypes of this shard gathered into component order
  DevBuf<int32_t> old_of_new_r(ctx, R), n_bor(ctx, R), n_lex(ctx, R), n_bc(ctx, R), n_up(ctx, in.up_id ? R : 1);
  DevBuf<uint32_t> n_gptr(ctx, (size_t)R + 1);
  DevBuf<int32_t> n_gmut(ctx, (size_t)nnz), n_gq(ctx, (size_t)nnz);
  PCB_CUDA(cudaMemsetAsync(n_gptr.p, 0, ((size_t)R + 1) * sizeof(uint32_t), ctx->stream));
  PCB_LAUNCH(ctx, read_renumber_kernel, grid_for(R, 256), 256, 0, in.bucket_ptr, in.bucket_reads, R, U, new_of_old_b.p, nb_ptr.p,
             in.lexrank, in.bc_id, in.up_id, in.geno_ptr, old_of_new_r.p, n_bor.p, n_lex.p, n_bc.p, n_up.p, n_gptr.p);
  exclusive_scan_u32(ctx, n_gptr.p, n_gptr.p, (size_t)Rm + 1);
  if (Rm > 0)
    PCB_LAUNCH(ctx, geno_gather_kernel, grid_for(Rm, 256), 256, 0, old_of_new_r.p, in.geno_ptr, in.geno_mut, in.geno_q, Rm, n_gptr.p,
               n_gmut.p, n_gq.p, counters.p);
  DevBuf<ulonglong2> n_planes(ctx, on_demand ? U : 1);
  DevBuf<int32_t> n_blen(ctx, on_demand ? U : 1);
  if (on_demand)
    PCB_LAUNCH(ctx, planes_gather_kernel, grid_for(U, 256), 256, 0, old_of_new_b.p, in.planes, in.blen, U, n_planes.p, n_blen.p,
               counters.p);

  // ---- adjacency (new bucket ids)
  DevBuf<int32_t> adj_ptr(ctx, (size_t)U + 1), adj_nbr(ctx, (size_t)(2 * Pm)), adj_d(ctx, (size_t)(2 * Pm));
  {
    DevBuf<uint32_t> cnt(ctx, (size_t)U + 1);
    PCB_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)U + 1) * sizeof(uint32_t), ctx->stream));
    if (Pm > 0) {
      DevBuf<uint64_t> k0(ctx, 2 * Pm), k1(ctx, 2 * Pm);
      DevBuf<uint32_t> v0(ctx, 2 * Pm), v1(ctx, 2 * Pm);
      PCB_LAUNCH(ctx, adj_keys_kernel, grid_for(Pm, 256), 256, 0, mq.p, mr.p, md.p, Pm, bbits, k0.p, v0.p, cnt.p);
      int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, 2 * Pm, 2 * bbits);
      PCB_LAUNCH(ctx, adj_unpack_kernel, grid_for(2 * Pm, 256), 256, 0, w ? k1.p : k0.p, w ? v1.p : v0.p, 2 * Pm, bbits,
                 adj_nbr.p, adj_d.p);
    }
    exclusive_scan_u32(ctx, cnt.p, cnt.p, (size_t)U + 1);
    PCB_CUDA(cudaMemcpyAsync(adj_ptr.p, cnt.p, ((size_t)U + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }

  // ---- visiting pairs per component (stable: the caller's visiting order survives inside a component)
  DevBuf<int32_t> cpair_ptr(ctx, (size_t)n_comp + 2), cq(ctx, (size_t)Pm), cr(ctx, (size_t)Pm), ce(ctx, (size_t)Pm);
  {
    DevBuf<uint32_t> cnt(ctx, (size_t)n_comp + 2);
    PCB_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)n_comp + 2) * sizeof(uint32_t), ctx->stream));
    if (Pm > 0) {
      DevBuf<uint64_t> k0(ctx, Pm), k1(ctx, Pm);
      DevBuf<uint32_t> v0(ctx, Pm), v1(ctx, Pm);
      PCB_LAUNCH(ctx, cpair_keys_kernel, grid_for(Pm, 256), 256, 0, mq.p, me.p, Pm, comp_of_new_b.p, n_comp, in.max_diff, k0.p, v0.p,
                 cnt.p);
      int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, Pm, bits_for((uint64_t)n_comp) + 2);
      PCB_LAUNCH(ctx, cpair_gather_kernel, grid_for(Pm, 256), 256, 0, w ? v1.p : v0.p, Pm, mq.p, mr.p, me.p, cq.p, cr.p, ce.p);
    }
    exclusive_scan_u32(ctx, cnt.p, cnt.p, (size_t)n_comp + 2);
    PCB_CUDA(cudaMemcpyAsync(cpair_ptr.p, cnt.p, ((size_t)n_comp + 2) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }

  // ---- scratch
  DevBuf<int32_t> slot_cap(ctx, (size_t)n_comp + 1);
  DevBuf<int64_t> slot_off(ctx, (size_t)n_comp + 1), list_off(ctx, (size_t)n_comp + 1), call_off(ctx, (size_t)n_comp + 1);
  PCB_LAUNCH(ctx, comp_scratch_kernel, grid_for((size_t)n_comp + 1, 256), 256, 0, comp_ptr.p, nb_ptr.p, n_gptr.p, n_comp, comp_label.p,
             lab_call_off.p, slot_cap.p, slot_off.p, list_off.p, call_off.p);
  exclusive_scan_i64(ctx, slot_off.p, slot_off.p, (size_t)n_comp + 1);
  exclusive_scan_i64(ctx, list_off.p, list_off.p, (size_t)n_comp + 1);
  int64_t n_slots = 0, n_list = 0;
  PCB_CUDA(cudaMemcpyAsync(&n_slots, slot_off.p + n_comp, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(&n_list, list_off.p + n_comp, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  DevBuf<GenoSlot> slots(ctx, (size_t)n_slots);
  DevBuf<int32_t> lists(ctx, (size_t)n_list + 1);
  PCB_CUDA(cudaMemsetAsync(slots.p, 0, (size_t)n_slots * sizeof(GenoSlot), ctx->stream));     // all-zero = empty

  // ---- cluster state and outputs in the new numbering
  DevBuf<int32_t> rd_slot(ctx, R), rd_next(ctx, R), sl_parent(ctx, R), sl_head(ctx, R), sl_tail(ctx, R), sl_size(ctx, R), status(ctx, 4);
  DevBuf<int64_t> sl_key(ctx, R);
  DevBuf<int32_t> t_root(ctx, R), t_pos(ctx, R), t_size(ctx, R), t_bc(ctx, R), t_up(ctx, R), t_calln(ctx, R);
  DevBuf<int64_t> t_key(ctx, R), t_calloff(ctx, R);
  PCB_CUDA(cudaMemsetAsync(rd_slot.p, 0xFF, (size_t)R * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(rd_next.p, 0xFF, (size_t)R * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(status.p, 0, 4 * sizeof(int32_t), ctx->stream));
  // everything starts as "not mine": the per-row fields of non-representative reads (new numbering, this shard's reads),
  // the call array, and -- when the library is sharded -- the caller's arrays for the reads other shards own
  PCB_LAUNCH(ctx, init_outputs_kernel, grid_for((size_t)(Rm > nnz ? Rm : nnz) + 1, 256), 256, 0, Rm, nnz, t_root.p, t_pos.p, t_size.p,
             t_key.p, t_bc.p, t_up.p, t_calloff.p, t_calln.p, out.call_mut);
  if (in.shard_count > 1)
    PCB_LAUNCH(ctx, init_outputs_kernel, grid_for((size_t)R, 256), 256, 0, R, (int64_t)0, out.read_root, out.read_pos, out.row_size,
               out.row_key, out.row_bc, out.row_up, out.row_call_off, out.row_call_n, out.call_mut);

  DevBuf<int32_t> declined(ctx, (size_t)n_comp + 1), n_declined(ctx, 1);
  MergeCtx c;
  memset(&c, 0, sizeof(c));
  c.R = R; c.U = U;
  c.bucket_ptr = (const int32_t *)nb_ptr.p; c.bucket_reads = nullptr; c.bucket_of_read = n_bor.p;
  c.lexrank = n_lex.p; c.bc_id = n_bc.p; c.up_id = in.up_id ? n_up.p : nullptr;
  c.geno_ptr = (const int32_t *)n_gptr.p; c.geno_mut = n_gmut.p; c.geno_q = n_gq.p; c.mut_rank = in.mut_rank;
  c.adj_ptr = adj_ptr.p; c.adj_nbr = adj_nbr.p; c.adj_d = adj_d.p;
  c.planes = on_demand ? (const uint64_t *)n_planes.p : nullptr; c.blen = on_demand ? n_blen.p : nullptr;
  c.compute_k = on_demand ? in.compute_k : -1; c.wide = in.wide;
  c.comp_ptr = comp_ptr.p; c.comp_buckets = nullptr;
  c.cpair_ptr = cpair_ptr.p; c.cpair_q = cq.p; c.cpair_r = cr.p; c.cpair_ed = ce.p;
  c.scr_slot_off = slot_off.p; c.scr_slot_cap = slot_cap.p; c.scr_list_off = list_off.p; c.call_off = call_off.p;
  c.slots = slots.p; c.lists = lists.p;
  c.rd_slot = rd_slot.p; c.rd_next = rd_next.p; c.sl_parent = sl_parent.p; c.sl_head = sl_head.p;
  c.sl_tail = sl_tail.p; c.sl_size = sl_size.p; c.sl_key = sl_key.p;
  c.read_root = t_root.p; c.read_pos = t_pos.p; c.row_size = t_size.p; c.row_key = t_key.p;
  c.row_bc = t_bc.p; c.row_up = t_up.p; c.row_call_off = t_calloff.p; c.row_call_n = t_calln.p;
  c.call_mut = out.call_mut; c.status = status.p;
  c.prm.maxDiff = in.max_diff; c.prm.minQual = in.min_qual; c.prm.minMatches = in.min_matches; c.prm.minJaccard = in.min_jaccard;
  stage_end(ctx, PCB_STAT_T_PREP_NS);
  // Launch shapes (profiles/r01_replay_shapes.md, r02_replay.md).  Default: components of at most 32 reads -- all but a
  // handful on every BASELINE config -- go to the one-lane-per-read kernel; the bigger ones (numbered first) and
  // whatever that kernel declines go to the generic warp kernel.  ctx->replay_shape forces the round-1 shapes
  // (one thread per component / the generic warp kernel for everything), which the parity tests keep exercising.
  int32_t n_warp = (int32_t)h_n_heavy;
  if (ctx->replay_shape == PCB_REPLAY_WARP) n_warp = n_comp;
  if (n_warp > 0)
    PCB_LAUNCH(ctx, component_kernel_warp, grid_for((size_t)n_warp * 32, COMP_BLOCK), COMP_BLOCK, 0, c, 0, n_warp);
  if (n_warp < n_comp && ctx->replay_shape == PCB_REPLAY_THREAD) {
    const int per_warp = 4;               // best of a 1..32 sweep (divergent lanes serialise)
    size_t n_threads = ((size_t)(n_comp - n_warp) + per_warp - 1) / per_warp * 32;
    // three launches, one per phase: each kernel's code fits the instruction cache and its threads share a loop nest
    PCB_LAUNCH(ctx, component_kernel_thread<1>, grid_for(n_threads, THREAD_BLOCK), THREAD_BLOCK, 0, c, n_warp, n_comp, 0, 1, per_warp);
    PCB_LAUNCH(ctx, component_kernel_thread<2>, grid_for(n_threads, THREAD_BLOCK), THREAD_BLOCK, 0, c, n_warp, n_comp, 0, 1, per_warp);
    PCB_LAUNCH(ctx, component_kernel_thread<3>, grid_for(n_threads, THREAD_BLOCK), THREAD_BLOCK, 0, c, n_warp, n_comp, 0, 1, per_warp);
  } else if (n_warp < n_comp) {
    PCB_CUDA(cudaMemsetAsync(n_declined.p, 0, sizeof(int32_t), ctx->stream));
    PCB_LAUNCH(ctx, component_kernel_small, grid_for((size_t)(n_comp - n_warp) * 32, SMALL_BLOCK), SMALL_BLOCK, 0, c, n_warp, n_comp,
               declined.p, n_declined.p);
    PCB_LAUNCH(ctx, component_kernel_warp_list, (unsigned)ctx->sm_count * 2u, COMP_BLOCK, 0, c, declined.p, n_declined.p);
  }
  stage_end(ctx, PCB_STAT_T_REPLAY_NS);
  if (Rm > 0)
    PCB_LAUNCH(ctx, scatter_outputs_kernel, grid_for(Rm, 256), 256, 0, Rm, old_of_new_r.p, old_of_new_b.p, t_root.p, t_pos.p, t_size.p,
               t_key.p, t_bc.p, t_up.p, t_calloff.p, t_calln.p, out);
  stage_end(ctx, PCB_STAT_T_SCATTER_NS);

  int32_t h_status[4];
  PCB_CUDA(cudaMemcpyAsync(h_status, status.p, sizeof(h_status), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats[PCB_STAT_LINKS] = h_status[2];
  ctx->stats[PCB_STAT_TESTS] = h_status[3];
  if (h_status[0] == 3)
    throw Error(PCB_E_UPTAG, "a clustered read (row of read " + std::to_string(h_status[1]) +
                                 ") has no uptag record: the reference raises KeyError at pacybara_cluster.py:519");
}

}  // namespace pcb
