// merge.cu -- cluster merge on the device (rows a4..a13): preprocessing kernels that turn the
// bucket-pair list into per-component work, then one thread per connected component replaying
// the reference's sequential algorithm (merge_core.h).
//
//   adjacency     symmetric (a, nbr, d) sorted by (a, nbr): the lookup table of :236,:266
//   components    lock-free union-find over the pairs a pass visits (1 <= ed <= maxDiff);
//                 label = smallest bucket of the component
//   grouping      buckets and visiting pairs gathered per component with the stable radix sort,
//                 which keeps the reference's visiting order inside every component
//   replay        process_component(): seed step, greedy merge, consensus
// Everything except the replay is HBM-bound streaming; the replay is latency-bound pointer work
// whose parallelism is the number of components (hundreds of thousands at cfg 2).
#include <cstring>

#include "common.cuh"
#include "merge_core.h"

namespace pcb {

__global__ void bucket_of_read_kernel(const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads,
                                      int32_t R, int32_t U, int32_t *__restrict__ bucket_of_read) {
  int32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= R) return;
  int32_t lo = 0, hi = U;                       // last b with bucket_ptr[b] <= x
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (bucket_ptr[mid] <= x) lo = mid; else hi = mid;
  }
  bucket_of_read[bucket_reads[x]] = lo;
}

__global__ void adj_keys_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ pr, const int32_t *__restrict__ pd,
                                int64_t n_pairs, int shift, uint64_t *__restrict__ key, uint32_t *__restrict__ val,
                                uint32_t *__restrict__ adj_cnt) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  uint32_t a = (uint32_t)pq[p], b = (uint32_t)pr[p];
  key[2 * p] = ((uint64_t)a << shift) | b;
  key[2 * p + 1] = ((uint64_t)b << shift) | a;
  val[2 * p] = val[2 * p + 1] = (uint32_t)pd[p];
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
                                int64_t n_pairs, int32_t *parent) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || ped[p] < 1) return;
  int32_t a = pq[p], b = pr[p];
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    int32_t hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(&parent[hi], hi, lo) == hi) return;       // roots only ever point to a smaller root
  }
}

__global__ void uf_label_kernel(int32_t *parent, int32_t U, uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  key[b] = (uint64_t)(uint32_t)uf_find(parent, b);
  val[b] = (uint32_t)b;
}

__global__ void comp_flag_kernel(const uint64_t *__restrict__ key, int32_t U, uint32_t *__restrict__ flag) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < U) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// rank = exclusive scan of the flags: component index of sorted position i is rank[i] + flag - 1
__global__ void comp_assign_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, const uint32_t *__restrict__ rank,
                                   int32_t U, int32_t n_comp, int32_t *__restrict__ comp_buckets, int32_t *__restrict__ comp_ptr,
                                   int32_t *__restrict__ comp_of_bucket) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  bool head = i == 0 || key[i] != key[i - 1];
  int32_t c = (int32_t)rank[i] + (head ? 1 : 0) - 1;
  int32_t b = (int32_t)val[i];
  comp_buckets[i] = b;
  comp_of_bucket[b] = c;
  if (head) comp_ptr[c] = i;
  if (i == U - 1) comp_ptr[n_comp] = U;
}

__global__ void cpair_keys_kernel(const int32_t *__restrict__ pq, const int32_t *__restrict__ ped, int64_t n_pairs,
                                  const int32_t *__restrict__ comp_of_bucket, int32_t n_comp, int32_t max_diff,
                                  uint64_t *__restrict__ key, uint32_t *__restrict__ val, uint32_t *__restrict__ cnt) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  int32_t ed = ped[p];
  bool visit = ed >= 1 && ed <= max_diff;
  uint32_t c = visit ? (uint32_t)comp_of_bucket[pq[p]] : (uint32_t)n_comp;
  key[p] = ((uint64_t)c << 2) | (uint64_t)(visit ? ed : 0);
  val[p] = (uint32_t)p;
  if (visit) atomicAdd(&cnt[c], 1u);
}

__global__ void cpair_gather_kernel(const uint32_t *__restrict__ order, int64_t n, const int32_t *__restrict__ pq,
                                    const int32_t *__restrict__ pr, const int32_t *__restrict__ ped,
                                    int32_t *__restrict__ cq, int32_t *__restrict__ cr, int32_t *__restrict__ ce) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t p = order[i];
  cq[i] = pq[p]; cr[i] = pr[p]; ce[i] = ped[p];
}

__global__ void comp_load_kernel(const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads,
                                 const int32_t *__restrict__ geno_ptr, const int32_t *__restrict__ comp_of_bucket, int32_t U,
                                 unsigned long long *__restrict__ comp_reads, unsigned long long *__restrict__ comp_nnz) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  unsigned long long nnz = 0;
  for (int32_t x = bucket_ptr[b]; x < bucket_ptr[b + 1]; x++) {
    int32_t r = bucket_reads[x];
    nnz += (unsigned long long)(geno_ptr[r + 1] - geno_ptr[r]);
  }
  int32_t c = comp_of_bucket[b];
  atomicAdd(&comp_reads[c], (unsigned long long)(bucket_ptr[b + 1] - bucket_ptr[b]));
  atomicAdd(&comp_nnz[c], nnz);
}

__global__ void comp_scratch_kernel(const unsigned long long *__restrict__ comp_reads, const unsigned long long *__restrict__ comp_nnz,
                                    int32_t n_comp, int32_t *__restrict__ slot_cap, int64_t *__restrict__ slot_sz,
                                    int64_t *__restrict__ list_sz, int64_t *__restrict__ call_sz) {
  int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_comp) return;
  if (c == n_comp) { slot_sz[c] = list_sz[c] = call_sz[c] = 0; return; }
  unsigned long long reads = comp_reads[c], nnz = comp_nnz[c];
  unsigned long long need = 2ull * (reads > nnz ? reads : nnz);
  if (need < 2) need = 2;
  unsigned long long cap = 2;
  while (cap < need) cap <<= 1;
  slot_cap[c] = (int32_t)cap;
  slot_sz[c] = (int64_t)cap;
  list_sz[c] = (int64_t)(2ull * (reads + nnz));
  call_sz[c] = (int64_t)nnz;
}

__global__ void __launch_bounds__(64) component_kernel(MergeCtx c, int32_t n_comp, int32_t shard_rank, int32_t shard_count) {
  int32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp >= n_comp) return;
  if (comp % shard_count != shard_rank) return;
  process_component(c, comp);
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
  const int64_t nnz = (int64_t)read_scalar(ctx, in.geno_ptr + R);

  // outputs start as "not mine"
  fill_i32(ctx, out.read_root, R, PCB_FILL); fill_i32(ctx, out.read_pos, R, PCB_FILL);
  fill_i32(ctx, out.row_size, R, PCB_FILL); fill_i32(ctx, out.row_bc, R, PCB_FILL);
  fill_i32(ctx, out.row_up, R, PCB_FILL); fill_i32(ctx, out.row_call_n, R, PCB_FILL);
  fill_i64(ctx, out.row_key, R, INT64_MIN); fill_i64(ctx, out.row_call_off, R, INT64_MIN);
  fill_i32(ctx, out.call_mut, (size_t)nnz, PCB_FILL);

  DevBuf<int32_t> bucket_of_read(ctx, R);
  PCB_LAUNCH(ctx, bucket_of_read_kernel, grid_for(R, 256), 256, 0, in.bucket_ptr, in.bucket_reads, R, U, bucket_of_read.p);

  const int bbits = bits_for((uint64_t)(U > 1 ? U - 1 : 1));
  // ---- adjacency
  DevBuf<int32_t> adj_ptr(ctx, (size_t)U + 1), adj_nbr(ctx, (size_t)(2 * P)), adj_d(ctx, (size_t)(2 * P));
  {
    DevBuf<uint32_t> cnt(ctx, (size_t)U + 1);
    PCB_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)U + 1) * sizeof(uint32_t), ctx->stream));
    if (P > 0) {
      DevBuf<uint64_t> k0(ctx, 2 * P), k1(ctx, 2 * P);
      DevBuf<uint32_t> v0(ctx, 2 * P), v1(ctx, 2 * P);
      PCB_LAUNCH(ctx, adj_keys_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, in.pair_d, P, bbits, k0.p, v0.p, cnt.p);
      int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, 2 * P, 2 * bbits);
      PCB_LAUNCH(ctx, adj_unpack_kernel, grid_for(2 * P, 256), 256, 0, w ? k1.p : k0.p, w ? v1.p : v0.p, 2 * P, bbits,
                 adj_nbr.p, adj_d.p);
    }
    exclusive_scan_u32(ctx, cnt.p, cnt.p, (size_t)U + 1);
    PCB_CUDA(cudaMemcpyAsync(adj_ptr.p, cnt.p, ((size_t)U + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }

  // ---- components
  DevBuf<int32_t> parent(ctx, U), comp_buckets(ctx, U), comp_of_bucket(ctx, U), comp_ptr(ctx, (size_t)U + 1);
  int32_t n_comp = 0;
  {
    PCB_LAUNCH(ctx, uf_init_kernel, grid_for(U, 256), 256, 0, parent.p, U);
    if (P > 0) PCB_LAUNCH(ctx, uf_union_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_r, in.pair_ed, P, parent.p);
    DevBuf<uint64_t> k0(ctx, U), k1(ctx, U);
    DevBuf<uint32_t> v0(ctx, U), v1(ctx, U), rank(ctx, U);
    PCB_LAUNCH(ctx, uf_label_kernel, grid_for(U, 256), 256, 0, parent.p, U, k0.p, v0.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, U, bbits);
    uint64_t *sk = w ? k1.p : k0.p;
    uint32_t *sv = w ? v1.p : v0.p;
    PCB_LAUNCH(ctx, comp_flag_kernel, grid_for(U, 256), 256, 0, sk, U, rank.p);
    uint32_t last_flag = read_scalar(ctx, rank.p + (U - 1));
    exclusive_scan_u32(ctx, rank.p, rank.p, U);
    n_comp = (int32_t)(read_scalar(ctx, rank.p + (U - 1)) + last_flag);
    PCB_LAUNCH(ctx, comp_assign_kernel, grid_for(U, 256), 256, 0, sk, sv, rank.p, U, n_comp, comp_buckets.p, comp_ptr.p,
               comp_of_bucket.p);
  }
  ctx->stats[PCB_STAT_COMPONENTS] = n_comp;

  // ---- visiting pairs per component
  DevBuf<int32_t> cpair_ptr(ctx, (size_t)n_comp + 2), cq(ctx, (size_t)P), cr(ctx, (size_t)P), ce(ctx, (size_t)P);
  {
    DevBuf<uint32_t> cnt(ctx, (size_t)n_comp + 2);
    PCB_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)n_comp + 2) * sizeof(uint32_t), ctx->stream));
    if (P > 0) {
      DevBuf<uint64_t> k0(ctx, P), k1(ctx, P);
      DevBuf<uint32_t> v0(ctx, P), v1(ctx, P);
      PCB_LAUNCH(ctx, cpair_keys_kernel, grid_for(P, 256), 256, 0, in.pair_q, in.pair_ed, P, comp_of_bucket.p, n_comp,
                 in.max_diff, k0.p, v0.p, cnt.p);
      int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, P, bits_for((uint64_t)n_comp) + 2);
      PCB_LAUNCH(ctx, cpair_gather_kernel, grid_for(P, 256), 256, 0, w ? v1.p : v0.p, P, in.pair_q, in.pair_r, in.pair_ed,
                 cq.p, cr.p, ce.p);
    }
    exclusive_scan_u32(ctx, cnt.p, cnt.p, (size_t)n_comp + 2);
    PCB_CUDA(cudaMemcpyAsync(cpair_ptr.p, cnt.p, ((size_t)n_comp + 2) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }

  // ---- scratch
  DevBuf<int32_t> slot_cap(ctx, (size_t)n_comp + 1);
  DevBuf<int64_t> slot_off(ctx, (size_t)n_comp + 1), list_off(ctx, (size_t)n_comp + 1), call_off(ctx, (size_t)n_comp + 1);
  {
    DevBuf<unsigned long long> comp_reads(ctx, (size_t)n_comp + 1), comp_nnz(ctx, (size_t)n_comp + 1);
    PCB_CUDA(cudaMemsetAsync(comp_reads.p, 0, ((size_t)n_comp + 1) * sizeof(unsigned long long), ctx->stream));
    PCB_CUDA(cudaMemsetAsync(comp_nnz.p, 0, ((size_t)n_comp + 1) * sizeof(unsigned long long), ctx->stream));
    PCB_LAUNCH(ctx, comp_load_kernel, grid_for(U, 256), 256, 0, in.bucket_ptr, in.bucket_reads, in.geno_ptr, comp_of_bucket.p, U,
               comp_reads.p, comp_nnz.p);
    PCB_LAUNCH(ctx, comp_scratch_kernel, grid_for((size_t)n_comp + 1, 256), 256, 0, comp_reads.p, comp_nnz.p, n_comp, slot_cap.p,
               slot_off.p, list_off.p, call_off.p);
    exclusive_scan_i64(ctx, slot_off.p, slot_off.p, (size_t)n_comp + 1);
    exclusive_scan_i64(ctx, list_off.p, list_off.p, (size_t)n_comp + 1);
    exclusive_scan_i64(ctx, call_off.p, call_off.p, (size_t)n_comp + 1);
  }
  int64_t n_slots = read_scalar(ctx, slot_off.p + n_comp);
  int64_t n_list = read_scalar(ctx, list_off.p + n_comp);
  DevBuf<GenoSlot> slots(ctx, (size_t)n_slots);
  DevBuf<int32_t> lists(ctx, (size_t)n_list + 1);
  PCB_CUDA(cudaMemsetAsync(slots.p, 0xFF, (size_t)n_slots * sizeof(GenoSlot), ctx->stream));

  // ---- cluster state
  DevBuf<int32_t> rd_slot(ctx, R), rd_next(ctx, R), sl_parent(ctx, R), sl_head(ctx, R), sl_tail(ctx, R), sl_size(ctx, R), status(ctx, 4);
  DevBuf<int64_t> sl_key(ctx, R);
  PCB_CUDA(cudaMemsetAsync(rd_slot.p, 0xFF, (size_t)R * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(rd_next.p, 0xFF, (size_t)R * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(status.p, 0, 4 * sizeof(int32_t), ctx->stream));

  MergeCtx c;
  memset(&c, 0, sizeof(c));
  c.R = R; c.U = U;
  c.bucket_ptr = in.bucket_ptr; c.bucket_reads = in.bucket_reads; c.bucket_of_read = bucket_of_read.p;
  c.lexrank = in.lexrank; c.bc_id = in.bc_id; c.up_id = in.up_id;
  c.geno_ptr = in.geno_ptr; c.geno_mut = in.geno_mut; c.geno_q = in.geno_q; c.mut_rank = in.mut_rank;
  c.adj_ptr = adj_ptr.p; c.adj_nbr = adj_nbr.p; c.adj_d = adj_d.p;
  c.planes = (const uint64_t *)in.planes; c.blen = in.blen;
  c.compute_k = in.planes ? in.compute_k : -1; c.wide = in.wide;
  c.comp_ptr = comp_ptr.p; c.comp_buckets = comp_buckets.p;
  c.cpair_ptr = cpair_ptr.p; c.cpair_q = cq.p; c.cpair_r = cr.p; c.cpair_ed = ce.p;
  c.scr_slot_off = slot_off.p; c.scr_slot_cap = slot_cap.p; c.scr_list_off = list_off.p; c.call_off = call_off.p;
  c.slots = slots.p; c.lists = lists.p;
  c.rd_slot = rd_slot.p; c.rd_next = rd_next.p; c.sl_parent = sl_parent.p; c.sl_head = sl_head.p;
  c.sl_tail = sl_tail.p; c.sl_size = sl_size.p; c.sl_key = sl_key.p;
  c.read_root = out.read_root; c.read_pos = out.read_pos; c.row_size = out.row_size; c.row_key = out.row_key;
  c.row_bc = out.row_bc; c.row_up = out.row_up; c.row_call_off = out.row_call_off; c.row_call_n = out.row_call_n;
  c.call_mut = out.call_mut; c.status = status.p;
  c.prm.maxDiff = in.max_diff; c.prm.minQual = in.min_qual; c.prm.minMatches = in.min_matches; c.prm.minJaccard = in.min_jaccard;
  PCB_LAUNCH(ctx, component_kernel, grid_for(n_comp, 64), 64, 0, c, n_comp, in.shard_rank, in.shard_count);

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
