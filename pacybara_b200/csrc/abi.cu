// abi.cu -- the extern "C" boundary declared in include/pacybara_b200.h.
// Host pointers are staged through stream-ordered device buffers; device pointers are used as
// they are.  Every entry point converts exceptions into a PCB_E_* code plus pcb_last_error().
#include <cstring>

#include "common.cuh"

using namespace pcb;

static thread_local std::string g_create_error;

namespace {

template <typename T>
struct InArr {
  const T *p = nullptr;
  DevBuf<T> own;
  const T *src_ = nullptr;
  size_t n_ = 0;
  // late = true: allocate now (on the main stream) but leave the copy to copy_late()
  InArr(pcb_ctx *ctx, int mem, const T *src, size_t n, bool late = false) {
    if (!src) return;
    if (mem == PCB_MEM_DEVICE) { p = src; return; }
    own.alloc(ctx, n);
    p = own.p;
    if (late) { src_ = src; n_ = n; return; }
    if (n) PCB_CUDA(cudaMemcpyAsync(own.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  }
  void copy_late(cudaStream_t s) {
    if (src_ && n_) PCB_CUDA(cudaMemcpyAsync(own.p, src_, n_ * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

template <typename T>
struct OutArr {
  T *p = nullptr;
  T *host = nullptr;
  DevBuf<T> own;
  pcb_ctx *ctx;
  OutArr(pcb_ctx *c, int mem, T *dst, size_t n) : ctx(c) {
    if (!dst) return;
    if (mem == PCB_MEM_DEVICE) { p = dst; return; }
    own.alloc(c, n);
    p = own.p; host = dst;
  }
  void commit(size_t n) {
    if (host && n) PCB_CUDA(cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  }
};

// After a call that used the workspace arena: drop everything; if the call needed more than the arena holds, make it
// that large (plus a margin) for the next one.  The stream has been synchronised: nothing is in flight.
void arena_end(pcb_ctx *ctx) {
  ctx->arena_on = false;
  ctx->arena_off = 0;
  if (ctx->arena_need > ctx->arena_cap) {
    const size_t want = ctx->arena_need + ctx->arena_need / 8 + (1 << 20);
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr; ctx->arena_cap = 0;
    cudaMemPool_t pool;                                         // what the pool served this call is the arena's job from now on
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    if (cudaMalloc((void **)&ctx->arena, want) == cudaSuccess) ctx->arena_cap = want;
    else { ctx->arena = nullptr; cudaGetLastError(); }          // no room for it: the stream-ordered pool keeps serving
  }
}

// arena: the call's temporaries come from the context's workspace arena (not for calls that leave device objects behind)
template <typename F>
int guarded(pcb_ctx *ctx, F f, bool arena = true) {
  if (!ctx) return PCB_E_STATE;
  try {
    PCB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena_on = arena; ctx->arena_off = 0; ctx->arena_need = 0;
    f();
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
    PCB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    arena_end(ctx);
    ctx->err.clear();
    return PCB_OK;
  } catch (const Error &e) {
    ctx->err = e.what();
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    cudaGetLastError();
    arena_end(ctx);
    return e.code;
  } catch (const std::exception &e) {
    ctx->err = e.what();
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    arena_end(ctx);
    return PCB_E_CUDA;
  }
}

__global__ void pairs_from_edges_kernel(const int32_t *__restrict__ ed, int64_t n, int32_t max_diff, int32_t *__restrict__ ped) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ped[i] = (ed[i] >= 1 && ed[i] <= max_diff) ? ed[i] : 0;
}

__global__ void edge_keys_kernel(const int32_t *__restrict__ ei, const int32_t *__restrict__ ej, const int32_t *__restrict__ ed,
                                 int64_t n, int shift, uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = ((uint64_t)(uint32_t)ei[i] << shift) | (uint32_t)ej[i]; val[i] = (uint32_t)ed[i]; }
}
__global__ void edge_unpack_kernel(const uint64_t *__restrict__ key, const uint32_t *__restrict__ val, int64_t n, int shift,
                                   int32_t *__restrict__ ei, int32_t *__restrict__ ej, int32_t *__restrict__ ed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    ei[i] = (int32_t)(key[i] >> shift);
    ej[i] = (int32_t)(key[i] & ((1ull << shift) - 1ull));
    ed[i] = (int32_t)val[i];
  }
}

// PCB_CLUSTER_COMPACT_ROWS: per-read row arrays -> one entry per row, in ascending order of the representative
__global__ void row_flag_kernel(const int32_t *__restrict__ row_size, const int32_t *__restrict__ row_call_n, int32_t R,
                                uint32_t *__restrict__ flag, uint32_t *__restrict__ ncall) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  bool rep = r < R && row_size[r] > 0;
  flag[r] = rep ? 1u : 0u;
  ncall[r] = rep && row_call_n[r] > 0 ? (uint32_t)row_call_n[r] : 0u;
}
__global__ void row_compact_kernel(int32_t R, const uint32_t *__restrict__ row_of, const uint32_t *__restrict__ call_at,
                                   const int32_t *__restrict__ t_size, const int64_t *__restrict__ t_key, const int32_t *__restrict__ t_bc,
                                   const int32_t *__restrict__ t_up, const int64_t *__restrict__ t_calloff, const int32_t *__restrict__ t_calln,
                                   const int32_t *__restrict__ t_call, int32_t *__restrict__ o_rep, int32_t *__restrict__ o_size,
                                   int64_t *__restrict__ o_key, int32_t *__restrict__ o_bc, int32_t *__restrict__ o_up,
                                   int64_t *__restrict__ o_calloff, int32_t *__restrict__ o_calln, int32_t *__restrict__ o_call) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || row_of[r + 1] == row_of[r]) return;
  uint32_t i = row_of[r];
  o_rep[i] = r;
  o_size[i] = t_size[r];
  o_key[i] = t_key[r];
  o_bc[i] = t_bc[r];
  o_up[i] = t_up[r];
  o_calln[i] = t_calln[r];
  uint32_t at = call_at[r];
  o_calloff[i] = (int64_t)at;
  int64_t from = t_calloff[r];
  for (int32_t k = 0; k < t_calln[r]; k++) o_call[at + k] = t_call[from + k];
}

__global__ void own_flag_kernel(const int32_t *__restrict__ read_root, int32_t R, uint32_t *__restrict__ flag) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  flag[r] = (r < R && read_root[r] != PCB_FILL) ? 1u : 0u;
}
__global__ void own_compact_kernel(int32_t R, const uint32_t *__restrict__ at, const int32_t *__restrict__ read_root,
                                   const int32_t *__restrict__ read_pos, const int32_t *__restrict__ bucket_of_read,
                                   int32_t *__restrict__ o_read, int32_t *__restrict__ o_root, int32_t *__restrict__ o_pos,
                                   int32_t *__restrict__ o_bucket) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || at[r + 1] == at[r]) return;
  o_read[at[r]] = r; o_root[at[r]] = read_root[r]; o_pos[at[r]] = read_pos[r];
  if (o_bucket) o_bucket[at[r]] = bucket_of_read[r];
}

struct RowArrays {           // device pointers
  int32_t *size; int64_t *key; int32_t *bc, *up; int64_t *calloff; int32_t *calln, *call;
};
// per-read row arrays -> one entry per row (ascending representative); returns after the counts have arrived
void compact_rows(pcb_ctx *ctx, int32_t R, const RowArrays &t, int32_t *o_rep, const RowArrays &o, uint32_t *n_rows, uint32_t *n_calls) {
  DevBuf<uint32_t> row_of(ctx, (size_t)R + 1), call_at(ctx, (size_t)R + 1);
  PCB_LAUNCH(ctx, row_flag_kernel, grid_for((size_t)R + 1, 256), 256, 0, t.size, t.calln, R, row_of.p, call_at.p);
  exclusive_scan_u32(ctx, row_of.p, row_of.p, (size_t)R + 1);
  exclusive_scan_u32(ctx, call_at.p, call_at.p, (size_t)R + 1);
  if (R) PCB_LAUNCH(ctx, row_compact_kernel, grid_for((size_t)R, 256), 256, 0, R, row_of.p, call_at.p, t.size, t.key, t.bc, t.up,
                    t.calloff, t.calln, t.call, o_rep, o.size, o.key, o.bc, o.up, o.calloff, o.calln, o.call);
  PCB_CUDA(cudaMemcpyAsync(n_rows, row_of.p + R, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(n_calls, call_at.p + R, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
}

}  // namespace

extern "C" {

int pcb_abi_version(void) { return PCB_ABI_VERSION; }

const char *pcb_last_error(const pcb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int pcb_ctx_create(int device, pcb_ctx **out) {
  if (!out) return PCB_E_STATE;
  *out = nullptr;
  pcb_ctx *ctx = nullptr;
  try {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw Error(PCB_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                  " (this library has no CPU implementation)");
    if (device < 0 || device >= n) throw Error(PCB_E_CUDA, "device index out of range");
    PCB_CUDA(cudaSetDevice(device));
    ctx = new pcb_ctx();
    ctx->device = device;
    PCB_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    PCB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    PCB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    PCB_CUDA(cudaEventCreateWithFlags(&ctx->ev_alloc, cudaEventDisableTiming));
    PCB_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
    PCB_CUDA(cudaEventCreate(&ctx->ev0));
    PCB_CUDA(cudaEventCreate(&ctx->ev1));
    for (int k = 0; k < 8; k++) PCB_CUDA(cudaEventCreate(&ctx->stage_ev[k]));
    cudaMemPool_t pool;
    PCB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    PCB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    *out = ctx;
    return PCB_OK;
  } catch (const Error &e) {
    g_create_error = e.what();
    delete ctx;
    cudaGetLastError();
    return e.code;
  }
}

// (e) what pcb_shard_pairs leaves for pcb_shard_merge: the bucket arrays and the packed barcodes of the whole read set
struct pcb_shard_state {
  pcb::DevBuf<int32_t> bptr, breads;
  pcb::PackedBarcodes uniq;
  int32_t R = 0, U = 0;
  bool valid = false;
};

void pcb_ctx_destroy(pcb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->edges.ei) cudaFreeAsync(ctx->edges.ei, ctx->stream);
  if (ctx->edges.ej) cudaFreeAsync(ctx->edges.ej, ctx->stream);
  if (ctx->edges.ed) cudaFreeAsync(ctx->edges.ed, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int k = 0; k < 8; k++) if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
  if (ctx->ev_alloc) cudaEventDestroy(ctx->ev_alloc);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->shard) { delete ctx->shard; ctx->shard = nullptr; }
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int64_t pcb_stat(const pcb_ctx *ctx, int which) {
  if (!ctx || which < 0 || which >= PCB_STAT_COUNT) return -1;
  return ctx->stats[which];
}

int pcb_set_option(pcb_ctx *ctx, int option, int64_t value) {
  if (!ctx) return PCB_E_STATE;
  if (option == PCB_OPT_REPLAY_SHAPE && value >= PCB_REPLAY_AUTO && value <= PCB_REPLAY_HUGE) {
    ctx->replay_shape = (int)value;
    return PCB_OK;
  }
  if (option == PCB_OPT_TRACE) { ctx->trace = value != 0; return PCB_OK; }
  if (option == 100 && value >= 0 && value <= 32768) { ctx->small_pad = (int)value; return PCB_OK; }      // measurement aid, not in the header
  if (option == 101 && value >= 1 && value <= 64) { ctx->huge_warps_per_sm = (int)value; return PCB_OK; }  // measurement aid: resident warps of the giant tier
  if (option == PCB_OPT_HUGE_BUCKETS && value >= 0 && value <= INT32_MAX) { ctx->huge_buckets = (int)value; return PCB_OK; }
  if (option == PCB_OPT_SMALL_WARPS && (value == 1 || value == 2 || value == 4)) { ctx->small_warps = (int)value; return PCB_OK; }
  if (option == PCB_OPT_SEED_BINS_LOG2 && value >= 8 && value <= 28) { ctx->seed_bins_log2 = (int)value; return PCB_OK; }
  if (option == PCB_OPT_PROBE_CLASSES) { ctx->probe_classes = value != 0; return PCB_OK; }
  if (option == PCB_OPT_PROBE_STAGING) { ctx->probe_staging = value != 0; return PCB_OK; }
  if (option == PCB_OPT_HUGE_READS && value >= 0 && value <= INT32_MAX) { ctx->huge_reads = (int)value; return PCB_OK; }
  if (option == PCB_OPT_HUGE_CAP && value >= 16 && value <= (1 << 20) && (value & (value - 1)) == 0) { ctx->huge_cap = (int)value; return PCB_OK; }
  if (option == PCB_OPT_HUGE_WINDOW && value >= 1 && value <= (1 << 24)) { ctx->huge_window = (int)value; return PCB_OK; }
  ctx->err = "pcb_set_option: unknown option or value";
  return PCB_E_INPUT;
}

int pcb_sync(pcb_ctx *ctx) {
  return guarded(ctx, [&] {});
}

void *pcb_stream(pcb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int pcb_bucket(pcb_ctx *ctx, int mem, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int32_t R,
               int32_t stride, int32_t *bucket_of_read, int32_t *n_buckets, int32_t *bucket_ptr,
               int32_t *bucket_reads, uint8_t *bucket_qsum) {
  return guarded(ctx, [&] {
    if (R < 0 || stride < 1) throw Error(PCB_E_INPUT, "bad shape");
    size_t cells = (size_t)R * stride;
    InArr<uint8_t> d_seq(ctx, mem, seq, cells), d_qual(ctx, mem, qual, cells);
    InArr<int32_t> d_len(ctx, mem, len, R);
    OutArr<int32_t> o_bor(ctx, mem, bucket_of_read, R), o_ptr(ctx, mem, bucket_ptr, (size_t)R + 1), o_reads(ctx, mem, bucket_reads, R);
    OutArr<uint8_t> o_qsum(ctx, mem, bucket_qsum, cells);
    if (bucket_qsum && !qual) throw Error(PCB_E_INPUT, "quality sums requested without qualities");
    BucketOut bo{o_bor.p, o_ptr.p, o_reads.p, o_qsum.p};
    stage_begin(ctx, PCB_STAT_T_BUCKET_NS, PCB_STAT_T_BUCKET_NS);
    int32_t U = bucket_reads_dev(ctx, d_seq.p, d_qual.p, d_len.p, R, stride, bo, nullptr);
    stage_end(ctx, PCB_STAT_T_BUCKET_NS);
    if (n_buckets) *n_buckets = U;
    o_bor.commit(R); o_ptr.commit((size_t)U + 1); o_reads.commit(R); o_qsum.commit((size_t)U * stride);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
  });
}

int pcb_edit_pairs(pcb_ctx *ctx, int mem, const uint8_t *seq, const int32_t *len, int32_t U, int32_t stride, int32_t k,
                   int32_t q_begin, int32_t q_end, int32_t t_begin, int32_t t_end, int64_t *n_edges) {
  return guarded(ctx, [&] {
    if (U < 0 || stride < 1) throw Error(PCB_E_INPUT, "bad shape");
    InArr<uint8_t> d_seq(ctx, mem, seq, (size_t)U * stride);
    InArr<int32_t> d_len(ctx, mem, len, U);
    PackedBarcodes bc;
    pack_barcodes(ctx, d_seq.p, d_len.p, nullptr, U, stride, bc);
    stage_begin(ctx, PCB_STAT_T_INDEX_NS, PCB_STAT_T_HITS_NS);
    edit_pairs_dev(ctx, bc, k, q_begin, q_end, t_begin, t_end);
    if (n_edges) *n_edges = ctx->edges.n;
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);
  });
}

int pcb_edit_pairs_fetch(pcb_ctx *ctx, int mem, int32_t *ei, int32_t *ej, int32_t *ed, int64_t cap) {
  return guarded(ctx, [&] {
    if (!ctx->edges.valid) throw Error(PCB_E_STATE, "pcb_edit_pairs_fetch before pcb_edit_pairs");
    int64_t n = ctx->edges.n;
    if (cap < n) throw Error(PCB_E_CAPACITY, "edge buffers hold " + std::to_string(cap) + ", need " + std::to_string(n));
    cudaMemcpyKind kind = mem == PCB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (n) {
      PCB_CUDA(cudaMemcpyAsync(ei, ctx->edges.ei, n * sizeof(int32_t), kind, ctx->stream));
      PCB_CUDA(cudaMemcpyAsync(ej, ctx->edges.ej, n * sizeof(int32_t), kind, ctx->stream));
      PCB_CUDA(cudaMemcpyAsync(ed, ctx->edges.ed, n * sizeof(int32_t), kind, ctx->stream));
    }
  });
}

int pcb_sort_edges(pcb_ctx *ctx, int mem, int32_t *ei, int32_t *ej, int32_t *ed, int64_t n, int32_t U) {
  return guarded(ctx, [&] {
    if (n <= 0) return;
    InArr<int32_t> d_i(ctx, mem, ei, n), d_j(ctx, mem, ej, n), d_d(ctx, mem, ed, n);
    int shift = bits_for((uint64_t)(U > 1 ? U - 1 : 1));
    DevBuf<uint64_t> k0(ctx, n), k1(ctx, n);
    DevBuf<uint32_t> v0(ctx, n), v1(ctx, n);
    PCB_LAUNCH(ctx, edge_keys_kernel, grid_for(n, 256), 256, 0, d_i.p, d_j.p, d_d.p, n, shift, k0.p, v0.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, n, 2 * shift);
    OutArr<int32_t> o_i(ctx, mem, ei, n), o_j(ctx, mem, ej, n), o_d(ctx, mem, ed, n);
    PCB_LAUNCH(ctx, edge_unpack_kernel, grid_for(n, 256), 256, 0, w ? k1.p : k0.p, w ? v1.p : v0.p, n, shift, o_i.p, o_j.p, o_d.p);
    o_i.commit(n); o_j.commit(n); o_d.commit(n);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_merge(pcb_ctx *ctx, int mem, int32_t R, int32_t U, const int32_t *bucket_ptr, const int32_t *bucket_reads,
              const int32_t *lexrank, const int32_t *bc_id, const int32_t *up_id, const int32_t *geno_ptr,
              const int32_t *geno_mut, const int32_t *geno_q, int64_t nnz, const int32_t *mut_rank, int32_t n_mut, int64_t n_pairs,
              const int32_t *pair_q, const int32_t *pair_r, const int32_t *pair_d, const int32_t *pair_ed,
              int32_t max_diff, int32_t min_qual, double min_jaccard, int32_t min_matches,
              const uint8_t *bucket_seq, const int32_t *bucket_len, int32_t stride, int32_t on_demand_k,
              int32_t shard_rank, int32_t shard_count, int32_t flags, int32_t *read_root, int32_t *read_pos, int32_t *row_size,
              int64_t *row_key, int32_t *row_bc, int32_t *row_up, int64_t *row_call_off, int32_t *row_call_n, int32_t *call_mut) {
  return guarded(ctx, [&] {
    if (R < 0 || U < 0 || n_pairs < 0 || nnz < 0) throw Error(PCB_E_INPUT, "bad shape");
    if (mem == PCB_MEM_HOST && R > 0 && geno_ptr[R] != nnz) throw Error(PCB_E_INPUT, "nnz is not geno_ptr[R]");
    InArr<int32_t> d_bptr(ctx, mem, bucket_ptr, (size_t)U + 1), d_breads(ctx, mem, bucket_reads, R), d_lex(ctx, mem, lexrank, R),
        d_bc(ctx, mem, bc_id, R), d_up(ctx, mem, up_id, R), d_gptr(ctx, mem, geno_ptr, (size_t)R + 1),
        d_gmut(ctx, mem, geno_mut, nnz), d_gq(ctx, mem, geno_q, nnz), d_rank(ctx, mem, mut_rank, n_mut),
        d_pq(ctx, mem, pair_q, n_pairs), d_pr(ctx, mem, pair_r, n_pairs), d_pd(ctx, mem, pair_d, n_pairs),
        d_pe(ctx, mem, pair_ed, n_pairs);
    OutArr<int32_t> o_root(ctx, mem, read_root, R), o_pos(ctx, mem, read_pos, R), o_size(ctx, mem, row_size, R),
        o_bc(ctx, mem, row_bc, R), o_up(ctx, mem, row_up, R), o_calln(ctx, mem, row_call_n, R), o_call(ctx, mem, call_mut, nnz);
    OutArr<int64_t> o_key(ctx, mem, row_key, R), o_calloff(ctx, mem, row_call_off, R);
    MergeIn in{R, U, n_mut, d_bptr.p, d_breads.p, d_lex.p, d_bc.p, d_up.p, d_gptr.p, d_gmut.p, d_gq.p, d_rank.p,
               n_pairs, d_pq.p, d_pr.p, d_pd.p, d_pe.p, max_diff, min_qual, min_matches, min_jaccard, shard_rank, shard_count};
    MergeOut out{o_root.p, o_pos.p, o_size.p, o_key.p, o_bc.p, o_up.p, o_calloff.p, o_calln.p, o_call.p};
    in.nnz = nnz;
    in.pairs_sorted = (flags & PCB_MERGE_PAIRS_SORTED) != 0;
    in.shard_layout = (flags & PCB_MERGE_SHARD_LAYOUT) != 0;
    PackedBarcodes bc;
    if (bucket_seq && bucket_len && on_demand_k >= 0 && U > 0) {
      InArr<uint8_t> d_bseq(ctx, mem, bucket_seq, (size_t)U * stride);
      InArr<int32_t> d_blen(ctx, mem, bucket_len, U);
      pack_barcodes(ctx, d_bseq.p, d_blen.p, nullptr, U, stride, bc);
      in.planes = bc.planes.p; in.blen = bc.len.p; in.compute_k = on_demand_k; in.wide = bc.max_len > 32;
    }
    stage_begin(ctx, PCB_STAT_T_PREP_NS, PCB_STAT_T_SCATTER_NS);
    merge_dev(ctx, in, out);
    stage_collect(ctx);
    o_root.commit(R); o_pos.commit(R); o_size.commit(R); o_bc.commit(R); o_up.commit(R); o_calln.commit(R);
    o_call.commit(nnz); o_key.commit(R); o_calloff.commit(R);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- (e) the sharded step as two calls around the one collective (pacybara_b200/multigpu.py)

namespace {
// the ranks' edge lists as the all-gather left them -- [world][3][m] int32, counts[r] valid columns -- as sort keys
__global__ void gathered_keys_kernel(const int32_t *__restrict__ buf, int32_t world, int64_t m, const int64_t *__restrict__ start,
                                     int64_t n, int shift, uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t r = 0;
  while (r + 1 < world && start[r + 1] <= i) r++;
  const int64_t c = i - start[r];
  const int32_t *b = buf + (size_t)r * 3 * (size_t)m;
  key[i] = ((uint64_t)(uint32_t)b[c] << shift) | (uint32_t)b[(size_t)m + c];
  val[i] = (uint32_t)b[2 * (size_t)m + c];
}
}  // namespace

int pcb_shard_pairs(pcb_ctx *ctx, int mem, const uint8_t *seq, const int32_t *len, int32_t R, int32_t stride, int32_t k,
                    int32_t q_group, int32_t n_q_groups, int32_t t_group, int32_t n_t_groups, int32_t *bucket_of_read,
                    int32_t *n_buckets, int64_t *n_edges) {
  return guarded(ctx, [&] {
    if (mem != PCB_MEM_DEVICE) throw Error(PCB_E_INPUT, "pcb_shard_pairs works on device arrays");
    if (R < 0 || stride < 1 || !bucket_of_read) throw Error(PCB_E_INPUT, "bad shape");
    if (n_q_groups < 1 || n_t_groups < 1 || q_group < 0 || q_group >= n_q_groups || t_group < 0 || t_group >= n_t_groups)
      throw Error(PCB_E_INPUT, "bad pair grid");
    if (!ctx->shard) ctx->shard = new pcb_shard_state();
    pcb_shard_state &st = *ctx->shard;
    st.valid = false;
    // what outlives the call comes from the pool, not from the workspace arena, and is kept from step to step
    ctx->arena_on = false;
    if (st.bptr.n < (size_t)R + 1) st.bptr.alloc(ctx, (size_t)R + 1);
    if (st.breads.n < (size_t)(R > 0 ? R : 1)) st.breads.alloc(ctx, R > 0 ? R : 1);
    if (st.uniq.planes.n < (size_t)R) st.uniq.planes.alloc(ctx, (size_t)R);
    if (st.uniq.len.n < (size_t)R) st.uniq.len.alloc(ctx, (size_t)R);
    ctx->arena_on = true;
    st.uniq.n = 0; st.uniq.has_hist = false; st.uniq.n_opaque = 0; st.uniq.min_len = st.uniq.max_len = 0;
    st.R = R;
    stage_begin(ctx);
    BucketOut bo{bucket_of_read, st.bptr.p, st.breads.p, nullptr};
    int32_t U = 0;
    if (R > 0) {
      BucketPending pending;
      bucket_launch(ctx, seq, len, R, stride, bo, &st.uniq, pending);
      stage_end(ctx, PCB_STAT_T_BUCKET_NS);
      PCB_CUDA(cudaStreamSynchronize(ctx->stream));
      U = bucket_finish(ctx, bo, &st.uniq, pending);
    } else {
      st.uniq.n = 0;
      stage_end(ctx, PCB_STAT_T_BUCKET_NS);
    }
    st.U = U;
    if (n_buckets) *n_buckets = U;
    const int32_t q_lo = (int32_t)(((int64_t)U * q_group) / n_q_groups), q_hi = (int32_t)(((int64_t)U * (q_group + 1)) / n_q_groups);
    const int32_t t_lo = (int32_t)(((int64_t)U * t_group) / n_t_groups), t_hi = (int32_t)(((int64_t)U * (t_group + 1)) / n_t_groups);
    edit_pairs_dev(ctx, st.uniq, k, q_lo, q_hi, t_lo, t_hi);
    stage_collect(ctx);
    if (n_edges) *n_edges = ctx->edges.n;
    st.valid = true;
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_shard_merge(pcb_ctx *ctx, int mem, const int32_t *gathered, int32_t world, int64_t m, const int64_t *counts,
                    const int32_t *bucket_of_read, const int32_t *lexrank, const int32_t *up_id, const int32_t *geno_ptr,
                    const int32_t *geno_mut, const int32_t *geno_q, int64_t nnz, const int32_t *mut_rank, int32_t n_mut,
                    int32_t max_diff, int32_t min_qual, double min_jaccard, int32_t min_matches, int32_t on_demand,
                    int32_t shard_rank, int32_t shard_count, int32_t flags, int64_t *n_edges, int32_t *read_root, int32_t *read_pos,
                    int32_t *row_size, int64_t *row_key, int32_t *row_bc, int32_t *row_up, int64_t *row_call_off,
                    int32_t *row_call_n, int32_t *call_mut) {
  return guarded(ctx, [&] {
    if (mem != PCB_MEM_DEVICE) throw Error(PCB_E_INPUT, "pcb_shard_merge works on device arrays");
    if (!ctx->shard || !ctx->shard->valid) throw Error(PCB_E_STATE, "pcb_shard_merge before pcb_shard_pairs");
    pcb_shard_state &st = *ctx->shard;
    if (world < 1 || m < 0 || nnz < 0 || !counts) throw Error(PCB_E_INPUT, "bad shape");
    const int32_t R = st.R, U = st.U;
    // ---- the gathered edges in canonical order (i, j)
    stage_begin(ctx, PCB_STAT_T_PREP_NS, PCB_STAT_T_SCATTER_NS);
    std::vector<int64_t> start((size_t)world + 1, 0);
    for (int32_t r = 0; r < world; r++) {
      if (counts[r] < 0 || counts[r] > m) throw Error(PCB_E_INPUT, "bad edge count");
      start[r + 1] = start[r] + counts[r];
    }
    const int64_t E = start[world];
    if (n_edges) *n_edges = E;
    DevBuf<int32_t> ei(ctx, (size_t)E), ej(ctx, (size_t)E), ed(ctx, (size_t)E), ped(ctx, (size_t)E);
    if (E > 0) {
      DevBuf<int64_t> d_start(ctx, (size_t)world + 1);
      PCB_CUDA(cudaMemcpyAsync(d_start.p, start.data(), ((size_t)world + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
      const int shift = bits_for((uint64_t)(U > 1 ? U - 1 : 1));
      DevBuf<uint64_t> k0(ctx, E), k1(ctx, E);
      DevBuf<uint32_t> v0(ctx, E), v1(ctx, E);
      PCB_LAUNCH(ctx, gathered_keys_kernel, grid_for(E, 256), 256, 0, gathered, world, m, d_start.p, E, shift, k0.p, v0.p);
      const int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, E, 2 * shift);
      PCB_LAUNCH(ctx, edge_unpack_kernel, grid_for(E, 256), 256, 0, w ? k1.p : k0.p, w ? v1.p : v0.p, E, shift, ei.p, ej.p, ed.p);
      PCB_LAUNCH(ctx, pairs_from_edges_kernel, grid_for(E, 256), 256, 0, ed.p, E, max_diff, ped.p);
      PCB_CUDA(cudaStreamSynchronize(ctx->stream));          // (start[] is host memory of this frame)
    }
    MergeIn in{R, U, n_mut, st.bptr.p, st.breads.p, lexrank, bucket_of_read, up_id, geno_ptr, geno_mut, geno_q, mut_rank,
               E, ei.p, ej.p, ed.p, ped.p, max_diff, min_qual, min_matches, min_jaccard, shard_rank, shard_count};
    if (on_demand) { in.planes = st.uniq.planes.p; in.blen = st.uniq.len.p; in.compute_k = 2 * max_diff; in.wide = st.uniq.max_len > 32; }
    in.nnz = nnz;
    in.pairs_sorted = true;
    in.shard_layout = (flags & PCB_MERGE_SHARD_LAYOUT) != 0;
    in.bucket_of_read = bucket_of_read;
    MergeOut out{read_root, read_pos, row_size, row_key, row_bc, row_up, row_call_off, row_call_n, call_mut};
    merge_dev(ctx, in, out);
    stage_collect(ctx);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_cluster_reads(pcb_ctx *ctx, int mem, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int32_t R,
                      int32_t stride, const int32_t *lexrank, const int32_t *up_id, const int32_t *geno_ptr,
                      const int32_t *geno_mut, const int32_t *geno_q, int64_t nnz, const int32_t *mut_rank, int32_t n_mut,
                      int32_t max_diff, int32_t min_qual, double min_jaccard, int32_t min_matches, int32_t flags,
                      int32_t *bucket_of_read, int32_t *n_buckets, int64_t *n_edges, int32_t *read_root,
                      int32_t *read_pos, int32_t *row_size, int64_t *row_key, int32_t *row_bc, int32_t *row_up,
                      int64_t *row_call_off, int32_t *row_call_n, int32_t *call_mut, int32_t *row_rep) {
  return guarded(ctx, [&] {
    if (R < 0 || stride < 1) throw Error(PCB_E_INPUT, "bad shape");
    if (max_diff < 0 || max_diff > 3) throw Error(PCB_E_INPUT, "maxDiff must be between 0 and 3");
    const bool compact = (flags & PCB_CLUSTER_COMPACT_ROWS) != 0;
    if (compact && !row_rep) throw Error(PCB_E_INPUT, "PCB_CLUSTER_COMPACT_ROWS needs row_rep");
    if (nnz < 0 || (mem == PCB_MEM_HOST && R > 0 && geno_ptr[R] != nnz)) throw Error(PCB_E_INPUT, "nnz is not geno_ptr[R]");
    size_t cells = (size_t)R * stride;
    (void)qual;   // qualities only feed the per-bucket sums of pcb_bucket; the clustering never reads them
    InArr<uint8_t> d_seq(ctx, mem, seq, cells);
    InArr<int32_t> d_len(ctx, mem, len, R);
    // what only the merge reads is copied on a second stream while bucketing and the pair search run
    InArr<int32_t> d_lex(ctx, mem, lexrank, R, true), d_up(ctx, mem, up_id, R, true),
        d_gptr(ctx, mem, geno_ptr, (size_t)R + 1, true), d_gmut(ctx, mem, geno_mut, nnz, true), d_gq(ctx, mem, geno_q, nnz, true),
        d_rank(ctx, mem, mut_rank, n_mut, true);
    if (mem == PCB_MEM_HOST) {
      PCB_CUDA(cudaEventRecord(ctx->ev_alloc, ctx->stream));                  // the buffers exist from here on
      PCB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_alloc, 0));
      d_lex.copy_late(ctx->copy_stream); d_up.copy_late(ctx->copy_stream); d_gptr.copy_late(ctx->copy_stream);
      d_gmut.copy_late(ctx->copy_stream); d_gq.copy_late(ctx->copy_stream); d_rank.copy_late(ctx->copy_stream);
      PCB_CUDA(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
    }
    OutArr<int32_t> o_bor(ctx, mem, bucket_of_read, R), o_root(ctx, mem, read_root, R), o_pos(ctx, mem, read_pos, R),
        o_size(ctx, mem, row_size, R), o_bc(ctx, mem, row_bc, R), o_up(ctx, mem, row_up, R),
        o_calln(ctx, mem, row_call_n, R), o_call(ctx, mem, call_mut, nnz);
    OutArr<int64_t> o_key(ctx, mem, row_key, R), o_calloff(ctx, mem, row_call_off, R);
    DevBuf<int32_t> bptr(ctx, (size_t)R + 1), breads(ctx, R), bor_tmp;
    int32_t *bor = o_bor.p;
    if (!bor) { bor_tmp.alloc(ctx, R); bor = bor_tmp.p; }
    // a1: buckets
    stage_begin(ctx);
    PackedBarcodes uniq;
    BucketOut bo{bor, bptr.p, breads.p, nullptr};
    int32_t U = 0;
    if (R > 0) {
      BucketPending pending;
      bucket_launch(ctx, d_seq.p, d_len.p, R, stride, bo, &uniq, pending);
      stage_end(ctx, PCB_STAT_T_BUCKET_NS);
      PCB_CUDA(cudaStreamSynchronize(ctx->stream));           // round trip 1 of the call: U, input errors, lengths seen
      U = bucket_finish(ctx, bo, &uniq, pending);
    } else {
      uniq.n = 0; uniq.planes.alloc(ctx, 0); uniq.len.alloc(ctx, 0);
      stage_end(ctx, PCB_STAT_T_BUCKET_NS);
    }
    if (n_buckets) *n_buckets = U;
    // a2+a3: either the full table up to 2*maxDiff (what pacybara.sh:466-468 asks of calcEdits.R), or only
    // the pairs a pass can visit (<= maxDiff) with the transitive-rule distances scored on demand
    const bool full = (flags & PCB_CLUSTER_FULL_TABLE) != 0;
    edit_pairs_dev(ctx, uniq, full ? 2 * max_diff : max_diff, 0, U);
    int64_t E = ctx->edges.n;
    if (n_edges) *n_edges = E;
    DevBuf<int32_t> ped(ctx, (size_t)E);
    if (E) PCB_LAUNCH(ctx, pairs_from_edges_kernel, grid_for(E, 256), 256, 0, ctx->edges.ed, E, max_diff, ped.p);
    if (mem == PCB_MEM_HOST) PCB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
    // a4..a13: merge; the barcode id of a read is its bucket
    MergeIn in{R, U, n_mut, bptr.p, breads.p, d_lex.p, bor, d_up.p, d_gptr.p, d_gmut.p, d_gq.p, d_rank.p,
               E, ctx->edges.ei, ctx->edges.ej, ctx->edges.ed, ped.p, max_diff, min_qual, min_matches, min_jaccard, 0, 1};
    if (!full) { in.planes = uniq.planes.p; in.blen = uniq.len.p; in.compute_k = 2 * max_diff; in.wide = uniq.max_len > 32; }
    in.nnz = nnz;
    in.pairs_sorted = true;            // the pair search delivers q < r, ascending by (q, r)
    in.bucket_of_read = bor;
    if (!compact) {
      MergeOut out{o_root.p, o_pos.p, o_size.p, o_key.p, o_bc.p, o_up.p, o_calloff.p, o_calln.p, o_call.p};
      merge_dev(ctx, in, out);
      stage_collect(ctx);
      o_bor.commit(R); o_root.commit(R); o_pos.commit(R); o_size.commit(R); o_bc.commit(R); o_up.commit(R);
      o_calln.commit(R); o_call.commit(nnz); o_key.commit(R); o_calloff.commit(R);
    } else {
      // per-read row arrays stay on the device; only the rows travel
      DevBuf<int32_t> t_size(ctx, R), t_bc(ctx, R), t_up(ctx, R), t_calln(ctx, R), t_call(ctx, (size_t)nnz);
      DevBuf<int64_t> t_key(ctx, R), t_calloff(ctx, R);
      MergeOut out{o_root.p, o_pos.p, t_size.p, t_key.p, t_bc.p, t_up.p, t_calloff.p, t_calln.p, t_call.p};
      merge_dev(ctx, in, out);
      OutArr<int32_t> o_rep(ctx, mem, row_rep, R);
      uint32_t h_rows = 0, h_calls = 0;
      compact_rows(ctx, R, RowArrays{t_size.p, t_key.p, t_bc.p, t_up.p, t_calloff.p, t_calln.p, t_call.p}, o_rep.p,
                   RowArrays{o_size.p, o_key.p, o_bc.p, o_up.p, o_calloff.p, o_calln.p, o_call.p}, &h_rows, &h_calls);
      o_bor.commit(R); o_root.commit(R); o_pos.commit(R);
      PCB_CUDA(cudaStreamSynchronize(ctx->stream));
      stage_collect(ctx);
      ctx->stats[PCB_STAT_ROWS] = h_rows;
      ctx->stats[PCB_STAT_CALLS] = h_calls;
      o_rep.commit(h_rows); o_size.commit(h_rows); o_bc.commit(h_rows); o_up.commit(h_rows); o_calln.commit(h_rows);
      o_key.commit(h_rows); o_calloff.commit(h_rows); o_call.commit(h_calls);
    }
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_compact_shard(pcb_ctx *ctx, int mem_in, int mem_out, int32_t R, int64_t nnz, const int32_t *read_root, const int32_t *read_pos,
                      const int32_t *row_size, const int64_t *row_key, const int32_t *row_bc, const int32_t *row_up,
                      const int64_t *row_call_off, const int32_t *row_call_n, const int32_t *call_mut, const int32_t *bucket_of_read,
                      int32_t *n_own, int32_t *own_read, int32_t *own_root, int32_t *own_pos, int32_t *own_bucket,
                      int32_t *n_rows, int64_t *n_calls, int32_t *row_rep, int32_t *c_size, int64_t *c_key, int32_t *c_bc,
                      int32_t *c_up, int64_t *c_call_off, int32_t *c_call_n, int32_t *c_call_mut) {
  return guarded(ctx, [&] {
    if (R < 0 || nnz < 0) throw Error(PCB_E_INPUT, "bad shape");
    if ((bucket_of_read == nullptr) != (own_bucket == nullptr)) throw Error(PCB_E_INPUT, "bucket_of_read and own_bucket go together");
    InArr<int32_t> i_bor(ctx, mem_in, bucket_of_read, R);
    OutArr<int32_t> o_bucket(ctx, mem_out, own_bucket, R);
    InArr<int32_t> i_root(ctx, mem_in, read_root, R), i_pos(ctx, mem_in, read_pos, R), i_size(ctx, mem_in, row_size, R),
        i_bc(ctx, mem_in, row_bc, R), i_up(ctx, mem_in, row_up, R), i_calln(ctx, mem_in, row_call_n, R), i_call(ctx, mem_in, call_mut, nnz);
    InArr<int64_t> i_key(ctx, mem_in, row_key, R), i_calloff(ctx, mem_in, row_call_off, R);
    OutArr<int32_t> o_read(ctx, mem_out, own_read, R), o_root(ctx, mem_out, own_root, R), o_pos(ctx, mem_out, own_pos, R),
        o_rep(ctx, mem_out, row_rep, R), o_size(ctx, mem_out, c_size, R), o_bc(ctx, mem_out, c_bc, R), o_up(ctx, mem_out, c_up, R),
        o_calln(ctx, mem_out, c_call_n, R), o_call(ctx, mem_out, c_call_mut, nnz);
    OutArr<int64_t> o_key(ctx, mem_out, c_key, R), o_calloff(ctx, mem_out, c_call_off, R);
    DevBuf<uint32_t> own_at(ctx, (size_t)R + 1);
    PCB_LAUNCH(ctx, own_flag_kernel, grid_for((size_t)R + 1, 256), 256, 0, i_root.p, R, own_at.p);
    exclusive_scan_u32(ctx, own_at.p, own_at.p, (size_t)R + 1);
    if (R) PCB_LAUNCH(ctx, own_compact_kernel, grid_for((size_t)R, 256), 256, 0, R, own_at.p, i_root.p, i_pos.p, i_bor.p, o_read.p, o_root.p,
                      o_pos.p, o_bucket.p);
    uint32_t h_own = 0, h_rows = 0, h_calls = 0;
    PCB_CUDA(cudaMemcpyAsync(&h_own, own_at.p + R, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    compact_rows(ctx, R, RowArrays{const_cast<int32_t *>(i_size.p), const_cast<int64_t *>(i_key.p), const_cast<int32_t *>(i_bc.p),
                                   const_cast<int32_t *>(i_up.p), const_cast<int64_t *>(i_calloff.p), const_cast<int32_t *>(i_calln.p),
                                   const_cast<int32_t *>(i_call.p)},
                 o_rep.p, RowArrays{o_size.p, o_key.p, o_bc.p, o_up.p, o_calloff.p, o_calln.p, o_call.p}, &h_rows, &h_calls);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_own) *n_own = (int32_t)h_own;
    if (n_rows) *n_rows = (int32_t)h_rows;
    if (n_calls) *n_calls = (int64_t)h_calls;
    o_read.commit(h_own); o_root.commit(h_own); o_pos.commit(h_own); o_bucket.commit(h_own);
    o_rep.commit(h_rows); o_size.commit(h_rows); o_bc.commit(h_rows); o_up.commit(h_rows); o_calln.commit(h_rows);
    o_key.commit(h_rows); o_calloff.commit(h_rows); o_call.commit(h_calls);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_match_barcodes(pcb_ctx *ctx, int mem, const uint8_t *lib_seq, const int32_t *lib_len, int32_t L, int32_t lib_stride,
                       const uint8_t *q_seq, const int32_t *q_len, int32_t Q, int32_t q_stride, int32_t max_err,
                       int32_t *best_d, int32_t *n_best, int32_t *first_hit, int64_t *n_pairs) {
  return guarded(ctx, [&] {
    if (L < 0 || Q < 0 || lib_stride < 1 || q_stride < 1) throw Error(PCB_E_INPUT, "bad shape");
    InArr<uint8_t> d_lseq(ctx, mem, lib_seq, (size_t)L * lib_stride), d_qseq(ctx, mem, q_seq, (size_t)Q * q_stride);
    InArr<int32_t> d_llen(ctx, mem, lib_len, L), d_qlen(ctx, mem, q_len, Q);
    OutArr<int32_t> o_best(ctx, mem, best_d, Q), o_n(ctx, mem, n_best, Q), o_first(ctx, mem, first_hit, Q);
    if (Q > 0 && (!o_best.p || !o_n.p || !o_first.p)) throw Error(PCB_E_INPUT, "all three outputs are required");
    PackedBarcodes lib, qs;
    lib.n = 0; qs.n = 0;
    if (L) pack_barcodes(ctx, d_lseq.p, d_llen.p, nullptr, L, lib_stride, lib);
    if (Q) pack_barcodes(ctx, d_qseq.p, d_qlen.p, nullptr, Q, q_stride, qs, true);       // N and friends: characters that match nothing
    match_barcodes_dev(ctx, lib, qs, max_err, o_best.p, o_n.p, o_first.p);
    if (n_pairs) *n_pairs = ctx->edges.n;
    o_best.commit(Q); o_n.commit(Q); o_first.commit(Q);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_tag_rows(pcb_ctx *ctx, int mem, int32_t n_rows, const int32_t *row_bc, const int32_t *row_up, const int32_t *row_size,
                 int32_t n_bc_ids, int32_t n_up_ids, double min_ratio, uint8_t *collision, uint8_t *up_collision, uint8_t *usable) {
  return guarded(ctx, [&] {
    if (n_rows < 0 || n_bc_ids < 0 || n_up_ids < 0) throw Error(PCB_E_INPUT, "bad shape");
    if (!(min_ratio >= 0.5 && min_ratio <= 1.0)) throw Error(PCB_E_INPUT, "minRatio must be between 0.5 and 1 (pacybara_softfilter.R:36-39)");
    InArr<int32_t> d_bc(ctx, mem, row_bc, n_rows), d_up(ctx, mem, row_up, n_rows), d_size(ctx, mem, row_size, n_rows);
    OutArr<uint8_t> o_c(ctx, mem, collision, n_rows), o_u(ctx, mem, up_collision, n_rows), o_ok(ctx, mem, usable, n_rows);
    tag_rows_dev(ctx, n_rows, d_bc.p, d_up.p, d_size.p, n_bc_ids, n_up_ids, min_ratio, o_c.p, o_u.p, o_ok.p);
    o_c.commit(n_rows); o_u.commit(n_rows); o_ok.commit(n_rows);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int pcb_merge_libraries(pcb_ctx *ctx, int mem, int32_t n1, const int32_t *vbc1, const int32_t *geno1, const int32_t *size1,
                        int32_t n2, const int32_t *vbc2, const int32_t *geno2, const int32_t *size2, int32_t n_ids,
                        int32_t *size1_out, int32_t *target2, int32_t *klass2) {
  return guarded(ctx, [&] {
    if (n1 < 0 || n2 < 0 || n_ids < 0) throw Error(PCB_E_INPUT, "bad shape");
    InArr<int32_t> d_v1(ctx, mem, vbc1, n1), d_g1(ctx, mem, geno1, n1), d_s1(ctx, mem, size1, n1), d_v2(ctx, mem, vbc2, n2),
        d_g2(ctx, mem, geno2, n2), d_s2(ctx, mem, size2, n2);
    OutArr<int32_t> o_s(ctx, mem, size1_out, n1), o_t(ctx, mem, target2, n2), o_k(ctx, mem, klass2, n2);
    if ((n1 > 0 && !o_s.p) || (n2 > 0 && (!o_t.p || !o_k.p))) throw Error(PCB_E_INPUT, "all three outputs are required");
    merge_libraries_dev(ctx, n1, d_v1.p, d_g1.p, d_s1.p, n2, d_v2.p, d_g2.p, d_s2.p, n_ids, o_s.p, o_t.p, o_k.p);
    o_s.commit(n1); o_t.commit(n2); o_k.commit(n2);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- a14: the writer's rows as a ragged concatenation (format.cu)
int pcb_concat_tokens(pcb_ctx *ctx, int mem, const uint8_t *src, int64_t n_src, const int64_t *tok_off, const int32_t *tok_len, int64_t n_tok,
                      const int32_t *item_tok, const uint8_t *item_sep, int64_t n_items, uint8_t *out, int64_t out_cap, int64_t *out_len) {
  if (out_len) *out_len = 0;
  return guarded(ctx, [&] {
    if (n_src < 0 || n_tok < 0 || n_items < 0 || out_cap < 0 || !out_len) throw Error(PCB_E_INPUT, "bad shape");
    if (n_items > 0 && (!tok_off || !tok_len || !item_tok || !item_sep)) throw Error(PCB_E_INPUT, "token and item arrays are required");
    InArr<uint8_t> d_src(ctx, mem, src, n_src), d_sep(ctx, mem, item_sep, n_items);
    InArr<int64_t> d_off(ctx, mem, tok_off, n_tok);
    InArr<int32_t> d_len(ctx, mem, tok_len, n_tok), d_tok(ctx, mem, item_tok, n_items);
    OutArr<uint8_t> o_out(ctx, mem, out, out_cap);
    const int64_t total = concat_tokens_dev(ctx, d_src.p, n_src, d_off.p, d_len.p, n_tok, d_tok.p, d_sep.p, n_items, o_out.p, out_cap);
    *out_len = total;
    if (total > out_cap) throw Error(PCB_E_CAPACITY, "the output needs " + std::to_string(total) + " bytes");
    o_out.commit(total);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- f2: tokenizers (ingest.cu)

int pcb_text_open(pcb_ctx *ctx, int mem, const uint8_t *bytes, int64_t n_bytes, pcb_text **out, int64_t *n_lines, int64_t *n_exotic) {
  if (out) *out = nullptr;
  return guarded(ctx, [&] {
    if (!out || n_bytes < 0 || (n_bytes && !bytes)) throw Error(PCB_E_INPUT, "bad arguments");
    DevBuf<uint8_t> own;
    if (mem != PCB_MEM_DEVICE) {
      own.alloc(ctx, (size_t)n_bytes);
      if (n_bytes) PCB_CUDA(cudaMemcpyAsync(own.p, bytes, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    int64_t exotic = 0;
    pcb_text *x = text_open_dev(ctx, bytes, std::move(own), n_bytes, &exotic);
    int64_t nl, nnz; int32_t a, b, c;
    text_sizes(x, &nl, &a, &nnz, &b, &c);
    if (n_lines) *n_lines = nl;
    if (n_exotic) *n_exotic = exotic;
    *out = x;
  }, false);       // text handles outlive the call: no workspace arena
}

void pcb_text_close(pcb_ctx *ctx, pcb_text *text) {
  if (!ctx || !text) return;
  cudaSetDevice(ctx->device);
  text_free(text);
}

int pcb_fastq_records(pcb_ctx *ctx, pcb_text *fastq, int32_t *n_records, int32_t *max_len, int32_t *max_id_len, int64_t *n_bad) {
  return guarded(ctx, [&] {
    if (!fastq) throw Error(PCB_E_INPUT, "no text");
    int32_t n = 0, ml = 0, mi = 0; int64_t bad = 0;
    fastq_records_dev(ctx, fastq, &n, &ml, &mi, &bad);
    if (n_records) *n_records = n;
    if (max_len) *max_len = ml;
    if (max_id_len) *max_id_len = mi;
    if (n_bad) *n_bad = bad;
  }, false);       // text handles outlive the call: no workspace arena
}

int pcb_fastq_fetch(pcb_ctx *ctx, pcb_text *fastq, int mem, int32_t stride, uint8_t *seq, uint8_t *qual, int32_t *length,
                    int64_t *id_off, int32_t *id_len) {
  return guarded(ctx, [&] {
    if (!fastq) throw Error(PCB_E_INPUT, "no text");
    int64_t nl, nnz; int32_t R, b, c;
    text_sizes(fastq, &nl, &R, &nnz, &b, &c);
    if (R < 0) throw Error(PCB_E_STATE, "pcb_fastq_fetch before pcb_fastq_records");
    if (stride <= 0) throw Error(PCB_E_INPUT, "bad stride");
    size_t cells = (size_t)R * stride;
    OutArr<uint8_t> o_seq(ctx, mem, seq, cells), o_qual(ctx, mem, qual, cells);
    OutArr<int32_t> o_len(ctx, mem, length, R), o_idl(ctx, mem, id_len, R);
    OutArr<int64_t> o_ido(ctx, mem, id_off, R);
    if (R > 0 && (!o_seq.p || !o_qual.p || !o_len.p || !o_idl.p || !o_ido.p)) throw Error(PCB_E_INPUT, "all five outputs are required");
    fastq_fetch_dev(ctx, fastq, stride, o_seq.p, o_qual.p, o_len.p, o_ido.p, o_idl.p);
    o_seq.commit(cells); o_qual.commit(cells); o_len.commit(R); o_idl.commit(R); o_ido.commit(R);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }, false);       // text handles outlive the call: no workspace arena
}

int pcb_rank_strings(pcb_ctx *ctx, const pcb_text *text, int mem, const int64_t *off, const int32_t *len, int32_t n, int32_t max_len,
                     int32_t *rank) {
  return guarded(ctx, [&] {
    if (!text || n < 0 || max_len < 0) throw Error(PCB_E_INPUT, "bad arguments");
    InArr<int64_t> d_off(ctx, mem, off, n);
    InArr<int32_t> d_len(ctx, mem, len, n);
    OutArr<int32_t> o_rank(ctx, mem, rank, n);
    rank_strings_dev(ctx, text, d_off.p, d_len.p, n, max_len, o_rank.p);
    o_rank.commit(n);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }, false);       // text handles outlive the call: no workspace arena
}

int pcb_geno_join(pcb_ctx *ctx, pcb_text *geno, const pcb_text *fastq, int mem, const int64_t *id_off, const int32_t *id_len, int32_t R,
                  int64_t *nnz, int32_t *n_mut, int64_t counters[4]) {
  return guarded(ctx, [&] {
    if (!geno || !fastq || R < 0 || !counters) throw Error(PCB_E_INPUT, "bad arguments");
    InArr<int64_t> d_off(ctx, mem, id_off, R);
    InArr<int32_t> d_len(ctx, mem, id_len, R);
    int64_t z = 0; int32_t m = 0;
    geno_join_dev(ctx, geno, fastq, d_off.p, d_len.p, R, &z, &m, counters);
    if (nnz) *nnz = z;
    if (n_mut) *n_mut = m;
  }, false);       // text handles outlive the call: no workspace arena
}

int pcb_geno_fetch(pcb_ctx *ctx, pcb_text *geno, int mem, int32_t *geno_ptr, int32_t *geno_mut, int32_t *geno_q, int64_t *mut_off,
                   int32_t *mut_len) {
  return guarded(ctx, [&] {
    if (!geno) throw Error(PCB_E_INPUT, "no text");
    int64_t nl, nnz; int32_t a, M, R;
    text_sizes(geno, &nl, &a, &nnz, &M, &R);
    if (nnz < 0 || R < 0) throw Error(PCB_E_STATE, "pcb_geno_fetch before pcb_geno_join");
    OutArr<int32_t> o_ptr(ctx, mem, geno_ptr, (size_t)R + 1), o_mut(ctx, mem, geno_mut, nnz), o_q(ctx, mem, geno_q, nnz), o_ml(ctx, mem, mut_len, M);
    OutArr<int64_t> o_mo(ctx, mem, mut_off, M);
    if (!o_ptr.p || (nnz > 0 && (!o_mut.p || !o_q.p)) || (M > 0 && (!o_ml.p || !o_mo.p))) throw Error(PCB_E_INPUT, "all five outputs are required");
    geno_fetch_dev(ctx, geno, o_ptr.p, o_mut.p, o_q.p, o_mo.p, o_ml.p);
    o_ptr.commit((size_t)R + 1); o_mut.commit(nnz); o_q.commit(nnz); o_ml.commit(M); o_mo.commit(M);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }, false);       // text handles outlive the call: no workspace arena
}

int pcb_uptag_join(pcb_ctx *ctx, pcb_text *uptag, const pcb_text *fastq, int mem, const int64_t *id_off, const int32_t *id_len,
                   int32_t R, int64_t *n_records, int32_t *n_tags, int64_t counters[4]) {
  return guarded(ctx, [&] {
    if (!uptag || !fastq || R < 0 || !counters) throw Error(PCB_E_INPUT, "bad arguments");
    InArr<int64_t> d_off(ctx, mem, id_off, R);
    InArr<int32_t> d_len(ctx, mem, id_len, R);
    int64_t nr = 0; int32_t nt = 0;
    uptag_join_dev(ctx, uptag, fastq, d_off.p, d_len.p, R, &nr, &nt, counters);
    if (n_records) *n_records = nr;
    if (n_tags) *n_tags = nt;
  }, false);       // text handles outlive the call: no workspace arena
}

int pcb_uptag_fetch(pcb_ctx *ctx, pcb_text *uptag, int mem, int32_t *up_id, int64_t *tag_off, int32_t *tag_len) {
  return guarded(ctx, [&] {
    if (!uptag) throw Error(PCB_E_INPUT, "no text");
    int64_t nl, nnz; int32_t a, M, R;
    text_sizes(uptag, &nl, &a, &nnz, &M, &R);
    if (R < 0 || M < 0) throw Error(PCB_E_STATE, "pcb_uptag_fetch before pcb_uptag_join");
    OutArr<int32_t> o_up(ctx, mem, up_id, R), o_tl(ctx, mem, tag_len, M);
    OutArr<int64_t> o_to(ctx, mem, tag_off, M);
    if ((R > 0 && !o_up.p) || (M > 0 && (!o_tl.p || !o_to.p))) throw Error(PCB_E_INPUT, "all three outputs are required");
    uptag_fetch_dev(ctx, uptag, o_up.p, o_to.p, o_tl.p);
    o_up.commit(R); o_tl.commit(M); o_to.commit(M);
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }, false);       // text handles outlive the call: no workspace arena
}

}  // extern "C"
