// common.cuh -- context, error handling, stream-ordered device buffers, launch helper.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pacybara_b200.h"

namespace pcb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define PCB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      throw pcb::Error(PCB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + \
                                       __FILE__ + ":" + std::to_string(__LINE__) + ")");       \
  } while (0)

}  // namespace pcb

// Result of the last pcb_edit_pairs call, kept on the device until fetched.
struct pcb_edges {
  int32_t *ei = nullptr, *ej = nullptr, *ed = nullptr;
  int64_t n = 0;
  bool valid = false;
};

struct pcb_text;     // a text file on the device (ingest.cu)
struct pcb_shard_state;     // what pcb_shard_pairs leaves for pcb_shard_merge (abi.cu)

// packed length of a barcode that does not fit the 2-bit planes (non-ACGT character, more than 64 bases, empty): it is
// bucketed by the exact bytes (planes = 128-bit hash) and takes no part in the pair search
#define PCB_LEN_OPAQUE 127

struct pcb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // host -> device copies that overlap the first stages (abi.cu)
  cudaEvent_t ev_alloc = nullptr, ev_copied = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t stage_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int stage_id[8] = {0};      // stat index the interval ENDING at mark k is charged to
  int n_marks = 0;
  std::string err;
  int64_t stats[PCB_STAT_COUNT] = {0};
  int replay_shape = PCB_REPLAY_AUTO;      // pcb_set_option(PCB_OPT_REPLAY_SHAPE)
  int huge_reads = 256;                    // PCB_OPT_HUGE_READS: components above it are replayed by the whole device (merge_huge.cuh)
  int huge_buckets = 16;                   // PCB_OPT_HUGE_BUCKETS: ... or with more than this many buckets (and more than 128 reads); 0 = off
  int huge_cap = 1024;                     // PCB_OPT_HUGE_CAP: slots of one warp's scratch table in that tier
  int huge_window = 32768;                 // PCB_OPT_HUGE_WINDOW: rows per round
  int huge_warps_per_sm = 40;
  int small_pad = 0;                       // (measurement aid) extra dynamic shared memory per block of the small kernel: lowers its occupancy
  int small_warps = 1;                     // PCB_OPT_SMALL_WARPS: warps per block of the small-component replay kernel (1, 2 or 4)
  int seed_bins_log2 = 26;                 // PCB_OPT_SEED_BINS_LOG2: the seed length is capped so that the index has at most 2^this bins
  int probe_classes = 1;                   // PCB_OPT_PROBE_CLASSES: split long bins by target class (probe.cuh: owns_pair); 0 = parity rule only
  int probe_staging = 0;                   // PCB_OPT_PROBE_STAGING: long bins through shared memory by TMA bulk copies (probe.cuh)
  bool medium_attr_set = false;            // the medium replay kernel's shared-memory opt-in (merge.cu)
  int trace = 0;                           // pcb_set_option(PCB_OPT_TRACE): host-side lap times on stderr
  // Workspace arena: the temporaries of one call are carved out of one long-lived allocation with a bump pointer and
  // all dropped when the call returns.  The stream-ordered pool is fast on average, but its occasional slow paths
  // (a block that has to be split, remapped or grown: 10..50 ms for the 0.5 GB buffers of a 2*10^7-read call, seen in
  // profiles/r02f) sit on the critical path of a step; the arena has none once it has seen the largest call.
  char *arena = nullptr;
  size_t arena_cap = 0, arena_off = 0, arena_need = 0;
  bool arena_on = false;
  pcb_edges edges;
  pcb_shard_state *shard = nullptr;
};

namespace pcb {

// Stream-ordered allocation out of the device's default pool (kept warm: the release
// threshold is raised at context creation, so steady-state steps never reach the OS).
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  bool in_arena = false;
  DevBuf() {}
  DevBuf(pcb_ctx *ctx, size_t count) { alloc(ctx, count); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s), in_arena(o.in_arena) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; s = o.s; in_arena = o.in_arena; o.p = nullptr; o.n = 0; }
    return *this;
  }
  void alloc(pcb_ctx *ctx, size_t count) {
    release();
    s = ctx->stream; n = count;
    const size_t bytes = (((count ? count : 1) * sizeof(T)) + 255) & ~(size_t)255;
    if (ctx->arena_on) {
      ctx->arena_need += bytes;
      if (ctx->arena_off + bytes <= ctx->arena_cap) {
        p = (T *)(ctx->arena + ctx->arena_off);
        ctx->arena_off += bytes;
        in_arena = true;
        return;
      }
    }
    in_arena = false;
    PCB_CUDA(cudaMallocAsync((void **)&p, bytes, s));
  }
  void release() {
    if (p && !in_arena) cudaFreeAsync(p, s);
    p = nullptr; n = 0; in_arena = false;
  }
  T *detach() { T *q = p; p = nullptr; n = 0; return q; }
  ~DevBuf() { release(); }
  operator T *() const { return p; }
};

// development aid (PCB_OPT_TRACE): where the HOST spends its time inside a stage
struct HostTrace {
  pcb_ctx *ctx; const char *what;
  std::chrono::steady_clock::time_point t;
  HostTrace(pcb_ctx *c, const char *w) : ctx(c), what(w), t(std::chrono::steady_clock::now()) {}
  void lap(const char *name) {
    if (!ctx->trace) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[pcb trace] %s / %s: %.3f ms\n", what, name, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

inline unsigned grid_for(size_t n, unsigned block) {
  size_t g = (n + block - 1) / block;
  return (unsigned)(g ? g : 1);
}

#define PCB_LAUNCH(ctx, kernel, grid, block, smem, ...)                 \
  do {                                                                  \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);    \
    (ctx)->stats[PCB_STAT_KERNEL_LAUNCHES]++;                           \
    PCB_CUDA(cudaGetLastError());                                       \
  } while (0)

// Stage timing: stage_begin() once, stage_end(stat) after each stage; stage_collect() after the stream
// has drained turns the event intervals into stats[stat] (ns).
inline void stage_begin(pcb_ctx *ctx, int first = PCB_STAT_T_BUCKET_NS, int last = PCB_STAT_T_SCATTER_NS) {
  ctx->n_marks = 0;
  for (int s = first; s <= last; s++) ctx->stats[s] = 0;      // the stages this call is going to time
  if (ctx->stage_ev[0]) { cudaEventRecord(ctx->stage_ev[0], ctx->stream); ctx->n_marks = 1; }
}
inline void stage_end(pcb_ctx *ctx, int stat) {
  if (ctx->n_marks == 0 || ctx->n_marks >= 8) return;
  cudaEventRecord(ctx->stage_ev[ctx->n_marks], ctx->stream);
  ctx->stage_id[ctx->n_marks] = stat;
  ctx->n_marks++;
}
inline void stage_collect(pcb_ctx *ctx) {
  for (int k = 1; k < ctx->n_marks; k++) {
    float ms = 0.f;
    if (ctx->stage_id[k] >= 0 && cudaEventElapsedTime(&ms, ctx->stage_ev[k - 1], ctx->stage_ev[k]) == cudaSuccess)
      ctx->stats[ctx->stage_id[k]] += (int64_t)(ms * 1e6);
  }
  ctx->n_marks = 0;
}

template <typename T>
inline T read_scalar(pcb_ctx *ctx, const T *dev) {
  T v;
  PCB_CUDA(cudaMemcpyAsync(&v, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return v;
}

// ---- primitives (prims.cu)
void exclusive_scan_u32(pcb_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n);   // out may alias in
void exclusive_scan_i64(pcb_ctx *ctx, const int64_t *in, int64_t *out, size_t n);
// Stable LSD radix sort of (key, val) pairs on the low `bits` bits of the key.  Both buffer pairs
// must hold n items; returns 0 if the result is in (k0, v0), 1 if in (k1, v1).
int radix_sort_pairs(pcb_ctx *ctx, uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, size_t n, int bits);
void fill_i32(pcb_ctx *ctx, int32_t *p, size_t n, int32_t v);
void fill_i64(pcb_ctx *ctx, int64_t *p, size_t n, int64_t v);
void iota_u32(pcb_ctx *ctx, uint32_t *p, size_t n);
inline int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) b++;
  return b;
}

// ---- stages on device memory (bucket.cu, edit_pairs.cu, merge.cu)
struct PackedBarcodes {      // planes of n barcodes: x = plane L (bit 0 of each base), y = plane H
  DevBuf<ulonglong2> planes;
  DevBuf<int32_t> len;
  DevBuf<uint64_t> other;    // optional: positions holding a character that is no base (f3 queries); planes are 0 there
  int32_t n = 0, min_len = 0, max_len = 0;   // lengths over the barcodes that fit the planes (len <= 64)
  int32_t n_opaque = 0;                      // barcodes kept as opaque strings (len == PCB_LEN_OPAQUE, planes = hash)
  uint32_t len_hist[66] = {0};   // barcodes per length, when the packing counted them (has_hist)
  bool has_hist = false;
};
void pack_launch(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, const int32_t *rows, int32_t n, int32_t stride,
                 PackedBarcodes &out, bool allow_other, DevBuf<int32_t> &flags);
void pack_finish(PackedBarcodes &out, const int32_t *h_flags, bool allow_other);
// ASCII rows (optionally gathered through `rows`) -> planes; throws PCB_E_INPUT on bad input
// allow_other: characters outside ACGT are recorded in out.other instead of being refused
void pack_barcodes(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, const int32_t *rows, int32_t n,
                   int32_t stride, PackedBarcodes &out, bool allow_other = false);

struct BucketOut {           // all device pointers, caller-owned
  int32_t *bucket_of_read, *bucket_ptr, *bucket_reads;
  uint8_t *bucket_qsum;      // may be null
};
int32_t bucket_reads_dev(pcb_ctx *ctx, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int32_t R,
                         int32_t stride, const BucketOut &out, PackedBarcodes *unique_out);
// the same in two halves, for callers that have more to launch before they synchronise the stream
struct BucketPending {
  int32_t R = 0;
  PackedBarcodes pk;
  DevBuf<int32_t> flags, counters, mid, big;
  int32_t h_flags[4 + 66];
  int32_t h_counters[4];
};
void bucket_launch(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, int32_t R, int32_t stride, const BucketOut &out,
                   PackedBarcodes *unique_out, BucketPending &st);
int32_t bucket_finish(pcb_ctx *ctx, const BucketOut &out, PackedBarcodes *unique_out, BucketPending &st);

void edit_pairs_dev(pcb_ctx *ctx, const PackedBarcodes &bc, int32_t k, int32_t q_begin, int32_t q_end, int32_t t_begin = 0, int32_t t_end = -1);
// f3: nearest library barcodes of every query within k; the (query, row, d) rows stay in ctx->edges
void match_barcodes_dev(pcb_ctx *ctx, const PackedBarcodes &library, const PackedBarcodes &queries, int32_t k, int32_t *best_d,
                        int32_t *n_best, int32_t *first_hit);

struct MergeIn {             // device pointers
  int32_t R, U, n_mut;
  const int32_t *bucket_ptr, *bucket_reads, *lexrank, *bc_id, *up_id;
  const int32_t *geno_ptr, *geno_mut, *geno_q, *mut_rank;
  int64_t n_pairs;
  const int32_t *pair_q, *pair_r, *pair_d, *pair_ed;
  int32_t max_diff, min_qual, min_matches;
  double min_jaccard;
  int32_t shard_rank, shard_count;
  // on-demand distances (see merge_core.h); planes == nullptr / compute_k < 0: the pair list is all there is
  const ulonglong2 *planes = nullptr;
  const int32_t *blen = nullptr;
  int32_t compute_k = -1, wide = 0;
  int64_t nnz = -1;            // number of genotype entries if the caller already knows it (saves a round trip)
  bool pairs_sorted = false;   // pair_q < pair_r everywhere and the pairs ascend by (q, r): how the pair search delivers them
  const int32_t *bucket_of_read = nullptr;   // [R], if the caller has it (it is derived from the CSR otherwise)
  bool shard_layout = false;   // call regions laid out by component label, identically on every shard
};
struct MergeOut {            // device pointers, caller-owned
  int32_t *read_root, *read_pos, *row_size;
  int64_t *row_key;
  int32_t *row_bc, *row_up;
  int64_t *row_call_off;
  int32_t *row_call_n, *call_mut;
};
void merge_dev(pcb_ctx *ctx, const MergeIn &in, const MergeOut &out);
int64_t concat_tokens_dev(pcb_ctx *ctx, const uint8_t *src, int64_t n_src, const int64_t *tok_off, const int32_t *tok_len, int64_t n_tok,
                          const int32_t *item_tok, const uint8_t *item_sep, int64_t n_items, uint8_t *out, int64_t out_cap);
void merge_libraries_dev(pcb_ctx *ctx, int32_t n1, const int32_t *vbc1, const int32_t *geno1, const int32_t *size1, int32_t n2,
                         const int32_t *vbc2, const int32_t *geno2, const int32_t *size2, int32_t n_ids, int32_t *size1_out,
                         int32_t *target2, int32_t *klass2);
void tag_rows_dev(pcb_ctx *ctx, int32_t n_rows, const int32_t *row_bc, const int32_t *row_up, const int32_t *row_size,
                  int32_t n_bc_ids, int32_t n_up_ids, double min_ratio, uint8_t *collision, uint8_t *up_collision, uint8_t *usable);


// ---- f2: text tokenizers (ingest.cu); all array arguments are device pointers
pcb_text *text_open_dev(pcb_ctx *ctx, const uint8_t *bytes, DevBuf<uint8_t> &&own, int64_t n, int64_t *n_exotic);
void fastq_records_dev(pcb_ctx *ctx, pcb_text *x, int32_t *n_records, int32_t *max_len, int32_t *max_id_len, int64_t *n_bad);
void fastq_fetch_dev(pcb_ctx *ctx, pcb_text *x, int32_t stride, uint8_t *seq, uint8_t *qual, int32_t *length, int64_t *id_off,
                     int32_t *id_len);
void rank_strings_dev(pcb_ctx *ctx, const pcb_text *x, const int64_t *off, const int32_t *len, int32_t n, int32_t max_len, int32_t *rank);
void geno_join_dev(pcb_ctx *ctx, pcb_text *g, const pcb_text *fq, const int64_t *id_off, const int32_t *id_len, int32_t R,
                   int64_t *nnz, int32_t *n_mut, int64_t counters_out[4]);
void geno_fetch_dev(pcb_ctx *ctx, pcb_text *g, int32_t *geno_ptr, int32_t *geno_mut, int32_t *geno_q, int64_t *mut_off, int32_t *mut_len);
void uptag_join_dev(pcb_ctx *ctx, pcb_text *u, const pcb_text *fq, const int64_t *id_off, const int32_t *id_len, int32_t R,
                    int64_t *n_records, int32_t *n_tags, int64_t counters_out[4]);
void uptag_fetch_dev(pcb_ctx *ctx, pcb_text *u, int32_t *up_id, int64_t *tag_off, int32_t *tag_len);
void text_sizes(const pcb_text *x, int64_t *n_lines, int32_t *n_rec, int64_t *nnz, int32_t *n_distinct, int32_t *n_reads);
void text_free(pcb_text *x);

}  // namespace pcb
