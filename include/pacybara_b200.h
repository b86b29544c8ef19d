/*
 * pacybara_b200.h -- C ABI of the B200-native barcode-clustering hot path.
 *
 * The reference (rothlab/pacybara) has no FFI: its boundary for this path is three
 * processes connected by files.  Each entry point below replaces the
 * compute of one of those processes; the file parsing/writing stays in the host layer
 * (the Python modules of pacybara_b200), which binds this header with ctypes.
 *
 *   pcb_bucket        <- src/pacybara_precluster.py:28-39,68-76   (addRecord + output order)
 *   pcb_edit_pairs    <- src/pacybara.sh:448-468                  (seqret + bowtie2-build +
 *                        bowtie2 --all + awk + src/pacybara_calcEdits.R:128-172)
 *   pcb_merge         <- src/pacybara_cluster.py:385-477,515-537  (seed + main clustering,
 *                        consensus genotype / barcode)
 *   pcb_cluster_reads <- the three of them back to back, nothing leaving the device
 *   pcb_shard_pairs, pcb_shard_merge <- (multi-GPU) one rank's share of a step: the two calls around the edge all-gather
 *   pcb_compact_shard <- (multi-GPU) the part of a merge result one rank owns, compacted
 *   pcb_concat_tokens <- src/pacybara_cluster.py:546-567 (writeOutput: the rows of clusters.csv.gz as text)
 * and, either side of the path:
 *   pcb_tag_rows       <- src/pacybara_translate.R:58-68, src/pacybara_softfilter.R:46-63      (f1)
 *   pcb_text_open, pcb_fastq_*, pcb_geno_*, pcb_uptag_*, pcb_rank_strings
 *                      <- the line loops of src/pacybara_precluster.py:47-63 and
 *                         src/pacybara_cluster.py:302-331, makePID's string order :152-157      (f2)
 *   pcb_match_barcodes <- src/barseq_caller.R:57,242 (yogiseq::new.bc.matcher)                 (f3)
 *   pcb_merge_libraries <- src/pacybara_merge.R:48-101                                         (f4)
 *
 * Conventions
 *   - every function returns 0 on success or a PCB_E_* code; pcb_last_error() gives the text;
 *   - `mem` says where ALL array arguments of that call live: PCB_MEM_HOST (the library
 *     copies in and out on its own stream) or PCB_MEM_DEVICE (pointers into the memory of the
 *     context's device; nothing is copied, results stay there);
 *   - the caller owns every array and sizes it as documented; scalars returned through
 *     pointers are always host memory;
 *   - barcodes are ASCII strings, one per row of a [n, stride] uint8 matrix, len[i] <= stride characters.  Those of
 *     1..64 bases over {A,C,G,T} are packed into 2-bit planes and take part in everything.  Any other barcode (another
 *     character such as N, more than 64 characters, or none) is kept as the opaque string the reference treats every
 *     barcode as (src/pacybara_precluster.py:28-39): it is bucketed with its exact copies (by a 128-bit hash of its
 *     bytes) and clusters like any bucket, but takes no part in the pair search -- no distance is reported for it.
 *     Only a negative length or one beyond `stride` is refused (PCB_E_INPUT);
 *   - one context per GPU and per host thread; calls on one context are serialised.
 *
 * There is no CPU implementation behind this ABI: without a CUDA device every entry point
 * that computes fails with PCB_E_CUDA.
 */
#ifndef PACYBARA_B200_H
#define PACYBARA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCB_ABI_VERSION 3

enum {
  PCB_OK = 0,
  PCB_E_CUDA = 1,      /* CUDA runtime / no device */
  PCB_E_INPUT = 2,     /* malformed input (a length outside the matrix, bad parameters, ...) */
  PCB_E_CAPACITY = 3,  /* caller buffer too small; the needed size is reported */
  PCB_E_STATE = 4,     /* call order (e.g. fetch before compute) */
  PCB_E_UPTAG = 5      /* a clustered read has no uptag (the reference's KeyError, :519) */
};

enum { PCB_MEM_HOST = 0, PCB_MEM_DEVICE = 1 };

#define PCB_FILL INT32_MIN          /* per-read outputs of components another shard owns */
#define PCB_NEEDS_ALIGNMENT (-2)    /* consensus the reference hands to muscle (:523-530) */

/* counters readable after a call (pcb_stat) */
enum {
  PCB_STAT_PAIRS_SCORED = 0,   /* candidate pairs that entered the Myers recurrence */
  PCB_STAT_CELLS = 1,          /* DP cells the recurrence computed for them (rows x columns done; an early exit stops short) */
  PCB_STAT_EDGES = 2,          /* unique pairs with dist <= k */
  PCB_STAT_SEED_LEN = 3,       /* q of the pigeonhole seeds used */
  PCB_STAT_KERNEL_LAUNCHES = 4,/* kernels launched by this context so far */
  PCB_STAT_COMPONENTS = 5,
  PCB_STAT_LINKS = 6,          /* accepted merges in the main clustering */
  PCB_STAT_TESTS = 7,          /* cluster-pair tests evaluated in the main clustering */
  PCB_STAT_BUCKETS = 8,
  PCB_STAT_PROBE_NS = 9,       /* device time of the last distance kernel, ns (CUDA events) */
  /* device time of the stages of the last call, ns, from CUDA events on the context's stream */
  PCB_STAT_T_BUCKET_NS = 10,   /* pack + table + numbering + read grouping */
  PCB_STAT_T_INDEX_NS = 11,    /* seed index (count, scan, fill) */
  PCB_STAT_T_HITS_NS = 12,     /* hit sort + de-duplication */
  PCB_STAT_T_PREP_NS = 13,     /* merge preprocessing: components, grouping, adjacency rows, visiting lists */
  PCB_STAT_T_REPLAY_NS = 14,   /* the replay kernels (they write the outputs in the caller's numbering themselves) */
  PCB_STAT_T_SCATTER_NS = 15,  /* what follows the replay: status round trip, late components */
  PCB_STAT_ROWS = 16,          /* rows of the last pcb_cluster_reads with PCB_CLUSTER_COMPACT_ROWS */
  PCB_STAT_CALLS = 17,         /* consensus genotype entries of those rows */
  PCB_STAT_CANDIDATES = 18,    /* candidate pairs the seed index produced (before the composition / length filters) */
  PCB_STAT_HUGE_COMPONENTS = 19, /* components replayed by the whole device (giant tier) in the last merge */
  PCB_STAT_HUGE_ROUNDS = 20,   /* rounds of deterministic reservations their main loop took */
  PCB_STAT_COUNT = 21
};

typedef struct pcb_ctx pcb_ctx;

int pcb_abi_version(void);
/* Text of the last error on this context (or of the last failed pcb_ctx_create when ctx is NULL). */
const char *pcb_last_error(const pcb_ctx *ctx);
int pcb_ctx_create(int device, pcb_ctx **out);
void pcb_ctx_destroy(pcb_ctx *ctx);
int64_t pcb_stat(const pcb_ctx *ctx, int which);
/*
 * Options of a context.  PCB_OPT_REPLAY_SHAPE: which launch shape replays the components in pcb_merge /
 * pcb_cluster_reads -- PCB_REPLAY_AUTO (default: one warp per component with one lane per read for components of
 * at most 32 reads, the generic warp kernel up to PCB_OPT_HUGE_READS reads, the whole device beyond),
 * PCB_REPLAY_THREAD (one thread per component), PCB_REPLAY_WARP (the generic warp kernel for everything) or
 * PCB_REPLAY_HUGE (every component through the giant tier: seed step per bucket, main loop in rounds of deterministic
 * reservations, consensus per cluster).  All shapes give identical results; the knob exists for measurements and for
 * the parity tests, which run every shape.
 * Giant tier: PCB_OPT_HUGE_READS = reads above which a component is replayed by the whole device (default 256);
 * PCB_OPT_HUGE_CAP = slots of a warp's scratch table, a power of two >= 16 (default 1024; units that need more wait
 * for a few warps with scratch sized on demand); PCB_OPT_HUGE_WINDOW = rows per round (default 32768).
 */
enum { PCB_OPT_REPLAY_SHAPE = 0, PCB_OPT_TRACE = 1 /* 1: host-side lap times of the stages on stderr (development aid) */,
       PCB_OPT_HUGE_READS = 2, PCB_OPT_HUGE_CAP = 3, PCB_OPT_HUGE_WINDOW = 4,
       PCB_OPT_PROBE_STAGING = 5 /* 1: the pair search walks long bins through shared memory (cp.async.bulk + mbarrier)
                                    instead of streaming loads; same results, kept for the measurement in profiles/ */,
       PCB_OPT_PROBE_CLASSES = 6 /* 1 (default): long bins of the seed index are grouped by target class so that a probe
                                    skips most entries of pairs it does not own; 0: whole bins, parity rule */,
       PCB_OPT_SEED_BINS_LOG2 = 7 /* the seed length of the pair search is capped so that its index has at most 2^value
                                     bins (8..28, default 26): shorter seeds = a smaller table and longer bins */,
       PCB_OPT_SMALL_WARPS = 8 /* warps per block of the small-component replay kernel: 1, 2 or 4 */,
       PCB_OPT_HUGE_BUCKETS = 9 /* a component of more than 128 reads also goes to the giant tier when it has more than this
                                   many buckets (default 16); 0 = only the read count decides */ };
enum { PCB_REPLAY_AUTO = 0, PCB_REPLAY_THREAD = 1, PCB_REPLAY_WARP = 2, PCB_REPLAY_HUGE = 3 };
int pcb_set_option(pcb_ctx *ctx, int option, int64_t value);
/* Block until everything queued on the context's stream has finished. */
int pcb_sync(pcb_ctx *ctx);
/* The context's cudaStream_t, for callers that time with their own CUDA events. */
void *pcb_stream(pcb_ctx *ctx);

/*
 * a1 -- exact-barcode bucketing (src/pacybara_precluster.py:28-39, 68-76).
 *   seq, qual      [R, stride] ASCII (qual = Phred+33), len [R]
 * outputs (caller-sized for the worst case U = R):
 *   bucket_of_read [R]        bucket index of each read; buckets are numbered in first-seen order
 *   n_buckets      host int   U
 *   bucket_ptr     [R+1]      CSR over buckets (first U+1 entries valid)
 *   bucket_reads   [R]        reads of each bucket in arrival order
 *   bucket_qsum    [R, stride] per-position min(93, sum of Phred) of each bucket (first U rows)
 * The barcode of bucket b is the row seq[bucket_reads[bucket_ptr[b]]].
 */
int pcb_bucket(pcb_ctx *ctx, int mem, const uint8_t *seq, const uint8_t *qual, const int32_t *len,
               int32_t R, int32_t stride, int32_t *bucket_of_read, int32_t *n_buckets,
               int32_t *bucket_ptr, int32_t *bucket_reads, uint8_t *bucket_qsum);

/*
 * a2+a3 -- every unordered pair of distinct barcodes with Levenshtein distance <= k
 * (replaces bowtie2 all-vs-all + pacybara_calcEdits.R; dist := unit-cost Levenshtein).
 *   seq [U, stride], len [U]: the unique barcodes;  0 <= k <= 8
 *   q_begin, q_end / t_begin, t_end: every unordered pair has exactly one "query" side and one "target" side;
 *     this call scores the pairs whose query lies in [q_begin, q_end) AND whose target lies in [t_begin, t_end).
 *     Disjoint ranges that cover [0, U) on either side (the other left at [0, U)) produce disjoint edge sets whose
 *     union is the full result (multi-GPU sharding).  Only the targets of the range are indexed, so sharding by
 *     target also divides the index build.
 *   n_edges: host int64, number of pairs found.  The pairs stay in the context until
 *   pcb_edit_pairs_fetch copies them out: ei < ej, sorted by (ei, ej), ed = distance.
 */
int pcb_edit_pairs(pcb_ctx *ctx, int mem, const uint8_t *seq, const int32_t *len, int32_t U, int32_t stride,
                   int32_t k, int32_t q_begin, int32_t q_end, int32_t t_begin, int32_t t_end, int64_t *n_edges);
int pcb_edit_pairs_fetch(pcb_ctx *ctx, int mem, int32_t *ei, int32_t *ej, int32_t *ed, int64_t cap);
/* Sort n edges in place by (ei, ej): puts the concatenation of several shards' edge lists
 * (after the all-gather) back into the canonical row order the merge visits them in. */
int pcb_sort_edges(pcb_ctx *ctx, int mem, int32_t *ei, int32_t *ej, int32_t *ed, int64_t n, int32_t U);

/*
 * a4..a13 -- seed clustering, greedy genotype-aware merge and consensus
 * (src/pacybara_cluster.py:385-455, 464-477, 515-537).
 *
 * inputs
 *   bucket_ptr [U+1], bucket_reads [R]   buckets and their reads in arrival order
 *   lexrank [R]     rank of the read id in string order (makePID, :152-157)
 *   bc_id [R]       id of the read's barcode string;  up_id [R] id of its uptag string, -1 if the
 *                   read has none, or NULL when no uptag file is given (:537)
 *   geno_ptr [R+1], geno_mut, geno_q [nnz]  genotype calls per read, distinct mutation ids within
 *                   a read, input order (nnz = geno_ptr[R], passed so that device arrays need no round trip);
 *                   mut_rank [n_mut] rank in (digit key, string) order (:475), a permutation of 0..n_mut-1
 *   pair_q, pair_r, pair_d, pair_ed [n_pairs]  unique unordered bucket pairs in visiting order:
 *                   (q, r) direction of the row that introduced the pair in the pass that visits it,
 *                   d = distance the lookup table ends with (:266), ed = visiting pass 1..maxDiff or 0
 *                   for lookup-only pairs (host helper: hostdata.dedupe_ed_rows / pairs_from_edges)
 *   bucket_seq [U, stride], bucket_len [U], on_demand_k  optional (NULL / -1 to disable): when given,
 *                   a distance the transitive rule (:442) looks up and the pair list does not hold is
 *                   scored on the spot from the barcodes and used if <= on_demand_k.  With the pair
 *                   list complete up to max_diff and on_demand_k = 2*max_diff this gives exactly the
 *                   clusters of the full table (lookups never leave a connected component).  The
 *                   drop-in pacybara_cluster.py never sets it: there the EDFILE is the only authority.
 *   shard_rank / shard_count  a connected component belongs to shard (smallest bucket of the component mod
 *                   shard_count); entries of reads owned by other shards are left at PCB_FILL
 *   flags           PCB_MERGE_PAIRS_SORTED: pair_q < pair_r everywhere and the pairs ascend by (q, r) -- the order
 *                   pcb_edit_pairs delivers and pcb_sort_edges restores; the adjacency and the visiting lists then need
 *                   no sort.  PCB_MERGE_SHARD_LAYOUT: lay the call regions out by component label, identically on
 *                   every shard, so that shards can be combined with an element-wise max (costs a pass over all reads;
 *                   without it a row's calls sit wherever the component found room: row_call_off / row_call_n / the
 *                   calls are reproducible, the layout of call_mut is not)
 * outputs (all [R] unless noted; a "row" is a cluster or an unclustered read and is named by
 * its representative read = first member)
 *   read_root       representative read of the row the read belongs to
 *   read_pos        position of the read in the row's member list (reference list order)
 *   row_size        members (valid at representative reads, 0 at other reads)
 *   row_key         int64: creation order of the surviving cluster id, -1 for singletons
 *   row_bc, row_up  consensus barcode / uptag id, or PCB_NEEDS_ALIGNMENT
 *   row_call_off (int64), row_call_n, call_mut [nnz]  consensus / singleton genotype
 */
#define PCB_MERGE_PAIRS_SORTED 1
#define PCB_MERGE_SHARD_LAYOUT 2
int pcb_merge(pcb_ctx *ctx, int mem, int32_t R, int32_t U,
              const int32_t *bucket_ptr, const int32_t *bucket_reads, const int32_t *lexrank,
              const int32_t *bc_id, const int32_t *up_id,
              const int32_t *geno_ptr, const int32_t *geno_mut, const int32_t *geno_q, int64_t nnz,
              const int32_t *mut_rank, int32_t n_mut,
              int64_t n_pairs, const int32_t *pair_q, const int32_t *pair_r, const int32_t *pair_d,
              const int32_t *pair_ed,
              int32_t max_diff, int32_t min_qual, double min_jaccard, int32_t min_matches,
              const uint8_t *bucket_seq, const int32_t *bucket_len, int32_t stride, int32_t on_demand_k,
              int32_t shard_rank, int32_t shard_count, int32_t flags,
              int32_t *read_root, int32_t *read_pos, int32_t *row_size, int64_t *row_key,
              int32_t *row_bc, int32_t *row_up, int64_t *row_call_off, int32_t *row_call_n,
              int32_t *call_mut);

/*
 * The whole hot path on one device: bucketing -> distances -> merge, with nothing returning to
 * the host in between.  bc_id is the bucket index.  Inputs as pcb_bucket plus the read-level
 * arrays of pcb_merge; outputs as pcb_bucket (bucket_of_read, n_buckets) and pcb_merge.
 * n_edges (host int64, may be NULL) receives the number of distance pairs found.  `qual` is
 * accepted for symmetry with pcb_bucket and may be NULL: the clustering never reads qualities.
 *
 * flags = 0: the pair search runs at k = max_diff (the pairs a pass can visit); the distances up
 *   to 2*max_diff that the transitive rule (:442) looks up -- always between two buckets of one
 *   connected component -- are scored on demand inside the merge.  Same clusters, far fewer
 *   candidates.
 * flags = PCB_CLUSTER_FULL_TABLE: the pair search runs at k = 2*max_diff, i.e. the table
 *   pacybara.sh:466-468 asks of calcEdits.R is materialised (fetch it with pcb_edit_pairs_fetch).
 * flags |= PCB_CLUSTER_COMPACT_ROWS: the row_* outputs hold one entry per ROW instead of one per read: the
 *   n_rows = pcb_stat(PCB_STAT_ROWS) rows in ascending order of their representative read, which row_rep[i]
 *   names (required with this flag, NULL otherwise); row_call_off[i] points into a call_mut that holds only the
 *   pcb_stat(PCB_STAT_CALLS) consensus entries, row after row.  read_root / read_pos / bucket_of_read stay per read.
 *   With host memory only those entries cross the bus (a cluster table has far fewer rows than reads).
 */
#define PCB_CLUSTER_FULL_TABLE 1
#define PCB_CLUSTER_COMPACT_ROWS 2
int pcb_cluster_reads(pcb_ctx *ctx, int mem, const uint8_t *seq, const uint8_t *qual, const int32_t *len,
                      int32_t R, int32_t stride, const int32_t *lexrank, const int32_t *up_id,
                      const int32_t *geno_ptr, const int32_t *geno_mut, const int32_t *geno_q, int64_t nnz,
                      const int32_t *mut_rank, int32_t n_mut,
                      int32_t max_diff, int32_t min_qual, double min_jaccard, int32_t min_matches, int32_t flags,
                      int32_t *bucket_of_read, int32_t *n_buckets, int64_t *n_edges,
                      int32_t *read_root, int32_t *read_pos, int32_t *row_size, int64_t *row_key,
                      int32_t *row_bc, int32_t *row_up, int64_t *row_call_off, int32_t *row_call_n,
                      int32_t *call_mut, int32_t *row_rep);

/*
 * The part of a merge result that belongs to this shard, compacted (multi-GPU: every rank keeps and ships only its own
 * components; N = 1: all reads are owned).  Inputs: the per-read outputs of pcb_merge / pcb_cluster_reads, in `mem_in`.
 * Outputs, in `mem_out` (buffers of R entries, c_call_mut of nnz):
 *   own_read / own_root / own_pos [n_own]   the reads whose read_root is not PCB_FILL, ascending, with their row and position
 *   own_bucket [n_own]                      bucket_of_read of those reads (both NULL to skip): row_bc names buckets, and
 *                                           the consensus barcode of a row is the barcode of one of its own members
 *   row_rep, c_* [n_rows]                   one entry per row in ascending order of the representative read; c_call_off
 *                                           points into c_call_mut [n_calls], which holds the rows' calls back to back
 * With mem_in = PCB_MEM_DEVICE and mem_out = PCB_MEM_HOST only the compacted entries cross the bus.
 */
int pcb_compact_shard(pcb_ctx *ctx, int mem_in, int mem_out, int32_t R, int64_t nnz,
                      const int32_t *read_root, const int32_t *read_pos, const int32_t *row_size, const int64_t *row_key,
                      const int32_t *row_bc, const int32_t *row_up, const int64_t *row_call_off, const int32_t *row_call_n,
                      const int32_t *call_mut, const int32_t *bucket_of_read,
                      int32_t *n_own, int32_t *own_read, int32_t *own_root, int32_t *own_pos, int32_t *own_bucket,
                      int32_t *n_rows, int64_t *n_calls, int32_t *row_rep, int32_t *c_size, int64_t *c_key, int32_t *c_bc,
                      int32_t *c_up, int64_t *c_call_off, int32_t *c_call_n, int32_t *c_call_mut);

/*
 * (e) One rank's share of a multi-GPU step as two calls around the one collective (pacybara_b200/multigpu.py): every rank
 * buckets the whole read set, searches ITS block of the (query groups x target groups) pair grid, all ranks all-gather
 * their edge lists (NCCL, by the caller), and every rank merges the components it owns.  Device arrays only.
 *   pcb_shard_pairs   buckets seq/len [R] (bucket_of_read [R] out, *n_buckets), then the pairs within k whose query side
 *                     lies in slice q_group of n_q_groups of [0, U) and whose target side in slice t_group of n_t_groups;
 *                     *n_edges pairs are left for pcb_edit_pairs_fetch.  The bucket arrays and the packed barcodes stay
 *                     in the context for pcb_shard_merge.
 *   pcb_shard_merge   gathered [world][3][m] int32: what the all-gather of the ranks' (ei, ej, ed) rows, padded to m
 *                     columns, left; counts [world] (HOST array) = valid columns per rank.  Sorts the edges by (i, j) and
 *                     runs pcb_merge's work for shard shard_rank of shard_count on the buckets of pcb_shard_pairs
 *                     (on_demand != 0: transitive-rule distances beyond max_diff are scored from the retained barcodes,
 *                     i.e. the search ran at k = max_diff; 0: the search ran at k = 2 * max_diff).  Outputs as pcb_merge.
 */
int pcb_shard_pairs(pcb_ctx *ctx, int mem, const uint8_t *seq, const int32_t *len, int32_t R, int32_t stride, int32_t k,
                    int32_t q_group, int32_t n_q_groups, int32_t t_group, int32_t n_t_groups, int32_t *bucket_of_read,
                    int32_t *n_buckets, int64_t *n_edges);
int pcb_shard_merge(pcb_ctx *ctx, int mem, const int32_t *gathered, int32_t world, int64_t m, const int64_t *counts,
                    const int32_t *bucket_of_read, const int32_t *lexrank, const int32_t *up_id, const int32_t *geno_ptr,
                    const int32_t *geno_mut, const int32_t *geno_q, int64_t nnz, const int32_t *mut_rank, int32_t n_mut,
                    int32_t max_diff, int32_t min_qual, double min_jaccard, int32_t min_matches, int32_t on_demand,
                    int32_t shard_rank, int32_t shard_count, int32_t flags, int64_t *n_edges, int32_t *read_root, int32_t *read_pos,
                    int32_t *row_size, int64_t *row_key, int32_t *row_bc, int32_t *row_up, int64_t *row_call_off,
                    int32_t *row_call_n, int32_t *call_mut);

/*
 * a14 -- the writer (src/pacybara_cluster.py:546-567).  A row of clusters.csv.gz is made of slices of texts the caller
 * already holds (read ids, mutation names, barcodes, uptags) and one-byte separators.
 *   src [n_src]                    the bytes those slices come from (any concatenation of the caller's texts)
 *   tok_off / tok_len [n_tok]      the slices ("tokens"): src[tok_off[t] .. tok_off[t] + tok_len[t])
 *   item_tok / item_sep [n_items]  the output, item by item: token item_tok[k], followed by the byte item_sep[k] unless 0
 *   out [out_cap], *out_len        the concatenation and its length.  PCB_E_CAPACITY (with *out_len = bytes needed) when
 *                                  it does not fit; PCB_E_INPUT for a token outside src or an item naming no token.
 * pacybara_b200/tablefmt.py lists the items of a cluster table (`vbc , ubc , id | id ... , size , mut ; mut ... \n`).
 */
int pcb_concat_tokens(pcb_ctx *ctx, int mem, const uint8_t *src, int64_t n_src, const int64_t *tok_off, const int32_t *tok_len,
                      int64_t n_tok, const int32_t *item_tok, const uint8_t *item_sep, int64_t n_items, uint8_t *out, int64_t out_cap,
                      int64_t *out_len);

/*
 * f4 -- merging two cluster libraries by virtual barcode and genotype (src/pacybara_merge.R:48-101).
 *   vbc1/geno1/size1 [n1], vbc2/geno2/size2 [n2]: per row of the first / second translated cluster table the id of its
 *   virtualBarcode (0 .. n_ids-1, one id space for both libraries), the id of its hgvsc genotype string, and its size.
 * outputs
 *   klass2 [n2]     PCB_LIBMERGE_NEW (barcode not in the first library: the row is appended), PCB_LIBMERGE_MERGE (a row
 *                   of the first library has the same barcode and genotype: sizes add up, read lists are joined) or
 *                   PCB_LIBMERGE_CONFLICT (the barcode is there with other genotypes only: appended, a collision)
 *   target2 [n2]    for merges the row of the first library the row goes into -- when several rows share barcode and
 *                   genotype, the one that is largest at that moment, first of equals (which.max) -- else -1
 *   size1_out [n1]  sizes of the first library's rows after all merges
 */
enum { PCB_LIBMERGE_NEW = 0, PCB_LIBMERGE_MERGE = 1, PCB_LIBMERGE_CONFLICT = 2 };
int pcb_merge_libraries(pcb_ctx *ctx, int mem, int32_t n1, const int32_t *vbc1, const int32_t *geno1, const int32_t *size1,
                        int32_t n2, const int32_t *vbc2, const int32_t *geno2, const int32_t *size2, int32_t n_ids,
                        int32_t *size1_out, int32_t *target2, int32_t *klass2);

/*
 * f3 -- barcode -> library matching with maxErr: what barseq_caller.R asks of
 * yogiseq::new.bc.matcher(bcRef$barcode, errCutoff = maxErr)$findMatches (src/barseq_caller.R:57,242).  yogiseq is
 * not vendored in the reference, so its matcher is restated from that call site -- per read: the library rows at the
 * smallest distance, that distance, and how many rows share it -- with the distance of this library (unit-cost
 * Levenshtein, the same Myers kernel and seed index as pcb_edit_pairs): parity unpinned, see DESIGN.md.
 *   library: lib_seq [L, lib_stride], lib_len [L]  (ACGT, 1..64 nt);  queries: q_seq [Q, q_stride], q_len [Q]  (1..64
 *   characters; any character outside ACGT, e.g. N, equals nothing and costs one edit like in a DP over the strings)
 *   best_d [Q]     smallest distance <= max_err to any library row, -1 if there is none
 *   n_best [Q]     library rows at that distance (duplicated library barcodes count once per row), 0 if none
 *   first_hit [Q]  the smallest such row, -1 if none
 *   n_pairs        all (query, row) pairs within max_err; fetch them with pcb_edit_pairs_fetch as (query, row, d),
 *                  sorted by (query, row)
 */
int pcb_match_barcodes(pcb_ctx *ctx, int mem, const uint8_t *lib_seq, const int32_t *lib_len, int32_t L, int32_t lib_stride,
                       const uint8_t *q_seq, const int32_t *q_len, int32_t Q, int32_t q_stride, int32_t max_err,
                       int32_t *best_d, int32_t *n_best, int32_t *first_hit, int64_t *n_pairs);

/*
 * f1 (first row after the hot path) -- barcode-collision tags and the dominance test of the
 * soft filter, on a cluster table given as per-row ids:
 *   collision[i]     = the virtual barcode of row i occurs in more than one row   (tagDups, src/pacybara_translate.R:58-68)
 *   up_collision[i]  = the uptag barcode of row i occurs in more than one row     (same, on upBarcode)
 *   usable[i]        = size_i / (sum of the sizes of all rows with the same uptag) > min_ratio  and  size_i > 1
 *                                                                                 (src/pacybara_softfilter.R:46-63)
 * row_bc / row_up: ids of the barcode strings (0 <= id < n_bc_ids / n_up_ids); a negative id stands for a string
 * that occurs nowhere else.  The ratio is the same fp64 divide and strict > as in R.  Outputs are bytes (0/1).
 */
int pcb_tag_rows(pcb_ctx *ctx, int mem, int32_t n_rows, const int32_t *row_bc, const int32_t *row_up,
                 const int32_t *row_size, int32_t n_bc_ids, int32_t n_up_ids, double min_ratio,
                 uint8_t *collision, uint8_t *up_collision, uint8_t *usable);

/* ------------------------------------------------------------------------------------------------------------
 * f2 -- tokenizers for the text files either side of the path (csrc/ingest.cu).
 *
 * They replace the reference's Python line loops
 *     extracted-barcode FASTQ   src/pacybara_precluster.py:47-63
 *     genoExtract.csv           src/pacybara_cluster.py:302-313
 *     uptag FASTQ               src/pacybara_cluster.py:318-331
 *     read-id string order      src/pacybara_cluster.py:152-157 (makePID)
 * with passes over the decompressed bytes on the device.  A pcb_text owns (or borrows, PCB_MEM_DEVICE) the bytes
 * of one file plus its newline index; it must be closed before its context is destroyed.  Whatever the passes do
 * not model exactly (non-ASCII or NUL bytes, a lone CR, text before the first header, malformed fields, duplicate
 * read ids, a mutation named twice in one read, a 64-bit hash collision) is counted in the n_exotic / n_bad /
 * counters outputs and the result must then be discarded in favour of a line parser: nothing is guessed.
 * `mem` describes the array arguments of each call (the bytes given to pcb_text_open included).
 */
typedef struct pcb_text pcb_text;

int pcb_text_open(pcb_ctx *ctx, int mem, const uint8_t *bytes, int64_t n_bytes, pcb_text **out, int64_t *n_lines,
                  int64_t *n_exotic);
void pcb_text_close(pcb_ctx *ctx, pcb_text *text);

/* FASTQ records as the reference's state machine finds them.  n_bad counts records whose quality and sequence
 * lengths differ or whose header holds no id, plus text before the first header. */
int pcb_fastq_records(pcb_ctx *ctx, pcb_text *fastq, int32_t *n_records, int32_t *max_len, int32_t *max_id_len,
                      int64_t *n_bad);
/* seq, qual: [n_records, stride] zero padded (stride a multiple of 4, >= max_len); id_off/id_len: the read id inside
 * the file's bytes. */
int pcb_fastq_fetch(pcb_ctx *ctx, pcb_text *fastq, int mem, int32_t stride, uint8_t *seq, uint8_t *qual, int32_t *length,
                    int64_t *id_off, int32_t *id_len);
/* rank[i] = position of string i (bytes off[i] .. off[i]+len[i] of `text`) in byte-wise string order; equal strings
 * keep their index order. */
int pcb_rank_strings(pcb_ctx *ctx, const pcb_text *text, int mem, const int64_t *off, const int32_t *len, int32_t n,
                     int32_t max_len, int32_t *rank);
/* Genotypes of the R reads whose ids are (id_off, id_len) in `fastq`, from the genoExtract text `geno`.
 * counters[0] malformed lines, [1] reads without a record (the reference's KeyError), [2] reads naming a mutation
 * twice, [3] duplicate read ids / hash collisions. */
int pcb_geno_join(pcb_ctx *ctx, pcb_text *geno, const pcb_text *fastq, int mem, const int64_t *id_off, const int32_t *id_len,
                  int32_t R, int64_t *nnz, int32_t *n_mut, int64_t counters[4]);
/* geno_ptr [R+1], geno_mut / geno_q [nnz] (mutation ids are dense, 0 <= id < n_mut), mut_off / mut_len [n_mut]:
 * where each mutation's name stands in the genoExtract bytes. */
int pcb_geno_fetch(pcb_ctx *ctx, pcb_text *geno, int mem, int32_t *geno_ptr, int32_t *geno_mut, int32_t *geno_q,
                   int64_t *mut_off, int32_t *mut_len);
/* Uptags: n_records = (id, tag) records in the file, n_tags = distinct tags among the R reads.  counters as above
 * ([0]: tags longer than 65535). */
int pcb_uptag_join(pcb_ctx *ctx, pcb_text *uptag, const pcb_text *fastq, int mem, const int64_t *id_off,
                   const int32_t *id_len, int32_t R, int64_t *n_records, int32_t *n_tags, int64_t counters[4]);
/* up_id [R] (-1: the read has no uptag record), tag_off / tag_len [n_tags] into the uptag bytes. */
int pcb_uptag_fetch(pcb_ctx *ctx, pcb_text *uptag, int mem, int32_t *up_id, int64_t *tag_off, int32_t *tag_len);

#ifdef __cplusplus
}
#endif
#endif /* PACYBARA_B200_H */
