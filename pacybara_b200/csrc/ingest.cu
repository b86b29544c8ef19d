// ingest.cu -- f2: device tokenizers for the text formats either side of the path.
//
// The reference loads its inputs with Python line loops (one dict insert per read):
//   extracted-barcode FASTQ   src/pacybara_precluster.py:47-63   (state machine with the '@' reset)
//   genoExtract.csv           src/pacybara_cluster.py:302-313    ("rid","mut:q;mut:q" | "=", last record of a read wins)
//   uptag FASTQ               src/pacybara_cluster.py:318-331    (line after a header, last record of a read wins)
// and makePID (:152-157) orders reads by their id strings.  Here the decompressed bytes of a file go to the
// device once (pcb_text_open: newline index) and every later step is a data-parallel pass over lines:
//
//   * a FASTQ record is a local pattern: line L is a quality line iff line L-3 starts with '@' and lines
//     L-2, L-1, L do not -- exactly the lines at which the reference's counter reaches 4 (it restarts at every
//     '@' line, which is also why a quality string starting with '@' drops its read);
//   * joins by read id go through an open-addressing table of 64-bit string hashes, every match verified
//     byte by byte; "last record wins" is an atomicMax over line numbers;
//   * mutation and uptag strings are interned the same way (first occurrence = representative), ids are
//     dense in order of first occurrence (deterministic: every rank of a multi-GPU job numbers them alike), and a
//     verification pass proves that no two different strings share an id;
//   * the string order of the read ids is a stable LSD radix sort over big-endian 8-byte words.
//
// Anything the line passes do not model exactly (bytes >= 0x80 or NUL, a lone '\r', text before the first
// header, malformed fields, duplicate ids, a 64-bit hash collision) is COUNTED and reported, never guessed at:
// the host layer then runs its line-by-line parser (hostdata.py), which raises the reference's own errors.
#include "common.cuh"

struct pcb_text {
  const uint8_t *t = nullptr;
  pcb::DevBuf<uint8_t> own;
  int64_t n = 0, n_lines = 0;
  pcb::DevBuf<int64_t> line_start;      // n_lines + 1; line i is [line_start[i], line_start[i+1] - 1)
  // FASTQ records
  pcb::DevBuf<int64_t> rec_line;        // quality line of record r
  int32_t n_rec = -1;
  int64_t n_bad = 0;
  int32_t max_len = 0;
  // last join (genotypes or uptags)
  pcb::DevBuf<int32_t> g_ptr, g_mut, g_q, up_id, rep_len;
  pcb::DevBuf<int64_t> rep_off;
  int64_t nnz = -1;
  int32_t n_distinct = -1, n_reads = -1;
  int kind = 0;                         // 1: genotypes, 2: uptags
};

namespace pcb {
namespace {

constexpr int CHUNK = 16;

__device__ __forceinline__ bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

__device__ __forceinline__ int load_chunk(const uint8_t *__restrict__ t, int64_t n, int64_t c, uint8_t *v) {
  int64_t b = c * CHUNK;
  int m = (int)(n - b < CHUNK ? n - b : CHUNK);
  if (m == CHUNK && (((uintptr_t)(t + b)) & 15) == 0) {
    uint4 q = *reinterpret_cast<const uint4 *>(t + b);
    memcpy(v, &q, 16);
  } else {
    for (int i = 0; i < m; i++) v[i] = t[b + i];
  }
  v[m] = (b + m < n) ? t[b + m] : (uint8_t)'\n';
  return m;
}

// ---- newline index
__global__ void nl_count_kernel(const uint8_t *__restrict__ t, int64_t n, int64_t n_chunks, uint32_t *__restrict__ cnt,
                                unsigned long long *__restrict__ exotic) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  uint8_t v[CHUNK + 1];
  int m = load_chunk(t, n, c, v);
  uint32_t k = 0, odd = 0;
  for (int i = 0; i < m; i++) {
    k += v[i] == '\n';
    odd += (v[i] >= 0x80) | (v[i] == 0) | (v[i] == '\r' && v[i + 1] != '\n');
  }
  cnt[c] = k;
  if (odd) atomicAdd(exotic, (unsigned long long)odd);
}

__global__ void nl_fill_kernel(const uint8_t *__restrict__ t, int64_t n, int64_t n_chunks, const uint32_t *__restrict__ base,
                               int64_t *__restrict__ line_start) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  if (c == 0) line_start[0] = 0;
  uint8_t v[CHUNK + 1];
  int m = load_chunk(t, n, c, v);
  int64_t at = (int64_t)base[c] + 1;
  for (int i = 0; i < m; i++)
    if (v[i] == '\n') line_start[at++] = c * CHUNK + i + 1;
}

__global__ void set_i64_kernel(int64_t *p, int64_t v) { *p = v; }

struct TextView {
  const uint8_t *t;
  const int64_t *ls;
  int64_t n_lines;
};

// line i without its terminator and trailing whitespace (str.rstrip())
__device__ __forceinline__ void span(const TextView &v, int64_t i, int64_t &b, int64_t &e) {
  b = v.ls[i];
  e = v.ls[i + 1] - 1;
  while (e > b && is_space(v.t[e - 1])) e--;
}
__device__ __forceinline__ bool is_header(const TextView &v, int64_t i) {
  int64_t b, e;
  span(v, i, b, e);
  return e > b && v.t[b] == '@';
}

// ---- FASTQ
__global__ void fastq_flag_kernel(TextView v, uint32_t *__restrict__ flag, unsigned long long *__restrict__ exotic) {
  int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= v.n_lines) return;
  bool rec = L >= 3 && is_header(v, L - 3) && !is_header(v, L - 2) && !is_header(v, L - 1) && !is_header(v, L);
  flag[L] = rec ? 1u : 0u;
  if (L == 0 && !is_header(v, 0)) atomicAdd(exotic, 1ull);       // text before the first header: host parser
  // the reference extracts the id of EVERY line that starts with '@' and raises if there is none (:52-57),
  // whether or not a record follows
  int64_t b, e;
  span(v, L, b, e);
  if (e > b && v.t[b] == '@' && (e - b == 1 || v.t[b + 1] == ' ')) atomicAdd(exotic, 1ull);
}

enum { META_MAX_LEN = 0, META_MAX_ID = 1, META_BAD = 2, META_COUNT = 4 };

__global__ void fastq_mark_kernel(TextView v, const uint32_t *__restrict__ pos, int64_t *__restrict__ rec_line,
                                  int32_t *__restrict__ meta) {
  int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool rec = L < v.n_lines && pos[L + 1] != pos[L];
  int32_t slen = 0, idlen = 0, bad = 0;
  if (rec) {
    rec_line[pos[L]] = L;
    int64_t b, e, qb, qe, hb, he;
    span(v, L - 2, b, e);
    span(v, L, qb, qe);
    span(v, L - 3, hb, he);
    slen = (int32_t)(e - b);
    int64_t p = hb + 1;
    while (p < he && v.t[p] != ' ') p++;
    idlen = (int32_t)(p - hb - 1);
    bad = (qe - qb != e - b) || idlen == 0 || (e - b) > 0x7fffffff;
  }
  slen = __reduce_max_sync(0xffffffffu, slen);
  idlen = __reduce_max_sync(0xffffffffu, idlen);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    if (slen) atomicMax(&meta[META_MAX_LEN], slen);
    if (idlen) atomicMax(&meta[META_MAX_ID], idlen);
    if (bad) atomicAdd(&meta[META_BAD], bad);
  }
}

// one thread per (record, 4-byte word) of the two output matrices
__global__ void fastq_fetch_kernel(TextView v, const int64_t *__restrict__ rec_line, int32_t n_rec, int32_t stride,
                                   uint8_t *__restrict__ seq, uint8_t *__restrict__ qual, int32_t *__restrict__ length,
                                   int64_t *__restrict__ id_off, int32_t *__restrict__ id_len) {
  const int32_t wpr = stride / 4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_rec * wpr) return;
  int32_t r = (int32_t)(i / wpr), w = (int32_t)(i % wpr);
  int64_t L = rec_line[r];
  int64_t b, e;
  span(v, L - 2, b, e);
  int64_t qb = v.ls[L];
  int32_t len = (int32_t)(e - b);
  uint32_t s = 0, q = 0;
  for (int k = 0; k < 4; k++) {
    int32_t c = w * 4 + k;
    if (c < len) {
      s |= (uint32_t)v.t[b + c] << (8 * k);
      q |= (uint32_t)v.t[qb + c] << (8 * k);
    }
  }
  reinterpret_cast<uint32_t *>(seq)[i] = s;
  reinterpret_cast<uint32_t *>(qual)[i] = q;
  if (w == 0) {
    length[r] = len;
    int64_t hb, he;
    span(v, L - 3, hb, he);
    int64_t p = hb + 1;
    while (p < he && v.t[p] != ' ') p++;
    id_off[r] = hb + 1;
    id_len[r] = (int32_t)(p - hb - 1);
  }
}

// ---- string hashing, tables
__device__ __forceinline__ uint64_t hash_bytes(const uint8_t *__restrict__ p, int32_t len) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (int32_t i = 0; i < len; i++) h = (h ^ p[i]) * 0x100000001b3ull;
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
  return h | 1ull;                                               // 0 is the empty slot
}
__device__ __forceinline__ bool same_bytes(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int32_t len) {
  for (int32_t i = 0; i < len; i++)
    if (a[i] != b[i]) return false;
  return true;
}

__global__ void id_insert_kernel(const uint8_t *__restrict__ t, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                 int32_t n, unsigned long long *__restrict__ keys, int32_t *__restrict__ vals, uint64_t mask,
                                 unsigned long long *__restrict__ dup) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = hash_bytes(t + off[i], len[i]);
  for (uint64_t s = (key >> 1) & mask;; s = (s + 1) & mask) {
    unsigned long long old = atomicCAS(&keys[s], 0ull, key);
    if (old == 0ull) { vals[s] = i; return; }
    if (old == key) { atomicAdd(dup, 1ull); return; }            // the same id twice (or a hash collision): host parser decides
  }
}

struct IdTable {
  const unsigned long long *keys;
  const int32_t *vals;
  uint64_t mask;
  const uint8_t *t;            // text the ids live in
  const int64_t *off;
  const int32_t *len;
};
__device__ __forceinline__ int32_t id_lookup(const IdTable &T, const uint8_t *__restrict__ q, int32_t qlen) {
  unsigned long long key = hash_bytes(q, qlen);
  for (uint64_t s = (key >> 1) & T.mask;; s = (s + 1) & T.mask) {
    unsigned long long k = T.keys[s];
    if (k == 0ull) return -1;
    if (k == key) {
      int32_t i = T.vals[s];
      if (T.len[i] == qlen && same_bytes(T.t + T.off[i], q, qlen)) return i;
    }
  }
}

// interning: slot of each string; the representative of a slot is its first occurrence (offset << 16 | length)
__global__ void intern_kernel(const uint8_t *__restrict__ t, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                              int64_t n, unsigned long long *__restrict__ keys, unsigned long long *__restrict__ rep, uint64_t mask,
                              int32_t *__restrict__ slot_of) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (len[i] < 0) { slot_of[i] = -1; return; }
  unsigned long long key = hash_bytes(t + off[i], len[i]);
  unsigned long long me = ((unsigned long long)off[i] << 16) | (unsigned long long)(len[i] & 0xffff);
  for (uint64_t s = (key >> 1) & mask;; s = (s + 1) & mask) {
    unsigned long long old = atomicCAS(&keys[s], 0ull, key);
    if (old == 0ull || old == key) {
      atomicMin(&rep[s], me);
      slot_of[i] = (int32_t)s;
      return;
    }
  }
}
__global__ void intern_flag_kernel(const unsigned long long *__restrict__ keys, uint64_t cap, uint32_t *__restrict__ flag) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < cap) flag[s] = keys[s] != 0ull;
}
// ids are dense in order of FIRST OCCURRENCE (the representative's offset): independent of which thread won which
// slot, so every run -- and every rank of a multi-GPU job -- numbers the strings alike
__global__ void intern_keys_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ rep, uint64_t cap,
                                   const uint32_t *__restrict__ at, uint64_t *__restrict__ skey, uint32_t *__restrict__ sval) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap || keys[s] == 0ull) return;
  skey[at[s]] = rep[s] >> 16;
  sval[at[s]] = (uint32_t)s;
}
__global__ void intern_reps_kernel(const uint64_t *__restrict__ skey, const uint32_t *__restrict__ sval, const unsigned long long *__restrict__ rep,
                                   int32_t n_distinct, uint32_t *__restrict__ id_of_slot, int64_t *__restrict__ rep_off,
                                   int32_t *__restrict__ rep_len) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_distinct) return;
  uint32_t s = sval[i];
  id_of_slot[s] = (uint32_t)i;
  rep_off[i] = (int64_t)skey[i];
  rep_len[i] = (int32_t)(rep[s] & 0xffff);
}
// slot -> dense id, and the proof that the string really is its representative
__global__ void intern_remap_kernel(const uint8_t *__restrict__ t, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                    int64_t n, const uint32_t *__restrict__ id_of_slot, const int64_t *__restrict__ rep_off,
                                    const int32_t *__restrict__ rep_len, int32_t *__restrict__ slot_to_id,
                                    unsigned long long *__restrict__ mismatch) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t s = slot_to_id[i];
  if (s < 0) return;
  int32_t id = (int32_t)id_of_slot[s];
  slot_to_id[i] = id;
  if (rep_len[id] != len[i] || !same_bytes(t + rep_off[id], t + off[i], len[i])) atomicAdd(mismatch, 1ull);
}

// ---- genoExtract.csv
__device__ __forceinline__ void strip_quotes(const uint8_t *t, int64_t &b, int64_t &e) {
  while (b < e && t[b] == '"') b++;
  while (e > b && t[e - 1] == '"') e--;
}

// number of "mut:q" pieces of a genotype field, -1 if the reference's parser would raise on it (or if it
// needs Python's int() to decide)
__device__ int32_t geno_pieces(const uint8_t *__restrict__ t, int64_t b, int64_t e) {
  if (e - b == 1 && t[b] == '=') return 0;
  int32_t n = 0;
  int64_t p = b;
  for (;;) {
    int64_t q = p;
    int colons = 0;
    int64_t colon_at = -1;
    while (q < e && t[q] != ';') {
      if (t[q] == ':') { colons++; colon_at = q; }
      q++;
    }
    if (colons != 1) return -1;
    int64_t d = colon_at + 1;
    if (d < q && (t[d] == '+' || t[d] == '-')) d++;
    if (d >= q) return -1;
    for (; d < q; d++)
      if (t[d] < '0' || t[d] > '9') return -1;
    if (q - colon_at > 10 || colon_at - p > 0xffff) return -1;      // quality: at most 9 characters, name: 16 bits of length
    n++;
    if (q >= e) break;
    p = q + 1;
  }
  return n;
}

enum { GC_BAD = 0, GC_MISSING = 1, GC_DUP = 2, GC_MISMATCH = 3, GC_RECORDS = 4, GC_COUNT = 8 };

__global__ void geno_line_kernel(TextView v, IdTable ids, int32_t *__restrict__ line_n, int64_t *__restrict__ field_b,
                                 int32_t *__restrict__ field_len, int32_t *__restrict__ line_of_read,
                                 unsigned long long *__restrict__ counters) {
  int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= v.n_lines) return;
  int64_t b, e;
  span(v, L, b, e);
  int64_t c1 = b;
  while (c1 < e && v.t[c1] != ',') c1++;
  if (c1 >= e) { atomicAdd(&counters[GC_BAD], 1ull); line_n[L] = 0; return; }      // no second field: IndexError
  int64_t c2 = c1 + 1;
  while (c2 < e && v.t[c2] != ',') c2++;
  int64_t rb = b, re = c1, gb = c1 + 1, ge = c2;
  strip_quotes(v.t, rb, re);
  strip_quotes(v.t, gb, ge);
  int32_t n = geno_pieces(v.t, gb, ge);
  if (n < 0) { atomicAdd(&counters[GC_BAD], 1ull); n = 0; }
  line_n[L] = n;
  field_b[L] = gb;
  field_len[L] = (int32_t)(ge - gb);
  int32_t r = id_lookup(ids, v.t + rb, (int32_t)(re - rb));
  if (r >= 0) atomicMax(&line_of_read[r], (int32_t)L);
}

__global__ void geno_count_kernel(const int32_t *__restrict__ line_of_read, const int32_t *__restrict__ line_n, int32_t R,
                                  uint32_t *__restrict__ cnt, unsigned long long *__restrict__ counters) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  if (r == R) { cnt[r] = 0; return; }
  int32_t L = line_of_read[r];
  if (L < 0) { atomicAdd(&counters[GC_MISSING], 1ull); cnt[r] = 0; return; }
  cnt[r] = (uint32_t)line_n[L];
}

__global__ void geno_fill_kernel(const uint8_t *__restrict__ t, const int32_t *__restrict__ line_of_read, const int64_t *__restrict__ field_b,
                                 const int32_t *__restrict__ field_len, const uint32_t *__restrict__ ptr, int32_t R,
                                 int32_t *__restrict__ g_ptr, int64_t *__restrict__ piece_off, int32_t *__restrict__ piece_len,
                                 int32_t *__restrict__ g_q) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  g_ptr[r] = (int32_t)ptr[r];
  if (r == R) return;
  int32_t n = (int32_t)(ptr[r + 1] - ptr[r]);
  if (n == 0) return;
  int32_t L = line_of_read[r];
  int64_t p = field_b[L], e = p + field_len[L];
  for (int32_t k = 0; k < n; k++) {
    int64_t c = p;
    while (t[c] != ':') c++;
    int64_t d = c + 1;
    bool neg = false;
    if (t[d] == '+' || t[d] == '-') { neg = t[d] == '-'; d++; }
    int64_t val = 0;
    while (d < e && t[d] != ';') { val = val * 10 + (t[d] - '0'); d++; }
    piece_off[ptr[r] + k] = p;
    piece_len[ptr[r] + k] = (int32_t)(c - p);
    g_q[ptr[r] + k] = (int32_t)(neg ? -val : val);
    p = d + 1;
  }
}

// the same mutation twice in one read (the reference's dict keeps the first position and the last quality)
__global__ void geno_dup_kernel(const int32_t *__restrict__ g_ptr, const int32_t *__restrict__ g_mut, int32_t R,
                                unsigned long long *__restrict__ counters) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int32_t a = g_ptr[r], b = g_ptr[r + 1];
  for (int32_t i = a + 1; i < b; i++)
    for (int32_t j = a; j < i; j++)
      if (g_mut[i] == g_mut[j]) { atomicAdd(&counters[GC_DUP], 1ull); return; }
}

// ---- uptag FASTQ
__global__ void uptag_line_kernel(TextView v, IdTable ids, int32_t *__restrict__ line_of_read, unsigned long long *__restrict__ counters) {
  int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= v.n_lines || L < 1) return;
  if (!is_header(v, L - 1) || is_header(v, L)) return;
  int64_t hb, he;
  span(v, L - 1, hb, he);
  int64_t p = hb + 1;
  while (p < he && v.t[p] != ' ') p++;
  if (p == hb + 1) return;                                       // rid == "": not stored (:328)
  atomicAdd(&counters[GC_RECORDS], 1ull);
  int32_t r = id_lookup(ids, v.t + hb + 1, (int32_t)(p - hb - 1));
  if (r >= 0) atomicMax(&line_of_read[r], (int32_t)L);
}
__global__ void uptag_items_kernel(TextView v, const int32_t *__restrict__ line_of_read, int32_t R, int64_t *__restrict__ off,
                                   int32_t *__restrict__ len, unsigned long long *__restrict__ counters) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int32_t L = line_of_read[r];
  if (L < 0) { off[r] = 0; len[r] = -1; return; }
  int64_t b, e;
  span(v, L, b, e);
  off[r] = b;
  len[r] = (int32_t)(e - b);
  if (e - b > 0xffff) atomicAdd(&counters[GC_BAD], 1ull);
}

// ---- string order
__device__ __forceinline__ uint64_t word_of(const uint8_t *__restrict__ t, int64_t off, int32_t len, int32_t w) {
  uint64_t k = 0;
  for (int i = 0; i < 8; i++) {
    int32_t c = w * 8 + i;
    k = (k << 8) | (c < len ? t[off + c] : 0);
  }
  return k;
}
__global__ void word_differs_kernel(const uint8_t *__restrict__ t, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                    int32_t n, int32_t n_words, uint32_t *__restrict__ differs) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int32_t w = 0; w < n_words; w++)
    if (word_of(t, off[i], len[i], w) != word_of(t, off[0], len[0], w)) differs[w] = 1u;
}
__global__ void word_keys_kernel(const uint8_t *__restrict__ t, const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                 int32_t n, int32_t w, const uint32_t *__restrict__ cur, uint64_t *__restrict__ key,
                                 uint32_t *__restrict__ val) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = cur[i];
  key[i] = word_of(t, off[s], len[s], w);
  val[i] = s;
}
__global__ void rank_scatter_kernel(const uint32_t *__restrict__ order, int32_t n, int32_t *__restrict__ rank) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rank[order[i]] = i;
}

uint64_t table_capacity(size_t items) {
  uint64_t cap = 1024;
  while (cap < 2 * (uint64_t)items) cap <<= 1;
  return cap;
}

struct Interned {
  int32_t n_distinct = 0;
  DevBuf<int64_t> rep_off;
  DevBuf<int32_t> rep_len;
};
// items (off, len; len < 0: none) -> ids in `ids_out` (-1: none), representatives, mismatch count added to *mismatch
Interned intern_strings(pcb_ctx *ctx, const uint8_t *t, int64_t text_bytes, const int64_t *off, const int32_t *len, int64_t n,
                        int32_t *ids_out, unsigned long long *mismatch) {
  Interned out;
  uint64_t cap = table_capacity((size_t)n);
  DevBuf<unsigned long long> keys(ctx, cap), rep(ctx, cap);
  DevBuf<uint32_t> id_of_slot(ctx, cap + 1);
  PCB_CUDA(cudaMemsetAsync(keys.p, 0, cap * sizeof(unsigned long long), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(rep.p, 0xff, cap * sizeof(unsigned long long), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(id_of_slot.p, 0, (cap + 1) * sizeof(uint32_t), ctx->stream));
  if (n) PCB_LAUNCH(ctx, intern_kernel, grid_for((size_t)n, 256), 256, 0, t, off, len, n, keys.p, rep.p, cap - 1, ids_out);
  PCB_LAUNCH(ctx, intern_flag_kernel, grid_for(cap, 256), 256, 0, keys.p, cap, id_of_slot.p);
  exclusive_scan_u32(ctx, id_of_slot.p, id_of_slot.p, cap + 1);
  out.n_distinct = (int32_t)read_scalar(ctx, id_of_slot.p + cap);
  out.rep_off.alloc(ctx, (size_t)out.n_distinct);
  out.rep_len.alloc(ctx, (size_t)out.n_distinct);
  if (out.n_distinct) {
    const size_t M = (size_t)out.n_distinct;
    DevBuf<uint64_t> k0(ctx, M), k1(ctx, M);
    DevBuf<uint32_t> v0(ctx, M), v1(ctx, M);
    PCB_LAUNCH(ctx, intern_keys_kernel, grid_for(cap, 256), 256, 0, keys.p, rep.p, cap, id_of_slot.p, k0.p, v0.p);
    int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, M, bits_for((uint64_t)(text_bytes > 1 ? text_bytes - 1 : 1)));
    PCB_LAUNCH(ctx, intern_reps_kernel, grid_for(M, 256), 256, 0, w ? k1.p : k0.p, w ? v1.p : v0.p, rep.p, out.n_distinct,
               id_of_slot.p, out.rep_off.p, out.rep_len.p);
  }
  if (n) PCB_LAUNCH(ctx, intern_remap_kernel, grid_for((size_t)n, 256), 256, 0, t, off, len, n, id_of_slot.p, out.rep_off.p,
                    out.rep_len.p, ids_out, mismatch);
  return out;
}

struct BuiltIdTable {
  DevBuf<unsigned long long> keys;
  DevBuf<int32_t> vals;
  uint64_t cap = 0;
  unsigned long long dups = 0;
};
void build_id_table(pcb_ctx *ctx, const uint8_t *t, const int64_t *off, const int32_t *len, int32_t n, BuiltIdTable &T,
                    unsigned long long *dup_counter) {
  T.cap = table_capacity((size_t)n);
  T.keys.alloc(ctx, T.cap);
  T.vals.alloc(ctx, T.cap);
  PCB_CUDA(cudaMemsetAsync(T.keys.p, 0, T.cap * sizeof(unsigned long long), ctx->stream));
  if (n) PCB_LAUNCH(ctx, id_insert_kernel, grid_for((size_t)n, 256), 256, 0, t, off, len, n, T.keys.p, T.vals.p, T.cap - 1, dup_counter);
}

}  // namespace

// ---- entry points on device memory (abi.cu wraps them)

pcb_text *text_open_dev(pcb_ctx *ctx, const uint8_t *bytes, DevBuf<uint8_t> &&own, int64_t n, int64_t *n_exotic) {
  if (n < 0 || n >= (1ll << 46)) throw Error(PCB_E_INPUT, "text size out of range");
  pcb_text *x = new pcb_text();
  try {
    x->own = std::move(own);
    x->t = x->own.p ? x->own.p : bytes;
    x->n = n;
    int64_t n_chunks = (n + CHUNK - 1) / CHUNK;
    DevBuf<uint32_t> cnt(ctx, (size_t)n_chunks + 1);
    DevBuf<unsigned long long> exotic(ctx, 1);
    PCB_CUDA(cudaMemsetAsync(exotic.p, 0, sizeof(unsigned long long), ctx->stream));
    PCB_CUDA(cudaMemsetAsync(cnt.p + n_chunks, 0, sizeof(uint32_t), ctx->stream));
    if (n_chunks) PCB_LAUNCH(ctx, nl_count_kernel, grid_for((size_t)n_chunks, 256), 256, 0, x->t, n, n_chunks, cnt.p, exotic.p);
    exclusive_scan_u32(ctx, cnt.p, cnt.p, (size_t)n_chunks + 1);
    int64_t n_nl = (int64_t)read_scalar(ctx, cnt.p + n_chunks);
    uint8_t last = '\n';
    if (n) last = read_scalar(ctx, x->t + n - 1);
    x->n_lines = n_nl + (last != '\n' ? 1 : 0);
    if (x->n_lines >= 0x7fffffffll) throw Error(PCB_E_INPUT, "more than 2^31 lines");
    x->line_start.alloc(ctx, (size_t)x->n_lines + 1);
    if (n_chunks) PCB_LAUNCH(ctx, nl_fill_kernel, grid_for((size_t)n_chunks, 256), 256, 0, x->t, n, n_chunks, cnt.p, x->line_start.p);
    else PCB_LAUNCH(ctx, set_i64_kernel, 1, 1, 0, x->line_start.p, (int64_t)0);
    if (last != '\n') PCB_LAUNCH(ctx, set_i64_kernel, 1, 1, 0, x->line_start.p + x->n_lines, n + 1);
    *n_exotic = (int64_t)read_scalar(ctx, exotic.p);
    return x;
  } catch (...) {
    delete x;
    throw;
  }
}

void fastq_records_dev(pcb_ctx *ctx, pcb_text *x, int32_t *n_records, int32_t *max_len, int32_t *max_id_len, int64_t *n_bad) {
  TextView v{x->t, x->line_start.p, x->n_lines};
  DevBuf<uint32_t> flag(ctx, (size_t)x->n_lines + 1);
  DevBuf<unsigned long long> exotic(ctx, 1);
  DevBuf<int32_t> meta(ctx, META_COUNT);
  PCB_CUDA(cudaMemsetAsync(exotic.p, 0, sizeof(unsigned long long), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(meta.p, 0, META_COUNT * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(flag.p + x->n_lines, 0, sizeof(uint32_t), ctx->stream));
  if (x->n_lines) PCB_LAUNCH(ctx, fastq_flag_kernel, grid_for((size_t)x->n_lines, 256), 256, 0, v, flag.p, exotic.p);
  exclusive_scan_u32(ctx, flag.p, flag.p, (size_t)x->n_lines + 1);
  x->n_rec = (int32_t)read_scalar(ctx, flag.p + x->n_lines);
  x->rec_line.alloc(ctx, (size_t)x->n_rec);
  if (x->n_lines) PCB_LAUNCH(ctx, fastq_mark_kernel, grid_for((size_t)x->n_lines, 256), 256, 0, v, flag.p, x->rec_line.p, meta.p);
  int32_t h[META_COUNT];
  PCB_CUDA(cudaMemcpyAsync(h, meta.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long ex = read_scalar(ctx, exotic.p);
  *n_records = x->n_rec;
  *max_len = h[META_MAX_LEN];
  *max_id_len = h[META_MAX_ID];
  *n_bad = (int64_t)h[META_BAD] + (int64_t)ex;
  x->n_bad = *n_bad;
  x->max_len = *max_len;
}

void fastq_fetch_dev(pcb_ctx *ctx, pcb_text *x, int32_t stride, uint8_t *seq, uint8_t *qual, int32_t *length, int64_t *id_off,
                     int32_t *id_len) {
  if (x->n_rec < 0) throw Error(PCB_E_STATE, "pcb_fastq_fetch before pcb_fastq_records");
  if (x->n_bad) throw Error(PCB_E_INPUT, "the FASTQ has records this tokenizer does not model; use the line parser");
  if (stride <= 0 || stride % 4) throw Error(PCB_E_INPUT, "stride must be a positive multiple of 4");
  if (stride < x->max_len) throw Error(PCB_E_CAPACITY, "stride " + std::to_string(stride) + " is shorter than the longest sequence (" + std::to_string(x->max_len) + ")");
  TextView v{x->t, x->line_start.p, x->n_lines};
  size_t n = (size_t)x->n_rec * (stride / 4);
  if (n) PCB_LAUNCH(ctx, fastq_fetch_kernel, grid_for(n, 256), 256, 0, v, x->rec_line.p, x->n_rec, stride, seq, qual, length, id_off, id_len);
}

void rank_strings_dev(pcb_ctx *ctx, const pcb_text *x, const int64_t *off, const int32_t *len, int32_t n, int32_t max_len,
                      int32_t *rank) {
  if (n <= 0) return;
  int32_t n_words = (max_len + 7) / 8;
  if (n_words < 1) n_words = 1;
  DevBuf<uint32_t> differs(ctx, (size_t)n_words);
  PCB_CUDA(cudaMemsetAsync(differs.p, 0, n_words * sizeof(uint32_t), ctx->stream));
  PCB_LAUNCH(ctx, word_differs_kernel, grid_for((size_t)n, 256), 256, 0, x->t, off, len, n, n_words, differs.p);
  std::vector<uint32_t> h((size_t)n_words);
  PCB_CUDA(cudaMemcpyAsync(h.data(), differs.p, n_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  DevBuf<uint64_t> k0(ctx, (size_t)n), k1(ctx, (size_t)n);
  DevBuf<uint32_t> v0(ctx, (size_t)n), v1(ctx, (size_t)n);
  iota_u32(ctx, v0.p, (size_t)n);
  uint32_t *cur = v0.p;
  for (int32_t w = n_words - 1; w >= 0; w--) {
    if (!h[w]) continue;                                         // every string has the same 8 bytes here
    PCB_LAUNCH(ctx, word_keys_kernel, grid_for((size_t)n, 256), 256, 0, x->t, off, len, n, w, cur, k0.p, v0.p);
    int where = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, (size_t)n, 64);
    cur = where ? v1.p : v0.p;
  }
  PCB_LAUNCH(ctx, rank_scatter_kernel, grid_for((size_t)n, 256), 256, 0, cur, n, rank);
}

// counters out: [0] malformed lines, [1] reads without a record, [2] reads naming a mutation twice,
// [3] hash collisions / duplicate read ids; all must be 0 for the result to stand
void geno_join_dev(pcb_ctx *ctx, pcb_text *g, const pcb_text *fq, const int64_t *id_off, const int32_t *id_len, int32_t R,
                   int64_t *nnz, int32_t *n_mut, int64_t counters_out[4]) {
  if (g->n_lines >= 0x7fffffffll) throw Error(PCB_E_INPUT, "too many lines");
  TextView v{g->t, g->line_start.p, g->n_lines};
  DevBuf<unsigned long long> counters(ctx, GC_COUNT);
  PCB_CUDA(cudaMemsetAsync(counters.p, 0, GC_COUNT * sizeof(unsigned long long), ctx->stream));
  BuiltIdTable T;
  build_id_table(ctx, fq->t, id_off, id_len, R, T, counters.p + GC_MISMATCH);
  IdTable ids{T.keys.p, T.vals.p, T.cap - 1, fq->t, id_off, id_len};
  DevBuf<int32_t> line_n(ctx, (size_t)g->n_lines), field_len(ctx, (size_t)g->n_lines), line_of_read(ctx, (size_t)R);
  DevBuf<int64_t> field_b(ctx, (size_t)g->n_lines);
  fill_i32(ctx, line_of_read.p, (size_t)R, -1);
  if (g->n_lines) PCB_LAUNCH(ctx, geno_line_kernel, grid_for((size_t)g->n_lines, 256), 256, 0, v, ids, line_n.p, field_b.p, field_len.p,
                             line_of_read.p, counters.p);
  DevBuf<uint32_t> ptr(ctx, (size_t)R + 1);
  PCB_LAUNCH(ctx, geno_count_kernel, grid_for((size_t)R + 1, 256), 256, 0, line_of_read.p, line_n.p, R, ptr.p, counters.p);
  exclusive_scan_u32(ctx, ptr.p, ptr.p, (size_t)R + 1);
  g->nnz = (int64_t)read_scalar(ctx, ptr.p + R);
  if (g->nnz >= 0x7fffffffll) throw Error(PCB_E_INPUT, "more than 2^31 genotype entries");
  g->g_ptr.alloc(ctx, (size_t)R + 1);
  g->g_mut.alloc(ctx, (size_t)g->nnz);
  g->g_q.alloc(ctx, (size_t)g->nnz);
  DevBuf<int64_t> piece_off(ctx, (size_t)g->nnz);
  DevBuf<int32_t> piece_len(ctx, (size_t)g->nnz);
  PCB_LAUNCH(ctx, geno_fill_kernel, grid_for((size_t)R + 1, 256), 256, 0, g->t, line_of_read.p, field_b.p, field_len.p, ptr.p, R,
             g->g_ptr.p, piece_off.p, piece_len.p, g->g_q.p);
  Interned in = intern_strings(ctx, g->t, g->n, piece_off.p, piece_len.p, g->nnz, g->g_mut.p, counters.p + GC_MISMATCH);
  if (R) PCB_LAUNCH(ctx, geno_dup_kernel, grid_for((size_t)R, 256), 256, 0, g->g_ptr.p, g->g_mut.p, R, counters.p);
  g->n_distinct = in.n_distinct;
  g->rep_off = std::move(in.rep_off);
  g->rep_len = std::move(in.rep_len);
  g->n_reads = R;
  g->kind = 1;
  unsigned long long h[GC_COUNT];
  PCB_CUDA(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  *nnz = g->nnz;
  *n_mut = g->n_distinct;
  counters_out[0] = (int64_t)h[GC_BAD];
  counters_out[1] = (int64_t)h[GC_MISSING];
  counters_out[2] = (int64_t)h[GC_DUP];
  counters_out[3] = (int64_t)h[GC_MISMATCH];
}

void geno_fetch_dev(pcb_ctx *ctx, pcb_text *g, int32_t *geno_ptr, int32_t *geno_mut, int32_t *geno_q, int64_t *mut_off, int32_t *mut_len) {
  if (g->kind != 1) throw Error(PCB_E_STATE, "pcb_geno_fetch before pcb_geno_join");
  cudaStream_t s = ctx->stream;
  PCB_CUDA(cudaMemcpyAsync(geno_ptr, g->g_ptr.p, ((size_t)g->n_reads + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (g->nnz) {
    PCB_CUDA(cudaMemcpyAsync(geno_mut, g->g_mut.p, (size_t)g->nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    PCB_CUDA(cudaMemcpyAsync(geno_q, g->g_q.p, (size_t)g->nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  }
  if (g->n_distinct) {
    PCB_CUDA(cudaMemcpyAsync(mut_off, g->rep_off.p, (size_t)g->n_distinct * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    PCB_CUDA(cudaMemcpyAsync(mut_len, g->rep_len.p, (size_t)g->n_distinct * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  }
}

// counters out: [0] tags too long for the interning, [1] 0, [2] 0, [3] hash collisions / duplicate read ids
void uptag_join_dev(pcb_ctx *ctx, pcb_text *u, const pcb_text *fq, const int64_t *id_off, const int32_t *id_len, int32_t R,
                    int64_t *n_records, int32_t *n_tags, int64_t counters_out[4]) {
  if (u->n_lines >= 0x7fffffffll) throw Error(PCB_E_INPUT, "too many lines");
  TextView v{u->t, u->line_start.p, u->n_lines};
  DevBuf<unsigned long long> counters(ctx, GC_COUNT);
  PCB_CUDA(cudaMemsetAsync(counters.p, 0, GC_COUNT * sizeof(unsigned long long), ctx->stream));
  BuiltIdTable T;
  build_id_table(ctx, fq->t, id_off, id_len, R, T, counters.p + GC_MISMATCH);
  IdTable ids{T.keys.p, T.vals.p, T.cap - 1, fq->t, id_off, id_len};
  DevBuf<int32_t> line_of_read(ctx, (size_t)R), len(ctx, (size_t)R);
  DevBuf<int64_t> off(ctx, (size_t)R);
  fill_i32(ctx, line_of_read.p, (size_t)R, -1);
  if (u->n_lines) PCB_LAUNCH(ctx, uptag_line_kernel, grid_for((size_t)u->n_lines, 256), 256, 0, v, ids, line_of_read.p, counters.p);
  if (R) PCB_LAUNCH(ctx, uptag_items_kernel, grid_for((size_t)R, 256), 256, 0, v, line_of_read.p, R, off.p, len.p, counters.p);
  u->up_id.alloc(ctx, (size_t)R);
  Interned in = intern_strings(ctx, u->t, u->n, off.p, len.p, R, u->up_id.p, counters.p + GC_MISMATCH);
  u->n_distinct = in.n_distinct;
  u->rep_off = std::move(in.rep_off);
  u->rep_len = std::move(in.rep_len);
  u->n_reads = R;
  u->kind = 2;
  unsigned long long h[GC_COUNT];
  PCB_CUDA(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_records = (int64_t)h[GC_RECORDS];
  *n_tags = u->n_distinct;
  counters_out[0] = (int64_t)h[GC_BAD];
  counters_out[1] = 0;
  counters_out[2] = 0;
  counters_out[3] = (int64_t)h[GC_MISMATCH];
}

void uptag_fetch_dev(pcb_ctx *ctx, pcb_text *u, int32_t *up_id, int64_t *tag_off, int32_t *tag_len) {
  if (u->kind != 2) throw Error(PCB_E_STATE, "pcb_uptag_fetch before pcb_uptag_join");
  cudaStream_t s = ctx->stream;
  if (u->n_reads) PCB_CUDA(cudaMemcpyAsync(up_id, u->up_id.p, (size_t)u->n_reads * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (u->n_distinct) {
    PCB_CUDA(cudaMemcpyAsync(tag_off, u->rep_off.p, (size_t)u->n_distinct * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    PCB_CUDA(cudaMemcpyAsync(tag_len, u->rep_len.p, (size_t)u->n_distinct * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  }
}

void text_sizes(const pcb_text *x, int64_t *n_lines, int32_t *n_rec, int64_t *nnz, int32_t *n_distinct, int32_t *n_reads) {
  *n_lines = x->n_lines; *n_rec = x->n_rec; *nnz = x->nnz; *n_distinct = x->n_distinct; *n_reads = x->n_reads;
}
void text_free(pcb_text *x) { delete x; }

}  // namespace pcb
