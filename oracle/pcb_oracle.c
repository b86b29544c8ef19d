/*
 * pcb_oracle.c -- CPU restatement of the reference's barcode-clustering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pacybara_b200/ may import, link or call
 * this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs have something to check the CUDA path against
 * on a box where /root/reference (Python) is not present.
 *
 * Parity pinning: this restatement is checked against outputs of the reference's own
 * scripts (src/pacybara_precluster.py, src/pacybara_cluster.py, run unmodified with
 * PYTHONHASHSEED=0 by oracle/make_golden.py) stored under tests/golden/.  The pairwise
 * distance itself is "parity unpinned" at the reference boundary: the reference obtains
 * it from bowtie2 + pacybara_calcEdits.R (src/pacybara.sh:448-468), neither of which is
 * vendored or runnable here; by SURVEY.md 8(c) the distance is DEFINED as unit-cost
 * Levenshtein and orc_lev() below is the textbook DP.
 *
 * Everything works on integer read / bucket / mutation ids; strings stay in Python.
 * Citations are to /root/reference/src/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ utilities */

typedef struct { int64_t *a; int64_t n, cap; } VecI64;
static void vi64_push(VecI64 *v, int64_t x) {
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->a = (int64_t *)realloc(v->a, v->cap * sizeof(int64_t)); }
  v->a[v->n++] = x;
}
typedef struct { int32_t *a; int64_t n, cap; } VecI32;
static void vi32_push(VecI32 *v, int32_t x) {
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->a = (int32_t *)realloc(v->a, v->cap * sizeof(int32_t)); }
  v->a[v->n++] = x;
}

/* open-addressing map  uint64 -> int32 (keys are never 0xFFFF...F) */
typedef struct { uint64_t *k; int32_t *v; uint64_t mask; int64_t n; } Map;
#define MAP_EMPTY 0xFFFFFFFFFFFFFFFFull
static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
static void map_init(Map *m, int64_t cap_hint) {
  uint64_t c = 64; while ((int64_t)c < cap_hint * 2) c <<= 1;
  m->k = (uint64_t *)malloc(c * sizeof(uint64_t)); m->v = (int32_t *)malloc(c * sizeof(int32_t));
  memset(m->k, 0xFF, c * sizeof(uint64_t)); m->mask = c - 1; m->n = 0;
}
static void map_free(Map *m) { free(m->k); free(m->v); m->k = NULL; m->v = NULL; }
static int32_t *map_slot(Map *m, uint64_t key, int *found);
static void map_grow(Map *m) {
  Map o = *m; uint64_t c = (o.mask + 1) * 2;
  m->k = (uint64_t *)malloc(c * sizeof(uint64_t)); m->v = (int32_t *)malloc(c * sizeof(int32_t));
  memset(m->k, 0xFF, c * sizeof(uint64_t)); m->mask = c - 1; m->n = 0;
  for (uint64_t i = 0; i <= o.mask; i++) if (o.k[i] != MAP_EMPTY) { int f; *map_slot(m, o.k[i], &f) = o.v[i]; }
  free(o.k); free(o.v);
}
/* returns the value slot for key, inserting the key if absent (found=0, value uninitialised) */
static int32_t *map_slot(Map *m, uint64_t key, int *found) {
  if ((m->n + 1) * 10 > (int64_t)(m->mask + 1) * 7) map_grow(m);
  uint64_t i = mix64(key) & m->mask;
  for (;;) {
    if (m->k[i] == key) { *found = 1; return &m->v[i]; }
    if (m->k[i] == MAP_EMPTY) { m->k[i] = key; m->n++; *found = 0; return &m->v[i]; }
    i = (i + 1) & m->mask;
  }
}
static int map_get(const Map *m, uint64_t key, int32_t *out) {
  uint64_t i = mix64(key) & m->mask;
  for (;;) {
    if (m->k[i] == key) { *out = m->v[i]; return 1; }
    if (m->k[i] == MAP_EMPTY) return 0;
    i = (i + 1) & m->mask;
  }
}

/* ------------------------------------------------------------------ a1: bucketing
 * pacybara_precluster.py:28-39 (addRecord) and :68-76 (first-seen output order).
 * seq/qual are [R, stride] ASCII; qsum gets, per bucket, the capped Phred sums (not +33).
 * Returns the number of buckets; bucket ids are first-seen ranks. */
int32_t orc_precluster(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int32_t R, int32_t stride,
                       int32_t *bucket_of_read, int32_t *first_read, uint8_t *qsum) {
  Map m; map_init(&m, R);
  int32_t U = 0;
  for (int32_t r = 0; r < R; r++) {
    const uint8_t *s = seq + (int64_t)r * stride; int32_t n = len[r];
    uint64_t h = 1469598103934665603ULL ^ (uint64_t)n;
    for (int32_t i = 0; i < n; i++) { h ^= s[i]; h *= 1099511628211ULL; }
    /* probe on the hash, confirm on the bytes (chained through salted rehash on mismatch) */
    int32_t b = -1;
    for (uint64_t salt = 0;; salt++) {
      int found; int32_t *slot = map_slot(&m, mix64(h + salt * 0x9E3779B97F4A7C15ULL) >> 1, &found);
      if (!found) { *slot = U; b = U; first_read[U] = r; U++; break; }
      int32_t f = first_read[*slot];
      if (len[f] == n && memcmp(seq + (int64_t)f * stride, s, n) == 0) { b = *slot; break; }
    }
    bucket_of_read[r] = b;
    const uint8_t *q = qual + (int64_t)r * stride;
    uint8_t *acc = qsum + (int64_t)b * stride;
    if (first_read[b] == r) {                       /* :38-39  quals[seq] = qualNums */
      for (int32_t i = 0; i < n; i++) acc[i] = (uint8_t)(q[i] - 33);
      for (int32_t i = n; i < stride; i++) acc[i] = 0;
    } else {                                        /* :36-37  min(93, new + old) */
      for (int32_t i = 0; i < n; i++) { int v = (int)acc[i] + (int)q[i] - 33; acc[i] = (uint8_t)(v > 93 ? 93 : v); }
    }
  }
  map_free(&m);
  return U;
}

/* ------------------------------------------------------------------ a2/a3: distance
 * Unit-cost Levenshtein, textbook O(mn) DP (dist := Levenshtein). */
int32_t orc_lev(const uint8_t *a, int32_t m, const uint8_t *b, int32_t n) {
  int32_t prev[256], cur[256];
  if (n > 255 || m > 255) return -1;
  for (int32_t j = 0; j <= n; j++) prev[j] = j;
  for (int32_t i = 1; i <= m; i++) {
    cur[0] = i;
    for (int32_t j = 1; j <= n; j++) {
      int32_t s = prev[j - 1] + (a[i - 1] != b[j - 1]);
      int32_t d = prev[j] + 1, e = cur[j - 1] + 1;
      if (d < s) s = d;
      if (e < s) s = e;
      cur[j] = s;
    }
    memcpy(prev, cur, (n + 1) * sizeof(int32_t));
  }
  return prev[n];
}

/* Same value as orc_lev when it is <= k, else k+1: diagonal band |i-j| <= k only. */
int32_t orc_lev_banded(const uint8_t *a, int32_t m, const uint8_t *b, int32_t n, int32_t k) {
  if (abs(m - n) > k) return k + 1;
  if (n > 255 || m > 255) return -1;
  const int32_t BIG = 1 << 20;
  int32_t prev[258], cur[258];
  for (int32_t j = 0; j <= n; j++) prev[j] = j <= k ? j : BIG;
  for (int32_t i = 1; i <= m; i++) {
    int32_t lo = i - k < 1 ? 1 : i - k, hi = i + k > n ? n : i + k;
    cur[lo - 1] = (lo == 1 && i <= k) ? i : BIG;
    int32_t best = cur[lo - 1];
    for (int32_t j = lo; j <= hi; j++) {
      int32_t s = prev[j - 1] + (a[i - 1] != b[j - 1]);
      int32_t d = (j <= i - 1 + k) ? prev[j] + 1 : BIG;
      int32_t e = cur[j - 1] + 1;
      if (d < s) s = d;
      if (e < s) s = e;
      cur[j] = s;
      if (s < best) best = s;
    }
    if (hi < n) cur[hi + 1] = BIG;
    if (best > k) return k + 1;
    memcpy(prev + (lo - 1), cur + (lo - 1), (hi - lo + 2 + (hi < n)) * sizeof(int32_t));
  }
  return prev[n] <= k ? prev[n] : k + 1;
}

/* All unordered pairs i<j of distinct barcodes with Levenshtein <= k, in (i,j) order.
 * Exhaustive: every pair goes through the full DP.  Returns the count (may exceed cap;
 * only the first cap are stored). */
int64_t orc_edit_pairs_brute(const uint8_t *seq, const int32_t *len, int32_t U, int32_t stride, int32_t k,
                             int32_t *out_i, int32_t *out_j, int32_t *out_d, int64_t cap) {
  int64_t n = 0;
  for (int32_t i = 0; i < U; i++)
    for (int32_t j = i + 1; j < U; j++) {
      if (abs(len[i] - len[j]) > k) continue;
      int32_t d = orc_lev(seq + (int64_t)i * stride, len[i], seq + (int64_t)j * stride, len[j]);
      if (d <= k) { if (n < cap) { out_i[n] = i; out_j[n] = j; out_d[n] = d; } n++; }
    }
  return n;
}

/* f3: nearest library barcodes of every query (what src/barseq_caller.R:57,242 asks of yogiseq's bc.matcher, restated
 * from the call site: the rows at the smallest Levenshtein distance <= k, that distance, their number).  Exhaustive:
 * every (query, row) pair goes through the full DP. */
void orc_match_brute(const uint8_t *lib, const int32_t *lib_len, int32_t L, int32_t lib_stride,
                     const uint8_t *qs, const int32_t *q_len, int32_t Q, int32_t q_stride, int32_t k,
                     int32_t *best_d, int32_t *n_best, int32_t *first_hit) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int32_t q = 0; q < Q; q++) {
    int32_t best = -1, cnt = 0, first = -1;
    for (int32_t j = 0; j < L; j++) {
      if (abs(q_len[q] - lib_len[j]) > k) continue;
      int32_t d = orc_lev(qs + (int64_t)q * q_stride, q_len[q], lib + (int64_t)j * lib_stride, lib_len[j]);
      if (d > k) continue;
      if (best < 0 || d < best) { best = d; cnt = 1; first = j; }
      else if (d == best) cnt++;
    }
    best_d[q] = best; n_best[q] = cnt; first_hit[q] = first;
  }
}

typedef struct { uint64_t key; int32_t j; } SeedEnt;
static int seedent_cmp(const void *a, const void *b) {
  const SeedEnt *x = (const SeedEnt *)a, *y = (const SeedEnt *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return (x->j > y->j) - (x->j < y->j);
}
static int i32_cmp(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }

static uint64_t seed_key(const uint8_t *s, int32_t q, int32_t pos) {
  uint64_t v = 0;
  for (int32_t t = 0; t < q; t++) v = v * 5 + (uint64_t)(s[t] == 'A' ? 0 : s[t] == 'C' ? 1 : s[t] == 'G' ? 2 : s[t] == 'T' ? 3 : 4);
  return (v << 8) | (uint64_t)pos;
}

/* Same result as orc_edit_pairs_brute, candidates from the pigeonhole principle: split the
 * query into k+1 pieces; if lev <= k one piece occurs unchanged in the target, displaced by
 * at most k.  Verified with the banded DP.  Multi-threaded over queries when built with
 * OpenMP (the reference arm of bench.py uses every host core it can). */
int64_t orc_edit_pairs_seeded(const uint8_t *seq, const int32_t *len, int32_t U, int32_t stride, int32_t k,
                              int32_t *out_i, int32_t *out_j, int32_t *out_d, int64_t cap) {
  if (U == 0) return 0;
  int32_t minlen = 1 << 30;
  for (int32_t i = 0; i < U; i++) if (len[i] < minlen) minlen = len[i];
  int32_t q = minlen / (k + 1);
  if (q < 2) return orc_edit_pairs_brute(seq, len, U, stride, k, out_i, out_j, out_d, cap);
  if (q > 24) q = 24;
  /* index: every q-gram of every barcode, keyed by (q-gram, position) */
  int64_t nent = 0;
  for (int32_t j = 0; j < U; j++) nent += len[j] - q + 1;
  SeedEnt *ent = (SeedEnt *)malloc((nent + 1) * sizeof(SeedEnt));
  int64_t e = 0;
  for (int32_t j = 0; j < U; j++)
    for (int32_t p = 0; p + q <= len[j]; p++) { ent[e].key = seed_key(seq + (int64_t)j * stride + p, q, p); ent[e].j = j; e++; }
  qsort(ent, nent, sizeof(SeedEnt), seedent_cmp);
  int64_t *cnt = (int64_t *)calloc(U + 1, sizeof(int64_t));
  VecI32 *res_j = (VecI32 *)calloc(U, sizeof(VecI32)), *res_d = (VecI32 *)calloc(U, sizeof(VecI32));
#pragma omp parallel
  {
    VecI32 cand = {0, 0, 0};
#pragma omp for schedule(dynamic, 256)
    for (int32_t i = 0; i < U; i++) {
      const uint8_t *s = seq + (int64_t)i * stride; int32_t m = len[i];
      cand.n = 0;
      for (int32_t t = 0; t <= k; t++) {
        int32_t st = (int32_t)((int64_t)t * m / (k + 1));           /* piece t starts here, is >= q long */
        for (int32_t sh = -k; sh <= k; sh++) {
          int32_t p = st + sh; if (p < 0) continue;
          uint64_t key = seed_key(s + st, q, p);
          int64_t lo = 0, hi = nent;
          while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (ent[mid].key < key) lo = mid + 1; else hi = mid; }
          for (; lo < nent && ent[lo].key == key; lo++) if (ent[lo].j > i) vi32_push(&cand, ent[lo].j);
        }
      }
      qsort(cand.a, cand.n, sizeof(int32_t), i32_cmp);
      int32_t last = -1;
      for (int64_t c = 0; c < cand.n; c++) {
        int32_t j = cand.a[c]; if (j == last) continue; last = j;
        int32_t d = orc_lev_banded(s, m, seq + (int64_t)j * stride, len[j], k);
        if (d <= k) { vi32_push(&res_j[i], j); vi32_push(&res_d[i], d); }
      }
      cnt[i + 1] = res_j[i].n;
    }
    free(cand.a);
  }
  int64_t n = 0;
  for (int32_t i = 0; i < U; i++) {
    for (int64_t c = 0; c < res_j[i].n; c++) {
      if (n < cap) { out_i[n] = i; out_j[n] = res_j[i].a[c]; out_d[n] = res_d[i].a[c]; }
      n++;
    }
    free(res_j[i].a); free(res_d[i].a);
  }
  free(res_j); free(res_d); free(cnt); free(ent);
  return n;
}

/* ------------------------------------------------------------------ a4-a14: cluster merge
 * A literal, read-pair-level restatement of pacybara_cluster.py.  Reads, buckets and
 * mutations are integer ids; `lexrank[r]` is the rank of read r's id in Python string
 * order (makePID, :152-157, compares the id strings). */

typedef struct {
  int32_t R, U, maxDiff, minQual, minMatches;
  double minJaccard;
  const int32_t *bucket_ptr, *bucket_reads, *lexrank, *geno_ptr, *geno_mut, *geno_q, *mut_rank;
  /* state */
  int32_t *rid2cid;           /* -1 = not in rid2cid (:179) */
  VecI32 *cid2rids;           /* indexed by cid, n<0 => deleted (:180) */
  int32_t lastCID, cidCap;
  Map known;                  /* knownRIDPairs (:150) */
  Map edistLookup;            /* :236 */
  VecI64 *edistPairs;         /* insertion-ordered keys per distance (:235) */
  Map *edistSeen;
  int64_t nLinks;
} Ctx;

static uint64_t makePID(const Ctx *c, int32_t r1, int32_t r2) {               /* :152-157 */
  int32_t a = r1, b = r2;
  if (!(c->lexrank[r1] < c->lexrank[r2])) { a = r2; b = r1; }
  return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
}
static void markAsKnown(Ctx *c, int32_t r1, int32_t r2) { int f; *map_slot(&c->known, makePID(c, r1, r2), &f) = 1; } /* :159 */

static int32_t newCID(Ctx *c) {
  c->lastCID++;
  if (c->lastCID >= c->cidCap) {
    int32_t nc = c->cidCap * 2;
    c->cid2rids = (VecI32 *)realloc(c->cid2rids, nc * sizeof(VecI32));
    memset(c->cid2rids + c->cidCap, 0, (nc - c->cidCap) * sizeof(VecI32));
    c->cidCap = nc;
  }
  return c->lastCID;
}

static void addLink(Ctx *c, int32_t rid1, int32_t rid2) {                      /* :185-220 */
  c->nLinks++;
  if (c->rid2cid[rid1] >= 0) {
    if (c->rid2cid[rid2] >= 0) {
      if (c->rid2cid[rid1] != c->rid2cid[rid2]) {
        int32_t cid1 = c->rid2cid[rid1], cid2 = c->rid2cid[rid2];
        VecI32 *l2 = &c->cid2rids[cid2];
        for (int64_t i = 0; i < l2->n; i++) c->rid2cid[l2->a[i]] = cid1;
        for (int64_t i = 0; i < l2->n; i++) vi32_push(&c->cid2rids[cid1], l2->a[i]);
        free(l2->a); l2->a = NULL; l2->n = -1; l2->cap = 0;                    /* del cid2rids[cid2] */
      }
    } else {
      int32_t cid = c->rid2cid[rid1]; c->rid2cid[rid2] = cid; vi32_push(&c->cid2rids[cid], rid2);
    }
  } else {
    if (c->rid2cid[rid2] >= 0) {
      int32_t cid = c->rid2cid[rid2]; c->rid2cid[rid1] = cid; vi32_push(&c->cid2rids[cid], rid1);
    } else {
      int32_t cid = newCID(c);
      c->rid2cid[rid1] = cid; c->rid2cid[rid2] = cid;
      vi32_push(&c->cid2rids[cid], rid1); vi32_push(&c->cid2rids[cid], rid2);
    }
  }
}

/* varMatrix (:338-347): muts = union of the reads' keys (here: ascending id), varMat[m][read] */
typedef struct { int32_t M, n; int32_t *muts; int32_t *mat; } VarMat;
static void varMatrix(const Ctx *c, const int32_t *rids, int32_t n, int32_t penalty, VarMat *vm) {
  int64_t tot = 0;
  for (int32_t i = 0; i < n; i++) tot += c->geno_ptr[rids[i] + 1] - c->geno_ptr[rids[i]];
  int32_t *all = (int32_t *)malloc((tot + 1) * sizeof(int32_t)); int64_t t = 0;
  for (int32_t i = 0; i < n; i++) for (int32_t e = c->geno_ptr[rids[i]]; e < c->geno_ptr[rids[i] + 1]; e++) all[t++] = c->geno_mut[e];
  qsort(all, tot, sizeof(int32_t), i32_cmp);
  int32_t M = 0;
  for (int64_t i = 0; i < tot; i++) if (i == 0 || all[i] != all[i - 1]) all[M++] = all[i];
  vm->M = M; vm->n = n; vm->muts = all;
  vm->mat = (int32_t *)malloc(((int64_t)M * n + 1) * sizeof(int32_t));
  for (int64_t i = 0; i < (int64_t)M * n; i++) vm->mat[i] = penalty;
  for (int32_t i = 0; i < n; i++)
    for (int32_t e = c->geno_ptr[rids[i]]; e < c->geno_ptr[rids[i] + 1]; e++) {
      int32_t mu = c->geno_mut[e], lo = 0, hi = M;
      while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (all[mid] < mu) lo = mid + 1; else hi = mid; }
      vm->mat[(int64_t)lo * n + i] = c->geno_q[e];                 /* later duplicate keys win, as in a dict */
    }
}
static void vm_free(VarMat *vm) { free(vm->muts); free(vm->mat); }
static void filterSeqError(const Ctx *c, VarMat *vm) {                         /* :350-356 */
  int32_t keep = 0;
  for (int32_t m = 0; m < vm->M; m++) {
    int32_t observed = 0; int64_t sum = 0;
    for (int32_t i = 0; i < vm->n; i++) { int32_t q = vm->mat[(int64_t)m * vm->n + i]; if (q > 0) observed++; sum += q; }
    if (observed > 1 || sum > c->minQual) {
      if (keep != m) { vm->muts[keep] = vm->muts[m]; memmove(vm->mat + (int64_t)keep * vm->n, vm->mat + (int64_t)m * vm->n, vm->n * sizeof(int32_t)); }
      keep++;
    }
  }
  vm->M = keep;
}
/* calcJaccard (:358-367) on two columns given as callbacks over m */
static void jaccard_cols(const int64_t *qs1, const int64_t *qs2, int32_t M, double *jacc, int32_t *inters, int32_t *uni) {
  int32_t in = 0, un = 0;
  for (int32_t m = 0; m < M; m++) {
    if (qs1[m] > 0 && qs2[m] > 0) in++;
    if (qs1[m] > 0 || qs2[m] > 0) un++;
  }
  *jacc = un > 0 ? (double)in / (double)un : 1.0; *inters = in; *uni = un;
}

static void edist_insert(Ctx *c, int32_t ed, int32_t r1, int32_t r2) {         /* :264-266 / :292-294 */
  uint64_t pid = makePID(c, r1, r2); int f;
  map_slot(&c->edistSeen[ed], pid, &f);
  if (!f) vi64_push(&c->edistPairs[ed], (int64_t)pid);
  *map_slot(&c->edistLookup, pid, &f) = ed;
}

/* Outputs: rows = surviving clusters in cid creation order, then singletons in read order.
 *   row_ptr[n_rows+1], row_members[R]      members in list order
 *   row_bc[n_rows], row_up[n_rows]         consensus ids; -2 = reference would call muscle (:530)
 *   call_ptr[n_rows+1], call_mut[<=nnz]    consensus / singleton genotype
 *   stats[0]=n_clusters, [1]=n_links, [2]=n main-step pair visits
 * bc_id / up_id: per-read ids of the barcode / uptag strings (up_id may be NULL).
 * Returns n_rows, or <0 on error. */
int32_t orc_cluster(int32_t R, int32_t U, const int32_t *bucket_ptr, const int32_t *bucket_reads,
                    const int32_t *lexrank, const int32_t *bc_id, const int32_t *up_id,
                    const int32_t *geno_ptr, const int32_t *geno_mut, const int32_t *geno_q, const int32_t *mut_rank,
                    int64_t n_ed, const int32_t *ed_q, const int32_t *ed_r, const int32_t *ed_d,
                    int32_t maxDiff, int32_t minQual, double minJaccard, int32_t minMatches,
                    int32_t *row_ptr, int32_t *row_members, int32_t *row_bc, int32_t *row_up,
                    int32_t *call_ptr, int32_t *call_mut, int64_t *stats) {
  Ctx cx; memset(&cx, 0, sizeof(cx));
  Ctx *c = &cx;
  c->R = R; c->U = U; c->maxDiff = maxDiff; c->minQual = minQual; c->minMatches = minMatches; c->minJaccard = minJaccard;
  c->bucket_ptr = bucket_ptr; c->bucket_reads = bucket_reads; c->lexrank = lexrank;
  c->geno_ptr = geno_ptr; c->geno_mut = geno_mut; c->geno_q = geno_q; c->mut_rank = mut_rank;
  c->rid2cid = (int32_t *)malloc((R + 1) * sizeof(int32_t));
  for (int32_t r = 0; r < R; r++) c->rid2cid[r] = -1;
  c->cidCap = 1024; c->cid2rids = (VecI32 *)calloc(c->cidCap, sizeof(VecI32));
  map_init(&c->known, R); map_init(&c->edistLookup, R);
  int32_t nD = 2 * maxDiff + 1;
  c->edistPairs = (VecI64 *)calloc(nD, sizeof(VecI64));
  c->edistSeen = (Map *)calloc(nD, sizeof(Map));
  for (int32_t d = 0; d < nD; d++) map_init(&c->edistSeen[d], 1024);

  /* LOAD: edit distances (:250-266) */
  for (int64_t e = 0; e < n_ed; e++) {
    int32_t ed = ed_d[e];
    if (ed < 0) continue;
    if (ed <= maxDiff * 2)
      for (int32_t x = bucket_ptr[ed_q[e]]; x < bucket_ptr[ed_q[e] + 1]; x++)
        for (int32_t y = bucket_ptr[ed_r[e]]; y < bucket_ptr[ed_r[e] + 1]; y++)
          edist_insert(c, ed, bucket_reads[x], bucket_reads[y]);
  }
  /* LOAD: barcode groups (:268-298): in-bucket pairs at distance 0 */
  for (int32_t b = 0; b < U; b++) {
    int32_t lo = bucket_ptr[b], hi = bucket_ptr[b + 1];
    if (hi - lo > 1)
      for (int32_t x = lo; x < hi; x++) for (int32_t y = x + 1; y < hi; y++) edist_insert(c, 0, bucket_reads[x], bucket_reads[y]);
  }

  /* SEED CLUSTERING (:385-400) */
  for (int32_t b = 0; b < U; b++) {
    int32_t lo = bucket_ptr[b], n = bucket_ptr[b + 1] - lo;
    if (n < 2) continue;
    const int32_t *rids = bucket_reads + lo;
    VarMat vm; varMatrix(c, rids, n, 0, &vm); filterSeqError(c, &vm);
    int64_t *qs1 = (int64_t *)malloc((vm.M + 1) * sizeof(int64_t)), *qs2 = (int64_t *)malloc((vm.M + 1) * sizeof(int64_t));
    for (int32_t i = 1; i < n; i++) {
      for (int32_t m = 0; m < vm.M; m++) qs1[m] = vm.mat[(int64_t)m * n + i];
      for (int32_t j = 0; j < i; j++) {
        markAsKnown(c, rids[i], rids[j]);
        for (int32_t m = 0; m < vm.M; m++) qs2[m] = vm.mat[(int64_t)m * n + j];
        double jacc; int32_t inters, uni; jaccard_cols(qs1, qs2, vm.M, &jacc, &inters, &uni);
        if (uni == 0 || (jacc > minJaccard && inters >= minMatches)) addLink(c, rids[i], rids[j]);
      }
    }
    free(qs1); free(qs2); vm_free(&vm);
  }

  /* MAIN CLUSTERING (:408-455) */
  int64_t visits = 0;
  VecI32 both = {0, 0, 0};
  for (int32_t ed = 1; ed <= maxDiff; ed++) {
    for (int64_t p = 0; p < c->edistPairs[ed].n; p++) {
      uint64_t pid = (uint64_t)c->edistPairs[ed].a[p]; int32_t dummy;
      if (map_get(&c->known, pid, &dummy)) continue;                         /* :415 */
      int32_t rid1 = (int32_t)(pid >> 32), rid2 = (int32_t)(pid & 0xFFFFFFFFu);   /* :418 */
      markAsKnown(c, rid1, rid2);
      visits++;
      int32_t cid1 = c->rid2cid[rid1], cid2 = c->rid2cid[rid2];
      if (cid1 >= 0 && cid1 == cid2) continue;                               /* :424 */
      const int32_t *cl1 = &rid1, *cl2 = &rid2; int32_t size1 = 1, size2 = 1;
      if (cid1 >= 0) { cl1 = c->cid2rids[cid1].a; size1 = (int32_t)c->cid2rids[cid1].n; }
      if (cid2 >= 0) { cl2 = c->cid2rids[cid2].a; size2 = (int32_t)c->cid2rids[cid2].n; }
      double sizeDispro = fabs(log2((double)size1 / (double)size2));        /* :435 */
      if (sizeDispro < ed) continue;
      int32_t maxED = 0, missing = 0;                                        /* :442 */
      for (int32_t i = 0; i < size1 && !missing; i++)
        for (int32_t j = 0; j < size2; j++) {
          int32_t d;
          if (!map_get(&c->edistLookup, makePID(c, cl1[i], cl2[j]), &d)) { missing = 1; break; }
          if (d > maxED) maxED = d;
        }
      if (missing || maxED > 2 * ed) continue;                               /* :443 */
      both.n = 0;
      for (int32_t i = 0; i < size1; i++) vi32_push(&both, cl1[i]);
      for (int32_t j = 0; j < size2; j++) vi32_push(&both, cl2[j]);
      VarMat vm; varMatrix(c, both.a, size1 + size2, 0, &vm); filterSeqError(c, &vm);   /* :447-448 */
      int64_t *qs1 = (int64_t *)calloc(vm.M + 1, sizeof(int64_t)), *qs2 = (int64_t *)calloc(vm.M + 1, sizeof(int64_t));
      for (int32_t m = 0; m < vm.M; m++) {
        for (int32_t i = 0; i < size1; i++) qs1[m] += vm.mat[(int64_t)m * vm.n + i];
        for (int32_t i = size1; i < size1 + size2; i++) qs2[m] += vm.mat[(int64_t)m * vm.n + i];
      }
      double jacc; int32_t inters, uni; jaccard_cols(qs1, qs2, vm.M, &jacc, &inters, &uni);
      free(qs1); free(qs2); vm_free(&vm);
      if (uni == 0 || jacc > minJaccard) addLink(c, rid1, rid2);             /* :452-453 */
    }
  }
  free(both.a);

  /* CONSENSUS + OUTPUT (:464-477, :515-537, :546-567) */
  int32_t n_rows = 0, n_clusters = 0; int64_t nm = 0, nc = 0; int32_t err = 0;
  row_ptr[0] = 0; call_ptr[0] = 0;
  for (int32_t cid = 1; cid <= c->lastCID && !err; cid++) {
    VecI32 *l = &c->cid2rids[cid];
    if (l->n < 0) continue;
    int32_t n = (int32_t)l->n;
    for (int32_t i = 0; i < n; i++) row_members[nm++] = l->a[i];
    /* genotype vote (:468-476) */
    VarMat vm; varMatrix(c, l->a, n, -minQual, &vm);
    int32_t ncall = 0; int32_t *calls = (int32_t *)malloc((vm.M + 1) * sizeof(int32_t));
    for (int32_t m = 0; m < vm.M; m++) {
      int64_t sum = 0; for (int32_t i = 0; i < n; i++) sum += vm.mat[(int64_t)m * n + i];
      if (sum > 0) calls[ncall++] = vm.muts[m];
    }
    for (int32_t a = 1; a < ncall; a++) {            /* sort by the (digit key, string) rank */
      int32_t v = calls[a], b = a - 1;
      while (b >= 0 && mut_rank[calls[b]] > mut_rank[v]) { calls[b + 1] = calls[b]; b--; }
      calls[b + 1] = v;
    }
    for (int32_t a = 0; a < ncall; a++) call_mut[nc++] = calls[a];
    free(calls); vm_free(&vm);
    /* consensus barcode, twice (:515-537) */
    for (int pass = 0; pass < 2; pass++) {
      const int32_t *ids = pass == 0 ? bc_id : up_id;
      int32_t *dst = pass == 0 ? row_bc : row_up;
      if (ids == NULL) { dst[n_rows] = row_bc[n_rows]; continue; }      /* finalUptags = finalBCs (:537) */
      int32_t best = -1, bestCnt = 0, nTop = 0;
      for (int32_t i = 0; i < n; i++) {
        int32_t id = ids[l->a[i]];
        if (id < 0) { err = -3; break; }                                   /* KeyError in the reference */
        int32_t seen = 0;
        for (int32_t j = 0; j < i; j++) if (ids[l->a[j]] == id) { seen = 1; break; }
        if (seen) continue;
        int32_t cnt = 0; for (int32_t j = i; j < n; j++) if (ids[l->a[j]] == id) cnt++;
        if (cnt > bestCnt) { bestCnt = cnt; best = id; nTop = 1; } else if (cnt == bestCnt) nTop++;
      }
      if (nTop == 1) dst[n_rows] = best;
      else if (n == 2) dst[n_rows] = ids[l->a[0]];
      else dst[n_rows] = -2;
    }
    n_rows++; n_clusters++;
    row_ptr[n_rows] = (int32_t)nm; call_ptr[n_rows] = (int32_t)nc;
  }
  for (int32_t r = 0; r < R && !err; r++) {          /* singletons (:560-567) */
    if (c->rid2cid[r] >= 0) continue;
    row_members[nm++] = r;
    for (int32_t e = geno_ptr[r]; e < geno_ptr[r + 1]; e++) {
      /* dict semantics: a repeated key keeps its first position and its last value */
      int32_t dup = 0; for (int32_t f = geno_ptr[r]; f < e; f++) if (geno_mut[f] == geno_mut[e]) { dup = 1; break; }
      if (dup) continue;
      int32_t q = geno_q[e]; for (int32_t f = e + 1; f < geno_ptr[r + 1]; f++) if (geno_mut[f] == geno_mut[e]) q = geno_q[f];
      if (q >= minQual) call_mut[nc++] = geno_mut[e];
    }
    row_bc[n_rows] = bc_id[r]; row_up[n_rows] = up_id ? up_id[r] : bc_id[r];
    n_rows++;
    row_ptr[n_rows] = (int32_t)nm; call_ptr[n_rows] = (int32_t)nc;
  }
  if (stats) { stats[0] = n_clusters; stats[1] = c->nLinks; stats[2] = visits; }

  for (int32_t cid = 0; cid < c->cidCap; cid++) if (c->cid2rids[cid].n >= 0) free(c->cid2rids[cid].a);
  free(c->cid2rids); free(c->rid2cid);
  map_free(&c->known); map_free(&c->edistLookup);
  for (int32_t d = 0; d < nD; d++) { free(c->edistPairs[d].a); map_free(&c->edistSeen[d]); }
  free(c->edistPairs); free(c->edistSeen);
  return err ? err : n_rows;
}

/* torchrun exports OMP_NUM_THREADS=1; the reference arm wants every host core it can use */
void orc_set_threads(int32_t n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int32_t orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
