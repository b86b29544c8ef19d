// bucket.cu -- exact-barcode bucketing on the device (row a1 of the scope table).
//
// Replaces the dictionary of src/pacybara_precluster.py:28-39 and its first-seen output
// order (:68-76).  Reads are packed to 2-bit planes, inserted into an open-addressing table
// keyed by (planes, length) whose 8-byte slots hold (smallest read index, count); buckets are
// numbered by the smallest read index they contain (= first-seen order); every read takes a place
// in its bucket's segment with one atomic and the segments are then put into ascending read order
// (= arrival order) bucket by bucket -- no sort over all reads; the per-position qualities are
// summed with the reference's cap of 93.  The whole stage is launched without a host round trip.
// HBM-bound, and at the 2*10^7-read size bound by RANDOM sector accesses: about 6 per read
// (slot, the slot owner's planes and length, slot again for the bucket number, the bucket's cursor
// and pointer, the place) against 15 with the three-array table and the radix sort of round 1.
#include "common.cuh"
#include "myers.cuh"

namespace pcb {

// flags[0] = first offending row + 1 (0xFFFFFFFF: none), flags[1] = min length, flags[2] = max length, flags[3] = opaque barcodes,
// flags[4 + m] = barcodes of length m (m = 1..64): the pair search picks its seed lengths from that histogram
constexpr int PACK_FLAGS = 4 + 66;
constexpr int PACK_BLOCK = 128;
// STAGED: the block's rows are one contiguous byte range (rows == nullptr); it is brought into shared memory with 16-byte
// streaming loads and every thread then reads its own row from there -- a thread reading its 25-byte row straight from
// global memory issues 25 one-byte loads at a 25-byte stride (450 GB/s in round 1, profiles/r01q_launches.md).
template <bool STAGED>
__global__ void __launch_bounds__(PACK_BLOCK) pack_kernel(const uint8_t *__restrict__ seq, const int32_t *__restrict__ len,
                                                          const int32_t *__restrict__ rows, int32_t n, int32_t stride,
                                                          ulonglong2 *__restrict__ planes, int32_t *__restrict__ out_len, int32_t *flags,
                                                          uint64_t *__restrict__ other) {
  extern __shared__ uint4 pack_sh[];
  __shared__ uint32_t sh_hist[66];
  if (threadIdx.x < 66) sh_hist[threadIdx.x] = 0;
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint8_t *s = nullptr;
  if (STAGED) {
    const size_t row0 = (size_t)blockIdx.x * PACK_BLOCK;
    const int32_t n_rows = n - (int32_t)row0 < PACK_BLOCK ? n - (int32_t)row0 : PACK_BLOCK;
    const uint8_t *g = seq + row0 * (size_t)stride;
    const uintptr_t a0 = (uintptr_t)g & ~(uintptr_t)15;            // (stays inside the allocation: those are 16-byte granular)
    const int32_t head = (int32_t)((uintptr_t)g - a0);
    const int32_t n_vec = (head + n_rows * stride + 15) >> 4;
    for (int32_t v = threadIdx.x; v < n_vec; v += PACK_BLOCK) pack_sh[v] = __ldcs((const uint4 *)a0 + v);
    s = (const uint8_t *)pack_sh + head + (size_t)threadIdx.x * stride;
  }
  __syncthreads();
  if (i < n) {
    int32_t row = rows ? rows[i] : i;
    int32_t m = len[row];
    if (!STAGED) s = seq + (size_t)row * stride;
    bool bad = m < 0 || m > stride;
    // The reference treats barcodes as opaque strings (src/pacybara_precluster.py:28-39).  One that does not fit the
    // 2-bit planes -- a character outside ACGT, more than 64 bases, or none -- is kept as such: its planes hold a
    // 128-bit hash of its bytes and its packed length the marker PCB_LEN_OPAQUE, so that it buckets with its exact
    // copies and nothing else, and the pair search leaves it alone.
    bool opaque = !bad && other == nullptr && (m < 1 || m > 64);
    if (other != nullptr && (m < 1 || m > 64)) bad = true;
    uint64_t L = 0, H = 0, X = 0;
    if (!bad && !opaque) {
      // base -> (L, H) bit without a compare chain: A 0x41, C 0x43, G 0x47, T 0x54 have H = bit 2 and L = bit 1 ^ bit 2;
      // the character is a base iff it is the one its own code names.  32-bit halves: a 64-bit shift is several instructions.
      uint32_t l0 = 0, h0 = 0, x0 = 0, l1 = 0, h1 = 0, x1 = 0;
      const int32_t m0 = m < 32 ? m : 32;
      for (int32_t t = 0; t < m0; t++) {
        const uint32_t ch = s[t], hb = (ch >> 2) & 1u, lb = ((ch >> 1) ^ (ch >> 2)) & 1u;
        const uint32_t ok = ((0x54474341u >> (8u * (lb | (hb << 1)))) & 0xFFu) == ch ? 1u : 0u;
        l0 |= (lb & ok) << t; h0 |= (hb & ok) << t; x0 |= (ok ^ 1u) << t;
      }
      for (int32_t t = 32; t < m; t++) {
        const uint32_t ch = s[t], hb = (ch >> 2) & 1u, lb = ((ch >> 1) ^ (ch >> 2)) & 1u;
        const uint32_t ok = ((0x54474341u >> (8u * (lb | (hb << 1)))) & 0xFFu) == ch ? 1u : 0u;
        l1 |= (lb & ok) << (t - 32); h1 |= (hb & ok) << (t - 32); x1 |= (ok ^ 1u) << (t - 32);
      }
      L = ((uint64_t)l1 << 32) | l0; H = ((uint64_t)h1 << 32) | h0; X = ((uint64_t)x1 << 32) | x0;
      if (X != 0 && other == nullptr) { opaque = true; L = H = X = 0; }       // a character that is no base: an opaque barcode
    }
    if (opaque) {
      uint64_t h1 = 0x9E3779B97F4A7C15ull ^ (uint64_t)m, h2 = 0xC2B2AE3D27D4EB4Full + (uint64_t)m;
      for (int32_t t = 0; t < m; t++) {
        h1 = (h1 ^ s[t]) * 0x100000001B3ull; h1 ^= h1 >> 29;
        h2 = (h2 + s[t]) * 0xD6E8FEB86659FD93ull; h2 ^= h2 >> 32;
      }
      L = h1; H = h2;
    }
    planes[i] = make_ulonglong2(L, H);
    if (other) other[i] = X;
    out_len[i] = opaque ? PCB_LEN_OPAQUE : m;
    if (bad) atomicMin((uint32_t *)&flags[0], (uint32_t)i + 1u);
    else if (opaque) atomicAdd(&flags[3], 1);
    else atomicAdd(&sh_hist[m], 1u);
  }
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x <= 64 && sh_hist[threadIdx.x]) {
    atomicAdd((uint32_t *)&flags[4 + threadIdx.x], sh_hist[threadIdx.x]);
    atomicMin(&flags[1], (int32_t)threadIdx.x);
    atomicMax(&flags[2], (int32_t)threadIdx.x);
  }
}

// launches the packing; the flags stay on the device until pack_finish() has them (after a stream synchronisation)
void pack_launch(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, const int32_t *rows, int32_t n, int32_t stride,
                 PackedBarcodes &out, bool allow_other, DevBuf<int32_t> &flags) {
  out.planes.alloc(ctx, (size_t)n);
  out.len.alloc(ctx, (size_t)n);
  if (allow_other) out.other.alloc(ctx, (size_t)n);
  out.n = n;
  flags.alloc(ctx, PACK_FLAGS);
  PCB_CUDA(cudaMemsetAsync(flags.p, 0, PACK_FLAGS * sizeof(int32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(flags.p, 0xFF, sizeof(int32_t), ctx->stream));                 // flags[0] as unsigned: 0xFFFFFFFF = none
  PCB_CUDA(cudaMemsetAsync(flags.p + 1, 0x7F, sizeof(int32_t), ctx->stream));             // min length: a large value
  const size_t stage_bytes = (size_t)PACK_BLOCK * (size_t)stride + 32;
  if (n > 0 && rows == nullptr && stage_bytes <= 40 * 1024)
    PCB_LAUNCH(ctx, pack_kernel<true>, grid_for(n, PACK_BLOCK), PACK_BLOCK, stage_bytes, seq, len, rows, n, stride, out.planes.p, out.len.p,
               flags.p, allow_other ? out.other.p : nullptr);
  else if (n > 0)
    PCB_LAUNCH(ctx, pack_kernel<false>, grid_for(n, PACK_BLOCK), PACK_BLOCK, 0, seq, len, rows, n, stride, out.planes.p, out.len.p, flags.p,
               allow_other ? out.other.p : nullptr);
}
void pack_finish(PackedBarcodes &out, const int32_t *h, bool allow_other) {
  if (h[0] != -1)
    throw Error(PCB_E_INPUT, "barcode " + std::to_string((uint32_t)h[0] - 1u) +
                                 (allow_other ? " is not 1..64 characters long" : " has a negative length or is longer than the rows of the matrix"));
  out.max_len = out.n ? h[2] : 0;                      // over the barcodes that fit the planes; 0 if none does
  out.min_len = out.max_len ? h[1] : 0;
  out.n_opaque = h[3];
  for (int m = 0; m < 66; m++) out.len_hist[m] = (uint32_t)h[4 + m];
  out.has_hist = true;
}

void pack_barcodes(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, const int32_t *rows, int32_t n,
                   int32_t stride, PackedBarcodes &out, bool allow_other) {
  DevBuf<int32_t> flags;
  pack_launch(ctx, seq, len, rows, n, stride, out, allow_other, flags);
  int32_t h[PACK_FLAGS];
  PCB_CUDA(cudaMemcpyAsync(h, flags.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  pack_finish(out, h, allow_other);
}

__device__ __forceinline__ uint32_t hash_barcode(ulonglong2 p, int32_t m) {
  uint64_t x = p.x * 0x9E3779B97F4A7C15ull ^ (p.y + (uint64_t)m) * 0xC2B2AE3D27D4EB4Full;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29;
  return (uint32_t)x;
}

// Open-addressing table over the barcodes.  A slot is 8 bytes: the smallest read index that landed in it so far -- which
// also names the slot's barcode, any read of a bucket does -- and how many reads did.  One 32-byte sector per probe.
struct BucketSlot { int32_t first; uint32_t count; };
constexpr int32_t SLOT_EMPTY = 0x7F7F7F7F;          // what cudaMemset(0x7F) leaves; larger than any read index

__global__ void bucket_insert_kernel(const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int32_t R,
                                     BucketSlot *slot, uint32_t mask, int32_t *__restrict__ slot_of_read) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const ulonglong2 p = planes[r];
  const int32_t m = len[r];
  uint32_t s = hash_barcode(p, m) & mask;
  for (;;) {
    int32_t f = *(volatile int32_t *)&slot[s].first;
    if (f == SLOT_EMPTY) {
      const int32_t old = atomicCAS(&slot[s].first, SLOT_EMPTY, r);
      f = old == SLOT_EMPTY ? r : old;
    }
    if (f == r) break;
    const ulonglong2 po = planes[f];
    if (po.x == p.x && po.y == p.y && len[f] == m) { if (r < f) atomicMin(&slot[s].first, r); break; }
    s = (s + 1) & mask;
  }
  atomicAdd(&slot[s].count, 1u);
  slot_of_read[r] = (int32_t)s;
}

// first reads of the buckets (= first-seen order of the barcodes)
__global__ void bucket_flag_first_kernel(const BucketSlot *__restrict__ slot, uint32_t cap, uint32_t *__restrict__ is_first) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const int32_t f = slot[s].first;
  if (f != SLOT_EMPTY) is_first[f] = 1u;
}

// rank[r] = number of first reads before r: for a first read that is its bucket index.  Per occupied slot: the bucket
// number replaces `first` in the slot; the bucket's size, packed barcode and length go to the bucket arrays.
__global__ void bucket_number_kernel(BucketSlot *__restrict__ slot, uint32_t cap, const uint32_t *__restrict__ rank,
                                     const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len,
                                     uint32_t *__restrict__ bucket_cnt, ulonglong2 *__restrict__ uplanes, int32_t *__restrict__ ulen) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const int32_t f = slot[s].first;
  if (f == SLOT_EMPTY) return;
  const uint32_t b = rank[f];
  slot[s].first = (int32_t)b;
  bucket_cnt[b] = slot[s].count - 0x7F7F7F7Fu;          // the table was filled with 0x7F bytes: counts started there
  if (uplanes) { uplanes[b] = planes[f]; ulen[b] = len[f]; }
}

// every read takes a place in its bucket's segment (in whatever order the atomics resolve; bucket_order_kernel fixes it)
__global__ void bucket_place_kernel(const int32_t *__restrict__ slot_of_read, const BucketSlot *__restrict__ slot, int32_t R,
                                    const int32_t *__restrict__ bucket_ptr, uint32_t *__restrict__ cursor,
                                    int32_t *__restrict__ bucket_of_read, int32_t *__restrict__ bucket_reads) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int32_t b = slot[slot_of_read[r]].first;
  bucket_of_read[r] = b;
  bucket_reads[bucket_ptr[b] + (int32_t)atomicAdd(&cursor[b], 1u)] = r;
}

// Arrival order inside a bucket = ascending read index.  Buckets of up to 32 reads (nearly all): one thread sorts its
// segment; up to 4096: on the `mid` list for one block each; more: on the `big` list, sorted by the host with the radix sort.
constexpr int ORDER_THREAD = 32, ORDER_BLOCK = 4096;
// counters: [0] = U (number of buckets), [1] = mid buckets, [2] = big buckets
__global__ void bucket_order_kernel(const uint32_t *__restrict__ rank, int32_t R, const int32_t *__restrict__ bucket_ptr,
                                    int32_t *__restrict__ bucket_reads, int32_t *__restrict__ counters, int32_t *__restrict__ mid,
                                    int32_t *__restrict__ big) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t U = (int32_t)rank[R];
  if (b == 0) counters[0] = U;
  if (b >= U) return;
  const int32_t lo = bucket_ptr[b], n = bucket_ptr[b + 1] - lo;
  if (n <= 1) return;
  if (n > ORDER_BLOCK) { big[atomicAdd(&counters[2], 1)] = b; return; }
  if (n > ORDER_THREAD) { mid[atomicAdd(&counters[1], 1)] = b; return; }
  int32_t v[ORDER_THREAD];
  for (int32_t i = 0; i < n; i++) {                      // insertion sort
    const int32_t x = bucket_reads[lo + i];
    int32_t k = i - 1;
    while (k >= 0 && v[k] > x) { v[k + 1] = v[k]; k--; }
    v[k + 1] = x;
  }
  for (int32_t i = 0; i < n; i++) bucket_reads[lo + i] = v[i];
}

__global__ void __launch_bounds__(256) bucket_order_mid_kernel(const int32_t *__restrict__ bucket_ptr, int32_t *__restrict__ bucket_reads,
                                                              const int32_t *__restrict__ counters, const int32_t *__restrict__ mid) {
  __shared__ int32_t sh[ORDER_BLOCK];
  const int32_t n_mid = counters[1];
  for (int32_t it = blockIdx.x; it < n_mid; it += gridDim.x) {
    const int32_t b = mid[it], lo = bucket_ptr[b], n = bucket_ptr[b + 1] - lo;
    int32_t m = 64;
    while (m < n) m <<= 1;
    for (int32_t i = threadIdx.x; i < m; i += blockDim.x) sh[i] = i < n ? bucket_reads[lo + i] : INT32_MAX;
    __syncthreads();
    for (int32_t k = 2; k <= m; k <<= 1)                 // bitonic sort
      for (int32_t j = k >> 1; j > 0; j >>= 1) {
        for (int32_t i = threadIdx.x; i < m; i += blockDim.x) {
          const int32_t l = i ^ j;
          if (l > i) {
            const int32_t a = sh[i], c = sh[l];
            const bool up = (i & k) == 0;
            if ((a > c) == up) { sh[i] = c; sh[l] = a; }
          }
        }
        __syncthreads();
      }
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) bucket_reads[lo + i] = sh[i];
    __syncthreads();
  }
}

__global__ void u32_keys_kernel(const int32_t *__restrict__ v, int64_t n, uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = (uint64_t)(uint32_t)v[i]; val[i] = 0u; }
}
__global__ void u32_unkeys_kernel(const uint64_t *__restrict__ key, int64_t n, int32_t *__restrict__ v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)key[i];
}

// One thread per (bucket, position): min(93, sum of Phred) -- the running min(93, old + new) of
// pacybara_precluster.py:37 equals the capped sum because every term is non-negative.
__global__ void bucket_qsum_kernel(const uint8_t *__restrict__ qual, const int32_t *__restrict__ len,
                                   const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads,
                                   int32_t U, int32_t stride, uint8_t *__restrict__ qsum) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)U * stride) return;
  int32_t b = (int32_t)(i / stride), pos = (int32_t)(i % stride);
  int32_t lo = bucket_ptr[b], hi = bucket_ptr[b + 1];
  int32_t m = len[bucket_reads[lo]];
  uint32_t acc = 0;
  if (pos < m)
    for (int32_t x = lo; x < hi; x++) {
      acc += (uint32_t)qual[(size_t)bucket_reads[x] * stride + pos] - 33u;
      if (acc >= 93u) { acc = 93u; if (x > lo) break; }
    }
  qsum[i] = (uint8_t)acc;
}

// Launches the bucketing of R reads; nothing is synchronised.  What the host has to know -- the packing flags and the
// counters (number of buckets, buckets left for the block / radix orderings) -- stays in st until bucket_finish().
void bucket_launch(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, int32_t R, int32_t stride, const BucketOut &out,
                   PackedBarcodes *unique_out, BucketPending &st) {
  HostTrace tr(ctx, "bucket_launch");
  st.R = R;
  pack_launch(ctx, seq, len, nullptr, R, stride, st.pk, false, st.flags);
  uint32_t cap = 64;
  while (cap < 2u * (uint32_t)R) cap <<= 1;
  tr.lap("pack");
  DevBuf<BucketSlot> slot(ctx, cap);
  DevBuf<int32_t> slot_of_read(ctx, R);
  DevBuf<uint32_t> rank(ctx, (size_t)R + 1), bucket_cnt(ctx, (size_t)R + 2), cursor(ctx, (size_t)R + 1);
  st.counters.alloc(ctx, 4); st.mid.alloc(ctx, (size_t)R / ORDER_THREAD + 1); st.big.alloc(ctx, (size_t)R / ORDER_BLOCK + 1);
  tr.lap("allocs");
  PCB_CUDA(cudaMemsetAsync(slot.p, 0x7F, (size_t)cap * sizeof(BucketSlot), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(rank.p, 0, ((size_t)R + 1) * sizeof(uint32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(bucket_cnt.p, 0, ((size_t)R + 2) * sizeof(uint32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(cursor.p, 0, ((size_t)R + 1) * sizeof(uint32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(st.counters.p, 0, 4 * sizeof(int32_t), ctx->stream));
  tr.lap("memsets");
  PCB_LAUNCH(ctx, bucket_insert_kernel, grid_for(R, 256), 256, 0, st.pk.planes.p, st.pk.len.p, R, slot.p, cap - 1, slot_of_read.p);
  PCB_LAUNCH(ctx, bucket_flag_first_kernel, grid_for(cap, 256), 256, 0, slot.p, cap, rank.p);
  exclusive_scan_u32(ctx, rank.p, rank.p, (size_t)R + 1);               // rank[R] = U
  tr.lap("insert..scan");
  if (unique_out) {                          // (a caller that keeps them across calls brings buffers that are large enough)
    if (unique_out->planes.n < (size_t)R) unique_out->planes.alloc(ctx, (size_t)R);
    if (unique_out->len.n < (size_t)R) unique_out->len.alloc(ctx, (size_t)R);
  }
  tr.lap("unique allocs");
  PCB_LAUNCH(ctx, bucket_number_kernel, grid_for(cap, 256), 256, 0, slot.p, cap, rank.p, st.pk.planes.p, st.pk.len.p, bucket_cnt.p,
             unique_out ? unique_out->planes.p : nullptr, unique_out ? unique_out->len.p : nullptr);
  exclusive_scan_u32(ctx, bucket_cnt.p, (uint32_t *)out.bucket_ptr, (size_t)R + 1);      // entries past U all hold R
  PCB_LAUNCH(ctx, bucket_place_kernel, grid_for(R, 256), 256, 0, slot_of_read.p, slot.p, R, out.bucket_ptr, cursor.p, out.bucket_of_read,
             out.bucket_reads);
  PCB_LAUNCH(ctx, bucket_order_kernel, grid_for(R, 128), 128, 0, rank.p, R, out.bucket_ptr, out.bucket_reads, st.counters.p, st.mid.p,
             st.big.p);
  PCB_LAUNCH(ctx, bucket_order_mid_kernel, (unsigned)ctx->sm_count, 256, 0, out.bucket_ptr, out.bucket_reads, st.counters.p, st.mid.p);
  tr.lap("number..order");
  PCB_CUDA(cudaMemcpyAsync(st.h_flags, st.flags.p, sizeof(st.h_flags), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(st.h_counters, st.counters.p, sizeof(st.h_counters), cudaMemcpyDeviceToHost, ctx->stream));
  tr.lap("d2h");
}

// After the stream has been synchronised: input errors, U, and the (rare) buckets of more than 4096 reads, whose
// segments go through the radix sort one by one.
int32_t bucket_finish(pcb_ctx *ctx, const BucketOut &out, PackedBarcodes *unique_out, BucketPending &st) {
  pack_finish(st.pk, st.h_flags, false);
  const int32_t U = st.h_counters[0], n_big = st.h_counters[2];
  if (n_big > 0) {
    std::vector<int32_t> h_big((size_t)n_big), h_ptr(2);
    PCB_CUDA(cudaMemcpyAsync(h_big.data(), st.big.p, (size_t)n_big * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PCB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int32_t b : h_big) {
      PCB_CUDA(cudaMemcpyAsync(h_ptr.data(), out.bucket_ptr + b, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      PCB_CUDA(cudaStreamSynchronize(ctx->stream));
      const int64_t n = h_ptr[1] - h_ptr[0];
      DevBuf<uint64_t> k0(ctx, n), k1(ctx, n);
      DevBuf<uint32_t> v0(ctx, n), v1(ctx, n);
      PCB_LAUNCH(ctx, u32_keys_kernel, grid_for(n, 256), 256, 0, out.bucket_reads + h_ptr[0], n, k0.p, v0.p);
      int w = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, n, bits_for((uint64_t)(st.R > 1 ? st.R - 1 : 1)));
      PCB_LAUNCH(ctx, u32_unkeys_kernel, grid_for(n, 256), 256, 0, w ? k1.p : k0.p, n, out.bucket_reads + h_ptr[0]);
    }
  }
  if (unique_out) {
    unique_out->n = U;
    unique_out->min_len = st.pk.min_len; unique_out->max_len = st.pk.max_len;
    // the histogram counts reads, not unique barcodes: same set of lengths, and only the choice of the usual length
    // (a performance matter for the pair search) looks at the counts
    for (int m = 0; m < 66; m++) unique_out->len_hist[m] = st.pk.len_hist[m];
    unique_out->has_hist = true;
  }
  ctx->stats[PCB_STAT_BUCKETS] = U;
  return U;
}

int32_t bucket_reads_dev(pcb_ctx *ctx, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int32_t R,
                         int32_t stride, const BucketOut &out, PackedBarcodes *unique_out) {
  if (R == 0) {
    PCB_CUDA(cudaMemsetAsync(out.bucket_ptr, 0, sizeof(int32_t), ctx->stream));
    if (unique_out) { unique_out->n = 0; unique_out->planes.alloc(ctx, 0); unique_out->len.alloc(ctx, 0); }
    return 0;
  }
  BucketPending st;
  bucket_launch(ctx, seq, len, R, stride, out, unique_out, st);
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  const int32_t U = bucket_finish(ctx, out, unique_out, st);
  if (out.bucket_qsum)
    PCB_LAUNCH(ctx, bucket_qsum_kernel, grid_for((size_t)U * stride, 256), 256, 0, qual, len, out.bucket_ptr,
               out.bucket_reads, U, stride, out.bucket_qsum);
  return U;
}

}  // namespace pcb
