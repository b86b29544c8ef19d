// bucket.cu -- exact-barcode bucketing on the device (row a1 of the scope table).
//
// Replaces the dictionary of src/pacybara_precluster.py:28-39 and its first-seen output
// order (:68-76).  Reads are packed to 2-bit planes, inserted into an open-addressing table
// keyed by (planes, length), buckets are numbered by the smallest read index they contain
// (= first-seen order), reads are grouped by a stable radix sort (= arrival order inside each
// bucket), and the per-position qualities are summed with the reference's cap of 93.
// All of it is HBM-bound: per read W+4 bytes in (planes + length), 8 bytes out, plus
// 2 * stride bytes of quality traffic.
#include "common.cuh"
#include "myers.cuh"

namespace pcb {

// flags[0] = first offending row + 1 (0: none), flags[1] = min length, flags[2] = max length
__global__ void pack_kernel(const uint8_t *__restrict__ seq, const int32_t *__restrict__ len,
                            const int32_t *__restrict__ rows, int32_t n, int32_t stride,
                            ulonglong2 *__restrict__ planes, int32_t *__restrict__ out_len, int32_t *flags,
                            uint64_t *__restrict__ other) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t row = rows ? rows[i] : i;
  int32_t m = len[row];
  bool bad = m < 1 || m > 64 || m > stride;
  uint64_t L = 0, H = 0, X = 0;
  if (!bad) {
    const uint8_t *s = seq + (size_t)row * stride;
    for (int32_t t = 0; t < m; t++) {
      uint32_t c = base_code(s[t]);
      if (c > 3u) { bad |= other == nullptr; X |= 1ull << t; c = 0u; }
      L |= (uint64_t)(c & 1u) << t;
      H |= (uint64_t)((c >> 1) & 1u) << t;
    }
  }
  planes[i] = make_ulonglong2(L, H);
  if (other) other[i] = X;
  out_len[i] = m;
  if (bad) { atomicMin((uint32_t *)&flags[0], (uint32_t)i + 1u); return; }
  atomicMin(&flags[1], m);
  atomicMax(&flags[2], m);
}

void pack_barcodes(pcb_ctx *ctx, const uint8_t *seq, const int32_t *len, const int32_t *rows, int32_t n,
                   int32_t stride, PackedBarcodes &out, bool allow_other) {
  out.planes.alloc(ctx, (size_t)n);
  out.len.alloc(ctx, (size_t)n);
  if (allow_other) out.other.alloc(ctx, (size_t)n);
  out.n = n;
  DevBuf<int32_t> flags(ctx, 4);
  int32_t init[4] = {-1, 1 << 30, 0, 0};     // flags[0] as unsigned: 0xFFFFFFFF = none
  PCB_CUDA(cudaMemcpyAsync(flags.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  if (n > 0) PCB_LAUNCH(ctx, pack_kernel, grid_for(n, 128), 128, 0, seq, len, rows, n, stride, out.planes.p, out.len.p, flags.p,
                        allow_other ? out.other.p : nullptr);
  int32_t h[4];
  PCB_CUDA(cudaMemcpyAsync(h, flags.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h[0] != -1)
    throw Error(PCB_E_INPUT, "barcode " + std::to_string((uint32_t)h[0] - 1u) +
                                 (allow_other ? " is not 1..64 characters long" : " is not 1..64 bases over ACGT (no side path exists for such barcodes)"));
  out.min_len = n ? h[1] : 0;
  out.max_len = n ? h[2] : 0;
}

__device__ __forceinline__ uint32_t hash_barcode(ulonglong2 p, int32_t m) {
  uint64_t x = p.x * 0x9E3779B97F4A7C15ull ^ (p.y + (uint64_t)m) * 0xC2B2AE3D27D4EB4Full;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29;
  return (uint32_t)x;
}

// Every read finds (or claims) the table slot of its barcode; the slot remembers the smallest
// read index that landed in it and how many did.
__global__ void bucket_insert_kernel(const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len, int32_t R,
                                     int32_t *owner, int32_t *first, uint32_t *count, uint32_t mask,
                                     int32_t *__restrict__ slot_of_read) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  ulonglong2 p = planes[r];
  int32_t m = len[r];
  uint32_t s = hash_barcode(p, m) & mask;
  for (;;) {
    int32_t o = owner[s];
    if (o < 0) {
      o = atomicCAS(&owner[s], -1, r);
      if (o < 0) o = r;
    }
    if (o == r) break;
    ulonglong2 po = planes[o];
    if (po.x == p.x && po.y == p.y && len[o] == m) break;
    s = (s + 1) & mask;
  }
  atomicMin(&first[s], r);
  atomicAdd(&count[s], 1u);
  slot_of_read[r] = (int32_t)s;
}

__global__ void bucket_flag_first_kernel(const int32_t *__restrict__ slot_of_read, const int32_t *__restrict__ first, int32_t R,
                                         uint32_t *__restrict__ is_first) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) is_first[r] = first[slot_of_read[r]] == r ? 1u : 0u;
}

// rank[r] = number of first reads before r: for a first read that is its bucket index
__global__ void bucket_number_kernel(const int32_t *__restrict__ slot_of_read, const int32_t *__restrict__ first,
                                     const uint32_t *__restrict__ rank, const uint32_t *__restrict__ count, int32_t R,
                                     int32_t *__restrict__ bucket_of_read, uint32_t *__restrict__ bucket_cnt,
                                     uint64_t *__restrict__ sort_key) {
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int32_t s = slot_of_read[r];
  int32_t f = first[s];
  uint32_t b = rank[f];
  bucket_of_read[r] = (int32_t)b;
  sort_key[r] = b;
  if (f == r) bucket_cnt[b] = count[s];
}

__global__ void copy_u32_to_i32_kernel(const uint32_t *__restrict__ in, int32_t *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
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

__global__ void gather_first_kernel(const int32_t *__restrict__ bucket_ptr, const int32_t *__restrict__ bucket_reads, int32_t U,
                                    const ulonglong2 *__restrict__ planes, const int32_t *__restrict__ len,
                                    ulonglong2 *__restrict__ uplanes, int32_t *__restrict__ ulen) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= U) return;
  int32_t r = bucket_reads[bucket_ptr[b]];
  uplanes[b] = planes[r];
  ulen[b] = len[r];
}

int32_t bucket_reads_dev(pcb_ctx *ctx, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int32_t R,
                         int32_t stride, const BucketOut &out, PackedBarcodes *unique_out) {
  if (R == 0) {
    PCB_CUDA(cudaMemsetAsync(out.bucket_ptr, 0, sizeof(int32_t), ctx->stream));
    if (unique_out) { unique_out->n = 0; unique_out->planes.alloc(ctx, 0); unique_out->len.alloc(ctx, 0); }
    return 0;
  }
  PackedBarcodes pk;
  pack_barcodes(ctx, seq, len, nullptr, R, stride, pk);
  uint32_t cap = 64;
  while (cap < 2u * (uint32_t)R) cap <<= 1;
  DevBuf<int32_t> owner(ctx, cap), first(ctx, cap), slot_of_read(ctx, R);
  DevBuf<uint32_t> count(ctx, cap), rank(ctx, (size_t)R + 1), bucket_cnt(ctx, (size_t)R + 1);
  PCB_CUDA(cudaMemsetAsync(owner.p, 0xFF, cap * sizeof(int32_t), ctx->stream));
  fill_i32(ctx, first.p, cap, INT32_MAX);
  PCB_CUDA(cudaMemsetAsync(count.p, 0, cap * sizeof(uint32_t), ctx->stream));
  PCB_LAUNCH(ctx, bucket_insert_kernel, grid_for(R, 256), 256, 0, pk.planes.p, pk.len.p, R, owner.p, first.p, count.p,
             cap - 1, slot_of_read.p);
  PCB_LAUNCH(ctx, bucket_flag_first_kernel, grid_for(R, 256), 256, 0, slot_of_read.p, first.p, R, rank.p);
  // U = number of first reads: scan one element past the end (a zero) and read the total there
  PCB_CUDA(cudaMemsetAsync(rank.p + R, 0, sizeof(uint32_t), ctx->stream));
  exclusive_scan_u32(ctx, rank.p, rank.p, (size_t)R + 1);
  int32_t U = (int32_t)read_scalar(ctx, rank.p + R);
  DevBuf<uint64_t> k0(ctx, R), k1(ctx, R);
  DevBuf<uint32_t> v0(ctx, R), v1(ctx, R);
  PCB_CUDA(cudaMemsetAsync(bucket_cnt.p, 0, ((size_t)R + 1) * sizeof(uint32_t), ctx->stream));
  PCB_LAUNCH(ctx, bucket_number_kernel, grid_for(R, 256), 256, 0, slot_of_read.p, first.p, rank.p, count.p, R,
             out.bucket_of_read, bucket_cnt.p, k0.p);
  exclusive_scan_u32(ctx, bucket_cnt.p, bucket_cnt.p, (size_t)U + 1);
  PCB_LAUNCH(ctx, copy_u32_to_i32_kernel, grid_for((size_t)U + 1, 256), 256, 0, bucket_cnt.p, out.bucket_ptr, (size_t)U + 1);
  iota_u32(ctx, v0.p, R);
  int where = radix_sort_pairs(ctx, k0.p, v0.p, k1.p, v1.p, R, bits_for((uint64_t)(U > 1 ? U - 1 : 1)));
  PCB_LAUNCH(ctx, copy_u32_to_i32_kernel, grid_for(R, 256), 256, 0, where ? v1.p : v0.p, out.bucket_reads, (size_t)R);
  if (out.bucket_qsum)
    PCB_LAUNCH(ctx, bucket_qsum_kernel, grid_for((size_t)U * stride, 256), 256, 0, qual, len, out.bucket_ptr,
               out.bucket_reads, U, stride, out.bucket_qsum);
  if (unique_out) {
    unique_out->planes.alloc(ctx, (size_t)U);
    unique_out->len.alloc(ctx, (size_t)U);
    unique_out->n = U;
    unique_out->min_len = pk.min_len;
    unique_out->max_len = pk.max_len;
    PCB_LAUNCH(ctx, gather_first_kernel, grid_for(U, 256), 256, 0, out.bucket_ptr, out.bucket_reads, U, pk.planes.p,
               pk.len.p, unique_out->planes.p, unique_out->len.p);
  }
  ctx->stats[PCB_STAT_BUCKETS] = U;
  return U;
}

}  // namespace pcb
