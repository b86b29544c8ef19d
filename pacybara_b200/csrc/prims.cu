// prims.cu -- hand-written device primitives used by every stage: exclusive scan and a
// stable LSD radix sort.  Both are HBM-bound streaming kernels (bytes per item).
#include "common.cuh"

namespace pcb {

// ------------------------------------------------------------------------------------ scan
// Reduce-then-scan over tiles of 2048 items: (1) tile sums, (2) recursive scan of the sums,
// (3) tile-local scan seeded with the tile offset.  2 reads + 1 write per item.
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_sums(const T *__restrict__ in, T *__restrict__ sums, size_t n) {
  __shared__ T warp_sum[SCAN_BLOCK / 32];
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  T s = 0;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_BLOCK) {
    size_t idx = base + i;
    if (idx < n) s += in[idx];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    T t = 0;
    for (int w = 0; w < SCAN_BLOCK / 32; w++) t += warp_sum[w];
    sums[blockIdx.x] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tiles(const T *in, T *out, const T *__restrict__ tile_off, size_t n) {
  __shared__ T sh[SCAN_TILE];
  __shared__ T warp_tot[SCAN_BLOCK / 32];
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_BLOCK) {
    size_t idx = base + i;
    sh[i] = idx < n ? in[idx] : (T)0;
  }
  __syncthreads();
  T loc[SCAN_ITEMS];
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { loc[k] = sh[threadIdx.x * SCAN_ITEMS + k]; s += loc[k]; }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  T run = tile_off[blockIdx.x] + inc - s;
  for (int w = 0; w < warp; w++) run += warp_tot[w];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { sh[threadIdx.x * SCAN_ITEMS + k] = run; run += loc[k]; }
  __syncthreads();
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_BLOCK) {
    size_t idx = base + i;
    if (idx < n) out[idx] = sh[i];
  }
}

template <typename T>
static void exclusive_scan(pcb_ctx *ctx, const T *in, T *out, size_t n) {
  if (n == 0) return;
  size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  DevBuf<T> sums(ctx, tiles);
  if (tiles == 1) {
    PCB_CUDA(cudaMemsetAsync(sums.p, 0, sizeof(T), ctx->stream));
  } else {
    PCB_LAUNCH(ctx, scan_tile_sums<T>, (unsigned)tiles, SCAN_BLOCK, 0, in, sums.p, n);
    exclusive_scan<T>(ctx, sums.p, sums.p, tiles);
  }
  PCB_LAUNCH(ctx, scan_tiles<T>, (unsigned)tiles, SCAN_BLOCK, 0, in, out, sums.p, n);
}

// Single-pass scan for 32-bit counts (the common case: every sort pass, every CSR build): tiles take a
// ticket, publish their aggregate, and look back over the descriptors of the tiles before them until they
// meet an inclusive prefix ("decoupled look-back").  One read + one write per item, one launch.
// Descriptor: status in the high word (0 = nothing yet, 1 = aggregate, 2 = inclusive prefix), value in the low.
__global__ void __launch_bounds__(SCAN_BLOCK) scan_lookback_u32(const uint32_t *in, uint32_t *out, size_t n,
                                                                 unsigned long long *desc, unsigned int *ticket) {
  __shared__ uint32_t sh[SCAN_TILE];
  __shared__ uint32_t warp_tot[SCAN_BLOCK / 32];
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const size_t base = (size_t)tile * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_BLOCK) {
    size_t idx = base + i;
    sh[i] = idx < n ? in[idx] : 0u;
  }
  __syncthreads();
  uint32_t loc[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { loc[k] = sh[threadIdx.x * SCAN_ITEMS + k]; s += loc[k]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t aggregate = 0;
    for (int w = 0; w < SCAN_BLOCK / 32; w++) aggregate += warp_tot[w];
    uint32_t running = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&desc[0], (2ull << 32) | aggregate);
    } else {
      if (lane == 0) atomicExch(&desc[tile], (1ull << 32) | aggregate);
      int look = (int)tile - 1;
      for (;;) {
        const int idx = look - lane;
        unsigned long long d = idx >= 0 ? *(volatile unsigned long long *)&desc[idx] : (2ull << 32);
        while (__any_sync(0xFFFFFFFFu, (d >> 32) == 0)) {
          if ((d >> 32) == 0) d = *(volatile unsigned long long *)&desc[idx];
        }
        const uint32_t status = (uint32_t)(d >> 32), val = (uint32_t)d;
        const unsigned pmask = __ballot_sync(0xFFFFFFFFu, status == 2u);
        const int first = pmask ? __ffs(pmask) - 1 : 31;           // nearest tile that already holds a prefix
        uint32_t v = lane <= first ? val : 0u;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        running += v;
        if (pmask) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(&desc[tile], (2ull << 32) | (unsigned long long)(running + aggregate));
    }
    if (lane == 0) s_prefix = running;
  }
  __syncthreads();
  uint32_t run = s_prefix + inc - s;
  for (int w = 0; w < warp; w++) run += warp_tot[w];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { sh[threadIdx.x * SCAN_ITEMS + k] = run; run += loc[k]; }
  __syncthreads();
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_BLOCK) {
    size_t idx = base + i;
    if (idx < n) out[idx] = sh[i];
  }
}

void exclusive_scan_u32(pcb_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n) {
  if (n == 0) return;
  // Measured (r01g): the single launch wins while launch latency dominates (cfg 2: 168 -> 115 launches per step,
  // -0.25 ms); on the 20 M-item scans of cfg 5 the three-kernel form was faster, so large inputs keep it.
  if (n > ((size_t)1 << 21)) { exclusive_scan<uint32_t>(ctx, in, out, n); return; }
  size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  DevBuf<unsigned long long> desc(ctx, tiles + 1);                 // last word: the ticket counter
  PCB_CUDA(cudaMemsetAsync(desc.p, 0, (tiles + 1) * sizeof(unsigned long long), ctx->stream));
  PCB_LAUNCH(ctx, scan_lookback_u32, (unsigned)tiles, SCAN_BLOCK, 0, in, out, n, desc.p, (unsigned int *)(desc.p + tiles));
}
void exclusive_scan_i64(pcb_ctx *ctx, const int64_t *in, int64_t *out, size_t n) { exclusive_scan<int64_t>(ctx, in, out, n); }

// ------------------------------------------------------------------------------------ sort
// 8 bits per pass.  A block owns a tile of 4096 consecutive items; warp w owns the w-th 512 of
// them and walks them 32 at a time, so "tile order, then warp order, then row order" is the
// input order and the sort is stable.  Per pass: histogram kernel -> scan of the
// [digit][block] table -> scatter kernel that ranks with __match_any_sync.
constexpr int SORT_BLOCK = 256;
constexpr int SORT_WARPS = SORT_BLOCK / 32;
// Rows of 32 items per warp and tile: 16 (tile 4096) keeps the [digit][block] table small for big inputs; small
// inputs take 4 (tile 1024) so that a 150 k-item sort still spreads over the whole chip instead of 37 blocks.
template <int ROWS>
__device__ __forceinline__ void sort_count_tile(const uint64_t *__restrict__ keys, size_t n, int shift,
                                                uint32_t (*cnt)[256]) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < SORT_WARPS * 256; d += SORT_BLOCK) (&cnt[0][0])[d] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * (SORT_BLOCK * ROWS) + (size_t)warp * (32 * ROWS);
#pragma unroll 4
  for (int r = 0; r < ROWS; r++) {
    size_t idx = base + r * 32 + lane;
    if (idx < n) atomicAdd(&cnt[warp][(keys[idx] >> shift) & 0xFF], 1u);
  }
  __syncthreads();
}

template <int ROWS>
__global__ void __launch_bounds__(SORT_BLOCK) sort_hist(const uint64_t *__restrict__ keys, size_t n, int shift,
                                                        uint32_t *__restrict__ ghist, uint32_t nblocks) {
  __shared__ uint32_t cnt[SORT_WARPS][256];
  sort_count_tile<ROWS>(keys, n, shift, cnt);
  uint32_t t = 0;
  for (int w = 0; w < SORT_WARPS; w++) t += cnt[w][threadIdx.x];
  ghist[(size_t)threadIdx.x * nblocks + blockIdx.x] = t;
}

// The [digit][block] table needs no general scan: one block per digit turns its row into exclusive prefixes over the
// tiles and leaves the digit's total; the scatter kernel adds the (256-entry) prefix over the digits itself.
constexpr int ROWSCAN_BLOCK = 256;
__global__ void __launch_bounds__(ROWSCAN_BLOCK) sort_rowscan(uint32_t *__restrict__ ghist, uint32_t nblocks,
                                                              uint32_t *__restrict__ digit_total) {
  __shared__ uint32_t warp_tot[ROWSCAN_BLOCK / 32];
  uint32_t *row = ghist + (size_t)blockIdx.x * nblocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t running = 0;
  for (uint32_t base = 0; base < nblocks; base += ROWSCAN_BLOCK) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? row[i] : 0u, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < ROWSCAN_BLOCK / 32; w++) { uint32_t t = warp_tot[w]; if (w < warp) before += t; all += t; }
    if (i < nblocks) row[i] = running + before + inc - v;
    running += all;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = running;
}

template <int ROWS>
__global__ void __launch_bounds__(SORT_BLOCK) sort_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                           uint64_t *__restrict__ okeys, uint32_t *__restrict__ ovals, size_t n,
                                                           int shift, const uint32_t *__restrict__ goff, uint32_t nblocks,
                                                           const uint32_t *__restrict__ digit_total) {
  __shared__ uint32_t cnt[SORT_WARPS][256];
  __shared__ uint32_t dig_warp[SORT_WARPS];
  // exclusive prefix of the digit totals: thread d owns digit d (SORT_BLOCK == 256)
  uint32_t dig_base;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = digit_total[threadIdx.x], inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) dig_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += dig_warp[w];
    dig_base = before + inc - v;
  }
  sort_count_tile<ROWS>(keys, n, shift, cnt);
  {
    uint32_t run = dig_base + goff[(size_t)threadIdx.x * nblocks + blockIdx.x];
    for (int w = 0; w < SORT_WARPS; w++) { uint32_t c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  size_t base = (size_t)blockIdx.x * (SORT_BLOCK * ROWS) + (size_t)warp * (32 * ROWS);
  for (int r = 0; r < ROWS; r++) {
    size_t idx = base + r * 32 + lane;
    bool valid = idx < n;
    uint64_t key = valid ? keys[idx] : 0;
    uint32_t val = valid ? vals[idx] : 0;
    uint32_t digit = valid ? (uint32_t)((key >> shift) & 0xFF) : 0xFFFFFFFFu;
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = valid ? cnt[warp][digit] + rank : 0;
    __syncwarp();
    if (valid && rank == 0) cnt[warp][digit] += __popc(peers);
    __syncwarp();
    if (valid) { okeys[pos] = key; ovals[pos] = val; }
  }
}

template <int ROWS>
static int radix_sort_rows(pcb_ctx *ctx, uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, size_t n, int bits) {
  const size_t tile = (size_t)SORT_BLOCK * ROWS;
  uint32_t nblocks = (uint32_t)((n + tile - 1) / tile);
  static_assert(SORT_BLOCK == 256, "one thread per digit");
  DevBuf<uint32_t> ghist(ctx, (size_t)256 * nblocks), digit_total(ctx, 256);
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    uint64_t *ki = cur ? k1 : k0, *ko = cur ? k0 : k1;
    uint32_t *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
    PCB_LAUNCH(ctx, sort_hist<ROWS>, nblocks, SORT_BLOCK, 0, ki, n, shift, ghist.p, nblocks);
    PCB_LAUNCH(ctx, sort_rowscan, 256, ROWSCAN_BLOCK, 0, ghist.p, nblocks, digit_total.p);
    PCB_LAUNCH(ctx, sort_scatter<ROWS>, nblocks, SORT_BLOCK, 0, ki, vi, ko, vo, n, shift, ghist.p, nblocks, digit_total.p);
    cur ^= 1;
  }
  return cur;
}

int radix_sort_pairs(pcb_ctx *ctx, uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, size_t n, int bits) {
  if (n == 0) return 0;
  if (n <= ((size_t)1 << 21)) return radix_sort_rows<4>(ctx, k0, v0, k1, v1, n, bits);
  return radix_sort_rows<16>(ctx, k0, v0, k1, v1, n, bits);
}

// ------------------------------------------------------------------------------------ small
template <typename T>
__global__ void fill_kernel(T *p, size_t n, T v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void iota_kernel(uint32_t *p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}
void fill_i32(pcb_ctx *ctx, int32_t *p, size_t n, int32_t v) {
  if (n) PCB_LAUNCH(ctx, fill_kernel<int32_t>, grid_for(n, 256), 256, 0, p, n, v);
}
void fill_i64(pcb_ctx *ctx, int64_t *p, size_t n, int64_t v) {
  if (n) PCB_LAUNCH(ctx, fill_kernel<int64_t>, grid_for(n, 256), 256, 0, p, n, v);
}
void iota_u32(pcb_ctx *ctx, uint32_t *p, size_t n) {
  if (n) PCB_LAUNCH(ctx, iota_kernel, grid_for(n, 256), 256, 0, p, n);
}

}  // namespace pcb
