// tagging.cu -- row f1: collision tags and the soft-filter dominance test on the cluster table.
// Replaces tagDups (src/pacybara_translate.R:58-68) and the dominance loop of src/pacybara_softfilter.R:46-63.
// Two streaming passes over the rows (HBM bound: 12 B in + 3 B out per row, plus the id-indexed counters).
#include "common.cuh"

namespace pcb {

__global__ void tag_count_kernel(const int32_t *__restrict__ row_bc, const int32_t *__restrict__ row_up, const int32_t *__restrict__ row_size,
                                 int32_t n_rows, uint32_t *__restrict__ bc_cnt, uint32_t *__restrict__ up_cnt,
                                 unsigned long long *__restrict__ up_total) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  int32_t b = row_bc[i], u = row_up[i];
  if (b >= 0) atomicAdd(&bc_cnt[b], 1u);
  if (u >= 0) { atomicAdd(&up_cnt[u], 1u); atomicAdd(&up_total[u], (unsigned long long)(row_size[i] > 0 ? row_size[i] : 0)); }
}

__global__ void tag_flag_kernel(const int32_t *__restrict__ row_bc, const int32_t *__restrict__ row_up, const int32_t *__restrict__ row_size,
                                int32_t n_rows, const uint32_t *__restrict__ bc_cnt, const uint32_t *__restrict__ up_cnt,
                                const unsigned long long *__restrict__ up_total, double min_ratio,
                                uint8_t *__restrict__ collision, uint8_t *__restrict__ up_collision, uint8_t *__restrict__ usable) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  int32_t b = row_bc[i], u = row_up[i], size = row_size[i];
  collision[i] = (b >= 0 && bc_cnt[b] > 1u) ? 1 : 0;
  up_collision[i] = (u >= 0 && up_cnt[u] > 1u) ? 1 : 0;
  double total = u >= 0 ? (double)up_total[u] : (double)size;
  usable[i] = (size > 1 && (double)size / total > min_ratio) ? 1 : 0;
}

void tag_rows_dev(pcb_ctx *ctx, int32_t n_rows, const int32_t *row_bc, const int32_t *row_up, const int32_t *row_size,
                  int32_t n_bc_ids, int32_t n_up_ids, double min_ratio, uint8_t *collision, uint8_t *up_collision, uint8_t *usable) {
  if (n_rows <= 0) return;
  DevBuf<uint32_t> bc_cnt(ctx, (size_t)n_bc_ids + 1), up_cnt(ctx, (size_t)n_up_ids + 1);
  DevBuf<unsigned long long> up_total(ctx, (size_t)n_up_ids + 1);
  PCB_CUDA(cudaMemsetAsync(bc_cnt.p, 0, ((size_t)n_bc_ids + 1) * sizeof(uint32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(up_cnt.p, 0, ((size_t)n_up_ids + 1) * sizeof(uint32_t), ctx->stream));
  PCB_CUDA(cudaMemsetAsync(up_total.p, 0, ((size_t)n_up_ids + 1) * sizeof(unsigned long long), ctx->stream));
  PCB_LAUNCH(ctx, tag_count_kernel, grid_for(n_rows, 256), 256, 0, row_bc, row_up, row_size, n_rows, bc_cnt.p, up_cnt.p, up_total.p);
  PCB_LAUNCH(ctx, tag_flag_kernel, grid_for(n_rows, 256), 256, 0, row_bc, row_up, row_size, n_rows, bc_cnt.p, up_cnt.p, up_total.p,
             min_ratio, collision, up_collision, usable);
}

}  // namespace pcb
