// libmerge.cu -- row f4: merging two cluster libraries by virtual barcode and genotype.
// Replaces the row loop and the merge loop of src/pacybara_merge.R:48-101: every row of the second library is NEW (its
// virtual barcode is not in the first library), a CONFLICT (the barcode is there with other genotypes only: it is added
// as a row of its own and left to the collision filter) or a MERGE into the row of the first library that carries the
// same barcode and genotype -- when several do, into the one that is largest at that moment (which.max: the first of
// the largest), sizes growing as the loop goes.  Barcodes and genotypes arrive as dense ids (interned on the host);
// rows of either library are grouped by barcode id with the stable radix sort, and one thread per barcode replays the
// second library's rows of that barcode in their order -- the only order the result depends on.
// HBM bound: 12 B per row in, 8 B per row of the second library + 4 B per row of the first out, plus the sort.
#include "common.cuh"

namespace pcb {

__global__ void libmerge_keys_kernel(const int32_t *__restrict__ vbc, int32_t n, uint64_t *__restrict__ key, uint32_t *__restrict__ val) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = (uint64_t)(uint32_t)vbc[i]; val[i] = (uint32_t)i; }
}

// row pointers of rows sorted by barcode id: ptr[v] = first sorted position with id >= v
__global__ void libmerge_ptr_kernel(const uint64_t *__restrict__ key, int32_t n, int32_t n_ids, int32_t *__restrict__ ptr) {
  int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  const int32_t lo = p == 0 ? 0 : (int32_t)key[p - 1] + 1, hi = p == n ? n_ids : (int32_t)key[p];
  for (int32_t v = lo; v <= hi; v++) ptr[v] = p;
}

__global__ void libmerge_replay_kernel(int32_t n_ids, const int32_t *__restrict__ ptr1, const uint32_t *__restrict__ rows1,
                                       const int32_t *__restrict__ ptr2, const uint32_t *__restrict__ rows2,
                                       const int32_t *__restrict__ geno1, const int32_t *__restrict__ geno2,
                                       const int32_t *__restrict__ size2, int32_t *__restrict__ size1_out,
                                       int32_t *__restrict__ target2, int32_t *__restrict__ klass2) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_ids) return;
  const int32_t a_lo = ptr1[v], a_hi = ptr1[v + 1];
  for (int32_t y = ptr2[v]; y < ptr2[v + 1]; y++) {                 // rows of the second library, in their order
    const int32_t r2 = (int32_t)rows2[y];
    if (a_lo == a_hi) { klass2[r2] = PCB_LIBMERGE_NEW; target2[r2] = -1; continue; }          // :66-69
    int32_t best = -1, best_size = 0;
    for (int32_t x = a_lo; x < a_hi; x++) {                         // rows of the first library with this barcode, ascending
      const int32_t r1 = (int32_t)rows1[x];
      if (geno1[r1] != geno2[r2]) continue;
      const int32_t s = size1_out[r1];
      if (best < 0 || s > best_size) { best = r1; best_size = s; }  // which.max: the first of the largest (:93-95)
    }
    if (best < 0) { klass2[r2] = PCB_LIBMERGE_CONFLICT; target2[r2] = -1; continue; }          // :79-83
    klass2[r2] = PCB_LIBMERGE_MERGE; target2[r2] = best;                                       // :72-78
    size1_out[best] += size2[r2];                                                              // :96
  }
}

void merge_libraries_dev(pcb_ctx *ctx, int32_t n1, const int32_t *vbc1, const int32_t *geno1, const int32_t *size1, int32_t n2,
                         const int32_t *vbc2, const int32_t *geno2, const int32_t *size2, int32_t n_ids, int32_t *size1_out,
                         int32_t *target2, int32_t *klass2) {
  if (n1 > 0) PCB_CUDA(cudaMemcpyAsync(size1_out, size1, (size_t)n1 * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  if (n2 <= 0) return;
  const int bits = bits_for((uint64_t)(n_ids > 1 ? n_ids - 1 : 1));
  DevBuf<int32_t> ptr1(ctx, (size_t)n_ids + 1), ptr2(ctx, (size_t)n_ids + 1);
  DevBuf<uint64_t> k0(ctx, (size_t)(n1 > n2 ? n1 : n2)), k1(ctx, (size_t)(n1 > n2 ? n1 : n2));
  DevBuf<uint32_t> a0(ctx, (size_t)n1), a1(ctx, (size_t)n1), b0(ctx, (size_t)n2), b1(ctx, (size_t)n2);
  const uint32_t *rows1 = a0.p, *rows2 = b0.p;
  if (n1 > 0) {
    PCB_LAUNCH(ctx, libmerge_keys_kernel, grid_for(n1, 256), 256, 0, vbc1, n1, k0.p, a0.p);
    int w = radix_sort_pairs(ctx, k0.p, a0.p, k1.p, a1.p, n1, bits);
    rows1 = w ? a1.p : a0.p;
    PCB_LAUNCH(ctx, libmerge_ptr_kernel, grid_for((size_t)n1 + 1, 256), 256, 0, w ? k1.p : k0.p, n1, n_ids, ptr1.p);
  } else {
    PCB_CUDA(cudaMemsetAsync(ptr1.p, 0, ((size_t)n_ids + 1) * sizeof(int32_t), ctx->stream));
  }
  PCB_LAUNCH(ctx, libmerge_keys_kernel, grid_for(n2, 256), 256, 0, vbc2, n2, k0.p, b0.p);
  int w = radix_sort_pairs(ctx, k0.p, b0.p, k1.p, b1.p, n2, bits);
  rows2 = w ? b1.p : b0.p;
  PCB_LAUNCH(ctx, libmerge_ptr_kernel, grid_for((size_t)n2 + 1, 256), 256, 0, w ? k1.p : k0.p, n2, n_ids, ptr2.p);
  PCB_LAUNCH(ctx, libmerge_replay_kernel, grid_for(n_ids, 128), 128, 0, n_ids, ptr1.p, rows1, ptr2.p, rows2, geno1, geno2, size2, size1_out,
             target2, klass2);
}

}  // namespace pcb
