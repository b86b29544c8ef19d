// format.cu -- row a14, the writer (src/pacybara_cluster.py:546-567), as a ragged concatenation on the device.
// A row of clusters.csv.gz is `virtualBarcode,upBarcode,read|read|...,size,mut;mut;...`: nothing but slices of texts the
// path already holds -- read ids out of the FASTQ, mutation names out of genoExtract, barcodes out of the sequence
// matrix, uptags out of the uptag FASTQ -- glued with one-byte separators.  The host lists, per output item, which
// slice (token) it is and which separator follows; the device sizes the items, scans, and copies the bytes.  With a
// million read ids per library this replaces a Python loop over a million strings (0.5 s of the 1.4 s from files to
// clusters.csv.gz at cfg 2).  HBM bound: every output byte is read once and written once, plus 13 B per item.
#include "common.cuh"

namespace pcb {

__global__ void concat_len_kernel(const int32_t *__restrict__ tok_len, int64_t n_tok, int64_t n_src, const int64_t *__restrict__ tok_off,
                                  const int32_t *__restrict__ item_tok, const uint8_t *__restrict__ item_sep, int64_t n_items,
                                  int64_t *__restrict__ item_at, int32_t *__restrict__ bad) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_items) return;
  if (k == n_items) { item_at[k] = 0; return; }
  const int32_t t = item_tok[k];
  int64_t len = 0;
  if (t < 0 || (int64_t)t >= n_tok) atomicExch(bad, 1);
  else {
    len = tok_len[t];
    if (len < 0 || tok_off[t] < 0 || tok_off[t] + len > n_src) { atomicExch(bad, 2); len = 0; }
  }
  item_at[k] = len + (item_sep[k] ? 1 : 0);
}

__global__ void concat_copy_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ tok_off, const int32_t *__restrict__ tok_len,
                                   const int32_t *__restrict__ item_tok, const uint8_t *__restrict__ item_sep, int64_t n_items,
                                   const int64_t *__restrict__ item_at, int64_t out_cap, uint8_t *__restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_items) return;
  const int64_t at = item_at[k], end = item_at[k + 1];
  if (end > out_cap) return;                               // (the host reports the capacity error from the total)
  const int32_t t = item_tok[k];
  const int32_t len = tok_len[t];
  const uint8_t *s = src + tok_off[t];
  for (int32_t x = 0; x < len; x++) out[at + x] = s[x];
  if (item_sep[k]) out[at + len] = item_sep[k];
}

// returns the number of bytes the items need; writes them when they fit out_cap
int64_t concat_tokens_dev(pcb_ctx *ctx, const uint8_t *src, int64_t n_src, const int64_t *tok_off, const int32_t *tok_len, int64_t n_tok,
                          const int32_t *item_tok, const uint8_t *item_sep, int64_t n_items, uint8_t *out, int64_t out_cap) {
  if (n_items == 0) return 0;
  DevBuf<int64_t> item_at(ctx, (size_t)n_items + 1);
  DevBuf<int32_t> bad(ctx, 1);
  PCB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int32_t), ctx->stream));
  PCB_LAUNCH(ctx, concat_len_kernel, grid_for((size_t)n_items + 1, 256), 256, 0, tok_len, n_tok, n_src, tok_off, item_tok, item_sep, n_items,
             item_at.p, bad.p);
  exclusive_scan_i64(ctx, item_at.p, item_at.p, (size_t)n_items + 1);
  int64_t total = 0;
  int32_t h_bad = 0;
  PCB_CUDA(cudaMemcpyAsync(&total, item_at.p + n_items, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PCB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_bad == 1) throw Error(PCB_E_INPUT, "an item names a token that does not exist");
  if (h_bad == 2) throw Error(PCB_E_INPUT, "a token reaches outside the source bytes");
  if (total <= out_cap)
    PCB_LAUNCH(ctx, concat_copy_kernel, grid_for((size_t)n_items, 256), 256, 0, src, tok_off, tok_len, item_tok, item_sep, n_items, item_at.p,
               out_cap, out);
  return total;
}

}  // namespace pcb
