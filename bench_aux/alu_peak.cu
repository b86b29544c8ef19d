// alu_peak.cu -- integer-ALU issue-rate microbenchmark (bench.py only; not part of the product).
// MEASURED_PEAKS.json has no INT32 figure, so the denominator of the distance kernel's ALU
// roofline is measured here at the clocks of the run.  Two kernels, instruction counts pinned
// with inline PTX (one lop3.b32 = one LOP3.LUT, one mad.lo = one IMAD):
//   mode 0: 8 independent LOP3 chains per thread          -> peak of the ALU pipe
//   mode 1: LOP3 chains interleaved with IMAD chains      -> peak with both INT-capable pipes
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) alu_chain_kernel(uint32_t *out, uint32_t seed, int iters) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1 + i); b[i] = seed ^ (blockIdx.x + 17 * i); }
  uint32_t c = seed | 1u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || (i & 1) == 0)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
      else
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b[i]));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// lane-instructions per second (32 per warp instruction) for the given mode
extern "C" int alu_peak_measure(int device, int mode, double *lane_instr_per_s, double *ms_out) {
  cudaSetDevice(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int blocks = sms * 8, threads = 256, iters = 8192;
  uint32_t *out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(uint32_t)) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    if (mode == 0) alu_chain_kernel<0><<<blocks, threads>>>(out, 0x9E3779B9u + rep, iters);
    else           alu_chain_kernel<1><<<blocks, threads>>>(out, 0x9E3779B9u + rep, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) return 2;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  *lane_instr_per_s = (double)blocks * threads * (double)iters * 8.0 / (best * 1e-3);
  *ms_out = best;
  cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}
