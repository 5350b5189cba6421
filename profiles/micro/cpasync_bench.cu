// Calibration: how long does a block-cooperative cp.async fetch of an L2-resident chunk take when every
// SM asks for the SAME lines (weights shared by all persistent CTAs) vs DIFFERENT lines?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp16(float* s, const float* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(g));
}

__global__ void k_fetch(const float* src, long long* out, int chunk_floats, int nchunks, int same, int reps) {
  extern __shared__ __align__(16) float sm[];
  const float* base = src + (same ? 0 : (size_t)blockIdx.x * chunk_floats * nchunks);
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    for (int c = 0; c < nchunks; ++c) {
      for (int i = threadIdx.x * 4; i < chunk_floats; i += blockDim.x * 4) cp16(sm + c * chunk_floats + i, base + c * chunk_floats + i);
      asm volatile("cp.async.commit_group;\n" ::);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = total / reps;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* src; long long* out; long long h[256];
  cudaMalloc(&src, (size_t)sms * 8 * 8192 * 4); cudaMemset(src, 0, (size_t)sms * 8 * 8192 * 4);
  cudaMalloc(&out, 256 * 8);
  cudaFuncSetAttribute(k_fetch, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192 * 4);
  for (int nt : {256, 512})
    for (int same : {1, 0})
      for (int chunk : {2048, 8192})
        for (int n : {1, 2, 4}) {
          if ((size_t)chunk * n * 4 > 200 * 1024) continue;
          k_fetch<<<sms, nt, (size_t)chunk * n * 4>>>(src, out, chunk, n, same, 50);
          cudaMemcpy(h, out, sms * 8, cudaMemcpyDeviceToHost);
          long long mx = 0, sum = 0; for (int i = 0; i < sms; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
          printf("threads %d %s chunk %5d floats x %d : avg %lld max %lld cycles  (%.1f B/clk/SM)\n", nt, same ? "SAME" : "DIFF",
                 chunk, n, sum / sms, mx, (double)chunk * n * 4 / (double)(sum / sms));
        }
  return 0;
}
