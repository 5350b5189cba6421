// Calibration microbenchmarks (not product code): achievable FFMA / LDS issue rates per SM as a
// function of resident warps, for the register-tiled inner loops used by the DSTD-GC kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_lds_bench fma_lds_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int TM, int TN>
__global__ void k_fma_only(float* out, int iters) {
  float acc[TM][TN], w[TM], x[TN];
  for (int i = 0; i < TM; ++i) w[i] = threadIdx.x * 0.001f + i;
  for (int j = 0; j < TN; ++j) x[j] = threadIdx.x * 0.002f + j;
  for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
    for (int j = 0; j < TN; ++j) x[j] += 1e-9f;     // keep the loop from being collapsed
  }
  float s = 0.f;
  for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// GEMM-like step: per k, TM/4 broadcast LDS.128 (weights) + one LDS.128/64 of TN activations, TM*TN FFMAs.
template <int TM, int TN, bool DB>
__global__ void k_gemm_step(float* out, int K, int Mp, int LD) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;                    // [K][Mp]
  float* X = sm + K * Mp;           // [K][LD]
  for (int i = threadIdx.x; i < K * Mp + K * LD; i += blockDim.x) sm[i] = 0.001f * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = (warp % (Mp / TM)) * TM, n0 = lane * TN;
  float acc[TM][TN];
  for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int rep = 0; rep < 64; ++rep) {
    const float* wp = W + m0;
    const float* xp = X + n0;
    if (!DB) {
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        float w[TM], x[TN];
#pragma unroll
        for (int q = 0; q < TM / 4; ++q) { float4 v = *(const float4*)(wp + 4 * q); w[4*q]=v.x; w[4*q+1]=v.y; w[4*q+2]=v.z; w[4*q+3]=v.w; }
        if (TN == 4) { float4 v = *(const float4*)xp; x[0]=v.x; x[1]=v.y; x[2]=v.z; x[3]=v.w; }
        else { for (int j = 0; j < TN; ++j) x[j] = xp[j]; }
        wp += Mp; xp += LD;
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
      }
    } else {
      float w0[TM], x0[TN], w1[TM], x1[TN];
      auto ld = [&](float (&w)[TM], float (&x)[TN]) {
#pragma unroll
        for (int q = 0; q < TM / 4; ++q) { float4 v = *(const float4*)(wp + 4 * q); w[4*q]=v.x; w[4*q+1]=v.y; w[4*q+2]=v.z; w[4*q+3]=v.w; }
        float4 v = *(const float4*)xp; x[0]=v.x; x[1]=v.y; x[2]=v.z; x[3]=v.w;
        wp += Mp; xp += LD;
      };
      auto fm = [&](float (&w)[TM], float (&x)[TN]) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
      };
      ld(w0, x0);
#pragma unroll 1
      for (int k = 0; k + 2 <= K; k += 2) { ld(w1, x1); fm(w0, x0); if (k + 2 < K) ld(w0, x0); fm(w1, x1); }
    }
  }
  float s = 0.f;
  for (int i = 0; i < TM; ++i) for (int j = 0; j < TN; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  int sms = 0, clk = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
  printf("SMs %d clock %d kHz\n", sms, clk);
  const int iters = 20000;
  for (int nt : {128, 256, 512, 1024}) {
    float ms = time_ms([&] { k_fma_only<8, 4><<<sms, nt>>>(out, iters); });
    double fma = (double)sms * nt * 32.0 * iters;
    printf("fma_only  8x4  warps/SM %2d : %.3f ms  -> %.1f FMA/clk/SM (at %.3f GHz nominal)\n", nt / 32, ms,
           fma / (ms * 1e-3) / sms / (clk * 1e3), clk * 1e-6);
    ms = time_ms([&] { k_fma_only<16, 4><<<sms, nt>>>(out, iters); });
    fma = (double)sms * nt * 64.0 * iters;
    printf("fma_only 16x4  warps/SM %2d : %.3f ms  -> %.1f FMA/clk/SM\n", nt / 32, ms, fma / (ms * 1e-3) / sms / (clk * 1e3));
  }
  const int K = 32, LD = 224;
  auto run = [&](const char* name, auto kern, int TM, int TN, int Mp, int nt) {
    size_t smem = (size_t)(K * Mp + K * LD) * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float ms = time_ms([&] { kern<<<sms, nt, smem>>>(out, K, Mp, LD); });
    double fma = (double)sms * nt * TM * TN * K * 64.0;
    printf("%-22s warps/SM %2d : %.3f ms -> %.1f FMA/clk/SM\n", name, nt / 32, ms, fma / (ms * 1e-3) / sms / (clk * 1e3));
  };
  for (int nt : {256, 512}) {
    run("gemm 16x4 plain", k_gemm_step<16, 4, false>, 16, 4, 64, nt);
    run("gemm 16x4 dblbuf", k_gemm_step<16, 4, true>, 16, 4, 64, nt);
    run("gemm  8x4 plain", k_gemm_step<8, 4, false>, 8, 4, 64, nt);
    run("gemm  8x4 dblbuf", k_gemm_step<8, 4, true>, 8, 4, 64, nt);
    run("gemm  8x1 plain", k_gemm_step<8, 1, false>, 8, 1, 64, nt);
    run("gemm  4x4 plain", k_gemm_step<4, 4, false>, 4, 4, 64, nt);
  }
  return 0;
}
