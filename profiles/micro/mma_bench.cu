// Calibration: issue rate of the legacy warp-level tensor-core path (mma.sync) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_mma_tf32(float* out, int iters) {
  float c[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_mma_bf16(float* out, int iters) {
  float c[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f003f00u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f003f00u + threadIdx.x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  int sms = 0, clk = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  for (int nt : {128, 256, 512}) {
    float ms = time_ms([&] { k_mma_tf32<8><<<sms, nt>>>(out, iters); });
    double mma = (double)sms * (nt / 32) * 8.0 * iters;
    printf("tf32 m16n8k8  warps/SM %2d: %.3f ms -> %.3f MMA/clk/SM = %.0f MAC/clk/SM (3xTF32 effective %.0f)\n", nt / 32, ms,
           mma / (ms * 1e-3) / sms / (clk * 1e3), 1024 * mma / (ms * 1e-3) / sms / (clk * 1e3), 1024 * mma / (ms * 1e-3) / sms / (clk * 1e3) / 3);
    ms = time_ms([&] { k_mma_bf16<8><<<sms, nt>>>(out, iters); });
    printf("bf16 m16n8k16 warps/SM %2d: %.3f ms -> %.3f MMA/clk/SM = %.0f MAC/clk/SM\n", nt / 32, ms,
           mma / (ms * 1e-3) / sms / (clk * 1e3), 2048 * mma / (ms * 1e-3) / sms / (clk * 1e3));
  }
  return 0;
}
