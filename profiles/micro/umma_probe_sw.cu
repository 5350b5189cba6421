// Probe: can the A operand of tcgen05.mma live in the SWIZZLE_128B K-major layout (128-byte rows) and still start at
// an arbitrary ROW (the tap shifts of the implicit-GEMM FPN kernel), and how fast are its operand reads compared with
// the no-swizzle layout (62.7 cycles per M128 K16 MMA, umma_probe.cu)?
//   A: [row][64 bf16] rows of 128 B, 16-byte chunk c of row r stored at chunk (c ^ (r & 7))  (absolute row index, base 1024-aligned)
//   B: no-swizzle [k-chunk][96 rows][8 bf16] as in the kernel
// Variants: start row shift in {0, 8, 3, 13}; descriptor base_offset 0 or (start_address >> 7) & 7.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe_sw umma_probe_sw.cu
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROWS = 160, NB = 96, KTOT = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t phase) {
  for (long long it = 0; it < 20000000LL; ++it) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// SWIZZLE_128B K-major: LBO field 1 (unused), SBO = 1024 B between 8-row groups, layout type 2 at bits 61-63
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A_log, const __nv_bfloat16* B_img, float* out,
                                                 int shift, int use_base_off, int nrep, long long* cycles, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                                        // ROWS x 128 B, 1024-aligned
  __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + ROWS * 128);
  __shared__ __align__(8) uint64_t bars[1];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_mma = smem_u32(&bars[0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_mma));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // A: logical [row][64] -> swizzled 16-byte chunks
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(A_log)[i];
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = tid; i < (KTOT / 8) * NB * 8; i += 128) sB[i] = B_img[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a0 = smem_u32(sA) + shift * 128;
  const uint32_t boff = use_base_off ? ((a0 >> 7) & 7) : 0;
  if (tid == 0) {
    for (int ks = 0; ks < KTOT / 16; ++ks)
      mma_bf16(tmem, desc_sw128(a0 + ks * 32, boff), desc_none(smem_u32(sB) + 2 * ks * NB * 16, NB * 16, 128), idesc, ks > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_mma) : "memory");
  }
  if (!mbar_wait(bar_mma, 0)) atomicOr(status, 2);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int blk = 0; blk < 3; ++blk) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + blk * 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * 96 + blk * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && nrep > 0) {           // rate: cycle through the 4 k-steps and 8 row starts, B descriptor varies too
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t0 = clock64();
    for (int i = 0; i < nrep; ++i) {
      const int ks = i & 3;
      mma_bf16(tmem, desc_sw128(a0 + ks * 32 + ((i >> 2) & 7) * 1024, boff), desc_none(smem_u32(sB) + 2 * ks * NB * 16, NB * 16, 128), idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_mma) : "memory");
    if (!mbar_wait(bar_mma, 1)) atomicOr(status, 4);
    cycles[0] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
}

int main() {
  std::vector<__nv_bfloat16> A(ROWS * KTOT), B((KTOT / 8) * NB * 8);
  std::vector<float> Af(ROWS * KTOT), Bf(NB * KTOT);
  srand(2);
  for (int i = 0; i < ROWS * KTOT; ++i) { float v = (float)(rand() % 7 - 3); Af[i] = v; A[i] = __float2bfloat16(v); }
  for (int n = 0; n < NB; ++n) for (int k = 0; k < KTOT; ++k) {
    float v = (float)(rand() % 5 - 2); Bf[n * KTOT + k] = v; B[((k / 8) * NB + n) * 8 + k % 8] = __float2bfloat16(v);
  }
  __nv_bfloat16 *dA, *dB; float* dO; long long* dC; int* dS;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dO, 128 * 96 * 4)); CK(cudaMalloc(&dC, 16)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = ROWS * 128 + (KTOT / 8) * NB * 16 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int shifts[4] = {0, 8, 3, 13};
  for (int bo = 0; bo < 2; ++bo) for (int si = 0; si < 4; ++si) {
    const int shift = shifts[si];
    CK(cudaMemset(dO, 0, 128 * 96 * 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dC, 0, 16));
    probe<<<1, 128, smem>>>(dA, dB, dO, shift, bo, 2048, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, bo, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(128 * 96); long long cyc; int st;
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 96; ++n) {
      float ref = 0;
      for (int k = 0; k < KTOT; ++k) ref += Af[(m + shift) * KTOT + k] * Bf[n * KTOT + k];
      if (fabs((double)ref - O[m * 96 + n]) > 1e-3) ++bad;
    }
    printf("SW128 A, row shift %2d, base_offset %s: status %d, mismatches %5d / 12288; %.2f cycles per M128 N96 K16 MMA\n",
           shift, bo ? "(addr>>7)&7" : "0", st, bad, cyc / 2048.0);
  }
  return 0;
}
