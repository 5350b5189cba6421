// Probe for the tcgen05 building blocks the tensor-core FPN kernel relies on (run on a B200):
//   * K-major, no-swizzle shared-memory descriptors over a [k-chunk][row][8 bf16] image whose rows sit at a
//     16-byte pitch (SBO = 128 B), so an operand may start at ANY row (the 3x3 taps become row shifts);
//   * cp.async.bulk (1-D TMA) + mbarrier complete_tx for the weight images;
//   * tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), M = 128, N = 32, tcgen05.commit, tcgen05.ld 32x32b;
//   * issue rate of back-to-back N = 32 MMAs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROWS = 160;          // A image rows (128 + max shift + slack)
constexpr int NB = 96;             // B rows: three 32-row blocks (the bf16 splits of one weight tile)
constexpr int KC = 2;              // 16-byte k-chunks per MMA (K = 16 bf16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
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

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A_img, const __nv_bfloat16* B_img, float* out,
                                                 int shift0, int shift1, int mode, int nrep, long long* cycles, int* status) {
  __shared__ __align__(128) __nv_bfloat16 sA[KC * ROWS * 8];
  __shared__ __align__(128) __nv_bfloat16 sB[1][KC * NB * 8];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_mma = smem_u32(&bars[0]), bar_tma = smem_u32(&bars[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_mma));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_tma));
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
  // A by ordinary stores (what the epilogue of the real kernel does), B by bulk copy
  for (int i = tid; i < KC * ROWS * 8; i += 128) sA[i] = A_img[i];
  if (tid == 0) {
    const uint32_t bytes = KC * NB * 8 * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_tma), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(&sB[0][0])), "l"(B_img), "r"(bytes), "r"(bar_tma) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // idesc: fp32 accumulate, bf16 x bf16, both K-major, N = 32, M = 128
  auto idesc_n = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
  const uint32_t a_chunk = ROWS * 16, b_chunk = NB * 16;
  uint32_t a_lbo = a_chunk, a_sbo = 128, b_lbo = b_chunk, b_sbo = 128;
  if (tid == 0) {
    if (!mbar_wait(bar_tma, 0)) atomicOr(status, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t dbq = make_desc(smem_u32(&sB[0][0]), b_lbo, b_sbo);
    mma_bf16(tmem, make_desc(smem_u32(sA) + shift0 * 16, a_lbo, a_sbo), dbq, idesc_n(96), 0);
    mma_bf16(tmem, make_desc(smem_u32(sA) + shift1 * 16, a_lbo, a_sbo), dbq, idesc_n(64), 1);
    mma_bf16(tmem, make_desc(smem_u32(sA) + (shift0 + shift1 + 3) * 16, a_lbo, a_sbo), dbq, idesc_n(32), 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_mma) : "memory");
  }
  if (!mbar_wait(bar_mma, 0)) atomicOr(status, 2);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t v[32];
    for (int blk = 0; blk < 3; ++blk) {
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // issue-rate measurement: nrep dependent-accumulate MMAs, one commit
  if (tid == 0 && nrep > 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t da = make_desc(smem_u32(sA), a_lbo, a_sbo), db = make_desc(smem_u32(&sB[0][0]), b_lbo, b_sbo);
    const long long t0 = clock64();
    for (int i = 0; i < nrep; ++i) {
      const uint64_t d2 = da + (uint64_t)(i & 7);
      if (mode == 0) { mma_bf16(tmem, d2, db, idesc_n(96), 1); mma_bf16(tmem, d2 + 9, db, idesc_n(64), 1); mma_bf16(tmem, d2 + 18, db, idesc_n(32), 1); }
      else mma_bf16(tmem, d2, db, idesc_n(mode), 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_mma) : "memory");
    const long long t1 = clock64();
    if (!mbar_wait(bar_mma, 1)) atomicOr(status, 4);
    const long long t2 = clock64();
    cycles[0] = t1 - t0; cycles[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
}

int main() {
  std::vector<__nv_bfloat16> A(KC * ROWS * 8), B(KC * NB * 8);
  std::vector<float> Af(ROWS * 16), Bf(NB * 16);
  srand(1);
  for (int r = 0; r < ROWS; ++r) for (int k = 0; k < 16; ++k) {
    float v = (float)(rand() % 7 - 3); Af[r * 16 + k] = v; A[((k / 8) * ROWS + r) * 8 + k % 8] = __float2bfloat16(v);
  }
  for (int n = 0; n < NB; ++n) for (int k = 0; k < 16; ++k) {
    float v = (float)(rand() % 5 - 2); Bf[n * 16 + k] = v; B[((k / 8) * NB + n) * 8 + k % 8] = __float2bfloat16(v);
  }
  __nv_bfloat16 *dA, *dB; float* dO; long long* dC; int* dS;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dO, 128 * 96 * 4)); CK(cudaMalloc(&dC, 16)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const int s0 = 0, s1 = 7;
  const int modes[4] = {0, 96, 64, 32};
  for (int mi = 0; mi < 4; ++mi) {
    const int mode = modes[mi];
    CK(cudaMemset(dO, 0, 128 * 96 * 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dC, 0, 16));
    probe<<<1, 128>>>(dA, dB, dO, s0, s1, mode, 2048, dC, dS);
    CK(cudaDeviceSynchronize());
    std::vector<float> O(128 * 96); long long cyc[2]; int st;
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    const int sh[3] = {s0, s1, s0 + s1 + 3};
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 96; ++n) {
      float ref = 0;
      for (int j = 0; j < 3; ++j) { if (n >= 96 - 32 * j) continue; for (int k = 0; k < 16; ++k) ref += Af[(m + sh[j]) * 16 + k] * Bf[n * 16 + k]; }
      double e = fabs((double)ref - O[m * 96 + n]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad;
    }
    printf("N = 96/64/32 triple: status %d, mismatches %d / %d, max err %.3g; timing mode %d x 2048: issue %lld cyc, complete %lld cyc (%.2f cyc/iteration)\n",
           st, bad, 128 * 96, maxerr, mode, cyc[0], cyc[1], cyc[1] / 2048.0);
  }
  return 0;
}
