// Thin inline-PTX wrappers for the sm_100a tensor-core path: tcgen05.mma / commit / ld, TMEM allocation is done at the
// call sites, mbarriers, cp.async.bulk (1-D TMA), proxy fences.  Shared by fpn_tc.cuh and the tensor-core channel
// mixes of dstd_block.cuh.  Descriptor encodings follow the UMMA matrix / instruction descriptor layout (K-major,
// no-swizzle canonical layout: 8-row x 16-byte core matrices, LBO = byte distance of the two k-chunks of a K = 16
// step, SBO = byte distance of consecutive 8-row groups); profiles/micro/umma_probe.cu pins them on a B200.
#pragma once
#ifndef CISTGCN_EMU
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "simt.h"

namespace cg {

// ---- PTX wrappers -------------------------------------------------------------------------------------
CG_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
CG_DEV uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {   // K-major, no swizzle, version 1
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
CG_DEV constexpr uint32_t umma_idesc_bf16(int n) {      // fp32 accumulate, bf16 x bf16, K-major both, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
CG_DEV constexpr uint32_t umma_idesc_f16(int n) {       // fp32 accumulate, fp16 x fp16, K-major both, M = 128
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
CG_DEV void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
CG_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if ((spin & 0xFFF) == 0xFFF && clock64() - t0 > 8000000000LL) __trap();   // a lost arrival must not hang the GPU
  }
}
CG_DEV void mbar_wait_timed(uint32_t bar, uint32_t parity, long long& acc) {     // accounting only in -DCISTGCN_PROFILE builds
  const long long t0 = CG_CLOCK();
  mbar_wait(bar, parity);
  acc += CG_CLOCK() - t0;
}
CG_DEV void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
CG_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
CG_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
CG_DEV void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
CG_DEV void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
CG_DEV bool elect_one() {            // one lane of a converged warp; keeps the surrounding control flow warp-uniform
  uint32_t pred = 0;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
CG_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
CG_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CG_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
CG_DEV void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
CG_DEV void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
}
CG_DEV void tmem_ld4(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
CG_DEV void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// 8 consecutive channels of one position -> the bf16 term and the fp16 remainder, one 16-byte chunk row each
CG_DEV void split_store8(const float* x, unsigned char* dst, int term_stride) {
  uint32_t t1[4], t2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float x0 = x[2 * e], x1 = x[2 * e + 1];
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const float r0 = fminf(fmaxf(x0 - __low2float(h), -65504.f), 65504.f);
    const float r1 = fminf(fmaxf(x1 - __high2float(h), -65504.f), 65504.f);
    const __half2 m = __floats2half2_rn(r0, r1);
    t1[e] = *reinterpret_cast<const uint32_t*>(&h);
    t2[e] = *reinterpret_cast<const uint32_t*>(&m);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
  *reinterpret_cast<uint4*>(dst + term_stride) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
}

}  // namespace cg
#endif  // CISTGCN_EMU
