// Portability shim: the kernels are plain CUDA C++; under -DCISTGCN_EMU (tests/emu only) the same
// sources are compiled by g++ against a SIMT emulator so kernel logic can be checked without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef CISTGCN_EMU
#include "simt_emu.h"
#define CG_LAUNCH(kfn, grid, block, smem, stream, ...) \
  simt_emu::launch(dim3(grid), dim3(block), (smem), [=]() { kfn(__VA_ARGS__); })
#define CG_DYN_SMEM(name) float* name = reinterpret_cast<float*>(simt_emu::dyn_smem())
#define CG_HOST_ONLY_CUDA(...)
#else
#include <cuda_runtime.h>
#define CG_LAUNCH(kfn, grid, block, smem, stream, ...) \
  kfn<<<dim3(grid), dim3(block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define CG_DYN_SMEM(name) extern __shared__ __align__(16) float name[]
#define CG_HOST_ONLY_CUDA(...) __VA_ARGS__
#endif

#define CG_DEV __device__ __forceinline__

// Cycle accounting of the debug hooks exists only in -DCISTGCN_PROFILE builds; elsewhere it folds to constants.
#if defined(CISTGCN_PROFILE) && !defined(CISTGCN_EMU)
#define CG_CLOCK() clock64()
#else
#define CG_CLOCK() 0LL
#endif

namespace cg {

// Launders a pointer so that the compiler cannot prove loads through it loop-invariant (it would hoist a whole
// weight matrix into registers -- and spill it -- out of a loop whose iterations re-read the same shared-memory rows).
template <class T>
CG_DEV const T* opaque_ptr(const T* p) {
#ifndef CISTGCN_EMU
  asm volatile("" : "+l"(p));
#endif
  return p;
}

// Same purpose for SHARED-memory operands: an opaque zero to add to the pointer.  Laundering the pointer itself makes it a
// generic pointer (LD with 64-bit address arithmetic instead of LDS); laundering an integer offset keeps its address space.
CG_DEV int opaque_zero() {
  int z = 0;
#ifndef CISTGCN_EMU
  asm volatile("" : "+r"(z));
#endif
  return z;
}

// ---- bf16 storage of inter-kernel activations (cistgcn_forward_bf16): plain bit manipulation, no cuda_bf16.h, so the
// same code runs under the SIMT emulator.  Round-to-nearest-even; NaN payloads are not preserved (activations are finite).
CG_DEV unsigned f32_bits(float v) { unsigned u; memcpy(&u, &v, 4); return u; }
CG_DEV float bits_f32(unsigned u) { float v; memcpy(&v, &u, 4); return v; }
CG_DEV unsigned short f32_to_bf16(float v) {
  unsigned u = f32_bits(v);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
CG_DEV float bf16_to_f32(unsigned short h) { return bits_f32((unsigned)h << 16); }
CG_DEV unsigned pack_bf16x2(float lo, float hi) { return (unsigned)f32_to_bf16(lo) | ((unsigned)f32_to_bf16(hi) << 16); }
// element i of an activation tensor stored as fp32 or bf16
CG_DEV float ld_act(const float* base, size_t i, bool bf16) {
  return bf16 ? bf16_to_f32(reinterpret_cast<const unsigned short*>(base)[i]) : __ldg(base + i);
}
CG_DEV void st_act(float* base, size_t i, float v, bool bf16) {
  if (bf16) reinterpret_cast<unsigned short*>(base)[i] = f32_to_bf16(v);
  else base[i] = v;
}
// four consecutive elements starting at element i (i % 4 == 0, base 16-byte aligned)
CG_DEV float4 ld_act4(const float* base, size_t i, bool bf16) {
  if (!bf16) return __ldg(reinterpret_cast<const float4*>(base + i));
  const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(base) + i);
  return make_float4(bits_f32(q.x << 16), bits_f32(q.x & 0xFFFF0000u), bits_f32(q.y << 16), bits_f32(q.y & 0xFFFF0000u));
}
CG_DEV void st_act4(float* base, size_t i, float4 v, bool bf16) {
  if (!bf16) { *reinterpret_cast<float4*>(base + i) = v; return; }
  uint2 q;
  q.x = pack_bf16x2(v.x, v.y);
  q.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base) + i) = q;
}

CG_DEV float prelu(float v, float a) { return v >= 0.f ? v : a * v; }
CG_DEV float sigmoidf(float v) { return 1.f / (1.f + expf(-v)); }

CG_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CG_DEV float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Loads TM consecutive floats (TM in {1,2,4,8}); p must be aligned to min(TM,4) floats.
template <int TM>
CG_DEV void load_vec(const float* __restrict__ p, float (&w)[TM]) {
  if constexpr (TM % 4 == 0) {
#pragma unroll
    for (int i = 0; i < TM / 4; ++i) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  } else if constexpr (TM == 2) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    w[0] = v.x; w[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < TM; ++i) w[i] = __ldg(p + i);
  }
}

template <int N> struct IntC { static constexpr int value = N; };

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that bypass the register file ---------
CG_DEV void cp_async16(float* smem_dst, const float* __restrict__ gsrc) {
#ifdef CISTGCN_EMU
  smem_dst[0] = gsrc[0]; smem_dst[1] = gsrc[1]; smem_dst[2] = gsrc[2]; smem_dst[3] = gsrc[3];
#else
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
#endif
}
// 4-byte variant (any 4-byte aligned addresses): gathers of strided rows
CG_DEV void cp_async4(float* smem_dst, const float* __restrict__ gsrc) {
#ifdef CISTGCN_EMU
  smem_dst[0] = gsrc[0];
#else
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
#endif
}
CG_DEV void cp_async_commit() {
#ifndef CISTGCN_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
CG_DEV void cp_async_wait_all() {
#ifndef CISTGCN_EMU
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}
// Block-cooperative async copy of n floats (n % 4 == 0, both pointers 16-byte aligned).
template <int NT>
CG_DEV void copy_async(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x * 4; i < n; i += NT * 4) cp_async16(dst + i, src + i);
}

// Vector loads from shared memory (n floats, n in {1,2,4,8,16}); p aligned to min(n,4) floats.
template <int N_>
CG_DEV void lds_vec(const float* p, float (&w)[N_]) {
  if constexpr (N_ % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N_ / 4; ++i) {
      const float4 v = *(reinterpret_cast<const float4*>(p) + i);
      w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  } else if constexpr (N_ == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    w[0] = v.x; w[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < N_; ++i) w[i] = p[i];
  }
}

// L2 prefetch of `bytes` bytes at p, one 128-byte line per thread and round (NT threads cooperate); a hint, no dependency.
template <int NT>
CG_DEV void prefetch_l2_range(const void* p, size_t bytes) {
#ifndef CISTGCN_EMU
  const char* c = reinterpret_cast<const char*>(p);
  for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)NT * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
#else
  (void)p; (void)bytes;
#endif
}

// Warp-level variant: the 32 lanes cover `bytes` bytes at p with one 128-byte line each per round.
CG_DEV void prefetch_l2_warp(const void* p, int bytes) {
#ifndef CISTGCN_EMU
  const char* c = reinterpret_cast<const char*>(p);
  for (int off = (threadIdx.x & 31) * 128; off < bytes; off += 32 * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
#else
  (void)p; (void)bytes;
#endif
}

// ---- TF32 helpers ------------------------------------------------------------------------------------------------
// Nearest TF32 (ties away from zero), returned as an fp32 value with 13 zero low bits.  Integer rounding on the bit
// pattern: ptxas expands cvt.rna.tf32.f32 on sm_100a into the same add / mask plus an Inf / NaN guard (5 instructions);
// the tiles hold finite activations and weights, so the guard is dropped.
CG_DEV float tf32_rna(float x) { return bits_f32((f32_bits(x) + 0x1000u) & 0xFFFFE000u); }
// x = hi + lo: hi = rna_tf32(x); lo = x - hi is exact in fp32 and is handed to the tensor core as it is -- the MMA reads
// only its upper 19 bits, i.e. truncates it to TF32 (error < 2^-10 |lo| <= 2^-21 |x|, the size of the dropped lo*lo term).
CG_DEV void tf32_split(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = x - hi;
}
// c += a (16x8, row) * b (8x8, col); fragments as in PTX mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32
CG_DEV void mma_tf32(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
#ifdef CISTGCN_EMU
  simt_emu::mma_m16n8k8(c, a, b);
#else
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
#endif
}

}  // namespace cg
