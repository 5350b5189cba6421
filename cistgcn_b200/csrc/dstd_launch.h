// Launchers of the fused DSTD-GC kernel, one translation unit per (T, V, threads) instantiation so that
// nvcc can compile them in parallel (dstd_inst_*.cu).  Return 0 or a CUDA error code.
//
// Two CTA shapes: 512 threads x 1 CTA per SM (wide blocks: the fp32 working set of E >= 32 fills the SM's
// shared memory) and 256 threads x 2 CTAs per SM (narrow blocks: two samples in flight per SM hide each
// other's barrier and latency stalls).  cistgcn_api.cu picks by what the shared-memory plan allows.
#pragma once
#include "dstd_block.cuh"
#include "host_util.h"

namespace cg {

constexpr int DSTD_NT_WIDE = 512;      // 1 CTA / SM, up to 227 KB of shared memory
constexpr int DSTD_NT_NARROW = 256;    // 2 CTAs / SM, up to 113 KB each
constexpr int DSTD_SMEM_NARROW_BYTES = (233472 - 2 * 1024) / 2;

#define CG_DECL_DSTD(T, V) \
  int launch_dstd_##T##_##V##_256(const DstdArgs& a, void* stream); \
  int launch_dstd_##T##_##V##_512(const DstdArgs& a, void* stream);
CG_DECL_DSTD(10, 22)
CG_DECL_DSTD(10, 18)
CG_DECL_DSTD(22, 25)
CG_DECL_DSTD(18, 25)
#undef CG_DECL_DSTD
int launch_dstd_10_22_512_tc(const DstdArgs& a, void* stream);     // tensor-core channel mixes (DstdArgs::tc)
int launch_dstd_10_18_512_tc(const DstdArgs& a, void* stream);

template <int T, int V, int NT, bool TC = false>
inline int launch_dstd_impl(const DstdArgs& a, void* stream) {
  auto kfn = dstd_block_kernel<T, V, NT, TC>;
  const size_t smem = (size_t)a.smem_floats * sizeof(float);
  int err = 0;
  const int per_sm = prepared_blocks_per_sm(kfn, NT, smem, &err);
  if (err) return err;
  const int grid = grid_for(a.batch, per_sm);
  CG_LAUNCH(kfn, grid, NT, smem, stream, a);
  return last_launch_error();
}

}  // namespace cg
