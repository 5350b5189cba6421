// Launchers of the fused DSTD-GC kernel, one translation unit per (T, V) instantiation so that nvcc
// can compile them in parallel (dstd_inst_*.cu).  Return 0 or a CUDA error code.
#pragma once
#include "dstd_block.cuh"
#include "host_util.h"

namespace cg {

#ifndef CISTGCN_DSTD_NT
#define CISTGCN_DSTD_NT 512
#endif
constexpr int DSTD_NT = CISTGCN_DSTD_NT;

int launch_dstd_10_22(const DstdArgs& a, void* stream);
int launch_dstd_10_18(const DstdArgs& a, void* stream);
int launch_dstd_22_25(const DstdArgs& a, void* stream);
int launch_dstd_18_25(const DstdArgs& a, void* stream);

template <int T, int V>
inline int launch_dstd_impl(const DstdArgs& a, void* stream) {
  auto kfn = dstd_block_kernel<T, V, DSTD_NT>;
  const size_t smem = (size_t)a.smem_floats * sizeof(float);
  int err = 0;
  const int per_sm = prepared_blocks_per_sm(kfn, DSTD_NT, smem, &err);
  if (err) return err;
  const int grid = grid_for(a.batch, per_sm);
  CG_LAUNCH(kfn, grid, DSTD_NT, smem, stream, a);
  return last_launch_error();
}

}  // namespace cg
