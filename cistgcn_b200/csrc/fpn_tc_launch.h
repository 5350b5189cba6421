// Launchers of the tensor-core FPN kernel (fpn_tc.cuh), one translation unit per joint count.
#pragma once
#ifndef CISTGCN_EMU
#include "fpn_tc.cuh"
#include "host_util.h"

namespace cg {

int launch_fpn_tc_22(const FpnTcArgs& a, void* stream);
int launch_fpn_tc_18(const FpnTcArgs& a, void* stream);
int launch_fpn_tc_22_bf16(const FpnTcArgs& a, void* stream);      // single-term bf16 operands, bf16 input (cistgcn_forward_bf16)
int launch_fpn_tc_18_bf16(const FpnTcArgs& a, void* stream);

template <int V, bool BF16 = false>
inline int launch_fpn_tc_impl(const FpnTcArgs& a, void* stream) {
  auto kfn = fpn_tc_kernel<V, BF16>;
  const size_t smem = (size_t)FtcGeom<V>::SMEM_BYTES;
  int err = 0;
  prepared_blocks_per_sm(kfn, FTC_NT, smem, &err);       // raises the opt-in shared-memory limit once
  if (err) return err;
  const int grid = grid_for(a.batch, 1);                  // one CTA per SM: it owns all 512 TMEM columns
  CG_LAUNCH(kfn, grid, FTC_NT, smem, stream, a);
  return last_launch_error();
}

}  // namespace cg
#endif
