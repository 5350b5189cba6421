// Explicit instantiations: reduce + adjacency stages (+ the narrow 3-channel mix stage) of the three-stage DSTD-GC path for (T, V) = (18, 25).
#include "dstd_adj.cuh"
#include "dstd_mix_narrow.cuh"
#include "dstd_reduce.cuh"
namespace cg {
int launch_reduce_18_25_1(const ReduceArgs& a, void* stream) { return launch_reduce_impl<18, 25, 1>(a, stream); }
int launch_reduce_18_25_2(const ReduceArgs& a, void* stream) { return launch_reduce_impl<18, 25, 2>(a, stream); }
int launch_adj_18_25(const AdjArgs& a, void* stream) { return launch_adj_impl<18, 25>(a, stream); }
int launch_mix_narrow_18_25(const MixArgs& a, void* stream) { return launch_mix_narrow_impl<18, 25, 3>(a, stream); }
}  // namespace cg
