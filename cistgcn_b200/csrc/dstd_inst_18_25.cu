// Explicit instantiation of the fused DSTD-GC kernel for (T, V) = (18, 25).
#include "dstd_launch.h"
namespace cg {
int launch_dstd_18_25(const DstdArgs& a, void* stream) { return launch_dstd_impl<18, 25>(a, stream); }
}  // namespace cg
