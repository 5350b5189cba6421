// Explicit instantiation of the fused DSTD-GC kernel for (T, V) = (10, 22).
#include "dstd_launch.h"
namespace cg {
int launch_dstd_10_22(const DstdArgs& a, void* stream) { return launch_dstd_impl<10, 22>(a, stream); }
}  // namespace cg
