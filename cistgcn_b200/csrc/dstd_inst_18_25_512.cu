// Explicit instantiation of the fused DSTD-GC kernel for (T, V) = (18, 25), 512 threads per CTA.
#include "dstd_launch.h"
namespace cg {
int launch_dstd_18_25_512(const DstdArgs& a, void* stream) { return launch_dstd_impl<18, 25, 512>(a, stream); }
}  // namespace cg
