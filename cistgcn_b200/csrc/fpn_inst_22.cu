// Explicit instantiation of the FPN-chain kernel for V = 22 joints.
#include "fpn_launch.h"
namespace cg {
int launch_fpn_22(const FpnArgs& a, void* stream) { return launch_fpn_impl<22>(a, stream); }
}  // namespace cg
