// Explicit instantiation of the FPN-chain kernel for V = 18 joints.
#include "fpn_launch.h"
namespace cg {
int launch_fpn_18(const FpnArgs& a, void* stream) { return launch_fpn_impl<18>(a, stream); }
}  // namespace cg
