// Explicit instantiation: tensor-core (3xTF32 mma.sync) mix stage of the three-stage DSTD-GC path, (T, V) = (10, 22).
#include "dstd_mix_mma.cuh"
namespace cg {
int launch_mix_mma_10_22(const MixArgs& a, void* stream) { return launch_mix_mma_impl<10, 22>(a, stream); }
}  // namespace cg
