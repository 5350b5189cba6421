// Host-side launch helpers shared by every translation unit of the library.
#pragma once
#include "simt.h"

namespace cg {

#ifdef CISTGCN_EMU
inline int sm_count() { return 2; }
template <class K> inline int prepare_kernel(K, size_t) { return 0; }
template <class K> inline int blocks_per_sm(K, int, size_t) { return 1; }
inline int last_launch_error() { return 0; }
inline const char* launch_error_string(int) { return "emulator"; }
#else
inline int sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}
template <class K> inline int prepare_kernel(K kfn, size_t smem_bytes) {
  return (int)cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}
template <class K> inline int blocks_per_sm(K kfn, int nt, size_t smem_bytes) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kfn, nt, smem_bytes) != cudaSuccess || n < 1) n = 1;
  return n;
}
inline int last_launch_error() { return (int)cudaGetLastError(); }
inline const char* launch_error_string(int e) { return cudaGetErrorString((cudaError_t)e); }
#endif

inline int grid_for(long long batch, int per_sm) {
  const long long g = (long long)sm_count() * per_sm;
  return (int)(batch < g ? batch : g);
}

}  // namespace cg
