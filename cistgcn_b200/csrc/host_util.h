// Host-side launch helpers shared by every translation unit of the library.
#pragma once
#include <mutex>

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

// cudaFuncSetAttribute + the occupancy query cost ~10 us per launch; both only depend on (kernel, threads,
// shared-memory size).  The opt-in shared-memory limit is a process-wide, sticky property of the function, so
// it is only ever RAISED (under a lock); the occupancy answers are remembered per host thread.
struct PreparedKernels {            // process-wide table, keyed by the kernel's address (one per template instance)
  struct Limit { const void* fn; int dev; size_t raised_to; };
  std::mutex mu;
  Limit limits[64];
  int n_limits = 0;
  static PreparedKernels& get() { static PreparedKernels t; return t; }
};

template <class K> inline int prepared_blocks_per_sm(K kfn, int nt, size_t smem_bytes, int* err) {
  struct Entry { const void* fn; int nt; size_t smem; int dev; int blocks; };
  thread_local Entry cache[32];
  thread_local int used = 0;
  const void* key = reinterpret_cast<const void*>(kfn);
  int dev = 0;
#ifndef CISTGCN_EMU
  cudaGetDevice(&dev);
#endif
  *err = 0;
  {
    PreparedKernels& t = PreparedKernels::get();
    std::lock_guard<std::mutex> lock(t.mu);
    PreparedKernels::Limit* lim = nullptr;
    for (int i = 0; i < t.n_limits; ++i)
      if (t.limits[i].fn == key && t.limits[i].dev == dev) lim = &t.limits[i];
    if (!lim && t.n_limits < 64) { lim = &t.limits[t.n_limits++]; *lim = {key, dev, 0}; }
    if (!lim || smem_bytes > lim->raised_to) {
      *err = prepare_kernel(kfn, smem_bytes);
      if (*err) return 0;
      if (lim) lim->raised_to = smem_bytes;
    }
  }
  for (int i = 0; i < used; ++i)
    if (cache[i].fn == key && cache[i].nt == nt && cache[i].smem == smem_bytes && cache[i].dev == dev) return cache[i].blocks;
  const int blocks = blocks_per_sm(kfn, nt, smem_bytes);
  if (used < 32) cache[used++] = Entry{key, nt, smem_bytes, dev, blocks};
  return blocks;
}

inline int cached_sm_count() {
  thread_local int dev_cached = -1, n_cached = 0;
  int dev = 0;
#ifndef CISTGCN_EMU
  cudaGetDevice(&dev);
#endif
  if (dev != dev_cached) { n_cached = sm_count(); dev_cached = dev; }
  return n_cached;
}

inline int grid_for(long long batch, int per_sm) {
  const long long g = (long long)cached_sm_count() * per_sm;
  return (int)(batch < g ? batch : g);
}


// Host: launch of a warp-per-sample kernel (`nwarps` samples per CTA pass).  Returns 0 or a CUDA error code.
template <class K, class ARGS>
inline int launch_warp_per_sample(K kfn, const ARGS& a, void* stream) {
  const int nt = 32 * a.nwarps;
  const size_t smem = (size_t)a.smem_floats * sizeof(float);
  int err = 0;
  const int per_sm = prepared_blocks_per_sm(kfn, nt, smem, &err);
  if (err) return err;
  const long long ctas = ((long long)a.batch + a.nwarps - 1) / a.nwarps;
  const int grid = grid_for(ctas, per_sm);
  CG_LAUNCH(kfn, grid, nt, smem, stream, a);
  return last_launch_error();
}

}  // namespace cg
