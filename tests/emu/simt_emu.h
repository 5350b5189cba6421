// SIMT emulator for kernel-logic tests on machines without a GPU.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the unmodified kernel sources of cistgcn_b200/csrc with g++ (-DCISTGCN_EMU) and runs one
// CTA at a time with one OS thread per CUDA thread: __syncthreads() is a std::barrier, warp shuffles
// go through a per-warp exchange buffer.  It exists so that `pytest -m "not gpu"` can check the
// kernels' indexing and algebra against the oracle; it is never built into, loaded by, or reachable
// from the cistgcn_b200 package (the product loads only the nvcc-built library and fails without it).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <thread>
#include <vector>

struct uint3_emu { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __shared__ static

namespace simt_emu {
struct WarpCtx {
  std::barrier<> bar{32};
  uint64_t slot[32];
  float mma_a[32][4], mma_b[32][2];      // fragment exchange of the emulated mma.sync
};
struct BlockCtx {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<WarpCtx>> warps;
  void* smem = nullptr;
};
inline thread_local BlockCtx* tl_block = nullptr;
inline thread_local WarpCtx* tl_warp = nullptr;
inline thread_local int tl_lane = 0;
inline void* dyn_smem() { return tl_block->smem; }

// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 with the PTX fragment layouts: lane = 4 g + q holds
//   a0 = A[g][q], a1 = A[g+8][q], a2 = A[g][q+4], a3 = A[g+8][q+4];  b0 = B[q][g], b1 = B[q+4][g];
//   c0 = C[g][2q], c1 = C[g][2q+1], c2 = C[g+8][2q], c3 = C[g+8][2q+1].
// Operands are truncated to TF32 (the tensor core ignores the 13 low mantissa bits); products and sums in fp32.
inline float tf32_trunc(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  std::memcpy(&x, &u, 4);
  return x;
}
inline void mma_m16n8k8(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  WarpCtx* w = tl_warp;
  const int lane = tl_lane, g = lane >> 2, q = lane & 3;
  for (int i = 0; i < 4; ++i) w->mma_a[lane][i] = tf32_trunc(a[i]);
  for (int i = 0; i < 2; ++i) w->mma_b[lane][i] = tf32_trunc(b[i]);
  w->bar.arrive_and_wait();
  for (int h = 0; h < 2; ++h)
    for (int j = 0; j < 2; ++j) {
      const int n = 2 * q + j;
      float s = c[2 * h + j];
      for (int k = 0; k < 8; ++k) {
        const float av = w->mma_a[g * 4 + (k & 3)][h + ((k >> 2) << 1)];
        const float bv = w->mma_b[n * 4 + (k & 3)][k >> 2];
        s += av * bv;
      }
      c[2 * h + j] = s;
    }
  w->bar.arrive_and_wait();
}
}  // namespace simt_emu

inline thread_local uint3_emu threadIdx{0, 0, 0};
inline thread_local uint3_emu blockIdx{0, 0, 0};
inline thread_local uint3_emu blockDim{1, 1, 1};
inline thread_local uint3_emu gridDim{1, 1, 1};

static inline void __syncthreads() { simt_emu::tl_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt_emu::tl_warp->bar.arrive_and_wait(); }

template <class T>
static inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  auto* w = simt_emu::tl_warp;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w->slot[simt_emu::tl_lane] = raw;
  w->bar.arrive_and_wait();
  uint64_t got = w->slot[src_lane & 31];
  w->bar.arrive_and_wait();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_shfl(v, lane); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, simt_emu::tl_lane ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  int s = simt_emu::tl_lane + d;
  return emu_shfl(v, s < 32 ? s : simt_emu::tl_lane);
}
template <class T> static inline T __ldg(const T* p) { return *p; }

static inline double atomicAdd(double* p, double v) {
  auto* a = reinterpret_cast<std::atomic<uint64_t>*>(p);
  uint64_t old = a->load();
  for (;;) {
    double cur;
    std::memcpy(&cur, &old, 8);
    double nv = cur + v;
    uint64_t nraw;
    std::memcpy(&nraw, &nv, 8);
    if (a->compare_exchange_weak(old, nraw)) return cur;
  }
}
static inline float atomicAdd(float* p, float v) {
  auto* a = reinterpret_cast<std::atomic<uint32_t>*>(p);
  uint32_t old = a->load();
  for (;;) {
    float cur;
    std::memcpy(&cur, &old, 4);
    float nv = cur + v;
    uint32_t nraw;
    std::memcpy(&nraw, &nv, 4);
    if (a->compare_exchange_weak(old, nraw)) return cur;
  }
}

namespace simt_emu {
// A persistent team of `nt` OS threads per CTA size: a launch hands them the kernel body and they run the blocks one
// after another (all threads of the team inside the same block, so barriers and shuffles behave).  Spawning threads per
// launch made the layer-by-layer training path (thousands of launches) take minutes.
struct Team {
  unsigned nt;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  uint64_t generation = 0;
  unsigned finished = 0;
  const std::function<void()>* body = nullptr;
  dim3 grid;
  size_t smem_bytes = 0;
  BlockCtx ctx;

  explicit Team(unsigned n) : nt(n) {
    ctx.bar = std::make_unique<std::barrier<>>(nt);
    for (unsigned w = 0; w < nt / 32; ++w) ctx.warps.push_back(std::make_unique<WarpCtx>());
    for (unsigned t = 0; t < nt; ++t) threads.emplace_back([this, t]() { worker(t); });
    for (auto& th : threads) th.detach();                     // the team lives until the process exits
  }

  void worker(unsigned t) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lock(mu);
        cv_start.wait(lock, [&] { return generation != seen; });
        seen = generation;
      }
      tl_block = &ctx;
      tl_warp = ctx.warps[t / 32].get();
      tl_lane = int(t % 32);
      threadIdx = {t, 0, 0};
      blockDim = {nt, 1, 1};
      gridDim = {grid.x, 1, 1};
      for (unsigned b = 0; b < grid.x; ++b) {
        if (t == 0) std::memset(ctx.smem, 0xCB, smem_bytes);  // poison: uninitialised shared memory shows up as NaN-ish junk
        ctx.bar->arrive_and_wait();
        blockIdx = {b, 0, 0};
        (*body)();
        ctx.bar->arrive_and_wait();                           // the block is done before its shared memory is reused
      }
      {
        std::lock_guard<std::mutex> lock(mu);
        if (++finished == nt) cv_done.notify_one();
      }
    }
  }

  void run(dim3 g, size_t smem, const std::function<void()>& fn) {
    ctx.smem = std::aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
    {
      std::lock_guard<std::mutex> lock(mu);
      body = &fn; grid = g; smem_bytes = smem; finished = 0;
      ++generation;
    }
    cv_start.notify_all();
    {
      std::unique_lock<std::mutex> lock(mu);
      cv_done.wait(lock, [&] { return finished == nt; });
    }
    std::free(ctx.smem);
    ctx.smem = nullptr;
  }
};

// Runs `body` for every (block, thread); blocks sequentially, threads of a block concurrently.  One launch at a time.
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const unsigned nt = block.x;
  if (nt % 32 != 0 || grid.y != 1 || grid.z != 1) std::abort();
  static std::mutex launch_mu;
  static std::map<unsigned, Team*> teams;                     // leaked on purpose (detached worker threads)
  std::lock_guard<std::mutex> lock(launch_mu);
  Team*& team = teams[nt];
  if (!team) team = new Team(nt);
  team->run(grid, smem_bytes, body);
}
}  // namespace simt_emu
