// GPU-side data pipeline for training batches (reference: loaders/h36m_motion_3d.py:94-108 __getitem__ -- window split,
// velocity / speed targets -- and environment/custom_transforms.py:10-419 -- the random flip, rotation, scale, noise and
// translation augmentations composed in loaders/loader.py:42-130 in exactly this order).  The reference does this per
// sample on the host with numpy / scipy inside a DataLoader with num_workers = 0 (train_h36m.yaml:86); here one warp owns
// one window, which lives in shared memory while the five transforms and the split are applied.
//
// The random draws stay on the host (cistgcn_b200/data.py draws them like the reference classes do: one uniform per
// "does it fire" decision and one per parameter); the kernel receives them as a per-window parameter record, so the
// arithmetic is deterministic and testable against the reference classes with a patched RNG.
#include "../../include/cistgcn_b200.h"
#include "host_util.h"
#include "simt.h"

namespace cgt { int fail_train(int code, const char* fmt, ...); }

namespace cga {

constexpr int NT = 128, NW = NT / 32;
constexpr int MAX_WINDOW_FLOATS = 64 * 36 * 3;      // frames x joints x 3 of one window (H36M: 35 x 32 x 3)

struct Args {
  const float* windows;     // (N, S, V, 3) resident dataset
  const long long* index;   // [B] window of every batch element
  const float* params;      // (B, CISTGCN_AUG_PARAMS)
  const float* noise;       // optional (B, V, 3) uniform(-1, 1) draws of RandomNoise
  float* sample;            // (B, Tin, V, 3)
  float* target;            // (B, S - Tin, V, 3)
  float* sample_vel;        // optional (B, Tin, V, 3)
  float* target_vel;        // optional (B, S - Tin, V, 3)
  float* target_gvel;       // optional (B, S - Tin, V, 1)
  long long B;
  int S, V, Tin;
};

// mean / min / max over (frames, joints) per coordinate: warp-cooperative, result in every lane
CG_DEV void coord_stats(const float* w, int n_pos, float (&mean)[3], float (&rng)[3]) {
  const int lane = threadIdx.x & 31;
  float s[3] = {0.f, 0.f, 0.f}, mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int p = lane; p < n_pos; p += 32) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { const float v = w[p * 3 + k]; s[k] += v; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s[k] = cg::warp_sum(s[k]);
    mx[k] = cg::warp_max(mx[k]);
    mn[k] = -cg::warp_max(-mn[k]);
    mean[k] = s[k] / n_pos;
    rng[k] = mx[k] - mn[k];
  }
}

__global__ void __launch_bounds__(NT) augment_kernel(const Args a) {
  CG_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = a.S, V = a.V, Tin = a.Tin, To = S - Tin, n_pos = S * V, n = n_pos * 3;
  float* w = smem + warp * (((n + 3) & ~3) + 4);
  for (long long b = (long long)blockIdx.x * NW + warp; b < a.B; b += (long long)gridDim.x * NW) {
    const float* src = a.windows + a.index[b] * n;
    const float* P = a.params + b * CISTGCN_AUG_PARAMS;
    for (int i = lane; i < n; i += 32) w[i] = __ldg(src + i);
    __syncwarp();
    float mean[3], rng[3];
    // ---- RandomFlip (custom_transforms.py:233-279): x_k <- centroid_k - (x_k - centroid_k), centroid of the incoming window
    if (P[CISTGCN_AUG_FLIP + 0] != 0.f || P[CISTGCN_AUG_FLIP + 1] != 0.f || P[CISTGCN_AUG_FLIP + 2] != 0.f) {
      coord_stats(w, n_pos, mean, rng);
      for (int p = lane; p < n_pos; p += 32) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (P[CISTGCN_AUG_FLIP + k] != 0.f) w[p * 3 + k] = mean[k] - (w[p * 3 + k] - mean[k]);
      }
      __syncwarp();
    }
    // ---- RandomRotation (:10-82): (x - centroid) R + centroid, R row-major (data @ R)
    if (P[CISTGCN_AUG_ROT_ON] != 0.f) {
      coord_stats(w, n_pos, mean, rng);
      float R[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = P[CISTGCN_AUG_ROT + i];
      for (int p = lane; p < n_pos; p += 32) {
        const float x0 = w[p * 3] - mean[0], x1 = w[p * 3 + 1] - mean[1], x2 = w[p * 3 + 2] - mean[2];
#pragma unroll
        for (int k = 0; k < 3; ++k) w[p * 3 + k] = (x0 * R[0 * 3 + k] + x1 * R[1 * 3 + k] + x2 * R[2 * 3 + k]) + mean[k];
      }
      __syncwarp();
    }
    // ---- RandomScale (:85-154): x_k <- x_k * s_k
    if (P[CISTGCN_AUG_SCALE_ON] != 0.f) {
      for (int i = lane; i < n; i += 32) w[i] *= P[CISTGCN_AUG_SCALE + i % 3];
      __syncwarp();
    }
    // ---- RandomNoise (:356-402): x <- x + amp * u[joint, k] * (max - min)_k, the same offset in every frame
    if (P[CISTGCN_AUG_NOISE] != 0.f && a.noise) {
      coord_stats(w, n_pos, mean, rng);
      const float amp = P[CISTGCN_AUG_NOISE];
      const float* u = a.noise + b * V * 3;
      for (int i = lane; i < n; i += 32) { const int jk = i % (V * 3); w[i] += amp * __ldg(u + jk) * rng[jk % 3]; }
      __syncwarp();
    }
    // ---- RandomTranslation (:157-230): x_k <- x_k + t_k * (max - min)_k
    if (P[CISTGCN_AUG_TRANS_ON] != 0.f) {
      coord_stats(w, n_pos, mean, rng);
      for (int i = lane; i < n; i += 32) w[i] += P[CISTGCN_AUG_TRANS + i % 3] * rng[i % 3];
      __syncwarp();
    }
    // ---- split + velocity targets (loaders/h36m_motion_3d.py:94-108)
    const int VJ = V * 3;
    float* so = a.sample + b * Tin * VJ;
    float* to = a.target + b * To * VJ;
    for (int i = lane; i < Tin * VJ; i += 32) so[i] = w[i];
    for (int i = lane; i < To * VJ; i += 32) to[i] = w[Tin * VJ + i];
    if (a.sample_vel) {                                     // velocities[:input_n]
      float* o = a.sample_vel + b * Tin * VJ;
      for (int i = lane; i < Tin * VJ; i += 32) o[i] = w[i + VJ] - w[i];
    }
    if (a.target_vel) {                                     // velocities[input_n - 1:].cumsum(0) = x[t + 1] - x[input_n - 1]
      float* o = a.target_vel + b * To * VJ;
      for (int jk = lane; jk < VJ; jk += 32) {
        float s = 0.f;
        for (int t = 0; t < To; ++t) { s += w[(Tin + t) * VJ + jk] - w[(Tin - 1 + t) * VJ + jk]; o[t * VJ + jk] = s; }
      }
    }
    if (a.target_gvel) {                                    // ||velocity||_2 per joint, cumulated over the frames
      float* o = a.target_gvel + b * To * V;
      for (int j = lane; j < V; j += 32) {
        float s = 0.f;
        for (int t = 0; t < To; ++t) {
          float q = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) { const float d = w[(Tin + t) * VJ + j * 3 + k] - w[(Tin - 1 + t) * VJ + j * 3 + k]; q += d * d; }
          s += sqrtf(q);
          o[t * V + j] = s;
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace cga

extern "C" int cistgcn_augment_windows_f32(const float* windows, const int64_t* index, const float* params, const float* noise,
                                           float* sample, float* target, float* sample_vel, float* target_vel, float* target_gvel,
                                           int64_t batch, int32_t seq_len, int32_t joints, int32_t input_n, void* stream) {
  if (!windows || !index || !params || !sample || !target) return cgt::fail_train(-1, "augment_windows: NULL buffer");
  if (batch < 0 || seq_len < 2 || joints < 1 || input_n < 1 || input_n >= seq_len || seq_len * joints * 3 > cga::MAX_WINDOW_FLOATS)
    return cgt::fail_train(-1, "augment_windows: bad geometry (seq %d, joints %d, input_n %d)", seq_len, joints, input_n);
  if (batch == 0) return 0;
  cga::Args a;
  a.windows = windows; a.index = reinterpret_cast<const long long*>(index); a.params = params; a.noise = noise; a.sample = sample;
  a.target = target; a.sample_vel = sample_vel; a.target_vel = target_vel; a.target_gvel = target_gvel; a.B = batch;
  a.S = seq_len; a.V = joints; a.Tin = input_n;
  const size_t smem = (size_t)cga::NW * (((seq_len * joints * 3 + 3) & ~3) + 4) * sizeof(float);
  auto kfn = cga::augment_kernel;
  int err = 0;
  const int per_sm = cg::prepared_blocks_per_sm(kfn, cga::NT, smem, &err);
  if (err) return cgt::fail_train(-3, "augment_windows: cudaFuncSetAttribute(%zu B): %s", smem, cg::launch_error_string(err));
  const long long ctas = (batch + cga::NW - 1) / cga::NW;
  CG_LAUNCH(kfn, cg::grid_for(ctas, per_sm), cga::NT, smem, stream, a);
  if (int e = cg::last_launch_error()) return cgt::fail_train(-4, "augment_kernel launch: %s", cg::launch_error_string(e));
  return 0;
}
