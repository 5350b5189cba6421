// Differentiable building blocks of the CIST-GCN training step and input-gradient path (declared in
// include/cistgcn_b200_train.h).  One hand-written kernel per layer type and direction; fp32 NCHW; no library calls.
// Reference semantics: torch.nn layers as used by models/CISTGCN/CISTGCN.py, environment/train.py:54-107.
//
// Parallelisation is deliberately plain (grid-stride elementwise kernels, one CTA per output channel / weight for the
// reductions): train mode needs whole-batch BatchNorm statistics between any two layers, so the step is a long chain of
// small launches (batch 128 in the reference, train_h36m.yaml:91) and is launch-bound, not bandwidth- or FLOP-bound.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/cistgcn_b200_train.h"
#include "host_util.h"
#include "simt.h"

namespace cgt {

int fail_train(int code, const char* fmt, ...);     // records the message for cistgcn_last_error() (cistgcn_api.cu)

#ifdef CISTGCN_EMU
constexpr int NT = 32;      // SIMT emulator (tests/emu): one OS thread per CUDA thread, keep the CTAs small
#else
constexpr int NT = 256;
#endif

inline int grid_1d(long long n, int per_thread = 1) {
  long long g = (n + (long long)NT * per_thread - 1) / ((long long)NT * per_thread);
#ifdef CISTGCN_EMU
  const long long cap = 2;                                   // the SIMT emulator spawns one OS thread per CUDA thread
#else
  const long long cap = (long long)cg::cached_sm_count() * 16;
#endif
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// CTAs of the one-CTA-per-item kernels (every kernel loops, so any grid size is correct)
inline int grid_items(long long items, int per_sm) {
#ifdef CISTGCN_EMU
  const long long cap = 2;
#else
  const long long cap = (long long)cg::cached_sm_count() * per_sm;
#endif
  const long long g = items < cap ? items : cap;
  return (int)(g < 1 ? 1 : g);
}

CG_DEV double block_sum_f64(double v, double* sh) {   // same in fp64 (heavily cancelling sums: PReLU slope gradients)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < NT / 32; ++i) s += sh[i];
  return s;
}

CG_DEV float block_sum(float v, float* sh) {         // sum over the CTA (NT threads); result in every thread
  v = cg::warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < NT / 32; ++i) s += sh[i];
  return s;
}

// ------------------------------------------------------------------------------------------------ conv2d
struct ConvP { long long B; int Ci, H, W, Co, kh, kw, ph, pw, dh, dw, Ho, Wo; };

// thread = (sample, group of CF_T output channels, output position): one x load per (ci, a, c) feeds CF_T FFMAs; the
// weights of a warp's channel group are warp-uniform (L1 broadcast)
constexpr int CF_T = 4;
__global__ void conv_fwd_kernel(ConvP p, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ y) {
  const int HoWo = p.Ho * p.Wo, cog = (p.Co + CF_T - 1) / CF_T, K = p.Ci * p.kh * p.kw;
  const long long n = p.B * cog * HoWo;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int pos = (int)(i % HoWo), g = (int)((i / HoWo) % cog);
    const long long b = i / ((long long)HoWo * cog);
    const int wo = pos % p.Wo, ho = pos / p.Wo, co0 = g * CF_T;
    float acc[CF_T];
#pragma unroll
    for (int j = 0; j < CF_T; ++j) acc[j] = (bias && co0 + j < p.Co) ? bias[co0 + j] : 0.f;
    const float* xb = x + b * p.Ci * p.H * p.W;
    const float* wp = w + (long long)co0 * K;
    for (int ci = 0; ci < p.Ci; ++ci)
      for (int a = 0; a < p.kh; ++a) {
        const int hi = ho - p.ph + a * p.dh;
        if (hi < 0 || hi >= p.H) continue;
        for (int c = 0; c < p.kw; ++c) {
          const int wi = wo - p.pw + c * p.dw;
          if (wi < 0 || wi >= p.W) continue;
          const float xv = xb[((long long)ci * p.H + hi) * p.W + wi];
          const int k = (ci * p.kh + a) * p.kw + c;
#pragma unroll
          for (int j = 0; j < CF_T; ++j) acc[j] = fmaf(co0 + j < p.Co ? wp[(long long)j * K + k] : 0.f, xv, acc[j]);
        }
      }
#pragma unroll
    for (int j = 0; j < CF_T; ++j)
      if (co0 + j < p.Co) y[(b * p.Co + co0 + j) * HoWo + pos] = acc[j];
  }
}

// thread = (sample, group of CF_T input channels, input position)
__global__ void conv_bwd_input_kernel(ConvP p, const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx) {
  const int HW = p.H * p.W, cig = (p.Ci + CF_T - 1) / CF_T, K = p.Ci * p.kh * p.kw, khw = p.kh * p.kw;
  const long long n = p.B * cig * HW;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int pos = (int)(i % HW), g = (int)((i / HW) % cig);
    const long long b = i / ((long long)HW * cig);
    const int wi = pos % p.W, hi = pos / p.W, ci0 = g * CF_T;
    float acc[CF_T];
#pragma unroll
    for (int j = 0; j < CF_T; ++j) acc[j] = 0.f;
    const float* dyb = dy + b * p.Co * p.Ho * p.Wo;
    for (int a = 0; a < p.kh; ++a) {
      const int ho = hi + p.ph - a * p.dh;
      if (ho < 0 || ho >= p.Ho) continue;
      for (int c = 0; c < p.kw; ++c) {
        const int wo = wi + p.pw - c * p.dw;
        if (wo < 0 || wo >= p.Wo) continue;
        const float* wp = w + ((long long)ci0 * p.kh + a) * p.kw + c;
        const float* dp = dyb + (long long)ho * p.Wo + wo;
        for (int co = 0; co < p.Co; ++co) {
          const float dv = dp[(long long)co * p.Ho * p.Wo];
#pragma unroll
          for (int j = 0; j < CF_T; ++j) acc[j] = fmaf(ci0 + j < p.Ci ? wp[(long long)co * K + j * khw] : 0.f, dv, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CF_T; ++j)
      if (ci0 + j < p.Ci) dx[(b * p.Ci + ci0 + j) * HW + pos] = acc[j];
  }
}

// dW / dbias.  A CTA owns a TCO x TK tile of the weight matrix W[co][k], k = (ci, a, c) flattened, for one chunk of the
// reduction range (b, ho, wo); threads run over the reduction index (coalesced dy / x rows) and keep the whole tile in
// registers: TCO + TK loads feed TCO * TK FFMAs (the first version re-streamed two full columns per weight element).
// Chunk partials go to part[item][chunk]; conv_bw_finish adds them in a fixed order (bit-reproducible, no atomics).
constexpr int CONV_BW_CHUNKS = 32;        // upper bound; small reductions use fewer
constexpr int BW_TCO = 8, BW_TK = 8;
__global__ void __launch_bounds__(NT) conv_bwd_weight_tile_kernel(ConvP p, const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ part, int nchunks, int with_bias) {
  __shared__ float sh[NT / 32][BW_TCO * BW_TK + BW_TCO];
  const int K = p.Ci * p.kh * p.kw;
  const int tiles_k = (K + BW_TK - 1) / BW_TK, tiles_co = (p.Co + BW_TCO - 1) / BW_TCO;
  const long long nw = (long long)p.Co * K;
  const long long red = p.B * p.Ho * p.Wo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int job = blockIdx.x; job < tiles_co * tiles_k * nchunks; job += gridDim.x) {
    const int chunk = job % nchunks, tile = job / nchunks, tk = tile % tiles_k, tco = tile / tiles_k;
    const long long r0 = red * chunk / nchunks, r1 = red * (chunk + 1) / nchunks;
    int xoff[BW_TK], oh[BW_TK], ow[BW_TK];
#pragma unroll
    for (int kk = 0; kk < BW_TK; ++kk) {
      const int k = tk * BW_TK + kk;
      if (k < K) {
        const int c = k % p.kw, a = (k / p.kw) % p.kh, ci = k / (p.kw * p.kh);
        oh[kk] = a * p.dh - p.ph; ow[kk] = c * p.dw - p.pw; xoff[kk] = ci * p.H * p.W;
      } else { oh[kk] = -(1 << 20); ow[kk] = 0; xoff[kk] = 0; }            // never inside the map: contributes 0
    }
    float acc[BW_TCO][BW_TK], bacc[BW_TCO];
#pragma unroll
    for (int j = 0; j < BW_TCO; ++j) { bacc[j] = 0.f;
#pragma unroll
      for (int kk = 0; kk < BW_TK; ++kk) acc[j][kk] = 0.f; }
    const int HoWo = p.Ho * p.Wo;
    for (long long r = r0 + threadIdx.x; r < r1; r += NT) {
      const int wo = (int)(r % p.Wo), ho = (int)((r / p.Wo) % p.Ho);
      const long long b = r / HoWo;
      const float* dyp = dy + (b * p.Co + tco * BW_TCO) * HoWo + ho * p.Wo + wo;
      const float* xb = x + b * p.Ci * p.H * p.W;
      float dv[BW_TCO], xv[BW_TK];
#pragma unroll
      for (int j = 0; j < BW_TCO; ++j) dv[j] = (tco * BW_TCO + j < p.Co) ? dyp[(long long)j * HoWo] : 0.f;
#pragma unroll
      for (int kk = 0; kk < BW_TK; ++kk) {
        const int hi = ho + oh[kk], wi = wo + ow[kk];
        xv[kk] = (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) ? xb[xoff[kk] + hi * p.W + wi] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < BW_TCO; ++j) {
        bacc[j] += dv[j];
#pragma unroll
        for (int kk = 0; kk < BW_TK; ++kk) acc[j][kk] = fmaf(dv[j], xv[kk], acc[j][kk]);
      }
    }
    // CTA reduction of the tile: warp shuffles, then the warps' rows in a fixed order
#pragma unroll
    for (int j = 0; j < BW_TCO; ++j) {
#pragma unroll
      for (int kk = 0; kk < BW_TK; ++kk) {
        const float v = cg::warp_sum(acc[j][kk]);
        if (lane == 0) sh[warp][j * BW_TK + kk] = v;
      }
      const float bv = cg::warp_sum(bacc[j]);
      if (lane == 0) sh[warp][BW_TCO * BW_TK + j] = bv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BW_TCO * BW_TK + BW_TCO; i += NT) {
      float v = 0.f;
      for (int w = 0; w < NT / 32; ++w) v += sh[w][i];
      if (i < BW_TCO * BW_TK) {
        const int co = tco * BW_TCO + i / BW_TK, k = tk * BW_TK + i % BW_TK;
        if (co < p.Co && k < K) part[((long long)co * K + k) * nchunks + chunk] = v;
      } else if (with_bias && tk == 0) {
        const int co = tco * BW_TCO + (i - BW_TCO * BW_TK);
        if (co < p.Co) part[(nw + co) * nchunks + chunk] = v;
      }
    }
    __syncthreads();
  }
}
__global__ void conv_bw_finish_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ dbias, long long nw,
                                      int n_items, int nchunks) {
  for (long long e = (long long)blockIdx.x * NT + threadIdx.x; e < n_items; e += (long long)gridDim.x * NT) {
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += part[e * nchunks + c];
    if (e < nw) dw[e] = s; else dbias[e - nw] = s;
  }
}

// ------------------------------------------------------------------------------------------------ batch norm
// Statistics over (B, HW) per channel are cut into BN_CHUNKS chunks so that C * BN_CHUNKS CTAs work at once (C alone is
// 3 .. 64 CTAs on 148 SMs).  scratch layout: DOUBLE part[C][BN_CHUNKS][2] then float coef[C][2].  Every reduction of
// the BatchNorm pair runs in fp64, like ATen's CPU kernels (acc_type<float> = double there): the sums behind dbeta /
// dgamma / mean(dy) cancel heavily on the 1 .. 4 channel maps of Map2Adj, and in fp32 their error fed 10 .. 40 x the
// reference's own noise into the gradients upstream of them.
#ifdef CISTGCN_EMU
constexpr int BN_CHUNKS = 2;              // SIMT emulator: every chunk costs host-thread barriers
#else
constexpr int BN_CHUNKS = 16;
#endif
// chunk statistics, two passes inside the chunk: (mean_k, M2_k = sum (x - mean_k)^2)
__global__ void bn_stats_kernel(const float* __restrict__ x, double* __restrict__ part, long long B, int C, int HW) {
  __shared__ double sh[NT / 32];
  const long long n = B * HW;
  for (int job = blockIdx.x; job < C * BN_CHUNKS; job += gridDim.x) {
    const int c = job / BN_CHUNKS, k = job % BN_CHUNKS;
    const long long r0 = n * k / BN_CHUNKS, r1 = n * (k + 1) / BN_CHUNKS;
    double s = 0.0;
    for (long long r = r0 + threadIdx.x; r < r1; r += NT) s += (double)x[((r / HW) * C + c) * HW + r % HW];
    const double cnt = (double)(r1 - r0);
    const double mean = cnt > 0 ? block_sum_f64(s, sh) / cnt : 0.0;
    double q = 0.0;
    for (long long r = r0 + threadIdx.x; r < r1; r += NT) { const double d = (double)x[((r / HW) * C + c) * HW + r % HW] - mean; q = fma(d, d, q); }
    q = block_sum_f64(q, sh);
    if (threadIdx.x == 0) { part[(c * BN_CHUNKS + k) * 2] = mean; part[(c * BN_CHUNKS + k) * 2 + 1] = q; }
    __syncthreads();
  }
}
// merge the chunks (Chan's update, fixed order), update the running statistics, emit save_mean / save_invstd
__global__ void bn_finalize_kernel(const double* __restrict__ part, float* running_mean, float* running_var, float* save_mean,
                                   float* save_invstd, long long B, int C, int HW, int training, float momentum, float eps) {
  const long long n = B * HW;
  for (int c = blockIdx.x * NT + threadIdx.x; c < C; c += gridDim.x * NT) {
    float mean, invstd;
    if (training) {
      double na = 0.0, m = 0.0, M2 = 0.0;
      for (int k = 0; k < BN_CHUNKS; ++k) {
        const double nb = (double)(n * (k + 1) / BN_CHUNKS - n * k / BN_CHUNKS);
        if (nb <= 0.0) continue;
        const double mb = part[(c * BN_CHUNKS + k) * 2], qb = part[(c * BN_CHUNKS + k) * 2 + 1];
        const double delta = mb - m;
        m += delta * (nb / (na + nb));
        M2 += qb + delta * delta * (na * nb / (na + nb));
        na += nb;
      }
      mean = (float)m;
      const double var = M2 / (double)n;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(n > 1 ? M2 / (double)(n - 1) : var);
    } else {
      mean = running_mean[c];
      invstd = 1.f / sqrtf(running_var[c] + eps);
    }
    save_mean[c] = mean; save_invstd[c] = invstd;
  }
}
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ y,
                                long long n, int C, int HW) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int c = (int)((i / HW) % C);
    const float g = gamma[c] * save_invstd[c];
    y[i] = fmaf(x[i], g, beta[c] - save_mean[c] * g);
  }
}
// backward: chunk sums of dy and dy * xhat -> part[C][BN_CHUNKS][2]
__global__ void bn_bwd_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ save_mean,
                                      const float* __restrict__ save_invstd, double* __restrict__ part, long long B, int C, int HW) {
  __shared__ double sh[NT / 32];
  const long long n = B * HW;
  for (int job = blockIdx.x; job < C * BN_CHUNKS; job += gridDim.x) {
    const int c = job / BN_CHUNKS, k = job % BN_CHUNKS;
    const long long r0 = n * k / BN_CHUNKS, r1 = n * (k + 1) / BN_CHUNKS;
    const float mean = save_mean[c], invstd = save_invstd[c];
    double sdy = 0.0, sdyx = 0.0;
    for (long long r = r0 + threadIdx.x; r < r1; r += NT) {
      const long long i = ((r / HW) * C + c) * HW + r % HW;
      const double d = (double)dy[i];
      sdy += d;
      sdyx = fma(d, (double)((x[i] - mean) * invstd), sdyx);
    }
    sdy = block_sum_f64(sdy, sh);
    sdyx = block_sum_f64(sdyx, sh);
    if (threadIdx.x == 0) { part[(c * BN_CHUNKS + k) * 2] = sdy; part[(c * BN_CHUNKS + k) * 2 + 1] = sdyx; }
    __syncthreads();
  }
}
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ part, float* __restrict__ coef, float* dgamma, float* dbeta, long long B,
                                       int C, int HW) {
  const double n = (double)(B * HW);
  for (int c = blockIdx.x * NT + threadIdx.x; c < C; c += gridDim.x * NT) {
    double sdy = 0.0, sdyx = 0.0;
    for (int k = 0; k < BN_CHUNKS; ++k) { sdy += part[(c * BN_CHUNKS + k) * 2]; sdyx += part[(c * BN_CHUNKS + k) * 2 + 1]; }
    if (dgamma) { dgamma[c] = (float)sdyx; dbeta[c] = (float)sdy; }
    coef[2 * c] = (float)(sdy / n); coef[2 * c + 1] = (float)(sdyx / n);
  }
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                    const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                    const float* __restrict__ coef, float* __restrict__ dx, long long n, int C, int HW, int training) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int c = (int)((i / HW) % C);
    const float invstd = save_invstd[c], k = gamma[c] * invstd;
    dx[i] = training ? k * (dy[i] - coef[2 * c] - (x[i] - save_mean[c]) * invstd * coef[2 * c + 1]) : k * dy[i];
  }
}

// ------------------------------------------------------------------------------------------------ PReLU / activations / dropout
__global__ void prelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ slope, float* __restrict__ y, long long n,
                                 int C, int HW, int ns) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float a = slope[ns == 1 ? 0 : (int)((i / HW) % C)];
    const float v = x[i];
    y[i] = v >= 0.f ? v : a * v;
  }
}

// dx elementwise; dslope: one CTA per slope, reduction over its elements
__global__ void prelu_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ slope,
                                    float* __restrict__ dx, long long n, int C, int HW, int ns) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float a = slope[ns == 1 ? 0 : (int)((i / HW) % C)];
    dx[i] = x[i] >= 0.f ? dy[i] : a * dy[i];
  }
}
// partial sums per CTA -> part[slope][gridDim.y]; a second tiny kernel adds them in a fixed order (bit-reproducible)
__global__ void prelu_bwd_ds_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, double* __restrict__ part,
                                            long long B, int C, int HW, int ns) {
  __shared__ double sh[NT / 32];
  const int nchunk = CISTGCN_PRELU_SCRATCH_PER_SLOPE;
  for (int item = blockIdx.x; item < ns * nchunk; item += gridDim.x) {
  const int s = item / nchunk, chunk = item % nchunk;
  // the slope gradient is a sum of ~10^5 .. 10^6 products of both signs whose total is orders of magnitude below its
  // terms (the reference's own fp32 result is only good to ~3e-4 on some of them): the whole reduction runs in fp64
  double accd = 0.0;
  if (ns == 1) {
    const long long n = B * C * HW;
    for (long long i = (long long)chunk * NT + threadIdx.x; i < n; i += (long long)nchunk * NT)
      if (x[i] < 0.f) accd += (double)dy[i] * (double)x[i];
  } else {
    const long long n = B * HW;
    for (long long r = (long long)chunk * NT + threadIdx.x; r < n; r += (long long)nchunk * NT) {
      const long long i = ((r / HW) * C + s) * HW + r % HW;
      if (x[i] < 0.f) accd += (double)dy[i] * (double)x[i];
    }
  }
  const double acc = block_sum_f64(accd, sh);
  if (threadIdx.x == 0) part[(long long)s * nchunk + chunk] = acc;
  }
}
__global__ void sum_rows_kernel(const double* __restrict__ part, float* __restrict__ out, int rows, int cols) {
  for (int r = blockIdx.x * NT + threadIdx.x; r < rows; r += gridDim.x * NT) {
    double s = 0.0;
    for (int c = 0; c < cols; ++c) s += part[(long long)r * cols + c];
    out[r] = (float)s;
  }
}

__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int kind) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float v = x[i];
    y[i] = kind == 0 ? fmaxf(v, 0.f) : 1.f / (1.f + expf(-v));
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long n, int kind) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float v = y[i];
    dx[i] = kind == 0 ? (v > 0.f ? dy[i] : 0.f) : dy[i] * v * (1.f - v);
  }
}

CG_DEV unsigned long long mix64(unsigned long long z) {       // splitmix64 finaliser: counter-based, stateless
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// `step` (optional, device memory): a per-step counter mixed into the seed, so that a captured CUDA graph of the training
// step draws fresh masks on every replay without a host-side argument change
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p, unsigned long long seed,
                               const unsigned long long* __restrict__ step) {
  const float keep_scale = 1.f / (1.f - p);
  if (step) seed ^= mix64(*step * 0xD1342543DE82EF95ull + 1ull);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const unsigned long long h = mix64(seed ^ mix64((unsigned long long)i));
    const float u = (float)(h >> 40) * (1.f / 16777216.f);
    y[i] = u >= p ? x[i] * keep_scale : 0.f;
  }
}

__global__ void counter_bump_kernel(unsigned long long* c) { if (threadIdx.x == 0 && blockIdx.x == 0) *c += 1ull; }

// ------------------------------------------------------------------------------------------------ strided copy / flat axpby
struct Copy4 { long long ds[4], ss[4], sz[4]; };
__global__ void copy4d_kernel(Copy4 c, float* __restrict__ dst, const float* __restrict__ src, int accumulate) {
  const long long n = c.sz[0] * c.sz[1] * c.sz[2] * c.sz[3];
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long i3 = i % c.sz[3], i2 = (i / c.sz[3]) % c.sz[2], i1 = (i / (c.sz[3] * c.sz[2])) % c.sz[1];
    const long long i0 = i / (c.sz[3] * c.sz[2] * c.sz[1]);
    const float v = src[i0 * c.ss[0] + i1 * c.ss[1] + i2 * c.ss[2] + i3 * c.ss[3]];
    float* d = dst + i0 * c.ds[0] + i1 * c.ds[1] + i2 * c.ds[2] + i3 * c.ds[3];
    *d = accumulate ? *d + v : v;
  }
}
__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT)
    y[i] = b == 0.f ? a * x[i] : fmaf(a, x[i], b * y[i]);
}

// ------------------------------------------------------------------------------------------------ adjacency products
// domain 0: y[n,c,q,v] = sum_t x[n,c,t,v] A[n,v,t,q];  domain 1: y[n,c,t,w] = sum_v x[n,c,t,v] A[n,t,v,w]
__global__ void gcn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ A, float* __restrict__ y, long long B, int C,
                               int T, int V, int domain, int ab) {
  const long long n = B * C * T * V;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int v = (int)(i % V), t = (int)((i / V) % T);
    const long long nc = i / ((long long)V * T), b = nc / C;
    const float* xp = x + nc * T * V;
    float acc = 0.f;
    if (domain == 0) {            // output index (q = t, v)
      const float* ap = A + (ab ? b * (long long)V * T * T : 0) + (long long)v * T * T + t;
      for (int k = 0; k < T; ++k) acc = fmaf(xp[k * V + v], ap[k * T], acc);
    } else {                      // output index (t, w = v)
      const float* ap = A + (ab ? b * (long long)T * V * V : 0) + (long long)t * V * V + v;
      for (int k = 0; k < V; ++k) acc = fmaf(xp[t * V + k], ap[k * V], acc);
    }
    y[i] = acc;
  }
}
// dx: domain 0: dx[n,c,t,v] = sum_q dy[n,c,q,v] A[n,v,t,q];  domain 1: dx[n,c,t,v] = sum_w dy[n,c,t,w] A[n,t,v,w]
__global__ void gcn_bwd_x_kernel(const float* __restrict__ A, const float* __restrict__ dy, float* __restrict__ dx, long long B, int C,
                                 int T, int V, int domain, int ab) {
  const long long n = B * C * T * V;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int v = (int)(i % V), t = (int)((i / V) % T);
    const long long nc = i / ((long long)V * T), b = nc / C;
    const float* dp = dy + nc * T * V;
    float acc = 0.f;
    if (domain == 0) {
      const float* ap = A + (ab ? b * (long long)V * T * T : 0) + ((long long)v * T + t) * T;
      for (int q = 0; q < T; ++q) acc = fmaf(dp[q * V + v], ap[q], acc);
    } else {
      const float* ap = A + (ab ? b * (long long)T * V * V : 0) + ((long long)t * V + v) * V;
      for (int w = 0; w < V; ++w) acc = fmaf(dp[t * V + w], ap[w], acc);
    }
    dx[i] = acc;
  }
}
// dA (per sample): domain 0: dA[n,v,t,q] = sum_c x[n,c,t,v] dy[n,c,q,v];  domain 1: dA[n,t,v,w] = sum_c x[n,c,t,v] dy[n,c,t,w]
__global__ void gcn_bwd_a_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dA, long long B, int C,
                                 int T, int V, int domain) {
  const long long per = domain == 0 ? (long long)V * T * T : (long long)T * V * V;
  const long long n = B * per;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long b = i / per, r = i % per;
    const float* xp = x + b * C * T * V;
    const float* dp = dy + b * C * T * V;
    float acc = 0.f;
    if (domain == 0) {
      const int q = (int)(r % T), t = (int)((r / T) % T), v = (int)(r / ((long long)T * T));
      for (int c = 0; c < C; ++c) acc = fmaf(xp[(c * T + t) * V + v], dp[(c * T + q) * V + v], acc);
    } else {
      const int w = (int)(r % V), v = (int)((r / V) % V), t = (int)(r / ((long long)V * V));
      for (int c = 0; c < C; ++c) acc = fmaf(xp[(c * T + t) * V + v], dp[(c * T + t) * V + w], acc);
    }
    dA[i] = acc;
  }
}
// static A: one CTA per element of A, reduction over (b, c)
__global__ void gcn_bwd_a_static_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dA, long long B,
                                        int C, int T, int V, int domain) {
  __shared__ float sh[NT / 32];
  const long long per = domain == 0 ? (long long)V * T * T : (long long)T * V * V;
  for (long long r = blockIdx.x; r < per; r += gridDim.x) {
    int i1, i2;         // flattened (t, v) offsets of the x and dy factors
    if (domain == 0) {
      const int q = (int)(r % T), t = (int)((r / T) % T), v = (int)(r / ((long long)T * T));
      i1 = t * V + v; i2 = q * V + v;
    } else {
      const int w = (int)(r % V), v = (int)((r / V) % V), t = (int)(r / ((long long)V * V));
      i1 = t * V + v; i2 = t * V + w;
    }
    float acc = 0.f;
    for (long long bc = threadIdx.x; bc < B * C; bc += NT) acc = fmaf(x[bc * T * V + i1], dy[bc * T * V + i2], acc);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) dA[r] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ Map2Adj outer products
// dsp (B,V,T), dseq (B,T,V).  domain 0: o[n,v,t,q] = dsp[n,v,t] dseq[n,q,v];  domain 1: o[n,t,v,w] = dsp[n,v,t] dseq[n,t,w]
__global__ void outer_fwd_kernel(const float* __restrict__ dsp, const float* __restrict__ dseq, float* __restrict__ o, long long B,
                                 int T, int V, int domain) {
  const long long per = domain == 0 ? (long long)V * T * T : (long long)T * V * V;
  const long long n = B * per;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long b = i / per, r = i % per;
    const float* sp = dsp + b * V * T;
    const float* sq = dseq + b * T * V;
    if (domain == 0) {
      const int q = (int)(r % T), t = (int)((r / T) % T), v = (int)(r / ((long long)T * T));
      o[i] = sp[v * T + t] * sq[q * V + v];
    } else {
      const int w = (int)(r % V), v = (int)((r / V) % V), t = (int)(r / ((long long)V * V));
      o[i] = sp[v * T + t] * sq[t * V + w];
    }
  }
}
// thread per element of d_dsp (first B*V*T) and d_dseq (next B*T*V)
__global__ void outer_bwd_kernel(const float* __restrict__ dsp, const float* __restrict__ dseq, const float* __restrict__ d_o,
                                 float* __restrict__ d_dsp, float* __restrict__ d_dseq, long long B, int T, int V, int domain) {
  const long long half = B * V * T;
  const long long per = domain == 0 ? (long long)V * T * T : (long long)T * V * V;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < 2 * half; i += (long long)gridDim.x * NT) {
    float acc = 0.f;
    if (i < half) {                                   // d_dsp[n,v,t]
      const long long b = i / ((long long)V * T);
      const int t = (int)(i % T), v = (int)((i / T) % V);
      const float* go = d_o + b * per;
      const float* sq = dseq + b * T * V;
      if (domain == 0) { for (int q = 0; q < T; ++q) acc = fmaf(go[((long long)v * T + t) * T + q], sq[q * V + v], acc); }
      else { for (int w = 0; w < V; ++w) acc = fmaf(go[((long long)t * V + v) * V + w], sq[t * V + w], acc); }
      d_dsp[i] = acc;
    } else {                                          // d_dseq[n,q,v] (domain 0) / d_dseq[n,t,w] (domain 1)
      const long long j = i - half, b = j / ((long long)T * V);
      const int c2 = (int)(j % V), c1 = (int)((j / V) % T);
      const float* go = d_o + b * per;
      const float* sp = dsp + b * V * T;
      if (domain == 0) { for (int t = 0; t < T; ++t) acc = fmaf(go[((long long)c2 * T + t) * T + c1], sp[c2 * T + t], acc); }
      else { for (int v = 0; v < V; ++v) acc = fmaf(go[((long long)c1 * V + v) * V + c2], sp[v * T + c1], acc); }
      d_dseq[j] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ _get_stats_
// one CTA per sample; shared: m1[C], s1[C], m2[C*T], s2[C*T]
__global__ void stats_fwd_kernel(const float* __restrict__ x, float* __restrict__ stats, long long B, int C, int T, int V) {
  CG_DYN_SMEM(shm);
  float* m1 = shm; float* s1 = m1 + C; float* m2 = s1 + C; float* s2 = m2 + C * T;
  const int TV = T * V, NS = 2 + 2 * T;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const float* xb = x + b * C * TV;
    for (int r = threadIdx.x; r < C * T; r += NT) {
      const float* xp = xb + (long long)r * V;
      float s = 0.f;
      for (int v = 0; v < V; ++v) s += xp[v];
      const float mu = s / V;
      float q = 0.f;
      for (int v = 0; v < V; ++v) { const float d = xp[v] - mu; q = fmaf(d, d, q); }
      m2[r] = mu; s2[r] = sqrtf(q / (V - 1));
    }
    for (int c = threadIdx.x; c < C; c += NT) {
      const float* xp = xb + (long long)c * TV;
      float s = 0.f;
      for (int i = 0; i < TV; ++i) s += xp[i];
      const float mu = s / TV;
      float q = 0.f;
      for (int i = 0; i < TV; ++i) { const float d = xp[i] - mu; q = fmaf(d, d, q); }
      m1[c] = mu; s1[c] = sqrtf(q / (TV - 1));
    }
    __syncthreads();
    float* out = stats + b * NS;
    for (int j = threadIdx.x; j < NS; j += NT) {
      // j = 0: mean_c m1; 1..T: mean_c m2[., t]; T+1: std_c s1; T+2..: std_c s2[., t]
      const bool is_std = j > T;
      const int t = is_std ? j - T - 2 : j - 1;               // -1: channel-level entry
      const float* src = is_std ? (t < 0 ? s1 : s2) : (t < 0 ? m1 : m2);
      const int stride = t < 0 ? 1 : T, off = t < 0 ? 0 : t;
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += src[c * stride + off];
      const float mu = s / C;
      if (!is_std) { out[j] = mu; continue; }
      float q = 0.f;
      for (int c = 0; c < C; ++c) { const float d = src[c * stride + off] - mu; q = fmaf(d, d, q); }
      out[j] = sqrtf(q / (C - 1));
    }
    __syncthreads();
  }
}

// dx += d(stats)/dx ^T dstats.  shared: m1, s1, m2, s2 (recomputed), ds1[C], ds2[C*T]
__global__ void stats_bwd_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ dstats,
                                 float* __restrict__ dx, long long B, int C, int T, int V) {
  CG_DYN_SMEM(shm);
  float* m1 = shm; float* s1 = m1 + C; float* m2 = s1 + C; float* s2 = m2 + C * T; float* ds1 = s2 + C * T; float* ds2 = ds1 + C;
  const int TV = T * V, NS = 2 + 2 * T;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const float* xb = x + b * C * TV;
    const float* st = stats + b * NS;
    const float* g = dstats + b * NS;
    for (int r = threadIdx.x; r < C * T; r += NT) {
      const float* xp = xb + (long long)r * V;
      float s = 0.f;
      for (int v = 0; v < V; ++v) s += xp[v];
      const float mu = s / V;
      float q = 0.f;
      for (int v = 0; v < V; ++v) { const float d = xp[v] - mu; q = fmaf(d, d, q); }
      m2[r] = mu; s2[r] = sqrtf(q / (V - 1));
    }
    for (int c = threadIdx.x; c < C; c += NT) {
      const float* xp = xb + (long long)c * TV;
      float s = 0.f;
      for (int i = 0; i < TV; ++i) s += xp[i];
      const float mu = s / TV;
      float q = 0.f;
      for (int i = 0; i < TV; ++i) { const float d = xp[i] - mu; q = fmaf(d, d, q); }
      m1[c] = mu; s1[c] = sqrtf(q / (TV - 1));
    }
    __syncthreads();
    // gradients w.r.t. the per-channel stds: out = std_c(s): ds[c] = g * (s[c] - mean_c s) / ((C-1) * out)
    for (int c = threadIdx.x; c < C; c += NT) {
      float s = 0.f;
      for (int k = 0; k < C; ++k) s += s1[k];
      ds1[c] = g[T + 1] * (s1[c] - s / C) / ((C - 1) * st[T + 1]);
    }
    for (int r = threadIdx.x; r < C * T; r += NT) {
      const int t = r % T;
      float s = 0.f;
      for (int k = 0; k < C; ++k) s += s2[k * T + t];
      ds2[r] = g[T + 2 + t] * (s2[r] - s / C) / ((C - 1) * st[T + 2 + t]);
    }
    __syncthreads();
    float* dxb = dx + b * C * TV;
    for (int i = threadIdx.x; i < C * TV; i += NT) {
      const int c = i / TV, t = (i % TV) / V, r = c * T + t;
      const float xv = xb[i];
      float d = g[0] / (float)(C * TV) + g[1 + t] / (float)(C * V);
      d += ds1[c] * (xv - m1[c]) / ((TV - 1) * s1[c]);
      d += ds2[r] * (xv - m2[r]) / ((V - 1) * s2[r]);
      dxb[i] += d;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ squeeze / scale
// one warp per (b, c) row
__global__ void spatial_mean_fwd_kernel(const float* __restrict__ x, float* __restrict__ m, long long rows, int HW) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (NT / 32)) {
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s += x[r * HW + i];
    s = cg::warp_sum(s);
    if (lane == 0) m[r] = s / HW;
  }
}
__global__ void spatial_mean_bwd_kernel(const float* __restrict__ dm, float* __restrict__ dx, long long n, int HW, int accumulate) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float v = dm[i / HW] / HW;
    dx[i] = accumulate ? dx[i] + v : v;
  }
}
__global__ void scale_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y, long long n, int HW) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) y[i] = x[i] * g[i / HW];
}
// one warp per (b, c) row: dx = dy * g, dg = sum_hw dy * x
__global__ void scale_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ dy,
                                 float* __restrict__ dx, float* __restrict__ dg, long long rows, int HW) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (NT / 32)) {
    const float gv = g[r];
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) {
      const float d = dy[r * HW + i];
      dx[r * HW + i] = d * gv;
      s = fmaf(d, x[r * HW + i], s);
    }
    s = cg::warp_sum(s);
    if (lane == 0) dg[r] = s;
  }
}

// ------------------------------------------------------------------------------------------------ row max / cumsum
__global__ void rowmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ idx, long long R, int N) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * (NT / 32)) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < N; i += 32) { const float v = x[r * N + i]; if (v > best) { best = v; bi = i; } }
    for (int o = 16; o > 0; o >>= 1) {        // ties resolve to the smallest index, like torch.max
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { y[r] = best; idx[r] = bi; }
  }
}
__global__ void rowmax_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ idx, float* __restrict__ dx, long long R, int N) {
  for (long long r = (long long)blockIdx.x * NT + threadIdx.x; r < R; r += (long long)gridDim.x * NT) dx[r * N + idx[r]] = dy[r];
}
__global__ void cumsum_kernel(const float* __restrict__ x, float* __restrict__ y, long long B, int L, int N, int reverse) {
  const long long n = B * N;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long b = i / N, j = i % N;
    float s = 0.f;
    if (!reverse) for (int l = 0; l < L; ++l) { s += x[(b * L + l) * N + j]; y[(b * L + l) * N + j] = s; }
    else for (int l = L - 1; l >= 0; --l) { s += x[(b * L + l) * N + j]; y[(b * L + l) * N + j] = s; }
  }
}

// ------------------------------------------------------------------------------------------------ features (CISTGCN.py:568-577)
// thread per (b, v, k): x (B,T,V,3) -> f (B,10,T,V) channels [x(3), acc(3), vel(3), |vel|]
__global__ void features_fwd_kernel(const float* __restrict__ x, float* __restrict__ f, long long B, int T, int V) {
  const long long n = B * T * V;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int v = (int)(i % V), t = (int)((i / V) % T);
    const long long b = i / ((long long)V * T);
    const float* xb = x + b * T * V * 3;
    float* fb = f + b * 10 * T * V + (long long)t * V + v;
    float sp = 0.f;
    for (int k = 0; k < 3; ++k) {
      const float p0 = xb[(t * V + v) * 3 + k];
      float vel, acc;
      if (t < T - 1) {
        const float p1 = xb[((t + 1) * V + v) * 3 + k];
        vel = p1 - p0;
        const float veln = (t < T - 2) ? xb[((t + 2) * V + v) * 3 + k] - p1 : p1;
        acc = veln - vel;
      } else { vel = p0; acc = p0; }
      fb[(long long)k * T * V] = p0; fb[(long long)(3 + k) * T * V] = acc; fb[(long long)(6 + k) * T * V] = vel;
      sp = fmaf(vel, vel, sp);
    }
    fb[(long long)9 * T * V] = sqrtf(sp);
  }
}
// thread per input element (b, t, v, k): gathers the adjoint of the linear maps vel = D x (+ last row), acc = D vel (+ last row)
// and of speed = ||vel||.  gv[s] = d/d vel[s] = df_vel[s] + df_speed[s] * vel[s]/speed[s] + (D^T df_acc)[s].
__global__ void features_bwd_kernel(const float* __restrict__ x, const float* __restrict__ df, float* __restrict__ dx, long long B,
                                    int T, int V) {
  const long long n = B * T * V * 3;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int k = (int)(i % 3), v = (int)((i / 3) % V), t = (int)((i / (3LL * V)) % T);
    const long long b = i / (3LL * V * T);
    const float* xb = x + b * T * V * 3;
    const float* fb = df + b * 10 * T * V;
    const long long TVl = (long long)T * V;
    auto velk = [&](int s, int kk) { return s < T - 1 ? xb[((s + 1) * V + v) * 3 + kk] - xb[(s * V + v) * 3 + kk] : xb[(s * V + v) * 3 + kk]; };
    auto gvel = [&](int s) {          // gradient w.r.t. vel[s][v][k]
      float g = fb[(6 + k) * TVl + s * V + v];
      const float v0 = velk(s, 0), v1 = velk(s, 1), v2 = velk(s, 2);
      const float spd = sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
      if (spd > 0.f) g += fb[9 * TVl + s * V + v] * velk(s, k) / spd;
      // acc[s'] = vel[s'+1] - vel[s'] (s' < T-1), acc[T-1] = vel[T-1]
      if (s < T - 1) g -= fb[(3 + k) * TVl + s * V + v];
      if (s >= 1) g += fb[(3 + k) * TVl + (s - 1) * V + v];
      if (s == T - 1) g += fb[(3 + k) * TVl + s * V + v];
      return g;
    };
    // vel[s] = x[s+1] - x[s] (s < T-1), vel[T-1] = x[T-1]
    float g = fb[k * TVl + t * V + v];
    if (t < T - 1) g -= gvel(t);
    if (t >= 1) g += gvel(t - 1);
    if (t == T - 1) g += gvel(t);
    dx[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------ mpjpe backward / Adam
__global__ void mpjpe_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ dpred,
                                 long long n, float scale) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float d0 = pred[3 * i] - target[3 * i], d1 = pred[3 * i + 1] - target[3 * i + 1], d2 = pred[3 * i + 2] - target[3 * i + 2];
    const float nr = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const float k = nr > 0.f ? scale / nr : 0.f;           // torch: subgradient 0 at the origin
    dpred[3 * i] = k * d0; dpred[3 * i + 1] = k * d1; dpred[3 * i + 2] = k * d2;
  }
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float gs) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, gs * g[i]);
    const float mv = b1 * m[i] + (1.f - b1) * gv;
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv; v[i] = vv;
    p[i] = pv - (lr / bc1) * mv / (sqrtf(vv) / sqrtf(bc2) + eps);
  }
}

int launched(const char* what) {
  if (int e = cg::last_launch_error()) return fail_train(-4, "%s launch: %s", what, cg::launch_error_string(e));
  return 0;
}

}  // namespace cgt

using namespace cgt;

extern "C" {

static bool conv_params(const cistgcn_conv_shape* s, ConvP& p) {
  if (!s || s->B < 0 || s->Ci < 1 || s->Co < 1 || s->H < 1 || s->W < 1 || s->kh < 1 || s->kw < 1 || s->dh < 1 || s->dw < 1 ||
      s->ph < 0 || s->pw < 0)
    return false;
  p = {s->B, s->Ci, s->H, s->W, s->Co, s->kh, s->kw, s->ph, s->pw, s->dh, s->dw,
       s->H + 2 * s->ph - s->dh * (s->kh - 1), s->W + 2 * s->pw - s->dw * (s->kw - 1)};
  return p.Ho >= 1 && p.Wo >= 1;
}

int cistgcn_conv2d_fwd(const cistgcn_conv_shape* s, const float* x, const float* w, const float* bias, float* y, void* stream) {
  ConvP p;
  if (!conv_params(s, p) || !x || !w || !y) return fail_train(-1, "conv2d_fwd: bad arguments");
  if (p.B == 0) return 0;
  CG_LAUNCH(conv_fwd_kernel, grid_1d(p.B * ((p.Co + CF_T - 1) / CF_T) * p.Ho * p.Wo), NT, 0, stream, p, x, w, bias, y);
  return launched("conv_fwd_kernel");
}
int cistgcn_conv2d_bwd_input(const cistgcn_conv_shape* s, const float* dy, const float* w, float* dx, void* stream) {
  ConvP p;
  if (!conv_params(s, p) || !dy || !w || !dx) return fail_train(-1, "conv2d_bwd_input: bad arguments");
  if (p.B == 0) return 0;
  CG_LAUNCH(conv_bwd_input_kernel, grid_1d(p.B * ((p.Ci + CF_T - 1) / CF_T) * p.H * p.W), NT, 0, stream, p, dy, w, dx);
  return launched("conv_bwd_input_kernel");
}
size_t cistgcn_conv2d_bwd_weight_scratch_floats(const cistgcn_conv_shape* s) {
  if (!s) return 0;
  return ((size_t)s->Co * s->Ci * s->kh * s->kw + s->Co) * CONV_BW_CHUNKS;
}
int cistgcn_conv2d_bwd_weight(const cistgcn_conv_shape* s, const float* x, const float* dy, float* dw, float* dbias, float* scratch,
                              void* stream) {
  ConvP p;
  if (!conv_params(s, p) || !x || !dy || !dw || !scratch) return fail_train(-1, "conv2d_bwd_weight: bad arguments");
  const int K = p.Ci * p.kh * p.kw;
  const long long nw = (long long)p.Co * K;
  const int items = (int)(nw + (dbias ? p.Co : 0));
  const long long red = p.B * p.Ho * p.Wo;
  int nchunks = (int)(red / (2 * NT));                      // at least two passes of the CTA per chunk
  if (nchunks < 1) nchunks = 1;
  if (nchunks > CONV_BW_CHUNKS) nchunks = CONV_BW_CHUNKS;
#ifdef CISTGCN_EMU
  if (nchunks > 2) nchunks = 2;                              // SIMT emulator: every job costs host-thread barriers
#endif
  const long long jobs = (long long)((p.Co + BW_TCO - 1) / BW_TCO) * ((K + BW_TK - 1) / BW_TK) * nchunks;
  CG_LAUNCH(conv_bwd_weight_tile_kernel, grid_items(jobs, 16), NT, 0, stream, p, x, dy, scratch, nchunks, dbias ? 1 : 0);
  if (int rc = launched("conv_bwd_weight_tile_kernel")) return rc;
  CG_LAUNCH(conv_bw_finish_kernel, grid_1d(items), NT, 0, stream, (const float*)scratch, dw, dbias, nw, items, nchunks);
  return launched("conv_bw_finish_kernel");
}

size_t cistgcn_bn_scratch_floats(int32_t C) { return (size_t)C * (4 * BN_CHUNKS + 2); }   // fp64 partials + fp32 coefficients
int cistgcn_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float* y,
                   float* save_mean, float* save_invstd, float* scratch, int64_t B, int32_t C, int32_t HW, int32_t training,
                   float momentum, float eps, void* stream) {
  if (!x || !gamma || !beta || !running_mean || !running_var || !y || !save_mean || !save_invstd || B < 1 || C < 1 || HW < 1 ||
      (training && !scratch))
    return fail_train(-1, "bn_fwd: bad arguments");
  if (training) {
    CG_LAUNCH(bn_stats_kernel, grid_items((long long)C * BN_CHUNKS, 16), NT, 0, stream, x, reinterpret_cast<double*>(scratch), (long long)B, C, HW);
    if (int rc = launched("bn_stats_kernel")) return rc;
  }
  CG_LAUNCH(bn_finalize_kernel, grid_1d(C), NT, 0, stream, (const double*)reinterpret_cast<double*>(scratch), running_mean, running_var, save_mean, save_invstd,
            (long long)B, C, HW, training, momentum, eps);
  if (int rc = launched("bn_finalize_kernel")) return rc;
  const long long n = (long long)B * C * HW;
  CG_LAUNCH(bn_apply_kernel, grid_1d(n), NT, 0, stream, x, gamma, beta, (const float*)save_mean, (const float*)save_invstd, y, n, C, HW);
  return launched("bn_apply_kernel");
}
int cistgcn_bn_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                   float* dx, float* dgamma, float* dbeta, float* scratch, int64_t B, int32_t C, int32_t HW, int32_t training,
                   void* stream) {
  if (!x || !dy || !gamma || !save_mean || !save_invstd || !dx || !scratch || B < 1 || C < 1 || HW < 1 ||
      ((dgamma == nullptr) != (dbeta == nullptr)))
    return fail_train(-1, "bn_bwd: bad arguments");
  float* coef = scratch + (size_t)C * 4 * BN_CHUNKS;          // behind the fp64 partials
  if (training || dgamma) {
    CG_LAUNCH(bn_bwd_partial_kernel, grid_items((long long)C * BN_CHUNKS, 16), NT, 0, stream, x, dy, save_mean, save_invstd,
              reinterpret_cast<double*>(scratch), (long long)B, C, HW);
    if (int rc = launched("bn_bwd_partial_kernel")) return rc;
    CG_LAUNCH(bn_bwd_finalize_kernel, grid_1d(C), NT, 0, stream, (const double*)reinterpret_cast<double*>(scratch), coef, dgamma, dbeta, (long long)B, C, HW);
    if (int rc = launched("bn_bwd_finalize_kernel")) return rc;
  }
  const long long n = (long long)B * C * HW;
  CG_LAUNCH(bn_bwd_apply_kernel, grid_1d(n), NT, 0, stream, x, dy, gamma, save_mean, save_invstd, (const float*)coef, dx, n, C, HW, training);
  return launched("bn_bwd_apply_kernel");
}

int cistgcn_prelu_fwd(const float* x, const float* slope, float* y, int64_t B, int32_t C, int32_t HW, int32_t n_slopes, void* stream) {
  if (!x || !slope || !y || B < 0 || C < 1 || HW < 1 || (n_slopes != 1 && n_slopes != C)) return fail_train(-1, "prelu_fwd: bad arguments");
  const long long n = (long long)B * C * HW;
  if (n == 0) return 0;
  CG_LAUNCH(prelu_fwd_kernel, grid_1d(n), NT, 0, stream, x, slope, y, n, C, HW, n_slopes);
  return launched("prelu_fwd_kernel");
}
int cistgcn_prelu_bwd(const float* x, const float* dy, const float* slope, float* dx, float* dslope, float* scratch, int64_t B,
                      int32_t C, int32_t HW, int32_t n_slopes, void* stream) {
  if (!x || !dy || !slope || !dx || B < 1 || C < 1 || HW < 1 || (n_slopes != 1 && n_slopes != C) || (dslope && !scratch))
    return fail_train(-1, "prelu_bwd: bad arguments");
  const long long n = (long long)B * C * HW;
  CG_LAUNCH(prelu_bwd_dx_kernel, grid_1d(n), NT, 0, stream, x, dy, slope, dx, n, C, HW, n_slopes);
  if (int rc = launched("prelu_bwd_dx_kernel")) return rc;
  if (dslope) {
    constexpr int NCHUNK = CISTGCN_PRELU_SCRATCH_PER_SLOPE;       // partial sums per slope, added in a fixed order
    double* part = reinterpret_cast<double*>(scratch);        // scratch is 8-byte aligned (caller: 2 floats per partial)
    CG_LAUNCH(prelu_bwd_ds_partial_kernel, grid_items((long long)n_slopes * NCHUNK, 64), NT, 0, stream, x, dy, part, (long long)B, C, HW, n_slopes);
    if (int rc = launched("prelu_bwd_ds_partial_kernel")) return rc;
    CG_LAUNCH(sum_rows_kernel, grid_1d(n_slopes), NT, 0, stream, (const double*)part, dslope, n_slopes, NCHUNK);
    return launched("sum_rows_kernel");
  }
  return 0;
}

int cistgcn_act_fwd(const float* x, float* y, int64_t n, int32_t kind, void* stream) {
  if (!x || !y || n < 0 || kind < 0 || kind > 1) return fail_train(-1, "act_fwd: bad arguments");
  if (n == 0) return 0;
  CG_LAUNCH(act_fwd_kernel, grid_1d(n), NT, 0, stream, x, y, (long long)n, kind);
  return launched("act_fwd_kernel");
}
int cistgcn_act_bwd(const float* y, const float* dy, float* dx, int64_t n, int32_t kind, void* stream) {
  if (!y || !dy || !dx || n < 0 || kind < 0 || kind > 1) return fail_train(-1, "act_bwd: bad arguments");
  if (n == 0) return 0;
  CG_LAUNCH(act_bwd_kernel, grid_1d(n), NT, 0, stream, y, dy, dx, (long long)n, kind);
  return launched("act_bwd_kernel");
}
int cistgcn_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* step, void* stream) {
  if (!x || !y || n < 0 || !(p >= 0.f && p < 1.f)) return fail_train(-1, "dropout: bad arguments");
  if (n == 0) return 0;
  CG_LAUNCH(dropout_kernel, grid_1d(n), NT, 0, stream, x, y, (long long)n, p, (unsigned long long)seed,
            reinterpret_cast<const unsigned long long*>(step));
  return launched("dropout_kernel");
}
int cistgcn_counter_bump(uint64_t* counter, void* stream) {
  if (!counter) return fail_train(-1, "counter_bump: bad arguments");
  CG_LAUNCH(counter_bump_kernel, 1, 32, 0, stream, reinterpret_cast<unsigned long long*>(counter));
  return launched("counter_bump_kernel");
}

int cistgcn_copy4d(float* dst, const int64_t dst_strides[4], const float* src, const int64_t src_strides[4],
                   const int64_t sizes[4], int32_t accumulate, void* stream) {
  if (!dst || !src || !dst_strides || !src_strides || !sizes) return fail_train(-1, "copy4d: NULL argument");
  Copy4 c;
  long long n = 1;
  for (int i = 0; i < 4; ++i) { c.ds[i] = dst_strides[i]; c.ss[i] = src_strides[i]; c.sz[i] = sizes[i]; n *= sizes[i]; if (sizes[i] < 0) return fail_train(-1, "copy4d: negative size"); }
  if (n == 0) return 0;
  CG_LAUNCH(copy4d_kernel, grid_1d(n), NT, 0, stream, c, dst, src, accumulate);
  return launched("copy4d_kernel");
}
int cistgcn_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream) {
  if (!x || !y || n < 0) return fail_train(-1, "axpby: bad arguments");
  if (n == 0) return 0;
  CG_LAUNCH(axpby_kernel, grid_1d(n, 4), NT, 0, stream, a, x, b, y, (long long)n);
  return launched("axpby_kernel");
}

int cistgcn_gcn_fwd(const float* x, const float* A, float* y, int64_t B, int32_t C, int32_t T, int32_t V, int32_t domain,
                    int32_t a_batched, void* stream) {
  if (!x || !A || !y || B < 0 || C < 1 || T < 1 || V < 1 || (domain != 0 && domain != 1)) return fail_train(-1, "gcn_fwd: bad arguments");
  if (B == 0) return 0;
  CG_LAUNCH(gcn_fwd_kernel, grid_1d((long long)B * C * T * V), NT, 0, stream, x, A, y, (long long)B, C, T, V, domain, a_batched);
  return launched("gcn_fwd_kernel");
}
int cistgcn_gcn_bwd(const float* x, const float* A, const float* dy, float* dx, float* dA, int64_t B, int32_t C, int32_t T,
                    int32_t V, int32_t domain, int32_t a_batched, void* stream) {
  if (!x || !A || !dy || B < 1 || C < 1 || T < 1 || V < 1 || (domain != 0 && domain != 1)) return fail_train(-1, "gcn_bwd: bad arguments");
  if (dx) {
    CG_LAUNCH(gcn_bwd_x_kernel, grid_1d((long long)B * C * T * V), NT, 0, stream, A, dy, dx, (long long)B, C, T, V, domain, a_batched);
    if (int rc = launched("gcn_bwd_x_kernel")) return rc;
  }
  if (dA) {
    const long long per = domain == 0 ? (long long)V * T * T : (long long)T * V * V;
    if (a_batched) {
      CG_LAUNCH(gcn_bwd_a_kernel, grid_1d((long long)B * per), NT, 0, stream, x, dy, dA, (long long)B, C, T, V, domain);
      return launched("gcn_bwd_a_kernel");
    }
    CG_LAUNCH(gcn_bwd_a_static_kernel, grid_items(per, 32), NT, 0, stream, x, dy, dA, (long long)B, C, T, V, domain);
    return launched("gcn_bwd_a_static_kernel");
  }
  return 0;
}

int cistgcn_outer_fwd(const float* dsp, const float* dseq, float* o, int64_t B, int32_t T, int32_t V, int32_t domain, void* stream) {
  if (!dsp || !dseq || !o || B < 0 || T < 1 || V < 1 || (domain != 0 && domain != 1)) return fail_train(-1, "outer_fwd: bad arguments");
  if (B == 0) return 0;
  const long long per = domain == 0 ? (long long)V * T * T : (long long)T * V * V;
  CG_LAUNCH(outer_fwd_kernel, grid_1d((long long)B * per), NT, 0, stream, dsp, dseq, o, (long long)B, T, V, domain);
  return launched("outer_fwd_kernel");
}
int cistgcn_outer_bwd(const float* dsp, const float* dseq, const float* d_o, float* d_dsp, float* d_dseq, int64_t B, int32_t T,
                      int32_t V, int32_t domain, void* stream) {
  if (!dsp || !dseq || !d_o || !d_dsp || !d_dseq || B < 1 || T < 1 || V < 1 || (domain != 0 && domain != 1))
    return fail_train(-1, "outer_bwd: bad arguments");
  CG_LAUNCH(outer_bwd_kernel, grid_1d(2LL * B * T * V), NT, 0, stream, dsp, dseq, d_o, d_dsp, d_dseq, (long long)B, T, V, domain);
  return launched("outer_bwd_kernel");
}

int cistgcn_stats_fwd(const float* x, float* stats, int64_t B, int32_t C, int32_t T, int32_t V, void* stream) {
  if (!x || !stats || B < 0 || C < 2 || T < 1 || V < 2) return fail_train(-1, "stats_fwd: bad arguments (needs C >= 2, V >= 2)");
  if (B == 0) return 0;
  const size_t smem = (size_t)(2 * C + 2 * C * T) * sizeof(float);
  CG_LAUNCH(stats_fwd_kernel, grid_items(B, 8), NT, smem, stream, x, stats, (long long)B, C, T, V);
  return launched("stats_fwd_kernel");
}
int cistgcn_stats_bwd(const float* x, const float* stats, const float* dstats, float* dx, int64_t B, int32_t C, int32_t T,
                      int32_t V, void* stream) {
  if (!x || !stats || !dstats || !dx || B < 1 || C < 2 || T < 1 || V < 2) return fail_train(-1, "stats_bwd: bad arguments");
  const size_t smem = (size_t)(3 * C + 3 * C * T) * sizeof(float);
  CG_LAUNCH(stats_bwd_kernel, grid_items(B, 8), NT, smem, stream, x, stats, dstats, dx, (long long)B, C, T, V);
  return launched("stats_bwd_kernel");
}

int cistgcn_spatial_mean_fwd(const float* x, float* m, int64_t B, int32_t C, int32_t HW, void* stream) {
  if (!x || !m || B < 0 || C < 1 || HW < 1) return fail_train(-1, "spatial_mean_fwd: bad arguments");
  if (B == 0) return 0;
  CG_LAUNCH(spatial_mean_fwd_kernel, grid_1d((long long)B * C * 32), NT, 0, stream, x, m, (long long)B * C, HW);
  return launched("spatial_mean_fwd_kernel");
}
int cistgcn_spatial_mean_bwd(const float* dm, float* dx, int64_t B, int32_t C, int32_t HW, int32_t accumulate, void* stream) {
  if (!dm || !dx || B < 1 || C < 1 || HW < 1) return fail_train(-1, "spatial_mean_bwd: bad arguments");
  CG_LAUNCH(spatial_mean_bwd_kernel, grid_1d((long long)B * C * HW), NT, 0, stream, dm, dx, (long long)B * C * HW, HW, accumulate);
  return launched("spatial_mean_bwd_kernel");
}
int cistgcn_scale_fwd(const float* x, const float* g, float* y, int64_t B, int32_t C, int32_t HW, void* stream) {
  if (!x || !g || !y || B < 0 || C < 1 || HW < 1) return fail_train(-1, "scale_fwd: bad arguments");
  if (B == 0) return 0;
  CG_LAUNCH(scale_fwd_kernel, grid_1d((long long)B * C * HW), NT, 0, stream, x, g, y, (long long)B * C * HW, HW);
  return launched("scale_fwd_kernel");
}
int cistgcn_scale_bwd(const float* x, const float* g, const float* dy, float* dx, float* dg, int64_t B, int32_t C, int32_t HW,
                      void* stream) {
  if (!x || !g || !dy || !dx || !dg || B < 1 || C < 1 || HW < 1) return fail_train(-1, "scale_bwd: bad arguments");
  CG_LAUNCH(scale_bwd_kernel, grid_1d((long long)B * C * 32), NT, 0, stream, x, g, dy, dx, dg, (long long)B * C, HW);
  return launched("scale_bwd_kernel");
}

int cistgcn_rowmax_fwd(const float* x, float* y, int32_t* idx, int64_t R, int32_t N, void* stream) {
  if (!x || !y || !idx || R < 0 || N < 1) return fail_train(-1, "rowmax_fwd: bad arguments");
  if (R == 0) return 0;
  CG_LAUNCH(rowmax_fwd_kernel, grid_1d((long long)R * 32), NT, 0, stream, x, y, idx, (long long)R, N);
  return launched("rowmax_fwd_kernel");
}
int cistgcn_rowmax_bwd(const float* dy, const int32_t* idx, float* dx, int64_t R, int32_t N, void* stream) {
  if (!dy || !idx || !dx || R < 1 || N < 1) return fail_train(-1, "rowmax_bwd: bad arguments");
  CG_LAUNCH(rowmax_bwd_kernel, grid_1d(R), NT, 0, stream, dy, idx, dx, (long long)R, N);
  return launched("rowmax_bwd_kernel");
}
int cistgcn_cumsum(const float* x, float* y, int64_t B, int32_t L, int32_t N, int32_t reverse, void* stream) {
  if (!x || !y || B < 0 || L < 1 || N < 1) return fail_train(-1, "cumsum: bad arguments");
  if (B == 0) return 0;
  CG_LAUNCH(cumsum_kernel, grid_1d((long long)B * N), NT, 0, stream, x, y, (long long)B, L, N, reverse);
  return launched("cumsum_kernel");
}

int cistgcn_features_fwd(const float* x, float* f, int64_t B, int32_t T, int32_t V, void* stream) {
  if (!x || !f || B < 0 || T < 2 || V < 1) return fail_train(-1, "features_fwd: bad arguments");
  if (B == 0) return 0;
  CG_LAUNCH(features_fwd_kernel, grid_1d((long long)B * T * V), NT, 0, stream, x, f, (long long)B, T, V);
  return launched("features_fwd_kernel");
}
int cistgcn_features_bwd(const float* x, const float* df, float* dx, int64_t B, int32_t T, int32_t V, void* stream) {
  if (!x || !df || !dx || B < 1 || T < 2 || V < 1) return fail_train(-1, "features_bwd: bad arguments");
  CG_LAUNCH(features_bwd_kernel, grid_1d((long long)B * T * V * 3), NT, 0, stream, x, df, dx, (long long)B, T, V);
  return launched("features_bwd_kernel");
}

int cistgcn_mpjpe_bwd(const float* pred, const float* target, float* dpred, int64_t n_joints, float scale, void* stream) {
  if (!pred || !target || !dpred || n_joints < 0) return fail_train(-1, "mpjpe_bwd: bad arguments");
  if (n_joints == 0) return 0;
  CG_LAUNCH(mpjpe_bwd_kernel, grid_1d(n_joints), NT, 0, stream, pred, target, dpred, (long long)n_joints, scale);
  return launched("mpjpe_bwd_kernel");
}
int cistgcn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int32_t step, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || n < 0 || step < 1) return fail_train(-1, "adam_step: bad arguments");
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  CG_LAUNCH(adam_kernel, grid_1d(n, 2), NT, 0, stream, p, g, m, v, (long long)n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  return launched("adam_kernel");
}

}  // extern "C"
