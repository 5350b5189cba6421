// ContextLayer (CISTGCN.py:393-475), output assembly (:595-597) and MPJPE (losses/losses.py:50-61).
//
// One CTA per sample (persistent).  The three (hidden x Tout x 3V) context maps of the reference
// (422 KB per sample each in eager mode) are never materialised: every thread streams its channel's
// values through max / sum accumulators in registers.
#pragma once
#include "../../include/cistgcn_b200.h"
#include "dstd_block.cuh"
#include "simt.h"

namespace cg {

constexpr int TAIL_NT = 256;

struct TailArgs {
  int t[CT_COUNT];
  const float* w;
  const float* x;        // (B, Tin, V, 3)   model input (for x[:, -1:])
  const float* x7;       // (B, Tout, V, 3)
  const float* x8;       // (B, Tout, V, 3)  output DSTD-GC block result, already permuted back
  float* pred;           // (B, Tout, V, 3)
  const float* target;   // optional
  double* frame_sums;    // optional, [Tout]
  float* tap_joints; float* tap_disp; float* tap_sjn; float* tap_sjd;
  int batch;
  int smem_floats;
};

inline void tail_plan(TailArgs& a) {
  const int To = a.t[CT_TOUT], V = a.t[CT_V], H = a.t[CT_HID];
  a.smem_floats = pad4i(To * V * 3) + 3 * pad4i(4 * H) + pad4i(3 * H) + pad4i(3 * To) + pad4i(V) + pad4i(To) +
                  3 * pad4i(To * V) + pad4i(3 * To * V) + 4 * pad4i(To) + 2 * 8 + 2 * pad4i(To) + pad4i(To) + (TAIL_NT / 32) * 32;
}

__global__ void __launch_bounds__(TAIL_NT, 4) tail_kernel(const TailArgs a) {
  CG_DYN_SMEM(smem);
  constexpr int NT = TAIL_NT;
  const int tid = threadIdx.x;
  const float* __restrict__ W = a.w;
  const int* t = a.t;
  const int Tin = t[CT_TIN], To = t[CT_TOUT], V = t[CT_V], H = t[CT_HID], S1 = t[CT_SEH1], S2 = t[CT_SEH2];
  const int V3 = V * 3, NZ = To * V3, Top = pad8i(To);
  float* p = smem;
  float* z = p;      p += pad4i(NZ);          // x7 as (Tout, 3V)
  float* part1 = p;  p += pad4i(4 * H);       // partial max of context_conv1
  float* part2 = p;  p += pad4i(4 * H);       // partial max of context_conv2
  float* part3 = p;  p += pad4i(4 * H);       // partial sum of context_conv3
  float* yv = p;     p += pad4i(3 * H);       // y1 | y2 | ym
  float* y = p;      p += pad4i(3 * To);
  float* joints = p; p += pad4i(V);
  float* disp = p;   p += pad4i(To);
  float* n1 = p;     p += pad4i(To * V);
  float* n1s = p;    p += pad4i(To * V);
  float* n2 = p;     p += pad4i(To * V);
  float* f3 = p;     p += pad4i(3 * To * V);
  float* mean1 = p;  p += pad4i(To);
  float* gate1 = p;  p += pad4i(To);
  float* mean2 = p;  p += pad4i(To);
  float* gate2 = p;  p += pad4i(To);
  float* e1 = p;     p += 8;
  float* e2 = p;     p += 8;
  float* fsum = p;   p += pad4i(To);          // per-frame error partials of this sample
  float* d2 = p;     p += pad4i(To);          // norm_map.0 applied to the displacement vector
  float* zsl = p;    p += (NT / 32) * 32;     // per-warp, per-slice sums of z
  double facc = 0.0;                          // thread tid < To accumulates frame tid over this CTA's samples

  const bool slice_sums = H <= NT && NT % H == 0 && ((NT / H) & (NT / H - 1)) == 0 && NT / H <= 32;
  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    const float* x7 = a.x7 + (size_t)b * NZ;
    for (int i = tid; i < NZ; i += NT) z[i] = x7[i];
    for (int i = tid; i < To; i += NT) fsum[i] = 0.f;
    __syncthreads();
    // ---- context_conv1 (max over all positions), context_conv3 (mean), context_conv2 (max over columns)
    // context_conv1 is a per-channel scale of the ONE input map followed by BN + PReLU: g_c(z) = PReLU(s_c z + b_c) is
    // monotone in z (slope a >= 0) or V-shaped (a < 0), so its maximum over the positions is attained where z is
    // smallest or largest -- exactly, rounding included.  Only the extremes of z are needed, not a pass per channel.
    {
      // (the same pass sums z per float4 slice i % nsub -- a thread's float4s all fall into slice tid % nsub -- for
      // context_conv3 below, where every one of the H channels of a slice would otherwise add up the same values)
      float lo = INFINITY, hi = -INFINITY, szp = 0.f;
      const float4* z4 = reinterpret_cast<const float4*>(z);
      const int n4 = NZ >> 2;
      for (int i = tid; i < n4; i += NT) {
        const float4 q = z4[i];
        lo = fminf(fminf(lo, q.x), fminf(fminf(q.y, q.z), q.w));
        hi = fmaxf(fmaxf(hi, q.x), fmaxf(fmaxf(q.y, q.z), q.w));
        szp += (q.x + q.y) + (q.z + q.w);
      }
      if (tid == 0) for (int i = n4 << 2; i < NZ; ++i) { lo = fminf(lo, z[i]); hi = fmaxf(hi, z[i]); szp += z[i]; }   // slice 0 takes the tail
      lo = -warp_max(-lo);
      hi = warp_max(hi);
      if ((tid & 31) == 0) { yv[tid >> 5] = lo; yv[NT / 32 + (tid >> 5)] = hi; }      // yv is free until the next phase
      if (slice_sums) {
        for (int o = 16; o >= NT / H; o >>= 1) szp += __shfl_xor_sync(0xffffffffu, szp, o);
        if ((tid & 31) < NT / H) zsl[(tid >> 5) * 32 + (tid & 31)] = szp;
      }
    }
    __syncthreads();
    {
      const int ch = tid % H, sub = tid / H, nsub = NT / H;      // H = 64 -> 4 position slices per channel
      float zlo = yv[0], zhi = yv[NT / 32];
#pragma unroll
      for (int w = 1; w < NT / 32; ++w) { zlo = fminf(zlo, yv[w]); zhi = fmaxf(zhi, yv[NT / 32 + w]); }
      const float s1 = W[t[CT_C1_S] + ch], b1 = W[t[CT_C1_B] + ch], a1 = W[t[CT_C1_A]];
      const float s3 = W[t[CT_C3_S] + ch], b3 = W[t[CT_C3_B] + ch], a3 = W[t[CT_C3_A]];
      const float mx = fmaxf(prelu(fmaf(s1, zlo, b1), a1), prelu(fmaf(s1, zhi, b1), a1));
      // context_conv3: sum over the positions of PReLU(u), u = s z + b.  PReLU(u) = a u + (1 - a) max(u, 0), the first
      // part sums in closed form and max(s z + b, 0) = s max(z, th) + b (s > 0) or s min(z, th) + b (s < 0) with
      // th = -b / s: two instructions per element (min/max + add, plus the add of the plain sum) instead of six.
      // The slices are quarter-interleaved float4 groups; degenerate scales fall back to the direct form.
      float sm = 0.f;
      {
        const float th = -b3 / s3;
        const int n4 = NZ >> 2;
        if (s3 != 0.f && fabsf(th) < 1e30f) {
          float sz = 0.f, sc = 0.f;
          const float4* z4 = reinterpret_cast<const float4*>(z);
          const int cnt = 4 * ((n4 - sub + nsub - 1) / nsub) + (sub == 0 ? NZ - (n4 << 2) : 0);
          if (slice_sums) {
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) sz += zsl[w * 32 + sub];
          } else {
            for (int i = sub; i < n4; i += nsub) { const float4 q = z4[i]; sz += (q.x + q.y) + (q.z + q.w); }
            if (sub == 0) for (int i = n4 << 2; i < NZ; ++i) sz += z[i];
          }
          if (s3 > 0.f) {
            for (int i = sub; i < n4; i += nsub) {
              const float4 q = z4[i];
              sc += (fmaxf(q.x, th) + fmaxf(q.y, th)) + (fmaxf(q.z, th) + fmaxf(q.w, th));
            }
            if (sub == 0) for (int i = n4 << 2; i < NZ; ++i) sc += fmaxf(z[i], th);
          } else {
            for (int i = sub; i < n4; i += nsub) {
              const float4 q = z4[i];
              sc += (fminf(q.x, th) + fminf(q.y, th)) + (fminf(q.z, th) + fminf(q.w, th));
            }
            if (sub == 0) for (int i = n4 << 2; i < NZ; ++i) sc += fminf(z[i], th);
          }
          const float nb = (float)cnt * b3;
          sm = fmaf(a3, fmaf(s3, sz, nb), (1.f - a3) * fmaf(s3, sc, nb));
        } else {
          for (int i = sub; i < NZ; i += nsub) sm += prelu(fmaf(s3, z[i], b3), a3);
        }
      }
      // context_conv2: (Tout, 1) convolution + BN + PReLU, max over the columns.  The channel's Tout weights sit in
      // registers for the thread's contiguous run of columns, two columns per pass (one 64-bit load of z per two FMAs).
      float mx2 = -INFINITY;
      const float b2 = W[t[CT_C2_B] + ch], a2 = W[t[CT_C2_A]];
      const float* w2 = W + t[CT_C2_WT] + ch;
      {
        const int per = (V3 + nsub - 1) / nsub, per2 = (per + 1) & ~1;          // even-sized runs: 8-byte aligned pairs (V3 is even)
        const int c_begin = sub * per2, c_end = c_begin + per2 < V3 ? c_begin + per2 : V3;
        if ((V3 & 1) == 0 && To <= 32) {
          float wr[32];
#pragma unroll
          for (int r = 0; r < 32; ++r) wr[r] = r < To ? w2[r * pad8i(H)] : 0.f;
          int col = c_begin;
          for (; col + 3 < c_end; col += 4) {                                   // four columns per pass: two loads in flight per FMA quad
            float acc0 = b2, acc1 = b2, acc2 = b2, acc3 = b2;
#pragma unroll
            for (int r = 0; r < 32; ++r)
              if (r < To) {
                const float2 q = *reinterpret_cast<const float2*>(z + r * V3 + col);
                const float2 u = *reinterpret_cast<const float2*>(z + r * V3 + col + 2);
                acc0 = fmaf(wr[r], q.x, acc0); acc1 = fmaf(wr[r], q.y, acc1);
                acc2 = fmaf(wr[r], u.x, acc2); acc3 = fmaf(wr[r], u.y, acc3);
              }
            mx2 = fmaxf(mx2, fmaxf(fmaxf(prelu(acc0, a2), prelu(acc1, a2)), fmaxf(prelu(acc2, a2), prelu(acc3, a2))));
          }
          for (; col + 1 < c_end; col += 2) {
            float acc0 = b2, acc1 = b2;
#pragma unroll
            for (int r = 0; r < 32; ++r)
              if (r < To) {
                const float2 q = *reinterpret_cast<const float2*>(z + r * V3 + col);
                acc0 = fmaf(wr[r], q.x, acc0);
                acc1 = fmaf(wr[r], q.y, acc1);
              }
            mx2 = fmaxf(mx2, fmaxf(prelu(acc0, a2), prelu(acc1, a2)));
          }
          if (c_end > c_begin && ((c_end - c_begin) & 1)) {
            const int col = c_end - 1;
            float acc = b2;
            for (int r = 0; r < To; ++r) acc = fmaf(w2[r * pad8i(H)], z[r * V3 + col], acc);
            mx2 = fmaxf(mx2, prelu(acc, a2));
          }
        } else {
          for (int col = c_begin; col < c_end; ++col) {
            float acc = b2;
            for (int r = 0; r < To; ++r) acc = fmaf(w2[r * pad8i(H)], z[r * V3 + col], acc);
            mx2 = fmaxf(mx2, prelu(acc, a2));
          }
        }
      }
      if (sub < 4) { part1[sub * H + ch] = mx; part2[sub * H + ch] = mx2; part3[sub * H + ch] = sm; }
    }
    __syncthreads();
    for (int ch = tid; ch < H; ch += NT) {
      const int nsub = NT / H;
      float mx = -INFINITY, mx2 = -INFINITY, sm = 0.f;
      for (int s = 0; s < nsub; ++s) { mx = fmaxf(mx, part1[s * H + ch]); mx2 = fmaxf(mx2, part2[s * H + ch]); sm += part3[s * H + ch]; }
      yv[ch] = mx; yv[H + ch] = mx2; yv[2 * H + ch] = sm / NZ;
    }
    __syncthreads();
    for (int m = tid; m < 3 * To; m += NT) {                     // map{1,2,3}: Linear(H, Tout) + PReLU (:421-432, 468)
      const int g = m / To, o = m - g * To;
      const float* wt = W + t[CT_MAP_WT] + (size_t)g * H * Top + o;
      float acc = 0.f;
      for (int k = 0; k < H; ++k) acc = fmaf(wt[k * Top], yv[g * H + k], acc);
      y[m] = prelu(acc, W[t[CT_MAP_A] + g]);
    }
    __syncthreads();
    for (int m = tid; m < V + To; m += NT) {                     // fmap_s / fmap_t (+BN) (:434-440, 469-470)
      float acc;
      if (m < V) {
        const float* wt = W + t[CT_FS_WT] + m;
        acc = W[t[CT_FS_B] + m];
        for (int k = 0; k < 3 * To; ++k) acc = fmaf(wt[k * pad8i(V)], y[k], acc);
        joints[m] = acc;
        if (a.tap_joints) a.tap_joints[(size_t)b * V + m] = acc;
      } else {
        const int o = m - V;
        const float* wt = W + t[CT_FT_WT] + o;
        acc = W[t[CT_FT_B] + o];
        for (int k = 0; k < 3 * To; ++k) acc = fmaf(wt[k * Top], y[k], acc);
        disp[o] = acc;
        if (a.tap_disp) a.tap_disp[(size_t)b * To + o] = acc;
      }
    }
    __syncthreads();
    // ---- norm_map on seq_joints = disp (x) joints: Conv1d(k=1)+BN+PReLU, SE1d, Conv1d+BN+PReLU (:443-451, 471-472)
    // seq_joints[f][v] = disp[f] * joints[v] is rank one, so the first Conv1d factors: sum_f w[fo][f] disp[f] joints[v] =
    // joints[v] * d2[fo] -- To MACs per output frame instead of per output element
    for (int fo = tid; fo < To; fo += NT) {
      const float* wt = W + t[CT_N0_WT] + fo;
      float acc = 0.f;
      for (int f = 0; f < To; ++f) acc = fmaf(wt[f * Top], disp[f], acc);
      d2[fo] = acc;
    }
    __syncthreads();
    for (int i = tid; i < To * V; i += NT) {
      const int fo = i / V, v = i - fo * V;
      n1[i] = prelu(fmaf(joints[v], d2[fo], W[t[CT_N0_B] + fo]), W[t[CT_N0_A]]);
    }
    __syncthreads();
    for (int f = tid; f < To; f += NT) {
      float s = 0.f;
      for (int v = 0; v < V; ++v) s += n1[f * V + v];
      mean1[f] = s / V;
    }
    __syncthreads();
    for (int h = tid; h < S1; h += NT) {
      const float* wt = W + t[CT_NSE1_WT] + h;
      float acc = 0.f;
      for (int f = 0; f < To; ++f) acc = fmaf(wt[f * pad8i(S1)], mean1[f], acc);
      e1[h] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int f = tid; f < To; f += NT) {
      const float* wt = W + t[CT_NSE2_WT] + f;
      float acc = 0.f;
      for (int h = 0; h < S1; ++h) acc = fmaf(wt[h * Top], e1[h], acc);
      gate1[f] = sigmoidf(acc);
    }
    __syncthreads();
    for (int i = tid; i < To * V; i += NT) n1s[i] = n1[i] * gate1[i / V];
    __syncthreads();
    for (int i = tid; i < To * V; i += NT) {
      const int fo = i / V, v = i - fo * V;
      const float* wt = W + t[CT_N5_WT] + fo;
      float acc = W[t[CT_N5_B] + fo];
      for (int f = 0; f < To; ++f) acc = fmaf(wt[f * Top], n1s[f * V + v], acc);
      const float val = prelu(acc, W[t[CT_N5_A]]);
      n2[i] = val;
      if (a.tap_sjn) a.tap_sjn[(size_t)b * To * V + i] = val;
    }
    __syncthreads();
    // ---- fconv 1->3->3 (+BN+PReLU) (:454-460, 473)
    for (int i = tid; i < To * V; i += NT) {
      float f0[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) f0[k] = prelu(fmaf(W[t[CT_FC0_S] + k], n2[i], W[t[CT_FC0_B] + k]), W[t[CT_FC0_A]]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float acc = W[t[CT_FC3_B] + k];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc = fmaf(W[t[CT_FC3_WT] + j * 8 + k], f0[j], acc);
        const float val = prelu(acc, W[t[CT_FC3_A]]);
        f3[k * To * V + i] = val;
        if (a.tap_sjd) a.tap_sjd[((size_t)b * 3 + k) * To * V + i] = val;
      }
    }
    __syncthreads();
    // ---- SE2d with the frames as channels (:461, 474)
    for (int f = tid; f < To; f += NT) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k)
        for (int v = 0; v < V; ++v) s += f3[k * To * V + f * V + v];
      mean2[f] = s / V3;
    }
    __syncthreads();
    for (int h = tid; h < S2; h += NT) {
      const float* wt = W + t[CT_SE1_WT] + h;
      float acc = 0.f;
      for (int f = 0; f < To; ++f) acc = fmaf(wt[f * pad8i(S2)], mean2[f], acc);
      e2[h] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int f = tid; f < To; f += NT) {
      const float* wt = W + t[CT_SE2_WT] + f;
      float acc = 0.f;
      for (int h = 0; h < S2; ++h) acc = fmaf(wt[h * Top], e2[h], acc);
      gate2[f] = sigmoidf(acc);
    }
    __syncthreads();
    // ---- pred = x[:, -1:] + x8 + act (:595-597); optional per-frame MPJPE partial sums
    {
      const float* xl = a.x + ((size_t)b * Tin + (Tin - 1)) * V3;
      const float* x8 = a.x8 + (size_t)b * NZ;
      float* pr = a.pred + (size_t)b * NZ;
      const float* tg = a.target ? a.target + (size_t)b * NZ : nullptr;
      for (int i = tid; i < To * V; i += NT) {
        const int f = i / V, v = i - f * V;
        float e = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int idx = i * 3 + k;
          const float act = f3[k * To * V + i] * gate2[f];
          const float val = xl[v * 3 + k] + (x8[idx] + act);
          pr[idx] = val;
          if (tg) { const float dd = val - tg[idx]; e = fmaf(dd, dd, e); }
        }
        if (tg) atomicAdd(&fsum[f], sqrtf(e));
      }
    }
    __syncthreads();
    if (a.target && tid < To) facc += (double)fsum[tid];
    __syncthreads();
  }
  if (a.target && a.frame_sums && tid < To) atomicAdd(&a.frame_sums[tid], facc);
}

// Stand-alone MPJPE: err (B,T,V) and / or per-frame sums.
struct MpjpeArgs {
  const float* pred;
  const float* target;
  float* err;
  double* frame_sums;
  long long n;      // B*T*V
  int T, V;
};

__global__ void __launch_bounds__(256) mpjpe_kernel(const MpjpeArgs a) {
  __shared__ double fs[64];
  const int tid = threadIdx.x;
  for (int i = tid; i < 64; i += 256) fs[i] = 0.0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + tid; i < a.n; i += stride) {
    const float* p = a.pred + i * 3;
    const float* q = a.target + i * 3;
    const float d0 = p[0] - q[0], d1 = p[1] - q[1], d2 = p[2] - q[2];
    const float e = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
    if (a.err) a.err[i] = e;
    if (a.frame_sums) {
      const int f = (int)((i / a.V) % a.T);
      if (a.T <= 64) atomicAdd(&fs[f], (double)e);
      else atomicAdd(&a.frame_sums[f], (double)e);
    }
  }
  __syncthreads();
  if (a.frame_sums && a.T <= 64)
    for (int i = tid; i < a.T; i += 256) atomicAdd(&a.frame_sums[i], fs[i]);
}

}  // namespace cg
