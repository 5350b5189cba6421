// Fused DSTD-GC block (reference: models/CISTGCN/CISTGCN.py:273-390, restated in SURVEY.md App. A).
//
// One CTA owns one sample at a time (persistent loop over the batch).  The sample's normalised
// activation tile XN (Ci x T*V) stays resident in shared memory for the whole block; every stage
// (statistics, context gates, Map2Adj adjacency generation, TxT / VxV adjacency products, 1x1 channel
// mixes with folded BatchNorm + PReLU + residual, gating, compressor, squeeze-excitation, block
// residual) reads and writes shared memory only.  HBM sees the tile once in and once out.
//
// Shared-memory map (floats; offsets come from dstd_plan()):
//   XN  [Ci][TV]                      resident input tile (after global_norm)
//   A   [Cmax][TV]  | B [..]          two work tiles: Map2Adj hidden maps -> g1/g2 -> u1/u2 -> c
//                                     (B also hosts the expansor's hidden map "mid")
//   ADJ [TV*max(T,V)]                 row statistics scratch, then o / Adj_s (stored [t][q][v]), then o / Adj_t
//   SM                                small vectors (stats, gate activations, dseq/dsp, SE)
#pragma once
#include "../../include/cistgcn_b200.h"
#include "simt.h"

namespace cg {

struct DstdArgs {
  int d[CB_COUNT];
  const float* w;
  const float* in;
  float* out;
  float* tap_adj_s;
  float* tap_adj_t;
  float* tap_w1;
  float* tap_w2;
  int batch;
  int o_xn, o_ab, tile, o_adj, o_sm, smem_floats;
};

__host__ __device__ inline int pad4i(int n) { return (n + 3) & ~3; }
__host__ __device__ inline int pad8i(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

// Fills the shared-memory plan fields of `a` from the descriptor.
inline void dstd_plan(DstdArgs& a) {
  const int* d = a.d;
  const int Ci = d[CB_CI], Co = d[CB_CO], T = d[CB_T], V = d[CB_V], Ch = d[CB_CH], Cg = d[CB_CG], Hs = d[CB_HS];
  const int TV = T * V, cmax = imax(Ci, Co), big = TV * imax(T, V);
  a.tile = pad4i(cmax * TV);
  a.o_xn = 0;
  a.o_ab = pad4i(Ci * TV);
  const int ab = a.tile + imax(a.tile, pad4i(big));
  a.o_adj = a.o_ab + ab;
  const int adj = imax(pad4i(big), pad4i(2 * Ci * T + 2 * Ci));
  a.o_sm = a.o_adj + adj;
  const int sm = pad4i(2 + 2 * T) + pad4i(2 * Cg * V) + 3 * pad4i(2 * Co) + pad4i(2 * Ch * V) + pad4i(2 * Ch * T) +
                 2 * pad4i(2 * TV) + 2 * pad4i(Co) + pad4i(Hs);
  a.smem_floats = a.o_sm + sm;
}

// ---------------------------------------------------------------------------------------------
// out(m, n) = sum_k Wt[k*Mp + m] * X(k, n), X given as up to two stacked row blocks in shared memory.
// Work unit = one warp x (TM output rows) x (32*TN columns); lane owns columns nb + 32*j, so the
// activation loads are conflict-free and the weight loads are warp-uniform broadcasts.
// INPLACE: the epilogue may overwrite X; a pass then holds whole column groups only and every
// thread meets the two barriers of every pass (requires ceil(M/TM) <= NT/32).
// ---------------------------------------------------------------------------------------------
template <int TM, int TN, int NT, bool INPLACE, class EPI>
CG_DEV void gemm_rows(const float* __restrict__ Wt, int Mp, int M, int N,
                      const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, EPI epi) {
  constexpr int NW = NT / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtiles = (M + TM - 1) / TM;
  const int ngroups = (N + 32 * TN - 1) / (32 * TN);
  const int total = mtiles * ngroups;
  const int per_pass = INPLACE ? (NW / mtiles) * mtiles : NW;
  for (int base = 0; base < total; base += per_pass) {
    const int item = base + warp;
    const bool active = warp < per_pass && item < total;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    int m0 = 0, nb = lane;
    if (active) {
      m0 = (item % mtiles) * TM;
      nb = (item / mtiles) * 32 * TN + lane;
      const float* wp = Wt + m0;
      int ncl[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) ncl[j] = (nb + 32 * j < N) ? nb + 32 * j : 0;   // clamp: keeps loads in range
#pragma unroll 4
      for (int k = 0; k < K1; ++k) {
        float w[TM], x[TN];
        load_vec<TM>(wp + (size_t)k * Mp, w);
#pragma unroll
        for (int j = 0; j < TN; ++j) x[j] = X1[k * ld1 + ncl[j]];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
      }
      wp += (size_t)K1 * Mp;
#pragma unroll 4
      for (int k = 0; k < K2; ++k) {
        float w[TM], x[TN];
        load_vec<TM>(wp + (size_t)k * Mp, w);
#pragma unroll
        for (int j = 0; j < TN; ++j) x[j] = X2[k * ld2 + ncl[j]];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
      }
    }
    if (INPLACE) __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
          if (m0 + i < M && nb + 32 * j < N) epi(m0 + i, nb + 32 * j, acc[i][j]);
    }
    if (INPLACE) __syncthreads();
  }
}

// Picks the widest column tile that still gives every warp work, then runs gemm_rows.
template <int NT, bool INPLACE, class EPI>
CG_DEV void gemm_rows_auto(const float* __restrict__ Wt, int Mp, int M, int N,
                           const float* X1, int ld1, int K1, const float* X2, int ld2, int K2, EPI epi) {
  constexpr int NW = NT / 32;
  const int mtiles = (M + 7) / 8;
  if (mtiles * ((N + 127) / 128) >= NW)
    gemm_rows<8, 4, NT, INPLACE>(Wt, Mp, M, N, X1, ld1, K1, X2, ld2, K2, epi);
  else if (mtiles * ((N + 63) / 64) >= NW)
    gemm_rows<8, 2, NT, INPLACE>(Wt, Mp, M, N, X1, ld1, K1, X2, ld2, K2, epi);
  else
    gemm_rows<8, 1, NT, INPLACE>(Wt, Mp, M, N, X1, ld1, K1, X2, ld2, K2, epi);
}

// ---------------------------------------------------------------------------------------------
// Collapsing convolution ((R,1) or (1,R) kernels):
//   out(m, n) = sum_{c<C, r<R} Wt[(c*R + r)*Mp + m] * X[c*ldc + r*N + n],   n < N (compile time).
// Lane -> (m sub-tile, n): 32/N sub-tiles share a warp so short rows still fill the lanes.
// `rot` rotates the warp assignment so that back-to-back calls land on different warps.
// ---------------------------------------------------------------------------------------------
template <int TM, int N, int NT, class EPI>
CG_DEV void kconv(const float* __restrict__ Wt, int Mp, int M, int C, int R,
                  const float* X, int ldc, int rot, EPI epi) {
  constexpr int NW = NT / 32;
  constexpr int MS = (32 / N) > 0 ? (32 / N) : 1;
  static_assert(N <= 32, "kconv: row length must fit a warp");
  const int lane = threadIdx.x & 31;
  const int warp = ((threadIdx.x >> 5) + NW - (rot % NW)) % NW;
  const int msub = lane / N, n = lane % N;
  const bool lane_ok = lane < MS * N;
  const int mtiles = (M + TM - 1) / TM;
  const int witems = (mtiles + MS - 1) / MS;
  for (int item = warp; item < witems; item += NW) {
    const int mt = item * MS + msub;
    const bool ok = lane_ok && mt < mtiles;
    const int m0 = ok ? mt * TM : 0;
    float acc[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* xp = X + c * ldc + n;
      const float* wp = Wt + (size_t)(c * R) * Mp + m0;
#pragma unroll 2
      for (int r = 0; r < R; ++r) {
        float w[TM];
        load_vec<TM>(wp + (size_t)r * Mp, w);
        const float x = xp[r * N];
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[i] = fmaf(w[i], x, acc[i]);
      }
    }
    if (ok) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
        if (m0 + i < M) epi(m0 + i, n, acc[i]);
    }
  }
}

template <int N, int NT, class EPI>
CG_DEV void kconv_auto(const float* __restrict__ Wt, int Mp, int M, int C, int R,
                       const float* X, int ldc, int rot, EPI epi) {
  if (M >= 64) kconv<8, N, NT>(Wt, Mp, M, C, R, X, ldc, rot, epi);
  else if (M >= 32) kconv<4, N, NT>(Wt, Mp, M, C, R, X, ldc, rot, epi);
  else kconv<2, N, NT>(Wt, Mp, M, C, R, X, ldc, rot, epi);
}

// ---------------------------------------------------------------------------------------------
// adjacency products (ConvTemporalGraphical, CISTGCN.py:110,117,123)
// ---------------------------------------------------------------------------------------------
// "space" domain: g1[c][q][v] = sum_t XN[c][t][v] * Adj_s[v][t][q], Adj_s held as adjT[(t*T+q)*V + v].
template <int T, int V, int TC, int NT>
CG_DEV void gcn_space(const float* XN, const float* adjT, float* G, int C) {
  constexpr int TV = T * V;
  const int nct = (C + TC - 1) / TC;
  for (int item = threadIdx.x; item < nct * V; item += NT) {
    const int v = item % V, c0 = (item / V) * TC;
    float acc[TC][T];
#pragma unroll
    for (int i = 0; i < TC; ++i)
#pragma unroll
      for (int q = 0; q < T; ++q) acc[i][q] = 0.f;
    for (int t = 0; t < T; ++t) {
      float xv[TC];
#pragma unroll
      for (int i = 0; i < TC; ++i) xv[i] = (c0 + i < C) ? XN[(c0 + i) * TV + t * V + v] : 0.f;
#pragma unroll
      for (int q = 0; q < T; ++q) {
        const float aq = adjT[(t * T + q) * V + v];
#pragma unroll
        for (int i = 0; i < TC; ++i) acc[i][q] = fmaf(xv[i], aq, acc[i][q]);
      }
    }
#pragma unroll
    for (int i = 0; i < TC; ++i)
      if (c0 + i < C) {
#pragma unroll
        for (int q = 0; q < T; ++q) G[(c0 + i) * TV + q * V + v] = acc[i][q];
      }
  }
}

// "time" domain: g2[c][t][w] = sum_v XN[c][t][v] * Adj_t[t][v][w]  (natural layout).
template <int T, int V, int TC, int NT>
CG_DEV void gcn_time(const float* XN, const float* adj, float* G, int C) {
  constexpr int TV = T * V, VV = V * V;
  constexpr int TW = (V % 11 == 0) ? 11 : ((V % 9 == 0) ? 9 : ((V % 5 == 0) ? 5 : 1));
  constexpr int NWG = V / TW;
  const int nct = (C + TC - 1) / TC;
  for (int item = threadIdx.x; item < nct * T * NWG; item += NT) {
    const int w0 = (item % NWG) * TW, t = (item / NWG) % T, c0 = (item / (NWG * T)) * TC;
    float acc[TC][TW];
#pragma unroll
    for (int i = 0; i < TC; ++i)
#pragma unroll
      for (int j = 0; j < TW; ++j) acc[i][j] = 0.f;
    for (int v = 0; v < V; ++v) {
      float xv[TC];
#pragma unroll
      for (int i = 0; i < TC; ++i) xv[i] = (c0 + i < C) ? XN[(c0 + i) * TV + t * V + v] : 0.f;
      const float* ap = adj + t * VV + v * V + w0;
#pragma unroll
      for (int j = 0; j < TW; ++j) {
        const float aj = ap[j];
#pragma unroll
        for (int i = 0; i < TC; ++i) acc[i][j] = fmaf(xv[i], aj, acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < TC; ++i)
      if (c0 + i < C) {
#pragma unroll
        for (int j = 0; j < TW; ++j) G[(c0 + i) * TV + t * V + w0 + j] = acc[i][j];
      }
  }
}

// ---------------------------------------------------------------------------------------------
template <int T, int V, int NT>
__global__ void __launch_bounds__(NT) dstd_block_kernel(const DstdArgs a) {
  CG_DYN_SMEM(smem);
  constexpr int TV = T * V, TT = T * T, VV = V * V;
  const int tid = threadIdx.x;
  const int* d = a.d;
  const float* __restrict__ W = a.w;
  const int Ci = d[CB_CI], Co = d[CB_CO], Ch = d[CB_CH], Cg = d[CB_CG], Hs = d[CB_HS];
  const bool has_res = d[CB_HAS_RES] != 0, interp = d[CB_INTERP] != 0;
  const int Cop = pad8i(Co);

  float* XN = smem + a.o_xn;
  float* A = smem + a.o_ab;
  float* Bt = A + a.tile;
  float* ADJ = smem + a.o_adj;
  float* p = smem + a.o_sm;
  float* stats = p;   p += pad4i(2 + 2 * T);
  float* h1 = p;      p += pad4i(2 * Cg * V);
  float* h2 = p;      p += pad4i(2 * Co);
  float* zg = p;      p += pad4i(2 * Co);
  float* wg = p;      p += pad4i(2 * Co);
  float* dseqp = p;   p += pad4i(2 * Ch * V);
  float* dspp = p;    p += pad4i(2 * Ch * T);
  float* dseq = p;    p += pad4i(2 * TV);
  float* dsp = p;     p += pad4i(2 * TV);
  float* semean = p;  p += pad4i(Co);
  float* gate = p;    p += pad4i(Co);
  float* hid = p;
  // row statistics live in the (still unused) adjacency region
  float* rowmean = ADJ;
  float* rowvar = rowmean + Ci * T;
  float* chmean = rowvar + Ci * T;
  float* chstd = chmean + Ci;

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    // ---------------- P1: load + global_norm (:375); block 0 builds the 10 features (:568-577)
    if (d[CB_IN_MODE] == 1) {
      const float* src = a.in + (size_t)b * d[CB_IN_SB];
      float* raw = A;
      for (int i = tid; i < TV * 3; i += NT) raw[i] = src[i];
      __syncthreads();
      const float* gs = W + d[CB_GN_S];
      const float* gb = W + d[CB_GN_B];
      for (int n = tid; n < TV; n += NT) {
        const int t = n / V;
        float f[10];
        float sp = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float p0 = raw[n * 3 + k];
          float vel, acc;
          if (t < T - 1) {
            const float p1 = raw[(n + V) * 3 + k];
            vel = p1 - p0;
            const float veln = (t < T - 2) ? raw[(n + 2 * V) * 3 + k] - p1 : p1;   // vel[T-1] = x[T-1]
            acc = veln - vel;
          } else {
            vel = p0;      // vel[:, -1] = x[:, -1]
            acc = p0;      // acc[:, -1] = vel[:, -1]
          }
          f[k] = p0; f[3 + k] = acc; f[6 + k] = vel;
          sp = fmaf(vel, vel, sp);
        }
        f[9] = sqrtf(sp);
#pragma unroll
        for (int c = 0; c < 10; ++c) XN[c * TV + n] = fmaf(gs[c], f[c], gb[c]);
      }
    } else {
      const float* src = a.in + (size_t)b * d[CB_IN_SB];
      const int sc = d[CB_IN_SC], st = d[CB_IN_ST], sv = d[CB_IN_SV];
      const float* gs = W + d[CB_GN_S];
      const float* gb = W + d[CB_GN_B];
      for (int i = tid; i < Ci * TV; i += NT) {
        const int c = i / TV, n = i - c * TV, t = n / V, v = n - t * V;
        XN[i] = fmaf(gs[c], src[c * sc + t * st + v * sv], gb[c]);
      }
    }
    __syncthreads();

    // ---------------- P2: row statistics | gate conv (T,1) | Map2Adj first 1x1 convs  (all read XN only)
    for (int r = tid; r < Ci * T; r += NT) {                  // r = c*T + t, row of V joints
      const float* xp = XN + (r / T) * TV + (r % T) * V;
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) s += xp[v];
      const float mu = s / V;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { const float dd = xp[v] - mu; q = fmaf(dd, dd, q); }
      rowmean[r] = mu;
      rowvar[r] = q / (V - 1);                                 // Bessel-corrected, like torch.std
    }
    {
      const float* gb = W + d[CB_G0_B];
      const float* ga = W + d[CB_G0_A];
      kconv_auto<V, NT>(W + d[CB_G0_WT], pad8i(2 * Cg), 2 * Cg, Ci, T, XN, TV, 0,
                        [&](int m, int v, float acc) { h1[m * V + v] = prelu(acc + gb[m], ga[m / Cg]); });
    }
    if (interp) {
      const float* ab = W + d[CB_A0_B];
      const float* aa = W + d[CB_A0_A];
      gemm_rows<8, 2, NT, false>(W + d[CB_A0_WT], pad8i(4 * Ch), 4 * Ch, TV, XN, TV, Ci, nullptr, 0, 0,
                                 [&](int m, int n, float acc) {
                                   const int br = m / Ch, r = m - br * Ch;
                                   const float val = prelu(acc + ab[m], aa[br]);
                                   float* tl = A + (br >> 1) * a.tile;
                                   if ((br & 1) == 0) tl[r * TV + n] = val;                // time_compress map  [c][t][v]
                                   else { const int t = n / V, v = n - t * V;
                                          tl[(Ch + r) * TV + v * T + t] = val; }           // joint_compress map [c][v][t]
                                 });
    }
    __syncthreads();

    // ---------------- P3: channel statistics | gate conv (1,V) | Map2Adj collapsing convs
    for (int c = tid; c < Ci; c += NT) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += rowmean[c * T + t];
      const float cm = s / T;
      float ss = 0.f;
      for (int t = 0; t < T; ++t) { const float dm = rowmean[c * T + t] - cm; ss += (V - 1) * rowvar[c * T + t] + V * dm * dm; }
      chmean[c] = cm;
      chstd[c] = sqrtf(ss / (TV - 1));
    }
    for (int t = tid - 64; t >= 0 && t < T; t += NT) {        // (threads 64.. so they overlap the loop above)
      float s = 0.f, s2 = 0.f;
      for (int c = 0; c < Ci; ++c) { s += rowmean[c * T + t]; s2 += sqrtf(rowvar[c * T + t]); }
      const float m2 = s2 / Ci;
      float q = 0.f;
      for (int c = 0; c < Ci; ++c) { const float dd = sqrtf(rowvar[c * T + t]) - m2; q = fmaf(dd, dd, q); }
      stats[1 + t] = s / Ci;
      stats[2 + T + t] = sqrtf(q / (Ci - 1));
    }
    for (int m = tid; m < 2 * Co; m += NT) {
      const int g = m / Co, o = m - g * Co, K = Cg * V;
      const float* wt = W + d[CB_G4_WT] + (size_t)g * K * Cop + o;
      const float* hp = h1 + g * K;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(wt[(size_t)k * Cop], hp[k], acc);
      h2[m] = prelu(acc + W[d[CB_G4_B] + m], W[d[CB_G4_A] + g]);
    }
    if (interp) {
#pragma unroll
      for (int L = 0; L < 2; ++L) {
        const float* tl = A + L * a.tile;
        const float* tb = W + d[CB_TC3_B_S + L];
        const float* jb = W + d[CB_JC3_B_S + L];
        float* dq = dseqp + L * Ch * V;
        float* dp = dspp + L * Ch * T;
        kconv_auto<V, NT>(W + d[CB_TC3_WT_S + L], pad8i(Ch), Ch, Ch, T, tl, TV, 2 + 3 * L,
                          [&](int m, int v, float acc) { dq[m * V + v] = acc + tb[m]; });
        kconv_auto<T, NT>(W + d[CB_JC3_WT_S + L], pad8i(Ch), Ch, Ch, V, tl + Ch * TV, TV, 4 + 3 * L,
                          [&](int m, int t, float acc) { dp[m * T + t] = acc + jb[m]; });
      }
    }
    __syncthreads();

    // ---------------- P4: scalar statistics + gate MLP layer 0 | dim_seq / dim_space
    if (tid == 0) {
      float s = 0.f, s2 = 0.f;
      for (int c = 0; c < Ci; ++c) { s += chmean[c]; s2 += chstd[c]; }
      const float m2 = s2 / Ci;
      float q = 0.f;
      for (int c = 0; c < Ci; ++c) { const float dd = chstd[c] - m2; q = fmaf(dd, dd, q); }
      stats[0] = s / Ci;
      stats[1 + T] = sqrtf(q / (Ci - 1));
    }
    if (interp) {
      for (int i = tid; i < 2 * TV; i += NT) {                // dim_seq[L][t'][v] = sum_o W6[o][t'] * dseqp[L][o][v]
        const int L = i / TV, r = i - L * TV, tq = r / V, v = r - tq * V;
        const float* wt = W + d[CB_TC6_WT_S + L] + tq;
        const float* xp = dseqp + L * Ch * V + v;
        float acc = 0.f;
        for (int o = 0; o < Ch; ++o) acc = fmaf(wt[o * pad8i(T)], xp[o * V], acc);
        dseq[i] = acc;
      }
      for (int i = tid; i < 2 * TV; i += NT) {                // dim_space[L][v'][t] = sum_o W6[o][v'] * dspp[L][o][t]
        const int L = i / TV, r = i - L * TV, vq = r / T, t = r - vq * T;
        const float* wt = W + d[CB_JC6_WT_S + L] + vq;
        const float* xp = dspp + L * Ch * T + t;
        float acc = 0.f;
        for (int o = 0; o < Ch; ++o) acc = fmaf(wt[o * pad8i(V)], xp[o * T], acc);
        dsp[i] = acc;
      }
    }
    __syncthreads();
    for (int m = tid; m < 2 * Co; m += NT) {                  // map_{s,t}.0 on cat(h2, stats)  (:341-344, 378, 380)
      const int g = m / Co, o = m - g * Co;
      const float* wt = W + d[CB_M0_WT] + (size_t)g * (Co + 2 + 2 * T) * Cop + o;
      float acc = 0.f;
      for (int k = 0; k < Co; ++k) acc = fmaf(wt[(size_t)k * Cop], h2[g * Co + k], acc);
      for (int k = 0; k < 2 + 2 * T; ++k) acc = fmaf(wt[(size_t)(Co + k) * Cop], stats[k], acc);
      zg[m] = prelu(acc + W[d[CB_M0_B] + m], W[d[CB_M0_A] + g]);
    }
    // ---------------- P5: space-domain outer product o[v'][t][q] = dsp[v'][t] * dseq[q][v']  (:187)
    if (interp) {
      for (int i = tid; i < V * TT; i += NT) {
        const int vq = i / TT, r = i - vq * TT, t = r / T, q = r - t * T;
        ADJ[i] = dsp[vq * T + t] * dseq[q * V + vq];
      }
    } else {
      const float* as = W + d[CB_ADJ_S];                      // static (V,T,T) -> [t][q][v]
      for (int i = tid; i < V * TT; i += NT) { const int v = i / TT, r = i - v * TT; ADJ[r * V + v] = as[i]; }
    }
    __syncthreads();
    for (int m = tid; m < 2 * Co; m += NT) {                  // map_{s,t}.4 -> gates w1, w2 (:345, 381-382)
      const int g = m / Co, o = m - g * Co;
      const float* wt = W + d[CB_M4_WT] + (size_t)g * Co * Cop + o;
      float acc = 0.f;
      for (int k = 0; k < Co; ++k) acc = fmaf(wt[(size_t)k * Cop], zg[g * Co + k], acc);
      wg[m] = acc;
      float* tp = g == 0 ? a.tap_w1 : a.tap_w2;
      if (tp) tp[(size_t)b * Co + o] = acc;
    }
    // ---------------- P6/P7: expansor over the joint axis -> Adj_s (kept as [t][q][v])
    if (interp) {
      {
        const float* eb = W + d[CB_E0_B_S];
        const float ea = W[d[CB_E0_A_S]];
        gemm_rows<8, 1, NT, false>(W + d[CB_E0_WT_S], pad8i(V), V, TT, ADJ, TT, V, nullptr, 0, 0,
                                   [&](int m, int n, float acc) { Bt[m * TT + n] = prelu(acc + eb[m], ea); });
      }
      __syncthreads();
      {
        float* tp = a.tap_adj_s ? a.tap_adj_s + (size_t)b * V * TT : nullptr;
        gemm_rows<8, 1, NT, false>(W + d[CB_E4_WT_S], pad8i(V), V, TT, Bt, TT, V, nullptr, 0, 0,
                                   [&](int m, int n, float acc) { ADJ[n * V + m] = acc; if (tp) tp[m * TT + n] = acc; });
      }
      __syncthreads();
    }
    // ---------------- P8: g1 = XN x_t Adj_s -> A
    if (Ci >= 4) gcn_space<T, V, 4, NT>(XN, ADJ, A, Ci);
    else gcn_space<T, V, 1, NT>(XN, ADJ, A, Ci);
    __syncthreads();
    // ---------------- P9-P11: time-domain outer product + expansor over the frame axis -> Adj_t
    if (interp) {
      for (int i = tid; i < T * VV; i += NT) {                // o[t'][v][w] = dsp[v][t'] * dseq[t'][w]
        const int tq = i / VV, r = i - tq * VV, v = r / V, w = r - v * V;
        ADJ[i] = dsp[TV + v * T + tq] * dseq[TV + tq * V + w];
      }
      __syncthreads();
      {
        const float* eb = W + d[CB_E0_B_T];
        const float ea = W[d[CB_E0_A_T]];
        gemm_rows<4, 2, NT, false>(W + d[CB_E0_WT_T], pad8i(T), T, VV, ADJ, VV, T, nullptr, 0, 0,
                                   [&](int m, int n, float acc) { Bt[m * VV + n] = prelu(acc + eb[m], ea); });
      }
      __syncthreads();
      {
        float* tp = a.tap_adj_t ? a.tap_adj_t + (size_t)b * T * VV : nullptr;
        gemm_rows<4, 2, NT, false>(W + d[CB_E4_WT_T], pad8i(T), T, VV, Bt, VV, T, nullptr, 0, 0,
                                   [&](int m, int n, float acc) { ADJ[m * VV + n] = acc; if (tp) tp[m * VV + n] = acc; });
      }
    } else {
      const float* at = W + d[CB_ADJ_T];
      for (int i = tid; i < T * VV; i += NT) ADJ[i] = at[i];
    }
    __syncthreads();
    // ---------------- P12: g2 = XN x_v Adj_t -> B
    if (Ci >= 4) gcn_time<T, V, 4, NT>(XN, ADJ, Bt, Ci);
    else gcn_time<T, V, 1, NT>(XN, ADJ, Bt, Ci);
    __syncthreads();
    // ---------------- P13: x_k = PReLU(BN(W g_k + b) + res); u_k = PReLU(BN(w_k * x_k))   (:266-268, 388)
#pragma unroll
    for (int L = 0; L < 2; ++L) {
      float* G = L == 0 ? A : Bt;
      const float* tb = W + d[CB_TCN_B_S + L];
      const float ta = W[d[CB_TCN_A_S + L]];
      const float* ps = W + d[CB_P_S_S + L];
      const float* pb = W + d[CB_P_B_S + L];
      const float pa = W[d[CB_P_A_S + L]];
      const float* wk = wg + L * Co;
      gemm_rows_auto<NT, true>(W + d[CB_TCN_WT_S + L], Cop, Co, TV, G, TV, Ci, XN, TV, has_res ? Ci : 0,
                               [&](int m, int n, float acc) {
                                 float v = acc + tb[m];
                                 if (!has_res) v += XN[m * TV + n];
                                 v = prelu(v, ta);
                                 v = fmaf(ps[m], wk[m] * v, pb[m]);
                                 G[m * TV + n] = prelu(v, pa);
                               });
    }
    // ---------------- P14: compressor 1x1 over cat(u1, u2) + BN + PReLU -> A   (:305-307)
    {
      const float* cb = W + d[CB_CP_B];
      const float ca = W[d[CB_CP_A]];
      gemm_rows_auto<NT, true>(W + d[CB_CP_WT], Cop, Co, TV, A, TV, Co, Bt, TV, Co,
                               [&](int m, int n, float acc) { A[m * TV + n] = prelu(acc + cb[m], ca); });
    }
    // ---------------- P15-P17: squeeze-excitation (SE.py:37-41)
    for (int m = tid >> 5; m < Co; m += NT / 32) {
      float s = 0.f;
      for (int n = tid & 31; n < TV; n += 32) s += A[m * TV + n];
      s = warp_sum(s);
      if ((tid & 31) == 0) semean[m] = s / TV;
    }
    __syncthreads();
    for (int h = tid; h < Hs; h += NT) {
      const float* wt = W + d[CB_SE1_WT] + h;
      float acc = 0.f;
      for (int c = 0; c < Co; ++c) acc = fmaf(wt[c * pad8i(Hs)], semean[c], acc);
      hid[h] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int o = tid; o < Co; o += NT) {
      const float* wt = W + d[CB_SE2_WT] + o;
      float acc = 0.f;
      for (int h = 0; h < Hs; ++h) acc = fmaf(wt[h * Cop], hid[h], acc);
      gate[o] = sigmoidf(acc);
    }
    __syncthreads();
    // ---------------- P18: out = c * gate + residual(xn)   (:390)
    {
      float* dst = a.out + (size_t)b * d[CB_OUT_SB];
      const int sc = d[CB_OUT_SC], st = d[CB_OUT_ST], sv = d[CB_OUT_SV];
      if (has_res) {
        const float* rb = W + d[CB_RS_B];
        gemm_rows<8, 2, NT, false>(W + d[CB_RS_WT], Cop, Co, TV, XN, TV, Ci, nullptr, 0, 0,
                                   [&](int m, int n, float acc) {
                                     const int t = n / V, v = n - t * V;
                                     dst[m * sc + t * st + v * sv] = fmaf(A[m * TV + n], gate[m], acc + rb[m]);
                                   });
      } else {
        for (int i = tid; i < Co * TV; i += NT) {
          const int m = i / TV, n = i - m * TV, t = n / V, v = n - t * V;
          dst[m * sc + t * st + v * sv] = fmaf(A[i], gate[m], XN[i]);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace cg
