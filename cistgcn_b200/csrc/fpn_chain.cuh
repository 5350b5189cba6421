// Time-extrapolator stack: 4 x FPN (CISTGCN.py:38-79) with the caller's PReLU / residual (:584-586),
// then dim_conversor (:541-545) and the cumulative sum over output frames (:588-589), one kernel.
//
// Frames are the channel axis here; the 2-D map is (H = F feature rows) x (W = V joints).
// One CTA owns one sample at a time, two CTAs per SM.  The sample's map lives in shared memory,
// column-padded ([c][h][XS], data at columns 4..4+V, zeros around it; out-of-range rows read a shared
// zero row), is updated in place layer after layer and never returns to HBM until x7 is written.
//
// Hot loop (56 % of the model's FLOPs): the three dilation branches run one after another with ALL
// warps on the same branch, so the inner body (one input channel x one kernel row: 180 FFMAs) is a
// few KB of code shared by every warp and stays in the instruction cache.  A thread owns 5 output
// channels x 12 (or 10) joints of one row; per body it loads the input window once (5 x LDS.128) and 15
// folded weights (4 x LDG.128, warp-uniform per channel tile, L1-resident) for 180 FFMAs.  100 work items
// per branch = 4 warps x 25 lanes: one warp per scheduler, equal work.  Each
// branch's slice of the 1x1 `compress` convolution is accumulated right after the branch, so only one
// branch output (25 x F x V) is ever staged.
#pragma once
#include "../../include/cistgcn_b200.h"
#include "dstd_block.cuh"
#include "simt.h"

namespace cg {

constexpr int FPN_XS = 36;        // padded row stride of the resident map (floats)
constexpr int FPN_LPAD = 4;       // data starts at column 4 (keeps rows float4-aligned)
constexpr int FPN_NT = 128;       // 4 warps, one per scheduler; 2 CTAs per SM
constexpr int FPN_MAX_LAYERS = CISTGCN_MAX_FPN;

struct FpnArgs {
  int f[FPN_MAX_LAYERS][CF_COUNT];
  int t[CT_COUNT];
  int n_layers;
  const float* w;
  const float* in;     // (B, Tin, F, V)
  float* x7;           // (B, Tout, V, 3)
  int batch;
  int o_x, o_zero, o_b, o_out, o_ring, ring_floats, o_misc, smem_floats;
};

inline void fpn_plan(FpnArgs& a) {
  const int To = a.t[CT_TOUT], F = a.t[CT_F], V = a.t[CT_V];
  a.o_x = 0;
  a.o_zero = To * F * FPN_XS;
  a.o_b = a.o_zero + 40;
  a.o_out = a.o_b + pad4i(To * F * V);
  a.o_ring = a.o_out + pad4i(To * F * V);
  a.ring_floats = To * pad8i(To);
  a.o_misc = a.o_ring + 2 * a.ring_floats;
  a.smem_floats = a.o_misc + 2 * pad4i(To) + pad4i(3 * To * V);
}

// One dilation branch: conv3x3(dil = D, pad = D) + folded BN + PReLU for `To` output channels into
// Bout [To][F*V].  Work item = (channel tile of 5, row h, column tile of TN).
template <int V, int D, int TN, int NT>
CG_DEV void fpn_branch(const float* __restrict__ Wd, const float* __restrict__ bias, float slope,
                       const float* X, const float* zero_row, float* Bout, int Cin, int To, int F) {
  constexpr int VW = (TN % 4 == 0) ? 4 : 2;                         // window load width
  constexpr int WIN = ((FPN_LPAD + TN + 3) + VW - 1) / VW * VW;     // floats of row window per item
  constexpr int NWT = (V + TN - 1) / TN;                            // column tiles per row
  static_assert((NWT - 1) * TN + WIN <= FPN_XS, "row padding too small for this joint count");
  const int notile = To / 5;
  const int items = notile * F * NWT;
  // every warp gets the same number of items (100 items -> 4 x 25): equal work per scheduler
  const int ipw = (items + NT / 32 - 1) / (NT / 32);
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
  for (int it0 = ln; it0 < ipw; it0 += 32) {
    const int item = wid * ipw + it0;
    if (item >= items) break;
    // channel tile fastest: neighbouring lanes read the same input window (broadcast), fewer LSU wavefronts
    const int ot = item % notile, wt = (item / notile) % NWT, h = item / (notile * NWT);
    const int w0 = wt * TN;
    float acc[5][TN];
#pragma unroll
    for (int m = 0; m < 5; ++m)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[m][j] = 0.f;
    const float* wp = Wd + (size_t)ot * 16;
    const int wstride = notile * 16;
    const int nit = Cin * 3;                                   // (input channel, kernel row) pairs
    auto step = [&](int it, const float (&wv)[16]) {
      const int c = it / 3, kh = it - c * 3;
      const int hh = h + (kh - 1) * D;
      const float* row = ((hh >= 0 && hh < F) ? X + (c * F + hh) * FPN_XS : zero_row) + w0;
      float xr[WIN];
#pragma unroll
      for (int q = 0; q < WIN / VW; ++q) {
        float tmp[VW];
        lds_vec<VW>(row + VW * q, tmp);
#pragma unroll
        for (int e = 0; e < VW; ++e) xr[VW * q + e] = tmp[e];
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
          for (int j = 0; j < TN; ++j)
            acc[m][j] = fmaf(wv[kw * 5 + m], xr[FPN_LPAD + j + (kw - 1) * D], acc[m][j]);
    };
    // weights are double-buffered in registers: the (L1-resident) loads of step it+1 fly under step it
    float wa[16], wb[16];
    load_vec<16>(wp, wa);
    int it = 0;
#pragma unroll 1
    for (; it + 2 <= nit; it += 2) {
      load_vec<16>(wp + (size_t)(it + 1) * wstride, wb);
      step(it, wa);
      if (it + 2 < nit) load_vec<16>(wp + (size_t)(it + 2) * wstride, wa);
      step(it + 1, wb);
    }
    if (it < nit) step(it, wa);
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const int o = ot * 5 + m;
      const float bo = bias[o];
      float* dst = Bout + o * (F * V) + h * V + w0;
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (w0 + j < V) dst[j] = prelu(acc[m][j] + bo, slope);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(FPN_NT, 2) fpn_chain_kernel(const FpnArgs a) {
  CG_DYN_SMEM(smem);
  constexpr int NT = FPN_NT;
  constexpr int TN = (V == 22) ? 12 : (V == 18 ? 10 : 8);       // 22 -> 12 + 10, 18 -> 10 + 8: two column tiles per row
  constexpr int FVC = 10 * V;                                    // F is fixed at 10 (in_ch, CISTGCN.py:512)
  const int tid = threadIdx.x, warp = tid >> 5;
  const float* __restrict__ W = a.w;
  const int Tin = a.t[CT_TIN], To = a.t[CT_TOUT], F = a.t[CT_F];
  const int FV = F * V;
  const int Top = pad8i(To);
  float* X = smem + a.o_x;
  float* zero_row = smem + a.o_zero;
  float* Bb = smem + a.o_b;
  float* OUT = smem + a.o_out;
  float* ring = smem + a.o_ring;
  const int rb = a.ring_floats;
  float* avg = smem + a.o_misc;
  float* cst = avg + pad4i(To);
  float* y6 = cst + pad4i(To);              // dim_conversor output (To, V, 3) before the cumsum

  // zero the padding once: data columns are the only ones ever rewritten
  for (int i = tid; i < To * F * FPN_XS + 40; i += NT) X[i] = 0.f;
  __syncthreads();

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    {   // load (Tin, F, V) into the padded map
      const float* src = a.in + (size_t)b * Tin * FV;
      for (int i = tid; i < Tin * FV; i += NT) {
        const int r = i / V, v = i - r * V;
        X[r * FPN_XS + FPN_LPAD + v] = __ldg(src + i);
      }
    }
    __syncthreads();
    for (int l = 0; l < a.n_layers; ++l) {
      const int* f = a.f[l];
      const int Cin = f[CF_CIN];
      const float* cpw = W + f[CF_CP_WT];
      // prefetch the first compress slice while the pooling branch runs
      copy_async<NT>(ring, cpw, To * Top);
      cp_async_commit();
      // global-average branch (:69, 76) folded into a per-output constant
      for (int c = warp; c < Cin; c += NT / 32) {
        float s = 0.f;
        for (int i = tid & 31; i < FV; i += 32) { const int h = i / V, v = i - h * V; s += X[(c * F + h) * FPN_XS + FPN_LPAD + v]; }
        s = warp_sum(s);
        if ((tid & 31) == 0) avg[c] = s / FV;
      }
      __syncthreads();
      for (int o = tid; o < To; o += NT) {
        const float* wt = W + f[CF_CP_AVG_WT] + o;
        float acc = W[f[CF_CP_B] + o];
        for (int c = 0; c < Cin; ++c) acc = fmaf(wt[c * Top], avg[c], acc);
        cst[o] = acc;
      }
      const float oa = W[f[CF_OUT_A]];
      const bool resid = f[CF_RESID] != 0;
#pragma unroll 1
      for (int dil = 1; dil <= 3; ++dil) {
        if (dil == 1) fpn_branch<V, 1, TN, NT>(W + f[CF_W_D1], W + f[CF_B_D1], W[f[CF_A_D1]], X, zero_row, Bb, Cin, To, F);
        else if (dil == 2) fpn_branch<V, 2, TN, NT>(W + f[CF_W_D2], W + f[CF_B_D2], W[f[CF_A_D2]], X, zero_row, Bb, Cin, To, F);
        else fpn_branch<V, 3, TN, NT>(W + f[CF_W_D3], W + f[CF_B_D3], W[f[CF_A_D3]], X, zero_row, Bb, Cin, To, F);
        cp_async_wait_all();
        __syncthreads();                                   // branch output + this slice's weights are visible
        const float* wslice = ring + ((dil - 1) & 1) * rb;
        if (dil < 3) {                                     // next slice streams in during the compress + next branch
          copy_async<NT>(ring + (dil & 1) * rb, cpw + (size_t)dil * To * Top, To * Top);
          cp_async_commit();
        }
        // compress 1x1 (:77-78), accumulated branch by branch; the last slice applies the caller's
        // PReLU (+ residual) and writes the map back in place
        const WideOp ops[1] = {{nullptr, wslice, Bb, nullptr}};
        gemm_wide<4, 4, FVC, FVC, NT, false, 1>(ops, Top, To, To, 0, nullptr, 0,
          [&](int, int m, int n0, float (&v)[4]) {
            float* op = OUT + m * FV + n0;
            if (dil == 1) { store_vec<4>(op, v); return; }
            float s[4];
            lds_vec<4>(op, s);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] += v[j];
            if (dil == 2) { store_vec<4>(op, s); return; }
            const float cm = cst[m];
            int h = n0 / V, vv = n0 - h * V;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float* xp = X + (m * F + h) * FPN_XS + FPN_LPAD + vv;
              float val = prelu(s[j] + cm, oa);
              if (resid) val += *xp;
              *xp = val;
              if (++vv == V) { vv = 0; ++h; }
            }
          });
        __syncthreads();
      }
    }
    // dim_conversor on (F channels, To, V): conv1x1 F->3, BN, PReLU, conv1x1 3->3, PReLU(3)  (:541-545)
    {
      const float* w0 = W + a.t[CT_DC0_WT];
      const float* b0 = W + a.t[CT_DC0_B];
      const float a0 = W[a.t[CT_DC0_A]];
      const float* w3 = W + a.t[CT_DC3_WT];
      const float* a3 = W + a.t[CT_DC3_A];
      for (int i = tid; i < To * V; i += NT) {
        const int fr = i / V, v = i - fr * V;
        float y[3] = {b0[0], b0[1], b0[2]};
        for (int c = 0; c < F; ++c) {
          const float xv = X[(fr * F + c) * FPN_XS + FPN_LPAD + v];
#pragma unroll
          for (int k = 0; k < 3; ++k) y[k] = fmaf(w0[c * 8 + k], xv, y[k]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) y[k] = prelu(y[k], a0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float z = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) z = fmaf(w3[j * 8 + k], y[j], z);
          y6[i * 3 + k] = prelu(z, a3[k]);
        }
      }
    }
    __syncthreads();
    {   // x7 = cumsum over frames (:589)
      float* dst = a.x7 + (size_t)b * To * V * 3;
      for (int i = tid; i < V * 3; i += NT) {
        float s = 0.f;
        for (int fr = 0; fr < To; ++fr) { s += y6[fr * V * 3 + i]; dst[fr * V * 3 + i] = s; }
      }
    }
    __syncthreads();
  }
}

}  // namespace cg
