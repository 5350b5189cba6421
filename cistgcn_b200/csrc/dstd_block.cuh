// Fused DSTD-GC block (reference: models/CISTGCN/CISTGCN.py:273-390, restated in SURVEY.md App. A).
//
// One persistent CTA per SM owns one sample at a time.  The sample's normalised activation tile XN
// (Ci x T*V) stays resident in shared memory for the whole block; every stage (statistics, context
// gates, Map2Adj adjacency generation, TxT / VxV adjacency products, 1x1 channel mixes with folded
// BatchNorm + PReLU + residual, gating, compressor, squeeze-excitation, block residual) reads and
// writes shared memory only.  HBM sees the tile once in and once out.
//
// Weights: the CTA processes hundreds of samples with the same weights, so every operand of a
// GEMM-style loop that fits is copied into shared memory ONCE per launch (host-side plan, dstd_plan());
// what does not fit streams through a RING_SLOTS-deep cp.async ring; single-use operands (gate conv,
// gate MLPs) are read straight from L2 in fixed-trip batches of 8-16 independent loads per lane.
// No inner loop waits on one L2 round trip per iteration.
//
// Shared-memory map (floats):
//   XN  [Ci][TV]          resident input tile (after global_norm)
//   A | B [Cmax][TV]      work tiles: Map2Adj hidden maps -> g1/g2 -> u1/u2 -> c ; B also hosts the
//                         joint-axis expansor's hidden map where that runs as GEMMs
//   ADJ [TV*max(T,V)]     row statistics / split-K partials, then o / Adj_s ([t][q][v]), then o / Adj_t
//   SM                    small vectors (stats, gate activations, dseq/dsp, SE)
//   RING                  RING_SLOTS x ring_floats streaming slots (absent when nothing streams)
//   RES                   resident weights
#pragma once
#include "../../include/cistgcn_b200.h"
#include "simt.h"
#include "umma.cuh"

namespace cg {

#if defined(CISTGCN_EMU) || !defined(CISTGCN_PROFILE)
#define CG_STAMP(i)
#else
#define CG_STAMP(i) do { if (a.phase_clocks && blockIdx.x == 0 && threadIdx.x == 0 && b == (int)blockIdx.x + a.stamp_iter * (int)gridDim.x) a.phase_clocks[i] = CG_CLOCK(); } while (0)
#endif

struct DstdArgs {
  int d[CB_COUNT];       // descriptor; weight fields are float offsets into the blob
  int res[CB_COUNT];     // shared-memory float offset of the resident copy of field f, or -1
  int wsz[CB_COUNT];     // size in floats (multiple of 4) of weight field f, 0 if absent
  const float* w;
  const float* in;
  float* out;
  float* tap_adj_s;
  float* tap_adj_t;
  float* tap_w1;
  float* tap_w2;
  int in_bf16, out_bf16;   // activation tensors stored as bf16 (cistgcn_forward_bf16); strides stay in elements
  int batch;
  int o_xn, o_ab, tile, o_adj, o_sm, o_ring, ring_floats, smem_floats;
  int scratch_floats;    // capacity of the split-K partial scratch (the adjacency region)
  // tensor-core channel mixes (tc_gemm): staging operand [2 terms][tc_kc chunks][256 positions][16 B] + mbarrier / TMEM slot
  int g0_stage, o_gbar;  // gate-conv weights streamed by cp.async.bulk into the (idle) work tiles at the start of every sample
  int tc, o_stage, tc_kc, o_tcmisc, o_img;     // o_img: one weight-image buffer, refilled by cp.async.bulk between uses
  long long* phase_clocks;   // optional debug: first CTA / thread 0 stamps CG_CLOCK() at phase boundaries
  int stamp_iter;            // ... of its stamp_iter-th sample (0 = first: cold caches)
};

__host__ __device__ inline int pad4i(int n) { return (n + 3) & ~3; }
__host__ __device__ inline int pad8i(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }

constexpr int RING_SLOTS = 4;   // cp.async streaming ring: slots in flight
constexpr int KSPLIT_MAX = 8;   // split-K fan-out cap of the narrow GEMMs (also limited by the scratch capacity)

// Host: sizes of every weight field, the shared-memory layout and the residency plan.
// Returns false if even the mandatory small vectors do not fit.
inline bool dstd_plan(DstdArgs& a, int nt, int max_smem_floats, bool tc_allowed = true) {
  const int* d = a.d;
  const int Ci = d[CB_CI], Co = d[CB_CO], T = d[CB_T], V = d[CB_V], Ch = d[CB_CH], Cg = d[CB_CG], Hs = d[CB_HS];
  const bool has_res = d[CB_HAS_RES] != 0, interp = d[CB_INTERP] != 0;
  const int TV = T * V, cmax = imax(Ci, Co), big = imax(TV * imax(T, V), T * pad4i(V * V)), Cop = pad8i(Co);
  for (int f = 0; f < CB_COUNT; ++f) { a.wsz[f] = 0; a.res[f] = -1; }
  int* z = a.wsz;
  z[CB_GN_S] = z[CB_GN_B] = Ci;
  z[CB_G0_WT] = Ci * T * pad8i(2 * Cg); z[CB_G0_B] = 2 * Cg; z[CB_G0_A] = 2;
  z[CB_G4_WT] = 2 * Cg * V * Cop; z[CB_G4_B] = 2 * Co; z[CB_G4_A] = 2;
  z[CB_M0_WT] = 2 * (Co + 2 + 2 * T) * Cop; z[CB_M0_B] = 2 * Co; z[CB_M0_A] = 2;
  z[CB_M4_WT] = 2 * Co * Cop;
  if (interp) {
    z[CB_A0_WT] = Ci * pad8i(4 * Ch); z[CB_A0_B] = 4 * Ch; z[CB_A0_A] = 4;
    for (int L = 0; L < 2; ++L) {
      const int n = L == 0 ? V : T;
      z[CB_TC3_WT_S + L] = Ch * T * pad8i(Ch); z[CB_TC3_B_S + L] = Ch; z[CB_TC6_WT_S + L] = Ch * pad8i(T);
      z[CB_JC3_WT_S + L] = Ch * V * pad8i(Ch); z[CB_JC3_B_S + L] = Ch; z[CB_JC6_WT_S + L] = Ch * pad8i(V);
      z[CB_E0_WT_S + L] = n * pad8i(n); z[CB_E0_B_S + L] = n; z[CB_E0_A_S + L] = 1; z[CB_E4_WT_S + L] = n * pad8i(n);
    }
  }
  for (int L = 0; L < 2; ++L) {
    z[CB_TCN_WT_S + L] = Ci * (has_res ? 2 : 1) * Cop; z[CB_TCN_B_S + L] = Co; z[CB_TCN_A_S + L] = 1;
    z[CB_P_S_S + L] = Co; z[CB_P_B_S + L] = Co; z[CB_P_A_S + L] = 1;
  }
  z[CB_CP_WT] = 2 * Co * Cop; z[CB_CP_B] = Co; z[CB_CP_A] = 1;
  z[CB_SE1_WT] = Co * pad8i(Hs); z[CB_SE2_WT] = Hs * Cop;
  if (has_res) { z[CB_RS_WT] = Ci * Cop; z[CB_RS_B] = Co; }
  for (int f = 0; f < CB_COUNT; ++f) z[f] = pad4i(z[f]);

  a.tile = pad4i(cmax * TV);
  a.o_xn = 0;
  a.o_ab = pad4i(Ci * TV);
  // B also hosts the hidden map [V][T*T] of the joint-axis expansor where that runs as two GEMMs (few columns)
  const int hs_need = (T * T * 2 <= nt) ? pad4i(T * T * V) : 0;
  a.o_adj = a.o_ab + a.tile + imax(a.tile, hs_need);
  const int nw = nt / 32;
  // split-K partial sums of the narrow GEMMs (<= KSPLIT_MAX copies of an M x N output) and of the matvecs
  // (nw/2 copies of 2 x Co) share the adjacency region with the row statistics
  const int scratch = imax(4 * imax(2 * Cg * V, imax(Ch * V, Ch * T)), nw * Co);
  const int adj0 = imax(imax(pad4i(big), pad4i(T * T * (V | 1))), imax(pad4i(2 * Ci * T + 2 * Ci), pad4i(scratch)));
  int adj = adj0;
  a.o_sm = a.o_adj + adj;
  a.scratch_floats = adj;
  const int sm = pad4i(2 + 2 * T) + imax(pad4i(2 * Cg * V), pad4i(2 * Ch * V) + pad4i(2 * Ch * T)) + 3 * pad4i(2 * Co) +
                 2 * pad4i(2 * TV) + 2 * pad4i(Co) + pad4i(Hs) + 4;       // + one mbarrier (g0_stage)
  a.o_ring = a.o_sm + sm;
  const int widest = imax(pad8i(4 * Ch), imax(Cop, pad8i(2 * Cg)));      // longest streamed row
  const int vectors[] = {CB_GN_S, CB_GN_B, CB_G0_B, CB_G0_A, CB_G4_B, CB_G4_A, CB_M0_B, CB_M0_A, CB_A0_B, CB_A0_A,
                         CB_TC3_B_S, CB_TC3_B_T, CB_JC3_B_S, CB_JC3_B_T, CB_E0_B_S, CB_E0_B_T, CB_E0_A_S, CB_E0_A_T,
                         CB_TCN_B_S, CB_TCN_B_T, CB_TCN_A_S, CB_TCN_A_T, CB_P_S_S, CB_P_S_T, CB_P_B_S, CB_P_B_T,
                         CB_P_A_S, CB_P_A_T, CB_CP_B, CB_CP_A, CB_RS_B};
  // operands of GEMM-style loops (must sit in shared memory: resident, else streamed through the ring), by benefit
  const int gemm_ops[] = {CB_SE1_WT, CB_SE2_WT, CB_A0_WT, CB_TCN_WT_S, CB_TCN_WT_T, CB_CP_WT, CB_RS_WT, CB_E0_WT_S, CB_E0_WT_T,
                          CB_E4_WT_S, CB_E4_WT_T, CB_TC6_WT_S, CB_TC6_WT_T, CB_JC6_WT_S, CB_JC6_WT_T,
                          CB_TC3_WT_S, CB_TC3_WT_T, CB_JC3_WT_S, CB_JC3_WT_T};
  // operands read straight from L2 with deep load batches when not resident (gate conv, gate MLPs)
  const int l2_ops[] = {CB_M4_WT, CB_M0_WT, CB_G0_WT, CB_G4_WT};
  // tensor-core images of the channel mixes (512-thread CTAs only: one CTA per SM owns the SM's tensor memory)
  const int tc_fields[] = {CB_TC_A0, CB_TC_TCN_S, CB_TC_TCN_T, CB_TC_CP, CB_TC_RS};
  const int kc_a0 = ((Ci + 15) / 16) * 2;                        // k-chunks of one row block of Ci / Co channels
  const int kc_co = ((Co + 15) / 16) * 2;
  const int kc_tcn = has_res ? 2 * kc_a0 : kc_a0;
  const int kc_cp = 2 * kc_co;
  const int np_a0 = (4 * Ch + 15) / 16 * 16, np_co = (Co + 15) / 16 * 16;
  bool want_tc = false;
#ifndef CISTGCN_EMU
  want_tc = nt == 512 && tc_allowed && d[CB_TC_CP] > 0 && d[CB_TC_TCN_S] > 0 && d[CB_TC_TCN_T] > 0 && TV <= 256 &&
            (!interp || d[CB_TC_A0] > 0) && (!has_res || d[CB_TC_RS] > 0) && np_a0 <= 64 && np_co <= 64;
#endif
  bool ok = true;
  int cur = 0;
  for (int try_tc = want_tc ? 1 : 0; try_tc >= 0; --try_tc) {
    for (int f : tc_fields) z[f] = 0;
    a.tc = try_tc;
    a.tc_kc = 0;
    if (try_tc) {
      if (interp) z[CB_TC_A0] = kc_a0 * 4 * np_a0 * 4;
      z[CB_TC_TCN_S] = z[CB_TC_TCN_T] = kc_tcn * 4 * np_co * 4;
      z[CB_TC_CP] = kc_cp * 4 * np_co * 4;
      if (has_res) z[CB_TC_RS] = kc_a0 * 4 * np_co * 4;
      a.tc_kc = imax(kc_a0, kc_co);                                // the staging operand holds one row block at a time
    }
    // the staging operand aliases the adjacency region (dead during every channel mix), which grows to hold it
    adj = try_tc ? imax(adj0, 2 * a.tc_kc * 256 * 4) : adj0;
    a.o_stage = a.o_adj;
    a.o_sm = a.o_adj + adj;
    a.scratch_floats = adj;
    a.o_ring = a.o_sm + sm;
  for (int with_ring = 0; with_ring < 2; ++with_ring) {          // first try: everything GEMM-side resident, no ring
    for (int f = 0; f < CB_COUNT; ++f) a.res[f] = -1;
    ok = true;
    const int budget = max_smem_floats - a.o_ring;
    a.ring_floats = with_ring ? imax(widest, budget >= 24576 ? 2048 : (budget >= 8192 ? 1024 : 512)) : 0;
    cur = a.o_ring + RING_SLOTS * a.ring_floats;
    if (try_tc) {
      a.o_tcmisc = cur; cur += 8;
      cur = (cur + 31) & ~31;
      int img = 0;
      for (int f : tc_fields) img = imax(img, a.wsz[f]);
      a.o_img = cur; cur += img;                                   // images stream from L2 into this one buffer
    }
    auto take = [&](int f, bool mandatory) {
      if (a.wsz[f] == 0 || a.res[f] >= 0) return true;
      if (try_tc && (f == CB_A0_WT || f == CB_TCN_WT_S || f == CB_TCN_WT_T || f == CB_CP_WT || f == CB_RS_WT)) return true;
      if (cur + a.wsz[f] <= max_smem_floats) { a.res[f] = cur; cur += a.wsz[f]; return true; }
      if (mandatory) ok = false;
      return false;
    };
    for (int f : vectors) take(f, true);
    bool all_gemm = true;
    for (int f : gemm_ops) all_gemm = take(f, false) && all_gemm;
    if (!with_ring && !all_gemm) continue;                       // something must stream: plan again with a ring
    for (int f : l2_ops) take(f, false);
    break;
  }
    if (ok && cur <= max_smem_floats) break;                     // else: plan again without the tensor-core path
  }
  if (a.tc && cur < 30 * 1024) cur = 30 * 1024;                  // > half an SM: the CTA must own all 512 TMEM columns
  a.o_gbar = a.o_ring - 4;
  a.g0_stage = 0;
#ifndef CISTGCN_EMU
  // gate conv (T,1) weights that found no room: the two work tiles are idle until Map2Adj starts, so the weights
  // are bulk-copied into them at the top of every sample and the conv runs the resident-operand path
  a.g0_stage = a.res[CB_G0_WT] < 0 && d[CB_IN_MODE] != 1 && a.wsz[CB_G0_WT] <= a.o_adj - a.o_ab;
#endif
  a.smem_floats = cur;
  return ok && cur <= max_smem_floats;
}

// True when the plan keeps every GEMM-style operand resident (no ring needed).
inline bool dstd_all_gemm_resident(const DstdArgs& a) { return a.ring_floats == 0; }

// ---------------------------------------------------------------------------------------------
// Weight delivery: body(wc, k0, kc) is called by every thread for consecutive row chunks
// [k0, k0+kc) of a k-major matrix (row length Mp) with the rows available in shared memory at wc.
// Resident matrices are one chunk with no barrier; streamed ones go through the cp.async ring
// (chunk c+1 is in flight while chunk c is consumed; one barrier per chunk).
// ---------------------------------------------------------------------------------------------

CG_DEV void cp_async_wait_pending(int n) {      // wait until at most n of this thread's groups are in flight
#ifndef CISTGCN_EMU
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;\n" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;\n" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 6;\n" ::: "memory"); break;
  }
#else
  (void)n;
#endif
}

template <int NT, int SLOTS = RING_SLOTS, class BODY>
CG_DEV void for_weight_chunks(const float* __restrict__ g, const float* s, int K, int Mp, float* ring, int rb, BODY body) {
  if (s != nullptr) { body(s, 0, K); return; }
  // ring = SLOTS slots of rb floats; up to SLOTS-1 chunks are in flight ahead of the consumer
  const int kc = rb / Mp;
  const int nch = (K + kc - 1) / kc;
  int issued = 0;
  auto issue = [&]() {
    copy_async<NT>(ring + (issued % SLOTS) * rb, g + (size_t)issued * kc * Mp, imin(kc, K - issued * kc) * Mp);
    cp_async_commit();
    ++issued;
  };
  for (int i = 0; i < imin(nch, SLOTS - 1); ++i) issue();
  for (int c = 0; c < nch; ++c) {
    cp_async_wait_pending(issued - c - 1);
    __syncthreads();                 // chunk c visible to all; everyone is done with chunk c-1, whose slot is reused now
    if (issued < nch) issue();
    body(ring + (c % SLOTS) * rb, c * kc, imin(kc, K - c * kc));
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Wide GEMM: out(m, n) = sum_k W[k][m] * X[k][n], X = up to two stacked row blocks (row stride LD) in
// shared memory, N (a multiple of TN) columns; NP independent problems of the same shape may share
// the phase.  Work unit = one warp x TM rows x 32 lanes x TN contiguous columns: per k one vector LDS
// of activations (conflict-free) + TM/4 broadcast LDS.128 of weights feed TM*TN FFMAs; the k loop
// is unrolled 4x (measured: the compiler's own pipelining beats manual register double-buffering).
// The epilogue gets whole rows: epi(problem, m, n0, float (&v)[TN]).
// INPLACE: the epilogue may overwrite X; a pass then holds whole column groups and every thread meets
// the two barriers of the pass (needs ceil(M/TM) <= NT/32).
// ---------------------------------------------------------------------------------------------
struct WideOp {
  const float* wg;   // weights in the global blob
  const float* ws;   // resident copy in shared memory or nullptr (-> streamed through the ring)
  const float* X1;   // first K1 rows of activations
  const float* X2;   // next K2 rows (or nullptr)
};

template <int TM, int TN, int LD, int N, int NT, bool INPLACE, int NP, int SLOTS = RING_SLOTS, class EPI>
CG_DEV void gemm_wide(const WideOp (&ops)[NP], int Mp, int M, int K1, int K2, float* ring, int rb, EPI epi) {
  static_assert(N % TN == 0 && LD % TN == 0, "gemm_wide: columns must tile by TN");
  constexpr int NW = NT / 32;
  constexpr int NCOLS = N / TN;
  constexpr int NG = (NCOLS + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtiles = (M + TM - 1) / TM;
  const int total = mtiles * NG * NP;
  const int per_pass = INPLACE ? (NW / mtiles) * mtiles : NW;
  for (int base = 0; base < total; base += per_pass) {
    const int item = base + warp;
    const int unit = item / mtiles;                  // (column group, problem)
    const int prob = NP == 1 ? 0 : unit % NP;
    const int slot = (unit / NP) * 32 + lane;
    const bool active = warp < per_pass && item < total && slot < NCOLS;
    const int m0 = active ? (item % mtiles) * TM : 0;
    const int n0 = active ? slot * TN : 0;
    const WideOp op = (NP > 1 && active && prob == 1) ? ops[NP - 1] : ops[0];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    auto fma_step = [&](const float (&w)[TM], const float (&x)[TN]) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
    };
    auto run = [&](const float* xp, const float* wp, int n) {
#pragma unroll 4
      for (int kk = 0; kk < n; ++kk) {
        float w[TM], x[TN];
        lds_vec<TM>(wp, w);
        lds_vec<TN>(xp, x);
        wp += Mp;
        xp += LD;
        fma_step(w, x);
      }
    };
    auto body = [&](const float* wc, int k0, int kc) {
      if (!active) return;
      int kb = k0, ke = imin(k0 + kc, K1);
      if (kb < ke) run(op.X1 + kb * LD + n0, wc + (kb - k0) * Mp + m0, ke - kb);
      kb = imax(k0, K1); ke = k0 + kc;
      if (kb < ke) run(op.X2 + (kb - K1) * LD + n0, wc + (kb - k0) * Mp + m0, ke - kb);
    };
    if constexpr (NP == 1) for_weight_chunks<NT, SLOTS>(ops[0].wg, ops[0].ws, K1 + K2, Mp, ring, rb, body);
    else body(op.ws, 0, K1 + K2);                    // multi-problem phases need resident weights
    if (INPLACE) __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
        if (m0 + i < M) epi(prob, m0 + i, n0, acc[i]);
    }
    if (INPLACE) __syncthreads();
  }
}

// Row-tile dispatch: the widest TM that still gives every warp a work unit.
template <int TN, int LD, int N, int NT, bool INPLACE, int NP, class EPI>
CG_DEV void gemm_wide_auto(const WideOp (&ops)[NP], int Mp, int M, int K1, int K2, float* ring, int rb, EPI epi) {
  constexpr int NW = NT / 32;
  constexpr int NG = (((N / TN) + 31) / 32) * NP;
  if (M % 16 == 0 && (M / 16) * NG >= NW)
    gemm_wide<16, TN, LD, N, NT, INPLACE, NP>(ops, Mp, M, K1, K2, ring, rb, epi);
  else if (M > 32 || ((M + 7) / 8) * NG >= NW)
    gemm_wide<8, TN, LD, N, NT, INPLACE, NP>(ops, Mp, M, K1, K2, ring, rb, epi);
  else if (M > 16 || ((M + 3) / 4) * NG >= NW)
    gemm_wide<4, TN, LD, N, NT, INPLACE, NP>(ops, Mp, M, K1, K2, ring, rb, epi);
  else
    gemm_wide<2, TN, LD, N, NT, INPLACE, NP>(ops, Mp, M, K1, K2, ring, rb, epi);
}

// Vector store of TN contiguous floats to shared / global memory (p aligned to TN floats).
template <int TN>
CG_DEV void store_vec(float* p, const float (&v)[TN]) {
  if constexpr (TN == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (TN == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  else {
#pragma unroll
    for (int j = 0; j < TN; ++j) p[j] = v[j];
  }
}

// ---------------------------------------------------------------------------------------------
// Narrow GEMM (the collapsing (T,1) / (1,V) convolutions: X is a [K][N] view of a [C][T*V] tile,
// N = V or T <= 32):  out(m, n) = sum_k W[k][m] * X[k*N + n].
// Lane -> (row sub-tile, TN-wide column slot); the K range is split over the warps that are left
// once every row tile has a warp, partial sums meet in `partial` (>= NW*TM*64 floats).
// Ends with a barrier; epi runs once per output.
// ---------------------------------------------------------------------------------------------
template <int TM, int N, bool STRIDED, int NT, class EPI>
CG_DEV void gemm_narrow(const float* __restrict__ wg_, const float* ws, int Mp, int M, int K,
                        const float* X, int R, int CS, float* partial, int partial_cap, float* ring, int rb, EPI epi) {
  // X(k, n): contiguous mode  X[k*N + n]  (the (T,1) conv over a [c][t][v] tile, N = V);
  //          strided mode     X[(k / R)*CS + (k % R) + n*R]  (the (1,V) conv over the same layout: k = (c, v), n = t, R = V)
  static_assert(N <= 32, "gemm_narrow: row length must fit a warp");
  constexpr int NW = NT / 32;
  constexpr int TN = (!STRIDED && N % 2 == 0) ? 2 : 1;
  constexpr int NP = N / TN;
  constexpr int MS = 32 / NP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int msub = lane / NP, np = lane - msub * NP;
  const int mtiles = (M + TM - 1) / TM;
  const int mgroups = (mtiles + MS - 1) / MS;
  const int ksplit = imax(1, imin(imin(KSPLIT_MAX, NW / mgroups), partial_cap / (M * N)));
  for (int gbase = 0; gbase < mgroups; gbase += NW) {          // (one pass unless mgroups > NW)
    const int mg = gbase + warp % imin(mgroups, NW);
    const int ks = warp / imin(mgroups, NW);
    const int mt = mg * MS + msub;
    const bool active = ks < ksplit && mg < mgroups && lane < MS * NP && mt < mtiles;
    const int m0 = active ? mt * TM : 0;
    const int n0 = active ? np * TN : 0;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for_weight_chunks<NT>(wg_, ws, K, Mp, ring, rb, [&](const float* wc, int k0, int kc) {
      if (!active) return;
      const int r0 = (kc * ks) / ksplit, r1 = (kc * (ks + 1)) / ksplit;
      const float* wp = wc + r0 * Mp + m0;
      if constexpr (!STRIDED) {
        const float* xp = X + (k0 + r0) * N + n0;
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
          float w[TM], x[TN];
          lds_vec<TM>(wp, w);
          lds_vec<TN>(xp, x);
          wp += Mp;
          xp += N;
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
        }
      } else {
        int cc = (k0 + r0) / R, rr = (k0 + r0) - cc * R;
        const float* xp = X + cc * CS + rr + n0 * R;
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
          float w[TM];
          lds_vec<TM>(wp, w);
          const float x = *xp;
          wp += Mp;
          ++xp;
          if (++rr == R) { rr = 0; xp += CS - R; }
#pragma unroll
          for (int i = 0; i < TM; ++i) acc[i][0] = fmaf(w[i], x, acc[i][0]);
        }
      }
    });
    if (ksplit == 1) {
      if (active) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
          if (m0 + i < M) {
#pragma unroll
            for (int j = 0; j < TN; ++j) epi(m0 + i, n0 + j, acc[i][j]);
          }
      }
      __syncthreads();
    } else {
      if (active) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
          if (m0 + i < M) {
#pragma unroll
            for (int j = 0; j < TN; ++j) partial[(ks * M + m0 + i) * N + n0 + j] = acc[i][j];
          }
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < M * N; idx += NT) {
        float s = 0.f;
        for (int q = 0; q < ksplit; ++q) s += partial[q * M * N + idx];
        epi(idx / N, idx % N, s);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The four collapsing convolutions of Map2Adj (time_compress.3 and joint_compress.3 of both domains, CISTGCN.py:141-150)
// in ONE phase: the warps are split over the four problems and the K range of a problem over its warps, so all warps
// work at once and the phase costs one barrier and one reduction instead of four of each.
//   problem 2L   : out(m, v) = sum_{c,t} W[(c,t)][m] * map_tc[L][c][t][v]      (N = V, contiguous rows)
//   problem 2L+1 : out(m, t) = sum_{c,v} W[(c,v)][m] * map_jc[L][c][t][v]      (N = T, strided rows)
// Needs resident weights, ceil(Ch/8) row tiles per warp (Ch <= 8 * lanes-per-row-tile), partial >= (NT/128) * 2*Ch*(V+T).
// ---------------------------------------------------------------------------------------------
template <int T, int V>
struct Collapse4 {
  static constexpr int TNA = (V % 2 == 0) ? 2 : 1, NPA = V / TNA, MSA = 32 / NPA;     // contiguous problems: lane -> (row tile, column pair)
  static constexpr int NPB = T, MSB = 32 / NPB;                                         // strided problems: lane -> (row tile, column)
  static_assert(V <= 32 * TNA && T <= 32, "Collapse4: a row must fit a warp");
  CG_DEV static bool fits(int Ch, int nt, int partial_cap) {
    const int mt = (Ch + 7) / 8;
    return (nt / 32) % 4 == 0 && mt <= MSA && mt <= MSB && partial_cap >= (nt / 128) * 2 * Ch * (V + T);
  }
};

template <int T, int V, int NT, class EPI>
CG_DEV void collapse4(const float* w_tc0, const float* w_jc0, const float* w_tc1, const float* w_jc1, int Ch,
                      const float* maps, int tile, float* partial, EPI epi) {
  using C4 = Collapse4<T, V>;
  constexpr int TV = T * V, KSW = NT / 128;                  // warps per problem = split-K fan-out
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = warp / KSW, ks = warp - prob * KSW, L = prob >> 1;
  const int Mp = pad8i(Ch), mtiles = (Ch + 7) / 8;
  const int szA = Ch * V, szB = Ch * T;                      // outputs of a contiguous / strided problem
  const int pbase = KSW * ((prob >> 1) * (szA + szB) + (prob & 1) * szA);
  const float* tl = maps + L * tile;
  float acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
  if ((prob & 1) == 0) {
    const float* wc = L ? w_tc1 : w_tc0;
    const int msub = lane / C4::NPA, np = lane - msub * C4::NPA;
    const bool active = lane < C4::MSA * C4::NPA && msub < mtiles;
    const int m0 = active ? msub * 8 : 0, n0 = active ? np * C4::TNA : 0;
    const int K = Ch * T, r0 = (K * ks) / KSW, r1 = (K * (ks + 1)) / KSW;
    if (active) {
      const float* wp = wc + r0 * Mp + m0;
      const float* xp = tl + r0 * V + n0;
#pragma unroll 4
      for (int r = r0; r < r1; ++r) {
        float w[8], x[C4::TNA];
        lds_vec<8>(wp, w);
        lds_vec<C4::TNA>(xp, x);
        wp += Mp;
        xp += V;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < C4::TNA; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (m0 + i < Ch) {
#pragma unroll
          for (int j = 0; j < C4::TNA; ++j) partial[pbase + (ks * Ch + m0 + i) * V + n0 + j] = acc[i][j];
        }
    }
  } else {
    const float* wc = L ? w_jc1 : w_jc0;
    const float* X = tl + Ch * TV;
    const int msub = lane / C4::NPB, np = lane - msub * C4::NPB;
    const bool active = lane < C4::MSB * C4::NPB && msub < mtiles;
    const int m0 = active ? msub * 8 : 0, n0 = active ? np : 0;
    const int K = Ch * V, r0 = (K * ks) / KSW, r1 = (K * (ks + 1)) / KSW;
    if (active) {
      const float* wp = wc + r0 * Mp + m0;
      int cc = r0 / V, rr = r0 - cc * V;
      const float* xp = X + cc * TV + rr + n0 * V;
#pragma unroll 4
      for (int r = r0; r < r1; ++r) {
        float w[8];
        lds_vec<8>(wp, w);
        const float x = *xp;
        wp += Mp;
        ++xp;
        if (++rr == V) { rr = 0; xp += TV - V; }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = fmaf(w[i], x, acc[i][0]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (m0 + i < Ch) partial[pbase + (ks * Ch + m0 + i) * T + n0] = acc[i][0];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * (szA + szB); idx += NT) {
    const int Lq = idx / (szA + szB), r = idx - Lq * (szA + szB);
    const bool strided = r >= szA;
    const int o = strided ? r - szA : r, sz = strided ? szB : szA;
    const float* pp = partial + KSW * (Lq * (szA + szB) + (strided ? szA : 0)) + o;
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < KSW; ++q) sum += pp[q * sz];
    const int n = strided ? T : V;
    epi(2 * Lq + (strided ? 1 : 0), o / n, o % n, sum);
  }
  __syncthreads();
}

template <int N, bool STRIDED, int NT, class EPI>
CG_DEV void gemm_narrow_auto(const float* __restrict__ wg_, const float* ws, int Mp, int M, int K,
                             const float* X, int R, int CS, float* partial, int partial_cap, float* ring, int rb, EPI epi) {
  if (M >= 16) gemm_narrow<8, N, STRIDED, NT>(wg_, ws, Mp, M, K, X, R, CS, partial, partial_cap, ring, rb, epi);
  else if (M >= 8) gemm_narrow<4, N, STRIDED, NT>(wg_, ws, Mp, M, K, X, R, CS, partial, partial_cap, ring, rb, epi);
  else gemm_narrow<2, N, STRIDED, NT>(wg_, ws, Mp, M, K, X, R, CS, partial, partial_cap, ring, rb, epi);
}

// ---------------------------------------------------------------------------------------------
// Gate conv (T,1) with its weights read straight from L2 (used when the matrix is not resident):
//   out(m, v) = sum_k W[k][m] * X[k*V + v].
// Lanes own output rows (one coalesced 128-byte weight row per k), warps split K, every lane keeps V
// accumulators; weight loads go out in fixed-trip batches of 8 independent requests (no rolled remainder),
// activations are warp-uniform shared-memory broadcasts.  partial: >= ksplit*M*V floats.  Ends with a barrier.
// ---------------------------------------------------------------------------------------------
template <int V, int NT, class EPI>
CG_DEV void gate_conv_l2(const float* __restrict__ Wg, int Mp, int M, int K, const float* X,
                         float* partial, int partial_cap, EPI epi) {
  constexpr int NW = NT / 32, BATCH = 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksplit = imax(1, imin(NW, partial_cap / (M * V)));
  const int r0 = (K * warp) / ksplit, r1 = warp < ksplit ? (K * (warp + 1)) / ksplit : r0;
  for (int mb = 0; mb < M; mb += 32) {
    const int m = mb + lane;
    const int mc = m < M ? m : M - 1;
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    for (int k0 = r0; k0 < r1; k0 += BATCH) {
      float w[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j) w[j] = __ldg(Wg + (size_t)imin(k0 + j, r1 - 1) * Mp + mc);
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        if (k0 + j < r1) {
          const float* xr = X + (k0 + j) * V;
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = fmaf(w[j], xr[v], acc[v]);
        }
      }
    }
    if (m < M && warp < ksplit) {
#pragma unroll
      for (int v = 0; v < V; ++v) partial[(warp * M + m) * V + v] = acc[v];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < M * V; idx += NT) {
    float s0 = 0.f, s1 = 0.f;
    int q = 0;
    for (; q + 2 <= ksplit; q += 2) { s0 += partial[q * M * V + idx]; s1 += partial[(q + 1) * M * V + idx]; }
    if (q < ksplit) s0 += partial[q * M * V + idx];
    epi(idx / V, idx % V, s0 + s1);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Two-gate matvec with split-K over warps: y[g][m] = sum_k W_g[k][m] * x_g[k] (+ optional tail rows
// multiplying a second vector shared by both gates).  Weights come straight from L2 (or from their
// resident copy): every lane owns output columns (coalesced 128-byte rows) and keeps UNR independent
// loads in flight, so the single-use weight matrix streams at bandwidth instead of latency.
// partial: >= NW * Mtot floats.  Ends with a barrier.
// ---------------------------------------------------------------------------------------------
template <int NT, class EPI>
CG_DEV void gate_matvec(const float* wmat, int Mp, int M, int K1, const float* x1, int x1_stride,
                        int K2, const float* x2, float* partial, EPI epi) {
  constexpr int NW = NT / 32, HW = NW / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = warp / HW, ks = warp % HW;
  const int K = K1 + K2;
  const int r0 = (K * ks) / HW, r1 = (K * (ks + 1)) / HW;
  const float* wgt = wmat + (size_t)g * K * Mp;
  const float* xa = x1 + g * x1_stride;
  for (int mb = 0; mb < M; mb += 32) {
    const int m = mb + lane;
    const int mc = m < M ? m : M - 1;
    float acc = 0.f;
    const int e1 = imin(r1, K1);
    // fixed-trip batches of 16 independent loads (rows past the end are clamped and multiplied by 0): a
    // rolled remainder loop would pay one L2 round trip per leftover row
    for (int k0 = r0; k0 < e1; k0 += 16) {
      float w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = wgt[(size_t)imin(k0 + j, e1 - 1) * Mp + mc];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(w[j], (k0 + j < e1) ? xa[k0 + j] : 0.f, acc);
    }
    for (int k0 = imax(r0, K1); k0 < r1; k0 += 8) {
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = wgt[(size_t)imin(k0 + j, r1 - 1) * Mp + mc];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(w[j], (k0 + j < r1) ? x2[k0 + j - K1] : 0.f, acc);
    }
    if (m < M) partial[(ks * 2 + g) * M + m] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * M; idx += NT) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < HW; ++q) s += partial[q * 2 * M + idx];
    epi(idx / M, idx % M, s);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Last two layers of a gate MLP (map_{s,t}.0 + BN + PReLU, map_{s,t}.4; CISTGCN.py:341-352) for ONE gate by ONE warp:
// both are small (Co x (Co + 2 + 2T) and Co x Co), nothing else needs their result before the tcn epilogue, so warps
// 0 / 1 run them with warp-level synchronisation only while the rest of the CTA moves on to Map2Adj.  Lanes own output
// columns (coalesced weight rows, fixed-trip batches of 16 independent loads as in gate_matvec).
// ---------------------------------------------------------------------------------------------
CG_DEV float warp_matvec_col(const float* wgt, int Mp, int mc, int K1, const float* xa, int K2, const float* xb) {
  float acc = 0.f;
  for (int k0 = 0; k0 < K1; k0 += 16) {
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = wgt[(size_t)imin(k0 + j, K1 - 1) * Mp + mc];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaf(w[j], (k0 + j < K1) ? xa[k0 + j] : 0.f, acc);
  }
  for (int k0 = 0; k0 < K2; k0 += 8) {
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = wgt[(size_t)(K1 + imin(k0 + j, K2 - 1)) * Mp + mc];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(w[j], (k0 + j < K2) ? xb[k0 + j] : 0.f, acc);
  }
  return acc;
}

template <class EPI>
CG_DEV void gate_mlp_tail(int g, const float* m0w, const float* m0b, float m0a, const float* m4w, int Mp, int Co, int nstats,
                          const float* h2g, const float* stats, float* zgg, EPI epi) {
  const int lane = threadIdx.x & 31;
  const float* w0 = m0w + (size_t)g * (Co + nstats) * Mp;
  for (int mb = 0; mb < Co; mb += 32) {
    const int m = mb + lane, mc = m < Co ? m : Co - 1;
    const float acc = warp_matvec_col(w0, Mp, mc, Co, h2g, nstats, stats);
    if (m < Co) zgg[m] = prelu(acc + m0b[g * Co + m], m0a);
  }
  __syncwarp();
  const float* w4 = m4w + (size_t)g * Co * Mp;
  for (int mb = 0; mb < Co; mb += 32) {
    const int m = mb + lane, mc = m < Co ? m : Co - 1;
    const float acc = warp_matvec_col(w4, Mp, mc, Co, zgg, 0, nullptr);
    if (m < Co) epi(m, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// adjacency products (ConvTemporalGraphical, CISTGCN.py:110,117,123)
// ---------------------------------------------------------------------------------------------
// "space" domain: g1[c][q][v] = sum_t XN[c][t][v] * Adj_s[v][t][q], Adj_s held as adjT[(t*T+q)*VP + v], VP = V|1.
template <int T, int V, int TC, int NT, int LD = T * V>
CG_DEV void gcn_space(const float* XN, const float* adjT, float* G, int C) {
  constexpr int TV = LD, VP = V | 1;                       // TV: row stride of the XN / G tiles
  const int nct = (C + TC - 1) / TC;
  for (int item = threadIdx.x; item < nct * V; item += NT) {
    const int v = item % V, c0 = (item / V) * TC;
    float acc[TC][T];
#pragma unroll
    for (int i = 0; i < TC; ++i)
#pragma unroll
      for (int q = 0; q < T; ++q) acc[i][q] = 0.f;
#pragma unroll 2
    for (int t = 0; t < T; ++t) {
      float xv[TC];
#pragma unroll
      for (int i = 0; i < TC; ++i) xv[i] = (c0 + i < C) ? XN[(c0 + i) * TV + t * V + v] : 0.f;
#pragma unroll
      for (int q = 0; q < T; ++q) {
        const float aq = adjT[(t * T + q) * VP + v];
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
template <int T, int V, int TC, int NT, int LD = T * V>
CG_DEV void gcn_time(const float* XN, const float* adj, float* G, int C) {
  constexpr int TV = LD, VVP = (V * V + 3) & ~3;           // TV: row stride of the XN / G tiles; VVP: padded row stride of the frame-axis adjacency
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
#pragma unroll 2
    for (int v = 0; v < V; ++v) {
      float xv[TC];
#pragma unroll
      for (int i = 0; i < TC; ++i) xv[i] = (c0 + i < C) ? XN[(c0 + i) * TV + t * V + v] : 0.f;
      const float* ap = adj + t * VVP + v * V + w0;
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
// Adjacency expansor (Map2Adj.expansor, CISTGCN.py:165-170) as a per-column two-layer MLP.  The two N x N
// convolutions mix the leading axis of the outer-product map o and every column of o is independent:
//   out[:, c] = W4^T * PReLU(W0^T * o[:, c] + b0)        (weights k-major [k][pad8(N)], N = V or T <= 32)
// so neither o nor the hidden map is materialised: a thread (or G adjacent lanes sharing a column: lane g owns the
// output rows [g*R, (g+1)*R)) forms o[k, c] on the fly, keeps its hidden rows in registers and writes only the result.
// With G > 1 the hidden column goes through `hs` ([NCOLS][N] floats) and a warp barrier.
// ---------------------------------------------------------------------------------------------
// NW floats (a multiple of 4) from a 16-byte aligned row in shared or global memory
template <int NW>
CG_DEV void load_row(const float* p, float (&w)[NW]) {
#pragma unroll
  for (int i = 0; i < NW / 4; ++i) {
    const float4 v = *(reinterpret_cast<const float4*>(p) + i);
    w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
  }
}

template <int N, int NCOLS, int NT, int G, class OFN, class STORE>
CG_DEV void expansor_fused(const float* __restrict__ w0, const float* __restrict__ b0, float a0,
                           const float* __restrict__ w4, float* hs, OFN o_at, STORE store) {
  static_assert(G == 1 || G == 2 || G == 4, "expansor_fused: lanes per column");
  constexpr int NPW = (N + 7) & ~7;                 // row stride of the k-major weights
  constexpr int R = (N + G - 1) / G;                // output rows per lane
  const int lane = threadIdx.x & 31;
  for (int it0 = threadIdx.x - lane; it0 < NCOLS * G; it0 += NT) {
    const int item = it0 + lane;
    const bool active = item < NCOLS * G;
    const int col = active ? item / G : 0, g = item % G;
    const int m0 = g * R;
    float h[R];
#pragma unroll
    for (int j = 0; j < R; ++j) h[j] = 0.f;
    if (active) {
#pragma unroll 2
      for (int k = 0; k < N; ++k) {
        const float ok = o_at(k, col);
        if constexpr (G == 1) {                    // whole weight row: warp-uniform 128-bit loads (rows are 32-byte aligned)
          float wrow[NPW];
          load_row<NPW>(w0 + k * NPW, wrow);
#pragma unroll
          for (int j = 0; j < N; ++j) h[j] = fmaf(wrow[j], ok, h[j]);
        } else {
          const float* wr = w0 + k * NPW + m0;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (m0 + j < N) h[j] = fmaf(wr[j], ok, h[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (m0 + j < N) h[j] = prelu(h[j] + b0[m0 + j], a0);
    }
    float o[R];
#pragma unroll
    for (int j = 0; j < R; ++j) o[j] = 0.f;
    if constexpr (G == 1) {
      if (active) {
#pragma unroll
        for (int k = 0; k < N; ++k) {              // h[k] must stay in registers: fully unrolled
          float wrow[NPW];
          load_row<NPW>(w4 + k * NPW, wrow);
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = fmaf(wrow[j], h[k], o[j]);
        }
      }
    } else {
      if (active) {
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (m0 + j < N) hs[col * N + m0 + j] = h[j];
      }
      __syncwarp();
      if (active) {
        const float* hc = hs + col * N;
#pragma unroll 2
        for (int k = 0; k < N; ++k) {
          const float hk = hc[k];
          const float* wr = w4 + k * NPW + m0;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (m0 + j < N) o[j] = fmaf(wr[j], hk, o[j]);
        }
      }
      __syncwarp();
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (m0 + j < N) store(m0 + j, col, o[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
#ifndef CISTGCN_EMU
// ---------------------------------------------------------------------------------------------
// Tensor-core channel mix (512-thread CTAs): out(m, p) = sum_k W[m][k] * X[k][p] over the TV positions of the
// sample, as tcgen05 MMAs with M = 128 positions (two tiles), N = channels, accumulators in tensor memory.
//   1. every thread converts its share of the fp32 activation rows (X1: K1 rows, X2: K2 rows, row stride TV) into
//      the staging operand [term][k-chunk][position][8 x 16 bit]: term 0 = bf16(x), term 1 = fp16(x - bf16(x));
//   2. one elected lane of warp 0 issues, per tile and k-step, x_bf16 * w1, * w2, * w3 and x_fp16 * w_fp16, all
//      accumulated into the same Np columns, against the weight image (pack.py tc_image) that cp.async.bulk streamed
//      into the image buffer a phase earlier;
//   3. after the commit's mbarrier fires, the 16 warps read their TMEM lane quadrant (thread = position, 16 channels
//      per tcgen05.ld) and hand (first channel, position, 16 values) to the epilogue.
// Same split scheme and accuracy as the FPN kernel (fpn_tc.cuh): ~1e-6 relative.
// ---------------------------------------------------------------------------------------------
struct TcState {
  uint32_t tmem;          // TMEM base (512 columns)
  uint32_t bar;           // mbarrier: MMA completion (shared address)
  uint32_t parity;        // phase parity of the next wait on it
  uint32_t img_bar;       // mbarrier: weight image landed (cp.async.bulk complete_tx)
  uint32_t img_parity;
  unsigned char* stage;   // staging operand (aliases the adjacency region)
  const float* img;       // the one weight-image buffer
  int kc_cap;             // chunks per term the staging buffer holds (term stride = kc_cap * 4096 B)
  long long* dbg;         // optional cycle counters of thread 0 (cistgcn_debug_phase_clocks): conversion, MMA wait, epilogue, calls
};

// Thread 0: start the bulk copy of the next weight image (the previous user's MMAs have completed).
CG_DEV void tc_prefetch(const TcState& tc, const float* gsrc, int floats) {
  if (threadIdx.x == 0) {
    mbar_expect_tx(tc.img_bar, (uint32_t)floats * 4u);
    bulk_g2s(smem_u32(tc.img), gsrc, (uint32_t)floats * 4u, tc.img_bar);
  }
}

// Operand conversion and MMA issue are shared, non-inlined routines and the epilogue walks its channels four at a
// time in a rolled loop: the fused kernel is far larger than the instruction cache, so straight-line code that runs
// once per phase costs more in instruction fetch than it saves in issue slots.
template <int TV, int NT>
__device__ __noinline__ void tc_convert(unsigned char* stage, int term_stride, const float* X, int K) {
  const int kc = ((K + 15) / 16) * 2;
  for (int item = threadIdx.x; item < kc * TV; item += NT) {       // position fastest: conflict-free loads, 16-byte stores
    const int c = item / TV, p = item - c * TV;
    float x[8];
    const int k0 = c * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = (k0 + e < K) ? X[(k0 + e) * TV + p] : 0.f;
    split_store8(x, stage + (size_t)c * 4096 + p * 16, term_stride);
  }
  fence_proxy_async();
}

// One elected lane of warp 0: per tile and k-step the three bf16 weight terms and the fp16 pair accumulate into the
// same Np tensor-memory columns (tensor-memory reads of the epilogue are the scarcer resource here, unlike in fpn_tc.cuh).
static __device__ __noinline__ void tc_issue(uint32_t tmem, uint32_t stage_addr, int term_stride, uint32_t img_addr, int img_chunk,
                                      int kc, int Np, int first, uint32_t bar) {
  if (elect_one()) {
    const uint32_t i1 = umma_idesc_bf16(Np), i2 = umma_idesc_f16(Np);
    const uint32_t bch = (uint32_t)(4 * Np * 16);                  // bytes of one k-chunk of the image
    const uint64_t da0 = umma_desc(stage_addr, 4096, 128);
    const uint64_t db0 = umma_desc(img_addr, bch, 128) + (uint64_t)((img_chunk * bch) >> 4);
    const uint64_t brow = (uint64_t)((Np * 16) >> 4);              // Np rows of the image
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      const uint32_t acc = tmem + (uint32_t)(t * Np);
#pragma unroll 1
      for (int ks = 0; ks < kc / 2; ++ks) {
        const uint64_t da = da0 + (uint64_t)((2 * ks * 4096 + t * 2048) >> 4);
        const uint64_t db = db0 + (uint64_t)((2 * ks * bch) >> 4);
        umma_bf16(acc, da, db, i1, (first == 0) || ks != 0);
        umma_bf16(acc, da, db + brow, i1, 1);
        umma_bf16(acc, da, db + 2 * brow, i1, 1);
        umma_bf16(acc, da + (uint64_t)(term_stride >> 4), db + 3 * brow, i2, 1);
      }
    }
    umma_commit(bar);
  }
  __syncwarp();
}

template <int TV, int NT, class EPI>
CG_DEV void tc_gemm(TcState& tc, const float* X1, int K1, const float* X2, int K2, int M,
                    const float* next_img, int next_floats, EPI epi) {
  static_assert(NT == 512, "tc_gemm: written for 16 warps");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Np = (M + 15) & ~15;
  const int term_stride = tc.kc_cap * 4096;
  const int npass = K2 > 0 ? 2 : 1;
  int img_chunk = 0;
#pragma unroll 1
  for (int pass = 0; pass < npass; ++pass) {
    const float* X = pass ? X2 : X1;
    const int K = pass ? K2 : K1;
    const int kc = ((K + 15) / 16) * 2;
    if (pass) {                                        // the first row block's MMAs must have read the staging operand
      mbar_wait(tc.bar, tc.parity);
      tc.parity ^= 1;
    }
    const long long tq0 = CG_CLOCK();
    tc_convert<TV, NT>(tc.stage, term_stride, X, K);
    __syncthreads();
    if (tc.dbg && tid == 0) tc.dbg[0] += CG_CLOCK() - tq0;
    if (warp == 0) {
      if (pass == 0) mbar_wait(tc.img_bar, tc.img_parity);
      tc_fence_after();
      tc_issue(tc.tmem, smem_u32(tc.stage), term_stride, smem_u32(tc.img), img_chunk, kc, Np, pass == 0, tc.bar);
    }
    img_chunk += kc;
  }
  tc.img_parity ^= 1;
  const long long tq1 = CG_CLOCK();
  mbar_wait(tc.bar, tc.parity);
  tc.parity ^= 1;
  tc_fence_after();
  if (next_img) tc_prefetch(tc, next_img, next_floats);     // the image buffer is free: fetch the next user's weights
  const long long tq2 = CG_CLOCK();
  {
    const int q = warp & 3, g = warp >> 2;
    const int t = g & 1;
    const int p = t * 128 + q * 32 + lane;
    const uint32_t base = tc.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * Np);
    // warps g>>1 = 0 / 1 split the channels in halves; four channels per tcgen05.ld, rolled loop
    const int half = Np / 2, c_begin = (g >> 1) * half;
#pragma unroll 1
    for (int c = c_begin; c < c_begin + half; c += 4) {
      uint32_t v[4];
      tmem_ld4(base + c, v);
      tmem_ld_wait();
      if (p < TV && c < M) {
        float f[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
        epi(c, p, f);                          // channels c .. c+3 of position p (the callee masks m >= M)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tc.dbg && tid == 0) { tc.dbg[1] += tq2 - tq1; tc.dbg[2] += CG_CLOCK() - tq2; tc.dbg[3] += 1; }
}
#endif  // CISTGCN_EMU

// TC = true: the instantiation that carries the tensor-core channel mixes (kept apart: their code costs the FP32-FMA
// variant registers and instruction-cache space even when it never runs).
template <int T, int V, int NT, bool TC = false>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) dstd_block_kernel(const DstdArgs a) {
  CG_DYN_SMEM(smem);
  constexpr int TV = T * V, TT = T * T, VV = V * V, VP = V | 1;   // VP: odd row stride of the transposed Adj_s
  constexpr int NW = NT / 32;
  constexpr int TNW = (TV % 4 == 0) ? 4 : 2;            // column vector width of the wide GEMMs
  constexpr int VVP = (VV + 3) & ~3;                    // row stride of Adj_t and of its hidden map: V*V padded to a float4
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int* d = a.d;
  const float* __restrict__ W = a.w;
  const int Ci = d[CB_CI], Co = d[CB_CO], Ch = d[CB_CH], Cg = d[CB_CG], Hs = d[CB_HS];
  const bool has_res = d[CB_HAS_RES] != 0, interp = d[CB_INTERP] != 0;
  const int Cop = pad8i(Co);

  float* XN = smem + a.o_xn;
  float* A = smem + a.o_ab;
  float* Bt = A + a.tile;
  float* ADJ = smem + a.o_adj;
  float* ring = smem + a.o_ring;
  const int rb = a.ring_floats;
  float* p = smem + a.o_sm;
  float* stats = p;   p += pad4i(2 + 2 * T);
  float* h1 = p;                                    // gate hidden map; dead after P4, so it shares its slot
  float* dseqp = p;                                 // with the Map2Adj collapsed maps of P6-P7
  float* dspp = dseqp + pad4i(2 * Ch * V);
  p += imax(pad4i(2 * Cg * V), pad4i(2 * Ch * V) + pad4i(2 * Ch * T));
  float* h2 = p;      p += pad4i(2 * Co);
  float* zg = p;      p += pad4i(2 * Co);
  float* wg = p;      p += pad4i(2 * Co);
  float* dseq = p;    p += pad4i(2 * TV);
  float* dsp = p;     p += pad4i(2 * TV);
  float* semean = p;  p += pad4i(Co);
  float* gate = p;    p += pad4i(Co);
  float* hid = p;
  // row statistics / split-K partials live in the (still unused) adjacency region
  float* rowmean = ADJ;
  float* rowvar = rowmean + Ci * T;
  float* partial = ADJ;

  // operand accessors: resident copy in shared memory if planned, else the global blob
  auto RS = [&](int f) -> const float* { return a.res[f] >= 0 ? smem + a.res[f] : nullptr; };
  auto P = [&](int f) -> const float* { return a.res[f] >= 0 ? smem + a.res[f] : W + d[f]; };
  auto G = [&](int f) -> const float* { return W + d[f]; };

  // ---------------- once per launch: resident weights -> shared memory
  for (int f = 0; f < CB_COUNT; ++f)
    if (a.res[f] >= 0) copy_async<NT>(smem + a.res[f], W + d[f], a.wsz[f]);
  cp_async_commit();
  cp_async_wait_all();
#ifndef CISTGCN_EMU
  const bool g0_stage = a.g0_stage != 0;
  const uint32_t gbar = smem_u32(smem + a.o_gbar);
  uint32_t gpar = 0;
  if (g0_stage && tid == 0) { mbar_init(gbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  TcState tc = {0, 0, 0, 0, 0, nullptr, nullptr, 0, nullptr};
  const float* tc_first = nullptr;       // first image a sample needs, and its size
  int tc_first_floats = 0;
  bool use_tc = false;
  if constexpr (TC) {
    use_tc = a.tc != 0;
    if (use_tc) {
      uint32_t* slot = reinterpret_cast<uint32_t*>(smem + a.o_tcmisc) + 2;      // [mbarrier 8 B][slot 4 B][pad][mbarrier 8 B]
      tc.bar = smem_u32(smem + a.o_tcmisc);
      tc.img_bar = tc.bar + 16;
      tc.stage = reinterpret_cast<unsigned char*>(smem + a.o_stage);
      tc.img = smem + a.o_img;
      tc.kc_cap = a.tc_kc;
      tc_first = W + (interp ? d[CB_TC_A0] : d[CB_TC_TCN_S]);
      tc_first_floats = interp ? a.wsz[CB_TC_A0] : a.wsz[CB_TC_TCN_S];
      tc.dbg = (a.phase_clocks && blockIdx.x == 0) ? a.phase_clocks + 16 : nullptr;
      if (tid == 0) { mbar_init(tc.bar, 1); mbar_init(tc.img_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
      if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      tc.tmem = *slot;
      tc_prefetch(tc, tc_first, tc_first_floats);
    }
  }
#else
  constexpr bool use_tc = false;
#endif
  __syncthreads();

  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    CG_STAMP(0);
#ifndef CISTGCN_EMU
    if (g0_stage && tid == 0) {           // A | B are idle until P5 (the previous sample's last read is behind a barrier)
      mbar_expect_tx(gbar, (uint32_t)a.wsz[CB_G0_WT] * 4u);
      bulk_g2s(smem_u32(A), W + d[CB_G0_WT], (uint32_t)a.wsz[CB_G0_WT] * 4u, gbar);
    }
#endif
    // ---------------- P1: load + global_norm (:375); block 0 builds the 10 features (:568-577)
    if (d[CB_IN_MODE] == 1) {
      const float* src = a.in + (size_t)b * d[CB_IN_SB];
      float* raw = A;
      for (int i = tid; i < TV * 3; i += NT) raw[i] = __ldg(src + i);
      __syncthreads();
      const float* gs = P(CB_GN_S);
      const float* gb = P(CB_GN_B);
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
      const bool ibf = a.in_bf16 != 0;
      const size_t ibase = (size_t)b * d[CB_IN_SB];
      const int sc = d[CB_IN_SC], st = d[CB_IN_ST], sv = d[CB_IN_SV];
      const float* gs = P(CB_GN_S);
      const float* gb = P(CB_GN_B);
      if (sv == 1 && st == V && sc == TV && (TV % 4) == 0 && (d[CB_IN_SB] % 4) == 0) {       // contiguous tile: 128-bit loads
        for (int i = tid; i < Ci * TV / 4; i += NT) {
          const int c = (i * 4) / TV;
          float4 v4 = ld_act4(a.in, ibase + (size_t)i * 4, ibf);
          const float g0 = gs[c], b0 = gb[c];
          v4.x = fmaf(g0, v4.x, b0); v4.y = fmaf(g0, v4.y, b0); v4.z = fmaf(g0, v4.z, b0); v4.w = fmaf(g0, v4.w, b0);
          reinterpret_cast<float4*>(XN)[i] = v4;
        }
      } else {
        for (int i = tid; i < Ci * TV; i += NT) {
          const int c = i / TV, n = i - c * TV, t = n / V, v = n - t * V;
          XN[i] = fmaf(gs[c], ld_act(a.in, ibase + (size_t)c * sc + t * st + v * sv, ibf), gb[c]);
        }
      }
    }
    __syncthreads();
    CG_STAMP(1);

    // ---------------- P2: statistics (:360-371), all Bessel-corrected like torch.std
    for (int r = tid; r < Ci * T; r += NT) {                  // r = c*T + t, row of V joints
      const float* xp = XN + (r / T) * TV + (r % T) * V;
      float xv[V];
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { xv[v] = xp[v]; s += xv[v]; }
      const float mu = s / V;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { const float dd = xv[v] - mu; q = fmaf(dd, dd, q); }
      rowmean[r] = mu;
      rowvar[r] = q / (V - 1);
    }
    __syncthreads();
    if (warp == NW - 1) {
      // channel level: mean / std over (T,V) per channel, then mean / std over channels (lanes own channels)
      float sm = 0.f, ssd = 0.f;
      float sdv[2] = {0.f, 0.f};
      for (int c = lane, it = 0; c < Ci; c += 32, ++it) {
        float rm[T];
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) { rm[t] = rowmean[c * T + t]; s += rm[t]; }
        const float cm = s / T;
        float ss = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) { const float dm = rm[t] - cm; ss += (V - 1) * rowvar[c * T + t] + V * dm * dm; }
        const float sd = sqrtf(ss / (TV - 1));
        sm += cm;
        ssd += sd;
        if (it < 2) sdv[it] = sd;
      }
      sm = warp_sum(sm);
      ssd = warp_sum(ssd);
      const float m2 = ssd / Ci;
      float q = 0.f;
      for (int c = lane, it = 0; c < Ci; c += 32, ++it) { const float dd = sdv[it < 2 ? it : 1] - m2; q = fmaf(dd, dd, q); }
      q = warp_sum(q);
      if (lane == 0) { stats[0] = sm / Ci; stats[1 + T] = sqrtf(q / (Ci - 1)); }
    } else {
      // frame level: per t, mean over channels of the row means and std over channels of the row stds
      for (int t = warp; t < T; t += NW - 1) {
        float s = 0.f, s2 = 0.f;
        float sdv[2] = {0.f, 0.f};
        for (int c = lane, it = 0; c < Ci; c += 32, ++it) {
          s += rowmean[c * T + t];
          const float sd = sqrtf(rowvar[c * T + t]);
          s2 += sd;
          if (it < 2) sdv[it] = sd;
        }
        s = warp_sum(s);
        s2 = warp_sum(s2);
        const float m2 = s2 / Ci;
        float q = 0.f;
        for (int c = lane, it = 0; c < Ci; c += 32, ++it) { const float dd = sdv[it < 2 ? it : 1] - m2; q = fmaf(dd, dd, q); }
        q = warp_sum(q);
        if (lane == 0) { stats[1 + t] = s / Ci; stats[2 + T + t] = sqrtf(q / (Ci - 1)); }
      }
    }
    __syncthreads();      // row statistics are dead from here on: ADJ becomes split-K scratch
    CG_STAMP(2);

    // ---------------- P3: gate conv (T,1) + BN + PReLU -> h1  (:323-326)
    {
      const float* gb = P(CB_G0_B);
      const float* ga = P(CB_G0_A);
      auto epi = [&](int m, int v, float acc) { h1[m * V + v] = prelu(acc + gb[m], ga[m / Cg]); };
#ifndef CISTGCN_EMU
      if (g0_stage) { mbar_wait(gbar, gpar); gpar ^= 1; }
      const float* g0s = g0_stage ? A : RS(CB_G0_WT);
#else
      const float* g0s = RS(CB_G0_WT);
#endif
      if (g0s)
        gemm_narrow_auto<V, false, NT>(G(CB_G0_WT), g0s, pad8i(2 * Cg), 2 * Cg, Ci * T, XN, T, TV, partial, a.scratch_floats, ring, rb, epi);
      else          // not resident: deep-batched L2 reads; the (still unused) work tiles host the split-K partials
        gate_conv_l2<V, NT>(G(CB_G0_WT), pad8i(2 * Cg), 2 * Cg, Ci * T, XN, A, a.o_adj - a.o_ab, epi);
    }
    CG_STAMP(3);
    // ---------------- P4: gate conv (1,V) -> h2 ; MLP -> w1, w2  (:327-352, 378-384)
    {
      const float* b4 = P(CB_G4_B);
      const float* a4 = P(CB_G4_A);
      gate_matvec<NT>(P(CB_G4_WT), Cop, Co, Cg * V, h1, Cg * V, 0, nullptr, partial,
                      [&](int g, int o, float acc) { h2[g * Co + o] = prelu(acc + b4[g * Co + o], a4[g]); });
      // map_{s,t}: one warp per gate, no block barrier; the other warps go on to Map2Adj (wg is needed in P13 only)
      if (warp < 2) {
        const int g = warp;
        gate_mlp_tail(g, P(CB_M0_WT), P(CB_M0_B), P(CB_M0_A)[g], P(CB_M4_WT), Cop, Co, 2 + 2 * T, h2 + g * Co, stats, zg + g * Co,
                      [&](int o, float acc) {
                        wg[g * Co + o] = acc;
                        float* tp = g == 0 ? a.tap_w1 : a.tap_w2;
                        if (tp) tp[(size_t)b * Co + o] = acc;
                      });
      }
    }
    CG_STAMP(4);

    if (interp) {
      // ---------------- P5: Map2Adj first 1x1 convs (4 stacked) + BN + PReLU -> A (dsgn maps), B (tsgn maps)
      {
        const float* ab = P(CB_A0_B);
        const float* aa = P(CB_A0_A);
#ifndef CISTGCN_EMU
        if constexpr (TC) {
          if (use_tc)
            tc_gemm<TV, NT>(tc, XN, Ci, nullptr, 0, 4 * Ch, W + d[CB_TC_TCN_S], a.wsz[CB_TC_TCN_S], [&](int m0, int pp, float (&v)[4]) {
              const float* sab = smem + a.res[CB_A0_B];       // vectors are always resident
              const float* saa = smem + a.res[CB_A0_A];
              // all loads first, then all stores: the compiler cannot prove the tiles and the vectors disjoint
              const int br0 = m0 / Ch, r0 = m0 - br0 * Ch;
              int br = br0, r = r0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                v[i] = prelu(v[i] + sab[imin(m0 + i, 4 * Ch - 1)], saa[imin(br, 3)]);
                if (++r == Ch) { r = 0; ++br; }
              }
              br = br0; r = r0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (m0 + i < 4 * Ch) (A + (br >> 1) * a.tile)[((br & 1) * Ch + r) * TV + pp] = v[i];
                if (++r == Ch) { r = 0; ++br; }
              }
            });
        }
#endif
        const WideOp ops[1] = {{G(CB_A0_WT), RS(CB_A0_WT), XN, nullptr}};
        if (!use_tc) gemm_wide_auto<TNW, TV, TV, NT, false, 1>(ops, pad8i(4 * Ch), 4 * Ch, Ci, 0, ring, rb,
          [&](int, int m, int n0, float (&v)[TNW]) {
            const int br = m / Ch, r = m - br * Ch;
            const float bias = ab[m], sl = aa[br];
            float* tl = A + (br >> 1) * a.tile;
#pragma unroll
            for (int j = 0; j < TNW; ++j) v[j] = prelu(v[j] + bias, sl);
            store_vec<TNW>(tl + ((br & 1) * Ch + r) * TV + n0, v);       // both maps stay [c][t][v]
          });
      }
      __syncthreads();
      CG_STAMP(5);
      // ---------------- P6: collapsing convs (T,1) / (1,V) + BN  (:141-142, 149-150)
      if (RS(CB_TC3_WT_S) && RS(CB_JC3_WT_S) && RS(CB_TC3_WT_T) && RS(CB_JC3_WT_T) && Collapse4<T, V>::fits(Ch, NT, a.scratch_floats)) {
        const float* tb0 = P(CB_TC3_B_S); const float* tb1 = P(CB_TC3_B_T);
        const float* jb0 = P(CB_JC3_B_S); const float* jb1 = P(CB_JC3_B_T);
        collapse4<T, V, NT>(RS(CB_TC3_WT_S), RS(CB_JC3_WT_S), RS(CB_TC3_WT_T), RS(CB_JC3_WT_T), Ch, A, a.tile, partial,
          [&](int prob, int m, int n, float acc) {
            const int L = prob >> 1;
            if (prob & 1) dspp[L * Ch * T + m * T + n] = acc + (L ? jb1 : jb0)[m];
            else dseqp[L * Ch * V + m * V + n] = acc + (L ? tb1 : tb0)[m];
          });
      } else {
#pragma unroll
      for (int L = 0; L < 2; ++L) {
        const float* tl = A + L * a.tile;
        const float* tb = P(CB_TC3_B_S + L);
        const float* jb = P(CB_JC3_B_S + L);
        float* dq = dseqp + L * Ch * V;
        float* dp = dspp + L * Ch * T;
        gemm_narrow_auto<V, false, NT>(G(CB_TC3_WT_S + L), RS(CB_TC3_WT_S + L), pad8i(Ch), Ch, Ch * T, tl, T, TV, partial, a.scratch_floats, ring, rb,
                                [&](int m, int v, float acc) { dq[m * V + v] = acc + tb[m]; });
        gemm_narrow_auto<T, true, NT>(G(CB_JC3_WT_S + L), RS(CB_JC3_WT_S + L), pad8i(Ch), Ch, Ch * V, tl + Ch * TV, V, TV, partial, a.scratch_floats, ring, rb,
                                [&](int m, int t, float acc) { dp[m * T + t] = acc + jb[m]; });
      }
      }
      CG_STAMP(6);
      // ---------------- P7: dim_seq / dim_space (last 1x1 of each compress branch)  (:144, 152)
      for (int i = tid; i < 2 * TV; i += NT) {                // dim_seq[L][t'][v] = sum_o W6[o][t'] * dseqp[L][o][v]
        const int L = i / TV, r = i - L * TV, tq = r / V, v = r - tq * V;
        const float* wt = P(CB_TC6_WT_S + L) + tq;
        const float* xp = dseqp + L * Ch * V + v;
        float acc = 0.f;
#pragma unroll 4
        for (int o = 0; o < Ch; ++o) acc = fmaf(wt[o * pad8i(T)], xp[o * V], acc);
        dseq[i] = acc;
      }
      for (int i = tid; i < 2 * TV; i += NT) {                // dim_space[L][v'][t] = sum_o W6[o][v'] * dspp[L][o][t]
        const int L = i / TV, r = i - L * TV, vq = r / T, t = r - vq * T;
        const float* wt = P(CB_JC6_WT_S + L) + vq;
        const float* xp = dspp + L * Ch * T + t;
        float acc = 0.f;
#pragma unroll 4
        for (int o = 0; o < Ch; ++o) acc = fmaf(wt[o * pad8i(V)], xp[o * T], acc);
        dsp[i] = acc;
      }
      __syncthreads();
      CG_STAMP(7);
      // ---------------- P8-P9: space-domain outer product o[v'][t][q] = dsp[v'][t] * dseq[q][v'] (:187) and the expansor over
      // the joint axis (:165-170), fused per column (t, q) -> Adj_s, kept as [t][q][v] with the odd row stride VP
      if constexpr (TT * 2 > NT) {                            // enough columns to give every lane its own
        float* tp = a.tap_adj_s ? a.tap_adj_s + (size_t)b * V * TT : nullptr;
        expansor_fused<V, TT, NT, 1>(P(CB_E0_WT_S), P(CB_E0_B_S), P(CB_E0_A_S)[0], P(CB_E4_WT_S), Bt,
          [&](int k, int col) { const int t = col / T, q = col - t * T; return dsp[k * T + t] * dseq[q * V + k]; },
          [&](int m, int col, float val) { ADJ[col * VP + m] = val; if (tp) tp[m * TT + col] = val; });
        __syncthreads();
      } else {                                                // few columns (T*T = 100): materialise o and run the two mixes as GEMMs
        constexpr int TNS = (TT % 4 == 0) ? 4 : 2;
        for (int i = tid; i < V * TT; i += NT) {
          const int vq = i / TT, r = i - vq * TT, t = r / T, q = r - t * T;
          ADJ[i] = dsp[vq * T + t] * dseq[q * V + vq];
        }
        __syncthreads();
        {
          const float* eb = P(CB_E0_B_S);
          const float ea = P(CB_E0_A_S)[0];
          const WideOp ops[1] = {{G(CB_E0_WT_S), RS(CB_E0_WT_S), ADJ, nullptr}};
          gemm_wide_auto<TNS, TT, TT, NT, false, 1>(ops, pad8i(V), V, V, 0, ring, rb,
            [&](int, int m, int n0, float (&v)[TNS]) {
              const float bias = eb[m];
#pragma unroll
              for (int j = 0; j < TNS; ++j) v[j] = prelu(v[j] + bias, ea);
              store_vec<TNS>(Bt + m * TT + n0, v);
            });
        }
        __syncthreads();
        {
          float* tp = a.tap_adj_s ? a.tap_adj_s + (size_t)b * V * TT : nullptr;
          const WideOp ops[1] = {{G(CB_E4_WT_S), RS(CB_E4_WT_S), Bt, nullptr}};
          gemm_wide_auto<TNS, TT, TT, NT, false, 1>(ops, pad8i(V), V, V, 0, ring, rb,
            [&](int, int m, int n0, float (&v)[TNS]) {
#pragma unroll
              for (int j = 0; j < TNS; ++j) ADJ[(n0 + j) * VP + m] = v[j];
              if (tp) {
#pragma unroll
                for (int j = 0; j < TNS; ++j) tp[m * TT + n0 + j] = v[j];
              }
            });
        }
        __syncthreads();
      }
    } else {
      const float* as = W + d[CB_ADJ_S];                      // static (V,T,T) -> [t][q][v]
      for (int i = tid; i < V * TT; i += NT) { const int v = i / TT, r = i - v * TT; ADJ[r * VP + v] = __ldg(as + i); }
      __syncthreads();
    }
    CG_STAMP(8);
    // ---------------- P10: g1 = XN x_t Adj_s -> A
    if (Ci >= 4 && ((Ci + 3) / 4) * V >= NT) gcn_space<T, V, 4, NT>(XN, ADJ, A, Ci);     // enough items for every thread
    else if (Ci >= 2) gcn_space<T, V, 2, NT>(XN, ADJ, A, Ci);
    else gcn_space<T, V, 1, NT>(XN, ADJ, A, Ci);
    __syncthreads();
    CG_STAMP(9);
    // ---------------- P11: time-domain outer product + expansor over the frame axis -> Adj_t
    if (interp) {
      // o[t'][v][w] = dsp[v][t'] * dseq[t'][w] and the expansor over the frame axis, fused per column (v, w)
      constexpr int GT = (VV * 4 <= NT) ? 4 : ((VV * 2 <= NT) ? 2 : 1);
      float* tp = a.tap_adj_t ? a.tap_adj_t + (size_t)b * T * VV : nullptr;
      expansor_fused<T, VV, NT, GT>(P(CB_E0_WT_T), P(CB_E0_B_T), P(CB_E0_A_T)[0], P(CB_E4_WT_T), Bt,
        [&](int k, int col) { const int v = col / V, w = col - v * V; return dsp[TV + v * T + k] * dseq[TV + k * V + w]; },
        [&](int m, int col, float val) { ADJ[m * VVP + col] = val; if (tp) tp[m * VV + col] = val; });
      if constexpr (VVP > VV) {                               // padding columns of the float4-aligned rows
        constexpr int PADC = VVP - VV;
        for (int i = tid; i < T * PADC; i += NT) ADJ[(i / PADC) * VVP + VV + i % PADC] = 0.f;
      }
    } else {
      const float* at = W + d[CB_ADJ_T];
      for (int i = tid; i < T * VV; i += NT) ADJ[(i / VV) * VVP + i % VV] = __ldg(at + i);
    }
    __syncthreads();
    CG_STAMP(10);
    // ---------------- P12: g2 = XN x_v Adj_t -> B
    if (Ci >= 4 && ((Ci + 3) / 4) * T * 2 >= NT) gcn_time<T, V, 4, NT>(XN, ADJ, Bt, Ci);
    else if (Ci >= 2) gcn_time<T, V, 2, NT>(XN, ADJ, Bt, Ci);
    else gcn_time<T, V, 1, NT>(XN, ADJ, Bt, Ci);
    __syncthreads();
    CG_STAMP(11);
    // ---------------- P13: x_k = PReLU(BN(W g_k + b) + res); u_k = PReLU(BN(w_k * x_k))   (:266-268, 388)
    {
      const float* tb0 = P(CB_TCN_B_S); const float* tb1 = P(CB_TCN_B_T);
      const float* ps0 = P(CB_P_S_S);   const float* ps1 = P(CB_P_S_T);
      const float* pb0 = P(CB_P_B_S);   const float* pb1 = P(CB_P_B_T);
      const float ta0 = P(CB_TCN_A_S)[0], ta1 = P(CB_TCN_A_T)[0], pa0 = P(CB_P_A_S)[0], pa1 = P(CB_P_A_T)[0];
      auto tcn_epi = [&](int L, int m, int n0, float (&v)[TNW]) {
        float* Gt = L == 0 ? A : Bt;
        const float tbm = (L ? tb1 : tb0)[m], ta = L ? ta1 : ta0;
        const float sc = (L ? ps1 : ps0)[m] * wg[L * Co + m], pbm = (L ? pb1 : pb0)[m], pa = L ? pa1 : pa0;
        float r[TNW];
        if (!has_res) lds_vec<TNW>(XN + m * TV + n0, r);
#pragma unroll
        for (int j = 0; j < TNW; ++j) {
          float x = v[j] + tbm;
          if (!has_res) x += r[j];
          x = prelu(x, ta);
          v[j] = prelu(fmaf(sc, x, pbm), pa);
        }
        store_vec<TNW>(Gt + m * TV + n0, v);
      };
      const int K2 = has_res ? Ci : 0;
#ifndef CISTGCN_EMU
      if constexpr (TC) {
        if (use_tc) {
#pragma unroll 1
          for (int L = 0; L < 2; ++L) {
            float* Gt = L == 0 ? A : Bt;
            const float ta = L ? ta1 : ta0, pa = L ? pa1 : pa0;
            const int nf = L == 0 ? CB_TC_TCN_T : CB_TC_CP;
            const float* stb = smem + a.res[CB_TCN_B_S + L];
            const float* sps = smem + a.res[CB_P_S_S + L];
            const float* spb = smem + a.res[CB_P_B_S + L];
            tc_gemm<TV, NT>(tc, Gt, Ci, XN, K2, Co, W + d[nf], a.wsz[nf], [&](int m0, int pp, float (&v)[4]) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {                 // all loads first, then all stores
                const int m = imin(m0 + i, Co - 1);
                float x = v[i] + stb[m];
                if (!has_res) x += XN[m * TV + pp];
                x = prelu(x, ta);
                v[i] = prelu(fmaf(sps[m] * wg[L * Co + m], x, spb[m]), pa);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (m0 + i < Co) Gt[(m0 + i) * TV + pp] = v[i];
            });
          }
        }
      }
#endif
      if (use_tc) {
      } else if (RS(CB_TCN_WT_S) && RS(CB_TCN_WT_T)) {            // both resident: one phase for both domains
        const WideOp ops[2] = {{G(CB_TCN_WT_S), RS(CB_TCN_WT_S), A, XN}, {G(CB_TCN_WT_T), RS(CB_TCN_WT_T), Bt, XN}};
        gemm_wide_auto<TNW, TV, TV, NT, true, 2>(ops, Cop, Co, Ci, K2, ring, rb, tcn_epi);
      } else {
#pragma unroll
        for (int L = 0; L < 2; ++L) {
          const WideOp ops[1] = {{G(CB_TCN_WT_S + L), RS(CB_TCN_WT_S + L), L == 0 ? A : Bt, XN}};
          gemm_wide_auto<TNW, TV, TV, NT, true, 1>(ops, Cop, Co, Ci, K2, ring, rb,
            [&](int, int m, int n0, float (&v)[TNW]) { tcn_epi(L, m, n0, v); });
        }
      }
    }
    CG_STAMP(12);
    // ---------------- P14: compressor 1x1 over cat(u1, u2) + BN + PReLU -> A   (:305-307)
    {
      const float* cb = P(CB_CP_B);
      const float ca = P(CB_CP_A)[0];
#ifndef CISTGCN_EMU
      if constexpr (TC) {
        if (use_tc)
          tc_gemm<TV, NT>(tc, A, Co, Bt, Co, Co, has_res ? W + d[CB_TC_RS] : tc_first, has_res ? a.wsz[CB_TC_RS] : tc_first_floats,
                          [&](int m0, int pp, float (&v)[4]) {
            const float* scb = smem + a.res[CB_CP_B];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = prelu(v[i] + scb[imin(m0 + i, Co - 1)], ca);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (m0 + i < Co) A[(m0 + i) * TV + pp] = v[i];
          });
      }
#endif
      const WideOp ops[1] = {{G(CB_CP_WT), RS(CB_CP_WT), A, Bt}};
      if (!use_tc) gemm_wide_auto<TNW, TV, TV, NT, true, 1>(ops, Cop, Co, Co, Co, ring, rb,
        [&](int, int m, int n0, float (&v)[TNW]) {
          const float bias = cb[m];
#pragma unroll
          for (int j = 0; j < TNW; ++j) v[j] = prelu(v[j] + bias, ca);
          store_vec<TNW>(A + m * TV + n0, v);
        });
    }
    CG_STAMP(13);
    // ---------------- P15-P17: squeeze-excitation (SE.py:37-41)
    for (int m = warp; m < Co; m += NW) {
      float s = 0.f;
      for (int n = lane; n < TV; n += 32) s += A[m * TV + n];
      s = warp_sum(s);
      if (lane == 0) semean[m] = s / TV;
    }
    __syncthreads();
    for (int h = warp; h < Hs; h += NW) {
      const float* wt = P(CB_SE1_WT) + h;
      float acc = 0.f;
      for (int c = lane; c < Co; c += 32) acc = fmaf(wt[c * pad8i(Hs)], semean[c], acc);
      acc = warp_sum(acc);
      if (lane == 0) hid[h] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int o = tid; o < Co; o += NT) {
      const float* wt = P(CB_SE2_WT) + o;
      float acc = 0.f;
      for (int h = 0; h < Hs; ++h) acc = fmaf(wt[h * Cop], hid[h], acc);
      gate[o] = sigmoidf(acc);
    }
    __syncthreads();
    CG_STAMP(14);
    // ---------------- P18: out = c * gate + residual(xn)   (:390)
    {
      const bool obf = a.out_bf16 != 0;
      const size_t obase = (size_t)b * d[CB_OUT_SB];
      float* dst = a.out + obase;                              // fp32 view (tensor-core epilogue: fp32 outputs only)
      const int sc = d[CB_OUT_SC], st = d[CB_OUT_ST], sv = d[CB_OUT_SV];
      const bool contiguous = sv == 1 && st == V && sc == TV && (TV % 4) == 0 && (d[CB_OUT_SB] % 4) == 0;
      if (has_res) {
        const float* rbias = P(CB_RS_B);
#ifndef CISTGCN_EMU
        if constexpr (TC) {
          if (use_tc)
            tc_gemm<TV, NT>(tc, XN, Ci, nullptr, 0, Co, tc_first, tc_first_floats, [&](int m0, int pp, float (&v)[4]) {
              const float* srb = smem + a.res[CB_RS_B];
              const int t = pp / V, vv = pp - t * V;
              float* dp = dst + t * st + vv * sv;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int m = imin(m0 + i, Co - 1);
                v[i] = fmaf(A[m * TV + pp], gate[m], v[i] + srb[m]);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (m0 + i < Co) dp[(m0 + i) * sc] = v[i];
            });
        }
#endif
        const WideOp ops[1] = {{G(CB_RS_WT), RS(CB_RS_WT), XN, nullptr}};
        if (!use_tc) gemm_wide_auto<TNW, TV, TV, NT, false, 1>(ops, Cop, Co, Ci, 0, ring, rb,
          [&](int, int m, int n0, float (&v)[TNW]) {
            const float gm = gate[m], bias = rbias[m];
            float c[TNW];
            lds_vec<TNW>(A + m * TV + n0, c);
#pragma unroll
            for (int j = 0; j < TNW; ++j) v[j] = fmaf(c[j], gm, v[j] + bias);
            if (contiguous) {
              if constexpr (TNW == 4) st_act4(a.out, obase + (size_t)m * TV + n0, make_float4(v[0], v[1], v[2], v[3]), obf);
              else { st_act(a.out, obase + (size_t)m * TV + n0, v[0], obf); st_act(a.out, obase + (size_t)m * TV + n0 + 1, v[1], obf); }
            } else {
              int t = n0 / V, vv = n0 - t * V;
#pragma unroll
              for (int j = 0; j < TNW; ++j) {
                st_act(a.out, obase + (size_t)m * sc + t * st + vv * sv, v[j], obf);
                if (++vv == V) { vv = 0; ++t; }
              }
            }
          });
      } else if (contiguous) {
        for (int i = tid; i < Co * TV / 4; i += NT) {
          const float gm = gate[(i * 4) / TV];
          const float4 c4 = reinterpret_cast<const float4*>(A)[i];
          const float4 x4 = reinterpret_cast<const float4*>(XN)[i];
          st_act4(a.out, obase + (size_t)i * 4,
                  make_float4(fmaf(c4.x, gm, x4.x), fmaf(c4.y, gm, x4.y), fmaf(c4.z, gm, x4.z), fmaf(c4.w, gm, x4.w)), obf);
        }
      } else {
        for (int i = tid; i < Co * TV; i += NT) {
          const int m = i / TV, n = i - m * TV, t = n / V, v = n - t * V;
          st_act(a.out, obase + (size_t)m * sc + t * st + v * sv, fmaf(A[i], gate[m], XN[i]), obf);
        }
      }
    }
#ifndef CISTGCN_EMU
    if (g0_stage) fence_proxy_async();     // generic reads of the work tiles before the next sample's bulk copy into them
#endif
    __syncthreads();
    CG_STAMP(15);
  }
#ifndef CISTGCN_EMU
  if constexpr (TC) {
    if (use_tc) {
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tc.tmem), "r"(512u));
      }
    }
  }
#endif
}

}  // namespace cg
