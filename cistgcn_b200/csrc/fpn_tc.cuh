// Time-extrapolator stack on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract as fpn_chain.cuh (4 x FPN, CISTGCN.py:38-79, with the caller's PReLU / residual :584-586,
// then dim_conversor :541-545 and the cumulative sum :588-589), but the three dilated 3x3 convolutions and
// the 1x1 `compress` run as implicit GEMMs issued with tcgen05.mma, accumulators in tensor memory.
//
// fp32 accuracy on the 16-bit tensor pipe: an activation is carried as x = x1 + x2 with x1 = bf16(x) (fp32 range)
// and x2 = fp16(x - x1) (|x - x1| <= 2^-9 |x|, so fp16's narrow range is safe and the pair holds ~20 bits); a
// weight as w = w1 + w2 + w3 (three bf16 terms, 24 bits) plus wh = fp16(w).  The weight image of a tap puts
// [w1 | w2 | w3 | wh] side by side as 128 rows, so one k-step is two MMAs: x1 * [w1|w2|w3] (kind::f16 with bf16
// operands, N = 96) and x2 * wh (fp16 operands, N = 32, accumulated into the first 32 columns); the epilogue adds
// the three 32-column blocks.  Accumulation is fp32 in TMEM.  What is dropped is below 2^-20 |x||w|.
// (An M128 K16 MMA costs >= ~63 cycles here whatever N <= 96 is -- the A operand streams from shared memory at
// ~64 B/clk, no-swizzle and SWIZZLE_128B alike (profiles/micro/umma_probe*.cu) -- so the MMA count, two per k-step,
// sets the pace; the first version carried three bf16 terms on both sides in three MMAs per k-step.)
//
// Implicit GEMM without im2col: the sample's map lives in shared memory as [term][k-chunk of 8 channels]
// [position][8 x 16 bit] with positions at a 16-byte pitch (row pitch W + 3: the three pad columns serve as
// the right padding of one row and the left padding of the next).  That is the K-major, no-swizzle
// canonical layout with SBO = 128 B, whose operand may start at ANY position: a 3x3 tap with dilation d is
// the same matrix started (kh-1)*d rows and (kw-1)*d columns away (profiles/micro/umma_probe.cu pins this).
// M = 128 positions per MMA, two tiles cover the 10 x (V+3) map.
//
// Roles (one CTA of 384 threads per SM, persistent over samples):
//   warp 0   weight producer: cp.async.bulk (1-D TMA) of per-tap slices from L2 into a 10-slot ring,
//            mbarrier complete_tx; the ring is refilled behind the second tile's pass over a branch
//   warp 1   MMA issuer (warp-uniform control flow, one elected lane issues): conv(d, tile) -> acc[tile] in TMEM; compress(d, tile) accumulates into
//            cacc[tile]; schedule conv(d,0) cmp(d-1,1) conv(d,1) cmp(d,0) keeps the pipe busy while the
//            epilogue of the other tile runs
//   warps 2-3  input loader: next sample's (Tin, F, V) map -> the two terms in a separate input buffer (layer 0
//            reads it, so the next sample's first layer overlaps this sample's last epilogue), average branch of layer 0
//   warps 4-11 epilogue, one warpgroup per tile (TMEM lane = position): branch bias + PReLU -> the two terms into
//            the staging operand of `compress`; compress bias + average branch + PReLU (+ residual) -> next
//            layer's map in place; last layer -> dim_conversor, cumsum, x7
#pragma once
#ifndef CISTGCN_EMU
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/cistgcn_b200.h"
#include "fpn_chain.cuh"
#include "simt.h"
#include "umma.cuh"

namespace cg {

constexpr int FTC_NT = 384;
constexpr int FTC_F = 10;                              // feature rows (in_ch, CISTGCN.py:512)
constexpr int FTC_SLOTS = 10;                          // 9 taps + the branch's compress slice
constexpr int FTC_BROWS = 128;                         // B rows: 3 bf16 terms + 1 fp16 copy, 32 output channels each
constexpr int FTC_BCHUNK = FTC_BROWS * 16;             // bytes of one k-chunk of a weight slice
constexpr int FTC_SLOT_BYTES = 4 * FTC_BCHUNK;         // 8192
constexpr int FTC_STAGE_CHUNK = 256 * 16;              // staging operand: [term][chunk][256 positions][8 x 16 bit]
constexpr int FTC_TMEM_COLS = 512;
constexpr int FTC_PRM_BIAS = 0, FTC_PRM_SLOPE = 96, FTC_PRM_OUT_A = 99, FTC_PRM_CPB = 100, FTC_PRM_WAVG = 132;

enum { FB_FULL = 0, FB_EMPTY = 10, FB_ACC_FULL = 20, FB_ACC_EMPTY = 22, FB_STG_FULL = 24, FB_STG_EMPTY = 26,
       FB_CACC_FULL = 28, FB_A_READY = 30, FB_IN_READY = 31, FB_IN_FREE = 32, FB_COUNT = 33 };

struct FpnTcArgs {
  int f[FPN_MAX_LAYERS][CF_COUNT];
  int t[CT_COUNT];
  int n_layers;
  const float* w;
  const void* in;      // (B, Tin, F, V): fp32, or bf16 in the BF16 instantiation
  float* x7;           // (B, Tout, V, 3)
  int batch;
  long long* dbg;      // optional (cistgcn_debug_phase_clocks): CTA 0 writes its wait / total cycle counters, 16 x int64
};

template <int V>
struct FtcGeom {
  static constexpr int WP = V + 3;                     // row pitch in positions
  static constexpr int Q0 = 3 * WP + 3;                // array row of map position (0, 0)
  static constexpr int NPOS = 6 * WP + 262;            // rows reachable by tile + tap shifts
  static constexpr int SA = NPOS * 16;                 // bytes of one [term][chunk] array
  static constexpr int O_A = 0;                                      // map: [2 terms][4 chunks][NPOS][16 B]
  static constexpr int O_IN = 8 * SA;                                // next sample's input: [2 terms][2 chunks][NPOS][16 B]
  static constexpr int O_RING = (12 * SA + 127) / 128 * 128;
  static constexpr int O_STAGE = O_RING + FTC_SLOTS * FTC_SLOT_BYTES;
  static constexpr int O_SCR = O_STAGE + 8 * FTC_STAGE_CHUNK;        // final map, fp32 [position][To]
  static constexpr int O_Y6 = O_SCR + 256 * 25 * 4;                  // 256 positions x (To <= 25) channels
  static constexpr int O_MISC = O_Y6 + (25 * V * 3 * 4 + 15) / 16 * 16;
  static constexpr int O_BAR = O_MISC + 768 * 4;                     // channel-sum partials and average-branch constants
  static constexpr int O_PRM = O_BAR + FB_COUNT * 8 + 16;            // per-layer epilogue constants + dim_conversor weights
  static constexpr int SMEM_BYTES = O_PRM + (4 * 136 + 120) * 4;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared-memory plan exceeds the 227 KB opt-in limit");
  static_assert(10 * WP <= 256, "map does not fit two 128-position tiles");
};

// 16 accumulator columns of this thread's position: the three 32-column term blocks added, small terms first
CG_DEV void ftc_load_sum(uint32_t addr, float (&o)[16]) {
  uint32_t b0[16], b1[16], b2[16];
  tmem_ld16(addr, b0);
  tmem_ld16(addr + 32, b1);
  tmem_ld16(addr + 64, b2);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = (__uint_as_float(b1[i]) + __uint_as_float(b2[i])) + __uint_as_float(b0[i]);
}

// BF16 = false: activation = bf16 term + fp16 remainder (fp32-accurate products); true: the bf16 term only
template <bool BF16>
CG_DEV void ftc_split_store8(const float* x, unsigned char* dst, int term_stride) {
  if constexpr (!BF16) { split_store8(x, dst, term_stride); }
  else {
    uint32_t t1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
      t1[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
  }
}

template <bool BF16>
CG_DEV void ftc_load_acc(uint32_t addr, float (&o)[16]) {
  if constexpr (!BF16) { ftc_load_sum(addr, o); }
  else {
    uint32_t b0[16];
    tmem_ld16(addr, b0);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(b0[i]);
  }
}

template <bool BF16>
CG_DEV void ftc_add_terms8(const unsigned char* src, int term_stride, float* x) {
  const uint4 q = *reinterpret_cast<const uint4*>(src);
  if constexpr (BF16) {
    const uint32_t wq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { x[2 * e] += __uint_as_float(wq[e] << 16); x[2 * e + 1] += __uint_as_float(wq[e] & 0xFFFF0000u); }
    return;
  }
  const uint4 r = *reinterpret_cast<const uint4*>(src + term_stride);
  const uint32_t wq[4] = {q.x, q.y, q.z, q.w}, wr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wr[e]));
    x[2 * e] += __uint_as_float(wq[e] << 16) + f.x;
    x[2 * e + 1] += __uint_as_float(wq[e] & 0xFFFF0000u) + f.y;
  }
}

// BF16 = true (cistgcn_forward_bf16): single-term bf16 operands -- ONE N = 32 MMA per k-step instead of two (N = 96 + 32),
// a quarter of the weight-ring traffic (the compact image CF_TC_W16), bf16 input activations.
template <int V, bool BF16 = false>
__global__ void __launch_bounds__(FTC_NT, 1) fpn_tc_kernel(const FpnTcArgs a) {
  using G = FtcGeom<V>;
  extern __shared__ __align__(128) unsigned char ftc_smem[];
  constexpr int WP = G::WP, Q0 = G::Q0, SA = G::SA, F = FTC_F, FV = F * V;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.n_layers, Tin = a.t[CT_TIN], To = a.t[CT_TOUT];
  const uint32_t sbase = smem_u32(ftc_smem);
  unsigned char* sA = ftc_smem + G::O_A;
  unsigned char* sStage = ftc_smem + G::O_STAGE;
  float* scr = reinterpret_cast<float*>(ftc_smem + G::O_SCR);
  float* y6 = reinterpret_cast<float*>(ftc_smem + G::O_Y6);
  float* misc = reinterpret_cast<float*>(ftc_smem + G::O_MISC);
  float* chsum = misc;              // [2][8 warps][32] per-warp channel sums of the next layer's input (fixed-order
                                    // reduction: results are bit-reproducible, no atomics)
  float* cstv = misc + 512;         // [2][32] compress bias + average branch, per layer parity
  float* cst_in = misc + 576;       // [2][32] same for layer 0, per sample parity (the loader runs a sample ahead)
  float* chsum_in = misc + 640;     // [2][2 warps][32] loader partials, per sample parity
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ftc_smem + G::O_BAR + FB_COUNT * 8);
  // epilogue constants in shared memory: with 224 KB of it carved out the L1 is tiny and an __ldg is an L2 round trip
  float* sprm = reinterpret_cast<float*>(ftc_smem + G::O_PRM);   // [layer][136]: bias[3][32], slope[3], out slope, compress bias[32]
  float* sdc = sprm + 4 * 136;                                   // dim_conversor: w0[F][8] @0, b0[3] @80, a0 @84, w3[3][8] @88, a3[3] @112
  auto bar = [&](int i) { return sbase + G::O_BAR + 8 * i; };

  if (tid == 0) {
    for (int s = 0; s < FTC_SLOTS; ++s) { mbar_init(bar(FB_FULL + s), 1); mbar_init(bar(FB_EMPTY + s), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(FB_ACC_FULL + t), 1);
      mbar_init(bar(FB_ACC_EMPTY + t), 128);
      mbar_init(bar(FB_STG_FULL + t), 128);
      mbar_init(bar(FB_STG_EMPTY + t), 1);
      mbar_init(bar(FB_CACC_FULL + t), 1);
    }
    mbar_init(bar(FB_A_READY), 256);
    mbar_init(bar(FB_IN_READY), 64);
    mbar_init(bar(FB_IN_FREE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the map (padding rows / columns / channels stay zero for the kernel's lifetime) and the small state
  for (int i = tid; i < 12 * SA / 16; i += FTC_NT) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 768; i += FTC_NT) misc[i] = 0.f;
  for (int i = tid; i < L * 136; i += FTC_NT) {
    const int l = i / 136, j = i - l * 136;
    sprm[i] = j < FTC_PRM_WAVG ? __ldg(a.w + a.f[l][CF_TC_PRM] + j) : 0.f;
  }
  for (int i = tid; i < 120; i += FTC_NT) {
    float v = 0.f;
    if (i < F * 8) v = __ldg(a.w + a.t[CT_DC0_WT] + i);
    else if (i >= 80 && i < 83) v = __ldg(a.w + a.t[CT_DC0_B] + i - 80);
    else if (i == 84) v = __ldg(a.w + a.t[CT_DC0_A]);
    else if (i >= 88 && i < 112) v = __ldg(a.w + a.t[CT_DC3_WT] + i - 88);
    else if (i >= 112 && i < 115) v = __ldg(a.w + a.t[CT_DC3_A] + i - 112);
    sdc[i] = v;
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)FTC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float inv_fv = 1.f / (float)FV;
  constexpr int BCH = BF16 ? 32 * 16 : FTC_BCHUNK;        // bytes of one k-chunk of a weight slice in the ring

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t u = 0;
      for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          const uint32_t tap_bytes = (uint32_t)a.f[l][CF_TC_KC] * BCH;
          const unsigned char* src = reinterpret_cast<const unsigned char*>(a.w + a.f[l][BF16 ? CF_TC_W16 : CF_TC_W]);
          for (int d = 0; d < 3; ++d, ++u) {
            for (int s = 0; s < FTC_SLOTS; ++s) {
              mbar_wait(bar(FB_EMPTY + s), (u & 1) ^ 1);
              const uint32_t bytes = s < 9 ? tap_bytes : (uint32_t)(4 * BCH);
              mbar_expect_tx(bar(FB_FULL + s), bytes);
              bulk_g2s(sbase + G::O_RING + s * FTC_SLOT_BYTES, src, bytes, bar(FB_FULL + s));
              src += bytes;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule (warp-uniform control flow keeps descriptors in
    // uniform registers), one elected lane issues =====
    {
      constexpr uint32_t I96 = umma_idesc_bf16(BF16 ? 32 : 96), I32H = umma_idesc_f16(32);
      constexpr uint64_t BH = (uint64_t)((96 * 16) >> 4);              // the fp16 rows of a weight slice
      const uint64_t dA0 = umma_desc(sbase + G::O_A, SA, 128);
      const uint64_t dI0 = umma_desc(sbase + G::O_IN, SA, 128);
      const uint64_t dB0 = umma_desc(sbase + G::O_RING, BCH, 128);
      const uint64_t dS0 = umma_desc(sbase + G::O_STAGE, FTC_STAGE_CHUNK, 128);
      const uint32_t acc0 = tmem, cacc0 = tmem + 192;
      uint32_t u0 = 0, ar = 0, it = 0;
      long long w_in = 0, w_ar = 0, w_acc = 0, w_full = 0, w_stg = 0;
      const long long t_begin = CG_CLOCK();
      for (int b = blockIdx.x; b < a.batch; b += gridDim.x, ++it) {
        mbar_wait_timed(bar(FB_IN_READY), it & 1, w_in);
        tc_fence_after();
        for (int l = 0; l < L; ++l, u0 += 3) {
          const int nks = a.f[l][CF_TC_KC] / 2;
          if (l > 0) { mbar_wait_timed(bar(FB_A_READY), ar & 1, w_ar); ++ar; tc_fence_after(); }
          auto conv = [&](int d, int t) {
            const uint32_t ud = u0 + d;
            mbar_wait_timed(bar(FB_ACC_EMPTY + t), (ud & 1) ^ 1, w_acc);
            tc_fence_after();
            const uint32_t acc = acc0 + t * 96;
            for (int tap = 0; tap < 9; ++tap) {
              if (t == 0) { mbar_wait_timed(bar(FB_FULL + tap), ud & 1, w_full); tc_fence_after(); }
              const int kh = tap / 3, kw = tap - kh * 3;
              const int row = Q0 + t * 128 + (kh - 1) * (d + 1) * WP + (kw - 1) * (d + 1);
              const uint64_t da_tap = (l == 0 ? dI0 : dA0) + (uint64_t)(row);          // 16 bytes per position
              const uint64_t db_tap = dB0 + (uint64_t)((tap * FTC_SLOT_BYTES) >> 4);
              if (elect_one()) {
                for (int ks = 0; ks < nks; ++ks) {
                  const uint64_t da = da_tap + (uint64_t)((2 * ks * SA) >> 4);
                  const uint64_t db = db_tap + (uint64_t)((2 * ks * BCH) >> 4);
                  umma_bf16(acc, da, db, I96, (tap | ks) != 0);
                  if constexpr (!BF16) umma_bf16(acc, da + (uint64_t)(((l == 0 ? 2 : 4) * SA) >> 4), db + BH, I32H, 1);
                }
                if (t == 1) umma_commit(bar(FB_EMPTY + tap));      // both tiles have read this tap: refill it
              }
              __syncwarp();
            }
            if (elect_one()) {
              umma_commit(bar(FB_ACC_FULL + t));
              if (l == 0 && d == 2 && t == 1) umma_commit(bar(FB_IN_FREE));      // the input buffer may take the next sample
            }
            __syncwarp();
          };
          auto cmp = [&](int d, int t) {
            const uint32_t ud = u0 + d;
            if (t == 0) mbar_wait_timed(bar(FB_FULL + 9), ud & 1, w_full);
            mbar_wait_timed(bar(FB_STG_FULL + t), ud & 1, w_stg);
            tc_fence_after();
            const uint32_t cacc = cacc0 + t * 96;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t da = dS0 + (uint64_t)((2 * ks * FTC_STAGE_CHUNK + t * 128 * 16) >> 4);
                const uint64_t db = dB0 + (uint64_t)((9 * FTC_SLOT_BYTES + 2 * ks * BCH) >> 4);
                umma_bf16(cacc, da, db, I96, (d | ks) != 0);
                if constexpr (!BF16) umma_bf16(cacc, da + (uint64_t)((4 * FTC_STAGE_CHUNK) >> 4), db + BH, I32H, 1);
              }
              umma_commit(bar(FB_STG_EMPTY + t));
              if (t == 1) umma_commit(bar(FB_EMPTY + 9));
              if (d == 2) umma_commit(bar(FB_CACC_FULL + t));
            }
            __syncwarp();
          };
          for (int d = 0; d < 3; ++d) {
            conv(d, 0);
            if (d > 0) cmp(d - 1, 1);
            conv(d, 1);
            cmp(d, 0);
          }
          cmp(2, 1);
        }
      }
      if (a.dbg && blockIdx.x == 0 && lane == 0) {
        a.dbg[0] = CG_CLOCK() - t_begin; a.dbg[1] = w_in; a.dbg[2] = w_ar; a.dbg[3] = w_acc; a.dbg[4] = w_full; a.dbg[5] = w_stg;
        a.dbg[6] = it;
      }
    }
  } else if (warp < 4) {
    // ===== input loader (64 threads) =====
    const int lt = tid - 64;
    uint32_t it = 0;
    const float* prm0 = a.w + a.f[0][CF_TC_PRM];
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x, ++it) {
      float xin[4][16];
      if constexpr (BF16) {
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.in) + (size_t)b * Tin * FV;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = lt + 64 * k;
#pragma unroll
          for (int c = 0; c < 16; ++c) xin[k][c] = (idx < FV && c < Tin) ? __bfloat162float(src[c * FV + idx]) : 0.f;
        }
      } else {
        const float* src = reinterpret_cast<const float*>(a.in) + (size_t)b * Tin * FV;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = lt + 64 * k;
#pragma unroll
          for (int c = 0; c < 16; ++c) xin[k][c] = (idx < FV && c < Tin) ? __ldg(src + c * FV + idx) : 0.f;
        }
      }
      if (it > 0) mbar_wait(bar(FB_IN_FREE), (it - 1) & 1); // the previous sample's first layer has read the buffer
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = lt + 64 * k;
        if (idx < FV) {
          const int h = idx / V, w = idx - h * V;
          unsigned char* dst = ftc_smem + G::O_IN + (size_t)(Q0 + h * WP + w) * 16;
          ftc_split_store8<BF16>(&xin[k][0], dst, 2 * SA);
          ftc_split_store8<BF16>(&xin[k][8], dst + SA, 2 * SA);
        }
      }
      float* cs = chsum_in + (it & 1) * 64;
      float mine = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float s = warp_sum((xin[0][c] + xin[1][c]) + (xin[2][c] + xin[3][c]));
        if (lane == c) mine = s;
      }
      cs[(warp - 2) * 32 + lane] = mine;
      named_bar(3, 64);
      if (warp == 2) {
        const float tot = (cs[lane] + cs[32 + lane]) * inv_fv;       // lane c holds the mean of input channel c
        float acc = __ldg(prm0 + FTC_PRM_CPB + lane);
        for (int c = 0; c < Tin; ++c) acc = fmaf(__ldg(prm0 + FTC_PRM_WAVG + c * 32 + lane), __shfl_sync(0xffffffffu, tot, c), acc);
        cst_in[(it & 1) * 32 + lane] = acc;
      }
      fence_proxy_async();
      mbar_arrive(bar(FB_IN_READY));
    }
  } else {
    // ===== epilogue: warpgroup t owns tile t, thread = position =====
    const int t = (warp - 4) >> 2, q = warp & 3;
    const int m = q * 32 + lane;
    const int p = t * 128 + m;
    const int ph = p / WP, pw = p - ph * WP;
    const bool valid = ph < F && pw < V;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t acc = tmem + t * 96 + lane_addr, cacc = tmem + 192 + t * 96 + lane_addr;
    unsigned char* stage_row = sStage + (size_t)p * 16;
    unsigned char* a_row = sA + (size_t)(Q0 + p) * 16;
    const int et = tid - 128;
    uint32_t u = 0, lv = 0, it = 0;
    long long w_in = 0, w_accf = 0, w_stge = 0, w_cacc = 0, w_nb = 0, c_ld = 0, c_e1 = 0, c_e2 = 0;
    const long long t_begin = CG_CLOCK();
    for (int b = blockIdx.x; b < a.batch; b += gridDim.x, ++it) {
      mbar_wait_timed(bar(FB_IN_READY), it & 1, w_in);      // cst_in of this sample is visible
      for (int l = 0; l < L; ++l) {
        const float* prm = sprm + l * 136;
        for (int d = 0; d < 3; ++d, ++u) {
          const float slope = prm[FTC_PRM_SLOPE + d];
          mbar_wait_timed(bar(FB_ACC_FULL + t), u & 1, w_accf);
          tc_fence_after();
          mbar_wait_timed(bar(FB_STG_EMPTY + t), (u & 1) ^ 1, w_stge);
          const long long te1 = CG_CLOCK();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float o[16];
            { const long long tl = CG_CLOCK(); ftc_load_acc<BF16>(acc + half * 16, o); c_ld += CG_CLOCK() - tl; }
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = prelu(o[i] + prm[FTC_PRM_BIAS + d * 32 + half * 16 + i], slope);
            ftc_split_store8<BF16>(&o[0], stage_row + (2 * half) * FTC_STAGE_CHUNK, 4 * FTC_STAGE_CHUNK);
            ftc_split_store8<BF16>(&o[8], stage_row + (2 * half + 1) * FTC_STAGE_CHUNK, 4 * FTC_STAGE_CHUNK);
          }
          tc_fence_before();
          mbar_arrive(bar(FB_ACC_EMPTY + t));
          fence_proxy_async();
          mbar_arrive(bar(FB_STG_FULL + t));
          c_e1 += CG_CLOCK() - te1;
        }
        // ---- compress epilogue: bias + average branch, caller's PReLU (+ residual) ----
        const bool last = l == L - 1;
        const bool resid = a.f[l][CF_RESID] != 0;
        const float oa = prm[FTC_PRM_OUT_A];
        float xo[32];                                       // residual: this layer's input at this position
#pragma unroll
        for (int i = 0; i < 32; ++i) xo[i] = 0.f;
        if (resid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) ftc_add_terms8<BF16>(a_row + c * SA, 4 * SA, &xo[8 * c]);
        }
        mbar_wait_timed(bar(FB_CACC_FULL + t), lv & 1, w_cacc);
        ++lv;
        tc_fence_after();
        { const long long tb = CG_CLOCK(); named_bar(1, 256); w_nb += CG_CLOCK() - tb; }   // this layer's cst is in place
        const float* cst = l == 0 ? cst_in + (it & 1) * 32 : cstv + (l & 1) * 32;
        float* csn = chsum + ((l + 1) & 1) * 256;
        const long long te2 = CG_CLOCK();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float o[16];
          ftc_load_acc<BF16>(cacc + half * 16, o);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float val = prelu(o[i] + cst[half * 16 + i], oa) + xo[half * 16 + i];
            xo[half * 16 + i] = valid ? val : 0.f;          // xo now holds the layer's output at this position
          }
          if (!last) {
            ftc_split_store8<BF16>(&xo[half * 16], a_row + (2 * half) * SA, 4 * SA);
            ftc_split_store8<BF16>(&xo[half * 16 + 8], a_row + (2 * half + 1) * SA, 4 * SA);
          }
        }
        tc_fence_before();
        if (!last) {
          fence_proxy_async();
          mbar_arrive(bar(FB_A_READY));                     // the next layer's convolutions may start
          c_e2 += CG_CLOCK() - te2;
          float mine = 0.f;                                 // channel sums for its average branch, off the critical path
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float sm = warp_sum(xo[i]);
            if (lane == i) mine = sm;
          }
          csn[(warp - 4) * 32 + lane] = mine;
        } else if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < To) scr[p * To + i] = xo[i];
        }
        named_bar(1, 256);                                  // channel sums complete / final map complete
        if (!last) {
          if (warp == 4) {
            const float* prn = a.w + a.f[l + 1][CF_TC_PRM];
            float tot = 0.f;
#pragma unroll
            for (int wq = 0; wq < 8; ++wq) tot += csn[wq * 32 + lane];
            tot *= inv_fv;                                  // lane c holds the mean of channel c
            float accv = __ldg(prn + FTC_PRM_CPB + lane);
            for (int c = 0; c < To; ++c) accv = fmaf(__ldg(prn + FTC_PRM_WAVG + c * 32 + lane), __shfl_sync(0xffffffffu, tot, c), accv);
            cstv[((l + 1) & 1) * 32 + lane] = accv;
          }
        } else {
          // dim_conversor on (F channels, To, V): conv1x1 F->3, BN, PReLU, conv1x1 3->3, PReLU(3)  (:541-545)
          const float* w0 = sdc;
          const float a0 = sdc[84];
          const float* w3 = sdc + 88;
          for (int i = et; i < To * V; i += 256) {
            const int fr = i / V, v = i - fr * V;
            float y[3] = {sdc[80], sdc[81], sdc[82]};
            for (int c = 0; c < F; ++c) {
              const float xv = scr[(c * WP + v) * To + fr];
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
              y6[i * 3 + k] = prelu(z, sdc[112 + k]);
            }
          }
          named_bar(1, 256);
          float* dst = a.x7 + (size_t)b * To * V * 3;           // x7 = cumsum over frames (:589)
          for (int i = et; i < V * 3; i += 256) {
            float s = 0.f;
            for (int fr = 0; fr < To; ++fr) { s += y6[fr * V * 3 + i]; dst[fr * V * 3 + i] = s; }
          }
        }
      }
    }
    if (a.dbg && blockIdx.x == 0 && (tid == 128 || tid == 256)) {
      long long* o = a.dbg + (tid == 128 ? 8 : 16);
      o[0] = CG_CLOCK() - t_begin; o[1] = w_in; o[2] = w_accf; o[3] = w_stge; o[4] = w_cacc; o[5] = w_nb;
      if (tid == 128) { a.dbg[24] = c_ld; a.dbg[25] = c_e1; a.dbg[26] = c_e2; }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)FTC_TMEM_COLS));
  }
}

}  // namespace cg
#endif  // CISTGCN_EMU
