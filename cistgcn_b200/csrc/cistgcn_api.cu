// C-ABI of the B200-native CIST-GCN forward path (declared in include/cistgcn_b200.h).
// Host-side orchestration only: descriptor validation, shared-memory planning, launches.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cistgcn_b200.h"
#include "dstd_launch.h"
#include "fpn_launch.h"
#include "fpn_tc_launch.h"
#include "host_util.h"
#include "simt.h"
#include "tail.cuh"

#ifndef CISTGCN_MAX_CHUNK
#define CISTGCN_MAX_CHUNK 32768
#endif

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

constexpr int kMaxSmemBytes = 227 * 1024;

using cg::sm_count;
using cg::grid_for;

int check_launch(const char* what) {
  if (int e = cg::last_launch_error()) return fail(-4, "%s launch: %s", what, cg::launch_error_string(e));
  return 0;
}

// ---- optional per-kernel CUDA-event timing (bench.py's live roofline numbers) ------------------
enum { KIND_DSTD = 0, KIND_FPN = 1, KIND_TAIL = 2, KIND_MPJPE = 3, KIND_COUNT = CISTGCN_PROFILE_KINDS };
#ifndef CISTGCN_EMU
struct ProfSlot { cudaEvent_t beg, end; int kind; };
struct Profiler {
  bool on = false;
  std::vector<ProfSlot> slots;     // events are recycled after every read
  size_t used = 0;
  long long launches[KIND_COUNT] = {0, 0, 0, 0};
  ProfSlot* next(int kind) {
    if (used == slots.size()) {
      if (slots.size() >= 16384) return nullptr;
      ProfSlot s;
      if (cudaEventCreate(&s.beg) != cudaSuccess || cudaEventCreate(&s.end) != cudaSuccess) return nullptr;
      slots.push_back(s);
    }
    slots[used].kind = kind;
    return &slots[used++];
  }
};
Profiler g_prof;
struct ProfScope {
  ProfSlot* s = nullptr;
  cudaStream_t st;
  ProfScope(int kind, void* stream) : st((cudaStream_t)stream) {
    if (!g_prof.on) return;
    g_prof.launches[kind]++;
    s = g_prof.next(kind);
    if (s) cudaEventRecord(s->beg, st);
  }
  ~ProfScope() { if (s) cudaEventRecord(s->end, st); }
};
#else
struct ProfScope { ProfScope(int, void*) {} };
#endif

// ---------------------------------------------------------------------------------------------
long long* g_phase_clocks = nullptr;   // debug hook, see cistgcn_debug_phase_clocks
int g_stamp_iter = 0;

int g_dstd_path = 0;  // 0: FP32-FMA channel mixes (default, faster at K <= 64), 1: tcgen05 channel mixes where the plan fits

int dstd_done(int e) {
  if (e) return fail(-4, "dstd_block_kernel launch: %s", cg::launch_error_string(e));
  return 0;
}

int launch_dstd(const int32_t* desc, const float* weights, const float* in, float* out, long long batch,
                const cistgcn_block_taps* taps, void* stream) {
  cg::DstdArgs a;
  memcpy(a.d, desc, sizeof(a.d));
  a.w = weights; a.in = in; a.out = out; a.batch = (int)batch;
  a.tap_adj_s = taps ? taps->adj_s : nullptr;
  a.tap_adj_t = taps ? taps->adj_t : nullptr;
  a.tap_w1 = taps ? taps->w1 : nullptr;
  a.tap_w2 = taps ? taps->w2 : nullptr;
  const int T = a.d[CB_T], V = a.d[CB_V], Ci = a.d[CB_CI], Co = a.d[CB_CO];
  if (Ci < 2 || Co < 1) return fail(-2, "DSTD-GC block needs Ci >= 2 (got %d -> %d)", Ci, Co);
  if (Co > 64 || Ci > 64) return fail(-2, "DSTD-GC block: %d -> %d channels exceed the 64 supported", Ci, Co);
  if (a.d[CB_IN_MODE] == 1 && Ci != 10) return fail(-2, "feature-building input mode needs Ci == 10");
  if (!a.d[CB_INTERP]) { a.tap_adj_s = nullptr; a.tap_adj_t = nullptr; }
  a.phase_clocks = g_phase_clocks;
  a.stamp_iter = g_stamp_iter;
  // narrow blocks: 256 threads, two CTAs per SM (if the whole plan fits half an SM's shared memory);
  // wide blocks: 512 threads, one CTA per SM
  int nt = cg::DSTD_NT_NARROW;
  if (!cg::dstd_plan(a, nt, cg::DSTD_SMEM_NARROW_BYTES / 4)) {
    nt = cg::DSTD_NT_WIDE;
    const bool tc_ok = g_dstd_path == 1 && T == 10 && (V == 22 || V == 18);     // shapes with a tensor-core instantiation
    if (!cg::dstd_plan(a, nt, kMaxSmemBytes / 4, tc_ok) || (size_t)a.smem_floats * 4 > (size_t)kMaxSmemBytes)
      return fail(-2, "DSTD-GC block (%d->%d, T=%d, V=%d) needs %zu B of shared memory (> %d)", Ci, Co, T, V,
                  (size_t)a.smem_floats * 4, kMaxSmemBytes);
  }
  ProfScope prof(KIND_DSTD, stream);
  if (nt == 512 && a.tc) return dstd_done(V == 22 ? cg::launch_dstd_10_22_512_tc(a, stream) : cg::launch_dstd_10_18_512_tc(a, stream));
#define CG_TRY_DSTD(TT, VV) \
  if (T == TT && V == VV) return dstd_done(nt == 256 ? cg::launch_dstd_##TT##_##VV##_256(a, stream) : cg::launch_dstd_##TT##_##VV##_512(a, stream));
  CG_TRY_DSTD(10, 22)
  CG_TRY_DSTD(10, 18)
  CG_TRY_DSTD(22, 25)
  CG_TRY_DSTD(18, 25)
#undef CG_TRY_DSTD
  return fail(-2, "DSTD-GC block: (T, V) = (%d, %d) has no compiled kernel "
                  "(built: (10,22), (10,18), (22,25), (18,25))", T, V);
}

int g_fpn_path = 0;   // 0: tensor-core kernel whenever the shape fits, 1: FP32-FMA kernel (cistgcn_set_fpn_path)

#ifndef CISTGCN_EMU
// tcgen05 kernel (fpn_tc.cuh): 32-wide channel tiles, two 128-position tiles, weights packed with CF_TC_*
bool fpn_tc_supported(const int32_t* fpn_descs, int n_fpn, const int32_t* tail_desc) {
  const int To = tail_desc[CT_TOUT], V = tail_desc[CT_V], Tin = tail_desc[CT_TIN];
  if (To > 25 || Tin > 16 || tail_desc[CT_F] != cg::FTC_F || (V != 22 && V != 18) || n_fpn > 4) return false;
  for (int l = 0; l < n_fpn; ++l) {
    const int32_t* f = fpn_descs + l * CF_COUNT;
    if (f[CF_TC_KC] != (l == 0 ? 2 : 4) || f[CF_TC_W] <= 0 || f[CF_TC_PRM] <= 0) return false;
    if ((f[CF_RESID] != 0) != (l > 0)) return false;       // layer 0 reads the separate input buffer, no residual there
  }
  return true;
}

int launch_fpn_tc(const int32_t* fpn_descs, int n_fpn, const int32_t* tail_desc, const float* weights,
                  const float* in, float* x7, long long batch, void* stream) {
  cg::FpnTcArgs a;
  memcpy(a.f, fpn_descs, sizeof(int32_t) * CF_COUNT * n_fpn);
  memcpy(a.t, tail_desc, sizeof(a.t));
  a.n_layers = n_fpn; a.w = weights; a.in = in; a.x7 = x7; a.batch = (int)batch;
  a.dbg = g_phase_clocks;
  int e;
  {
    ProfScope prof(KIND_FPN, stream);
    e = a.t[CT_V] == 22 ? cg::launch_fpn_tc_22(a, stream) : cg::launch_fpn_tc_18(a, stream);
  }
  if (e) return fail(-4, "fpn_tc_kernel launch: %s", cg::launch_error_string(e));
  return 0;
}
#endif

int launch_fpn(const int32_t* fpn_descs, int n_fpn, const int32_t* tail_desc, const float* weights,
               const float* in, float* x7, long long batch, void* stream) {
  if (n_fpn < 1 || n_fpn > cg::FPN_MAX_LAYERS) return fail(-2, "FPN chain: %d layers unsupported", n_fpn);
#ifndef CISTGCN_EMU
  if (g_fpn_path == 0 && fpn_tc_supported(fpn_descs, n_fpn, tail_desc))
    return launch_fpn_tc(fpn_descs, n_fpn, tail_desc, weights, in, x7, batch, stream);
#endif
  cg::FpnArgs a;
  memcpy(a.f, fpn_descs, sizeof(int32_t) * CF_COUNT * n_fpn);
  memcpy(a.t, tail_desc, sizeof(a.t));
  a.n_layers = n_fpn; a.w = weights; a.in = in; a.x7 = x7; a.batch = (int)batch;
  const int To = a.t[CT_TOUT], V = a.t[CT_V], Tin = a.t[CT_TIN];
  if (To % 5 != 0) return fail(-2, "FPN chain: output_n = %d must be a multiple of 5", To);
  if (Tin > To) return fail(-2, "FPN chain: input_n > output_n unsupported");
  if (a.t[CT_F] != 10) return fail(-2, "FPN chain: feature width %d unsupported (in_ch is fixed at 10)", a.t[CT_F]);
  for (int l = 0; l < n_fpn; ++l) {
    if (a.f[l][CF_COUT] != To || a.f[l][CF_CIN] != (l == 0 ? Tin : To)) return fail(-2, "FPN chain: layer %d channel mismatch", l);
  }
  cg::fpn_plan(a);
  const size_t smem = (size_t)a.smem_floats * sizeof(float);
  if (smem > (size_t)kMaxSmemBytes) return fail(-2, "FPN chain needs %zu B of shared memory", smem);
  int e;
  {
    ProfScope prof(KIND_FPN, stream);
    if (V == 22) e = cg::launch_fpn_22(a, stream);
    else if (V == 18) e = cg::launch_fpn_18(a, stream);
    else return fail(-2, "FPN chain: joints = %d has no compiled kernel (built: 22, 18)", V);
  }
  if (e) return fail(-4, "fpn_chain_kernel launch: %s", cg::launch_error_string(e));
  return 0;

}

int launch_tail(const int32_t* tail_desc, const float* weights, const float* x, const float* x7, const float* x8,
                float* pred, const float* target, double* frame_sums, long long batch, const cistgcn_taps* taps,
                long long tap_offset, void* stream) {
  cg::TailArgs a;
  memcpy(a.t, tail_desc, sizeof(a.t));
  a.w = weights; a.x = x; a.x7 = x7; a.x8 = x8; a.pred = pred; a.target = target; a.frame_sums = frame_sums;
  a.batch = (int)batch;
  const int To = a.t[CT_TOUT], V = a.t[CT_V], H = a.t[CT_HID];
  if (H < 64 || cg::TAIL_NT % H != 0) return fail(-2, "ContextLayer hidden_dim = %d unsupported (64, 128, 256)", H);
  if (a.t[CT_SEH1] > 8 || a.t[CT_SEH2] > 8) return fail(-2, "SE hidden width > 8 unsupported");
  a.tap_joints = taps && taps->ctx_joints ? taps->ctx_joints + tap_offset * V : nullptr;
  a.tap_disp = taps && taps->ctx_displacements ? taps->ctx_displacements + tap_offset * To : nullptr;
  a.tap_sjn = taps && taps->ctx_seq_joints_n ? taps->ctx_seq_joints_n + tap_offset * To * V : nullptr;
  a.tap_sjd = taps && taps->ctx_seq_joints_dims ? taps->ctx_seq_joints_dims + tap_offset * 3 * To * V : nullptr;
  cg::tail_plan(a);
  const size_t smem = (size_t)a.smem_floats * sizeof(float);
  auto kfn = cg::tail_kernel;
  int perr = 0;
  const int per_sm = cg::prepared_blocks_per_sm(kfn, cg::TAIL_NT, smem, &perr);
  if (perr) return fail(-3, "cudaFuncSetAttribute(%zu B): %s", smem, cg::launch_error_string(perr));
  const int grid = grid_for(batch, per_sm);
  ProfScope prof(KIND_TAIL, stream);
  CG_LAUNCH(kfn, grid, cg::TAIL_NT, smem, stream, a);
  return check_launch("tail_kernel");
}

struct PlanView {
  const int32_t* hdr;
  const int32_t* in_blocks;
  const int32_t* fpn;
  const int32_t* tail;
  const int32_t* out_blocks;
  int n_in, n_fpn, n_out, len;
};

int parse_plan(const int32_t* plan, int plan_len, PlanView& pv) {
  if (!plan) return fail(-1, "plan is NULL");
  if (plan_len >= 0 && plan_len < CP_HEADER_COUNT) return fail(-1, "plan too short");
  if (plan[CP_ABI] != CISTGCN_ABI_VERSION) return fail(-1, "plan ABI %d != library ABI %d", plan[CP_ABI], CISTGCN_ABI_VERSION);
  pv.hdr = plan;
  pv.n_in = plan[CP_N_IN_BLOCKS]; pv.n_fpn = plan[CP_N_FPN]; pv.n_out = plan[CP_N_OUT_BLOCKS];
  if (pv.n_in < 1 || pv.n_in > CISTGCN_MAX_BLOCKS || pv.n_out < 1 || pv.n_out > CISTGCN_MAX_BLOCKS || pv.n_fpn < 1 ||
      pv.n_fpn > CISTGCN_MAX_FPN)
    return fail(-1, "plan block counts out of range");
  pv.len = CP_HEADER_COUNT + (pv.n_in + pv.n_out) * CB_COUNT + pv.n_fpn * CF_COUNT + CT_COUNT;
  if (plan_len >= 0 && plan_len != pv.len) return fail(-1, "plan length %d, expected %d", plan_len, pv.len);
  pv.in_blocks = plan + CP_HEADER_COUNT;
  pv.fpn = pv.in_blocks + pv.n_in * CB_COUNT;
  pv.tail = pv.fpn + pv.n_fpn * CF_COUNT;
  pv.out_blocks = pv.tail + CT_COUNT;
  return 0;
}

size_t act_floats_per_sample(const int32_t* hdr) {
  const int T = hdr[CP_TIN], To = hdr[CP_TOUT], V = hdr[CP_V];
  return (size_t)hdr[CP_CMAX] * V * (T > To ? T : To);
}

}  // namespace

extern "C" {

const char* cistgcn_last_error(void) { return g_err.c_str(); }
int cistgcn_abi_version(void) { return CISTGCN_ABI_VERSION; }

int cistgcn_debug_phase_clocks(void* device_buffer) {
  g_phase_clocks = reinterpret_cast<long long*>(device_buffer);
  return 0;
}

int cistgcn_debug_stamp_iteration(int iteration) {
  g_stamp_iter = iteration < 0 ? 0 : iteration;
  return 0;
}

int cistgcn_set_fpn_path(int path) {
  if (path != 0 && path != 1) return fail(-1, "fpn path %d unknown (0 tensor-core when supported, 1 FP32-FMA)", path);
  g_fpn_path = path;
  return 0;
}

int cistgcn_set_dstd_path(int path) {
  if (path != 0 && path != 1) return fail(-1, "dstd path %d unknown (0 FP32-FMA channel mixes, 1 tensor-core channel mixes when the plan fits)", path);
  g_dstd_path = path;
  return 0;
}

int cistgcn_profile_enable(int on) {
#ifndef CISTGCN_EMU
  g_prof.on = on != 0;
  g_prof.used = 0;
  for (int k = 0; k < KIND_COUNT; ++k) g_prof.launches[k] = 0;
#else
  (void)on;
#endif
  return 0;
}

int cistgcn_profile_read(double* ms_by_kind, int64_t* launches_by_kind) {
  for (int k = 0; k < KIND_COUNT; ++k) { if (ms_by_kind) ms_by_kind[k] = 0.0; if (launches_by_kind) launches_by_kind[k] = 0; }
#ifndef CISTGCN_EMU
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(-4, "profile_read: device synchronize failed");
  for (size_t i = 0; i < g_prof.used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.slots[i].beg, g_prof.slots[i].end) == cudaSuccess && ms_by_kind)
      ms_by_kind[g_prof.slots[i].kind] += ms;
  }
  for (int k = 0; k < KIND_COUNT; ++k) { if (launches_by_kind) launches_by_kind[k] = g_prof.launches[k]; g_prof.launches[k] = 0; }
  g_prof.used = 0;
#endif
  return 0;
}

size_t cistgcn_workspace_bytes(const int32_t* plan, int64_t batch) {
  PlanView pv;
  if (parse_plan(plan, -1, pv)) return 0;
  const long long chunk = batch < CISTGCN_MAX_CHUNK ? (batch > 0 ? batch : 1) : CISTGCN_MAX_CHUNK;
  const size_t per = 2 * act_floats_per_sample(plan) + 2 * (size_t)plan[CP_TOUT] * plan[CP_V] * 3;
  return (size_t)chunk * per * sizeof(float) + 256;
}

int cistgcn_forward_f32(const int32_t* plan, int32_t plan_len, const float* weights, const float* x, float* pred,
                        const float* target, double* frame_sums, void* workspace, size_t workspace_bytes,
                        int64_t batch, const cistgcn_taps* taps, void* stream) {
  PlanView pv;
  if (int rc = parse_plan(plan, plan_len, pv)) return rc;
  if (batch < 0) return fail(-1, "negative batch");
  if (batch == 0) return 0;
  if (!weights || !x || !pred || !workspace) return fail(-1, "NULL buffer");
  if ((target == nullptr) != (frame_sums == nullptr)) return fail(-1, "target and frame_sums must be given together");
  if (workspace_bytes < cistgcn_workspace_bytes(plan, batch)) return fail(-1, "workspace too small");
  const int T = plan[CP_TIN], To = plan[CP_TOUT], V = plan[CP_V];
  const long long chunk_max = batch < CISTGCN_MAX_CHUNK ? batch : CISTGCN_MAX_CHUNK;
  const size_t act = act_floats_per_sample(plan);
  float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* act0 = ws;
  float* act1 = act0 + (size_t)chunk_max * act;
  float* x7 = act1 + (size_t)chunk_max * act;
  float* x8 = x7 + (size_t)chunk_max * To * V * 3;
  for (long long s = 0; s < batch; s += chunk_max) {
    const long long n = (batch - s) < chunk_max ? (batch - s) : chunk_max;
    const float* xs = x + (size_t)s * T * V * 3;
    const float* cur = xs;
    float* bufs[2] = {act0, act1};
    int flip = 0;
    for (int i = 0; i < pv.n_in; ++i) {
      const int32_t* bd = pv.in_blocks + i * CB_COUNT;
      cistgcn_block_taps bt = {nullptr, nullptr, nullptr, nullptr};
      if (taps) {
        const cistgcn_block_taps& src = taps->in_blocks[i];
        const size_t tv = (size_t)bd[CB_T] * bd[CB_V];
        bt.adj_s = src.adj_s ? src.adj_s + (size_t)s * tv * bd[CB_T] : nullptr;
        bt.adj_t = src.adj_t ? src.adj_t + (size_t)s * tv * bd[CB_V] : nullptr;
        bt.w1 = src.w1 ? src.w1 + (size_t)s * bd[CB_CO] : nullptr;
        bt.w2 = src.w2 ? src.w2 + (size_t)s * bd[CB_CO] : nullptr;
      }
      float* dst = bufs[flip];
      if (int rc = launch_dstd(bd, weights, cur, dst, n, taps ? &bt : nullptr, stream)) return rc;
      cur = dst;
      flip ^= 1;
    }
    if (int rc = launch_fpn(pv.fpn, pv.n_fpn, pv.tail, weights, cur, x7, n, stream)) return rc;
    cur = x7;
    for (int i = 0; i < pv.n_out; ++i) {
      const int32_t* bd = pv.out_blocks + i * CB_COUNT;
      cistgcn_block_taps bt = {nullptr, nullptr, nullptr, nullptr};
      if (taps) {
        const cistgcn_block_taps& src = taps->out_blocks[i];
        const size_t tv = (size_t)bd[CB_T] * bd[CB_V];
        bt.adj_s = src.adj_s ? src.adj_s + (size_t)s * tv * bd[CB_T] : nullptr;
        bt.adj_t = src.adj_t ? src.adj_t + (size_t)s * tv * bd[CB_V] : nullptr;
        bt.w1 = src.w1 ? src.w1 + (size_t)s * bd[CB_CO] : nullptr;
        bt.w2 = src.w2 ? src.w2 + (size_t)s * bd[CB_CO] : nullptr;
      }
      float* dst = (i == pv.n_out - 1) ? x8 : bufs[flip];
      if (int rc = launch_dstd(bd, weights, cur, dst, n, taps ? &bt : nullptr, stream)) return rc;
      cur = dst;
      flip ^= 1;
    }
    if (int rc = launch_tail(pv.tail, weights, xs, x7, x8, pred + (size_t)s * To * V * 3,
                             target ? target + (size_t)s * To * V * 3 : nullptr, frame_sums, n, taps, s, stream))
      return rc;
  }
  return 0;
}

int cistgcn_dstd_block_f32(const int32_t* block_desc, const float* weights, const float* in, float* out,
                           int64_t batch, const cistgcn_block_taps* taps, void* stream) {
  if (!block_desc || !weights || !in || !out) return fail(-1, "NULL buffer");
  if (batch <= 0) return batch == 0 ? 0 : fail(-1, "negative batch");
  return launch_dstd(block_desc, weights, in, out, batch, taps, stream);
}

int cistgcn_fpn_chain_f32(const int32_t* fpn_descs, int32_t n_fpn, const int32_t* tail_desc, const float* weights,
                          const float* in, float* x7, int64_t batch, void* stream) {
  if (!fpn_descs || !tail_desc || !weights || !in || !x7) return fail(-1, "NULL buffer");
  if (batch <= 0) return batch == 0 ? 0 : fail(-1, "negative batch");
  return launch_fpn(fpn_descs, n_fpn, tail_desc, weights, in, x7, batch, stream);
}

int cistgcn_tail_f32(const int32_t* tail_desc, const float* weights, const float* x, const float* x7,
                     const float* x8, float* pred, const float* target, double* frame_sums, int64_t batch,
                     const cistgcn_taps* taps, void* stream) {
  if (!tail_desc || !weights || !x || !x7 || !x8 || !pred) return fail(-1, "NULL buffer");
  if ((target == nullptr) != (frame_sums == nullptr)) return fail(-1, "target and frame_sums must be given together");
  if (batch <= 0) return batch == 0 ? 0 : fail(-1, "negative batch");
  return launch_tail(tail_desc, weights, x, x7, x8, pred, target, frame_sums, batch, taps, 0, stream);
}

int cistgcn_mpjpe_f32(const float* pred, const float* target, int64_t batch, int32_t T, int32_t V, float* err,
                      double* frame_sums, void* stream) {
  if (!pred || !target) return fail(-1, "NULL buffer");
  if (batch <= 0) return batch == 0 ? 0 : fail(-1, "negative batch");
  cg::MpjpeArgs a;
  a.pred = pred; a.target = target; a.err = err; a.frame_sums = frame_sums;
  a.n = (long long)batch * T * V; a.T = T; a.V = V;
  long long blocks = (a.n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  auto kfn = cg::mpjpe_kernel;
  ProfScope prof(KIND_MPJPE, stream);
  CG_LAUNCH(kfn, (int)(blocks < cap ? blocks : cap), 256, 0, stream, a);
  return check_launch("mpjpe_kernel");
}

}  // extern "C"
