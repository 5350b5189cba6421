/*
 * cistgcn_b200 -- C-ABI of the differentiable CIST-GCN path: the building blocks of the reference's training step
 * (human_motion_prediction/environment/train.py:54-107: train-mode forward, losses.mpjpe with reduce_axis=[],
 * backward, Adam with L2-in-gradient weight decay, environment/utils.py:53-57) and of its eval-mode input-gradient path
 * (environment/adversarial_attacks.py:184, 422, 495-511).
 *
 * Train mode breaks the per-sample fusion of the inference kernels (include/cistgcn_b200.h): every one of the model's
 * ~160 BatchNorms needs statistics over the WHOLE batch between two layers (CISTGCN.py:143, 236, 325, 375 ...; the
 * reference has no SyncBN, statistics are per replica).  The differentiable path therefore runs layer by layer: each
 * entry point below is one hand-written sm_100a kernel (forward or backward of one layer type) on fp32 NCHW tensors,
 * and cistgcn_b200/train.py replays the reference's layer sequence (CISTGCN.py:567-597) through them and keeps the tape.
 * Nothing here calls cuDNN / cuBLAS / ATen.
 *
 * Conventions: plain pointers and sizes; all tensors fp32, contiguous, caller-owned, on the current device; `stream` is
 * a cudaStream_t passed as void*; return 0 or a negative code (message: cistgcn_last_error()).  "bwd" entry points WRITE
 * their outputs unless the name / comment says they accumulate.
 */
#ifndef CISTGCN_B200_TRAIN_H
#define CISTGCN_B200_TRAIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* nn.Conv2d / nn.Conv1d(k=1) / nn.Linear (H = W = kh = kw = 1), stride 1: x (B,Ci,H,W), w (Co,Ci,kh,kw), y (B,Co,Ho,Wo),
 * Ho = H + 2*ph - dh*(kh-1), Wo = W + 2*pw - dw*(kw-1).  bias / dbias may be NULL. */
typedef struct cistgcn_conv_shape {
  int64_t B;
  int32_t Ci, H, W, Co, kh, kw, ph, pw, dh, dw;
} cistgcn_conv_shape;
int cistgcn_conv2d_fwd(const cistgcn_conv_shape* s, const float* x, const float* w, const float* bias, float* y, void* stream);
int cistgcn_conv2d_bwd_input(const cistgcn_conv_shape* s, const float* dy, const float* w, float* dx, void* stream);
/* dW (and dbias) from partial sums over chunks of the (b, ho, wo) range, added in a fixed order; `scratch` holds them:
 * cistgcn_conv2d_bwd_weight_scratch_floats(s) floats. */
size_t cistgcn_conv2d_bwd_weight_scratch_floats(const cistgcn_conv_shape* s);
int cistgcn_conv2d_bwd_weight(const cistgcn_conv_shape* s, const float* x, const float* dy, float* dw, float* dbias, float* scratch,
                              void* stream);

/* nn.BatchNorm{1,2}d over x (B,C,HW).  training != 0: batch statistics (biased variance) normalise, running_mean / running_var
 * are updated in place with `momentum` (unbiased variance), save_mean / save_invstd [C] are written for the backward.
 * training == 0: running statistics normalise, nothing is updated (save_* are written with the running values).
 * `scratch`: cistgcn_bn_scratch_floats(C) floats, 8-byte aligned (chunked batch statistics kept in fp64, merged with
 * Chan's update in a fixed order; every BatchNorm reduction accumulates in fp64 like ATen's CPU kernels). */
size_t cistgcn_bn_scratch_floats(int32_t C);
int cistgcn_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float* y,
                   float* save_mean, float* save_invstd, float* scratch, int64_t B, int32_t C, int32_t HW, int32_t training,
                   float momentum, float eps, void* stream);
/* dgamma / dbeta may be NULL (input-gradient path).  training == 0: dx = dy * gamma * invstd. */
int cistgcn_bn_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                   float* dx, float* dgamma, float* dbeta, float* scratch, int64_t B, int32_t C, int32_t HW, int32_t training,
                   void* stream);

/* nn.PReLU with n_slopes in {1, C} on x (B,C,HW).  dslope [n_slopes] may be NULL; otherwise `scratch` must hold
 * n_slopes * CISTGCN_PRELU_SCRATCH_PER_SLOPE DOUBLES, 8-byte aligned (fp64 partial sums, added in a fixed order:
 * bit-reproducible; the slope gradients are heavily cancelling sums). */
#define CISTGCN_PRELU_SCRATCH_PER_SLOPE 64
int cistgcn_prelu_fwd(const float* x, const float* slope, float* y, int64_t B, int32_t C, int32_t HW, int32_t n_slopes, void* stream);
int cistgcn_prelu_bwd(const float* x, const float* dy, const float* slope, float* dx, float* dslope, float* scratch, int64_t B,
                      int32_t C, int32_t HW, int32_t n_slopes, void* stream);

/* elementwise: kind 0 relu, 1 sigmoid.  bwd takes the forward OUTPUT y. */
int cistgcn_act_fwd(const float* x, float* y, int64_t n, int32_t kind, void* stream);
int cistgcn_act_bwd(const float* y, const float* dy, float* dx, int64_t n, int32_t kind, void* stream);

/* nn.Dropout(p): y = x * keep(seed, i) / (1 - p), keep from a counter-based hash of (seed, *step, element index); the same
 * call with dy in place of x is the backward.  `step` (device memory, may be NULL) is a per-step counter read by the
 * kernel: a captured CUDA graph of the training step draws fresh masks on every replay.  cistgcn_counter_bump adds 1. */
int cistgcn_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* step, void* stream);
int cistgcn_counter_bump(uint64_t* counter, void* stream);

/* dst[i0,i1,i2,i3] (=|+=) src[i0,i1,i2,i3] with arbitrary element strides on both sides: permute, cat / split (channel
 * slices), broadcast (source stride 0), residual adds. */
int cistgcn_copy4d(float* dst, const int64_t dst_strides[4], const float* src, const int64_t src_strides[4],
                   const int64_t sizes[4], int32_t accumulate, void* stream);
/* y = a * x + b * y over n floats (flat buffers: gradient accumulation, averaging after the all-reduce). */
int cistgcn_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream);

/* ConvTemporalGraphical (CISTGCN.py:110, 117, 123).  domain 0 "space": y[n,c,q,v] = sum_t x[n,c,t,v] A[n,v,t,q];
 * domain 1 "time": y[n,c,t,w] = sum_v x[n,c,t,v] A[n,t,v,w].  a_batched == 0: A has no sample axis (static parameter);
 * its gradient is then summed over the batch. */
int cistgcn_gcn_fwd(const float* x, const float* A, float* y, int64_t B, int32_t C, int32_t T, int32_t V, int32_t domain,
                    int32_t a_batched, void* stream);
int cistgcn_gcn_bwd(const float* x, const float* A, const float* dy, float* dx, float* dA, int64_t B, int32_t C, int32_t T,
                    int32_t V, int32_t domain, int32_t a_batched, void* stream);

/* Map2Adj outer products (CISTGCN.py:183-187): dsp (B,V,T), dseq (B,T,V).
 * domain 0: o[n,v,t,q] = dsp[n,v,t] * dseq[n,q,v];  domain 1: o[n,t,v,w] = dsp[n,v,t] * dseq[n,t,w]. */
int cistgcn_outer_fwd(const float* dsp, const float* dseq, float* o, int64_t B, int32_t T, int32_t V, int32_t domain, void* stream);
int cistgcn_outer_bwd(const float* dsp, const float* dseq, const float* d_o, float* d_dsp, float* d_dseq, int64_t B, int32_t T,
                      int32_t V, int32_t domain, void* stream);

/* DSTD_GC._get_stats_ (CISTGCN.py:360-371): x (B,C,T,V) -> stats (B, 2+2T), unbiased std.  bwd ACCUMULATES into dx. */
int cistgcn_stats_fwd(const float* x, float* stats, int64_t B, int32_t C, int32_t T, int32_t V, void* stream);
int cistgcn_stats_bwd(const float* x, const float* stats, const float* dstats, float* dx, int64_t B, int32_t C, int32_t T,
                      int32_t V, void* stream);

/* squeeze / excite helpers: mean over HW -> (B,C);  y = x * g[b,c] (also the context gating w[..., None, None] * x). */
int cistgcn_spatial_mean_fwd(const float* x, float* m, int64_t B, int32_t C, int32_t HW, void* stream);
int cistgcn_spatial_mean_bwd(const float* dm, float* dx, int64_t B, int32_t C, int32_t HW, int32_t accumulate, void* stream);
int cistgcn_scale_fwd(const float* x, const float* g, float* y, int64_t B, int32_t C, int32_t HW, void* stream);
int cistgcn_scale_bwd(const float* x, const float* g, const float* dy, float* dx, float* dg, int64_t B, int32_t C, int32_t HW,
                      void* stream);

/* max over the last axis: x (R, N) -> y (R), idx (R) int32; bwd scatters dy into a ZEROED dx. */
int cistgcn_rowmax_fwd(const float* x, float* y, int32_t* idx, int64_t R, int32_t N, void* stream);
int cistgcn_rowmax_bwd(const float* dy, const int32_t* idx, float* dx, int64_t R, int32_t N, void* stream);

/* cumsum over axis 1 of (B, L, N) (CISTGCN.py:589); reverse != 0 is its adjoint. */
int cistgcn_cumsum(const float* x, float* y, int64_t B, int32_t L, int32_t N, int32_t reverse, void* stream);

/* feature build (CISTGCN.py:568-577): x (B,T,V,3) -> f (B,10,T,V); bwd: df -> dx. */
int cistgcn_features_fwd(const float* x, float* f, int64_t B, int32_t T, int32_t V, void* stream);
int cistgcn_features_bwd(const float* x, const float* df, float* dx, int64_t B, int32_t T, int32_t V, void* stream);

/* losses.mpjpe, reduce_axis=[] (losses.py:50-61): loss_sum (double, ACCUMULATED into) = sum_{b,t,v} ||pred - target||_2;
 * bwd: dpred = scale * (pred - target) / ||pred - target||_2  (scale = upstream gradient / (B*T*V)). */
int cistgcn_mpjpe_bwd(const float* pred, const float* target, float* dpred, int64_t n_joints, float scale, void* stream);

/* torch.optim.Adam semantics (amsgrad off): g += wd * p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 * p -= lr / (1 - b1^step) * m / (sqrt(v / (1 - b2^step)) + eps).  `grad_scale` multiplies g first (1/world after the
 * gradient all-reduce). */
int cistgcn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int32_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CISTGCN_B200_TRAIN_H */
