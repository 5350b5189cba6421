"""MPJPE on the GPU through the C-ABI (reference: losses/losses.py:50-61)."""
from __future__ import annotations

import torch

from . import _cabi


class _MpjpeErr(torch.autograd.Function):
    """Per-joint L2 error with a hand-written backward (csrc/train_ops.cu: mpjpe_bwd_kernel): d err / d pred =
    (pred - target) / ||pred - target||, and the negative of it for the target (losses.mpjpe is called as
    loss_func(target, output), environment/train.py:76)."""

    @staticmethod
    def forward(ctx, predicted, target):
        with torch.no_grad():
            err = mpjpe(predicted.detach(), target.detach(), reduce_axis=None)
        ctx.save_for_backward(predicted.detach(), target.detach())
        return err

    @staticmethod
    def backward(ctx, derr):
        import ctypes
        p, t = ctx.saved_tensors
        lib = _cabi.bind_train(_cabi.lib())
        unit = torch.empty_like(p)
        stream = torch.cuda.current_stream(p.device).cuda_stream
        with torch.cuda.device(p.device):
            _cabi.check(lib.cistgcn_mpjpe_bwd(p.contiguous().data_ptr(), t.contiguous().data_ptr(), unit.data_ptr(),
                                              p.numel() // 3, ctypes.c_float(1.0), stream), "cistgcn_mpjpe_bwd")
        g = torch.empty_like(unit)                            # g[n, :] = unit[n, :] * derr[n]  (scale kernel: rows = joints)
        d = derr.contiguous()
        with torch.cuda.device(p.device):
            _cabi.check(lib.cistgcn_scale_fwd(unit.data_ptr(), d.data_ptr(), g.data_ptr(), p.numel() // 3, 1, 3, stream),
                        "cistgcn_scale_fwd")
        return (g if ctx.needs_input_grad[0] else None), (-g if ctx.needs_input_grad[1] else None)


def mpjpe(predicted: torch.Tensor, target: torch.Tensor, w=None, dim=-1, reduce_axis=[]):
    """Mean per-joint position error.  Same call shape as the reference:
    reduce_axis [] / () -> scalar mean (training loss), (0, 2) -> per-frame (T,), None -> (B, T, V).
    Inputs: (B, T, V, 3) float32 CUDA tensors."""
    if predicted.shape != target.shape:
        raise AssertionError("predicted and target must have the same shape")   # reference asserts (:55)
    if predicted.dim() != 4 or predicted.shape[-1] != 3 or dim not in (-1, 3):
        raise ValueError("cistgcn_b200.mpjpe: expected (B, T, V, 3) tensors reduced over the last axis")
    if predicted.dtype != torch.float32 or target.dtype != torch.float32 or not predicted.is_cuda \
            or predicted.device != target.device:
        raise ValueError("cistgcn_b200.mpjpe: float32 CUDA tensors on one device required (no CPU fallback)")
    if torch.is_grad_enabled() and (predicted.requires_grad or target.requires_grad):
        err = _MpjpeErr.apply(predicted, target)              # (B, T, V), differentiable w.r.t. both arguments
        if reduce_axis is None:
            return err
        if isinstance(reduce_axis, int):
            reduce_axis = (reduce_axis,)
        return err.mean(tuple(reduce_axis)) if len(reduce_axis) else err.mean()
    lib = _cabi.lib()
    B, T, V, _ = predicted.shape
    p, t = predicted.contiguous(), target.contiguous()
    stream = torch.cuda.current_stream(p.device).cuda_stream
    if reduce_axis is None:
        err = torch.empty(B, T, V, device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            _cabi.check(lib.cistgcn_mpjpe_f32(p.data_ptr(), t.data_ptr(), B, T, V, err.data_ptr(), None, stream),
                        "cistgcn_mpjpe_f32")
        return err
    if isinstance(reduce_axis, int):
        reduce_axis = (reduce_axis,)
    axes = tuple(sorted(a % 3 for a in reduce_axis))
    if axes not in ((), (0, 2), (0, 1, 2)):
        # other reductions (e.g. [1, 2]: per-sample MPJPE, adversarial_attacks.py:193, 521): mean of the kernel's
        # per-joint errors over the requested axes
        return mpjpe(predicted, target, w, dim, None).mean(axes)
    sums = torch.zeros(T, device=p.device, dtype=torch.float64)
    with torch.cuda.device(p.device):
        _cabi.check(lib.cistgcn_mpjpe_f32(p.data_ptr(), t.data_ptr(), B, T, V, None, sums.data_ptr(), stream),
                    "cistgcn_mpjpe_f32")
    if axes == ():
        return (sums.sum() / (B * T * V)).to(torch.float32)
    if axes == (0, 2):
        return (sums / (B * V)).to(torch.float32)
    return (sums.sum() / (B * T * V)).to(torch.float32)
