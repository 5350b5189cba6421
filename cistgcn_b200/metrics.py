"""Fused evaluation metrics on the GPU (reference: environment/test.py:65-94 ``Metrics.compute`` and :125-129 the
scatter of the predicted joints into the full skeleton; losses/losses.py).  One kernel (csrc/eval_metrics.cu) reads the
prediction and the target once and accumulates every metric per output frame; the reference launches ~10 chains of
ATen kernels with a ``.cpu()`` sync each."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _cabi
from .pack import F as _F

METRICS = ("mpjpe", "pa_mpjpe", "n_mpjpe", "velocity", "bone_length", "weighted0", "weighted1")
# loaders/h36m_motion_3d.py:55-56: joints of the 22-joint prediction that are copied onto ignored joints of the 32-joint skeleton
H36M_DIM_REPEAT_22 = (9, 9, 14, 16, 19, 21)
H36M_DIM_REPEAT_32 = (16, 24, 20, 23, 28, 31)


def source_map(n_full: int, dim_used: Sequence[int], dim_repeat_full: Sequence[int] = (), dim_repeat_used: Sequence[int] = ()) -> torch.Tensor:
    """[n_full] int32 table: which predicted joint lands on each joint of the full skeleton (-1: keep the target's joint).
    Mirrors ``mygt[:, :, dim_used] = outputs; mygt[:, :, dim_repeat_32] = outputs[:, :, dim_repeat_22]`` (test.py:125-129)."""
    m = torch.full((n_full,), -1, dtype=torch.int32)
    for k, j in enumerate(dim_used):
        m[j] = k
    for j, k in zip(dim_repeat_full, dim_repeat_used):
        m[j] = k
    return m


class EvalMetrics:
    """Accumulates the evaluation metrics of ``Metrics.compute`` over batches, entirely on the device.

    compute(outputs, target, weights0=None, weights1=None): outputs (B, To, Vu, 3) from the model, target (B, To, Vf, 3).
    values(): dict name -> float32 (To,) tensor = the reference's ``reduce_axis=(0, 2)`` means (velocity: To - 1 entries)."""

    def __init__(self, output_n: int, src_map: torch.Tensor, bones: Optional[Sequence[Sequence[int]]] = None, device="cuda"):
        self.To = int(output_n)
        self.device = torch.device(device)
        self.src_map = src_map.to(self.device, torch.int32).contiguous()
        self.Vf = int(src_map.numel())
        self.bones = torch.tensor(bones, dtype=torch.int32, device=self.device).contiguous() if bones is not None and len(bones) else None
        self.n_bones = 0 if self.bones is None else int(self.bones.shape[0])
        self.sums = torch.zeros(len(METRICS), self.To, dtype=torch.float64, device=self.device)
        self.count = 0

    def compute(self, outputs: torch.Tensor, target: torch.Tensor, weights0: Optional[torch.Tensor] = None,
                weights1: Optional[torch.Tensor] = None, want_assembled: bool = False) -> Optional[torch.Tensor]:
        if outputs.dim() != 4 or target.dim() != 4 or outputs.shape[:2] != target.shape[:2] or target.shape[2] != self.Vf \
                or outputs.shape[1] != self.To or outputs.shape[3] != 3 or target.shape[3] != 3:
            raise ValueError("EvalMetrics.compute: outputs (B, To, Vu, 3) / target (B, To, Vf, 3) expected")
        if not outputs.is_cuda or outputs.dtype != torch.float32 or target.dtype != torch.float32:
            raise ValueError("EvalMetrics.compute: float32 CUDA tensors required (no CPU fallback)")
        lib = _cabi.lib()
        o, t = outputs.detach().contiguous(), target.detach().contiguous()
        B, _, Vu, _ = o.shape
        w0 = weights0.contiguous() if weights0 is not None else None
        w1 = weights1.contiguous() if weights1 is not None else None
        asm = torch.empty_like(t) if want_assembled else None
        stream = torch.cuda.current_stream(o.device).cuda_stream
        with torch.cuda.device(o.device):
            rc = lib.cistgcn_eval_metrics_f32(o.data_ptr(), t.data_ptr(), self.src_map.data_ptr(),
                                              self.bones.data_ptr() if self.bones is not None else None, self.n_bones,
                                              w0.data_ptr() if w0 is not None else None, w1.data_ptr() if w1 is not None else None,
                                              asm.data_ptr() if asm is not None else None, self.sums.data_ptr(), B, self.To, Vu,
                                              self.Vf, stream)
        _cabi.check(rc, "cistgcn_eval_metrics_f32")
        self.count += B
        return asm

    def values(self) -> Dict[str, torch.Tensor]:
        out = {}
        for i, name in enumerate(METRICS):
            den = self.count * (self.n_bones if name == "bone_length" else self.Vf)
            v = (self.sums[i] / max(den, 1)).to(torch.float32)
            out[name] = v[:-1] if name == "velocity" else v
        return out
