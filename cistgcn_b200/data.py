"""GPU-side training-batch pipeline (reference: loaders/h36m_motion_3d.py:94-108 ``__getitem__`` and the augmentations
of environment/custom_transforms.py composed by loaders/loader.py:42-130).  The dataset of windows stays resident on the
device; a batch is produced by ONE kernel (csrc/augment.cu): gather the windows, apply flip -> rotation -> scale -> noise
-> translation, split into sample / target and derive the velocity targets.  The random draws are made on the host the way
the reference classes make them (one uniform for "does it fire", one per parameter), so the statistics are the same; the
reference's numpy RNG stream itself is not reproduced."""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from .pack import F as _F

P = {k: v for k, v in _F.items() if k.startswith("CISTGCN_AUG_")}
NPARAM = P["CISTGCN_AUG_PARAMS"]


def rotvec_degrees_to_matrix(rx: float, ry: float, rz: float) -> np.ndarray:
    """scipy.spatial.transform.Rotation.from_rotvec([rx, ry, rz], degrees=True).as_matrix() (Rodrigues' formula), as used by
    RandomRotation (custom_transforms.py:65)."""
    v = np.deg2rad(np.array([rx, ry, rz], dtype=np.float64))
    th = float(np.linalg.norm(v))
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


class AugmentConfig:
    """The ``transformations`` section of the training yaml (train_h36m.yaml:52-80): ranges per transform, None = absent."""

    def __init__(self, flip: Optional[Sequence[bool]] = None, rotation: Optional[Sequence[Sequence[float]]] = None,
                 scale: Optional[Sequence[Sequence[float]]] = None, noise: Optional[float] = None,
                 translation: Optional[Sequence[Sequence[float]]] = None, prob_threshold: float = 0.5):
        self.flip, self.rotation, self.scale, self.noise, self.translation = flip, rotation, scale, noise, translation
        self.prob_threshold = prob_threshold

    def draw(self, batch: int, joints: int, rng: np.random.Generator):
        """(params (B, NPARAM) float32, noise (B, V, 3) float32 or None), drawn like the reference classes' __call__."""
        prm = np.zeros((batch, NPARAM), dtype=np.float32)
        noise = None
        th = self.prob_threshold
        if self.flip is not None:
            for k in range(3):
                if self.flip[k]:
                    prm[:, P["CISTGCN_AUG_FLIP"] + k] = rng.uniform(size=batch) > th
        if self.rotation is not None:
            fire = rng.uniform(size=batch) > th
            ang = np.stack([rng.uniform(r[0], r[1], size=batch) for r in self.rotation], 1)
            for b in np.nonzero(fire)[0]:
                prm[b, P["CISTGCN_AUG_ROT_ON"]] = 1.0
                prm[b, P["CISTGCN_AUG_ROT"]: P["CISTGCN_AUG_ROT"] + 9] = rotvec_degrees_to_matrix(*ang[b]).reshape(-1)
        if self.scale is not None:
            fire = rng.uniform(size=batch) > th
            prm[:, P["CISTGCN_AUG_SCALE_ON"]] = fire
            for k in range(3):
                prm[:, P["CISTGCN_AUG_SCALE"] + k] = rng.uniform(self.scale[k][0], self.scale[k][1], size=batch)
        if self.noise:
            fire = rng.uniform(size=batch) > th
            prm[:, P["CISTGCN_AUG_NOISE"]] = np.where(fire, self.noise, 0.0)
            noise = rng.uniform(-1, 1, size=(batch, joints, 3)).astype(np.float32)
        if self.translation is not None:
            fire = rng.uniform(size=batch) > th
            prm[:, P["CISTGCN_AUG_TRANS_ON"]] = fire
            for k in range(3):
                prm[:, P["CISTGCN_AUG_TRANS"] + k] = rng.uniform(self.translation[k][0], self.translation[k][1], size=batch)
        return prm, noise


class WindowBatcher:
    """Dataset of motion windows (N, S, V, 3) resident on the device + on-device batch assembly.

    batch(index, params, noise) -> dict with the keys of the reference's ``__getitem__``: "sample" (B, input_n, V, 3),
    "target", "sample_vel", "target_vel", "target_gvel" (B, S - input_n, V, 1)."""

    def __init__(self, windows: torch.Tensor, input_n: int):
        if windows.dim() != 4 or windows.shape[3] != 3 or windows.dtype != torch.float32 or not windows.is_cuda:
            raise ValueError("WindowBatcher: float32 CUDA tensor (N, S, V, 3) required (no CPU fallback)")
        self.windows = windows.contiguous()
        self.input_n = int(input_n)
        self.N, self.S, self.V, _ = windows.shape

    def batch(self, index: torch.Tensor, params: torch.Tensor, noise: Optional[torch.Tensor] = None,
              with_velocities: bool = True) -> Dict[str, torch.Tensor]:
        dev = self.windows.device
        idx = index.to(dev, torch.int64).contiguous()
        B, To = idx.numel(), self.S - self.input_n
        prm = params.to(dev, torch.float32).contiguous()
        if prm.shape != (B, NPARAM):
            raise ValueError(f"params must be (B, {NPARAM})")
        nz = noise.to(dev, torch.float32).contiguous() if noise is not None else None
        out = {"sample": torch.empty(B, self.input_n, self.V, 3, device=dev), "target": torch.empty(B, To, self.V, 3, device=dev)}
        if with_velocities:
            out["sample_vel"] = torch.empty(B, self.input_n, self.V, 3, device=dev)
            out["target_vel"] = torch.empty(B, To, self.V, 3, device=dev)
            out["target_gvel"] = torch.empty(B, To, self.V, 1, device=dev)
        lib = _cabi.lib()
        ptr = lambda k: out[k].data_ptr() if k in out else None
        with torch.cuda.device(dev):
            rc = lib.cistgcn_augment_windows_f32(self.windows.data_ptr(), idx.data_ptr(), prm.data_ptr(),
                                                 nz.data_ptr() if nz is not None else None, ptr("sample"), ptr("target"),
                                                 ptr("sample_vel"), ptr("target_vel"), ptr("target_gvel"), B, self.S, self.V,
                                                 self.input_n, torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "cistgcn_augment_windows_f32")
        return out

    def random_batch(self, batch: int, cfg: AugmentConfig, rng: np.random.Generator) -> Dict[str, torch.Tensor]:
        idx = torch.from_numpy(rng.integers(0, self.N, size=batch))
        prm, noise = cfg.draw(batch, self.V, rng)
        return self.batch(idx, torch.from_numpy(prm), torch.from_numpy(noise) if noise is not None else None)
