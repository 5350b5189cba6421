"""Generates tests/golden/eval_metrics.npz from the UNMODIFIED reference losses (losses/losses.py) -- build container only.

The reference's pa_mpjpe moves tensors with `.cuda()` (losses.py:105, 111); on this GPU-less container `.cuda()` is patched
to the identity for the duration of the call, nothing else is touched.  Inputs: a synthetic 32-joint H36M-like batch
(mm scale, so the `X0 ** 2 < 1e-6` quirk of :95 stays dormant as it does on real data), the 22 used joints replaced by a
perturbed prediction, scattered exactly like environment/test.py:125-129."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _reference as R  # noqa: E402

R._modules()
L = importlib.import_module("human_motion_prediction.losses.losses")
torch.Tensor.cuda = lambda self, *a, **k: self                     # see the docstring

g = torch.Generator().manual_seed(11)
B, To, Vf = 6, 25, 32
dim_used = [2, 3, 4, 5, 7, 8, 9, 10, 12, 13, 14, 15, 17, 18, 19, 21, 22, 25, 26, 27, 29, 30]     # utils/data_utils.py (H36M 22 of 32)
rep22, rep32 = [9, 9, 14, 16, 19, 21], [16, 24, 20, 23, 28, 31]
target = 50.0 + 350.0 * (torch.randn(B, 1, Vf, 3, generator=g) + (0.03 * torch.randn(B, To, Vf, 3, generator=g)).cumsum(1))
pred22 = target[:, :, dim_used] + 20.0 * torch.randn(B, To, 22, 3, generator=g)
outputs = target.clone()
outputs[:, :, dim_used] = pred22
outputs[:, :, rep32] = pred22[:, :, rep22]
w = torch.rand(B, To, Vf, generator=g)
bones = [(0, 1), (1, 2), (2, 3), (3, 4), (0, 6), (6, 7), (7, 8), (8, 9), (0, 12), (12, 13), (13, 14), (14, 15), (13, 17), (17, 18), (18, 19),
         (13, 25), (25, 26), (26, 27)]
bt = torch.tensor(bones)
dp = torch.norm(outputs[:, :, bt[:, 0]] - outputs[:, :, bt[:, 1]], dim=-1)
dt = torch.norm(target[:, :, bt[:, 0]] - target[:, :, bt[:, 1]], dim=-1)
ax = (0, 2)
out = {
    "target": target.numpy(), "pred22": pred22.numpy(), "assembled": outputs.numpy(), "weights": w.numpy(),
    "dim_used": np.array(dim_used), "rep22": np.array(rep22), "rep32": np.array(rep32), "bones": np.array(bones),
    "mpjpe": L.mpjpe(outputs, target, reduce_axis=ax).numpy(),
    "pa_mpjpe": L.pa_mpjpe(outputs.clone(), target.clone(), reduce_axis=ax).numpy(),
    "n_mpjpe": L.n_mpjpe(outputs, target, reduce_axis=ax).numpy(),
    "velocity": L.mean_velocity_error(outputs, target, reduce_axis=ax).numpy(),
    "weighted0": L.weighted_mpjpe(outputs, target, w=w, reduce_axis=ax).numpy(),
    # bone_length_error's arithmetic (:204-215) on an explicit bone list (body_utils.get_reduced_skeleton needs dataset tables)
    "bone_length": (dp - dt).abs().mean(ax).numpy(),
}
np.savez_compressed(os.path.join(HERE, "aux", "eval_metrics.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
