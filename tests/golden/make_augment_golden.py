"""Generates tests/golden/aux/augment.npz from the UNMODIFIED reference augmentation classes
(environment/custom_transforms.py) and the window split of loaders/h36m_motion_3d.py:94-108 -- build container only.

np.random.uniform is replaced by a scripted source for the duration of the calls, so that the same "random" parameters
can be handed to the CUDA kernel: the reference classes draw one uniform for "does it fire" and one per parameter."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _reference as R  # noqa: E402
from cistgcn_b200 import data as D  # noqa: E402

R._modules()
T = importlib.import_module("human_motion_prediction.environment.custom_transforms")

rng = np.random.default_rng(5)
N, S, V, Tin = 7, 35, 22, 10
windows = (50.0 + 350.0 * (rng.standard_normal((N, 1, V, 3)) + (0.03 * rng.standard_normal((N, S, V, 3))).cumsum(1))).astype(np.float32)
index = np.array([3, 0, 6, 6, 2, 5], dtype=np.int64)
B = len(index)
P = D.P
params = np.zeros((B, D.NPARAM), dtype=np.float32)
noise = rng.uniform(-1, 1, size=(B, V, 3)).astype(np.float32)
outs = {k: [] for k in ("sample", "target", "sample_vel", "target_vel", "target_gvel")}
real_uniform = np.random.uniform
for b in range(B):
    fire = rng.uniform(size=8) > 0.35                      # which transforms fire for this window (a mix)
    flips = fire[:3]
    angles = rng.uniform([-5, -180, -5], [5, 180, 5])
    scales = rng.uniform(0.8, 1.2, size=3)
    trans = rng.uniform(-0.1, 0.1, size=3)
    amp = 0.02
    script = []
    for k in range(3):
        script.append(0.9 if flips[k] else 0.1)            # RandomFlip: uniform() > 0.5 per axis
    script.append(0.9 if fire[3] else 0.1)                 # RandomRotation fires?
    if fire[3]:
        script += list(angles)
    script.append(0.9 if fire[4] else 0.1)                 # RandomScale
    if fire[4]:
        script += list(scales)
    script.append(0.9 if fire[5] else 0.1)                 # RandomNoise
    if fire[5]:
        script.append(noise[b].astype(np.float64))
    script.append(0.9 if fire[6] else 0.1)                 # RandomTranslation
    if fire[6]:
        script += list(trans)
    q = list(script)

    def scripted(*a, **k):
        v = q.pop(0)
        return v

    np.random.uniform = scripted
    try:
        x = torch.from_numpy(windows[index[b]].copy())
        x = T.RandomFlip(True, True, True)(x)
        x = T.RandomRotation([-5, 5], [-180, 180], [-5, 5])(x)
        x = T.RandomScale([0.8, 1.2], [0.8, 1.2], [0.8, 1.2])(x)
        x = T.RandomNoise(amp)(x)
        x = T.RandomTranslation([-0.1, 0.1], [-0.1, 0.1], [-0.1, 0.1])(x)
    finally:
        np.random.uniform = real_uniform
    assert not q, q
    proc = x.float()
    vel = np.diff(proc, axis=0)                            # loaders/h36m_motion_3d.py:97-98
    gvel = np.linalg.norm(vel, axis=-1, keepdims=True)
    outs["sample"].append(proc[:Tin].numpy())
    outs["target"].append(proc[Tin:].numpy())
    outs["sample_vel"].append(np.asarray(vel[:Tin]))
    outs["target_vel"].append(np.asarray(vel[Tin - 1:]).cumsum(0))
    outs["target_gvel"].append(np.asarray(gvel[Tin - 1:]).cumsum(0))
    for k in range(3):
        params[b, P["CISTGCN_AUG_FLIP"] + k] = float(flips[k])
    if fire[3]:
        params[b, P["CISTGCN_AUG_ROT_ON"]] = 1
        params[b, P["CISTGCN_AUG_ROT"]: P["CISTGCN_AUG_ROT"] + 9] = D.rotvec_degrees_to_matrix(*np.float32(angles)).reshape(-1)
    if fire[4]:
        params[b, P["CISTGCN_AUG_SCALE_ON"]] = 1
        params[b, P["CISTGCN_AUG_SCALE"]: P["CISTGCN_AUG_SCALE"] + 3] = scales
    if fire[5]:
        params[b, P["CISTGCN_AUG_NOISE"]] = amp
    if fire[6]:
        params[b, P["CISTGCN_AUG_TRANS_ON"]] = 1
        params[b, P["CISTGCN_AUG_TRANS"]: P["CISTGCN_AUG_TRANS"] + 3] = trans
os.makedirs(os.path.join(HERE, "aux"), exist_ok=True)
np.savez_compressed(os.path.join(HERE, "aux", "augment.npz"), windows=windows, index=index, params=params, noise=noise, input_n=Tin,
                    **{k: np.stack(v).astype(np.float32) for k, v in outs.items()})
print({k: np.stack(v).shape for k, v in outs.items()}, params[:, [0, 1, 2, 3, 13, 17, 18]])
