"""GPU-side training-batch pipeline (csrc/augment.cu, cistgcn_b200/data.py) against golden batches produced by the
reference's own augmentation classes and window split (tests/golden/make_augment_golden.py).  CPU: the kernel runs on
the SIMT emulator; GPU: through the nvcc-built library and the WindowBatcher front end."""
import os

import numpy as np
import pytest
import torch

from cistgcn_b200 import _cabi
from cistgcn_b200 import data as D

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux", "augment.npz"))
KEYS = ("sample", "target", "sample_vel", "target_vel", "target_gvel")


def _run(lib, device):
    w = torch.from_numpy(G["windows"]).to(device)
    idx = torch.from_numpy(G["index"]).to(device)
    prm = torch.from_numpy(G["params"]).to(device)
    nz = torch.from_numpy(G["noise"]).to(device)
    N, S, V, _ = w.shape
    Tin, B = int(G["input_n"]), idx.numel()
    out = {"sample": torch.empty(B, Tin, V, 3, device=device), "target": torch.empty(B, S - Tin, V, 3, device=device),
           "sample_vel": torch.empty(B, Tin, V, 3, device=device), "target_vel": torch.empty(B, S - Tin, V, 3, device=device),
           "target_gvel": torch.empty(B, S - Tin, V, 1, device=device)}
    stream = torch.cuda.current_stream().cuda_stream if device != "cpu" else None
    rc = lib.cistgcn_augment_windows_f32(w.data_ptr(), idx.data_ptr(), prm.data_ptr(), nz.data_ptr(), *[out[k].data_ptr() for k in KEYS],
                                         B, S, V, Tin, stream)
    _cabi.check(rc, "cistgcn_augment_windows_f32", lib)
    return {k: v.cpu() for k, v in out.items()}


def _check(out):
    scale = float(np.abs(G["target"]).max())                     # mm-scale data: ~10^3
    for k in KEYS:
        ref = torch.from_numpy(G[k])
        err = (out[k] - ref).abs().max().item()
        assert err <= 2e-6 * scale * (25 if "vel" in k else 1), (k, err)      # cumulated targets sum up to 25 differences


def test_augment_pipeline_matches_reference_transforms_emulated():
    import _emu
    _check(_run(_emu.lib(), "cpu"))


def test_rotvec_matrix_matches_scipy():
    from scipy.spatial.transform import Rotation as R
    for v in ([3.0, -120.0, 4.5], [0.0, 0.0, 0.0], [-5.0, 180.0, 5.0]):
        assert np.allclose(D.rotvec_degrees_to_matrix(*v), R.from_rotvec(np.array(v), degrees=True).as_matrix(), atol=1e-12)


def test_augment_config_draws_like_the_reference_classes():
    cfg = D.AugmentConfig(flip=(True, True, True), rotation=([-5, 5], [-180, 180], [-5, 5]), translation=([-.1, .1],) * 3)
    prm, noise = cfg.draw(4000, 22, np.random.default_rng(0))
    P = D.P
    assert noise is None and prm.shape == (4000, D.NPARAM)
    for col in (P["CISTGCN_AUG_FLIP"], P["CISTGCN_AUG_ROT_ON"], P["CISTGCN_AUG_TRANS_ON"]):
        assert 0.45 < prm[:, col].mean() < 0.55                    # fires with probability 1 - prob_threshold
    fired = prm[prm[:, P["CISTGCN_AUG_ROT_ON"]] == 1][:, P["CISTGCN_AUG_ROT"]: P["CISTGCN_AUG_ROT"] + 9].reshape(-1, 3, 3)
    assert np.allclose(np.einsum("bij,bkj->bik", fired, fired), np.eye(3), atol=1e-5)     # proper rotations
    assert np.abs(prm[:, P["CISTGCN_AUG_TRANS"]: P["CISTGCN_AUG_TRANS"] + 3]).max() <= 0.1 + 1e-6


@pytest.mark.gpu
def test_augment_pipeline_matches_reference_transforms_gpu():
    _check(_run(_cabi.lib(), "cuda:0"))


@pytest.mark.gpu
def test_window_batcher_front_end():
    wb = D.WindowBatcher(torch.from_numpy(G["windows"]).to("cuda:0"), int(G["input_n"]))
    out = wb.batch(torch.from_numpy(G["index"]), torch.from_numpy(G["params"]), torch.from_numpy(G["noise"]))
    _check({k: v.cpu() for k, v in out.items()})
    rb = wb.random_batch(64, D.AugmentConfig(flip=(True, False, True), rotation=([-5, 5], [-180, 180], [-5, 5])), np.random.default_rng(1))
    assert rb["sample"].shape == (64, 10, 22, 3) and torch.isfinite(rb["target_gvel"]).all()
