"""Fused evaluation-metric kernel (csrc/eval_metrics.cu) against golden values produced by the reference's own
losses/losses.py (tests/golden/make_metrics_golden.py).  CPU: the kernel runs on the SIMT emulator; GPU: through the
nvcc-built library and the EvalMetrics front end."""
import os

import numpy as np
import pytest
import torch

from cistgcn_b200 import _cabi
from cistgcn_b200.metrics import METRICS, EvalMetrics, source_map

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux", "eval_metrics.npz"))


def _run(lib, device):
    tgt = torch.from_numpy(G["target"]).to(device)
    pred = torch.from_numpy(G["pred22"]).to(device)
    w = torch.from_numpy(G["weights"]).to(device)
    B, To, Vf, _ = tgt.shape
    smap = source_map(Vf, G["dim_used"].tolist(), G["rep32"].tolist(), G["rep22"].tolist()).to(device)
    bones = torch.from_numpy(G["bones"]).to(torch.int32).to(device).contiguous()
    sums = torch.zeros(len(METRICS), To, dtype=torch.float64, device=device)
    asm = torch.empty_like(tgt)
    stream = torch.cuda.current_stream().cuda_stream if device != "cpu" else None
    rc = lib.cistgcn_eval_metrics_f32(pred.data_ptr(), tgt.data_ptr(), smap.data_ptr(), bones.data_ptr(), bones.shape[0],
                                      w.data_ptr(), None, asm.data_ptr(), sums.data_ptr(), B, To, pred.shape[2], Vf, stream)
    _cabi.check(rc, "cistgcn_eval_metrics_f32", lib)
    return sums.cpu(), asm.cpu(), B, Vf, bones.shape[0]


def _check(sums, asm, B, Vf, nb):
    assert torch.equal(asm, torch.from_numpy(G["assembled"]))                     # the 32-joint scatter is exact
    for i, name in enumerate(METRICS):
        if name == "weighted1":
            assert float(sums[i].abs().max()) == 0.0
            continue
        ref = torch.from_numpy(G[name]).double()
        got = sums[i] / (B * (nb if name == "bone_length" else Vf))
        if name == "velocity":
            got = got[:-1]
        tol = 2e-5 * float(ref.abs().max()) if name != "pa_mpjpe" else 2e-4 * float(ref.abs().max())   # SVD conditioning
        assert (got - ref).abs().max().item() <= tol, (name, (got - ref).abs().max().item(), tol)


def test_eval_metrics_match_reference_losses_emulated():
    import _emu
    _check(*_run(_emu.lib(), "cpu"))


@pytest.mark.gpu
def test_eval_metrics_match_reference_losses_gpu():
    _check(*_run(_cabi.lib(), "cuda:0"))


@pytest.mark.gpu
def test_eval_metrics_front_end_accumulates_batches():
    tgt = torch.from_numpy(G["target"]).to("cuda:0")
    pred = torch.from_numpy(G["pred22"]).to("cuda:0")
    smap = source_map(tgt.shape[2], G["dim_used"].tolist(), G["rep32"].tolist(), G["rep22"].tolist())
    ev = EvalMetrics(25, smap, G["bones"].tolist())
    ev.compute(pred[:4], tgt[:4])
    asm = ev.compute(pred[4:], tgt[4:], want_assembled=True)
    assert torch.equal(asm.cpu(), torch.from_numpy(G["assembled"])[4:])
    vals = ev.values()
    assert torch.allclose(vals["mpjpe"].cpu(), torch.from_numpy(G["mpjpe"]), rtol=2e-5)
    assert torch.allclose(vals["pa_mpjpe"].cpu(), torch.from_numpy(G["pa_mpjpe"]), rtol=5e-4)
    assert vals["velocity"].shape == (24,)
