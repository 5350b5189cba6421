"""Pins oracle/cistgcn_oracle.py to the reference: committed golden vectors (always) and the live
reference module (when /root/reference is mounted, i.e. in the build container)."""
import pytest
import torch

import _golden as G
import _reference as R
from oracle import cistgcn_oracle as O


@pytest.mark.parametrize("name", G.names())
def test_oracle_matches_golden(name):
    g = G.load(name)
    taps = {}
    itp = g["interpretable"]
    with torch.no_grad():
        pred = O.forward(g["sd"], g["cfg"], g["x"], taps=taps,
                         interpretable_in=[itp] * 5, interpretable_out=[itp])
    assert pred.shape == g["pred"].shape
    assert (pred - g["pred"]).abs().max().item() <= 0.1 * G.tol(g["pred"])
    for k, ref in g["taps"].items():
        got = taps[k]
        assert got.shape == ref.shape, k
        assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), k
    assert abs(O.mpjpe(pred, g["target"]).item() - g["mpjpe_all"].item()) <= 1e-3 * max(1, g["pred"].abs().max().item() / 4)
    assert torch.allclose(O.mpjpe(g["pred"], g["target"], (0, 2)), g["mpjpe_frames"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(O.mpjpe(g["pred"], g["target"], None), g["mpjpe_none"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(O.mpjpe(g["pred"], g["target"]), g["mpjpe_all"], rtol=1e-6, atol=1e-6)


def test_feature_quirks():
    """App. E items 1-2: positions leak into the last velocity / acceleration slots."""
    x = torch.randn(2, 10, 5, 3)
    f = O.build_features(x)                        # (B, 10, T, V)
    assert torch.equal(f[:, 0:3].permute(0, 2, 3, 1), x)
    vel = f[:, 6:9].permute(0, 2, 3, 1)
    acc = f[:, 3:6].permute(0, 2, 3, 1)
    assert torch.equal(vel[:, -1], x[:, -1])
    assert torch.equal(acc[:, -1], x[:, -1])
    assert torch.allclose(acc[:, -2], x[:, -1] - (x[:, -1] - x[:, -2]), atol=1e-6)
    assert torch.allclose(f[:, 9], vel.norm(dim=-1), atol=1e-6)


def test_stats_unbiased():
    x = torch.randn(3, 4, 10, 22)
    s = O.get_stats(x)
    assert s.shape == (3, 22)
    assert torch.allclose(s[:, 11], x.reshape(3, 4, -1).std(-1, unbiased=True).std(1, unbiased=True))


@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("E,V", [(8, 22), (32, 22), (64, 18)])
@pytest.mark.parametrize("train", [False, True])
def test_oracle_matches_live_reference(E, V, train):
    ref = R.build(E, V)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    O.stress_init_(sd)
    ref.load_state_dict(sd)
    cfg = O.OracleConfig(joints=V, input_gcn=[E] * 4)
    x, _ = O.synth_inputs(6, cfg)
    if train:
        ref.train()
        for m in ref.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    upd = {}
    with torch.no_grad():
        p_ref = ref(x)[0]
        p_or = O.forward(sd, cfg, x, train=train, bn_updates=upd)
    assert (p_ref - p_or).abs().max().item() <= 0.1 * G.tol(p_ref)
    if train:
        new = ref.state_dict()
        assert len(upd) > 0
        for k, v in upd.items():
            assert torch.allclose(new[k], v, rtol=1e-5, atol=1e-6), k
