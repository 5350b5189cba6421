"""Kernel logic on CPU: the *same* CUDA sources compiled against the SIMT emulator (tests/emu),
driven through the same C-ABI, checked against the golden vectors and the oracle.  This is test
infrastructure; the product never loads the emulator build."""
import pytest
import torch

import _emu
import _golden as G
import _models as M
from oracle import cistgcn_oracle as O


@pytest.mark.parametrize("name", G.names())
def test_emulated_forward_matches_golden(name):
    g = G.load(name)
    m = M.make_opt(g["embed"], g["cfg"].joints, g["interpretable"])
    from cistgcn_b200 import CISTGCN
    model = CISTGCN(m.architecture_config, m.learning_config).eval()
    model.load_state_dict(g["sd"])
    pred, sums, taps = _emu.forward(g["sd"], model.geometry(), g["x"], g["target"])
    assert (pred - g["pred"]).abs().max().item() <= G.tol(g["pred"])
    B, To, V = pred.shape[0], pred.shape[1], pred.shape[2]
    scale = max(1.0, g["pred"].abs().max().item() / 4)
    assert abs((sums.sum() / (B * To * V)).item() - g["mpjpe_all"].item()) <= 1e-3 * scale
    assert torch.allclose((sums / (B * V)).float(), g["mpjpe_frames"], rtol=1e-5, atol=1e-3 * scale)
    # intermediates: relative 1e-4 against the reference's own attribute taps
    for k, ref in g["taps"].items():
        got = taps[k].reshape(ref.shape)
        assert (got - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), k


def test_emulated_e32_block_widths():
    """E=32 exercises the 2-m-tile / multi-pass in-place GEMM paths that E=8 does not."""
    model, sd, cfg = M.build(32, 18, "W2")
    x, tgt = O.synth_inputs(3, cfg)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x)
    pred, sums, _ = _emu.forward(sd, model.geometry(), x, tgt, want_taps=False)
    assert (pred - ref).abs().max().item() <= G.tol(ref)
    assert abs((sums.sum() / (3 * 25 * 18)).item() - O.mpjpe(ref, tgt).item()) <= 1e-3 * max(1, ref.abs().max().item() / 4)


@pytest.mark.parametrize("E,V", [(32, 18)])
def test_emulated_mix_mma_matches_ffma_and_oracle(E, V):
    """Stage 3 of the DSTD-GC path on 3xTF32 mma.sync (csrc/dstd_mix_mma.cuh; the emulator executes the PTX fragment
    layouts of m16n8k8 with TF32-truncated operands) against the FP32-FMA tile loops and the oracle."""
    from cistgcn_b200 import _cabi
    from cistgcn_b200.pack import F, pack_state_dict
    model, sd, cfg = M.build(E, V, "W2")
    x, _ = O.synth_inputs(2, cfg)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x)
    L = _emu.lib()
    geom = model.geometry()
    outs = []
    for flags in (0, _cabi.FLAG_DSTD_MIX_FFMA):
        pk = pack_state_dict(sd, geom, "cpu")
        pk.plan_c[F["CP_FLAGS"]] = flags
        pred = torch.empty(2, geom.output_n, geom.joints, 3)
        ws = torch.empty(L.cistgcn_workspace_bytes(pk.plan_c, 2), dtype=torch.uint8)
        rc = L.cistgcn_forward_f32(pk.plan_c, len(pk.plan), pk.blob.data_ptr(), x.contiguous().data_ptr(), pred.data_ptr(),
                                   None, None, ws.data_ptr(), ws.numel(), 2, None, None)
        _cabi.check(rc, "cistgcn_forward_f32[emu]", L)
        outs.append(pred)
    mma, ffma = outs
    assert (mma - ref).abs().max().item() <= G.tol(ref)
    assert (ffma - ref).abs().max().item() <= G.tol(ref)
    assert not torch.equal(mma, ffma)
    assert (mma - ffma).abs().max().item() <= 0.5 * G.tol(ref)


@pytest.mark.parametrize("E,V", [(8, 18)])
def test_emulated_adj_mma_expansor_matches_ffma_and_oracle(E, V):
    """Stage 2's Map2Adj expansor as chained 3xTF32 mma.sync GEMMs (csrc/dstd_adj.cuh: layer 1's accumulator fragment is
    layer 2's A fragment) against the FP32-FMA column loops (CISTGCN_FLAG_DSTD_ADJ_FFMA): the sample-specific
    adjacencies of every block (N = V, T and, in the output block, 25 and V) and the prediction."""
    from cistgcn_b200 import _cabi
    from cistgcn_b200.pack import F, pack_state_dict
    model, sd, cfg = M.build(E, V, "W2")
    x, _ = O.synth_inputs(2, cfg)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x)
    L = _emu.lib()
    geom = model.geometry()
    outs = []
    for flags in (0, _cabi.FLAG_DSTD_ADJ_FFMA):
        pk = pack_state_dict(sd, geom, "cpu")
        pk.plan_c[F["CP_FLAGS"]] = flags
        pred = torch.empty(2, geom.output_n, geom.joints, 3)
        ws = torch.empty(L.cistgcn_workspace_bytes(pk.plan_c, 2), dtype=torch.uint8)
        taps_struct, holders = _cabi.make_taps(geom, 2, "cpu")
        rc = L.cistgcn_forward_f32(pk.plan_c, len(pk.plan), pk.blob.data_ptr(), x.contiguous().data_ptr(), pred.data_ptr(),
                                   None, None, ws.data_ptr(), ws.numel(), 2, taps_struct, None)
        _cabi.check(rc, "cistgcn_forward_f32[emu]", L)
        outs.append((pred, holders))
    (mma, tm), (ffma, tf) = outs
    assert (mma - ref).abs().max().item() <= G.tol(ref)
    assert (ffma - ref).abs().max().item() <= G.tol(ref)
    n_adj = 0
    for k in tm:
        if not k.endswith("Adj"):
            continue
        n_adj += 1
        a, b = tm[k], tf[k]
        assert not torch.equal(a, b), k
        assert (a - b).abs().max().item() <= 2e-5 * max(1e-30, b.abs().max().item()), k
    assert n_adj >= 4


@pytest.mark.parametrize("E,V,interp", [(8, 18, True), (8, 22, False)])
def test_emulated_narrow_mix_matches_tile_kernel_and_oracle(E, V, interp):
    """Stage 3 of the 3 -> 3 output block: the warp-per-sample kernel that streams the adjacencies from global memory
    (csrc/dstd_mix_narrow.cuh) against the tile kernel (CISTGCN_FLAG_DSTD_MIX_FFMA) and the oracle, with sample-specific
    and with static adjacencies."""
    from cistgcn_b200 import _cabi
    from cistgcn_b200.pack import F, pack_state_dict
    model, sd, cfg = M.build(E, V, "W2", interp=interp)
    x, _ = O.synth_inputs(3, cfg)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x, interpretable_in=[interp] * 5, interpretable_out=[interp])
    L = _emu.lib()
    geom = model.geometry()
    outs = []
    for flags in (0, _cabi.FLAG_DSTD_MIX_FFMA):
        pk = pack_state_dict(sd, geom, "cpu")
        pk.plan_c[F["CP_FLAGS"]] = flags
        pred = torch.empty(3, geom.output_n, geom.joints, 3)
        ws = torch.empty(L.cistgcn_workspace_bytes(pk.plan_c, 3), dtype=torch.uint8)
        rc = L.cistgcn_forward_f32(pk.plan_c, len(pk.plan), pk.blob.data_ptr(), x.contiguous().data_ptr(), pred.data_ptr(),
                                   None, None, ws.data_ptr(), ws.numel(), 3, None, None)
        _cabi.check(rc, "cistgcn_forward_f32[emu]", L)
        outs.append(pred)
    narrow, tile = outs
    assert (narrow - ref).abs().max().item() <= G.tol(ref)
    assert (tile - ref).abs().max().item() <= G.tol(ref)
    assert not torch.equal(narrow, tile)
    assert (narrow - tile).abs().max().item() <= 0.5 * G.tol(ref)


@pytest.mark.parametrize("E,V,interp", [(32, 22, True), (16, 18, False)])
def test_emulated_reduce_mma_matches_ffma_and_oracle(E, V, interp):
    """Stage 1's stacked 1x1 convolutions (Map2Adj entry maps + the gate conv (T,1)) on 3xTF32 mma.sync
    (csrc/dstd_reduce.cuh, MMA variant) against the lane-per-channel FFMA loops (CISTGCN_FLAG_DSTD_REDUCE_FFMA): gate
    vectors w1 / w2, adjacencies and the prediction; two row tiles (E = 32), one (E = 16), and blocks without Map2Adj."""
    from cistgcn_b200 import _cabi
    from cistgcn_b200.pack import F, pack_state_dict
    model, sd, cfg = M.build(E, V, "W2", interp=interp)
    nb = 1 if E >= 32 else 2
    x, _ = O.synth_inputs(nb, cfg)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x, interpretable_in=[interp] * 5, interpretable_out=[interp])
    L = _emu.lib()
    geom = model.geometry()
    outs = []
    for flags in (0, _cabi.FLAG_DSTD_REDUCE_FFMA):
        pk = pack_state_dict(sd, geom, "cpu")
        pk.plan_c[F["CP_FLAGS"]] = flags
        pred = torch.empty(nb, geom.output_n, geom.joints, 3)
        ws = torch.empty(L.cistgcn_workspace_bytes(pk.plan_c, nb), dtype=torch.uint8)
        taps_struct, holders = _cabi.make_taps(geom, nb, "cpu")
        rc = L.cistgcn_forward_f32(pk.plan_c, len(pk.plan), pk.blob.data_ptr(), x.contiguous().data_ptr(), pred.data_ptr(),
                                   None, None, ws.data_ptr(), ws.numel(), nb, taps_struct, None)
        _cabi.check(rc, "cistgcn_forward_f32[emu]", L)
        outs.append((pred, holders))
    (mma, tm), (ffma, tf) = outs
    assert (mma - ref).abs().max().item() <= G.tol(ref)
    assert (ffma - ref).abs().max().item() <= G.tol(ref)
    assert not torch.equal(mma, ffma)
    for k in tm:
        if k.startswith("st_gcnns.") and (k.endswith(".w1") or k.endswith(".w2") or (interp and k.endswith("Adj"))):
            a, b = tm[k], tf[k]
            assert (a - b).abs().max().item() <= 5e-5 * max(1e-30, b.abs().max().item()), k
