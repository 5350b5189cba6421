"""Tensor-core FPN stack (csrc/fpn_tc.cuh): the bf16 three-term operand images built by pack.py (CPU), a CPU
restatement of the kernel's six-product arithmetic against the oracle (CPU), and both FPN kernels through
cistgcn_fpn_chain_f32 against the oracle's fpn_stack on a B200 (GPU).

Tolerance on x7 (fp32): max-abs <= 1e-4 * max(1, |ref|_inf / 4), the same rule as the predicted coordinates."""
import pytest
import torch
import torch.nn.functional as Fn

import _models as M
from cistgcn_b200 import pack as P
from oracle import cistgcn_oracle as O


def _tol(ref):
    return 1e-4 * max(1.0, ref.abs().max().item() / 4)


def test_bf16_split3_reconstructs_fp32():
    g = torch.Generator().manual_seed(3)
    w = torch.randn(4096, generator=g, dtype=torch.float64) * torch.logspace(-6, 3, 4096, dtype=torch.float64)
    b1, b2, b3 = P.bf16_split3(w)
    w32 = w.to(torch.float32).double()
    rec = b1.double() + b2.double() + b3.double()
    assert ((rec - w32).abs() <= w32.abs() * 2.0 ** -22).all()
    assert (b2.double().abs() <= b1.double().abs() * 2.0 ** -7).all()


def test_tc_slice_layout_round_trip():
    g = torch.Generator().manual_seed(4)
    W = torch.randn(25, 10, generator=g, dtype=torch.float64)
    words = P.tc_slice(W, 2)
    assert words.dtype == torch.int32 and words.numel() == 2 * 128 * 4
    raw = words.view(torch.int16).reshape(2, 128, 8)                     # [k-chunk][row][8 channels]
    img = raw[:, :96].contiguous().view(torch.bfloat16).double()         # rows 32 * term + cout: bf16 terms
    rec = torch.zeros(32, 16, dtype=torch.float64)
    for term in range(3):
        rec += img[:, 32 * term: 32 * term + 32, :].permute(1, 0, 2).reshape(32, 16)
    assert torch.allclose(rec[:25, :10], W.float().double(), rtol=2.0 ** -22, atol=0)
    assert rec[25:].abs().max() == 0 and rec[:, 10:].abs().max() == 0
    wh = raw[:, 96:].contiguous().view(torch.float16).double().permute(1, 0, 2).reshape(32, 16)   # rows 96 + cout: fp16
    assert torch.equal(wh[:25, :10], W.float().half().double()) and wh[25:].abs().max() == 0


def test_packed_plan_has_tensor_core_image():
    model, sd, cfg = M.build(8, 22, "W2")
    pk = P.pack_state_dict(sd, model.geometry(), "cpu")
    descs = list(pk.fpn_descs())
    n = P.F["CF_COUNT"]
    for l in range(4):
        f = descs[l * n: (l + 1) * n]
        assert f[P.F["CF_TC_KC"]] == (2 if l == 0 else 4)
        assert f[P.F["CF_TC_W"]] % 4 == 0 and f[P.F["CF_TC_W"]] > 0          # 16-byte aligned for cp.async.bulk
        prm = pk.blob[f[P.F["CF_TC_PRM"]]: f[P.F["CF_TC_PRM"]] + P.TC_PRM_FLOATS]
        assert torch.isfinite(prm).all()
        bias = sd[f"txcnns.{l}.compress.bias"]
        assert torch.allclose(prm[P.TC_PRM_CPB: P.TC_PRM_CPB + 25], bias.float())
    # raw bf16 words survive the float32 blob bit-exactly
    off = descs[P.F["CF_TC_W"]]
    hs = P.host_state(sd)                                  # fp64 host copies, as the packer sees them
    s, b = P._fold(hs, "txcnns.0.block1.1")
    W = s[:, None, None, None] * hs["txcnns.0.block1.0.weight"]
    want = P.tc_slice(W[:, :, 0, 0], 2)
    got = pk.blob[off: off + want.numel()].view(torch.int32)
    assert torch.equal(got, want)


def _three_term_conv(x, W, dil):
    """What the kernel computes for one dilated 3x3 branch: x = x1 (bf16) + x2 (fp16 remainder);
    x1 * (w1 + w2 + w3) with three bf16 weight terms, x2 * fp16(w); fp32 accumulation."""
    x1 = x.bfloat16().float()
    x2 = (x - x1).clamp(-65504, 65504).half().float()
    ws = [t.float() for t in P.bf16_split3(W)]
    out = Fn.conv2d(x1, ws[2], None, padding=dil, dilation=dil) + Fn.conv2d(x1, ws[1], None, padding=dil, dilation=dil)
    return out + (Fn.conv2d(x1, ws[0], None, padding=dil, dilation=dil) + Fn.conv2d(x2, W.half().float(), None, padding=dil, dilation=dil))


def test_three_term_bf16_arithmetic_matches_fp32_conv():
    """Pins the numerical scheme on CPU: the split operands reproduce the fp32 convolution to ~1e-6
    relative, far inside the 1e-4 budget (a single bf16 product would be off by ~4e-3)."""
    model, sd, cfg = M.build(8, 22, "W2")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 25, 10, 22, generator=g)
    for i in (1, 2, 3):
        W = sd[f"txcnns.1.block{i}.0.weight"]
        ref = Fn.conv2d(x.double(), W.double(), None, padding=i, dilation=i)
        got = _three_term_conv(x, W, i)
        one = Fn.conv2d(x.bfloat16().float(), W.bfloat16().float(), None, padding=i, dilation=i)
        scale = ref.abs().max().item()
        assert (got.double() - ref).abs().max().item() <= 4e-6 * scale
        assert (one.double() - ref).abs().max().item() > 1e-4 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("V,weights,scale", [(22, "W1", 1.0), (22, "W2", 1.0), (18, "W2", 1.0), (22, "W2", 300.0), (18, "W1", 300.0)])
@pytest.mark.parametrize("B", [1, 149, 600])
def test_fpn_kernels_match_oracle(V, weights, scale, B):
    from cistgcn_b200 import _cabi
    L = _cabi.lib()
    dev = "cuda:0"
    model, sd, cfg = M.build(8, V, weights)
    model = model.to(dev)
    pk = model.pack()
    g = torch.Generator().manual_seed(B + V)
    x5 = torch.randn(B, 10, 10, V, generator=g) * scale
    ref = O.fpn_stack_eval(sd, cfg, x5)
    xd = x5.to(dev)
    outs = {}
    for path, flags in ((0, 0), (1, _cabi.FLAG_FPN_FP32)):      # 0: tcgen05 kernel, 1: FP32-FMA kernel (per-call flag)
        x7 = torch.full((B, 25, V, 3), float("nan"), device=dev)
        rc = L.cistgcn_fpn_chain_f32(pk.fpn_descs(), 4, pk.tail_desc(), pk.blob.data_ptr(), xd.data_ptr(),
                                     x7.data_ptr(), B, flags, torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "cistgcn_fpn_chain_f32", L)
        torch.cuda.synchronize()
        outs[path] = x7.cpu()
    for path, x7 in outs.items():
        assert torch.isfinite(x7).all(), path
        assert (x7 - ref).abs().max().item() <= _tol(ref), path
    assert (outs[0] - outs[1]).abs().max().item() <= 0.25 * _tol(ref)


@pytest.mark.gpu
def test_fpn_chain_rejects_inconsistent_descriptors():
    """Descriptor consistency is checked before EITHER kernel is dispatched (a direct C-ABI caller with mismatched
    channel counts must get an error, not an out-of-range shared-memory access in the tensor-core kernel)."""
    from cistgcn_b200 import _cabi
    from cistgcn_b200.pack import F
    L = _cabi.lib()
    model, sd, cfg = M.build(8, 22, "W1")
    pk = model.to("cuda:0").pack()
    descs = pk.fpn_descs()
    descs[F["CF_CIN"]] = 7                                       # layer 0 claims 7 input frames, the tail says 10
    x5 = torch.zeros(2, 10, 10, 22, device="cuda:0")
    x7 = torch.zeros(2, 25, 22, 3, device="cuda:0")
    rc = L.cistgcn_fpn_chain_f32(descs, 4, pk.tail_desc(), pk.blob.data_ptr(), x5.data_ptr(), x7.data_ptr(), 2, 0, None)
    assert rc < 0 and b"mismatch" in L.cistgcn_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("E,V", [(32, 22), (32, 18)])
@pytest.mark.parametrize("weights,scale", [("W1", "unit"), ("W2", "unit"), ("W1", "mm")])
def test_dstd_tensor_core_channel_mixes_match_oracle(E, V, weights, scale):
    """CISTGCN_FLAG_DSTD_TC: the round-1 fused kernel with its 1x1 channel mixes (Map2Adj entry convs, tcn, compressor,
    residual conv) as tcgen05 MMAs; the whole forward must stay inside the fp32 tolerance and agree with the default
    (three-stage FP32) path to a fraction of it."""
    import _golden as G
    from cistgcn_b200 import _cabi
    dev = "cuda:0"
    model, sd, cfg = M.build(E, V, weights)
    x, _ = O.synth_inputs(48, cfg, scale=scale)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x)
    model = model.to(dev)
    model.kernel_flags = _cabi.FLAG_DSTD_FUSED | _cabi.FLAG_DSTD_TC
    tc = model(x.to(dev))[0].cpu()
    model.kernel_flags = _cabi.FLAG_DSTD_FUSED
    fused = model(x.to(dev))[0].cpu()
    model.kernel_flags = 0
    fm = model(x.to(dev))[0].cpu()
    assert torch.isfinite(tc).all()
    for name, out in (("tc", tc), ("fused", fused), ("split", fm)):
        assert (out - ref).abs().max().item() <= G.tol(ref), name
    assert not torch.equal(tc, fused)                    # the tensor-core path really ran
    assert (tc - fm).abs().max().item() <= 0.5 * G.tol(ref)
    assert (fused - fm).abs().max().item() <= 0.5 * G.tol(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("E,V", [(32, 22), (32, 18), (16, 22), (16, 18)])
@pytest.mark.parametrize("weights,scale", [("W1", "unit"), ("W2", "unit"), ("W1", "mm")])
def test_dstd_mix_mma_channel_mixes_match_oracle(E, V, weights, scale):
    """Default three-stage path: the channel mixes of stage 3 (tcn of both domains, domain residual conv, compressor,
    block residual conv) run as 3xTF32 mma.sync (csrc/dstd_mix_mma.cuh).  The forward must stay inside the fp32
    tolerance, agree with the FP32-FMA tile loops (CISTGCN_FLAG_DSTD_MIX_FFMA) to a fraction of it, and differ from them
    in the last bits (i.e. the tensor-core kernel really ran)."""
    import _golden as G
    from cistgcn_b200 import _cabi
    dev = "cuda:0"
    model, sd, cfg = M.build(E, V, weights)
    x, _ = O.synth_inputs(48, cfg, scale=scale)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x)
    model = model.to(dev)
    model.kernel_flags = 0
    mma = model(x.to(dev))[0].cpu()
    model.kernel_flags = _cabi.FLAG_DSTD_MIX_FFMA
    ffma = model(x.to(dev))[0].cpu()
    model.kernel_flags = 0
    assert torch.isfinite(mma).all()
    assert (mma - ref).abs().max().item() <= G.tol(ref)
    assert (ffma - ref).abs().max().item() <= G.tol(ref)
    assert not torch.equal(mma, ffma)
    assert (mma - ffma).abs().max().item() <= 0.5 * G.tol(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("E,V", [(32, 22), (16, 18), (8, 22), (64, 18)])
@pytest.mark.parametrize("weights,scale", [("W1", "unit"), ("W2", "unit"), ("W1", "mm")])
@pytest.mark.parametrize("which", ["adj", "reduce", "narrow"])
def test_dstd_stage_variants_match_oracle_and_each_other(E, V, weights, scale, which):
    """Three-stage path, one stage at a time switched back to its FP32-FMA / tile implementation:
      adj     Map2Adj expansor as chained 3xTF32 mma.sync GEMMs (csrc/dstd_adj.cuh) vs the FFMA column loops
      reduce  stacked 1x1 convolutions of stage 1 on 3xTF32 mma.sync (csrc/dstd_reduce.cuh, Ci >= 16) vs lane-per-channel FFMA
      narrow  3 -> 3 output block: warp-per-sample streaming kernel (csrc/dstd_mix_narrow.cuh) vs the tile kernel
    Both forwards stay inside the fp32 tolerance against the oracle, agree to a fraction of it, and (where the variant is
    reachable at this width) differ in the last bits, i.e. the other kernel really ran.  The adjacency taps are compared
    relatively (they are ~1e-9 at default init).  E = 64 blocks run on the fused kernel except the output block."""
    import _golden as G
    from cistgcn_b200 import _cabi
    dev = "cuda:0"
    flag = {"adj": _cabi.FLAG_DSTD_ADJ_FFMA, "reduce": _cabi.FLAG_DSTD_REDUCE_FFMA, "narrow": _cabi.FLAG_DSTD_MIX_FFMA}[which]
    model, sd, cfg = M.build(E, V, weights)
    x, _ = O.synth_inputs(40, cfg, scale=scale)
    with torch.no_grad():
        ref = O.forward(sd, cfg, x)
    model = model.to(dev)
    outs = []
    for flags in (0, flag):
        model.kernel_flags = flags
        model.enable_taps(True)
        pred = model(x.to(dev))[0].cpu()
        taps = {k: v.cpu().clone() for k, v in model.last_taps.items() if k.endswith("Adj")}
        model.enable_taps(False)
        outs.append((pred, taps))
    model.kernel_flags = 0
    (new, tn), (old, to) = outs
    assert torch.isfinite(new).all()
    assert (new - ref).abs().max().item() <= G.tol(ref)
    assert (old - ref).abs().max().item() <= G.tol(ref)
    assert (new - old).abs().max().item() <= 0.5 * G.tol(ref)
    reachable = {"adj": True, "narrow": True, "reduce": E in (16, 32)}[which]      # reduce MMA needs 16 <= Ci <= 32
    if reachable and weights == "W2":               # (at default init the adjacency products vanish in fp32: only W2 can tell)
        assert not torch.equal(new, old)
    if which == "adj":
        assert any(not torch.equal(tn[k], to[k]) for k in tn)
    for k in tn:
        a, b = tn[k], to[k]
        assert (a - b).abs().max().item() <= 1e-4 * max(1e-30, b.abs().max().item()), k
