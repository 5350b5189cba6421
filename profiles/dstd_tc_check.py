"""DSTD-GC block: tensor-core channel mixes (tc_gemm in dstd_block.cuh) vs the FP32-FMA loops, through
cistgcn_dstd_block_f32.  Prints max-abs differences per block and CUDA-event times at a bench-sized batch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _models as M  # noqa: E402
from cistgcn_b200 import _cabi  # noqa: E402
from cistgcn_b200.pack import F  # noqa: E402

L = _cabi.lib()
dev = "cuda:0"


def run(pk, i, x, path, out_numel):
    B = x.shape[0]
    out = torch.full((B, out_numel), float("nan"), device=dev)
    _cabi.check(L.cistgcn_set_dstd_path(path), "set_dstd_path", L)
    rc = L.cistgcn_dstd_block_f32(pk.block_desc("in", i), pk.blob.data_ptr(), x.data_ptr(), out.data_ptr(), B, None,
                                  torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "dstd_block", L)
    torch.cuda.synchronize()
    return out


for E, V, W in ((32, 22, "W2"), (32, 22, "W1"), (32, 18, "W2")):
    model, sd, cfg = M.build(E, V, W)
    model = model.to(dev)
    pk = model.pack()
    for i in (1, 3, 4):
        d = list(pk.block_desc("in", i))
        ci, co = d[F["CB_CI"]], d[F["CB_CO"]]
        for B in (1, 300):
            g = torch.Generator().manual_seed(B + i)
            x = torch.randn(B, ci, 10, V, generator=g).to(dev)
            ref = run(pk, i, x, 0, co * 10 * V)
            got = run(pk, i, x, 1, co * 10 * V)
            print(f"E={E} V={V} {W} block in{i} ({ci}->{co}) B={B}: |out|max {ref.abs().max().item():.4g}  "
                  f"max-abs diff tc vs ffma {(got - ref).abs().max().item():.3e}  finite {bool(torch.isfinite(got).all())}", flush=True)

model, sd, cfg = M.build(32, 22, "W1")
model = model.to(dev)
pk = model.pack()
B = 32768
for i in (1, 4):
    d = list(pk.block_desc("in", i))
    ci, co = d[F["CB_CI"]], d[F["CB_CO"]]
    x = torch.randn(B, ci, 10, 22, device=dev)
    for path in (0, 1):
        run(pk, i, x, path, co * 220)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run(pk, i, x, path, co * 220)
        e1.record()
        torch.cuda.synchronize()
        print(f"block in{i} path {path} ({'tensor-core' if path == 1 else 'fp32-fma'}): {e0.elapsed_time(e1) / 3:.2f} ms per {B} samples", flush=True)
L.cistgcn_set_dstd_path(0)
