"""FPN stack: tensor-core kernel (fpn_tc.cuh) vs FP32-FMA kernel (fpn_chain.cuh) through cistgcn_fpn_chain_f32.
Prints the max-abs difference on x7 and the CUDA-event time of both at a bench-sized batch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _models as M  # noqa: E402
from cistgcn_b200 import _cabi  # noqa: E402

L = _cabi.lib()
dev = "cuda:0"


def run(model, x, path):
    pk = model.pack()
    B, V = x.shape[0], x.shape[-1]
    x7 = torch.full((B, 25, V, 3), float("nan"), device=dev)
    _cabi.check(L.cistgcn_set_fpn_path(path), "set_fpn_path", L)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.cistgcn_fpn_chain_f32(pk.fpn_descs(), model.n_txcnn_layers, pk.tail_desc(), pk.blob.data_ptr(),
                                 x.data_ptr(), x7.data_ptr(), B, st)
    _cabi.check(rc, "fpn_chain", L)
    torch.cuda.synchronize()
    return x7


for V, W, scale in ((22, "W2", 1.0), (22, "W1", 1.0), (18, "W2", 1.0), (22, "W2", 300.0)):
    model, sd, cfg = M.build(8, V, W)
    model = model.to(dev)
    for B in (1, 3, 300):
        g = torch.Generator().manual_seed(B)
        x = (torch.randn(B, 10, 10, V, generator=g) * scale).to(dev)
        ref = run(model, x, 1)
        got = run(model, x, 0)
        err = (got - ref).abs().max().item()
        print(f"V={V} {W} scale={scale} B={B}: |x7|max {ref.abs().max().item():.4g}  max-abs diff tc vs ffma {err:.3e}  finite {bool(torch.isfinite(got).all())}", flush=True)

model, sd, cfg = M.build(8, 22, "W2")
model = model.to(dev)
B = 65536
x = torch.randn(B, 10, 10, 22, device=dev)
for path in (1, 0):
    run(model, x, path)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run(model, x, path)
    e1.record()
    torch.cuda.synchronize()
    print(f"path {path} ({'tensor-core' if path == 0 else 'fp32-fma'}): {e0.elapsed_time(e1) / 3:.2f} ms per {B} samples", flush=True)

# where the cycles go (CTA 0): MMA-issuer and epilogue wait counters
clk = torch.zeros(32, dtype=torch.int64, device=dev)
L.cistgcn_debug_phase_clocks(clk.data_ptr())
xs = x[: 148 * 16].contiguous()
run(model, xs, 0)
L.cistgcn_debug_phase_clocks(None)
c = clk.cpu().tolist()
n = max(c[6], 1)
print(f"MMA issuer, {n} samples: total {c[0] / n:.0f} cyc/sample; waits: input {c[1] / n:.0f}, map ready {c[2] / n:.0f}, "
      f"accumulator free {c[3] / n:.0f}, weights {c[4] / n:.0f}, staging {c[5] / n:.0f}")
for name, o in (("epilogue tile 0", 8), ("epilogue tile 1", 16)):
    print(f"{name}: total {c[o] / n:.0f} cyc/sample; waits: input {c[o + 1] / n:.0f}, conv accumulator {c[o + 2] / n:.0f}, "
          f"staging free {c[o + 3] / n:.0f}, compress accumulator {c[o + 4] / n:.0f}, group barrier {c[o + 5] / n:.0f}")
print(f"epilogue thread 128: branch epilogues {c[25] / n / 12:.0f} cyc each (TMEM loads {c[24] / n / 12:.0f}), compress epilogue to map-ready {c[26] / n / 3:.0f} cyc each")
