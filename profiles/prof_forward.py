"""Minimal forward driver for ncu captures: python profiles/prof_forward.py [E] [V] [B] [iters]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import _make_model  # noqa: E402
from oracle import cistgcn_oracle as O  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
V = int(sys.argv[2]) if len(sys.argv) > 2 else 22
B = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 8
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
model = _make_model(E, V).cuda()
x, _ = O.synth_inputs(B, O.OracleConfig(joints=V, input_gcn=[E] * 4))
x = x.cuda()
for _ in range(iters):
    out = model(x)[0]
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
