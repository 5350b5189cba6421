"""Runs only the FPN stack kernel (tensor-core path by default) on a bench-shaped batch: the ncu target.
usage: python profiles/fpn_only.py [batch] [path]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _models as M  # noqa: E402
from cistgcn_b200 import _cabi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
path = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = _cabi.lib()
dev = "cuda:0"
model, sd, cfg = M.build(32, 22, "W1")
model = model.to(dev)
pk = model.pack()
x = torch.randn(B, 10, 10, 22, device=dev)
x7 = torch.empty(B, 25, 22, 3, device=dev)
_cabi.check(L.cistgcn_set_fpn_path(path), "set_fpn_path", L)
for _ in range(3):
    rc = L.cistgcn_fpn_chain_f32(pk.fpn_descs(), 4, pk.tail_desc(), pk.blob.data_ptr(), x.data_ptr(), x7.data_ptr(), B,
                                 torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "fpn_chain", L)
torch.cuda.synchronize()
print("ok", float(x7.abs().max()))
