"""Runs one DSTD-GC block kernel (input block `i` of the E=32 model) on a bench-shaped batch: the ncu target.
usage: python profiles/dstd_only.py [block index] [batch] [dstd path]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _models as M  # noqa: E402
from cistgcn_b200 import _cabi  # noqa: E402
from cistgcn_b200.pack import F  # noqa: E402

i = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = _cabi.lib()
dev = "cuda:0"
model, sd, cfg = M.build(32, 22, "W1")
model = model.to(dev)
pk = model.pack()
d = pk.block_desc("in", i)
ci, co = d[F["CB_CI"]], d[F["CB_CO"]]
x = torch.randn(B, ci, 10, 22, device=dev)
out = torch.empty(B, co * 220, device=dev)
_cabi.check(L.cistgcn_set_dstd_path(path), "set_dstd_path", L)
for _ in range(3):
    rc = L.cistgcn_dstd_block_f32(d, pk.blob.data_ptr(), x.data_ptr(), out.data_ptr(), B, None, torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "dstd_block", L)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
