"""Per-block, per-stage times of the DSTD-GC path (CUDA events around every launch, cistgcn_profile_*).
usage: python profiles/dstd_stage_times.py [embed] [joints] [batch] [only block: in0..in4 | out0]
Also the ncu target for a single block (give the block name)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cistgcn_b200 import CISTGCN, _cabi  # noqa: E402
from cistgcn_b200.pack import F  # noqa: E402
from cistgcn_b200.synth import make_opt  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
V = int(sys.argv[2]) if len(sys.argv) > 2 else 22
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
only = sys.argv[4] if len(sys.argv) > 4 else None
L = _cabi.lib()
dev = "cuda:0"
opt = make_opt(E, V)
torch.manual_seed(0)
model = CISTGCN(opt.architecture_config, opt.learning_config).eval().to(dev)
pk = model.pack()
NK = _cabi.PROFILE_KINDS
names = [L.cistgcn_profile_kind_name(i).decode() for i in range(NK)]
stream = torch.cuda.current_stream().cuda_stream
blocks = [("in", i) for i in range(5)] + [("out", 0)]
for which, i in blocks:
    tag = f"{which}{i}"
    if only and tag != only:
        continue
    d = pk.block_desc(which, i)
    ci, co, T, Vb = d[F["CB_CI"]], d[F["CB_CO"]], d[F["CB_T"]], d[F["CB_V"]]
    raw = d[F["CB_IN_MODE"]] == 1
    x = torch.randn(B, 10, V, 3, device=dev) if raw else torch.randn(B, ci * T * Vb, device=dev)
    out = torch.empty(B, co * T * Vb, device=dev)
    for flags, label in ((0, "split"), (_cabi.FLAG_DSTD_MIX_FFMA | _cabi.FLAG_DSTD_ADJ_FFMA | _cabi.FLAG_DSTD_REDUCE_FFMA, "split, all FFMA"), (_cabi.FLAG_DSTD_FUSED, "fused")):
        if only and flags:
            continue
        nb = L.cistgcn_dstd_block_workspace_bytes(d, B) if not (flags & _cabi.FLAG_DSTD_FUSED) else 0
        ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
        def run():
            rc = L.cistgcn_dstd_block_f32(d, pk.blob.data_ptr(), x.data_ptr(), out.data_ptr(), B, None,
                                          ws.data_ptr(), ws.numel(), flags, stream)
            _cabi.check(rc, "dstd_block", L)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        L.cistgcn_profile_enable(1)
        reps = 3
        for _ in range(reps):
            run()
        kms = (ctypes.c_double * NK)()
        kln = (ctypes.c_int64 * NK)()
        L.cistgcn_profile_read(kms, kln)
        L.cistgcn_profile_enable(0)
        parts = {names[k]: kms[k] / reps for k in range(NK) if kln[k]}
        tot = sum(parts.values())
        cyc = tot * 1e-3 * 1.9e9 * 148 / B
        print(f"{tag} ({ci}->{co}, T={T}, V={Vb}) {label:5s}: total {tot:7.3f} ms  ({cyc / 1e3:6.1f} K cycles/sample/SM)  " +
              "  ".join(f"{k.replace('dstd_', '').replace('_kernel', '')} {v:.3f}" for k, v in parts.items()), flush=True)
