"""Per-phase cycle breakdown of the fused DSTD-GC kernel (debug hook cistgcn_debug_phase_clocks).
usage: python profiles/phase_clocks.py [E] [V] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import _make_model  # noqa: E402
from cistgcn_b200 import _cabi  # noqa: E402
from cistgcn_b200.pack import F  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
V = int(sys.argv[2]) if len(sys.argv) > 2 else 22
B = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 4
ITER = int(sys.argv[4]) if len(sys.argv) > 4 else 0       # which sample of CTA 0 to stamp (0 = cold, >=1 = warm)
PATH = int(sys.argv[5]) if len(sys.argv) > 5 else 0       # cistgcn_set_dstd_path: 0 FP32-FMA channel mixes, 1 tensor-core
NAMES = ["load+norm", "stats", "gate conv(T,1)", "gate matvecs", "map2adj 1x1", "collapse convs", "dimseq/dsp+outer_s",
         "expansor_s", "gcn_space", "outer_t+expansor_t", "gcn_time", "tcn x2", "compressor", "SE", "store"]
lib = _cabi.lib()
lib.cistgcn_set_dstd_path(PATH)
model = _make_model(E, V).cuda()
pk = model.pack("cuda")
clk = torch.zeros(24, dtype=torch.int64, device="cuda")
lib.cistgcn_debug_phase_clocks(clk.data_ptr())
lib.cistgcn_debug_stamp_iteration(ITER)
for which, i in (("in", 0), ("in", 1), ("in", 4), ("out", 0)):
    d = pk.block_desc(which, i)
    ci, co, T, Vb = d[F["CB_CI"]], d[F["CB_CO"]], d[F["CB_T"]], d[F["CB_V"]]
    n_in = max(d[F["CB_IN_SB"]], 1)
    n_out = max(d[F["CB_OUT_SB"]], 1)
    x = torch.randn(B * n_in, device="cuda")
    y = torch.empty(B * n_out, device="cuda")
    for _ in range(2):
        clk.zero_()
        rc = lib.cistgcn_dstd_block_f32(d, pk.blob.data_ptr(), x.data_ptr(), y.data_ptr(), B, None, None)
        _cabi.check(rc, "dstd")
        torch.cuda.synchronize()
    c = clk.cpu().tolist()
    tot = c[15] - c[0]
    print(f"block {which}{i} ({ci}->{co}, T={T}, V={Vb}): {tot} cycles / sample")
    for k, n in enumerate(NAMES):
        print(f"   {n:22s} {c[k + 1] - c[k]:8d}  {100.0 * (c[k + 1] - c[k]) / tot:5.1f}%")
    if c[19]:
        ns = -(-B // 148)        # samples CTA 0 processed (counters accumulate over all of them)
        print(f"   tensor-core channel mixes, per call ({c[19] / ns:.0f} calls / sample): conversion {c[16] / c[19]:.0f}, "
              f"MMA wait {c[17] / c[19]:.0f}, epilogue {c[18] / c[19]:.0f} cycles")
lib.cistgcn_debug_phase_clocks(None)
