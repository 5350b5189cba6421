"""Forward latency at the reference's own batch sizes (eval 256, train 128): wall clock per call incl. host overhead."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import _make_model
from oracle import cistgcn_oracle as O

for E in (32,):
    model = _make_model(E, 22).cuda()
    cfg = O.OracleConfig(joints=22, input_gcn=[E] * 4)
    for B in (1, 128, 256, 1024, 4096):
        x, _ = O.synth_inputs(B, cfg)
        x = x.cuda()
        for _ in range(5):
            model(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            out = model(x)[0]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            out = model(x)[0]
        ev1.record()
        torch.cuda.synchronize()
        print(f"E={E} B={B:5d}: wall {dt * 1e3:7.3f} ms/call ({B / dt:9.0f} seq/s), device-timeline {ev0.elapsed_time(ev1) / n:7.3f} ms/call")
