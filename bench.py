#!/usr/bin/env python
"""bench.py -- sequences/sec of the CIST-GCN forward hot path (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--embed 32 --joints 22 --batch 65536]
    python bench.py --impl reference ...      # the reference's own CPU implementation (oracle port)

A "step" is one forward pass over one batch of synthetic H36M-shaped sequences (configs[1] of
BASELINE.json: E=32, 22 joints, 10 -> 25 frames, batch 65536 per GPU, fp32).  N > 1: launched by
torchrun, one rank per GPU, the batch is sharded by rank ("weak": per-GPU work is fixed) and no
collective is on the data path; only the timing is max-reduced over ranks.  Prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sequences/sec CIST-GCN forward (H36M 10->25 frames, 22 joints)"
UNIT = "sequences/s"
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # CUDA-core FP32 FMA peak at max SM clock (74.5)
# algorithmic FLOPs / sequence (2*MAC of conv+mm+bmm, measured on the reference)
FLOPS_PER_SEQ = {(22, 8): 37.06e6, (22, 16): 40.45e6, (22, 32): 52.22e6, (22, 64): 95.68e6,
                 (18, 8): 29.59e6, (18, 16): 32.32e6, (18, 32): 41.87e6, (18, 64): 77.26e6}
FPN_FLOPS_PER_SEQ = {22: 29.5e6, 18: 29.5e6 * 18 / 22}


def bytes_per_seq(V, E, s=4):
    """one-kernel-per-block compulsory HBM traffic per sequence."""
    return s * V * (80 * E + 2683)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense bf16 TFLOP/s, sustained figure (the FPN kernel is timed inside a long step)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        for k in ("bf16_tflops_sustained", "bf16_tflops"):
            if k in d:
                return float(d[k]), f"measured {k} (MEASURED_PEAKS.json)"
    return 1400.0, "fallback sustained (B200_PROFILING.md)"


def fpn_tc_issued_flops_per_seq():
    """Tensor FLOPs the tcgen05 FPN kernel ISSUES per sequence (csrc/fpn_tc.cuh): per k-step and 128-position tile one
    M128 N96 K16 bf16 MMA + one M128 N32 K16 fp16 MMA; 2 tiles; layer 0 has 1 k-step per tap, layers 1-3 have 2;
    27 taps + 3 compress slices (2 k-steps) per layer."""
    ksteps_tiles = 2 * ((27 * 1 + 3 * 2) + 3 * (27 * 2 + 3 * 2))
    return ksteps_tiles * 2.0 * 128 * (96 + 32) * 16


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through `nvidia-smi -lms 200` while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) == 6:
                    self.rows.append(parts)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_baseline(E, V, budget_s, batch=256):
    """Oracle port (oracle/cistgcn_oracle.py = the reference's algorithm on PyTorch CPU kernels) timed on
    this box's host cores on a bounded sample of the same workload."""
    import torch
    import types
    from cistgcn_b200 import CISTGCN
    from oracle import cistgcn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = _make_model(E, V)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(joints=V, input_gcn=[E] * 4)
    x, _ = O.synth_inputs(batch, cfg)
    with torch.no_grad():
        for _ in range(2):
            O.forward(sd, cfg, x)
        times, t_end = [], time.time() + budget_s
        while len(times) < 3 or (time.time() < t_end and len(times) < 200):
            t0 = time.perf_counter()
            O.forward(sd, cfg, x)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": batch / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} forwards of batch {batch} (E={E}, V={V}, fp32, eval), median; "
                      f"{torch.get_num_threads()} torch threads"}


def _make_model(E, V):
    import types
    import torch
    from cistgcn_b200 import CISTGCN
    ns = types.SimpleNamespace
    mp = ns(input_n=10, output_n=25, joints=V, n_txcnn_layers=4, txc_kernel_size=3, reduction=8, hidden_dim=64,
            input_gcn=ns(model_complexity=[E] * 4, interpretable=[True] * 5),
            output_gcn=ns(model_complexity=[3], interpretable=[True]), clipping=15)
    torch.manual_seed(0)
    return CISTGCN(ns(model_params=mp), ns(dropout=0.1)).eval()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python
    reference package cannot travel to the GPU box).  Rank 0 alone works; other ranks exit 0."""
    if rank != 0:
        return
    import torch
    from oracle import cistgcn_oracle as O
    E, V = args.embed, args.joints
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = _make_model(E, V)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(joints=V, input_gcn=[E] * 4)
    step_seqs, sub = 1024, 256                       # one step = a bounded 1024-sequence sample, eval batch 256
    x, _ = O.synth_inputs(step_seqs, cfg)
    def step():
        with torch.no_grad():
            for i in range(0, step_seqs, sub):
                O.forward(sd, cfg, x[i:i + sub])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = step_seqs * args.steps / dt
    sample = f"{args.steps} steps x {step_seqs} sequences (batches of {sub}), E={E}, V={V}, fp32 eval, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CISTGCN embed={E} forward, H36M shape (10 in / 25 out, {V} joints), fp32; "
                               f"CPU sample of {step_seqs} sequences per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_native(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from cistgcn_b200 import _cabi
    from oracle import cistgcn_oracle as O

    E, V, B = args.embed, args.joints, args.batch
    # library chatter on stdout (e.g. the "NCCL version" banner) goes to stderr: stdout carries the ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()                                   # fails loudly if the extension is missing
    _cabi.check(lib.cistgcn_set_fpn_path(int(os.environ.get("CISTGCN_BENCH_FPN_PATH", "0"))), "cistgcn_set_fpn_path", lib)
    model = _make_model(E, V).to(dev)
    cfg = O.OracleConfig(joints=V, input_gcn=[E] * 4)
    # every rank owns its own shard of the global batch (different seed per rank); no data-path collective
    x_host, _ = O.synth_inputs(B, cfg, seed=123 + rank)
    x_pin = x_host.pin_memory()
    pred_pin = torch.empty(B, 25, V, 3).pin_memory()
    x = x_pin.to(dev)
    in_bytes, out_bytes = x.numel() * 4, pred_pin.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs already in HBM (input 173 MB > 126 MB L2, so nothing is L2-warm)
    for _ in range(max(args.warmup, 3)):
        model(x)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.cistgcn_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = model(x)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    kms = (ctypes.c_double * 4)()
    kln = (ctypes.c_int64 * 4)()
    lib.cistgcn_profile_read(kms, kln)
    lib.cistgcn_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end-to-end arm: host buffers; every step copies its inputs host->device and its predictions
    # device->host inside the timed region.  The step is split into sub-batches so the copies (second stream)
    # overlap the forward of the neighbouring sub-batch -- plain stream pipelining around the public forward().
    n_sub = 4 if B % 4 == 0 and B >= 4096 else 1
    sub = B // n_sub
    copy_stream = torch.cuda.Stream(device=dev)
    compute_stream = torch.cuda.current_stream(dev)
    x_dev = [torch.empty(sub, 10, V, 3, device=dev) for _ in range(n_sub)]

    def e2e_step():
        ev_in = [torch.cuda.Event() for _ in range(n_sub)]
        ev_out = [torch.cuda.Event() for _ in range(n_sub)]
        copy_stream.wait_stream(compute_stream)              # previous step's compute no longer reads x_dev
        with torch.cuda.stream(copy_stream):
            for i in range(n_sub):
                x_dev[i].copy_(x_pin[i * sub:(i + 1) * sub], non_blocking=True)
                ev_in[i].record(copy_stream)
        for i in range(n_sub):
            compute_stream.wait_event(ev_in[i])
            pr = model(x_dev[i])[0]
            ev_out[i].record(compute_stream)
            pr.record_stream(copy_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_out[i])
                pred_pin[i * sub:(i + 1) * sub].copy_(pr, non_blocking=True)
        compute_stream.wait_stream(copy_stream)              # the step ends when its last D2H has landed

    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    checksum = float(pred_pin[:16].double().abs().sum())          # the D2H result is really read

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    hbm_peak, peak_src = measured_peaks()
    launches = sum(int(v) for v in kln)                     # launches of OUR kernels inside the timed region
    fpn_tc = V in (22, 18) and os.environ.get("CISTGCN_BENCH_FPN_PATH", "0") == "0"
    names = ["dstd_block_kernel", "fpn_tc_kernel" if fpn_tc else "fpn_chain_kernel", "tail_kernel", "mpjpe_kernel"]
    kshare = {names[i]: {"ms_per_step": kms[i] / args.steps, "launches_per_step": int(kln[i]) / args.steps}
              for i in range(4) if kln[i]}
    dom = max(range(4), key=lambda i: kms[i])
    total_kms = sum(kms)
    n_launch_dom = max(int(kln[dom]), 1)
    seqs = B * args.steps                                   # sequences pushed through every kernel kind
    # Algorithmic (compulsory) HBM bytes and FLOPs per sequence of each kernel kind -- DESIGN.md section 5:
    # every kernel reads its input activation once and writes its output once; weights are amortised.
    TV = 10 * V
    dstd_bytes = 4 * ((3 * TV + E * TV) + 3 * 2 * E * TV + (E * TV + 10 * TV) + 2 * 75 * V)   # 5 input blocks + output block
    kind_bytes = [dstd_bytes, 4 * (10 * TV + 75 * V), 4 * (30 * V + 3 * 75 * V), 4 * 2 * 75 * V]
    fl_total = FLOPS_PER_SEQ.get((V, E), 0.0)
    kind_flops = [fl_total - FPN_FLOPS_PER_SEQ[V] - 0.7e6 * V / 22, FPN_FLOPS_PER_SEQ[V], 0.7e6 * V / 22, 0.0]
    dur_s = kms[dom] / 1e3 / n_launch_dom                   # average launch duration, CUDA events on the launch stream
    bytes_launch = kind_bytes[dom] * seqs / n_launch_dom
    flops_launch = kind_flops[dom] * seqs / n_launch_dom
    achieved = bytes_launch / dur_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.isfile(tpath):                               # ncu --set full capture of the same launch shape, bytes / sequence
        tj = json.load(open(tpath)).get(f"E{E}_V{V}", {}).get(names[dom])
        if tj is not None:
            traffic = tj * seqs / n_launch_dom
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "share_of_step": kms[dom] / total_kms if total_kms else None,
                "bytes_per_launch": bytes_launch, "launch_ms": dur_s * 1e3,
                "note": "the fp32 path is bound by the FP32 FMA pipe, not HBM (SURVEY.md 0.3 / 8d): see roofline_fp32_fma"}
    tf = flops_launch / dur_s / 1e12
    roofline_fma = {"bound": "fp32_fma", "kernel": names[dom], "achieved": tf, "peak": FP32_FMA_PEAK_TFLOPS,
                    "unit": "TFLOP/s", "frac": tf / FP32_FMA_PEAK_TFLOPS,
                    "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz; measured ceiling of a pure FFMA loop on this "
                                   "pool is 92/128 of it (profiles/microbench_r1.log)",
                    "whole_forward": {"achieved": fl_total * (value / world) / 1e12,
                                      "frac": fl_total * (value / world) / 1e12 / FP32_FMA_PEAK_TFLOPS,
                                      "hbm_frac_survey_bytes": bytes_per_seq(V, E) * (value / world) / 1e9 / hbm_peak}}
    roofline_tensor = None
    if fpn_tc and kln[1]:
        tpeak, tsrc = measured_tensor_peak()
        fpn_s = kms[1] / 1e3
        alg = FPN_FLOPS_PER_SEQ[V] * seqs / fpn_s / 1e12
        issued = fpn_tc_issued_flops_per_seq() * seqs / fpn_s / 1e12
        roofline_tensor = {"bound": "tensor", "kernel": "fpn_tc_kernel", "achieved": alg, "issued": issued, "peak": tpeak,
                           "unit": "TFLOP/s", "frac": alg / tpeak, "frac_issued": issued / tpeak, "peak_source": tsrc,
                           "share_of_step": kms[1] / total_kms if total_kms else None,
                           "launch_ms": kms[1] / max(int(kln[1]), 1),
                           "note": "achieved = the reference's algorithmic FPN FLOPs (SURVEY.md 8d) / kernel time; issued = tcgen05 "
                                   "FLOPs incl. channel padding to 32 and the split-operand products that buy fp32 accuracy "
                                   "(x1*[w1|w2|w3] + x2*wh); the pipe is fed from shared memory at N <= 96, see DESIGN.md"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CISTGCN embed={E} forward, H36M shape (10 in / 25 out frames, {V} joints x 3), "
                               f"batch {B} per GPU, fp32, random-init weights (BASELINE.json configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "embed": E, "joints": V,
                   "parallelism": f"batch-sharded x{world}, no data-path collective",
                   "l2_policy": "inputs_larger_than_L2 (173 MB input + activations per step > 126 MB L2)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "ms_per_step": ms_e2e / args.steps, "api": f"CISTGCN.forward on pinned host buffers, {n_sub} sub-batches, copies overlapped on a second stream",
                "checksum": checksum},
        "gpu_launches": launches * world,     # every rank launches the same kernels on its own GPU
        "kernels": kshare,
        "roofline": roofline,
        "roofline_fp32_fma": roofline_fma,
        "roofline_tensor": roofline_tensor,
        "clocks": clocks,
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(E, V, args.cpu_budget)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--embed", type=int, default=32)
    ap.add_argument("--joints", type=int, default=22)
    ap.add_argument("--batch", type=int, default=65536, help="sequences per GPU per step")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline timing (N=1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
