#!/bin/bash
# Multi-GPU measurements of round 2 (gpurun --gpus N): the driver's launch line for N > 1.
# usage: bash profiles/run_r2_multi.sh <N> <tag> [sections: fwd mpjpe train trainbig]
N=$1; TAG=$2; shift; shift
SECTIONS=${@:-fwd mpjpe train}
O=gpurun_out; mkdir -p $O
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@"; }
if has fwd;   then run --steps 10 --warmup 3 > $O/bench_${TAG}_fwd_e32_n$N.json 2> $O/bench_${TAG}_fwd_e32_n$N.err; echo "fwd rc=$?"; fi
if has mpjpe; then run --mode mpjpe --embed 64 --joints 18 --batch 262144 --scaling strong --steps 3 > $O/bench_${TAG}_mpjpe_e64v18_n$N.json 2> $O/bench_${TAG}_mpjpe_e64v18_n$N.err; echo "mpjpe rc=$?"; fi
if has train; then NCCL_DEBUG=INFO run --mode train --batch 128 --steps 20 > $O/bench_${TAG}_train_b128_n$N.json 2> $O/bench_${TAG}_train_b128_n$N.err; echo "train rc=$?"; fi
if has trainbig; then run --mode train --batch 4096 --steps 10 > $O/bench_${TAG}_train_b4096_n$N.json 2> $O/bench_${TAG}_train_b4096_n$N.err; echo "trainbig rc=$?"; fi
for f in $O/bench_${TAG}_*_n$N.json; do echo "$f: $(head -c 300 $f)"; done
