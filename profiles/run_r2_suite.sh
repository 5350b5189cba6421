#!/bin/bash
# Round-2 measurement suite (run on the GPU box via gpurun; outputs under gpurun_out/, copied to profiles/ by hand).
# usage: bash profiles/run_r2_suite.sh <tag> [sections...]   sections: tests fwd e16 mpjpe train bf16 sweep stages launches
TAG=${1:-r2}; shift
SECTIONS=${@:-tests fwd train}
O=gpurun_out
mkdir -p $O
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
if has tests; then
  rm -f $O/parity_r2_*.log
  python -m pytest tests -q -m gpu > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
fi
if has fwd;   then python bench.py > $O/bench_${TAG}_fwd_e32.json 2> $O/bench_${TAG}_fwd_e32.err; echo "fwd rc=$?"; fi
if has e16;   then python bench.py --embed 16 --steps 5 > $O/bench_${TAG}_fwd_e16.json 2> $O/bench_${TAG}_fwd_e16.err; echo "e16 rc=$?"; fi
if has e8;    then python bench.py --embed 8 --batch 256 --steps 20 > $O/bench_${TAG}_fwd_e8_b256.json 2> $O/bench_${TAG}_fwd_e8_b256.err; echo "e8 rc=$?"; fi
if has mpjpe; then python bench.py --mode mpjpe --embed 64 --joints 18 --batch 262144 --scaling strong --steps 3 > $O/bench_${TAG}_mpjpe_e64v18_n1.json 2> $O/bench_${TAG}_mpjpe_e64v18_n1.err; echo "mpjpe rc=$?"; fi
if has train; then python bench.py --mode train --batch 128 --steps 20 > $O/bench_${TAG}_train_b128_n1.json 2> $O/bench_${TAG}_train_b128_n1.err; echo "train rc=$?"; fi
if has trainbig; then python bench.py --mode train --batch 4096 --steps 10 > $O/bench_${TAG}_train_b4096_n1.json 2> $O/bench_${TAG}_train_b4096_n1.err; echo "trainbig rc=$?"; fi
if has bf16;  then python bench.py --dtype bf16 --embed 64 --steps 5 > $O/bench_${TAG}_bf16_e64.json 2> $O/bench_${TAG}_bf16_e64.err; echo "bf16 rc=$?"; fi
if has sweep; then
  : > $O/bench_${TAG}_bf16_sweep.json
  for B in 1024 4096 16384 65536 262144 1048576; do
    S=10; [ $B -ge 262144 ] && S=3
    python bench.py --dtype bf16 --embed 64 --batch $B --steps $S --cpu-budget 0 >> $O/bench_${TAG}_bf16_sweep.json 2>> $O/bench_${TAG}_bf16_sweep.err; echo "sweep $B rc=$?"
  done
fi
if has stages; then python profiles/dstd_stage_times.py 32 22 32768 > $O/stage_times_$TAG.log 2>&1; echo "stages rc=$?"; fi
if has launches; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --cpu-budget 0 > $O/ncu_launches_$TAG.log 2>&1; echo "launches rc=$?"
fi
