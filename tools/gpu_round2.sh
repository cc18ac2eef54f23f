#!/bin/bash
# gpurun -- 'bash tools/gpu_round2.sh <tag>': GPU parity suite, then the small-step lines (C1, C2, C5 sweep up to 1e6)
# in sync and async api modes, then the default bench line.
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log
L=$OUT/${TAG}_small_steps.jsonl; : > $L
run() { timeout 600 python bench.py --no-cpu-baseline --e2e-batch 1048576 --e2e-steps 3 "$@" >> $L 2>> $OUT/${TAG}_small_steps.err; }
for mode in sync async; do
  echo "# api-mode $mode" >> $L
  run --workload C1 --steps 200 --warmup 5 --api-mode $mode
  run --workload C1 --search linear --steps 200 --warmup 5 --api-mode $mode
  run --workload C2 --steps 50 --warmup 5 --api-mode $mode
  for n in 1000 10000 100000 1000000; do
    run --workload C5 --total $n --steps 500 --warmup 5 --api-mode $mode
  done
done
python tools/show_lines.py $L
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_lines.py $OUT/${TAG}_bench.json
