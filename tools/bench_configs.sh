#!/bin/bash
# Driver-grade lines for every BASELINE.json configuration (C1..C5) and the C5 batch sweep, one GPU:
#   gpurun -- 'bash tools/bench_configs.sh <outtag> [configs|sweep|all]'
TAG=${1:-r02}; WHAT=${2:-all}
OUT=gpurun_out/${TAG}_bench_configs.jsonl
SWEEP=gpurun_out/${TAG}_bench_c5_sweep.jsonl
mkdir -p gpurun_out
run() { timeout 600 python bench.py --no-cpu-baseline --e2e-batch 4194304 --e2e-steps 3 "$@" 2>> gpurun_out/${TAG}_bench_configs.err; }
if [ "$WHAT" != "sweep" ]; then
: > $OUT
run --workload C1 --steps 200 --warmup 5 >> $OUT
run --workload C1 --search linear --steps 200 --warmup 5 >> $OUT
run --workload C2 --steps 50 --warmup 5 >> $OUT
run --workload C2 --search linear --steps 3 --warmup 3 --total 1000000 >> $OUT
run --workload C3 --steps 8 --warmup 3 >> $OUT
run --workload C3 --method nearest --steps 8 --warmup 3 >> $OUT
run --workload C4 --method nearest --steps 8 --warmup 3 >> $OUT
run --workload C4 --mix deep --steps 8 --warmup 3 >> $OUT
run --workload C5 --steps 8 --warmup 3 >> $OUT
run --workload C5 --method nearest --steps 8 --warmup 3 >> $OUT
run --workload C5 --scale 20 --steps 8 --warmup 3 >> $OUT
fi
if [ "$WHAT" != "configs" ]; then
: > $SWEEP
for n in 1000 10000 100000 1000000 10000000 100000000 1000000000; do
  steps=$(( n < 1000000 ? 500 : (n < 100000000 ? 50 : 5) ))
  run --workload C5 --total $n --steps $steps --warmup 5 >> $SWEEP
  run --workload C5 --method nearest --total $n --steps $steps --warmup 5 >> $SWEEP
done
fi
wc -l $OUT $SWEEP 2>/dev/null
