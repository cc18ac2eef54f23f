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
run --workload C5 --sweep-totals 1e3,1e4,1e5,1e6,1e7,1e8,1e9 --steps 500 --warmup 5 >> $SWEEP
run --workload C5 --method nearest --sweep-totals 1e3,1e4,1e5,1e6,1e7,1e8,1e9 --steps 500 --warmup 5 >> $SWEEP
fi
wc -l $OUT $SWEEP 2>/dev/null
