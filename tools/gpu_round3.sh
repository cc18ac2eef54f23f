#!/bin/bash
# gpurun -- 'bash tools/gpu_round3.sh <tag>': GPU parity suite, default bench line, tuning experiments, small steps (async).
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_lines.py $OUT/${TAG}_bench.json
bash tools/gpu_exp.sh $TAG "" "--scale 60" "--scale 60 --opt block_order=2" "--spacing 0.3" "--method nearest" "--mix deep"
NDPB_LIB=$PWD/tools/bin/libndpb200_notma.so bash tools/gpu_exp.sh ${TAG}_notma "" 
L=$OUT/${TAG}_small_steps.jsonl; : > $L
run() { timeout 600 python bench.py --no-cpu-baseline --e2e-batch 1048576 --e2e-steps 3 "$@" >> $L 2>> $OUT/${TAG}_small_steps.err; }
run --workload C1 --steps 200 --warmup 5
run --workload C2 --steps 50 --warmup 5
for n in 1000 10000 100000 1000000; do run --workload C5 --total $n --steps 500 --warmup 5; done
python tools/show_lines.py $L
