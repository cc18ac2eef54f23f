#!/bin/bash
# gpurun -- 'bash tools/gpu_round4.sh <tag>': parity suite, cached-path timing, 60^5 table, ncu launch list + full capture
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log
for wl in C4 C5 C3; do timeout 300 python tools/bench_cached.py --workload $wl >> $OUT/${TAG}_cached.jsonl 2>> $OUT/${TAG}_cached.err; done; cat $OUT/${TAG}_cached.jsonl
bash tools/gpu_exp.sh $TAG "--scale 60" "--scale 60 --mix deep" "--scale 50"
bash tools/gpu_round.sh $TAG bench ncu
