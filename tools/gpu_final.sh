#!/bin/bash
# gpurun -- 'bash tools/gpu_final.sh <tag>': everything the round's final single-GPU numbers come from, on one box.
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_round.sh $TAG tests bench ncu
bash tools/bench_configs.sh $TAG all
for wl in C4 C5 C3; do timeout 300 python tools/bench_cached.py --workload $wl >> $OUT/${TAG}_cached.jsonl 2>> $OUT/${TAG}_cached.err; done; cat $OUT/${TAG}_cached.jsonl
bash tools/gpu_exp.sh $TAG "--scale 60" "--spacing 0.3" "--api-mode async --batch 8388608"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
