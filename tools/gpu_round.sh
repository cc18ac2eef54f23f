#!/bin/bash
# One gpurun call: GPU parity suite, bench lines, ncu launch list and one full capture of the top kernel.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh <tag> [steps...]'
# steps: tests bench old ncu   (default: all)
set -u
TAG=${1:-r02a}; shift || true
STEPS=${*:-tests bench old ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
for step in $STEPS; do
case $step in
tests)
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/${TAG}_status.txt ;;
bench)
  timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" | tee -a $OUT/${TAG}_status.txt
  tail -c 1500 $OUT/${TAG}_bench.json ;;
old)
  timeout 600 python bench.py --steps 8 --warmup 3 --opt fused=1 --no-cpu-baseline > $OUT/${TAG}_bench_threepass.json 2> $OUT/${TAG}_bench_threepass.err; echo "old rc=$?" | tee -a $OUT/${TAG}_status.txt ;;
deep)
  timeout 600 python bench.py --steps 8 --warmup 3 --mix deep --no-cpu-baseline > $OUT/${TAG}_bench_deep.json 2> $OUT/${TAG}_bench_deep.err; echo "deep rc=$?" | tee -a $OUT/${TAG}_status.txt ;;
ncu)
  CMD="python bench.py --steps 2 --warmup 1 --batch 8388608 --e2e-batch 1048576 --e2e-steps 1 --no-cpu-baseline"
  timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "ncu launches rc=$?" | tee -a $OUT/${TAG}_status.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 2 -f -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?" | tee -a $OUT/${TAG}_status.txt ;;
esac
done
cat $OUT/${TAG}_status.txt
