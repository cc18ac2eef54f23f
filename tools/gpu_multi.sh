#!/bin/bash
# One N-GPU gpurun call: group tests, [PCIe probe], torchrun bench + C5 sweep at the given rank counts, in-process group bench.
#   gpurun --gpus 8 -- 'bash tools/gpu_multi.sh r02s 8 "8" "2 4 8" [probe]'
#   $3 = rank counts of the default bench line, $4 = rank counts of the C5 batch sweep
TAG=$1; NMAX=$2; RANKS=${3:-$NMAX}; SWEEP_RANKS=${4:-$RANKS}; PROBE=${5:-}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "group or device_list" > $OUT/${TAG}_tests_group.log 2>&1; echo "group tests rc=$?"; tail -2 $OUT/${TAG}_tests_group.log
if [ -n "$PROBE" ]; then
  : > $OUT/${TAG}_probe_pcie.jsonl
  for n in $RANKS; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tools/probe_pcie.py 2>/dev/null >> $OUT/${TAG}_probe_pcie.jsonl
  done
  grep '"both"' $OUT/${TAG}_probe_pcie.jsonl | cut -c1-260
fi
for n in $RANKS; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n --steps 8 --warmup 3 > $OUT/${TAG}_bench_${n}gpu.json 2> $OUT/${TAG}_bench_${n}gpu.err; echo "bench $n rc=$?"
done
# C5 batch sweep (BASELINE.json configs[4]): one launch per rank count, one JSON line per total
for n in $SWEEP_RANKS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $n --workload C5 --sweep-totals 1e3,1e4,1e5,1e6,1e7,1e8,1e9 --steps 200 --warmup 5 --e2e-batch 4194304 --e2e-steps 3 > $OUT/${TAG}_bench_c5_sweep_${n}gpu.jsonl 2> $OUT/${TAG}_bench_c5_sweep_${n}gpu.err; echo "sweep $n rc=$?"
done
python tools/show_lines.py $OUT/${TAG}_bench_c5_sweep_*gpu.jsonl
timeout 600 python bench.py --gpus $NMAX --inprocess --steps 10 --warmup 2 --e2e-batch 16777216 > $OUT/${TAG}_bench_inprocess_${NMAX}gpu.json 2> $OUT/${TAG}_bench_inprocess_${NMAX}gpu.err; echo "inprocess rc=$?"
python tools/show_lines.py $OUT/${TAG}_bench_*gpu.json
