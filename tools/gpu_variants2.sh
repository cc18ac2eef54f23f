#!/bin/bash
# bench lines of several workloads per tuning variant: gpurun -- 'bash tools/gpu_variants2.sh <outtag> "tag1 tag2" "<bench args A>" "<bench args B>" ...'
OUTTAG=$1; TAGS=$2; shift 2
mkdir -p gpurun_out; L=gpurun_out/${OUTTAG}_variants2.jsonl; : > $L
for args in "$@"; do
  for t in $TAGS; do
    lib=tools/bin/libndpb200_$t.so; [ "$t" == "default" ] && lib=ndpolator_b200/libndpb200.so
    echo "# $t | $args" >> $L
    NDPB_LIB=$PWD/$lib timeout 300 python bench.py --warmup 3 --no-cpu-baseline --e2e-batch 1048576 --e2e-steps 1 $args >> $L 2>> gpurun_out/${OUTTAG}_variants2.err
  done
done
python tools/show_lines.py $L
