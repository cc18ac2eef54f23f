#!/bin/bash
# bench line per tuning variant (tools/bin/libndpb200_<tag>.so): gpurun -- 'bash tools/gpu_variants.sh <outtag> tag1 tag2 ... [-- extra bench args]'
OUTTAG=$1; shift
TAGS=(); EXTRA=()
while [ $# -gt 0 ]; do if [ "$1" == "--" ]; then shift; EXTRA=("$@"); break; fi; TAGS+=("$1"); shift; done
mkdir -p gpurun_out
: > gpurun_out/${OUTTAG}_variants.jsonl
for t in "${TAGS[@]}"; do
  lib=tools/bin/libndpb200_$t.so
  [ "$t" == "default" ] && lib=ndpolator_b200/libndpb200.so
  echo "== $t" >&2
  NDPB_LIB=$PWD/$lib timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-batch 1048576 --e2e-steps 1 "${EXTRA[@]}" 2> gpurun_out/${OUTTAG}_$t.err | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print(json.dumps({'variant': '$t', 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'whole_step_frac': d['roofline']['whole_step']['frac'], 'phase': d['roofline']['phase_ms_per_step'], 'checksum': d['checksum'], 'clocks': d['clocks']}))
" | tee -a gpurun_out/${OUTTAG}_variants.jsonl
done
