#!/bin/bash
# tuning sweep of the search passes on the default bench workload (results identical; see checksum)
out=gpurun_out/sweep_search3.jsonl; : > $out
run() { echo "# $*" >> $out; env "$1" timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 "${@:2}" >> $out 2>> gpurun_out/sweep_search.err; }
run X=0
for b in 2 4 8 32; do run NDPB_LIB=tools/bin/libndpb200_steal.so --opt search_hard_blocks_per_sm=$b; done
run NDPB_LIB=tools/bin/libndpb200_steal.so --opt search_hard_blocks_per_sm=2 --mix deep
python - <<'PY'
import json
for l in open('gpurun_out/sweep_search3.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        d = json.loads(l); print('   ', round(d['ms_per_step'], 3), d['roofline']['phase_ms_per_step'], d['checksum'])
    except Exception as e: print('   bad line', e)
PY
tail -5 gpurun_out/sweep_search.err
