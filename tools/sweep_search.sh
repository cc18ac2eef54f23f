#!/bin/bash
# Tuning sweep of the search passes on the default bench workload (profiles/r01p_sweep_search.txt).
# Results must stay identical: compare the checksum.  Variant libraries are built beforehand with
#   python -c "from ndpolator_b200 import build; build.build(force=True, extra_flags=['-DNDP_FAST_MARGIN'], lib_out='tools/bin/libndpb200_margin.so')"
# (other switches tried in round 1: -DNDP_FAST_OCC=3|4, -DNDP_HARD_OCC=3).
out=gpurun_out/sweep_search.jsonl; : > $out
run() { echo "# $*" >> $out; env "$1" timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 "${@:2}" >> $out 2>> gpurun_out/sweep_search.err; }
run X=0
run X=0 --opt search_fast_blocks_per_sm=8 --opt search_hard_blocks_per_sm=8
[ -f tools/bin/libndpb200_margin.so ] && run NDPB_LIB=tools/bin/libndpb200_margin.so
python - <<'PY'
import json
for l in open('gpurun_out/sweep_search.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        d = json.loads(l); print('   ', round(d['ms_per_step'], 3), d['roofline']['phase_ms_per_step'], d['checksum'])
    except Exception as e: print('   bad line', e)
PY
