#!/bin/bash
# tuning sweep of the two search passes on the default bench workload (results identical; see checksum)
out=gpurun_out/sweep_search2.jsonl; : > $out
run() { echo "# $*" >> $out; env "$1" timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 "${@:2}" >> $out 2>> gpurun_out/sweep_search.err; }
run X=0
run X=0 --opt search_fast_blocks_per_sm=8 --opt search_hard_blocks_per_sm=8
run X=0 --opt search_fast_blocks_per_sm=16
run X=0 --opt search_fast_blocks_per_sm=64
run X=0 --opt search_fast_blocks_per_sm=128
run X=0 --opt search_hard_blocks_per_sm=64
run X=0 --opt search_hard_blocks_per_sm=256
run X=0 --method nearest
run X=0 --method nearest --opt search_fast_blocks_per_sm=8 --opt search_hard_blocks_per_sm=8
run X=0 --mix deep
run X=0 --mix deep --opt search_fast_blocks_per_sm=8 --opt search_hard_blocks_per_sm=8
python - <<'PY'
import json
for l in open('gpurun_out/sweep_search2.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        d = json.loads(l); print('   ', round(d['ms_per_step'], 3), d['roofline']['phase_ms_per_step'], d['checksum'])
    except Exception as e: print('   bad line', e)
PY
# launch list of the bench command on HEAD, then one full-set capture of the four hot kernels
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/r01p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01p_ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_gather_staged|k_search' -s 8 -c 4 -o gpurun_out/r01p_prof python bench.py --steps 1 --warmup 3 --batch 8388608 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01p_ncu_full.log 2>&1
ls -la gpurun_out
