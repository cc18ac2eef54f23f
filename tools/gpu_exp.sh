#!/bin/bash
# One gpurun call of tuning experiments on one GPU: each line of the table is a bench.py run with extra flags.
#   gpurun --timeout 1700 -- 'bash tools/gpu_exp.sh <tag> "<flags A>" "<flags B>" ...'
TAG=$1; shift
OUT=gpurun_out/${TAG}_exp.jsonl
mkdir -p gpurun_out; : > $OUT
for flags in "$@"; do
  echo "# $flags" >> $OUT
  timeout 900 python bench.py --no-cpu-baseline --steps 5 --warmup 3 --e2e-batch 4194304 --e2e-steps 3 $flags >> $OUT 2>> gpurun_out/${TAG}_exp.err
  echo "rc=$? $flags"
done
python - <<PY
import json
for l in open('$OUT'):
    if l.startswith('#'):
        print(l.strip()); continue
    try:
        d = json.loads(l)
        r = d['roofline']
        print('   value %.4e  ms %.3f  frac %.3f  whole %.3f  slow %.2e  build %.2fs cold %.3fs  clk %s %s' % (
            d['value'], d['ms_per_step'], r['frac'] or 0, r['whole_step']['frac'], d['mix_measured']['left_to_exact_generic_kernel'],
            d['table']['build_s'], d['table']['cold_first_call_s'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print('   unreadable', e)
PY
