#!/usr/bin/env python
"""One line per bench.py JSON line of a .jsonl file: workload, queries/step, pts/s, ms/step, roofline fractions, e2e."""
import json
import sys
for f in sys.argv[1:]:
    print(f'== {f}')
    for l in open(f):
        l = l.strip()
        if not l.startswith('{'):
            if l:
                print(l)
            continue
        d = json.loads(l)
        r = d.get('roofline', {})
        c = d.get('clocks') or {}
        print('%-46s n/step %.0e  %.3e pts/s  %9.4f ms  kern %.3f  step %.3f  e2e %.3e  launches %d  clk %s/%s %s' % (
            d['config']['workload'].replace('extrapolation_method=', '').replace('search_algorithm=', '')[:46],
            d['config']['queries_per_step_all_gpus'], d['value'], d['ms_per_step'], r.get('frac') or 0,
            (r.get('whole_step') or {}).get('frac') or 0, d['e2e']['value'], d.get('gpu_launches', 0),
            c.get('sm_mhz'), c.get('samples'), ','.join(c.get('reasons') or [])))
