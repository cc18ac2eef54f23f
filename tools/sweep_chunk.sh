#!/bin/bash
# end-to-end (host buffers) throughput against the staged chunk size; default bench workload
out=gpurun_out/sweep_chunk.jsonl; : > $out
for c in 4194304 1048576 2097152 8388608; do
  echo "# chunk=$c" >> $out
  timeout 100 python bench.py --steps 1 --warmup 3 --batch 16777216 --no-cpu-baseline --e2e-steps 6 --opt chunk=$c >> $out 2>> gpurun_out/sweep_chunk.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep_chunk.jsonl'):
    if l.startswith('#'): print(l.strip(), end=' -> '); continue
    d = json.loads(l); print(round(d['e2e']['value'] / 1e9, 4), 'e9 pts/s e2e')
PY
