#!/bin/bash
# tuning experiment: stage times of prebuilt library variants (tools/_variant_*.so, not committed)
for so in "" tools/_variant_*.so; do
  RR_LIB_OVERRIDE=$so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$so', round(d['value']), {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
