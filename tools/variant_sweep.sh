#!/bin/bash
# tuning experiment: stage times of prebuilt library variants (tools/_variant_*.so, not committed); "" = the in-tree library.
# Every library twice: the two-chain pipeline (what ships) and RR_SERIAL=1 (one kernel at a time: clean per-stage times).
for so in "" tools/_variant_*.so; do
  for serial in 0 1; do
    RR_SERIAL=$serial RR_LIB_OVERRIDE=$so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$so serial=$serial', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
  done
done
