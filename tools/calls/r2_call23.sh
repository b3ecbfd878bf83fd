#!/bin/bash
# round 2, GPU call 23 (1 GPU): e2e loop with the next batch assembled before the wait; sub-batch count
O=gpurun_out; mkdir -p $O
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$1', round(d['value']), 'e2e', round(e['value']), round(e['ms_per_step'],3), e['equals_device_arm'], e['link_frac'])"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-dropin"
{ $B 2>/dev/null | show "S=2"; RR_SUB_BATCHES_ASYNC=1 $B 2>/dev/null | show "S=1"; RR_SUB_BATCHES_ASYNC=3 $B 2>/dev/null | show "S=3"; RR_SUB_BATCHES_ASYNC=4 $B 2>/dev/null | show "S=4"; $B 2>/dev/null | show "S=2"; } > $O/r2c23_ab.txt 2>&1; cat $O/r2c23_ab.txt
