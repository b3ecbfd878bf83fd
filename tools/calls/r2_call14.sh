#!/bin/bash
# round 2, GPU call 14 (1 GPU): stream priorities A/B
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
{
$B 2>&1 | show "prio=1"
RR_STREAM_PRIO=0 $B 2>&1 | show "prio=0"
$B 2>&1 | show "prio=1"
RR_STREAM_PRIO=0 $B 2>&1 | show "prio=0"
} > $O/r2c14_prio.txt 2>&1; cat $O/r2c14_prio.txt
