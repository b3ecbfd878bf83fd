#!/bin/bash
# round 2, GPU call 11 (1 GPU): rolling-window fog kernel: parity, then A/B against the tile kernel
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not full_size_one_frame") > $O/r2c11_tests.log 2>&1; tail -15 $O/r2c11_tests.log | cut -c1-300
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
{
$B 2>&1 | show "roll=1"
RR_FOG_ROLL=0 $B 2>&1 | show "roll=0"
RR_SERIAL=1 $B 2>&1 | show "roll=1 serial"
RR_SERIAL=1 RR_FOG_ROLL=0 $B 2>&1 | show "roll=0 serial"
} > $O/r2c11_ab.txt 2>&1; cat $O/r2c11_ab.txt
