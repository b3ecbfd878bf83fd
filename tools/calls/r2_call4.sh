#!/bin/bash
# round 2, GPU call 4 (1 GPU): parity with the TMA tile loads + integer rasteriser; A/B stage times of the TMA variants
O=gpurun_out; mkdir -p $O
(time timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q) > $O/r2c4_tests.log 2>&1; tail -6 $O/r2c4_tests.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
{
$B 2>&1 | show "tma=1 bulk=1"
RR_FOG_TMA=0 $B 2>&1 | show "tma=0 bulk=1"
RR_ENV_BULK=0 $B 2>&1 | show "tma=1 bulk=0"
RR_FOG_TMA=0 RR_ENV_BULK=0 $B 2>&1 | show "tma=0 bulk=0"
$B 2>&1 | show "tma=1 bulk=1 again"
} > $O/r2c4_ab.txt 2>&1
cat $O/r2c4_ab.txt
RR_FOG_TMA=0 RR_ENV_BULK=0 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_frames or stage_parity or compact" > $O/r2c4_tests_notma.log 2>&1; tail -3 $O/r2c4_tests_notma.log
