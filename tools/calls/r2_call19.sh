#!/bin/bash
# round 2, GPU call 19 (1 GPU): whole-batch double-buffered submissions: async parity tests, e2e A/B, drop-in
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin.py tests/test_launcher.py -m gpu -x -q -k "pipelined or png_image or generator_run or launcher or compact") > $O/r2c19_tests.log 2>&1; tail -4 $O/r2c19_tests.log | cut -c1-300
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$1', round(d['value']), 'e2e', round(e['value']), round(e['ms_per_step'],3), e['equals_device_arm'], e['link_frac'])"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-dropin"
{ $B 2>/dev/null | show "wide=1"; RR_WIDE=0 $B 2>/dev/null | show "wide=0"; $B 2>/dev/null | show "wide=1"; RR_WIDE=0 $B 2>/dev/null | show "wide=0"; } > $O/r2c19_ab.txt 2>&1; cat $O/r2c19_ab.txt
python tools/dropin_e2e.py 2048 64 0 2>&1 | grep -E "^\{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('dropin', round(d['value']), d['steady_frames_per_s'] and round(d['steady_frames_per_s']), d['waits_s'])"
