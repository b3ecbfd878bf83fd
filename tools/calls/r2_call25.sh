#!/bin/bash
# round 2, GPU call 25 (1 GPU): api.RainLanes: parity test, bench with 1 / 2 / 3 lanes (device-resident and end to end)
O=gpurun_out; mkdir -p $O
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "lanes or pipelined") > $O/r2c25_tests.log 2>&1; tail -3 $O/r2c25_tests.log | cut -c1-300
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$1', 'device', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), e['equals_device_arm'], 'link_frac', round(e['link_frac'],3), 'launches', d['gpu_launches'], 'dom', d['roofline']['kernel'], round(d['roofline']['kernel_ms_in_step'],3))"; }
B="python bench.py --steps 36 --warmup 3 --no-cpu-baseline --no-dropin"
{ for L in 1 2 3 4 2; do $B --lanes $L 2>$O/r2c25_err_$L.log | show "lanes=$L"; done; } > $O/r2c25_lanes.txt 2>&1; cat $O/r2c25_lanes.txt; tail -3 $O/r2c25_err_2.log
