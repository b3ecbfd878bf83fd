#!/bin/bash
# round 2, GPU call 17 (1 GPU): compositor / epilogue in frame groups on two streams: parity + sweep of the group count
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin.py -m gpu -x -q -k "not full_size_one_frame") > $O/r2c17_tests.log 2>&1; tail -4 $O/r2c17_tests.log | cut -c1-300
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
{ for g in 1 2 4 8 1 4; do RR_EPI_GROUPS=$g $B 2>&1 | show "groups=$g"; done; } > $O/r2c17_ab.txt 2>&1; cat $O/r2c17_ab.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-dropin 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', round(d['value']), round(d['e2e']['value']), d['e2e']['equals_device_arm'])"
