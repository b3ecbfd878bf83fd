#!/bin/bash
# round 2, GPU call 16 (1 GPU): register-resident interval merge in k_setup: parity + timing
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not full_size_one_frame") > $O/r2c16_tests.log 2>&1; tail -4 $O/r2c16_tests.log | cut -c1-300
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()}, {k: round(v,3) for k,v in d['stage_ms_solo'].items()})"; }
{ $B 2>&1 | show "setup-regs"; $B 2>&1 | show "setup-regs"; } > $O/r2c16_ab.txt 2>&1; cat $O/r2c16_ab.txt
