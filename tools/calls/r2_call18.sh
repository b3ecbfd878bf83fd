#!/bin/bash
# round 2, GPU call 18 (1 GPU): extinction plane on the side stream A/B; the complete GPU suite + smoke
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
{ for v in 1 0 1 0; do RR_FEXT_SIDE=$v $B 2>&1 | show "fext_side=$v"; done; } > $O/r2c18_ab.txt 2>&1; cat $O/r2c18_ab.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r2c18_tests.log 2>&1; tail -5 $O/r2c18_tests.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-400
