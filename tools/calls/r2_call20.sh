#!/bin/bash
# round 2, GPU call 20 (1 GPU): zero-bordered texture copies in the rasteriser: parity, A/B against the plain sampler
O=gpurun_out; mkdir -p $O
(time timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q) > $O/r2c20_tests.log 2>&1; tail -4 $O/r2c20_tests.log | cut -c1-300
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d.get('stage_ms_solo') or {}; print('$1', round(d['value']), round(d['ms_per_step'],3), 'raster', s.get('raster'), 'solo', {k: round(v,3) for k,v in s.items()})"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-dropin"
{ $B 2>/dev/null | show "padded"; RR_LIB_OVERRIDE=tools/_variant_nopad.so $B 2>/dev/null | show "plain"; $B 2>/dev/null | show "padded"; RR_LIB_OVERRIDE=tools/_variant_nopad.so $B 2>/dev/null | show "plain"; } > $O/r2c20_ab.txt 2>&1; cat $O/r2c20_ab.txt
