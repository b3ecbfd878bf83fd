#!/bin/bash
# round 2, GPU call 30 (1 GPU): rasteriser variants on top of the bordered textures (samples in flight, resident CTAs)
O=gpurun_out; mkdir -p $O
for so in "" tools/_variant_unroll3.so tools/_variant_unroll4.so tools/_variant_minb4.so tools/_variant_minb6.so ""; do
  RR_LIB_OVERRIDE=$so python bench.py --steps 40 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${so:-default}', round(d['value']), round(d['ms_per_step'],3), 'raster solo', round(d['stage_ms_solo']['raster'],3), 'sum solo', round(sum(d['stage_ms_solo'].values()),3))"
done > $O/r2c30_raster_variants.txt 2>&1; cat $O/r2c30_raster_variants.txt
