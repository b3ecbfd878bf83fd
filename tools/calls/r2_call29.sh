#!/bin/bash
# round 2, GPU call 29 (1 GPU): the final tree once more: smoke(), the default bench line, the quick GPU tests
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py 2>$O/r02f_bench.err | grep -E "^\{" | tail -1 > $O/r02f_bench.json
python -c "
import json; d=json.load(open('$O/r02f_bench.json')); p=d.get('dropin_png_e2e',{}); print('bench', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'dropin', round(p.get('value',0)), 'enqueue ms/step', round(d['host_enqueue_ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['roofline']['traffic'], d['clocks'])"
(time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_baseline_configs_at_full_size_one_frame_each) > $O/r02f_gpu_tests.log 2>&1; tail -3 $O/r02f_gpu_tests.log | cut -c1-200
