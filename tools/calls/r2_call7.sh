#!/bin/bash
# round 2, GPU call 7 (1 GPU): drop-in / launcher / simulator tests, C3 bench, drop-in e2e, full bench line
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_simulator.py tests/test_dropin.py tests/test_launcher.py -m gpu -x -q) > $O/r2c7_tests.log 2>&1; tail -8 $O/r2c7_tests.log | cut -c1-400
python bench.py --workload C3 --steps 10 --warmup 3 2> $O/r2c7_bench_c3.err | tail -1 > $O/r2c7_bench_c3.json; python -c "
import json; d=json.load(open('gpurun_out/r2c7_bench_c3.json')); print({k:d[k] for k in ('value','ms_per_step','sim_ms_per_step','render_ms_per_step','wall_ms_per_step','gpu_launches')})"; tail -2 $O/r2c7_bench_c3.err
python tools/dropin_e2e.py 2048 64 0 2>&1 | grep -E "^\{|Error|error" > $O/r2c7_dropin.jsonl
python tools/dropin_e2e.py 2048 32 0 2>&1 | grep -E "^\{|Error|error" >> $O/r2c7_dropin.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2c7_dropin.jsonl"):
    try:
        d=json.loads(l); print(round(d['value']), d['steady_frames_per_s'] and round(d['steady_frames_per_s']), d['batch'], d['io_threads'], d['host_cores'], d['setup_s'], d['waits_s'])
    except Exception as e: print(l[:300])
PY
python bench.py --steps 20 --warmup 3 2> $O/r2c7_bench.err | tail -1 > $O/r2c7_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2c7_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['e2e'].get('link_frac'), d.get('dropin_png_e2e'), d.get('cpu_baseline'), d['stage_ms'])"; tail -3 $O/r2c7_bench.err
