#!/bin/bash
# round 2, GPU call 6 (1 GPU): simulator tests (device-resident path, DSD chi-square), C3 bench, drop-in e2e with decode-ahead, full bench line
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_simulator.py tests/test_dropin.py -m gpu -x -q) > $O/r2c6_tests.log 2>&1; tail -25 $O/r2c6_tests.log | cut -c1-400
python bench.py --workload C3 --steps 10 --warmup 3 > $O/r2c6_bench_c3.json 2> $O/r2c6_bench_c3.err; tail -c 1500 $O/r2c6_bench_c3.json; tail -3 $O/r2c6_bench_c3.err
python tools/dropin_e2e.py 2048 64 0 2>&1 | grep -E "^\{|Error|error" > $O/r2c6_dropin.jsonl
python tools/dropin_e2e.py 2048 32 0 2>&1 | grep -E "^\{|Error|error" >> $O/r2c6_dropin.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2c6_dropin.jsonl"):
    try:
        d=json.loads(l); print(round(d['value']), d['steady_frames_per_s'] and round(d['steady_frames_per_s']), d['batch'], d['io_threads'], d['host_cores'], d['waits_s'])
    except Exception as e: print(l[:300])
PY
python bench.py --steps 20 --warmup 3 > $O/r2c6_bench.json 2> $O/r2c6_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2c6_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['e2e'].get('link_frac'), d.get('dropin_png_e2e'), d.get('cpu_baseline'), d['stage_ms'])"; tail -3 $O/r2c6_bench.err
