#!/bin/bash
# round 2, GPU call 5 (1 GPU): two-chain pipeline + GPU PNG: parity (fast subset + png + dropin), memcheck of png, bench, dropin e2e
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin.py -m gpu -x -q -k "not full_size_one_frame") > $O/r2c5_tests.log 2>&1; tail -12 $O/r2c5_tests.log | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "png_image_data" > $O/r2c5_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/r2c5_memcheck.log | head -5
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e --no-dropin"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
{ $B 2>&1 | show "two chains"; RR_SERIAL=1 $B 2>&1 | show "serial"; } > $O/r2c5_ab.txt 2>&1; cat $O/r2c5_ab.txt
python tools/dropin_e2e.py 2048 64 0,8 2>&1 | grep -E "^\{|Error|error" > $O/r2c5_dropin.jsonl
RAIN_B200_GPU_PNG=0 python tools/dropin_e2e.py 2048 64 0 2>&1 | grep -E "^\{|Error|error" >> $O/r2c5_dropin.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2c5_dropin.jsonl"):
    try:
        d=json.loads(l); print(round(d['value']), d['steady_frames_per_s'] and round(d['steady_frames_per_s']), d['batch'], d['io_threads'], d['host_cores'], d['waits_s'])
    except Exception as e: print(l[:300])
PY
