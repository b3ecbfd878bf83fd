#!/bin/bash
# round 2, GPU call 3 (1 GPU): drop-in tests and the PNG end-to-end pipeline at several thread counts
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_dropin.py tests/test_pngio.py -m gpu -x -q) > $O/r2c3_tests.log 2>&1; tail -15 $O/r2c3_tests.log
nproc
python tools/dropin_e2e.py 1024 64 0,16,8 2>&1 | grep -E "^\{|Error|error" > $O/r2c3_dropin.jsonl; cat $O/r2c3_dropin.jsonl
python tools/dropin_e2e.py 1024 32 0 2>&1 | grep -E "^\{|Error|error" >> $O/r2c3_dropin.jsonl; tail -1 $O/r2c3_dropin.jsonl
