#!/bin/bash
# last GPU call of a session: full parity suite, memcheck of the smoke render, default vs variant timing, variant parity
O=gpurun_out; mkdir -p $O
(time timeout 400 python -m pytest tests -m gpu -x -q) > $O/fc_tests.log 2>&1; grep -E "passed|failed" $O/fc_tests.log
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/fc_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|Error" $O/fc_memcheck.log | head -8
bash tools/variant_sweep.sh > $O/fc_sweep.log 2>&1; cat $O/fc_sweep.log
for so in tools/_variant_*.so; do
  RR_LIB_OVERRIDE=$so timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_frames or many_streaks or determinism or pipelined" > $O/fc_variant_tests.log 2>&1; echo "$so: $(grep -E 'passed|failed' $O/fc_variant_tests.log)"
done
