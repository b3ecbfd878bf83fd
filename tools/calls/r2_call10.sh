#!/bin/bash
# round 2, GPU call 10 (1 GPU): compositor software pipeline variants
O=gpurun_out; mkdir -p $O
bash tools/variant_sweep.sh > $O/r2c10_sweep.txt 2>&1; cat $O/r2c10_sweep.txt
RR_LIB_OVERRIDE=tools/_variant_pipe4.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_frames or many_streaks or determinism or edge" > $O/r2c10_tests_pipe.log 2>&1; tail -2 $O/r2c10_tests_pipe.log
