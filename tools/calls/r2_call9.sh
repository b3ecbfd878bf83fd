#!/bin/bash
# round 2, GPU call 9 (1 GPU): quad-packed raster textures, 3-double prefix entries, streaming stores: parity + A/B sweep
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not full_size_one_frame") > $O/r2c9_tests.log 2>&1; tail -4 $O/r2c9_tests.log | cut -c1-300
bash tools/variant_sweep.sh > $O/r2c9_sweep.txt 2>&1; cat $O/r2c9_sweep.txt
