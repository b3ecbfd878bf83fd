#!/bin/bash
O=gpurun_out; mkdir -p $O
bash tools/variant_sweep.sh > $O/r2c15_sweep.txt 2>&1; cat $O/r2c15_sweep.txt
