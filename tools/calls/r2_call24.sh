#!/bin/bash
# round 2, GPU call 24 (1 GPU): end to end over several contexts on one GPU
O=gpurun_out; mkdir -p $O
{
python tools/exp_lanes_e2e.py 1 2 60
python tools/exp_lanes_e2e.py 2 1 60
RR_SUB_BATCHES_ASYNC=1 python tools/exp_lanes_e2e.py 2 1 60
RR_SUB_BATCHES_ASYNC=1 python tools/exp_lanes_e2e.py 3 1 60
python tools/exp_lanes_e2e.py 2 2 60
RR_SUB_BATCHES_ASYNC=1 python tools/exp_lanes_e2e.py 2 2 60
python tools/exp_lanes_e2e.py 3 1 60
} 2>&1 | grep -E "^lanes|Error|error" > $O/r2c24_lanes.txt; cat $O/r2c24_lanes.txt
