#!/bin/bash
O=gpurun_out; mkdir -p $O
{ python tools/exp_two_contexts.py 1 64; python tools/exp_two_contexts.py 2 32; python tools/exp_two_contexts.py 4 16; python tools/exp_two_contexts.py 2 64; python tools/exp_two_contexts.py 3 64; } 2>&1 | grep -E "lanes|Error|error" > $O/r2c12_lanes.txt; cat $O/r2c12_lanes.txt
