#!/bin/bash
# round 2, GPU call 27 (1 GPU): compute-sanitizer over more of the suite (racecheck, initcheck, synccheck)
O=gpurun_out; mkdir -p $O
run() { tool=$1; tag=$2; shift 2; timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" > $O/r02e_${tool}_${tag}.log 2>&1; echo "$tool $tag rc=$?"; grep -E "passed|failed|SUMMARY" $O/r02e_${tool}_${tag}.log | tail -2 | cut -c1-160; }
run racecheck png tests/test_parity_gpu.py -m gpu -x -q -k "png_image"
run racecheck parity2 tests/test_parity_gpu.py -m gpu -x -q -k "many_streaks or render_scale_2 or stage_parity or degenerate or compact_boundary"
run racecheck sim tests/test_simulator.py -m gpu -x -q
run synccheck edge tests/test_parity_gpu.py -m gpu -x -q -k "edge_cases or bright_frames or png_image"
run initcheck edge tests/test_parity_gpu.py -m gpu -x -q -k "edge_cases"
