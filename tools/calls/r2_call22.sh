#!/bin/bash
# round 2, GPU call 22 (1 GPU): the round's final evidence: bench lines of both arms, ncu full + launch list, sanitizer pass
O=gpurun_out; mkdir -p $O
python bench.py --steps 20 --warmup 3 2>$O/r02e_bench.err | grep -E "^\{" | tail -1 > $O/r02e_bench.json
python -c "
import json; d=json.load(open('$O/r02e_bench.json')); print('bench', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'dropin', d.get('dropin_png_e2e',{}).get('value'), 'cpu', d['cpu_baseline']['value'], d['roofline'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep -E "^\{" | tail -1 > $O/r02e_bench_reference.json; cut -c1-300 $O/r02e_bench_reference.json
python bench.py --workload C3 --steps 10 --warmup 3 2>/dev/null | grep -E "^\{" | tail -1 > $O/r02e_bench_c3.json; cut -c1-200 $O/r02e_bench_c3.json
bash tools/final_profile.sh r02e
timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "edge_cases or compact_boundary or lanes" > $O/r02e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r02e_memcheck.log | cut -c1-200
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "edge_cases" > $O/r02e_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/r02e_racecheck.log | cut -c1-200
