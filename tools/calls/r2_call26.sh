#!/bin/bash
# round 2, GPU call 26 (1 GPU): after the k_fog_roll barrier fix and the idempotent set-up calls: racecheck, then the round's
# final evidence again (bench lines of both arms, ncu full + launch list, memcheck)
O=gpurun_out; mkdir -p $O
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "edge_cases or bright_frames" > $O/r02e_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 $O/r02e_racecheck.log | cut -c1-200
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "png_image" > $O/r02e_racecheck_png.log 2>&1; echo "racecheck png rc=$?"; tail -2 $O/r02e_racecheck_png.log | cut -c1-200
python bench.py --steps 40 --warmup 3 2>$O/r02e_bench.err | grep -E "^\{" | tail -1 > $O/r02e_bench.json
python -c "
import json; d=json.load(open('$O/r02e_bench.json')); p=d.get('dropin_png_e2e',{}); print('bench', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'dropin', p.get('value'), p.get('setup_s'), p.get('steady_frames_per_s'), 'cpu', d['cpu_baseline']['value'], d['roofline']['frac'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep -E "^\{" | tail -1 > $O/r02e_bench_reference.json; cut -c1-200 $O/r02e_bench_reference.json
python bench.py --workload C3 --steps 10 --warmup 3 2>/dev/null | grep -E "^\{" | tail -1 > $O/r02e_bench_c3.json; cut -c1-200 $O/r02e_bench_c3.json
bash tools/final_profile.sh r02e
timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "edge_cases or compact_boundary or lanes" > $O/r02e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r02e_memcheck.log | cut -c1-200
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r02e_gpu_tests.log 2>&1; tail -4 $O/r02e_gpu_tests.log | cut -c1-200
