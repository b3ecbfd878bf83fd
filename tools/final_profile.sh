#!/bin/bash
# One GPU-box call that produces everything kept under profiles/ for a round:  bash tools/final_profile.sh TAG
#   tests log, ncu --set full of one step (-> summary CSV + per-stage DRAM traffic), ncu launch list, the bench line,
#   the reference arm, the C3-C5 stress run.
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
(time timeout 600 python -m pytest tests -m gpu -x -q) > $O/${TAG}_tests.log 2>&1; grep -E "passed|failed" $O/${TAG}_tests.log
# init launches: 4 (env tables) + 3 (solid angles); a step = 15 launches; 3 warm-up steps
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ -s 52 -c 15 -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $O/${TAG}_prof.ncu-rep $O/${TAG}_ncu_full_summary.csv profiles/roofline_traffic.json; cp profiles/roofline_traffic.json $O/roofline_traffic.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 52 -c 30 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 2500 $O/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; cat $O/${TAG}_bench_reference.json | cut -c1-300
timeout 300 python tools/stress_configs.py > $O/${TAG}_stress.log 2>&1; tail -4 $O/${TAG}_stress.log
