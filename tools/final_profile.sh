#!/bin/bash
# One GPU-box call that produces the ncu evidence kept under profiles/ for a round:  bash tools/final_profile.sh TAG
#   ncu --set full of one step (-> summary CSV + per-stage DRAM traffic + hot source lines), ncu launch list of two steps.
# Launch bookkeeping (device-resident arm, --skip-e2e): 9 init launches (5 env tables, 3 solid angles, 1 extinction table), 1 texture-border build in the first step,
# 17 per step (stats 2, fog constants + extinction + rolling fog + tile fog 4, env map + prefix + ambient 3, plan + set-up 2,
# scan 1, raster + blur 2, composite + frame mean 2, epilogue 1), 3 warm-up steps.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 61 -c 17 -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --lanes 1 --skip-e2e --no-cpu-baseline --no-dropin > $O/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $O/${TAG}_prof.ncu-rep $O/${TAG}_ncu_full_summary.csv $O/${TAG}_roofline_traffic.json
python tools/ncu_hot_lines.py $O/${TAG}_prof.ncu-rep $O/${TAG}_hot_lines.md k_raster k_fog k_composite k_env_prefix k_setup k_blur k_env_map k_epilogue > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 61 -c 34 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --lanes 1 --skip-e2e --no-cpu-baseline --no-dropin > $O/${TAG}_ncu_launches.log 2>&1
grep -c "k_" $O/${TAG}_launches.csv
