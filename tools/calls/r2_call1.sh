#!/bin/bash
# round 2, GPU call 1 (2 GPUs): box topology, host-link ceiling with / without NUMA binding, r2-variants sweep, e2e at N=2
O=gpurun_out; mkdir -p $O
{ nvidia-smi topo -m; lscpu | head -30; echo "nproc $(nproc)"; cat /sys/devices/system/node/online; for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done; free -g | head -2; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; } > $O/r2c1_sysinfo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
python tools/pcie_ceiling.py --no-bind
python tools/pcie_ceiling.py
python tools/pcie_ceiling.py --wc
python tools/pcie_ceiling.py --direction 1
python tools/pcie_ceiling.py --direction 2
$TR --nproc-per-node 2 --master-port 29511 tools/pcie_ceiling.py --no-bind
$TR --nproc-per-node 2 --master-port 29512 tools/pcie_ceiling.py
$TR --nproc-per-node 2 --master-port 29513 tools/pcie_ceiling.py --wc
} 2>&1 | grep -E "pcie_ceiling|Error|error" > $O/r2c1_ceiling.txt
cat $O/r2c1_ceiling.txt | cut -c1-600
bash tools/variant_sweep.sh > $O/r2c1_sweep.log 2>&1; cat $O/r2c1_sweep.log
RAIN_B200_NUMA_BIND=0 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 > $O/r2c1_bench2_nobind.json
$TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 > $O/r2c1_bench2_bind.json
python - <<'PY'
import json
for f in ("nobind","bind"):
    try:
        d=json.load(open("gpurun_out/r2c1_bench2_%s.json"%f)); print(f, round(d["value"]), round(d["e2e"]["value"]), d["e2e"].get("numa_bind"))
    except Exception as e: print(f, "failed", e)
PY
