#!/bin/bash
# round 2, GPU call 21 (8 GPUs): the scaling points the driver measures at round end, with the host-link ceiling of the same box
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $O/r2c21_gpus.txt; nproc >> $O/r2c21_gpus.txt
timeout 300 $TR --nproc-per-node 8 --master-port 29531 tools/pcie_ceiling.py 2>/dev/null | grep -E "^\{" > $O/r2c21_pcie_n8.jsonl; tail -1 $O/r2c21_pcie_n8.jsonl | cut -c1-400
timeout 600 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-dropin 2>$O/r2c21_n8.err | grep -E "^\{" > $O/r2c21_bench_n8.json
timeout 600 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-dropin 2>$O/r2c21_n4.err | grep -E "^\{" > $O/r2c21_bench_n4.json
for n in 8 4; do python - <<PY
import json
d=json.loads(open("$O/r2c21_bench_n$n.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("N=$n device", round(d["value"]), "e2e", round(e["value"]), "link", e.get("host_link_gbs"), "frac", e.get("link_frac"))
PY
done
