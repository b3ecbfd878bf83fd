#!/bin/bash
# round 2, GPU call 28 (2 GPUs): the lanes bench under torchrun (what the driver's scaling run launches), both arms
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 40 --warmup 3 2>$O/r2c28_n2.err | grep -E "^\{" > $O/r02e_bench_n2.json; tail -3 $O/r2c28_n2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("$O/r02e_bench_n2.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("N=2 device", round(d["value"]), d["ms_per_step"], "e2e", round(e["value"]), "link", e.get("host_link_gbs"), "frac", e.get("link_frac"), "launches", d["gpu_launches"], d["config"]["lanes_per_gpu"])
PY
timeout 300 $TR --nproc-per-node 2 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | grep -E "^\{" | cut -c1-200
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_launcher.py -m gpu -x -q -k "two_contexts_on_two_devices or launcher" 2>&1 | tail -2
