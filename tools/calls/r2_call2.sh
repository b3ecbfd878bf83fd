#!/bin/bash
# round 2, GPU call 2 (2 GPUs): parity suite with the compact boundary formats, bench at N=1 and N=2
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/r2c2_tests.log 2>&1; tail -5 $O/r2c2_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2c2_bench1.json 2> $O/r2c2_bench1.err; tail -c 3000 $O/r2c2_bench1.json; tail -3 $O/r2c2_bench1.err
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 3 2>$O/r2c2_bench2.err | tail -1 > $O/r2c2_bench2.json
python - <<'PY'
import json
for f in ("1","2"):
    try:
        d=json.load(open("gpurun_out/r2c2_bench%s.json"%f)); e=d["e2e"]; print(f, round(d["value"]), round(e["value"]), e.get("host_link_gbs"), e.get("link_frac"), e.get("equals_device_arm"), d["stage_ms"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 $O/r2c2_bench2.err
