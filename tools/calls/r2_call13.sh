#!/bin/bash
# round 2, GPU call 13 (2 GPUs): FOGR_SEG sweep (GPU 0), two-device test, N=2 bench + host-link ceiling
O=gpurun_out; mkdir -p $O
bash tools/variant_sweep.sh > $O/r2c13_sweep.txt 2>&1; cat $O/r2c13_sweep.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_contexts or compact or png_image" > $O/r2c13_tests.log 2>&1; tail -3 $O/r2c13_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29521 tools/pcie_ceiling.py 2>&1 | grep pcie_ceiling > $O/r2c13_ceiling2.json; cut -c1-400 $O/r2c13_ceiling2.json
$TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 2>$O/r2c13_bench2.err | tail -1 > $O/r2c13_bench2.json
python -c "
import json; d=json.load(open('gpurun_out/r2c13_bench2.json')); e=d['e2e']; print('N=2', round(d['value']), round(e['value']), e.get('host_link_gbs'), e.get('link_frac'))"; tail -2 $O/r2c13_bench2.err
