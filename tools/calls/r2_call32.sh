#!/bin/bash
# round 2, GPU call 32 (1 GPU): drop-in PNG pass after the faster Huffman table build; PNG / drop-in GPU tests
O=gpurun_out; mkdir -p $O
python tools/dropin_e2e.py 2048 64 0 2>/dev/null | grep -E "^\{" > $O/r02g_dropin_e2e.jsonl
python -c "
import json
for l in open('$O/r02g_dropin_e2e.jsonl'):
    d=json.loads(l); print('dropin', round(d['value']), d['steady_frames_per_s'] and round(d['steady_frames_per_s']), d['waits_s'])"
timeout 100 python -m pytest tests/test_dropin.py tests/test_pngio.py -m gpu -x -q 2>&1 | tail -2
