#!/bin/bash
# round 2, GPU call 33 (1 GPU): drop-in PNG pass with the row-sink decoder (the round's last GPU seconds)
O=gpurun_out; mkdir -p $O
python tools/dropin_e2e.py 2048 64 0 2>/dev/null | grep -E "^\{" > $O/r02h_dropin_e2e.jsonl
python -c "
import json
for l in open('$O/r02h_dropin_e2e.jsonl'):
    d=json.loads(l); print('dropin', round(d['value']), d['steady_frames_per_s'] and round(d['steady_frames_per_s']), d['waits_s'])"
