#!/bin/bash
# prints value / unet step / per-kind in-graph ms of one bench run (extra env vars are passed through)
python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', round(d['value'],2), round(d['unet_step_ms'],3), round(d['roofline']['achieved'],1), d['roofline']['other_kinds_ms_in_graph'])"
