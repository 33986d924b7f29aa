#!/bin/bash
cd /root/repo
python tools/gn_parts_dbg.py enc 2>&1 | tail -40
python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "groupnorm" 2>&1 | tail -5
python tools/gn_parts_ablate.py 2>&1 | grep norm
