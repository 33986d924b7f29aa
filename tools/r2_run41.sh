#!/bin/bash
cd /root/repo
python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "groupnorm" 2>&1 | tail -3
python tools/gn_parts_ablate.py 2>&1 | grep norm
python -m pytest tests/test_boundary_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -5
