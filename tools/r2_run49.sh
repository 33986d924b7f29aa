#!/bin/bash
# the A/B switches still give green parity: self-contained GroupNorm kernels, explicit LayerNorm kernels
cd /root/repo
LS_GN_PARTS=0 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_boundary_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
LS_FOLD_LN=0 timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
