#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r2s_gputests.log
