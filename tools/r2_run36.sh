#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_whisper_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s 2>&1 | tail -40
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "pixel_pre" -s 2>&1 | tail -15
