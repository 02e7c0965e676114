#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_calib.py -q -s -p no:cacheprovider 2>&1 | tail -80 | tee gpurun_out/h_calib_tests.log
