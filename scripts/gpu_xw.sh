#!/bin/bash
# parity + timing of the warp-serial column sweeps (fast_xw.cu)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "confined" 2>&1 | tail -4 > gpurun_out/xw_tests.log
timeout 120 python scripts/quick_bench.py c,2048,2049 C,2048,2049 > gpurun_out/xw_on.log 2>&1
cat gpurun_out/xw_tests.log gpurun_out/xw_on.log
