#!/bin/bash
# A/B of the warp-serial column sweeps (fast_xw.cu) against the tile kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "confined" 2>&1 | tail -4 > gpurun_out/xw_tests.log
RUSTPDE_B200_XW=0 python scripts/quick_bench.py c,2048,2049 C,2048,2049 > gpurun_out/xw_off.log 2>&1
python scripts/quick_bench.py c,2048,2049 C,2048,2049 > gpurun_out/xw_on.log 2>&1
cat gpurun_out/xw_tests.log gpurun_out/xw_off.log gpurun_out/xw_on.log
