#!/bin/bash
# periodic projection + pressure update as row sweeps (pw_project): parity + timing against the tile kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "periodic and not slab" 2>&1 | tail -4 > gpurun_out/pwp_tests.log
timeout 300 python scripts/quick_bench.py p,8192,8193 P,8192,8193 p,2048,2049 P,2048,2049 > gpurun_out/pwp_default.log 2>&1
RUSTPDE_B200_PW=1 timeout 300 python scripts/quick_bench.py P,2048,2049 P,512,513 > gpurun_out/pwp_pw1.log 2>&1
cat gpurun_out/pwp_tests.log gpurun_out/pwp_default.log gpurun_out/pwp_pw1.log
