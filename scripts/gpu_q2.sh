#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "confined_specialised or (periodic and not slab and not row_sweeps)" 2>&1 | tail -3 > gpurun_out/q2_tests.log
timeout 300 python scripts/quick_bench.py c,2048,2049 C,2048,2049 p,8192,8193 P,8192,8193 > gpurun_out/q2_bench.log 2>&1
cat gpurun_out/q2_tests.log gpurun_out/q2_bench.log
