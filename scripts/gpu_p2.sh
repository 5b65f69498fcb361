#!/bin/bash
# power-of-two DCT period along x (nx = 2^k + 1): parity + timing next to the Bluestein grids of the same size
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pow2_period or tile_only or confined_specialised" 2>&1 | tail -4 > gpurun_out/p2_tests.log
timeout 300 python scripts/quick_bench.py c,1025,1025 c,1024,1025 c,2049,2049 C,2049,2049 c,2048,2049 > gpurun_out/p2_bench.log 2>&1
cat gpurun_out/p2_tests.log gpurun_out/p2_bench.log
