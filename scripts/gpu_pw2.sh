#!/bin/bash
# periodic tests with both per-mode kernel families + benches with the row-count heuristic
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "periodic" 2>&1 | tail -4 > gpurun_out/pw2_tests.log
python scripts/quick_bench.py p,512,513 p,2048,2049 P,2048,2049 > gpurun_out/pw2_q.log 2>&1
cat gpurun_out/pw2_tests.log; grep -E "steps/s|hholtz|poisson" gpurun_out/pw2_q.log
