#!/bin/bash
# GEMM A/B (m16n8k16 vs m8n8k4) + tensor-solve parity
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "fast_diag or standalone or confined_specialised" 2>&1 | tail -4 > gpurun_out/q_tests.log
RUSTPDE_B200_DMMA=884 python scripts/quick_bench.py C,2048,2049 2>&1 | grep -E "total|gemm" > gpurun_out/q.log
python scripts/quick_bench.py c,2048,2049 C,2048,2049 2>&1 | grep -E "steps/s|total|gemm" >> gpurun_out/q.log
cat gpurun_out/q_tests.log gpurun_out/q.log
