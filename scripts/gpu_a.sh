#!/bin/bash
# round-2 GPU session A: parity tests, bench N=1 (group barriers on/off), config-2 and config-3 workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_tests.log
python bench.py --steps 200 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
RUSTPDE_B200_KFLAGS=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_k1.json 2> gpurun_out/r2a_bench_k1.err
python bench.py --workload hholtz1024 > gpurun_out/r2a_hholtz.json 2> gpurun_out/r2a_hholtz.err
python bench.py --workload periodic512 --steps 1000 --no-cpu-baseline > gpurun_out/r2a_p512.json 2> gpurun_out/r2a_p512.err
python bench.py --workload periodic2048 --steps 200 --no-cpu-baseline > gpurun_out/r2a_p2048.json 2> gpurun_out/r2a_p2048.err
tail -5 gpurun_out/r2a_tests.log
cat gpurun_out/r2a_bench.json | cut -c1-600
