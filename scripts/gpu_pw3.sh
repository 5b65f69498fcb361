#!/bin/bash
# 2 GPUs: slab NCCL tests (both per-mode kernel families) + slab bench A/B
mkdir -p gpurun_out
python -m pytest tests/test_slab_nccl.py -q -x 2>&1 | tail -3 > gpurun_out/pw3_tests.log
RUSTPDE_B200_PW=1 python -m pytest tests/test_slab_nccl.py -q -x -k "512 or 2048" 2>&1 | tail -3 >> gpurun_out/pw3_tests.log
cat gpurun_out/pw3_tests.log
bash scripts/gpu_scale_ab.sh 2
