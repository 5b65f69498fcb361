#!/bin/bash
mkdir -p gpurun_out
python scripts/e2e_diag.py > gpurun_out/r2b_e2e_diag.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_slab2.json 2> gpurun_out/r2b_slab2.err
python -m pytest tests/test_slab_nccl.py -q 2>&1 | tail -5 > gpurun_out/r2b_slab_tests.log
cat gpurun_out/r2b_e2e_diag.log; tail -3 gpurun_out/r2b_slab2.err; cut -c1-1500 gpurun_out/r2b_slab2.json; cat gpurun_out/r2b_slab_tests.log
