#!/bin/bash
# usage: gpu_scale_ab.sh N  -- slab bench at N GPUs with the per-mode tile kernels (PW=0) and the row sweeps (PW=1), per-phase times
N=$1
mkdir -p gpurun_out
for pw in 0 1; do
  RUSTPDE_B200_PW=$pw NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$pw \
      bench.py --gpus $N --steps 15 --warmup 3 > gpurun_out/r3_scale_${N}_pw$pw.json 2> gpurun_out/r3_scale_${N}_pw$pw.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r3_scale_${N}_pw$pw.json').read().strip().splitlines()[-1])
print('PW=$pw', {k:d[k] for k in ('value','ms_per_step','n_gpus','speedup_vs_1gpu','ms_per_step_1gpu')}, d['roofline'].get('per_phase_ms_rank0'), d['slab_parity']['max_rel_err'])
PY
done
