#!/bin/bash
# usage: gpu_scale.sh N [tag]   -- the driver's own launch line for bench.py at N GPUs (slab-decomposed periodic 8192x8193)
N=$1
T=${2:-r3}
mkdir -p gpurun_out
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${T}_scale_$N.json 2> gpurun_out/${T}_scale_$N.err
tail -3 gpurun_out/${T}_scale_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_scale_$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','speedup_vs_1gpu','ms_per_step_1gpu','slab_parity','clocks')})
print(d['e2e']); print(d['roofline'])
PY
