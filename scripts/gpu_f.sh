#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "periodic or slab or adi or fast_diag or standalone" 2>&1 | tail -5 > gpurun_out/r2f_tests.log
for w in confined2048 periodic2048 periodic512; do
  python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2f_$w.json 2> gpurun_out/r2f_$w.err
done
python bench.py --workload periodic8192 --steps 20 --no-cpu-baseline > gpurun_out/r2f_p8192_1gpu.json 2> gpurun_out/r2f_p8192_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2f_slab2.json 2> gpurun_out/r2f_slab2.err
tail -3 gpurun_out/r2f_tests.log
python - <<'PY'
import json
for f in ['r2f_confined2048','r2f_periodic2048','r2f_periodic512','r2f_p8192_1gpu','r2f_slab2']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), 'ms', round(d['ms_per_step'],4), d.get('speedup_vs_1gpu'), d.get('ms_per_step_1gpu'), [(k['kernel'],k['ms']) for k in (d['roofline'].get('per_kernel') or [])])
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
