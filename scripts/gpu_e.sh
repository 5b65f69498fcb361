#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2e_tests.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2e_ref.json 2> gpurun_out/r2e_ref.err
for w in confined64 hholtz1024 periodic512 periodic2048 confined1024; do
  python bench.py --workload $w > gpurun_out/r2e_$w.json 2> gpurun_out/r2e_$w.err
done
python bench.py --impl reference --workload periodic512 --steps 20 > gpurun_out/r2e_ref_p512.json 2> gpurun_out/r2e_ref_p512.err
tail -4 gpurun_out/r2e_tests.log
python - <<'PY'
import json
for f in ['r2e_bench','r2e_ref','r2e_confined64','r2e_hholtz1024','r2e_periodic512','r2e_periodic2048','r2e_confined1024','r2e_ref_p512']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],3), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value'],2), 'cpu', d.get('cpu_baseline') and (round(d['cpu_baseline']['value'],3), d['cpu_baseline']['cores']))
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
