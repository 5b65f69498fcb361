#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c_tests.log
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
RUSTPDE_B200_NO_XS=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_noxs.json 2> gpurun_out/r2c_bench_noxs.err
tail -4 gpurun_out/r2c_tests.log
python - <<'PY'
import json
for f in ['r2c_bench','r2c_bench_noxs']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], [(k['kernel'],k['ms']) for k in d['roofline']['per_kernel']])
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
