#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_tests.log
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
RUSTPDE_B200_XS=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_xs0.json 2> gpurun_out/r2d_bench_xs0.err
RUSTPDE_B200_XS=1 RUSTPDE_B200_XS_CFG=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_xs1.json 2> gpurun_out/r2d_bench_xs1.err
tail -4 gpurun_out/r2d_tests.log
python - <<'PY'
import json
for f in ['r2d_bench','r2d_bench_xs0','r2d_bench_xs1']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], [(k['kernel'],k['ms']) for k in d['roofline']['per_kernel']])
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
