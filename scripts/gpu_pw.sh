#!/bin/bash
# warp-serial row sweeps of the periodic path (fast_pw.cu): parity + timing against the tile kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "periodic" 2>&1 | tail -4 > gpurun_out/pw_tests.log
for pw in 0 1; do
  RUSTPDE_B200_PW=$pw python scripts/quick_bench.py p,512,513 p,2048,2049 P,2048,2049 > gpurun_out/pw_$pw.log 2>&1
done
python bench.py --workload periodic8192 --steps 10 --no-cpu-baseline > gpurun_out/pw_p8192.json 2> gpurun_out/pw_p8192.err
cat gpurun_out/pw_tests.log gpurun_out/pw_0.log gpurun_out/pw_1.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/pw_p8192.json').read().strip().splitlines()[-1])
print('p8192', round(d['value'],2), 'ms', round(d['ms_per_step'],3), [(k['kernel'],round(k['ms'],3)) for k in (d['roofline'].get('per_kernel') or [])])
PY
