#!/bin/bash
# round-2 closing session on one B200: the full -m gpu suite, the bench lines of every single-GPU workload (both arms for
# the default one), the ncu launch list of the bench command and ncu --set full captures of the top kernels.
# Outputs: gpurun_out/r3_*  (summaries are copied to profiles/ by hand)
mkdir -p gpurun_out
T=${1:-r3}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for w in periodic2048 periodic512 confined64 confined1024 hholtz1024; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
done
timeout 300 python bench.py --workload periodic8192 --steps 20 --no-cpu-baseline > gpurun_out/${T}_p8192_1gpu.json 2> gpurun_out/${T}_p8192_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_confined2048.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
for k in xk_backward xk_forward xw_adi yk_backward; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/${T}_$k \
      python scripts/ncu_step.py c 2048 2049 1 > gpurun_out/${T}_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${T}_$k.ncu-rep > gpurun_out/${T}_ncu_${k}_confined2048.txt 2>> gpurun_out/${T}_ncu_$k.log
  rm -f gpurun_out/${T}_$k.ncu-rep
done
for k in pw_hholtz pw_divpois pw_project; do
  RUSTPDE_B200_PW=1 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/${T}_$k \
      python scripts/ncu_step.py p 2048 2049 1 > gpurun_out/${T}_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${T}_$k.ncu-rep > gpurun_out/${T}_ncu_${k}_periodic2048.txt 2>> gpurun_out/${T}_ncu_$k.log
  rm -f gpurun_out/${T}_$k.ncu-rep
done
cat gpurun_out/${T}_tests.log; tail -4 gpurun_out/${T}_smoke.log
python - <<PY
import json
for f in ['bench_default','bench_reference','periodic2048','periodic512','confined64','confined1024','hholtz1024','p8192_1gpu']:
    try:
        d=json.loads([l for l in open('gpurun_out/${T}_%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],2), d.get('unit'), 'ms', round(d.get('ms_per_step',0),4), 'e2e', (d.get('e2e') or {}).get('value'), 'res', (d.get('e2e_resident') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/${T}_%s.err'%f).read()[-800:])
PY
