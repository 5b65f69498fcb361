#!/bin/bash
# round-2 profile session: launch list of the bench command + ncu --set full captures of the top kernels (one update(), eager)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_confined2048.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
for k in xk_forward xk_backward yk_backward yk_adi yk_mode dgemm_dmma yk_conv xk_project yk_project; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r2_$k \
      python scripts/ncu_step.py c 2048 2049 1 > gpurun_out/r2_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2_$k.ncu-rep > gpurun_out/r2_ncu_${k}_confined2048.txt 2>> gpurun_out/r2_ncu_$k.log
  rm -f gpurun_out/r2_$k.ncu-rep
done
for k in xs_rhs_adi xs_project; do
  RUSTPDE_B200_XS=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r2_$k \
      python scripts/ncu_step.py c 2048 2049 1 > gpurun_out/r2_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2_$k.ncu-rep > gpurun_out/r2_ncu_${k}_confined2048.txt 2>> gpurun_out/r2_ncu_$k.log
  rm -f gpurun_out/r2_$k.ncu-rep
done
for k in pk_hholtz pk_divpois; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/r2_$k \
      python scripts/ncu_step.py p 2048 2049 1 > gpurun_out/r2_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2_$k.ncu-rep > gpurun_out/r2_ncu_${k}_periodic2048.txt 2>> gpurun_out/r2_ncu_$k.log
  rm -f gpurun_out/r2_$k.ncu-rep
done
ls -la gpurun_out/r2_ncu_*_*.txt; head -12 gpurun_out/r2_ncu_xk_forward_confined2048.txt
