#!/bin/bash
# end-of-round captures of the remaining kernels of the confined step (the four largest are in gpu_final.sh)
mkdir -p gpurun_out
T=${1:-r4}
for k in dgemm_dmma yk_conv yk_adi yk_mode yk_project yk_pres xk_div xk_project; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/${T}_$k \
      python scripts/ncu_step.py c 2048 2049 1 > gpurun_out/${T}_ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${T}_$k.ncu-rep > gpurun_out/${T}_ncu_${k}_confined2048.txt 2>> gpurun_out/${T}_ncu_$k.log
  rm -f gpurun_out/${T}_$k.ncu-rep
done
for k in dgemm_dmma yk_adi yk_mode; do head -14 gpurun_out/${T}_ncu_${k}_confined2048.txt | cut -c1-150; done
