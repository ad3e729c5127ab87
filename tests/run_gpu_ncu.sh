#!/bin/bash
# ncu launch list + one --set full capture of the dominant kernels (B200_PROFILING.md recipe).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/plain.log
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_conv_sm100|k_attention_sm100|k_gn_apply|k_gn_partial" -s 450 -c 60 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
