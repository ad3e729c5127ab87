#!/bin/bash
# ncu launch list + one --set full capture of the dominant kernels (B200_PROFILING.md recipe).
# The .ncu-rep stays on the box (too large for gpurun_out); CSV exports come back.
mkdir -p gpurun_out
TAG=${TAG:-r01b}
CMD="python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 340 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -c 300 gpurun_out/plain.log
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_conv_sm100|k_attention_sm100|k_gn_apply" -s 420 -c 40 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2> gpurun_out/ncu_export.log
ls -la /tmp/prof_$TAG.ncu-rep gpurun_out | tail -8
