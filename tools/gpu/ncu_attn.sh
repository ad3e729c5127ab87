#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/run_attn_once.py 4096 128"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_attention_sm100" -s 1 -c 1 -o /tmp/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
ncu -i /tmp/prof_attn.ncu-rep --page source --csv > gpurun_out/prof_attn128_source.csv 2> gpurun_out/ncu_export.log
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/prof_attn128_raw.csv 2>> gpurun_out/ncu_export.log
ls -la gpurun_out/prof_attn128*
