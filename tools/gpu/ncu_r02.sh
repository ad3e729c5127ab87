#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): (1) per-launch DRAM traffic of every conv launch of one forward,
# (2) launch list of the bench (shares), (3) one --set full capture of the dominant kernels.  Each ncu run follows a
# plain run of the same command that exited 0.
mkdir -p gpurun_out
TAG=${TAG:-r02}
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
python tools/ncu_conv_traffic.py run > gpurun_out/traffic_plain.log 2>&1 &&
N=$(python -c "import json; print(json.load(open('gpurun_out/conv_costs_r02.json'))['n'])") &&
ncu --metrics $M --clock-control none -k regex:k_conv_sm100 -s $((2 * N)) -c $N --csv --log-file gpurun_out/conv_traffic_r02.csv \
    python tools/ncu_conv_traffic.py run > gpurun_out/traffic_ncu.log 2>&1
tail -2 gpurun_out/traffic_plain.log; wc -l gpurun_out/conv_traffic_r02.csv
CMD="python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline --no-torch-gpu --profile-repeats 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -c 200 gpurun_out/plain.log; wc -l gpurun_out/launches_$TAG.csv
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_conv_sm100|k_attention_sm100|k_gn_apply|k_ddim" -s 1500 -c 60 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2> gpurun_out/ncu_export.log
ls -la /tmp/prof_$TAG.ncu-rep gpurun_out | tail -6
