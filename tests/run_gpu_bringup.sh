#!/bin/bash
# Bring-up helper: runs the GPU test groups in separate processes so one CUDA fault cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi -L
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -${TAILN:-25} | tee gpurun_out/$name.log; }
run k_misc tests/test_gpu_kernels.py -k "groupnorm or upsample or timestep or success"
run k_conv_simt tests/test_gpu_kernels.py -k "conv and simt"
run k_attn_simt tests/test_gpu_kernels.py -k "attention and simt"
run k_conv_sm100 tests/test_gpu_kernels.py -k "conv and sm100"
run k_attn_sm100 tests/test_gpu_kernels.py -k "attention and (sm100 or lazy)"
run parity_simt tests/test_gpu_parity.py -k "not sm100" -s
run parity_sm100 tests/test_gpu_parity.py -k "sm100" -s
