#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/gpu_suite.log
echo "=== decision agreement"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "decisions_agree or teacher" -s 2>&1 | grep -E "agreement|teacher|passed|failed" | tee gpurun_out/agree.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
rm -f gpurun_out/sweep.jsonl
echo "=== sweep dm2"; timeout 900 python tools/sweep.py --model dm2 --batches 1,8,64 --sizes 64,128,256 --steps 10,50 2>&1 | tail -20
echo "=== attack resnet50"; timeout 600 python tools/attack_bench.py --victim resnet50 --images 32 2>&1 | tail -2 | tee gpurun_out/attack_resnet50.log
echo "=== attack vit K=8"; timeout 600 python tools/attack_bench.py --victim vit_b_16 --images 8 --candidates 8 2>&1 | tail -2 | tee gpurun_out/attack_vit.log
