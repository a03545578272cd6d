#!/bin/bash
set -x
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "every_forward_path or fused" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python profiles/fwd_ab.py 0 1 3 5 7 > gpurun_out/${tag}_fwd_ab.json 2> gpurun_out/${tag}_fwd_ab.err; cat gpurun_out/${tag}_fwd_ab.json
AB_SORT=1 timeout 300 python profiles/fwd_ab.py 7 3 > gpurun_out/${tag}_fwd_ab_sort1.json 2>> gpurun_out/${tag}_fwd_ab.err
AB_SORT=2 timeout 300 python profiles/fwd_ab.py 7 3 > gpurun_out/${tag}_fwd_ab_sort2.json 2>> gpurun_out/${tag}_fwd_ab.err
cat gpurun_out/${tag}_fwd_ab_sort*.json
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd python profiles/fwd_ab.py 7 > gpurun_out/${tag}_ncu_k_fwd.log 2>&1
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd_m3 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_fwd_m3.log 2>&1
ls -la gpurun_out | tail -8
