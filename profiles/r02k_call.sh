#!/bin/bash
set -x
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "every_forward_path or fused or c2_full_size" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python profiles/fwd_ab.py 3 11 > gpurun_out/${tag}_ab.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab.json
FMPM_TMA=0 timeout 300 python profiles/fwd_ab.py 11 > gpurun_out/${tag}_ab_tma0.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab_tma0.json
AB_SORT=2 timeout 300 python profiles/fwd_ab.py 11 > gpurun_out/${tag}_ab_sort2.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab_sort2.json
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd python profiles/fwd_ab.py 11 > gpurun_out/${tag}_ncu_k_fwd.log 2>&1
tail -3 gpurun_out/${tag}_ab.err
ls -la gpurun_out | tail -3
