#!/bin/bash
set -x
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "every_forward_path or fused or forward_phases or substep_grad_matches or ragged or no_used or out_of_grid or reference_kernels or rigid_material" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python profiles/fwd_ab.py -1 0 1 3 7 > gpurun_out/${tag}_fwd_ab.json 2> gpurun_out/${tag}_fwd_ab.err; cat gpurun_out/${tag}_fwd_ab.json
AB_SORT=2 timeout 300 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_fwd_ab_sort2.json 2>> gpurun_out/${tag}_fwd_ab.err
AB_SORT=8 timeout 300 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_fwd_ab_sort8.json 2>> gpurun_out/${tag}_fwd_ab.err
cat gpurun_out/${tag}_fwd_ab_sort*.json
timeout 600 python bench.py --fuse-g2p2g 1 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-600 gpurun_out/${tag}_bench.json
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd_m3 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_fwd_m3.log 2>&1
ls -la gpurun_out | tail -5
