#!/bin/bash
set -x
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "every_forward_path or fused or forward_phases or substep_grad_matches or ragged or no_used or out_of_grid or sdf_colliders or reference_agents" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
FMPM_PDL=1 timeout 300 python profiles/fwd_ab.py -1 0 3 > gpurun_out/${tag}_ab_pdl1.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab_pdl1.json
FMPM_PDL=0 timeout 300 python profiles/fwd_ab.py -1 0 3 > gpurun_out/${tag}_ab_pdl0.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab_pdl0.json
AB_SORT=2 timeout 300 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ab_sort2.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab_sort2.json
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd_m3 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_fwd_m3.log 2>&1
AB_STEPS=6 timeout 600 ncu --set full --clock-control none -k regex:k_grid_op -s 60 -c 1 -f -o gpurun_out/${tag}_k_grid_op python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_grid_op.log 2>&1
ls -la gpurun_out | tail -5
