#!/bin/bash
set -x
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "every_forward_path or fused or forward_phases or substep_grad_matches or ragged or no_used or out_of_grid" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
for v in mb5 mb6 mb4; do
  FMPM_LIB=$PWD/gpurun_variants/$v.so timeout 300 python profiles/fwd_ab.py -1 0 3 > gpurun_out/${tag}_ab_$v.json 2>> gpurun_out/${tag}_ab.err; cat gpurun_out/${tag}_ab_$v.json
done
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd_m3 python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_fwd_m3.log 2>&1
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_g2p2g -s 60 -c 1 -f -o gpurun_out/${tag}_k_g2p2g python profiles/fwd_ab.py 0 > gpurun_out/${tag}_ncu_k_g2p2g.log 2>&1
AB_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_op -s 60 -c 1 -f -o gpurun_out/${tag}_k_grid_op python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_grid_op.log 2>&1
ls -la gpurun_out | tail -5
