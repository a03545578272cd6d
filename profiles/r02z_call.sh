#!/bin/bash
# round 2, second session, single GPU: A/B of k_fwd CTA shapes at 80 registers (3 warps x 7 / x 8, 2 x 12, 3 x 6 at 96 registers), in-kernel vs host frame bases,
# backward kernels with / without the L2 prefetch, the parity tests the changed kernels touch, the final bench line, ncu --set full of the final k_fwd
set -x
tag=${1:-r02z}
mkdir -p gpurun_out
for so in gpurun_variants/*.so; do
  echo "{\"variant\": \"$(basename $so .so)\"}" >> gpurun_out/${tag}_ab.jsonl
  FMPM_LIB="$PWD/$so" AB_STEPS=40 timeout 200 python profiles/fwd_ab.py 3 >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
done
cat gpurun_out/${tag}_ab.jsonl | cut -c1-120
for v in a_w3m7 f_nopf; do FMPM_LIB="$PWD/gpurun_variants/$v.so" PT_BWD=1 timeout 200 python profiles/phase_times.py 2>> gpurun_out/${tag}_ab.err | tail -1 >> gpurun_out/${tag}_bwd_ab.txt; done
cat gpurun_out/${tag}_bwd_ab.txt
timeout 900 python -m pytest tests -m gpu -q -k "substep_grad or dloss or jetbot or c3_latteart or c2_full or every_forward or rigid or reference_agents or reference_kernels" > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
AB_STEPS=6 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fwd -s 60 -c 1 -f -o gpurun_out/${tag}_k_fwd python profiles/fwd_ab.py 3 > gpurun_out/${tag}_ncu_k_fwd.log 2>&1
ls -la gpurun_out | tail -8
