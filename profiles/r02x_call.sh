#!/bin/bash
# round 2, second session, single GPU: A/B of the k_fwd build variants (gpurun_variants/*.so: host-computed frame bases, warps per CTA, persistent warps),
# the whole -m gpu suite on HEAD, the default bench line, the C3 / C4 config lines, launch list, ncu --set full of the two backward kernels
set -x
tag=${1:-r02x}
mkdir -p gpurun_out
for so in gpurun_variants/*.so; do
  echo "{\"variant\": \"$(basename $so .so)\"}" >> gpurun_out/${tag}_ab.jsonl
  FMPM_LIB="$PWD/$so" AB_STEPS=40 timeout 200 python profiles/fwd_ab.py 3 >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
done
cat gpurun_out/${tag}_ab.jsonl | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_all.log 2>&1; tail -6 gpurun_out/${tag}_pytest_all.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config C3 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; cut -c1-200 gpurun_out/${tag}_bench_c3.json; tail -3 gpurun_out/${tag}_bench_c3.err
timeout 300 python bench.py --config C4 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; cut -c1-200 gpurun_out/${tag}_bench_c4.json; tail -3 gpurun_out/${tag}_bench_c4.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_fwd.csv python bench.py --steps 3 --warmup 3 --bwd 0 --no-cpu > gpurun_out/${tag}_ncu_fwd.log 2>&1
PT_BWD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_particle_grad -s 4 -c 1 -f -o gpurun_out/${tag}_k_particle_grad python profiles/phase_times.py > gpurun_out/${tag}_ncu_pg.log 2>&1
PT_BWD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_g2p_grad_scatter -s 4 -c 1 -f -o gpurun_out/${tag}_k_g2p_grad_scatter python profiles/phase_times.py > gpurun_out/${tag}_ncu_sc.log 2>&1
tail -2 gpurun_out/${tag}_ncu_pg.log
ls -la gpurun_out | tail -12
