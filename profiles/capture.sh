#!/bin/bash
# Evidence capture on the GPU box (one GPU).  Writes into gpurun_out/; summaries are copied into profiles/ afterwards.
#   bash profiles/capture.sh <tag>
set -x
tag=${1:-r01b}
mkdir -p gpurun_out
# 1. bench, both arms (numbers quoted in profiles/README.md)
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
# 2. launch list of a short forward-only bench (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_fwd.csv python bench.py --steps 3 --warmup 3 --bwd 0 --no-cpu > gpurun_out/${tag}_ncu_fwd.log 2>&1
# 3. full capture of the dominant kernel (one launch, after the sort and the warm-up steps)
ncu --set full --clock-control none --import-source on -k regex:k_p2g -s 40 -c 1 -f -o gpurun_out/${tag}_k_p2g python bench.py --steps 3 --warmup 3 --bwd 0 --no-cpu > gpurun_out/${tag}_ncu_p2g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_g2p -s 40 -c 1 -f -o gpurun_out/${tag}_k_g2p python bench.py --steps 3 --warmup 3 --bwd 0 --no-cpu > gpurun_out/${tag}_ncu_g2p.log 2>&1
ls -la gpurun_out | tail -8
