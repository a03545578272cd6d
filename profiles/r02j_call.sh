#!/bin/bash
set -x
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "substep_grad_matches or c2_full_size or c3_latteart or full_size_directional or dloss_daction" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 900 python bench.py --scaling strong --steps 10 > gpurun_out/${tag}_bench_strong1.json 2> gpurun_out/${tag}_bench_strong1.err; cut -c1-400 gpurun_out/${tag}_bench_strong1.json; tail -3 gpurun_out/${tag}_bench_strong1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
ls -la gpurun_out | tail -3
