#!/bin/bash
set -x
tag=${1:-r02n}
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_all.log 2>&1; tail -8 gpurun_out/${tag}_pytest_all.log
timeout 900 python bench.py --scaling strong --steps 8 > gpurun_out/${tag}_bench_strong1.json 2> gpurun_out/${tag}_bench_strong1.err; cut -c1-300 gpurun_out/${tag}_bench_strong1.json; tail -3 gpurun_out/${tag}_bench_strong1.err
ls -la gpurun_out | tail -3
