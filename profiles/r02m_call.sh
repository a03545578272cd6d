#!/bin/bash
set -x
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest_all.log 2>&1; tail -5 gpurun_out/${tag}_pytest_all.log
timeout 900 python profiles/check_c5.py 40 > gpurun_out/${tag}_check_c5.json 2> gpurun_out/${tag}_check_c5.err; cat gpurun_out/${tag}_check_c5.json; tail -2 gpurun_out/${tag}_check_c5.err
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
ls -la gpurun_out | tail -3
