#!/bin/bash
# final validation of the round's HEAD, single GPU: the whole -m gpu suite, then the default bench line (e2e with the staged uploads)
set -x
tag=${1:-r02final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/${tag}_pytest_all.log 2>&1; tail -12 gpurun_out/${tag}_pytest_all.log
timeout 500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
