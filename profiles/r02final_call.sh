#!/bin/bash
# final validation of the round's HEAD on a 2-GPU box: the WHOLE -m gpu suite (the six 2-GPU x-slab tests included)
set -x
tag=${1:-r02final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/${tag}_pytest_all_2gpu.log 2>&1; tail -8 gpurun_out/${tag}_pytest_all_2gpu.log
