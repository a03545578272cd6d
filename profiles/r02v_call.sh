#!/bin/bash
# 2 GPUs: the six x-slab parity tests with the handshake inside k_grid_op_pull (default), weak scaling with it on / off
set -x
tag=${1:-r02v}; n=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "slab" > gpurun_out/${tag}_pytest_slab.log 2>&1; tail -4 gpurun_out/${tag}_pytest_slab.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench${n}_$2.json 2> gpurun_out/${tag}_bench${n}_$2.err; grep '^{' gpurun_out/${tag}_bench${n}_$2.json | cut -c1-160; tail -2 gpurun_out/${tag}_bench${n}_$2.err | cut -c1-300; }
run 29711 weak_fsync1
FMPM_SLAB_FSYNC=0 run 29712 weak_fsync0
run 29713 strong_fsync1 --scaling strong --steps 8
ls -la gpurun_out | tail -5
