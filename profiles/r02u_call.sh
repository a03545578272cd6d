#!/bin/bash
# 2-GPU call: the six x-slab parity tests, then the weak-scaling bench with the pull / push form of the ghost reduction and the strong-scaling (C5) bench
set -x
tag=${1:-r02u}; n=${2:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "slab" > gpurun_out/${tag}_pytest_slab.log 2>&1; tail -15 gpurun_out/${tag}_pytest_slab.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench${n}_$2.json 2> gpurun_out/${tag}_bench${n}_$2.err; grep '^{' gpurun_out/${tag}_bench${n}_$2.json | cut -c1-160; tail -2 gpurun_out/${tag}_bench${n}_$2.err | cut -c1-300; }
run 29711 weak_pull
FMPM_SLAB_PULL=0 run 29712 weak_push
run 29713 strong --scaling strong --steps 8
ls -la gpurun_out | tail -4
