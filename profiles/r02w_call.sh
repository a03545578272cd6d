#!/bin/bash
# 8-GPU call: weak scaling (pull form of the ghost reduction, default), strong scaling (C5), weak scaling with the push form for comparison
set -x
tag=${1:-r02w}; n=${2:-8}
mkdir -p gpurun_out
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench${n}_$2.json 2> gpurun_out/${tag}_bench${n}_$2.err; grep '^{' gpurun_out/${tag}_bench${n}_$2.json | cut -c1-160; tail -2 gpurun_out/${tag}_bench${n}_$2.err | cut -c1-300; }
run 29711 weak_pull
run 29713 strong --scaling strong --steps 8
FMPM_SLAB_PULL=0 run 29712 weak_push
ls -la gpurun_out | tail -4
