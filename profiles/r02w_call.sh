#!/bin/bash
# 4 GPUs (inner ranks with two neighbours on hardware): weak scaling with the default path (pull form, neighbour handshake, graph-replayed one-call steps), strong scaling (C5)
set -x
tag=${1:-r02w}; n=${2:-4}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench${n}_$2.json 2> gpurun_out/${tag}_bench${n}_$2.err; grep '^{' gpurun_out/${tag}_bench${n}_$2.json | cut -c1-160; tail -2 gpurun_out/${tag}_bench${n}_$2.err | cut -c1-300; }
run 29711 weak
run 29713 strong --scaling strong --steps 8
ls -la gpurun_out | tail -4
