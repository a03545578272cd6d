#!/bin/bash
# round 2, second session, 2 GPUs: the six x-slab parity tests (forward x3, sharded backward x3), weak scaling with the pull / push form of the ghost
# reduction, strong scaling (C5), and ONE slab of the weak workload alone on one GPU (what the per-GPU workload costs without any exchange)
set -x
tag=${1:-r02y}; n=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "slab" > gpurun_out/${tag}_pytest_slab.log 2>&1; tail -8 gpurun_out/${tag}_pytest_slab.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench${n}_$2.json 2> gpurun_out/${tag}_bench${n}_$2.err; grep '^{' gpurun_out/${tag}_bench${n}_$2.json | cut -c1-160; tail -2 gpurun_out/${tag}_bench${n}_$2.err | cut -c1-300; }
run 29711 weak_pull
FMPM_SLAB_PULL=0 run 29712 weak_push
SLAB_MIGRATE_EVERY=100000 run 29714 weak_pull_nocensus
run 29713 strong --scaling strong --steps 8
timeout 300 python bench.py --slab-shape $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench1_slab_shape.json 2> gpurun_out/${tag}_bench1_slab_shape.err; cut -c1-160 gpurun_out/${tag}_bench1_slab_shape.json; tail -2 gpurun_out/${tag}_bench1_slab_shape.err
ls -la gpurun_out | tail -6
