#!/bin/bash
# multi-GPU bench call: bash profiles/scale_call.sh TAG N   (weak scaling with halo 4 / 2, strong scaling C5)
set -x
tag=$1; n=$2
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench${n}_$2.json 2> gpurun_out/${tag}_bench${n}_$2.err; grep '^{' gpurun_out/${tag}_bench${n}_$2.json | cut -c1-200; tail -2 gpurun_out/${tag}_bench${n}_$2.err | cut -c1-300; }
run 29711 weak
SLAB_HALO=2 run 29712 weak_halo2
run 29713 strong --scaling strong --steps 10
if [ "$n" = "2" ]; then
  SLAB_SELFCHECK=1 SLAB_MODE=backward SLAB_EXCHANGE=peer SLAB_SYNC=signal timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/run_slab_gpu.py > gpurun_out/${tag}_slab_backward_selfcheck.log 2>&1; grep -E "slab backward|selfcheck|SLAB_" gpurun_out/${tag}_slab_backward_selfcheck.log
fi
ls -la gpurun_out | tail -4
