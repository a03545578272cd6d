#!/bin/bash
# 2-GPU call: the six x-slab parity tests (forward peer / signal / fused, backward peer / nccl / peer-signal) + the 2-GPU bench arm
set -x
tag=${1:-r02i}
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "slab" > gpurun_out/${tag}_pytest_slab.log 2>&1; tail -15 gpurun_out/${tag}_pytest_slab.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --bwd 0 "${@:3}" > gpurun_out/${tag}_bench2_$2.json 2> gpurun_out/${tag}_bench2_$2.err; cut -c1-300 gpurun_out/${tag}_bench2_$2.json; tail -3 gpurun_out/${tag}_bench2_$2.err; }
run 29611 signal_fused
SLAB_SYNC=barrier run 29612 barrier_fused
run 29613 signal_plain --fuse-g2p2g 0
ls -la gpurun_out | tail -4
