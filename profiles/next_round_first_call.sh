#!/bin/bash
# Everything queued at the end of round 1 (profiles/README.md, "Queued for the next GPU round") in ONE gpurun call, so that box
# acquisition is paid once.  One GPU; ~10 GPU-minutes.  Writes into gpurun_out/ (copy summaries into profiles/ afterwards).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/next_round_first_call.sh r02a'
# The 2-GPU items (x-slab backward parity) need a separate `gpurun --gpus 2` call:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'python -m pytest tests/test_gpu_parity.py -m gpu -q -k slab_sharded > gpurun_out/r02a_slab2.log 2>&1'
set -x
tag=${1:-r02a}
mkdir -p gpurun_out
# 1. the parity gate, incl. the tests that have never run on hardware (collected last, tests/conftest.py: smoke solver, circulation stack, reference
#    agent scenes, g2p2g paths incl. MAT_RIGID bodies, device Adam, 'locked' reference run)
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${tag}_pytest_gpu.log
# 2. bench: plain vs g2p2g-fused forward (same process conditions back to back); the line also carries e2e vs e2e_obs_bridge
python bench.py > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err
python bench.py --fuse-g2p2g 1 > gpurun_out/${tag}_bench_fused.json 2> gpurun_out/${tag}_bench_fused.err
python bench.py --fuse-g2p2g 1 --sort-every 2 --bwd 0 --no-cpu > gpurun_out/${tag}_bench_fused_sort2.json 2> /dev/null   # the fused scatter sees older cell orders: does a shorter sort period pay?
# 3. smoke solver phase timings (50 and 500 Jacobi sweeps at 128^3)
python profiles/smoke_times.py > gpurun_out/${tag}_smoke_times.json 2> gpurun_out/${tag}_smoke_times.err
python profiles/c1_rollout_times.py > gpurun_out/${tag}_c1_rollout.json 2> gpurun_out/${tag}_c1_rollout.err
# 4. launch list + full capture of the fused kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_fused.csv python bench.py --steps 3 --warmup 3 --bwd 0 --no-cpu --fuse-g2p2g 1 > gpurun_out/${tag}_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_g2p2g -s 40 -c 1 -f -o gpurun_out/${tag}_k_g2p2g python bench.py --steps 3 --warmup 3 --bwd 0 --no-cpu --fuse-g2p2g 1 > gpurun_out/${tag}_ncu_g2p2g.log 2>&1
ls -la gpurun_out | tail -12
