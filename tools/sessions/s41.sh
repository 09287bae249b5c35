#!/bin/bash
# session 41 (4 GPUs): the default bench line under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3) > gpurun_out/s41_bench_4gpu.json 2> gpurun_out/s41_bench_4gpu.err
tail -n 4 gpurun_out/s41_bench_4gpu.err
