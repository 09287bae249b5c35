#!/bin/bash
# session 30 (2 GPUs): the default bench line under torchrun, both arms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/s30_bench_2gpu.json 2> gpurun_out/s30_bench_2gpu.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1) > gpurun_out/s30_bench_ref_2gpu.json 2> gpurun_out/s30_bench_ref_2gpu.err
tail -3 gpurun_out/s30_bench_2gpu.err
