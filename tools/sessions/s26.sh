#!/bin/bash
# session 26: bulk DMMA kernel on partial tiles and in the sparse far updates; full GPU test pass; full default bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s26_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s26_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s26_bench_all.json 2> gpurun_out/s26_bench_all.err
B200_GEMM_BULK=0 timeout 300 python bench.py --workload sparse_c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s26_bench_c3_cpasync.json 2> gpurun_out/s26_bench_c3_cpasync.err
tail -3 gpurun_out/s26_pytest.log
