#!/bin/bash
# session 37: whole-GPU sparse schedule inside the per-iteration CUDA graph
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_sparse_ldlt.py tests/test_gpu_sparse_cond.py tests/test_gpu_dense_ldlt.py tests/test_gpu_mm_small.py -m gpu -x -q > gpurun_out/s37_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s37_pytest.log
timeout 300 python bench.py --workload sparse_c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s37_bench_c3.json 2> gpurun_out/s37_bench_c3.err
B200_SUITE_THREADS=12 timeout 600 python tools/mm_suite.py > gpurun_out/s37_mm_suite.json 2> gpurun_out/s37_mm_suite.err
tail -n 3 gpurun_out/s37_pytest.log; tail -n 3 gpurun_out/s37_mm_suite.err
