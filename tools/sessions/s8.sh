#!/bin/bash
# GPU session 8: partition kernels with common-size ring slots; sparse: prefetching left-looking fronts + pipelined extend-add
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_multistage.py tests/test_gpu_full_size.py tests/test_gpu_sparse_ldlt.py tests/test_gpu_sparse_cond.py tests/test_gpu_mm_small.py -m gpu -q > gpurun_out/s8_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/s8_pytest.log
B200_MS_TIMING=1 timeout 300 python bench.py --workload multistage --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/s8_ms_timing.json 2> gpurun_out/s8_ms_timing.err
timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s8_bench_ms.json 2> gpurun_out/s8_bench_ms.err
B200_MS_NO_PARTITION=1 timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s8_bench_ms_nopart.json 2> gpurun_out/s8_bench_ms_nopart.err
for v in left right; do
  B200_MF_PROF=1 B200_MF_BIG=$v timeout 300 python bench.py --workload sparse --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s8_bench_sparse_$v.json 2> gpurun_out/s8_bench_sparse_$v.err
done
