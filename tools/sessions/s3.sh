#!/bin/bash
# GPU session 3: partition with resource-aware K; first run of the multi-workload bench and the full MM-shaped suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python -m pytest tests/test_gpu_multistage.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/s3_pytest_ms.log 2>&1
echo "rc=$?" >> gpurun_out/s3_pytest_ms.log
for k in default 3 4 5 6 8; do
  if [ $k = default ]; then unset B200_MS_SEGMENTS; else export B200_MS_SEGMENTS=$k; fi
  timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s3_bench_ms_K$k.json 2> gpurun_out/s3_bench_ms_K$k.err
done
unset B200_MS_SEGMENTS
B200_MS_NO_PARTITION=1 timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s3_bench_ms_nopart.json 2> gpurun_out/s3_bench_ms_nopart.err
( time timeout 900 python bench.py --workload mm_suite ) > gpurun_out/s3_bench_mm_suite.json 2> gpurun_out/s3_bench_mm_suite.err
( time timeout 1200 python bench.py --steps 5 --warmup 3 ) > gpurun_out/s3_bench_all.json 2> gpurun_out/s3_bench_all.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/s3_bench_ref.json 2> gpurun_out/s3_bench_ref.err
