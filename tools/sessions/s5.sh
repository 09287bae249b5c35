#!/bin/bash
# GPU session 5: in-situ timing of the partitioned multistage kernels; .mat suites; tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python -m pytest tests/test_gpu_multistage.py tests/test_gpu_full_size.py::test_config4_full_size_batch_matches_oracle -m gpu -x -q > gpurun_out/s5_pytest_ms.log 2>&1
echo "rc=$?" >> gpurun_out/s5_pytest_ms.log
for k in 3 5 8; do
  B200_MS_TIMING=1 B200_MS_SEGMENTS=$k timeout 300 python bench.py --workload multistage --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/s5_ms_timing_K$k.json 2> gpurun_out/s5_ms_timing_K$k.err
done
for k in default 4; do
  if [ $k = default ]; then unset B200_MS_SEGMENTS; else export B200_MS_SEGMENTS=$k; fi
  timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s5_bench_ms_K$k.json 2> gpurun_out/s5_bench_ms_K$k.err
done
unset B200_MS_SEGMENTS
B200_MS_NO_PARTITION=1 timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s5_bench_ms_nopart.json 2> gpurun_out/s5_bench_ms_nopart.err
timeout 1500 python tools/mat_suite.py run --suite netlib_feas,netlib_infeas --out gpurun_out/r02_mat_suite_netlib.json > gpurun_out/s5_mat_netlib.log 2>&1
timeout 2400 python tools/mat_suite.py run --suite mm --out gpurun_out/r02_mat_suite_mm.json > gpurun_out/s5_mat_mm.log 2>&1
