#!/bin/bash
# GPU session 6: warp-level DMMA spike kernel + run-local meta; level-kernel probes of the three suite problems the oracle solves
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python -m pytest tests/test_gpu_multistage.py tests/test_gpu_full_size.py::test_config4_full_size_batch_matches_oracle -m gpu -x -q > gpurun_out/s6_pytest_ms.log 2>&1
echo "rc=$?" >> gpurun_out/s6_pytest_ms.log
for k in 4 5; do
  B200_MS_TIMING=1 B200_MS_SEGMENTS=$k timeout 300 python bench.py --workload multistage --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/s6_ms_timing_K$k.json 2> gpurun_out/s6_ms_timing_K$k.err
done
for k in default 4 6; do
  if [ $k = default ]; then unset B200_MS_SEGMENTS; else export B200_MS_SEGMENTS=$k; fi
  timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6_bench_ms_K$k.json 2> gpurun_out/s6_bench_ms_K$k.err
done
unset B200_MS_SEGMENTS
B200_MS_NO_PARTITION=1 timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6_bench_ms_nopart.json 2> gpurun_out/s6_bench_ms_nopart.err
B200_LDLT_LEVELS=1 timeout 600 python tools/mat_suite.py run --suite mm --only QSIERRA > gpurun_out/s6_probe_levels.log 2>&1
B200_LDLT_LEVELS=1 timeout 600 python tools/mat_suite.py run --suite netlib_feas --only share2b,sierra,finnis >> gpurun_out/s6_probe_levels.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s6_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/s6_pytest_all.log
