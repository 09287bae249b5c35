#!/bin/bash
# session 22: dense e2e after removing the device-wide synchronisations from setup / cleanup; sub-batch count
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for c in 4 6 8; do
  B200_E2E_CHUNKS=$c B200_TIMING=1 timeout 300 python bench.py --workload dense --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s22_bench_dense_c$c.json 2> gpurun_out/s22_bench_dense_c$c.err
done
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_capi_symbols.py tests/test_gpu_sparse_ldlt.py -m gpu -x -q > gpurun_out/s22_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s22_pytest.log
B200_SUITE_THREADS=12 timeout 600 python tools/mm_suite.py > gpurun_out/s22_mm_suite_t12.json 2> gpurun_out/s22_mm_suite_t12.err
tail -3 gpurun_out/s22_pytest.log
