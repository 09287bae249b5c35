#!/bin/bash
# session 28: look-ahead of the diagonal-tile update in the dense Cholesky
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > gpurun_out/s28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s28_pytest.log
timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s28_bench_dense.json 2> gpurun_out/s28_bench_dense.err
B200_CHOL_LOOKAHEAD=0 timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s28_bench_dense_nola.json 2> gpurun_out/s28_bench_dense_nola.err
B200_TIMELINE=1 timeout 300 python tools/timeline.py --workload dense --out gpurun_out/s28_timeline_dense.raw > gpurun_out/s28_timeline_dense.txt 2>&1
tail -3 gpurun_out/s28_pytest.log
