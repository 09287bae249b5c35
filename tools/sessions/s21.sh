#!/bin/bash
# session 21: static-shape spike stage, SpMV loads in groups of four; full GPU test pass
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s21_pytest.log
timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s21_bench_ms.json 2> gpurun_out/s21_bench_ms.err
B200_TIMELINE=1 timeout 300 python tools/timeline.py --workload multistage --out gpurun_out/s21_timeline_ms.raw > gpurun_out/s21_timeline_ms.txt 2>&1
timeout 300 python bench.py --workload sparse --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s21_bench_sparse.json 2> gpurun_out/s21_bench_sparse.err
tail -3 gpurun_out/s21_pytest.log
