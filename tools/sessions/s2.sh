#!/bin/bash
# GPU session 2: CUDA-graph iteration + parallel-in-horizon multistage
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 python -m pytest tests/test_gpu_multistage.py -m gpu -x -q > gpurun_out/s2_pytest_ms.log 2>&1
echo "rc=$?" >> gpurun_out/s2_pytest_ms.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/s2_pytest_all.log
for v in default nograph nopart nograph_nopart; do
  export B200_NO_GRAPH=0 B200_MS_NO_PARTITION=0
  case $v in nograph) export B200_NO_GRAPH=1;; nopart) export B200_MS_NO_PARTITION=1;; nograph_nopart) export B200_NO_GRAPH=1 B200_MS_NO_PARTITION=1;; esac
  timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_bench_ms_$v.json 2> gpurun_out/s2_bench_ms_$v.err
done
for k in 4 6 8 11 14 18; do
  B200_NO_GRAPH=0 B200_MS_NO_PARTITION=0 B200_MS_SEGMENTS=$k timeout 300 python bench.py --workload multistage --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_bench_ms_K$k.json 2> gpurun_out/s2_bench_ms_K$k.err
done
unset B200_MS_SEGMENTS
for v in 0 1; do
  B200_NO_GRAPH=$v timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_bench_dense_nograph$v.json 2> gpurun_out/s2_bench_dense_nograph$v.err
  B200_NO_GRAPH=$v timeout 300 python bench.py --workload sparse --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_bench_sparse_nograph$v.json 2> gpurun_out/s2_bench_sparse_nograph$v.err
done
