#!/bin/bash
# GPU session 10: split Cholesky (t64 update + solve-only panel) vs fused
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_dense_ldlt.py -m gpu -q > gpurun_out/s10_pytest_dense.log 2>&1
echo "rc=$?" >> gpurun_out/s10_pytest_dense.log
for v in 1 0; do
  B200_CHOL_SPLIT=$v timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s10_bench_dense_split$v.json 2> gpurun_out/s10_bench_dense_split$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"chol|gemm_nt_t64" --launch-skip 60 -c 60 --csv --log-file gpurun_out/s10_launches_chol.csv python bench.py --workload dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s10_ncu.log 2>&1
