#!/bin/bash
# GPU session 14: blocked-storage diag kernel (two CTAs per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_dense_ldlt.py tests/test_gpu_sparse_ldlt.py -m gpu -q > gpurun_out/s14_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/s14_pytest.log
for v in default aux0 split0; do
  export B200_CHOL_SPLIT=1 B200_CHOL_AUX=1
  case $v in aux0) export B200_CHOL_AUX=0;; split0) export B200_CHOL_SPLIT=0;; esac
  timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s14_bench_dense_$v.json 2> gpurun_out/s14_bench_dense_$v.err
done
export B200_CHOL_SPLIT=1 B200_CHOL_AUX=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"chol|gemm_nt_t64" --launch-skip 60 -c 48 --csv --log-file gpurun_out/s14_launches_chol.csv python bench.py --workload dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s14_ncu.log 2>&1
