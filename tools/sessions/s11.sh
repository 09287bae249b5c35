#!/bin/bash
# GPU session 11: Cholesky variants: fork/join diag, 64-row solve kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_dense_ldlt.py tests/test_gpu_sparse_ldlt.py -m gpu -q > gpurun_out/s11_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/s11_pytest.log
for v in default aux0 solve128 aux0_solve128 split0; do
  export B200_CHOL_SPLIT=1 B200_CHOL_AUX=1 B200_CHOL_SOLVE64=1
  case $v in aux0) export B200_CHOL_AUX=0;; solve128) export B200_CHOL_SOLVE64=0;; aux0_solve128) export B200_CHOL_AUX=0 B200_CHOL_SOLVE64=0;; split0) export B200_CHOL_SPLIT=0;; esac
  timeout 300 python bench.py --workload dense --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s11_bench_dense_$v.json 2> gpurun_out/s11_bench_dense_$v.err
done
