#!/bin/bash
# GPU session 12: round-2 validation set: full tests, default bench, launch lists + ncu full captures for profiles/
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_pytest_gpu.log
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02b_bench_all.json 2> gpurun_out/r02b_bench_all.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1
# launch lists (cold, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_launches_dense_c2.csv python bench.py --workload dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_dense.log 2>&1
B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"ms|k_|spmv" --launch-skip 200 -c 400 --csv --log-file gpurun_out/r02b_launches_multistage_c4.csv python bench.py --workload multistage --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_ms.log 2>&1
# full captures: dense assembly + Cholesky kernels, sparse factor kernel
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"gemm_nt_t64_kernel|chol_diag_kernel|chol_solve64_kernel" --launch-skip 30 --launch-count 8 -o gpurun_out/r02b_ncu_dense python bench.py --workload dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_dense_full.log 2>&1
ncu -i gpurun_out/r02b_ncu_dense.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_dense_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"mf_factor_kernel|mf_solve_ring_kernel" --launch-skip 4 --launch-count 2 -o gpurun_out/r02b_ncu_sparse python bench.py --workload sparse --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02b_ncu_sparse_full.log 2>&1
ncu -i gpurun_out/r02b_ncu_sparse.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_sparse_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_ncu_dense.ncu-rep gpurun_out/r02b_ncu_sparse.ncu-rep
