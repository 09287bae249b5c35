#!/bin/bash
# GPU session 15: validation of the final round-2 state: full tests, smoke, default bench, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r02c_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02c_bench_all.json 2> gpurun_out/r02c_bench_all.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"gemm_nt_t64_kernel|chol_diag_kernel|chol_solve64_kernel" --launch-skip 30 --launch-count 8 -o gpurun_out/r02c_ncu_dense python bench.py --workload dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02c_ncu_dense_full.log 2>&1
ncu -i gpurun_out/r02c_ncu_dense.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_dense_raw.csv 2>/dev/null
rm -f gpurun_out/r02c_ncu_dense.ncu-rep
