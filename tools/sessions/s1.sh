#!/bin/bash
# GPU session 1 (round 2): tests, iteration-parity tables, mf_factor phase clocks, sanitizer logs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_config3_full_size_factor_solve_matches_oracle_golden > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
timeout 600 python tools/mm_iter_parity.py --set small --out gpurun_out/r02_mm_iter_parity_small.json > gpurun_out/s1_parity_small.log 2>&1
timeout 600 python tools/mm_iter_parity.py --set small --names DUALC1,HS118,LOTSCHD,QAFIRO,QPCBLEND,QSC205 --modes sparse_ldlt_eq_cond,sparse_ldlt_ineq_cond,sparse_ldlt_cond --out gpurun_out/r02_mm_iter_parity_cond.json > gpurun_out/s1_parity_cond.log 2>&1
timeout 900 python tools/mm_iter_parity.py --set mid --out gpurun_out/r02_mm_iter_parity_mid.json > gpurun_out/s1_parity_mid.log 2>&1
B200_MF_PROF=1 timeout 300 python bench.py --workload sparse --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s1_bench_sparse_prof.json 2> gpurun_out/s1_bench_sparse_prof.err
timeout 300 python bench.py --workload multistage --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_ms.json 2> gpurun_out/s1_bench_ms.err
for t in "tests/test_gpu_dense.py -k known_answers" "tests/test_gpu_multistage.py -k mpc_matches" "tests/test_gpu_sparse_ldlt.py -k batched"; do
  n=$(echo $t | sed 's/[^a-z_]/_/g' | cut -c1-40)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $t -m gpu -x -q > gpurun_out/r02_sanitizer_memcheck_$n.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_sanitizer_memcheck_$n.log
done
