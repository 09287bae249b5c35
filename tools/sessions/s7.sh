#!/bin/bash
# GPU session 7: full test suite; ncu full capture of the partitioned multistage kernels; sparse left- vs right-looking big fronts
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s7_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/s7_pytest_all.log
for v in left right; do
  B200_MF_PROF=1 B200_MF_BIG=$v timeout 300 python bench.py --workload sparse --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s7_bench_sparse_$v.json 2> gpurun_out/s7_bench_sparse_$v.err
done
B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"msp_fwd_kernel|msp_spike_kernel|msp_bwd_kernel|msw_factor_chain_kernel" --launch-skip 40 --launch-count 6 -o gpurun_out/r02_ncu_ms_partition \
   python bench.py --workload multistage --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s7_ncu_ms.log 2>&1
ncu -i gpurun_out/r02_ncu_ms_partition.ncu-rep --page raw --csv > gpurun_out/r02_ncu_ms_partition_raw.csv 2>/dev/null
